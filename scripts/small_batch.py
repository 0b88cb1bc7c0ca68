"""A 10k x 1024 collection searched with 16-query batches (the text2sql / memory-store regime): used under
`ncu --metrics gpu__time_duration.sum` for the per-kernel breakdown of a small-collection batch."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from youtu_rag_b200 import native  # noqa: E402

n, d, nq, k = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000, 1024, int(sys.argv[2]) if len(sys.argv) > 2 else 16, 10
rng = np.random.default_rng(0)
ix = native.Index(d, "cosine", "bf16", 0, n)
ix.append(rng.standard_normal((n, d)).astype(np.float32))
q = rng.standard_normal((nq, d)).astype(np.float32)
for _ in range(6):
    ix.search(q, k)
print("ok", ix.launches())
