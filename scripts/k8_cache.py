"""C4b through the `where` API (10M x 1024 bf16, a metadata column selecting 10 % of the rows, 256-query batches, top-10):
first search of a filter (K4 + K8 count + gather + K2) against the same filter repeated (mask and compaction cached:
K2 on the gathered rows only).  VERDICT r1 next-round item 7."""
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from youtu_rag_b200 import native  # noqa: E402

n, d, nq, k = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000, 1024, 256, 10
ix = native.Index(d, "cosine", "bf16", 0, n)
g = torch.Generator(device="cuda")
for a in range(0, n, 125_000):
    m = min(125_000, n - a)
    g.manual_seed(a)
    blk = torch.randn(m, d, device="cuda", generator=g)
    torch.cuda.synchronize()
    ix.append_device(blk.data_ptr(), m)
kb = (np.arange(n, dtype=np.int64) * 2654435761 % 10)            # a "knowledge base id" per row, 10 % each
ix.column_write(0, native.COL_I64, 0, kb, np.ones(n, np.uint8))
q = np.random.default_rng(1).standard_normal((nq, d)).astype(np.float32)


def where_eq(v):
    return native.CompiledWhere([(0, native.OPS["$eq"], 0, 1)], [v], [0])


def timed(w, reps):
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        out = ix.search(q, k, where=w)
        ts.append(time.perf_counter() - t0)
    return out, ts


w3, w4 = where_eq(3), where_eq(4)
ix.search(q, k, where=where_eq(9))                                 # warm-up: allocations, first compaction buffers
out_a, t_first = timed(w3, 1)
out_b, t_rep = timed(w3, 20)
h = ix.cache_stats()
out_c, t_other = timed(w4, 1)
assert np.array_equal(out_a[0], out_b[0]) and (kb[out_a[0].ravel()] == 3).all() and (kb[out_c[0].ravel()] == 4).all()
print(f"rows={n} nq={nq} k={k}: first search of a filter {1e3 * t_first[0]:.3f} ms, repeated (cached mask + compaction) "
      f"median {1e3 * np.median(t_rep):.3f} ms (min {1e3 * min(t_rep):.3f}), another filter {1e3 * t_other[0]:.3f} ms; "
      f"cache hits (filter, compaction) = {h}")
