"""K1-loop vs K2 for small batches: device time per search for nq in {1,2,4,8,...} on a 1M x 1024 bf16 corpus."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from youtu_rag_b200 import native  # noqa: E402

n, d, k = 1_000_000, 1024, 10
ix = native.Index(d, "cosine", "bf16", 0, n)
g = torch.Generator(device="cuda"); g.manual_seed(0)
for a in range(0, n, 125_000):
    blk = torch.randn(125_000, d, device="cuda", generator=g)
    torch.cuda.synchronize()
    ix.append_device(blk.data_ptr(), 125_000)
st = torch.cuda.Stream()
out = torch.zeros(256 * k, dtype=torch.int64, device="cuda")
print("nq   K1-loop ms   K2 ms")
for nq in (1, 2, 3, 4, 8, 16, 64, 128, 256):
    q = torch.randn(nq, d, device="cuda")
    res = []
    for path in (native.PATH_K1, native.PATH_K2):
        ix.set_path(path)
        with torch.cuda.stream(st):
            for _ in range(5):
                ix.search_device(q.data_ptr(), nq, k, 0, out.data_ptr(), st.cuda_stream)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            reps = 20 if (path == native.PATH_K2 or nq <= 16) else 3
            for _ in range(reps):
                ix.search_device(q.data_ptr(), nq, k, 0, out.data_ptr(), st.cuda_stream)
            e1.record(st)
            e1.synchronize()
        res.append(e0.elapsed_time(e1) / reps)
    print(f"{nq:3d}   {res[0]:9.3f}   {res[1]:7.3f}")
