"""fp32-storage batches: K1Q (four queries per pass) vs a K1 loop, device time per batch on a 1M x 1024 fp32 corpus."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from youtu_rag_b200 import native  # noqa: E402

n, d, k = 1_000_000, 1024, 10
ix = native.Index(d, "cosine", "f32", 0, n)
g = torch.Generator(device="cuda"); g.manual_seed(0)
for a in range(0, n, 125_000):
    blk = torch.randn(125_000, d, device="cuda", generator=g)
    torch.cuda.synchronize()
    ix.append_device(blk.data_ptr(), 125_000)
st = torch.cuda.Stream()
out = torch.zeros(64 * k, dtype=torch.int64, device="cuda")
print("nq   K1-loop ms   K1Q ms   (4.096 GB per pass: %.3f ms at 6459 GB/s)" % (n * d * 4 / 6459e9 * 1e3))
for nq in (1, 2, 4, 8, 16):
    q = torch.randn(nq, d, device="cuda", generator=g)
    res = []
    for path in (native.PATH_K1, native.PATH_AUTO):
        ix.set_path(path)
        with torch.cuda.stream(st):
            for _ in range(3):
                ix.search_device(q.data_ptr(), nq, k, 0, out.data_ptr(), st.cuda_stream)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            for _ in range(10):
                ix.search_device(q.data_ptr(), nq, k, 0, out.data_ptr(), st.cuda_stream)
            e1.record(st)
            e1.synchronize()
        res.append(e0.elapsed_time(e1) / 10)
    print(f"{nq:3d}   {res[0]:9.3f}   {res[1]:7.3f}")
