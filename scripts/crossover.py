"""K1-loop vs K2 for batches over small and large collections: device time per search (all launches of one
`yrb_index_search_device` call, CUDA events on the launching stream) for N x nq.  Feeds the auto-dispatch rule in
capi.cu (scan_select) and the table in profiles/.  Usage: python scripts/crossover.py [N ...]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from youtu_rag_b200 import native  # noqa: E402

d, k = 1024, 10
sizes = [int(a) for a in sys.argv[1:]] or [1_000, 10_000, 100_000, 1_000_000]
st = torch.cuda.Stream()
out = torch.zeros(256 * k, dtype=torch.int64, device="cuda")
g = torch.Generator(device="cuda")
g.manual_seed(0)
print("      N   nq   K1-loop us     K2 us    auto us")
for n in sizes:
    ix = native.Index(d, "cosine", "bf16", 0, n)
    for a in range(0, n, 125_000):
        m = min(125_000, n - a)
        blk = torch.randn(m, d, device="cuda", generator=g)
        torch.cuda.synchronize()
        ix.append_device(blk.data_ptr(), m)
    for nq in (1, 2, 4, 8, 16, 32, 64, 128, 256):
        q = torch.randn(nq, d, device="cuda")
        res = []
        for path in (native.PATH_K1, native.PATH_K2, native.PATH_AUTO):
            ix.set_path(path)
            with torch.cuda.stream(st):
                for _ in range(5):
                    ix.search_device(q.data_ptr(), nq, k, 0, out.data_ptr(), st.cuda_stream)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                reps = 3 if (path == native.PATH_K1 and nq * n > 16_000_000) else 20
                e0.record(st)
                for _ in range(reps):
                    ix.search_device(q.data_ptr(), nq, k, 0, out.data_ptr(), st.cuda_stream)
                e1.record(st)
                e1.synchronize()
            res.append(1e3 * e0.elapsed_time(e1) / reps)
        print(f"{n:8d}  {nq:3d}   {res[0]:10.1f}  {res[1]:8.1f}  {res[2]:9.1f}", flush=True)
    ix.close()
