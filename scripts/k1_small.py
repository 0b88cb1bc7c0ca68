"""K1 on a small shard (default 125k x 1024 bf16 = one C2 shard of an 8-GPU box): device time per search, and with
YRB_K1_TRACE=1 the per-phase breakdown the library prints.  Used to size the launch's fixed costs."""
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from youtu_rag_b200 import native  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 125_000
sel = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0  # < 1: scan under a random row mask of this selectivity
d, k = 1024, 10
ix = native.Index(d, "cosine", "bf16", 0, n)
g = torch.Generator(device="cuda"); g.manual_seed(0)
for a in range(0, n, 125_000):
    m = min(125_000, n - a)
    blk = torch.randn(m, d, device="cuda", generator=g)
    torch.cuda.synchronize()
    ix.append_device(blk.data_ptr(), m)
st = torch.cuda.Stream()
out = torch.zeros(k, dtype=torch.int64, device="cuda")
q = torch.randn(1, d, device="cuda", generator=g)
mask_ptr = 0
if sel < 1.0:
    n_words = (n + 63) // 64 * 2
    bits = (torch.rand(n_words * 32, device="cuda", generator=g) < sel)
    bits[n:] = False
    w = (bits.view(n_words, 32).to(torch.int64) << torch.arange(32, device="cuda")).sum(1)
    mask = w.bitwise_and(0xFFFFFFFF).to(torch.uint32).contiguous()
    mask_ptr = mask.data_ptr()
torch.cuda.synchronize()
if os.environ.get("YRB_K1_TRACE"):
    for _ in range(3):
        ix.search_device(q.data_ptr(), 1, k, mask_ptr, out.data_ptr(), st.cuda_stream)
    sys.exit(0)
for reserved in (0, 2):
    ix.set_reserved_sms(reserved)
    with torch.cuda.stream(st):
        for _ in range(20):
            ix.search_device(q.data_ptr(), 1, k, mask_ptr, out.data_ptr(), st.cuda_stream)
        best = 1e9
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            for _ in range(200):
                ix.search_device(q.data_ptr(), 1, k, mask_ptr, out.data_ptr(), st.cuda_stream)
            e1.record(st)
            e1.synchronize()
            best = min(best, e0.elapsed_time(e1) / 200)
    ids = out.cpu().tolist()
    print(f"lib={os.environ.get('YRB200_LIB', 'tree')} rows={n} sel={sel} reserved={reserved}: {best * 1e3:.2f} us/search  "
          f"({sel * n * d * 2 / best / 1e6:.0f} GB/s)  keys={ids[:3]}")
