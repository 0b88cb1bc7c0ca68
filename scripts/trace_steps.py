"""Timeline of a few search steps (kernel start/duration per stream) via torch.profiler — a stand-in for nsys.

    torchrun --nproc-per-node 2 scripts/trace_steps.py [rows] [steps]
"""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from youtu_rag_b200 import native  # noqa: E402
from youtu_rag_b200.sharded import ShardedSearcher  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
n_local, d = rows // world, 1024
ix = native.Index(d, "cosine", "bf16", local, n_local)
g = torch.Generator(device="cuda"); g.manual_seed(rank)
for a in range(0, n_local, 125_000):
    blk = torch.randn(min(125_000, n_local - a), d, device="cuda", generator=g)
    torch.cuda.synchronize()
    ix.append_device(blk.data_ptr(), blk.shape[0])
bounds = [i * n_local for i in range(world + 1)]
s = ShardedSearcher(ix, bounds)
q = torch.randn(64, 1, d, device="cuda")
for i in range(20):
    s.search_device(q[i % 64], 10)
s.synchronize()
if world > 1:
    dist.barrier()
from torch.profiler import ProfilerActivity, profile  # noqa: E402

with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(steps):
        s.search_device(q[i % 64], 10)
    s.synchronize()
if rank == 0:
    ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    ev.sort(key=lambda e: e.time_range.start)
    t0 = ev[0].time_range.start
    print("start_us  dur_us  stream  kernel")
    for e in ev:
        print(f"{e.time_range.start - t0:9.1f} {e.time_range.end - e.time_range.start:7.1f}  {getattr(e, 'stream', '?')}  {e.name[:70]}")
    cpu = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CPU and ("nccl" in e.name.lower() or "all_gather" in e.name.lower())]
    for e in cpu[:6]:
        print("cpu", e.name[:60], round(e.time_range.end - e.time_range.start, 1), "us")
if world > 1:
    dist.destroy_process_group()
