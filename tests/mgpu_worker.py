"""Worker for tests/test_gpu_multi.py: one rank per GPU (torchrun), row-sharded search vs the oracle."""

import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

from oracle import exact_search as ox  # noqa: E402
from tests.helpers import check_topk, unit_rows  # noqa: E402
from youtu_rag_b200 import native  # noqa: E402
from youtu_rag_b200.sharded import ShardedSearcher, shard_bounds  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n, d = 50001, 256
    x = unit_rows(n, d, 1)
    x[[7, n - 1]] = x[7]                                  # equal scores on the first and the last shard
    bounds = shard_bounds(n, world)
    a, b = bounds[rank], bounds[rank + 1]
    ix = native.Index(d, "cosine", "bf16", local, b - a)
    ix.append(x[a:b])
    rows = ox.prepare(x, "cosine", "bf16")
    assert np.array_equal(ix.read_rows(np.arange(b - a)).view(np.uint32), rows[a:b].view(np.uint32))
    kinds = {}
    for want in ("p2p", "nccl"):
        _run(ShardedSearcher(ix, bounds, exchange=want), want, kinds, x, rows, n, d, a, b, world)
    dist.barrier()
    if rank == 0:
        print("MGPU_OK world", world, "exchanges", kinds)
    dist.destroy_process_group()


def _run(s, want, kinds, x, rows, n, d, a, b, world):
    kinds[want] = s.exchange_kind
    mask = np.random.default_rng(2).random(n) < 0.2
    mask[[7, n - 1]] = True
    local_words = torch.from_numpy(ox.pack_mask(np.concatenate([mask[a:b], np.zeros((-(b - a)) % 64, bool)])).view(np.int32)).cuda()
    for nq, k in ((1, 10), (3, 100), (40, 10), (256, 100)):
        qs = np.concatenate([x[7][None], unit_rows(nq, d, 3)])[:nq]
        for m, dm in ((None, None), (mask, local_words)):
            ids, scores, counts = s.search(qs, k, dm)
            for j in sorted({0, nq // 2, nq - 1}):
                c = int(counts[j])
                check_topk(ids[j, :c], scores[j, :c], rows, ox.prepare(qs[j], "cosine", "bf16")[0], k, "cosine", "bf16", mask=m)
            assert ids[0, :2].tolist() == [7, n - 1]
    s.synchronize()


if __name__ == "__main__":
    main()
