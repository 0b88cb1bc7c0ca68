"""The N>1 path on CPU: world_size-2 `gloo` processes exchange per-shard top-k keys and merge them.

Each rank computes its shard's local top-k with the ORACLE (no GPU here), packs them into the same
64-bit selection keys the kernels emit, all-gathers them and merges with the global
(score desc, id asc) rule; every rank must end up with the oracle's global top-k."""

import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import exact_search as ox
from tests.helpers import unit_rows
from youtu_rag_b200.sharded import exchange_host, merge_host, shard_bounds


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _local_keys(rows, q, k, mask):
    """oracle top-k of one shard → uint64 keys (key 0 = empty), as K1/K2 would emit them."""
    ids, scores = ox.exact_topk(rows, q, k, "cosine", mask)
    keys = np.zeros(k, np.uint64)
    s32 = scores.astype(np.float32)
    keys[: ids.shape[0]] = (ox.score_key_u32(s32).astype(np.uint64) << np.uint64(32)) | (~ids.astype(np.uint32)).astype(np.uint64)
    return keys


def _worker(rank, world, port, n, d, k, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        x = unit_rows(n, d, 1)
        x[[3, n - 2]] = x[3]                       # a tie that straddles the two shards
        rows = ox.prepare(x, "cosine", "bf16")
        qs = ox.prepare(np.concatenate([x[3][None], unit_rows(3, d, 2)]), "cosine", "bf16")
        mask = np.random.default_rng(3).random(n) < 0.3
        mask[[3, n - 2]] = True
        bounds = shard_bounds(n, world)
        a, b = bounds[rank], bounds[rank + 1]
        for m in (None, mask):
            local = np.stack([_local_keys(rows[a:b], q, k, None if m is None else m[a:b]) for q in qs])
            gathered = exchange_host(local)
            assert gathered.shape == (world, qs.shape[0], k)
            ids, scores, counts = merge_host(gathered, bounds, k)
            for j, q in enumerate(qs):
                want_ids, want_s = ox.exact_topk(rows, q, k, "cosine", m)
                assert counts[j] == want_ids.shape[0]
                assert np.array_equal(ids[j, : counts[j]], want_ids), (rank, j, ids[j], want_ids)
                np.testing.assert_allclose(scores[j, : counts[j]], want_s, rtol=1e-6, atol=1e-7)
            if m is None:
                assert ids[0, :2].tolist() == [3, n - 2]   # equal scores: lower GLOBAL id first
        out.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        out.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n,k", [(1001, 10), (64, 40)])
def test_two_rank_exchange_and_merge(n, k):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, 32, k, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = [out.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_shard_bounds():
    assert shard_bounds(10, 4) == [0, 3, 6, 8, 10]
    assert shard_bounds(8, 8) == list(range(9))
    assert shard_bounds(3, 4) == [0, 1, 2, 3, 3]
