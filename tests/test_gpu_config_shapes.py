"""BASELINE.json's configurations at FULL size against the sliced oracle: C3 (1M x 1024, 256-query
batch, top-100 — the K2 pair kernel, fused sampling, K3), the north-star target (10M x 1024, single query, top-10 — K1)
and C4 on one GPU (10M x 1024 + 10 % metadata mask), plus one C5 shard's shape (768-d, 1024 queries → four 256-query
chunks).  The oracle is oracle/exact_search.py over the stored rows read back bit-exactly, shortlisted per 250k-row
block with fp32 BLAS (bench.py's `local_shortlists` / `parity_report`, the check every bench run carries).  Slow: ~1 min."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def bench_mod():
    import bench

    return bench, bench.Env(1)


def _search(corpus, k, queries, mask_bits=None):
    if mask_bits is None:
        return corpus.index.search(queries, k)
    import torch

    nq = queries.shape[0]
    out = [torch.empty(nq * k, dtype=torch.int64, device="cuda"), torch.empty(nq * k, dtype=torch.float32, device="cuda"),
           torch.empty(nq, dtype=torch.int32, device="cuda")]
    dq = torch.from_numpy(np.ascontiguousarray(queries)).cuda()
    corpus.index.search_device_ids(dq.data_ptr(), nq, k, mask_bits[0].data_ptr(), out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr())
    torch.cuda.synchronize()
    return out[0].view(-1, k).cpu().numpy(), out[1].view(-1, k).cpu().numpy(), out[2].cpu().numpy()


def _probe(corpus, name, k, queries, mask_bits=None, n_check=8):
    ids, scores, counts = _search(corpus, k, queries, mask_bits)
    assert (counts == k).all()
    return {"name": name, "k": k, "queries": queries[:n_check], "mask": None if mask_bits is None else mask_bits[1],
            "ids": ids[:n_check], "scores": scores[:n_check], "counts": counts[:n_check]}


def _check(bench, corpus, probes):
    """ONE pass over the read-back corpus serves every probe."""
    rep = bench.parity_report(probes, [bench.local_shortlists(corpus, probes)])
    for name, w in rep["workloads"].items():
        assert w["ok"] and w["id_mismatches"] == 0 and w["recall_at_k"] == 1.0, (name, w)


def test_c3_full_shape(bench_mod):
    bench, env = bench_mod
    rows, dim, nq, k, _ = bench.WORKLOADS["c3"]
    corpus = bench.Corpus(env, rows, dim)
    _check(bench, corpus, [_probe(corpus, "c3", k, bench.host_queries(dim, nq)[0])])
    corpus.index.close()


def test_c5_shard_shape_scaled(bench_mod):
    """C5's per-GPU shape at a fifth of its rows (2.5M x 768, 1024-query batch, top-10): four 256-query chunks, 768-d
    tiles (12 k-blocks), k = 10 thresholds from one published score per pair."""
    bench, env = bench_mod
    corpus = bench.Corpus(env, 2_500_000, 768)
    _check(bench, corpus, [_probe(corpus, "c5s", 10, bench.host_queries(768, 1024)[0])])
    corpus.index.close()


def test_north_star_and_c4_full_shape(bench_mod):
    bench, env = bench_mod
    rows, dim, _, k, sel = bench.WORKLOADS["c4"]
    corpus = bench.Corpus(env, rows, dim)                        # 20.5 GB resident
    q = bench.host_queries(dim, 1)[:4, 0]                        # four single queries
    m = corpus.mask(sel)
    probes = [_probe(corpus, f"t10m_{j}", k, q[j:j + 1], n_check=1) for j in range(4)]                  # K1 over 10M rows
    probes += [_probe(corpus, f"c4_{j}", k, q[j:j + 1], mask_bits=m, n_check=1) for j in range(2)]      # K1 + the 10 % bitmask
    probes.append(_probe(corpus, "c4b", k, bench.host_queries(dim, 256)[1], mask_bits=m, n_check=4))    # K8 compaction + K2
    _check(bench, corpus, probes)
    corpus.index.close()
