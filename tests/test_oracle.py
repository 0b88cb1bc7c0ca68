"""The oracle against independent statements of the same maths (CPU only)."""

import numpy as np
import pytest
import torch
from sklearn.neighbors import NearestNeighbors

from oracle import exact_search as ox
from tests.helpers import unit_rows


def test_bf16_round_matches_torch():
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.standard_normal(100000).astype(np.float32) * 10.0 ** rng.integers(-8, 8, 100000),
                        np.array([0.0, -0.0, 1.0, 3.3895314e38, 1e-40, np.inf, -np.inf], np.float32)])
    want = torch.from_numpy(x).to(torch.bfloat16).to(torch.float32).numpy()
    got = ox.bf16_round(x)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert np.array_equal(ox.bf16_bits_to_f32(ox.bf16_bits(x)).view(np.uint32), want.view(np.uint32))


def test_prepare_cosine_unit_norm_and_zero_row():
    x = np.random.default_rng(1).standard_normal((50, 33)).astype(np.float32) * 7
    x[7] = 0
    y = ox.prepare(x, "cosine", "f32")
    n = np.linalg.norm(y.astype(np.float64), axis=1)
    assert np.allclose(np.delete(n, 7), 1.0, atol=1e-6) and n[7] == 0
    assert np.array_equal(ox.prepare(x, "dot", "f32"), x)
    assert np.array_equal(ox.prepare(x, "nonsense", "f32"), y)  # unknown metric → cosine (chroma_store.py:52)


@pytest.mark.parametrize("metric,sk", [("cosine", "cosine"), ("euclidean", "sqeuclidean")])
def test_exact_topk_vs_sklearn_brute(metric, sk):
    x = unit_rows(3000, 48, 2) * (1.0 if metric == "cosine" else 1.7)
    q = unit_rows(5, 48, 3)
    rows = ox.prepare(x, metric, "f32")
    nn = NearestNeighbors(n_neighbors=10, algorithm="brute", metric=sk).fit(rows.astype(np.float64))
    for j in range(q.shape[0]):
        qp = ox.prepare(q[j], metric, "f32")[0]
        ids, scores = ox.exact_topk(rows, qp, 10, metric)
        dist, ind = nn.kneighbors(qp[None].astype(np.float64))
        assert np.array_equal(ids, ind[0])
        np.testing.assert_allclose(scores, 1.0 - dist[0], rtol=0, atol=3e-7)   # score = 1 - distance (sklearn renormalises)


def test_exact_topk_ties_by_id_mask_and_short():
    x = unit_rows(200, 16, 4)
    x[150] = x[20]; x[60] = x[20]
    rows = ox.prepare(x, "cosine", "bf16")
    q = ox.prepare(x[20], "cosine", "bf16")[0]
    ids, scores = ox.exact_topk(rows, q, 5, "cosine")
    assert ids[:3].tolist() == [20, 60, 150] and scores[0] == scores[1] == scores[2]
    mask = np.zeros(200, bool); mask[[60, 150, 7]] = True
    ids, _ = ox.exact_topk(rows, q, 5, "cosine", mask)
    assert ids.tolist()[:2] == [60, 150] and set(ids.tolist()) == {60, 150, 7}   # fewer than k
    ids, _ = ox.exact_topk(rows, q, 5, "cosine", np.zeros(200, bool))
    assert ids.shape[0] == 0
    # blocked evaluation is independent of the block size
    a = ox.exact_topk(rows, q, 17, "dot", block=13)
    b = ox.exact_topk(rows, q, 17, "dot", block=1 << 20)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_mask_pack_roundtrip_lsb_first():
    m = np.random.default_rng(5).random(1000) < 0.1
    w = ox.pack_mask(m)
    assert w.dtype == np.uint32 and w.shape[0] == 32
    assert np.array_equal(ox.unpack_mask(w, 1000), m)
    one = np.zeros(70, bool); one[33] = True
    assert ox.pack_mask(one).tolist() == [0, 2, 0]


def test_score_key_monotone():
    s = np.sort(np.random.default_rng(6).standard_normal(10000).astype(np.float32))
    k = ox.score_key_u32(s)
    assert (np.diff(k.astype(np.int64)) >= 0).all()
    assert ox.score_key_u32(np.float32(-0.0)) < ox.score_key_u32(np.float32(0.0))


def test_faiss_flat_restatement_matches_exact():
    x = ox.l2_normalize(unit_rows(2000, 32, 8) * 3)
    q = unit_rows(1, 32, 9)[0] * 5
    ids, sim = ox.faiss_flat_search(x, q, 10, "cosine")
    ref_ids, ref_s = ox.exact_topk(x, ox.prepare(q, "cosine", "f32")[0], 10, "cosine")
    assert np.array_equal(ids, ref_ids)
    np.testing.assert_allclose(sim, ref_s, atol=1e-6)
    ids, sim = ox.faiss_flat_search(x, q, 10, "euclidean")
    ref_ids, ref_s = ox.exact_topk(x, q, 10, "euclidean")
    assert np.array_equal(ids, ref_ids)
    np.testing.assert_allclose(sim, 1.0 / (1.0 + (1.0 - ref_s)), rtol=1e-5)   # faiss_store.py:181-182


def test_c_restatement_agrees_with_numpy_oracle():
    """Third independent statement of the path (plain C, oracle/exact_search.c) vs the numpy oracle."""
    from oracle import c_oracle

    rng = np.random.default_rng(12)
    x = unit_rows(4000, 96, 13) * rng.uniform(0.5, 2.0, (4000, 1)).astype(np.float32)
    x[[10, 3000]] = x[10]
    x[77] = 0
    assert np.array_equal(c_oracle.normalize(x).view(np.uint32), ox.l2_normalize(x).view(np.uint32))
    sample = rng.standard_normal(2000).astype(np.float32) * 10.0 ** rng.integers(-6, 6, 2000)
    assert np.array_equal(c_oracle.bf16_bits(sample), ox.bf16_bits(sample))
    mask = rng.random(4000) < 0.3
    mask[[10, 3000]] = True
    for metric in ("cosine", "dot", "euclidean"):
        for dtype in ("bf16", "f32"):
            rows = ox.prepare(x, metric, dtype)
            stored = ox.bf16_bits(rows) if dtype == "bf16" else rows
            for q in (x[10], unit_rows(1, 96, 14)[0] * 1.7):
                qp = ox.prepare(q, metric, dtype)[0]
                for m in (None, mask):
                    want_ids, want_s = ox.exact_topk(rows, qp, 25, metric, m)
                    got_ids, got_s = c_oracle.topk(stored, qp, 25, metric, None if m is None else ox.pack_mask(m))
                    assert np.array_equal(got_ids, want_ids), (metric, dtype)
                    np.testing.assert_allclose(got_s, want_s, rtol=1e-12, atol=1e-12)
    ids, _ = c_oracle.topk(ox.bf16_bits(ox.prepare(x[:3], "cosine", "bf16")), ox.prepare(x[0], "cosine", "bf16")[0], 10, "cosine")
    assert ids.shape[0] == 3
