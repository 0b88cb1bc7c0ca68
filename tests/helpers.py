"""Shared test helpers: tolerance-aware top-k comparison against the oracle."""

from __future__ import annotations

import numpy as np

from oracle import exact_search as ox

# north_star tolerances: relative 1e-3 for bf16 storage, 1e-5 for fp32 (scores of unit vectors, so
# an absolute floor of the same size covers scores near zero).
TOL = {"bf16": 1e-3, "f32": 1e-5}
# fp32-accumulation noise floor between GPU (fp32 FMA) and oracle (fp64) on identical operands;
# two scores closer than this are "tied" and either order is accepted.
TIE_EPS = {"bf16": 2e-6, "f32": 2e-6}


def unit_rows(n: int, d: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, d)).astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    return x


def check_topk(ids, scores, stored_rows, q_prepared, k, metric, dtype, mask=None, tie_eps=None):
    """ids/scores: one query's result (valid prefix only).  stored_rows/q_prepared: the operands as
    stored (oracle `prepare` output or a read-back).  Asserts id-exact agreement with the oracle,
    accepting a different order/choice only among candidates tied within `tie_eps`."""
    ids = np.asarray(ids, dtype=np.int64)
    scores = np.asarray(scores, dtype=np.float64)
    tie_eps = TIE_EPS[dtype] if tie_eps is None else tie_eps
    full = ox.scores_f64(stored_rows, q_prepared, metric)
    n = full.shape[0]
    valid = np.ones(n, bool) if mask is None else np.asarray(mask, bool)
    ref_ids, ref_scores = ox.exact_topk(stored_rows, q_prepared, k, metric, mask)
    assert ids.shape[0] == ref_ids.shape[0], f"returned {ids.shape[0]} results, oracle {ref_ids.shape[0]}"
    if ids.shape[0] == 0:
        return
    assert len(set(ids.tolist())) == ids.shape[0], "duplicate ids in result"
    assert ((ids >= 0) & (ids < n)).all() and valid[ids].all(), "result contains filtered / out-of-range rows"
    # scores agree with the oracle's score of the same row
    tol = TOL[dtype]
    np.testing.assert_allclose(scores, full[ids], rtol=tol, atol=tol)
    # ordering: non-increasing, ties by id, up to tie_eps
    for a in range(ids.shape[0] - 1):
        assert full[ids[a]] >= full[ids[a + 1]] - tie_eps, f"order violated at {a}"
    if np.array_equal(ids, ref_ids):
        return
    # any disagreement must be explained by near-ties around the positions that differ
    for a in np.flatnonzero(ids != ref_ids):
        assert abs(full[ids[a]] - ref_scores[a]) <= tie_eps, (
            f"position {a}: got row {ids[a]} (oracle score {full[ids[a]]:.9f}), oracle row {ref_ids[a]} "
            f"({ref_scores[a]:.9f}) — not a tie")


def recall_at_k(ids, ref_ids) -> float:
    ref = set(np.asarray(ref_ids).tolist())
    return len(ref & set(np.asarray(ids).tolist())) / max(1, len(ref))
