"""bench.py's in-run parity check (`local_shortlists` + `parity_report`, the sliced oracle) exercised on the CPU over a
fake index: it must accept the exact answer — also when the rows are split over several "ranks" — and reject a wrong id,
a wrong score, a missing hit and a filtered-out row."""

import numpy as np

import bench
from oracle import exact_search as ox
from tests.fake_index import FakeIndex
from tests.helpers import unit_rows


class _Slice:
    """What bench.Corpus exposes to the parity code: an index over this rank's rows and where they start."""

    def __init__(self, x, base):
        self.index = FakeIndex(x.shape[1], "cosine", "bf16")
        self.index.append(x)
        self.dim, self.n_local, self.base = x.shape[1], x.shape[0], base


def _truth(x, q, k, mask=None):
    rows = ox.prepare(x, "cosine", "bf16")
    out = [ox.exact_topk(rows, ox.prepare(qq, "cosine", "bf16")[0], k, "cosine", mask) for qq in q]
    ids = np.stack([o[0] for o in out])
    return ids, np.stack([o[1] for o in out]).astype(np.float32), np.full(len(out), k, np.int32)


def test_parity_report_accepts_the_truth_and_rejects_errors():
    n, d, k = 6000, 64, 10
    x, q = unit_rows(n, d, 1), unit_rows(3, d, 2)
    mask = np.random.default_rng(3).random(n) < 0.2
    ids, sc, cnt = _truth(x, q, k)
    mids, msc, mcnt = _truth(x, q, k, mask)
    cuts = [0, 2500, 2500 + 1700, n]                           # three "ranks"
    slices = [_Slice(x[a:b], a) for a, b in zip(cuts, cuts[1:])]

    def report(probe_plain, probe_masked):
        probes = [probe_plain, probe_masked]
        gathered = []
        for sl, (a, b) in zip(slices, zip(cuts, cuts[1:])):
            local = [dict(p, mask=None if p["mask"] is None else p["mask"][a:b]) for p in probes]
            gathered.append(bench.local_shortlists(sl, local))
        return bench.parity_report(probes, gathered)

    plain = {"name": "plain", "k": k, "queries": q, "mask": None, "ids": ids, "scores": sc, "counts": cnt}
    masked = {"name": "masked", "k": k, "queries": q, "mask": mask, "ids": mids, "scores": msc, "counts": mcnt}
    rep = report(plain, masked)
    assert rep["ok"] and all(w["recall_at_k"] == 1.0 and w["id_mismatches"] == 0 for w in rep["workloads"].values()), rep

    bad = ids.copy()
    bad[1, 4] = (bad[1, 4] + 1) % n                            # a wrong id
    assert not report(dict(plain, ids=bad), masked)["ok"]
    worse = sc.copy()
    worse[0, 0] *= 1.01                                         # a score off by 1 %
    assert not report(dict(plain, scores=worse), masked)["ok"]
    short = cnt.copy()
    short[2] = k - 1                                            # a missing hit
    assert not report(dict(plain, counts=short), masked)["ok"]
    assert not report(plain, dict(masked, ids=ids))["workloads"]["masked"]["ok"]   # unfiltered answer under a filter
