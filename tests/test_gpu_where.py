"""K4 (device where-evaluator) against the row-at-a-time oracle.  Needs a B200."""

import json

import numpy as np
import pytest

from oracle import exact_search as ox
from oracle import where_eval as ow
from tests.golden_util import GOLDEN
from tests.helpers import unit_rows
from youtu_rag_b200 import native
from youtu_rag_b200.metadata import MetadataTable
from youtu_rag_b200.where import compile_where, normalize_filters

pytestmark = pytest.mark.gpu


def _setup(metas, d=16):
    ix = native.Index(d, "cosine", "bf16", 0, 0)
    ix.append(unit_rows(len(metas), d, 3))
    t = MetadataTable()
    t.append(metas)
    return ix, t


def _mask(ix, t, where):
    prog, cols = compile_where(where, t)
    t.sync(ix, cols)
    words, n = ix.where_mask(prog)
    m = ox.unpack_mask(words, ix.rows)
    assert int(m.sum()) == n
    return m


def test_where_on_golden_corpus():
    metas = GOLDEN["corpus"]["metadatas"]
    ix, t = _setup(metas)
    for flt in sorted({json.dumps(r["filters"], sort_keys=True) for r in GOLDEN["chroma"] if "error" not in r}):
        where = normalize_filters(json.loads(flt))
        assert np.array_equal(_mask(ix, t, where), ow.eval_where(where, metas)), flt


def test_where_random_trees():
    rng = np.random.default_rng(0)
    n = 5003
    metas = []
    for i in range(n):
        m = {}
        if rng.random() < 0.8: m["i"] = int(rng.integers(-5, 5))
        if rng.random() < 0.7: m["f"] = float(rng.integers(-4, 4)) / 2
        if rng.random() < 0.9: m["s"] = str(rng.choice(["a", "b", "c", "dd"]))
        if rng.random() < 0.3: m["b"] = bool(rng.integers(0, 2))
        if rng.random() < 0.5: m["mix"] = [int(rng.integers(0, 3)), float(rng.integers(0, 3)), "x"][int(rng.integers(0, 3))]
        m["stamp_min"] = 1_700_000_000 + int(rng.integers(0, 10**6)); m["stamp_max"] = m["stamp_min"] + int(rng.integers(0, 10**5))
        metas.append(m)
    ix, t = _setup(metas)

    def leaf():
        f = str(rng.choice(["i", "f", "s", "b", "mix", "nope"]))
        kind = rng.integers(0, 4)
        val = {"i": int(rng.integers(-5, 5)), "f": float(rng.integers(-4, 4)) / 2, "s": str(rng.choice(["a", "b", "zz"])),
               "b": bool(rng.integers(0, 2)), "mix": [1, 1.0, "x"][int(rng.integers(0, 3))], "nope": 1}[f]
        if kind == 0: return {f: val}
        if kind == 1: return {f: {str(rng.choice(["$eq", "$ne"])): val}}
        if kind == 2 and isinstance(val, (int, float)) and not isinstance(val, bool):
            return {f: {str(rng.choice(["$gt", "$gte", "$lt", "$lte"])): val}}
        return {f: {str(rng.choice(["$in", "$nin"])): [val, val] if isinstance(val, bool) else [val, type(val)(val * 2 if not isinstance(val, str) else val + "q")]}}

    def tree(depth):
        if depth == 0 or rng.random() < 0.3: return leaf()
        return {str(rng.choice(["$and", "$or"])): [tree(depth - 1) for _ in range(int(rng.integers(2, 4)))]}

    for _ in range(150):
        w = tree(3)
        assert np.array_equal(_mask(ix, t, w), ow.eval_where(w, metas)), w
    # time-range overlap filter as meta_retrieval_toolkit.py:237-255 builds it
    w = {"$or": [{"$and": [{"stamp_min": {"$lte": 1_700_300_000}}, {"stamp_max": {"$gte": 1_700_200_000}}]},
                 {"$and": [{"stamp_min": {"$lte": 1_700_900_000}}, {"stamp_max": {"$gte": 1_700_850_000}}]}]}
    assert np.array_equal(_mask(ix, t, w), ow.eval_where(w, metas))
    # None → everything live; tombstones are excluded
    ix.set_live([0, 5, 4000], False)
    m = _mask(ix, t, None)
    assert m.sum() == n - 3 and not m[[0, 5, 4000]].any()


def test_where_then_search_is_prefilter():
    n, d = 4000, 64
    x = unit_rows(n, d, 4)
    metas = [{"source": f"f{i % 50}", "year": 2000 + i % 25} for i in range(n)]
    ix = native.Index(d, "cosine", "bf16", 0, 0)
    ix.append(x)
    t = MetadataTable(); t.append(metas)
    where = {"$and": [{"source": {"$in": ["f3", "f4"]}}, {"year": {"$gte": 2012}}]}
    prog, cols = compile_where(where, t); t.sync(ix, cols)
    q = unit_rows(1, d, 5)[0]
    ids, scores, counts = ix.search(q, 10, where=prog)
    mask = ow.eval_where(where, metas)
    from tests.helpers import check_topk
    check_topk(ids[0, :counts[0]], scores[0, :counts[0]], ix.read_rows(np.arange(n)), ox.prepare(q, "cosine", "bf16")[0], 10,
               "cosine", "bf16", mask=mask)


def test_where_from_reference_filter_producers():
    """K4 on the `where` dicts the reference's toolkits emit (tests/golden/filter_producers.json),
    and a filtered search with each of them against the oracle."""
    from pathlib import Path

    from tests.helpers import check_topk

    g = json.loads((Path(__file__).parent / "golden" / "filter_producers.json").read_text())
    metas = g["metadatas"]
    ix, t = _setup(metas, d=32)
    rows = ix.read_rows(np.arange(ix.rows))
    q = unit_rows(1, 32, 5)[0]
    qp = ox.prepare(q, "cosine", "bf16")[0]
    for case in g["cases"]:
        if case["where"] is None:
            continue
        want = np.zeros(len(metas), bool)
        want[case["rows"]] = True
        assert np.array_equal(_mask(ix, t, case["where"]), want), case["producer"]
        prog, cols = compile_where(case["where"], t)
        t.sync(ix, cols)
        ids, scores, counts = ix.search(q, 5, where=prog)
        c = int(counts[0])
        assert c == min(5, int(want.sum()))
        if c:
            check_topk(ids[0, :c], scores[0, :c], rows, qp, 5, "cosine", "bf16", mask=want)


def test_repeated_filters_hit_the_caches_and_mutations_invalidate_them():
    """the same `where` over unchanged rows reuses its bitmask (K4) and, for a batch under a
    selective filter, the gathered rows (K8) — and any append / delete / metadata write drops both."""
    import asyncio

    from tests.helpers import unit_rows
    from youtu_rag_b200 import B200VectorStore, Chunk, VectorStoreConfig

    n, d = 70_000, 64                                   # >= 65536 rows: K8 territory
    x = unit_rows(n, d, 4)
    s = B200VectorStore(VectorStoreConfig(collection_name="col_cache", index_params={"storage_dtype": "bf16"}))
    asyncio.run(s.add_chunks([Chunk(id=f"c{i}", document_id=f"d{i % 10}", content="", chunk_index=i, metadata={"kb": i % 10}, embedding=x[i].tolist())
                              for i in range(n)]))
    q = unit_rows(4, d, 5)
    f = {"kb": 3}

    def ids(res):
        return [[c.id for c, _ in r] for r in res]

    first = asyncio.run(s.search_batch(q, top_k=5, filters=f))
    h0 = s.index.cache_stats()
    again = asyncio.run(s.search_batch(q, top_k=5, filters=f))
    h1 = s.index.cache_stats()
    assert ids(again) == ids(first) and h1[0] == h0[0] + 1 and h1[1] == h0[1] + 1          # both caches hit
    single = asyncio.run(s.search(q[0].tolist(), 5, f))
    assert [c.id for c, _ in single] == ids(first)[0] and s.index.cache_stats()[0] == h1[0] + 1
    victim = ids(first)[0][0]
    asyncio.run(s.delete([victim]))                                                          # tombstone → new epoch
    after = asyncio.run(s.search_batch(q, top_k=5, filters=f))
    assert s.index.cache_stats() == (h1[0] + 1, h1[1]) and victim not in ids(after)[0] and ids(after)[0][:4] == ids(first)[0][1:]
    asyncio.run(s.add_chunks([Chunk(id="new", document_id="d3", content="", chunk_index=0, metadata={"kb": 3}, embedding=q[1].tolist())]))
    assert ids(asyncio.run(s.search_batch(q, top_k=5, filters=f)))[1][0] == "new"            # the append is seen
    s.close()
