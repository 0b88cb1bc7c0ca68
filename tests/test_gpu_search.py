"""Parity of the CUDA search path (through the C ABI) with the oracle.  Needs a B200."""

import numpy as np
import pytest

from oracle import exact_search as ox
from tests.helpers import check_topk, unit_rows
from youtu_rag_b200 import native

pytestmark = pytest.mark.gpu


def build(x, metric="cosine", dtype="bf16", reserve=0):
    ix = native.Index(x.shape[1], metric, dtype, 0, reserve)
    ix.append(x)
    return ix


def stored(ix):
    return ix.read_rows(np.arange(ix.rows))


@pytest.mark.parametrize("dtype", ["bf16", "f32"])
@pytest.mark.parametrize("metric", ["cosine", "dot", "euclidean"])
def test_ingest_matches_oracle_bits(metric, dtype):
    """K5: normalise (fp64 norm) + round to storage dtype lands on the oracle's stored values."""
    x = unit_rows(3000, 200, 1) * np.random.default_rng(2).uniform(0.1, 9, (3000, 1)).astype(np.float32)
    x[11] = 0
    ix = build(x, metric, dtype)
    got, want = stored(ix), ox.prepare(x, metric, dtype)
    diff = got.view(np.uint32) != want.view(np.uint32)
    assert diff.mean() < 1e-6, f"{diff.sum()} stored elements differ from the oracle"


@pytest.mark.parametrize("dtype", ["bf16", "f32"])
@pytest.mark.parametrize("n,d", [(1, 8), (3, 64), (257, 100), (1000, 768), (4099, 1024), (2500, 2048), (700, 4100)])
def test_k1_small_shapes(n, d, dtype):
    x = unit_rows(n, d, n + d)
    ix = build(x, "cosine", dtype)
    ix.set_path(native.PATH_K1)
    rows = stored(ix)
    qs = unit_rows(3, d, 99)
    for k in (1, 5, 10, 32, 33, 100, 128):
        ids, scores, counts = ix.search(qs, k)
        for j in range(3):
            qp = ox.prepare(qs[j], "cosine", dtype)[0]
            c = int(counts[j])
            assert c == min(k, n) and (ids[j, c:] == -1).all()
            check_topk(ids[j, :c], scores[j, :c], rows, qp, k, "cosine", dtype)


@pytest.mark.parametrize("metric", ["dot", "euclidean"])
@pytest.mark.parametrize("dtype", ["bf16", "f32"])
def test_k1_metrics(metric, dtype):
    x = unit_rows(5000, 256, 5) * np.random.default_rng(6).uniform(0.5, 2, (5000, 1)).astype(np.float32)
    ix = build(x, metric, dtype)
    ix.set_path(native.PATH_K1)
    rows = stored(ix)
    q = unit_rows(2, 256, 7) * 1.3
    ids, scores, counts = ix.search(q, 10)
    for j in range(2):
        qp = ox.prepare(q[j], metric, dtype)[0]
        # euclidean uses ||q||^2 - 2 q.x + ||x||^2 in fp32: cancellation noise ~1e-6 of the norms
        check_topk(ids[j], scores[j], rows, qp, 10, metric, dtype, tie_eps=2e-5 if metric == "euclidean" else None)


def test_k1_ties_broken_by_row_id():
    x = unit_rows(6000, 128, 8)
    dup = [17, 900, 901, 4000, 5999]
    x[dup] = x[17]
    ix = build(x)
    ix.set_path(native.PATH_K1)
    ids, scores, _ = ix.search(x[17], 8)
    assert ids[0, :5].tolist() == dup and len(set(scores[0, :5].tolist())) == 1


@pytest.mark.parametrize("sel", [0.0, 0.0005, 0.1, 0.5, 1.0])
def test_k1_mask_prefilter(sel):
    n, d = 20011, 256
    x = unit_rows(n, d, 9)
    ix = build(x)
    ix.set_path(native.PATH_K1)
    rows = stored(ix)
    mask = np.random.default_rng(10).random(n) < sel
    q = unit_rows(1, d, 11)[0]
    qp = ox.prepare(q, "cosine", "bf16")[0]
    ids, scores, counts = ix.search(q, 10, mask=ox.pack_mask(mask))
    c = int(counts[0])
    assert c == min(10, int(mask.sum()))
    check_topk(ids[0, :c], scores[0, :c], rows, qp, 10, "cosine", "bf16", mask=mask)


def test_tombstones_and_growth():
    d = 64
    x = unit_rows(5000, d, 12)
    ix = native.Index(d, "cosine", "bf16", 0, 0)
    for a in range(0, 5000, 700):          # grows the device arrays several times
        ix.append(x[a:a + 700])
    assert ix.counts() == (5000, 5000)
    rows = stored(ix)
    assert np.array_equal(rows.view(np.uint32), ox.prepare(x, "cosine", "bf16").view(np.uint32))
    q = x[123]
    qp = ox.prepare(q, "cosine", "bf16")[0]
    ids, _, _ = ix.search(q, 5)
    assert ids[0, 0] == 123
    dead = ids[0, :3].tolist()
    ix.set_live(dead, False)
    assert ix.counts() == (5000, 4997)
    live = np.ones(5000, bool); live[dead] = False
    ids2, s2, c2 = ix.search(q, 5)
    check_topk(ids2[0], s2[0], rows, qp, 5, "cosine", "bf16", mask=live)
    # a filter mask is AND-ed with the tombstones
    m = np.zeros(5000, bool); m[dead] = True; m[[7, 8]] = True
    ids3, _, c3 = ix.search(q, 5, mask=ox.pack_mask(m))
    assert int(c3[0]) == 2 and set(ids3[0, :2].tolist()) == {7, 8}
    ix.set_live(dead, True)
    ids4, _, _ = ix.search(q, 5)
    assert np.array_equal(ids4, ids)
    ix.clear()
    assert ix.counts() == (0, 0)
    ids5, _, c5 = ix.search(q, 5)
    assert int(c5[0]) == 0 and (ids5 == -1).all()


def test_argument_errors():
    ix = build(unit_rows(10, 16, 1))
    with pytest.raises(native.NativeError):
        ix.search(unit_rows(1, 16, 2), 0)
    with pytest.raises(ValueError):
        ix.search(unit_rows(1, 17, 2), 1)
    with pytest.raises(native.NativeError):
        ix.read_rows([10])
    with pytest.raises(native.NativeError):
        ix.set_live([-1], False)


def test_search_is_idempotent_and_sorted_1m():
    """C2-sized property test (1M x 1024 bf16): results are sorted, stable across calls, and the
    oracle agrees on a 64k-row slice that contains every returned row."""
    import torch

    n, d = 1_000_000, 1024
    ix = native.Index(d, "cosine", "bf16", 0, n)
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    for b in range(0, n, 250_000):
        blk = torch.randn(250_000, d, device="cuda", generator=g)
        torch.cuda.synchronize()
        ix.append_device(blk.data_ptr(), 250_000)
    ix.set_path(native.PATH_K1)
    q = unit_rows(2, d, 1)
    a = ix.search(q, 10)
    b = ix.search(q, 10)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert (np.diff(a[1], axis=1) <= 0).all() and (a[2] == 10).all()
    for j in range(2):
        qp = ox.prepare(q[j], "cosine", "bf16")[0]
        # exact check against the oracle on the whole corpus, read back in slices
        best_ids, best_s = np.empty(0, np.int64), np.empty(0)
        for s0 in range(0, n, 125_000):
            rows = ix.read_rows(np.arange(s0, s0 + 125_000))
            i, s = ox.exact_topk(rows, qp, 10, "cosine")
            best_ids = np.concatenate([best_ids, i + s0]); best_s = np.concatenate([best_s, s])
        o = ox.order_desc_id_asc(best_s, best_ids)[:10]
        assert np.array_equal(a[0][j], best_ids[o]), (a[0][j], best_ids[o])
        np.testing.assert_allclose(a[1][j], best_s[o], rtol=1e-3)


@pytest.mark.parametrize("dtype", ["bf16", "f32"])
def test_k6_large_k_any_path(dtype):
    """k > 128 takes the key-vector + radix-select path; forcing it for small k must agree too."""
    n, d = 30011, 128
    x = unit_rows(n, d, 21)
    x[[100, 20000, 30010]] = x[100]
    ix = build(x, "cosine", dtype)
    rows = stored(ix)
    q = np.concatenate([x[100][None], unit_rows(1, d, 22)])
    mask = np.random.default_rng(23).random(n) < 0.3
    for k in (129, 500, 4096):
        for m in (None, mask):
            ids, scores, counts = ix.search(q, k, mask=None if m is None else ox.pack_mask(m))
            for j in range(2):
                c = int(counts[j])
                assert c == min(k, n if m is None else int(m.sum()))
                check_topk(ids[j, :c], scores[j, :c], rows, ox.prepare(q[j], "cosine", dtype)[0], k, "cosine", dtype, mask=m)
    ix.set_path(native.PATH_K6)
    a = ix.search(q, 10)
    ix.set_path(native.PATH_K1)
    b = ix.search(q, 10)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert a[0][0, :3].tolist() == [100, 20000, 30010]
    ix.set_path(native.PATH_AUTO)
    with pytest.raises(native.NativeError):
        ix.search(q, 5000)                       # above the supported n_results
    small = build(unit_rows(50, d, 24), "cosine", dtype)
    ids, _, counts = small.search(q, 4097)       # clamped to the collection size
    assert int(counts[0]) == 50 and (ids[0, 50:] == -1).all()


@pytest.mark.parametrize("dtype", ["bf16", "f32"])
@pytest.mark.parametrize("metric", ["cosine", "euclidean"])
def test_score_threshold_inside_the_scan_equals_filtering_the_result(metric, dtype):
    """f4 (base_retriever.py:71: keep a hit when score >= threshold): `min_score` is the initial bound of the running
    top-k lists / the epilogue's survivor test on every kernel family — K1 (k <= 32 and k = 100), K2 or K1Q (batches),
    K6 (k > 128), with a bitmask, and over a collection sharded inside the process — and returns exactly the hits a
    host-side filter of the plain result keeps, a hit whose score EQUALS the threshold included."""
    n, d = 30_000, 256
    x = unit_rows(n, d, 31)
    q = unit_rows(6, d, 32)
    m = np.random.default_rng(3).random(n) < 0.3
    ix = build(x, metric, dtype)
    sh = native.ShardedIndex(d, metric, dtype, [0, 0, 0], block_rows=1024)
    sh.append(x)
    for index in (ix, sh):
        for queries, k, kw in ((q[:1], 10, {}), (q[:1], 100, {}), (q, 10, {}), (q, 100, {}), (q[:1], 200, {}),
                               (q[:1], 10, {"mask": ox.pack_mask(m)}), (q, 10, {"mask": ox.pack_mask(m)})):
            ids, scores, counts = index.search(queries, k, **kw)
            for pos in (0, 4, k - 1):
                t = float(scores[0, pos])                                    # exactly a returned score: must be kept
                for thr in (t, float(np.nextafter(np.float32(t), np.float32(np.inf)))):
                    i2, s2, c2 = index.search(queries, k, min_score=thr, **kw)
                    for j in range(queries.shape[0]):
                        keep = scores[j, :counts[j]] >= np.float32(thr)
                        assert c2[j] == keep.sum(), (k, pos, j, c2[j], keep.sum())
                        np.testing.assert_array_equal(i2[j, :c2[j]], ids[j, :counts[j]][keep])
                        np.testing.assert_array_equal(s2[j, :c2[j]].view(np.uint32), scores[j, :counts[j]][keep].view(np.uint32))
                        assert (i2[j, c2[j]:] == -1).all()
    ids, scores, counts = ix.search(q, 10, min_score=1e9)                    # nothing qualifies
    assert (counts == 0).all() and (ids == -1).all()
    sh.close()


def test_retriever_threshold_runs_inside_the_store():
    """VectorRetriever hands its similarity threshold to a store that supports it; results equal the host-side rule."""
    import asyncio

    from youtu_rag_b200 import B200VectorStore, Chunk, RetrieverConfig, VectorRetriever, VectorStoreConfig
    from youtu_rag_b200.base import BaseEmbedder

    n, d = 5000, 64
    x = unit_rows(n, d, 41)
    qs = unit_rows(3, d, 42)

    class Emb(BaseEmbedder):
        async def embed_texts(self, texts):
            return [qs[int(t)].tolist() for t in texts]

        async def embed_query(self, query):
            return qs[int(query)].tolist()

    s = B200VectorStore(VectorStoreConfig(collection_name="col_thr", index_params={"storage_dtype": "f32"}))
    asyncio.run(s.add_chunks([Chunk(id=f"c{i}", document_id="d", content="", chunk_index=i, metadata={}, embedding=x[i].tolist()) for i in range(n)]))
    plain = asyncio.run(s.search(qs[0].tolist(), 20))
    thr = (plain[7][1] + plain[8][1]) / 2                                     # a Python double between two scores
    r = VectorRetriever(s, Emb(), RetrieverConfig(top_k=20, similarity_threshold=min(1.0, max(0.0, thr))))
    got = asyncio.run(r.retrieve("0"))
    assert [(g.chunk.id, g.rank) for g in got] == [(c.id, i + 1) for i, (c, sc) in enumerate(plain) if sc >= thr]
    batch = asyncio.run(r.batch_retrieve(["0", "1", "2"]))
    assert [(g.chunk.id, g.score) for g in batch[0]] == [(g.chunk.id, g.score) for g in got]
    for j in (1, 2):
        want = [(c.id, sc) for c, sc in asyncio.run(s.search(qs[j].tolist(), 20)) if sc >= thr]
        assert [(g.chunk.id, g.score) for g in batch[j]] == want
    s.close()
