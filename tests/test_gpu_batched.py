"""K2 (tcgen05 GEMM + fused top-k) parity with the oracle and with K1, through the C ABI.  Needs a B200."""

import numpy as np
import pytest

from oracle import exact_search as ox
from tests.helpers import check_topk, unit_rows
from youtu_rag_b200 import native

pytestmark = pytest.mark.gpu


def build(x, metric="cosine"):
    ix = native.Index(x.shape[1], metric, "bf16", 0, 0)
    ix.append(x)
    return ix


@pytest.mark.parametrize("n,d,nq", [(100, 64, 8), (1000, 128, 100), (5000, 768, 128), (20000, 1024, 129),
                                    (40000, 1024, 256), (33333, 200, 300), (19000, 2048, 16)])
def test_k2_matches_oracle(n, d, nq):
    x = unit_rows(n, d, n + d)
    ix = build(x)
    ix.set_path(native.PATH_K2)
    rows = ix.read_rows(np.arange(n))
    qs = unit_rows(nq, d, 7)
    for k in (1, 10, 100, 128):
        ids, scores, counts = ix.search(qs, k)
        for j in list(range(0, nq, max(1, nq // 9))) + [nq - 1]:
            c = int(counts[j])
            assert c == min(k, n)
            check_topk(ids[j, :c], scores[j, :c], rows, ox.prepare(qs[j], "cosine", "bf16")[0], k, "cosine", "bf16")


def test_k2_equals_k1_every_query():
    n, d, nq, k = 60000, 1024, 256, 100
    x = unit_rows(n, d, 3)
    x[[5, 700, 59999]] = x[5]                      # exact ties
    ix = build(x)
    qs = unit_rows(nq, d, 4)
    qs[0] = x[5]
    ix.set_path(native.PATH_K2)
    a = ix.search(qs, k)
    ix.set_path(native.PATH_K1)
    b = ix.search(qs, k)
    assert a[0][0, :3].tolist() == [5, 700, 59999]
    same = (a[0] == b[0]).all(axis=1)
    # tensor-core and FMA accumulation orders differ: allow swaps only between scores tied to ~1e-6
    for j in np.flatnonzero(~same):
        diff = np.flatnonzero(a[0][j] != b[0][j])
        assert np.abs(a[1][j][diff] - b[1][j][diff]).max() < 3e-6
    assert same.mean() > 0.9
    np.testing.assert_allclose(a[1], b[1], rtol=1e-4, atol=3e-6)


@pytest.mark.parametrize("sel", [0.0, 0.01, 0.1, 0.9])
def test_k2_mask(sel):
    n, d, nq = 30000, 256, 64
    x = unit_rows(n, d, 5)
    ix = build(x)
    ix.set_path(native.PATH_K2)
    rows = ix.read_rows(np.arange(n))
    mask = np.random.default_rng(6).random(n) < sel
    qs = unit_rows(nq, d, 8)
    ids, scores, counts = ix.search(qs, 10, mask=ox.pack_mask(mask))
    for j in (0, 31, 63):
        c = int(counts[j])
        assert c == min(10, int(mask.sum()))
        check_topk(ids[j, :c], scores[j, :c], rows, ox.prepare(qs[j], "cosine", "bf16")[0], 10, "cosine", "bf16", mask=mask)


def test_k2_dot_metric_and_tombstones():
    n, d, nq = 10000, 128, 32
    x = unit_rows(n, d, 9) * np.random.default_rng(1).uniform(0.5, 2, (n, 1)).astype(np.float32)
    ix = build(x, "dot")
    ix.set_path(native.PATH_K2)
    rows = ix.read_rows(np.arange(n))
    qs = unit_rows(nq, d, 10) * 1.5
    ids, scores, counts = ix.search(qs, 20)
    check_topk(ids[3], scores[3], rows, ox.prepare(qs[3], "dot", "bf16")[0], 20, "dot", "bf16")
    dead = ids[3, :5].tolist()
    ix.set_live(dead, False)
    live = np.ones(n, bool); live[dead] = False
    ids2, scores2, _ = ix.search(qs, 20)
    check_topk(ids2[3], scores2[3], rows, ox.prepare(qs[3], "dot", "bf16")[0], 20, "dot", "bf16", mask=live)


def test_k2_adversarial_order_forces_compaction():
    """Rows sorted by increasing score for query 0: every later tile beats the threshold, so candidate
    buffers overflow and the in-kernel compaction path has to be exact."""
    n, d = 148 * 128 * 6, 64
    rng = np.random.default_rng(11)
    q = unit_rows(1, d, 12)[0]
    noise = unit_rows(n, d, 13)
    w = np.linspace(-0.9, 0.9, n, dtype=np.float32)[:, None]
    x = w * q[None, :] + 0.3 * noise
    ix = build(x)
    ix.set_path(native.PATH_K2)
    rows = ix.read_rows(np.arange(n))
    qs = np.concatenate([q[None], unit_rows(15, d, 14)])
    ids, scores, counts = ix.search(qs, 100)
    for j in (0, 1, 15):
        check_topk(ids[j], scores[j], rows, ox.prepare(qs[j], "cosine", "bf16")[0], 100, "cosine", "bf16")


def test_auto_path_uses_k2_for_batches_and_store_batch_api():
    import asyncio

    from tests.golden_util import GOLDEN, golden_chunks
    from youtu_rag_b200 import B200VectorStore, VectorStoreConfig

    s = B200VectorStore(VectorStoreConfig(collection_name="col_b"))
    asyncio.run(s.add_chunks(golden_chunks()))
    qs = np.asarray(GOLDEN["queries"] * 4, np.float32)          # 16 queries → batched kernel
    batch = asyncio.run(s.search_batch(qs, top_k=5, filters={"source": {"$in": ["file0.pdf", "file3.pdf"]}}))
    for j in range(16):
        single = asyncio.run(s.search(qs[j].tolist(), top_k=5, filters={"source": {"$in": ["file0.pdf", "file3.pdf"]}}))
        assert [c.id for c, _ in batch[j]] == [c.id for c, _ in single]
        np.testing.assert_allclose([sc for _, sc in batch[j]], [sc for _, sc in single], atol=3e-6)


@pytest.mark.parametrize("dtype", ["bf16", "f32"])
def test_batched_euclidean(dtype):
    n, d, nq = 20000, 256, 40
    x = unit_rows(n, d, 31) * np.random.default_rng(32).uniform(0.5, 1.5, (n, 1)).astype(np.float32)
    ix = native.Index(d, "euclidean", dtype, 0, 0)
    ix.append(x)
    rows = ix.read_rows(np.arange(n))
    qs = unit_rows(nq, d, 33)
    ids, scores, counts = ix.search(qs, 10)          # bf16 → K2 with the euclidean epilogue; f32 → K1Q
    for j in (0, 17, 39):
        check_topk(ids[j], scores[j], rows, ox.prepare(qs[j], "euclidean", dtype)[0], 10, "euclidean", dtype, tie_eps=2e-5)


@pytest.mark.parametrize("nq", [3, 24])
def test_per_query_filters(nq):
    """One metadata filter per query: K1 loop for small batches, K2 with per-query masks for large ones."""
    import asyncio

    from oracle import where_eval as ow
    from youtu_rag_b200 import B200VectorStore, Chunk, VectorStoreConfig

    n, d = 6000, 64
    x = unit_rows(n, d, 41)
    metas = [{"table_name": f"t{i % 7}", "column_name": f"c{i % 3}", "type": "column_value", "v": i % 11} for i in range(n)]
    s = B200VectorStore(VectorStoreConfig(collection_name="col_pq"))
    asyncio.run(s.add_chunks([Chunk(id=f"r{i}", document_id="d", content="", chunk_index=i, metadata=metas[i],
                                    embedding=x[i].tolist()) for i in range(n)]))
    q = unit_rows(1, d, 42)[0]
    filters = [None if j % 5 == 4 else {"$and": [{"type": "column_value"}, {"table_name": f"t{j % 7}"}, {"column_name": f"c{j % 3}"}]}
               for j in range(nq)]
    filters[1] = {"v": {"$gte": 9}}
    got = asyncio.run(s.search_batch(np.tile(q, (nq, 1)), top_k=3, filters=filters))
    rows = s.index.read_rows(np.arange(n))
    full_meta = [{"document_id": "d", "chunk_index": i, **metas[i]} for i in range(n)]
    for j in range(nq):
        mask = ow.eval_where(filters[j], full_meta)
        ids = np.array([int(c.id[1:]) for c, _ in got[j]])
        sc = np.array([v for _, v in got[j]])
        check_topk(ids, sc, rows, ox.prepare(q, "cosine", "bf16")[0], 3, "cosine", "bf16", mask=mask)
        single = asyncio.run(s.search(q.tolist(), top_k=3, filters=filters[j]))
        assert [c.id for c, _ in single] == [c.id for c, _ in got[j]]


def test_k2_pair_kernel_matches_single_cta_kernel():
    """cta_group::2 pair kernel (path 4) vs the one-CTA kernel (path 2) vs the oracle, incl. masks and euclidean."""
    n, d, nq = 70000, 512, 200
    x = unit_rows(n, d, 51)
    x[[9, 40000]] = x[9]
    qs = unit_rows(nq, d, 52)
    qs[130] = x[9]
    mask = np.random.default_rng(53).random(n) < 0.2
    mask[[9, 40000]] = True
    for metric in ("cosine", "euclidean"):
        ix = build(x, metric)
        rows = ix.read_rows(np.arange(n))
        for m in (None, mask):
            pm = None if m is None else ox.pack_mask(m)
            ix.set_path(native.PATH_K2_PAIR)
            a = ix.search(qs, 50, mask=pm)
            ix.set_path(native.PATH_K2)
            b = ix.search(qs, 50, mask=pm)
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
            for j in (0, 127, 128, 130, 199):
                check_topk(a[0][j], a[1][j], rows, ox.prepare(qs[j], metric, "bf16")[0], 50, metric, "bf16", mask=m,
                           tie_eps=2e-5 if metric == "euclidean" else None)
            if metric == "cosine":
                assert a[0][130, :2].tolist() == [9, 40000]


@pytest.mark.parametrize("sel", [0.0005, 0.02, 0.2, 0.3])
@pytest.mark.parametrize("metric", ["cosine", "euclidean"])
def test_k2_compaction_for_selective_shared_mask(sel, metric):
    """>= 65536 rows and a shared mask passing <= 25 % of them: rows are gathered (K8) and K2 runs on the compact
    matrix; 30 % stays on the dense masked path; 0.05 % (< k rows pass) falls back as well.  All must equal the oracle."""
    n, d, nq, k = 100_003, 128, 40, 60
    x = unit_rows(n, d, 61) * (1.0 if metric == "cosine" else 1.3)
    mask = np.random.default_rng(62).random(n) < sel
    x[[11, 70_000]] = x[11]
    mask[[11, 70_000]] = True
    ix = build(x, metric)
    rows = ix.read_rows(np.arange(n))
    qs = unit_rows(nq, d, 63)
    qs[5] = x[11]
    ids, scores, counts = ix.search(qs, k, mask=ox.pack_mask(mask))
    for j in (0, 5, 39):
        c = int(counts[j])
        assert c == min(k, int(mask.sum()))
        check_topk(ids[j, :c], scores[j, :c], rows, ox.prepare(qs[j], metric, "bf16")[0], k, metric, "bf16", mask=mask,
                   tie_eps=2e-5 if metric == "euclidean" else None)
    if metric == "cosine":
        assert ids[5, :2].tolist() == [11, 70_000]
    # tombstones ride in the same mask
    dead = ids[0, :3].tolist()
    ix.set_live(dead, False)
    m2 = mask.copy(); m2[dead] = False
    ids2, scores2, counts2 = ix.search(qs, k, mask=ox.pack_mask(mask))
    c = int(counts2[0])
    check_topk(ids2[0, :c], scores2[0, :c], rows, ox.prepare(qs[0], metric, "bf16")[0], k, metric, "bf16", mask=m2,
               tie_eps=2e-5 if metric == "euclidean" else None)


@pytest.mark.parametrize("metric", ["cosine", "dot", "euclidean"])
@pytest.mark.parametrize("nq", [2, 4, 5, 9])
def test_k1q_fp32_batches_match_oracle_and_the_k1_loop(metric, nq):
    """fp32 storage: batches go through K1Q (four queries per pass over the rows).  Same per-row arithmetic as
    K1, so ids AND score bits must equal a K1 loop; both must match the oracle."""
    n, d = 30011, 1024 if nq == 4 else 200
    x = unit_rows(n, d, 41) * np.random.default_rng(42).uniform(0.5, 1.5, (n, 1)).astype(np.float32)
    x[[5, 7000, 29999]] = x[5]                       # ties broken by row id
    ix = native.Index(d, metric, "f32", 0, 0)
    ix.append(x)
    rows = ix.read_rows(np.arange(n))
    qs = unit_rows(nq, d, 43) * 1.2
    qs[0] = x[5]
    mask = np.random.default_rng(44).random(n) < 0.3
    mask[[5, 7000, 29999]] = True
    for k in (1, 10, 32):
        for m in (None, mask):
            pm = None if m is None else ox.pack_mask(m)
            ids, scores, counts = ix.search(qs, k, mask=pm)              # auto path → K1Q
            ix.set_path(native.PATH_K1)
            ids1, scores1, counts1 = ix.search(qs, k, mask=pm)           # forced: one K1 scan per query
            ix.set_path(native.PATH_AUTO)
            assert np.array_equal(ids, ids1) and np.array_equal(counts, counts1)
            if metric == "euclidean":   # ||q||^2 comes from K5 here and from K1's fused prologue there
                np.testing.assert_allclose(scores, scores1, rtol=0, atol=1e-6)
            else:
                assert np.array_equal(scores.view(np.uint32), scores1.view(np.uint32))
            for j in range(nq):
                c = int(counts[j])
                assert c == k
                check_topk(ids[j, :c], scores[j, :c], rows, ox.prepare(qs[j], metric, "f32")[0], k, metric, "f32", mask=m,
                           tie_eps=2e-5 if metric == "euclidean" else None)
    if metric == "cosine":
        assert ids[0, :3].tolist() == [5, 7000, 29999]


def test_k1q_fewer_passing_rows_than_k():
    n, d = 5000, 64
    x = unit_rows(n, d, 45)
    ix = native.Index(d, "cosine", "f32", 0, 0)
    ix.append(x)
    mask = np.zeros(n, bool)
    mask[[3, 77, 4999]] = True
    ids, scores, counts = ix.search(unit_rows(6, d, 46), 10, mask=ox.pack_mask(mask))
    assert counts.tolist() == [3] * 6
    assert all(sorted(ids[j, :3].tolist()) == [3, 77, 4999] and (ids[j, 3:] == -1).all() for j in range(6))


@pytest.mark.parametrize("metric", ["cosine", "euclidean"])
def test_batches_beyond_256_queries_sweep_the_corpus_once(metric):
    """More than 256 queries run as ONE launch of the pair kernel over up to four 256-query chunks per row tile (C5's
    shape): 300 / 700 / 1024 / 1300 queries (partial last chunks, a second launch), k = 10 and 100,
    with a shared bitmask, against the oracle — and bit-identical to the same batch issued as 256-query chunks."""
    n, d = 60_000, 192
    x = unit_rows(n, d, 61)
    ix = native.Index(d, metric, "bf16", 0, n)
    ix.append(x)
    rows = ix.read_rows(np.arange(n))
    m = np.random.default_rng(8).random(n) < 0.5
    l0 = ix.launches()
    for nq in (300, 700, 1024, 1300):
        q = unit_rows(nq, d, 62 + nq)
        for k, mask in ((10, None), (100, None), (10, m)):
            ids, scores, counts = ix.search(q, k, mask=None if mask is None else ox.pack_mask(mask))
            assert (counts == k).all()
            for j in sorted({0, 127, 128, 255, 256, 299, nq // 2, nq - 1}):
                check_topk(ids[j], scores[j], rows, ox.prepare(q[j], metric, "bf16")[0], k, metric, "bf16", mask)
            parts = [ix.search(q[a:a + 256], k, mask=None if mask is None else ox.pack_mask(mask)) for a in range(0, nq, 256)]
            np.testing.assert_array_equal(ids, np.concatenate([p[0] for p in parts]))
            np.testing.assert_array_equal(scores.view(np.uint32), np.concatenate([p[1] for p in parts]).view(np.uint32))
    assert ix.launches() > l0
    ix.close()
