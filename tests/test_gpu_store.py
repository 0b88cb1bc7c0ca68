"""B200VectorStore (the drop-in BaseVectorStore) against outputs of the reference's own glue and
against the oracle store.  Needs a B200."""

import asyncio

import numpy as np
import pytest

from tests.golden_util import GOLDEN, compare_with_golden, golden_chunks
from tests.oracle_store import OracleStore
from youtu_rag_b200 import (B200VectorStore, RetrieverConfig, VectorRetriever, VectorStoreConfig, VectorStoreFactory)
from youtu_rag_b200.base import BaseEmbedder

pytestmark = pytest.mark.gpu


def run(c):
    return asyncio.run(c)


def make(metric, dtype):
    cfg = VectorStoreConfig(backend="b200", collection_name="col_t", distance_metric=metric,
                            index_params={"storage_dtype": dtype})
    s = VectorStoreFactory.create(cfg)
    assert isinstance(s, B200VectorStore) and s.config.backend == "b200"
    run(s.add_chunks(golden_chunks()))
    return s


@pytest.mark.parametrize("metric", ["cosine", "dot", "euclidean"])
def test_store_reproduces_reference_glue_f32(metric):
    s = make(metric, "f32")
    for rec in GOLDEN["chroma"]:
        if rec["metric"] != metric:
            continue
        q = GOLDEN["queries"][rec["query"]]
        if "error" in rec:
            with pytest.raises(ValueError):
                run(s.search(q, rec["top_k"], rec["filters"]))
        else:
            compare_with_golden(run(s.search(q, rec["top_k"], rec["filters"])), rec["results"], tol=1e-5)


def test_store_bf16_within_north_star_tolerance():
    s = make("cosine", "bf16")
    o = OracleStore("cosine", "bf16")
    run(o.add_chunks(golden_chunks()))
    for rec in GOLDEN["chroma"]:
        if rec["metric"] != "cosine" or "error" in rec:
            continue
        q = GOLDEN["queries"][rec["query"]]
        got = run(s.search(q, rec["top_k"], rec["filters"]))
        want = run(o.search(q, rec["top_k"], rec["filters"]))
        # same stored bf16 operands on both sides → ids identical up to exact ties
        compare_with_golden(got, [{"id": c.id, "score": sc} for c, sc in want], tol=1e-5)
        # and within 1e-3 of the fp32 reference scores
        ref = {w["id"]: w["score"] for w in rec["results"]}
        for c, sc in got:
            if c.id in ref:
                assert abs(sc - ref[c.id]) <= 1e-3 * max(1.0, abs(ref[c.id])) + 2e-3


def test_store_mutations_match_reference_glue():
    s = make("cosine", "f32")
    mut = GOLDEN["chroma_mutations"]
    assert run(s.count()) == mut["count0"]
    assert run(s.delete_by_document_id("doc2")) == mut["deleted_doc2"]
    assert run(s.delete_by_metadata({"source": "file1.pdf", "index_type": "index_summary"})) == mut["deleted_meta"]
    run(s.delete(["doc0_chunk_0", "nope"]))
    assert run(s.count()) == mut["count1"]
    g = run(s.get_by_id("doc0_chunk_2"))
    assert {"id": g.id, "document_id": g.document_id, "chunk_index": g.chunk_index, "metadata": g.metadata} == mut["get"]
    emb = np.asarray(GOLDEN["corpus"]["embeddings"][2], np.float32)
    np.testing.assert_allclose(g.embedding, emb / np.linalg.norm(emb), atol=1e-6)
    assert run(s.get_by_id("doc0_chunk_0")) is None
    compare_with_golden(run(s.search(GOLDEN["queries"][0], top_k=5)), mut["search_after"], tol=1e-5)
    run(s.clear())
    assert run(s.count()) == mut["count2"] and run(s.search(GOLDEN["queries"][0], top_k=5)) == []
    run(s.add_chunks(golden_chunks()[:10]))            # usable again after clear
    assert run(s.count()) == 10 and len(run(s.search(GOLDEN["queries"][0], top_k=3))) == 3


def test_store_edge_cases():
    s = B200VectorStore(VectorStoreConfig(collection_name="col_e"))
    assert run(s.search([0.0] * 8, 3)) == [] and run(s.count()) == 0
    run(s.add_chunks([]))
    ch = golden_chunks()
    run(s.add_chunks(ch[:5]))
    run(s.add_chunks(ch[:7]))                           # existing ids ignored, new ones added
    assert run(s.count()) == 7
    with pytest.raises(ValueError):
        run(s.add_chunks([ch[9], ch[9]]))
    with pytest.raises(ValueError):
        run(s.search([0.0] * 3, 3))                     # wrong dimension
    with pytest.raises(ValueError):
        run(s.search(GOLDEN["queries"][0], 0))
    res = run(s.search(GOLDEN["queries"][0], 50))       # k > count
    assert len(res) == 7
    zero = run(s.search([0.0] * 16, 2))                 # zero query: all cosine scores 0, ids by row order
    assert [c.id for c, _ in zero] == [ch[0].id, ch[1].id] and all(sc == 0.0 for _, sc in zero)


class _Emb(BaseEmbedder):
    async def embed_texts(self, texts):
        return [GOLDEN["queries"][int(t)] for t in texts]

    async def embed_query(self, query):
        return GOLDEN["queries"][int(query)]


def test_retriever_over_b200_store_matches_reference_retriever():
    s = make("cosine", "f32")
    for rec in GOLDEN["retriever"]:
        r = VectorRetriever(s, _Emb(), RetrieverConfig(top_k=4, similarity_threshold=rec["config_threshold"]))
        single = run(r.retrieve("1", **rec["kwargs"]))
        batch = run(r.batch_retrieve(["0", "1", "2"], top_k=3, **rec["kwargs"]))
        for got, want in [(single, rec["single"])] + list(zip(batch, rec["batch"])):
            assert [x.rank for x in got] == [w["rank"] for w in want]
            compare_with_golden([(x.chunk, x.score) for x in got], want, tol=1e-5)


def test_persistence_roundtrip_is_bit_identical(tmp_path):
    """f1: add → delete → reopen from disk → same ids, same scores (bitwise), same metadata; then keep writing."""
    cfg = VectorStoreConfig(collection_name="col_p", persist_directory=str(tmp_path), distance_metric="cosine",
                            index_params={"persist": True})
    a = B200VectorStore(cfg)
    ch = golden_chunks()
    run(a.add_chunks(ch[:40]))
    run(a.add_chunks(ch[40:]))
    run(a.delete_by_document_id("doc2"))
    run(a.delete(["doc0_chunk_0"]))
    want = [run(a.search(q, 7, {"source": {"$in": ["file0.pdf", "file1.pdf"]}})) for q in GOLDEN["queries"]]
    n = run(a.count())
    a.close()
    b = B200VectorStore(cfg)
    assert run(b.count()) == n and run(b.get_by_id("doc0_chunk_0")) is None
    for q, w in zip(GOLDEN["queries"], want):
        got = run(b.search(q, 7, {"source": {"$in": ["file0.pdf", "file1.pdf"]}}))
        assert [(c.id, s, c.metadata, c.content) for c, s in got] == [(c.id, s, c.metadata, c.content) for c, s in w]
    run(b.add_chunks(ch[:3]))                 # doc0_chunk_0 was deleted → can be added again; the other two exist
    assert run(b.count()) == n + 1
    b.close()
    c = B200VectorStore(cfg)
    assert run(c.count()) == n + 1 and run(c.get_by_id("doc0_chunk_0")) is not None
    run(c.clear())
    assert not (tmp_path / "col_p.b200").exists()
    assert run(B200VectorStore(cfg).count()) == 0


@pytest.mark.parametrize("dtype", ["f32", "bf16"])
def test_c1_config_through_the_store(dtype):
    """BASELINE.json configs[0]: 10k chunks x 1024-d unit embeddings, single query, top-10 cosine, via the
    vector-store interface — against the exact oracle and the FAISS-semantics CPU restatement."""
    from oracle import exact_search as ox
    from tests.helpers import unit_rows
    from youtu_rag_b200.base import Chunk

    n, d = 10_000, 1024
    x = unit_rows(n, d, 0)
    s = B200VectorStore(VectorStoreConfig(collection_name="col_c1", distance_metric="cosine", index_params={"storage_dtype": dtype}))
    for a in range(0, n, 2500):
        run(s.add_chunks([Chunk(id=f"k{i}", document_id=f"doc{i // 100}", content=f"chunk {i}", chunk_index=i % 100,
                                embedding=x[i].tolist()) for i in range(a, a + 2500)]))
    assert run(s.count()) == n
    rows = ox.prepare(x, "cosine", dtype)
    for q in unit_rows(5, d, 1):
        got = run(s.search(q.tolist(), top_k=10))
        ids, scores = ox.exact_topk(rows, ox.prepare(q, "cosine", dtype)[0], 10, "cosine")
        assert [c.id for c, _ in got] == [f"k{i}" for i in ids]
        np.testing.assert_allclose([sc for _, sc in got], scores, rtol=1e-5 if dtype == "f32" else 1e-3, atol=1e-6)
        if dtype == "f32":
            f_ids, f_sim = ox.faiss_flat_search(ox.l2_normalize(x), q, 10, "cosine")     # the reference's exact variant
            assert [c.id for c, _ in got] == [f"k{i}" for i in f_ids]
            np.testing.assert_allclose([sc for _, sc in got], f_sim, rtol=1e-5, atol=1e-6)
