"""Pins the oracle and the host-side mirrors to outputs of the reference's own (unmodified) Python
glue, captured by tests/golden/make_golden.py.  CPU only."""

import asyncio

import numpy as np
import pytest

from oracle import exact_search as ox
from tests.golden_util import GOLDEN, compare_with_golden, golden_chunks
from tests.oracle_store import OracleStore
from youtu_rag_b200 import RetrieverConfig, VectorRetriever
from youtu_rag_b200.base import BaseEmbedder


@pytest.mark.parametrize("metric", ["cosine", "dot", "euclidean"])
def test_oracle_reproduces_reference_chroma_glue(metric):
    store = OracleStore(metric, "f32")
    asyncio.run(store.add_chunks(golden_chunks()))
    n = 0
    for rec in GOLDEN["chroma"]:
        if rec["metric"] != metric:
            continue
        q = GOLDEN["queries"][rec["query"]]
        if "error" in rec:
            with pytest.raises(ValueError):
                asyncio.run(store.search(q, rec["top_k"], rec["filters"]))
        else:
            got = asyncio.run(store.search(q, rec["top_k"], rec["filters"]))
            compare_with_golden(got, rec["results"], tol=2e-6)
        n += 1
    assert n == 136


def test_faiss_restatement_reproduces_reference_faiss_glue():
    x = np.asarray(GOLDEN["corpus"]["embeddings"], np.float32)
    metas = GOLDEN["corpus"]["metadatas"]
    ids = [f"doc{i // 8}_chunk_{i % 8}" for i in range(len(metas))]
    for rec in GOLDEN["faiss"]:
        rows = ox.l2_normalize(x) if rec["metric"] == "cosine" else x
        keep = None
        if rec["filters"]:
            keep = lambda i, f=rec["filters"]: all(metas[i].get(k) == v for k, v in f.items())
        got_ids, got_s = ox.faiss_flat_search(rows, np.asarray(GOLDEN["queries"][rec["query"]], np.float32), rec["top_k"],
                                              rec["metric"], keep)
        want = rec["results"]
        assert len(want) == got_ids.shape[0]
        np.testing.assert_allclose(got_s, [w["score"] for w in want], rtol=2e-5, atol=2e-6)
        for g, w, s in zip(got_ids.tolist(), want, got_s.tolist()):
            # fp32 BLAS may order exact duplicates (rows 3/5) either way, also across the k-th boundary
            assert ids[g] == w["id"] or abs(s - w["score"]) < 1e-6


class _Emb(BaseEmbedder):
    async def embed_texts(self, texts):
        return [GOLDEN["queries"][int(t)] for t in texts]

    async def embed_query(self, query):
        return GOLDEN["queries"][int(query)]


def test_retriever_mirror_reproduces_reference_retriever():
    store = OracleStore("cosine", "f32")
    asyncio.run(store.add_chunks(golden_chunks()))
    for rec in GOLDEN["retriever"]:
        r = VectorRetriever(store, _Emb(), RetrieverConfig(top_k=4, similarity_threshold=rec["config_threshold"]))
        single = asyncio.run(r.retrieve("1", **rec["kwargs"]))
        batch = asyncio.run(r.batch_retrieve(["0", "1", "2"], top_k=3, **rec["kwargs"]))
        for got, want in [(single, rec["single"])] + list(zip(batch, rec["batch"])):
            assert [x.rank for x in got] == [w["rank"] for w in want]
            compare_with_golden([(x.chunk, x.score) for x in got], want, tol=2e-6)


def test_retriever_with_a_reranker_reproduces_reference_retriever():
    """a7 with a reranker: 2k hits are fetched, thresholded, reranked to k and sliced (base_retriever.py:61-80); the
    same deterministic reranker ran under the reference's VectorRetriever when the golden was made."""
    from tests.golden_util import GoldenReranker

    store = OracleStore("cosine", "f32")
    asyncio.run(store.add_chunks(golden_chunks()))
    for rec in GOLDEN["retriever_rerank"]:
        cfg = RetrieverConfig(top_k=4, similarity_threshold=rec["config_threshold"], enable_reranking=True)
        r = VectorRetriever(store, _Emb(), cfg, reranker=GoldenReranker())
        single = asyncio.run(r.retrieve("2", **rec["kwargs"]))
        batch = asyncio.run(r.batch_retrieve(["0", "1"], top_k=3, **rec["kwargs"]))
        for got, want in [(single, rec["single"])] + list(zip(batch, rec["batch"])):
            assert [(x.chunk.id, x.rank, x.score) for x in got] == [(w["id"], w["rank"], w["score"]) for w in want]
    with pytest.raises(ValueError):
        VectorRetriever(store, _Emb(), RetrieverConfig(enable_reranking=True))     # no silent skip without a reranker


def test_context_assembler_reproduces_reference():
    from youtu_rag_b200.base import Chunk, RetrievalResult
    from youtu_rag_b200.postprocess import ContextAssembler, dedup_by_file, merge_results

    g = GOLDEN["assembler"]
    hits = [RetrievalResult(chunk=Chunk(id=h["id"], document_id=h["document_id"], content=h["content"],
                                        chunk_index=h["chunk_index"],
                                        metadata=dict(h["metadata_items"]) if h["metadata_items"] else None),
                            score=h["score"], rank=h["rank"])
            for h in g["hits"]]
    for c in g["cases"]:
        assert ContextAssembler(c["max_len"]).assemble(hits, c["include_metadata"], c["style"]) == c["text"], c
    assert ContextAssembler().assemble([]) == ""
    with pytest.raises(ValueError):
        ContextAssembler().assemble(hits, format_style="xml")
    files = dedup_by_file(hits, include_summary=True)
    assert [f["file_name"] for f in files] == ["d0", "file1.pdf", "file0.pdf", "d1"]  # no metadata → document id
    assert files[1]["chunk_id"] == "c1" and "chunk_index" not in files[1]["metadata"] and files[1]["summary"] == ""
    dup = hits[:3] + [RetrievalResult(chunk=hits[1].chunk, score=0.99, rank=1)]
    merged = merge_results(dup)
    assert [(r.chunk.id, r.score) for r in merged] == [("c0", 1.0), ("c1", 0.99), ("c2", hits[2].score)]


def test_filters_emitted_by_the_reference_toolkits_compile_and_evaluate():
    """the `where` dicts produced by the reference's own toolkit code (captured unmodified by
    tests/golden/make_filter_golden.py) pass the product's validator, compile to a K4 program, and that
    program — interpreted on the CPU — selects the rows the oracle selects."""
    import json
    from pathlib import Path

    import numpy as np

    from oracle import where_eval as ow
    from tests.test_where_host import interpret
    from youtu_rag_b200.metadata import MetadataTable
    from youtu_rag_b200.where import compile_where, normalize_filters, validate_where

    g = json.loads((Path(__file__).parent / "golden" / "filter_producers.json").read_text())
    metas = g["metadatas"]
    table = MetadataTable()
    table.append(metas)
    seen = 0
    for case in g["cases"]:
        where = case["where"]
        if where is None:
            assert case["rows"] is None
            continue
        validate_where(where)
        handed = normalize_filters(where)                 # what the store hands the engine (chroma_store.py:104-116)
        assert handed == where or set(where) == set(handed) and all(handed[k] == {"$eq": v} for k, v in where.items())
        assert np.array_equal(ow.eval_where(handed, metas), ow.eval_where(where, metas))
        want = np.zeros(len(metas), bool)
        want[case["rows"]] = True
        assert np.array_equal(ow.eval_where(where, metas), want), case["producer"]
        prog, _cols = compile_where(where, table)
        assert np.array_equal(interpret(prog, table), want), (case["producer"], where)
        seen += 1
    assert seen >= 13


def test_memory_and_skill_rescoring_matches_the_reference_toolkit():
    """rank_memories / rank_skills against VectorMemoryToolkit.search_memories / search_skills run
    unmodified over preset hits (tests/golden/make_memory_golden.py): same order, same relevance."""
    import json
    from datetime import datetime
    from pathlib import Path

    from youtu_rag_b200 import Chunk, rank_memories, rank_skills

    g = json.loads((Path(__file__).parent / "golden" / "memory_rescoring.json").read_text())
    now = datetime.fromisoformat(g["now"])

    def chunks(hits):
        return [(Chunk(id=h["id"], document_id="d", content=h["document"], chunk_index=0, metadata=dict(h["metadata"])),
                 1.0 - h["distance"]) for h in hits]

    import numpy as np

    from oracle import where_eval as ow
    from tests.test_where_host import interpret
    from youtu_rag_b200.metadata import MetadataTable
    from youtu_rag_b200.where import compile_where, validate_where

    for case in g["memories"]:
        where = case["query"]["where"]          # shorthand leaves inside $and (memory_toolkit.py:880-894)
        if where is not None:
            metas = [h["metadata"] for h in g["memories"][0]["hits"]]
            table = MetadataTable()
            table.append(metas)
            validate_where(where)
            prog, _ = compile_where(where, table)
            assert np.array_equal(interpret(prog, table), ow.eval_where(where, metas)), where
        got = rank_memories(chunks(case["hits"]), now=now)
        assert [c.id for c, _, _ in got] == [r["id"] for r in case["results"]], case["kwargs"]
        for (c, sim, rel), r in zip(got, case["results"]):
            assert abs(sim - (1.0 - r["distance"])) < 1e-12 and abs(rel - r["relevance"]) < 1e-12
    for case in g["skills"]:
        kw = case["kwargs"]
        assert case["query"]["n_results"] == 2 * kw["top_k"] and case["query"]["where"] is None   # callers over-fetch 2k
        tf = kw.get("tool_filter")
        got = rank_skills(chunks(case["hits"]), top_k=kw["top_k"], min_success_rate=kw.get("min_success_rate", 0.3),
                          tool_filter=[tf] if isinstance(tf, str) else tf, now=now)
        assert [c.id for c, _, _ in got] == [r["id"] for r in case["results"]], kw
        for (c, sim, rel), r in zip(got, case["results"]):
            assert abs(rel - r["relevance"]) < 1e-12


def test_postprocessing_matches_the_reference_toolkits():
    """per-file dedup of file-level search and the merge of several searches' hits, against
    KBSearchToolkit.kb_file_search / MetaRetrievalToolkit.merge_retrieval_results run unmodified
    (tests/golden/make_postprocess_golden.py); the filter kb_file_search composes must compile for K4."""
    import json
    from pathlib import Path

    from youtu_rag_b200 import Chunk, RetrievalResult
    from youtu_rag_b200.metadata import MetadataTable
    from youtu_rag_b200.postprocess import dedup_by_file, merge_results
    from youtu_rag_b200.where import compile_where, validate_where

    g = json.loads((Path(__file__).parent / "golden" / "postprocess.json").read_text())

    def rr(h):
        return RetrievalResult(chunk=Chunk(id=h["id"], document_id=h["document_id"], content=h["content"],
                                           chunk_index=h["chunk_index"], metadata=dict(h["metadata"])), score=h["score"], rank=h["rank"])

    hits = [rr(h) for h in g["hits"]]
    table = MetadataTable()
    table.append([h["metadata"] for h in g["hits"]])
    for case in g["file_search"]:
        kw, want = case["kwargs"], case["output"]
        top_k = kw["top_k"] if kw["top_k"] is not None else 4              # KBSelf.file_search_top_k in the generator
        assert case["retriever_top_k"] == top_k                            # auto_rerank=False: no over-fetch
        validate_where(case["filters"])                                     # {"$and": [base, {"index_type": {"$eq": "index_summary"}}]}
        compile_where(case["filters"], table)
        files = dedup_by_file(hits, include_summary=kw["include_summary"])
        assert want["total_files"] == len(files)
        shaped = []
        for idx, e in enumerate(files[:top_k], 1):
            x = {"rank": idx, "file_name": e["file_name"], "embedding_score": round(e["relevance_score"], 4), "metadata": e["metadata"]}
            if kw["include_summary"]:
                x["summary"] = e.get("summary", "")
            shaped.append(x)
        assert shaped == want["files"], kw
    for case in g["merge"]:
        got = merge_results([rr(h) for h in case["input"]])
        want = case["output"]["results"]
        assert [(r.chunk.id, r.rank, round(r.score, 4)) for r in got] == [(w["chunk_id"], w["rank"], w["similarity_score"]) for w in want]
        assert [r.chunk.metadata for r in got] == [w["metadata"] for w in want]
