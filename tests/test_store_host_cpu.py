"""Host logic of the product's stores on the CPU: B200VectorStore / B200MemoryVectorStore / VectorRetriever run
UNCHANGED over tests/fake_index.py (the oracle behind native.Index's interface), against the same golden outputs
of the reference's glue that the GPU tests use.  Covers id/row bookkeeping, metadata columns and their upload,
filter compilation, tombstones, upsert, result shaping — everything in store.py above the C ABI."""

import asyncio
import json
from pathlib import Path

import numpy as np
import pytest

from tests.fake_index import FakeIndex, FakeShardedIndex
from tests.golden_util import replay_memory_scenario, check_memory_outputs
from youtu_rag_b200 import B200MemoryVectorStore, B200VectorStore, Chunk, VectorStoreConfig, native


@pytest.fixture(autouse=True)
def _oracle_behind_the_index(monkeypatch):
    monkeypatch.setattr(native, "Index", FakeIndex)
    monkeypatch.setattr(native, "ShardedIndex", FakeShardedIndex)


# the GPU tests of the store, collected here without the gpu mark: same bodies, fake index underneath
from tests.test_gpu_store import (  # noqa: E402,F401
    test_persistence_roundtrip_is_bit_identical,
    test_retriever_over_b200_store_matches_reference_retriever,
    test_store_bf16_within_north_star_tolerance,
    test_store_edge_cases,
    test_store_mutations_match_reference_glue,
    test_store_reproduces_reference_glue_f32,
)


from tests.test_gpu_batched import test_per_query_filters  # noqa: E402,F401  (search_batch with one filter per query)
from tests.test_gpu_memory_store import test_memory_store_matches_reference_semantics, test_rescoring_formulas  # noqa: E402,F401


def test_retriever_with_a_reranker_over_the_product_store():
    """Same golden as tests/test_golden_cpu.py::test_retriever_with_a_reranker_reproduces_reference_retriever, but over
    B200VectorStore (search_batch path of batch_retrieve included)."""
    from tests.golden_util import GOLDEN, GoldenReranker, golden_chunks
    from tests.test_gpu_store import _Emb
    from youtu_rag_b200 import B200VectorStore, RetrieverConfig, VectorRetriever

    s = B200VectorStore(VectorStoreConfig(collection_name="col_rr", index_params={"storage_dtype": "f32"}))
    asyncio.run(s.add_chunks(golden_chunks()))
    for rec in GOLDEN["retriever_rerank"]:
        cfg = RetrieverConfig(top_k=4, similarity_threshold=rec["config_threshold"], enable_reranking=True)
        r = VectorRetriever(s, _Emb(), cfg, reranker=GoldenReranker())
        single = asyncio.run(r.retrieve("2", **rec["kwargs"]))
        batch = asyncio.run(r.batch_retrieve(["0", "1"], top_k=3, **rec["kwargs"]))
        for got, want in [(single, rec["single"])] + list(zip(batch, rec["batch"])):
            assert [(x.chunk.id, x.rank, x.score) for x in got] == [(w["id"], w["rank"], w["score"]) for w in want]


def test_memory_store_replays_the_reference_scenario(tmp_path):
    """every step of tests/golden/memory_store.json (the reference's MemoryVectorStore driven by
    tests/golden/make_memory_store_golden.py) gives the same outputs on B200MemoryVectorStore."""
    g = json.loads((Path(__file__).parent / "golden" / "memory_store.json").read_text())
    # the memory store persists by default, like the reference's PersistentClient (memory_store.py:183-199)
    store = B200MemoryVectorStore(VectorStoreConfig(backend="b200", collection_name="agent_memory", persist_directory=str(tmp_path),
                                                    index_params={"storage_dtype": "f32"}))
    got = asyncio.run(replay_memory_scenario(store, Chunk, g["specs"], g["steps"]))
    check_memory_outputs(g, got, tol=2e-6, emb_atol=1e-6)


@pytest.mark.parametrize("persist", [False, True])
def test_store_bookkeeping_against_a_dict_model(persist, tmp_path):
    """Random add / re-add / upsert / delete / delete_by_document_id / delete_by_metadata / clear sequences (with
    persistence: also close + reopen from disk at random points): after every step the store agrees with a plain
    dict model on count, get_by_id and a filtered exact search."""
    from oracle import exact_search as ox
    from oracle import where_eval as ow
    from youtu_rag_b200 import B200VectorStore

    rng = np.random.default_rng(17)
    d = 12
    cfg = VectorStoreConfig(collection_name="model", persist_directory=str(tmp_path),
                            index_params={"storage_dtype": "f32", "persist": persist})
    s = B200VectorStore(cfg)
    model: dict[str, dict] = {}          # id -> {"emb", "meta", "content", "seq"}  (seq = order of the row in the store)
    seq = 0

    def mk(i):
        return Chunk(id=f"c{i}", document_id=f"d{i % 4}", content=f"t{i}-{int(rng.integers(0, 99))}", chunk_index=i % 3,
                     metadata={"source": f"f{i % 3}.pdf", "v": int(rng.integers(0, 10)), "opt": None if i % 2 else "x"},
                     embedding=rng.standard_normal(d).astype(np.float32).tolist())

    def put(c, replace):
        nonlocal seq
        if c.id in model and not replace:
            return                                   # add_chunks ignores ids that already exist (collection.add)
        meta = {"document_id": c.document_id, "chunk_index": c.chunk_index, **{k: v for k, v in c.metadata.items() if v is not None}}
        model[c.id] = {"emb": np.asarray(c.embedding, np.float32), "meta": meta, "content": c.content, "seq": seq}
        seq += 1

    def check():
        assert asyncio.run(s.count()) == len(model)
        probe = [f"c{int(i)}" for i in rng.integers(0, 40, 4)]
        for cid in probe:
            got = asyncio.run(s.get_by_id(cid))
            if cid not in model:
                assert got is None
            else:
                assert (got.content, got.metadata) == (model[cid]["content"], model[cid]["meta"])
        flt = [None, {"source": "f1.pdf"}, {"v": {"$gte": 5}}, {"$and": [{"document_id": "d2"}, {"v": {"$lt": 8}}]}][int(rng.integers(0, 4))]
        q = rng.standard_normal(d).astype(np.float32)
        got = asyncio.run(s.search(q.tolist(), top_k=5, filters=flt))
        ids = sorted(model, key=lambda k: model[k]["seq"])                       # row order = insertion order of live rows
        if ids:
            rows = ox.prepare(np.stack([model[k]["emb"] for k in ids]), "cosine", "f32")
            mask = ow.eval_where(ow.normalize_filters(flt), [model[k]["meta"] for k in ids])
            wi, ws = ox.exact_topk(rows, ox.prepare(q, "cosine", "f32")[0], 5, "cosine", mask)
            assert [c.id for c, _ in got] == [ids[i] for i in wi.tolist()]
            np.testing.assert_allclose([sc for _, sc in got], ws, atol=1e-6)
        else:
            assert got == []

    for step in range(120):
        op = rng.choice(["add", "add", "readd", "upsert", "delete", "doc", "meta", "clear"], p=[.3, .2, .1, .12, .12, .07, .07, .02])
        if op in ("add", "readd"):
            cs = [mk(int(i)) for i in set(rng.integers(0, 40, int(rng.integers(1, 6))).tolist())]
            asyncio.run(s.add_chunks(cs))
            for c in cs:
                put(c, replace=False)
        elif op == "upsert":
            cs = [mk(int(i)) for i in set(rng.integers(0, 40, 3).tolist())]
            asyncio.run(s.upsert_chunks(cs))
            for c in cs:
                put(c, replace=True)
        elif op == "delete":
            ids = [f"c{int(i)}" for i in rng.integers(0, 40, 3)]
            asyncio.run(s.delete(ids))
            for i in ids:
                model.pop(i, None)
        elif op == "doc":
            doc = f"d{int(rng.integers(0, 5))}"
            n = asyncio.run(s.delete_by_document_id(doc))
            gone = [k for k, v in model.items() if v["meta"]["document_id"] == doc]
            assert n == len(gone)
            for k in gone:
                del model[k]
        elif op == "meta":
            f = {"source": f"f{int(rng.integers(0, 3))}.pdf", "chunk_index": int(rng.integers(0, 3))}
            n = asyncio.run(s.delete_by_metadata(f))
            gone = [k for k, v in model.items() if v["meta"]["source"] == f["source"] and v["meta"]["chunk_index"] == f["chunk_index"]]
            assert n == len(gone)
            for k in gone:
                del model[k]
        else:
            asyncio.run(s.clear())
            model.clear()
        if persist and rng.random() < 0.15:
            s.close()
            s = B200VectorStore(cfg)                 # reload: segments + tombstones
        check()


def test_delete_collection_and_orphan_cleanup(tmp_path):
    from tests.golden_util import GOLDEN, golden_chunks
    from youtu_rag_b200 import B200VectorStore

    cfg = VectorStoreConfig(collection_name="kb1", persist_directory=str(tmp_path), index_params={"persist": True, "storage_dtype": "f32"})
    s = B200VectorStore(cfg)
    asyncio.run(s.add_chunks(golden_chunks()[:20]))
    assert (tmp_path / "kb1.b200" / "manifest.json").exists()
    (tmp_path / "torn.b200").mkdir()                       # a writer died before its first manifest
    (tmp_path / "other_dir").mkdir()
    assert B200VectorStore.cleanup_orphaned_directories(str(tmp_path)) == {"deleted_count": 1, "deleted_dirs": ["torn.b200"]}
    assert (tmp_path / "kb1.b200").exists() and (tmp_path / "other_dir").exists()
    assert hasattr(s, "delete_collection")                 # what KnowledgeCleanupManager probes (cleanup_manager.py:651)
    s.delete_collection()
    assert not (tmp_path / "kb1.b200").exists()
    again = B200VectorStore(cfg)                            # a fresh, empty collection of the same name
    assert asyncio.run(again.count()) == 0 and asyncio.run(again.search(GOLDEN["queries"][0], 3)) == []
    assert B200VectorStore.cleanup_orphaned_directories(str(tmp_path / "missing")) == {"deleted_count": 0, "deleted_dirs": []}


def test_collection_names_follow_chromas_rule(tmp_path):
    """memory collections are named memory_<user id>; a name with a separator or '..' must never reach
    mkdir / rmtree.  Chroma (the store this replaces) rejects the same names at create_collection."""
    from youtu_rag_b200.persist import CollectionDir

    for bad in ("../evil", "a/b", "..", "ab", "x" * 513, "a..b", "-abc", "abc-", "memory_../../etc"):
        with pytest.raises(ValueError):
            B200VectorStore(VectorStoreConfig(collection_name=bad, persist_directory=str(tmp_path)))
        with pytest.raises(ValueError):
            CollectionDir(str(tmp_path), bad)
    mem = B200MemoryVectorStore(VectorStoreConfig(collection_name="agent_memory", persist_directory=str(tmp_path)))
    assert mem.delete_collection("memory_../../x") is False          # swallowed like any engine error, nothing touched
    B200VectorStore(VectorStoreConfig(collection_name="memory_user-1.a", persist_directory=str(tmp_path)))


def test_memory_collections_are_listed_and_deleted_on_disk_after_a_restart(tmp_path):
    """list_collections / delete_collection work on the persistent state (memory_store.py:617-643),
    not only on collections this process has opened; memories persist by default."""
    cfg = VectorStoreConfig(collection_name="agent_memory", persist_directory=str(tmp_path), index_params={"storage_dtype": "f32"})
    a = B200MemoryVectorStore(cfg)
    rng = np.random.default_rng(0)
    chunk = Chunk(id="m1", document_id="d", content="hello", chunk_index=0, metadata={}, embedding=rng.standard_normal(8).tolist())
    asyncio.run(a.add_chunks([chunk], collection_name="memory_u1"))
    asyncio.run(a.add_chunks([chunk], collection_name="memory_u2"))
    del a
    b = B200MemoryVectorStore(cfg)                                    # "restart": nothing opened yet
    assert b.list_collections() == ["memory_u1", "memory_u2"]
    assert b.delete_collection("memory_u1") is True
    assert not (tmp_path / "memory_u1.b200").exists()
    assert b.list_collections() == ["memory_u2"]
    assert asyncio.run(b.count("memory_u1")) == 0                     # recreated empty, not resurrected
    assert asyncio.run(b.count("memory_u2")) == 1                     # the other one was reloaded from disk


def test_add_is_all_or_nothing_and_numpy_metadata_is_coerced(tmp_path, monkeypatch):
    """numpy scalars in metadata are stored as plain Python values (json-serialisable); a segment that
    cannot be written rolls the device append back, so memory and disk agree; upsert persists the new row before
    the tombstone."""
    cfg = VectorStoreConfig(collection_name="kb_txn", persist_directory=str(tmp_path), index_params={"persist": True, "storage_dtype": "f32"})
    s = B200VectorStore(cfg)
    rng = np.random.default_rng(1)

    def mk(i, **meta):
        return Chunk(id=f"c{i}", document_id="d", content=f"t{i}", chunk_index=i, metadata=meta, embedding=rng.standard_normal(8).tolist())

    asyncio.run(s.add_chunks([mk(0, n=np.int64(7), f=np.float32(0.5), b=np.bool_(True))]))
    got = asyncio.run(s.get_by_id("c0")).metadata
    assert got["n"] == 7 and type(got["n"]) is int and type(got["f"]) is float and got["b"] is True
    # a failing disk write leaves nothing behind
    from youtu_rag_b200 import persist

    def boom(*a, **k):
        raise OSError("disk full")

    with monkeypatch.context() as mp:
        mp.setattr(persist.CollectionDir, "append_segment", boom)
        with pytest.raises(OSError):
            asyncio.run(s.add_chunks([mk(1), mk(2)]))
    assert asyncio.run(s.count()) == 1 and s.index.rows == 1 and asyncio.run(s.get_by_id("c1")) is None
    asyncio.run(s.add_chunks([mk(1)]))
    asyncio.run(s.upsert_chunks([mk(0, n=8)]))                        # replaces c0
    assert asyncio.run(s.count()) == 2 and asyncio.run(s.get_by_id("c0")).metadata["n"] == 8
    # crash between the new segment and the tombstone list: the reload still sees ONE c0, the newer one
    (tmp_path / "kb_txn.b200" / "deleted.json").unlink()
    s.close()
    r = B200VectorStore(cfg)
    assert asyncio.run(r.count()) == 2 and asyncio.run(r.get_by_id("c0")).metadata["n"] == 8
    hits = asyncio.run(r.search(rng.standard_normal(8).tolist(), top_k=10))
    assert sorted(c.id for c, _ in hits) == ["c0", "c1"]


def test_devices_in_index_params_select_the_sharded_index(tmp_path):
    """g1: `index_params.devices` with more than one GPU makes the store build a ShardedIndex over them (same host logic,
    global row ids); one device — or none — keeps the one-GPU index on that device."""
    from tests.golden_util import GOLDEN, compare_with_golden, golden_chunks

    s = B200VectorStore(VectorStoreConfig(collection_name="col_dev", persist_directory=str(tmp_path),
                                          index_params={"storage_dtype": "f32", "devices": [0, 1, 2, 3], "shard_block_rows": 64, "persist": True}))
    asyncio.run(s.add_chunks(golden_chunks()))
    assert isinstance(s.index, FakeShardedIndex) and s.index.devices == [0, 1, 2, 3] and s.index.block_rows == 64
    for rec in GOLDEN["chroma"]:
        if rec["metric"] == "cosine" and "error" not in rec:
            compare_with_golden(asyncio.run(s.search(GOLDEN["queries"][rec["query"]], rec["top_k"], rec["filters"])), rec["results"], tol=2e-6)
    s.close()
    one = B200VectorStore(VectorStoreConfig(collection_name="col_dev", persist_directory=str(tmp_path),
                                            index_params={"storage_dtype": "f32", "devices": [2], "persist": True}))
    assert type(one.index) is FakeIndex and one.index.device == 2 and asyncio.run(one.count()) == len(golden_chunks())   # reloaded
