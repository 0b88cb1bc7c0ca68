"""B200MemoryVectorStore against an oracle built from the reference's semantics.  Needs a B200."""

import asyncio
import json
from datetime import datetime, timedelta

import numpy as np
import pytest

from oracle import exact_search as ox
from oracle import where_eval as ow
from tests.helpers import unit_rows
from youtu_rag_b200 import B200MemoryVectorStore, Chunk, VectorStoreConfig, rank_memories, rank_skills

pytestmark = pytest.mark.gpu


def run(c):
    return asyncio.run(c)


def _memories(n, d, seed):
    x = unit_rows(n, d, seed)
    now = datetime(2026, 1, 1, 12, 0, 0)
    out = []
    for i in range(n):
        out.append(Chunk(id=f"m{i}", document_id=f"sess{i % 3}", content=f"memory {i}", chunk_index=0, embedding=x[i].tolist(),
                         metadata={"memory_type": ["episodic", "procedural", "working"][i % 3], "session_id": f"s{i % 4}",
                                   "importance_score": round(0.1 * (i % 10), 1), "success_rate": round(0.05 * (i % 20), 2),
                                   "created_at": now - timedelta(hours=i), "tool_sequence": [{"tool": "search", "n": i}],
                                   "metadata": {"k": i}, "none": None}))
    return x, out, now


def test_memory_store_matches_reference_semantics(tmp_path):
    n, d = 300, 32
    x, mem, now = _memories(n, d, 5)
    s = B200MemoryVectorStore(VectorStoreConfig(collection_name="agent_memory", persist_directory=str(tmp_path)))
    coll = s.get_collection_name("u1")
    assert coll == "memory_u1" and s.get_collection_name("u1", "procedural") == "memory_u1_procedural"
    run(s.add_chunks(mem, collection_name=coll))
    assert run(s.count(coll)) == n and run(s.count()) == 0        # collections are isolated
    q = unit_rows(1, d, 6)[0]
    # stored metadata is what the reference writes to Chroma: datetimes → iso strings, lists/dicts → JSON strings
    stored = [{"document_id": c.document_id, "chunk_index": 0,
               **{k: (v.isoformat() if isinstance(v, datetime) else json.dumps(v, ensure_ascii=False) if isinstance(v, (list, dict)) else v)
                  for k, v in c.metadata.items() if v is not None}} for c in mem]
    rows = ox.prepare(x, "cosine", "bf16")
    qp = ox.prepare(q, "cosine", "bf16")[0]
    got = run(s.search_memories(q.tolist(), "u1", session_id="s1", top_k=7, min_importance=0.3))
    where = {"$and": [{"session_id": {"$eq": "s1"}}, {"importance_score": {"$gte": 0.3}}, {"success_rate": {"$gte": 0.2}}]}
    ids, scores = ox.exact_topk(rows, qp, 7, "cosine", ow.eval_where(where, stored))
    assert [c.id for c, _ in got] == [f"m{i}" for i in ids]
    np.testing.assert_allclose([sc for _, sc in got], scores, rtol=1e-3, atol=1e-5)
    c0 = got[0][0]
    i0 = int(c0.id[1:])
    assert c0.metadata["tool_sequence"] == [{"tool": "search", "n": i0}] and c0.metadata["metadata"] == {"k": i0}
    assert c0.metadata["created_at"] == (now - timedelta(hours=i0)).isoformat() and "none" not in c0.metadata
    assert c0.document_id == f"sess{i0 % 3}"
    # a malformed filter is swallowed into [] like memory_store.py:324-326
    assert run(s.search(q.tolist(), 5, {"$and": [{"a": 1}]}, collection_name=coll)) == []
    # upsert replaces in place
    run(s.add_chunks([Chunk(id="m5", document_id="new", content="changed", chunk_index=0, embedding=q.tolist(),
                            metadata={"memory_type": "episodic", "success_rate": 1.0})], collection_name=coll))
    assert run(s.count(coll)) == n
    top = run(s.search(q.tolist(), 1, collection_name=coll))
    assert top[0][0].id == "m5" and top[0][0].content == "changed" and abs(top[0][1] - 1.0) < 4e-3
    g = run(s.get_by_id("m5", coll))
    assert g.document_id == "new" and g.embedding is not None
    # working memory: where + sort by created_at + last max_turns
    wm = run(s.get_working_memory("u1", "s2", max_turns=4))
    want = sorted([c for c in mem if c.metadata["memory_type"] == "working" and c.metadata["session_id"] == "s2" and c.id != "m5"],
                  key=lambda c: c.metadata["created_at"].isoformat())[-4:]
    assert [c.id for c in wm] == [c.id for c in want]
    # cleanup of low success rates lives in the procedural collection
    proc = s.get_collection_name("u1", "procedural")
    run(s.add_chunks([m for m in mem if m.metadata["memory_type"] == "procedural"], collection_name=proc))
    n_proc = run(s.count(proc))
    removed = run(s.cleanup_outdated_memories("u1", 0.2))
    assert removed == sum(1 for m in mem if m.metadata["memory_type"] == "procedural" and m.metadata["success_rate"] < 0.2)
    assert run(s.count(proc)) == n_proc - removed
    assert run(s.delete_by_document_id("sess1", coll)) > 0
    run(s.clear(coll))
    assert run(s.count(coll)) == 0 and s.delete_collection(proc) and proc not in s.list_collections()


def test_rescoring_formulas():
    now = datetime(2026, 1, 2)
    mk = lambda i, imp, age_h, **kw: (Chunk(id=i, document_id="", content="", chunk_index=0,
                                             metadata={"importance_score": imp, "created_at": (now - timedelta(hours=age_h)).isoformat(), **kw}), 0.0)
    a, b = mk("a", 0.9, 48), mk("b", 0.1, 0)
    res = rank_memories([(a[0], 0.5), (b[0], 0.6)], now=now)
    want_a, want_b = 0.5 * 0.5 + 0.3 * 0.9 + 0.2 * 0.25, 0.5 * 0.6 + 0.3 * 0.1 + 0.2 * 1.0
    assert [r[0].id for r in res] == (["a", "b"] if want_a > want_b else ["b", "a"])
    assert abs(dict((r[0].id, r[2]) for r in res)["a"] - want_a) < 1e-12
    s1 = mk("s1", 0.7, 24, success_count=3, failure_count=1, tags=json.dumps(["search"]))
    s2 = mk("s2", 0.7, 24, success_count=0, failure_count=4, tags=json.dumps(["search"]))
    s3 = mk("s3", 0.7, 24, tags=json.dumps(["python"]))
    out = rank_skills([(s1[0], 0.8), (s2[0], 0.9), (s3[0], 0.95)], top_k=5, min_success_rate=0.3, tool_filter=["search"], now=now)
    assert [r[0].id for r in out] == ["s1"]
    assert abs(out[0][2] - (0.4 * 0.8 + 0.3 * 0.7 + 0.2 * 0.75 + 0.1 * 0.5)) < 1e-12


def test_memory_store_replays_the_reference_scenario_on_the_device(tmp_path):
    """tests/golden/memory_store.json — the reference's own MemoryVectorStore driven through a fixed
    script — replayed on B200MemoryVectorStore with the real index underneath (fp32 storage: scores within 1e-5)."""
    import json
    from pathlib import Path

    from tests.golden_util import check_memory_outputs, replay_memory_scenario

    g = json.loads((Path(__file__).parent / "golden" / "memory_store.json").read_text())
    store = B200MemoryVectorStore(VectorStoreConfig(backend="b200", collection_name="agent_memory", persist_directory=str(tmp_path),
                                                    index_params={"storage_dtype": "f32"}))
    got = asyncio.run(replay_memory_scenario(store, Chunk, g["specs"], g["steps"]))
    check_memory_outputs(g, got, tol=1e-5, emb_atol=1e-6)


def test_a_thousand_tiny_collections_share_one_devices_scratch():
    """the memory toolkit keeps one collection per user / memory type.  Every index of a GPU shares
    one stream and one set of K2 candidate buffers (86 MB — round 1 allocated them per index), so 1000 ten-row
    collections, each searched with a batch, stay far below 2 GB of device memory."""
    import torch

    from youtu_rag_b200 import native

    d, n_coll = 1024, 1000
    rng = np.random.default_rng(0)
    torch.cuda.synchronize()
    free0, _ = torch.cuda.mem_get_info()
    colls = []
    q = rng.standard_normal((4, d)).astype(np.float32)
    for c in range(n_coll):
        ix = native.Index(d, "cosine", "bf16", 0, 0)
        x = rng.standard_normal((10, d)).astype(np.float32)
        ix.append(x)
        colls.append((ix, x))
    for ix, x in colls[::50] + colls[-3:]:
        ids, scores, counts = ix.search(q, 3)                     # a batch: the K2 path
        want = np.argsort(-(x / np.linalg.norm(x, axis=1, keepdims=True)) @ (q / np.linalg.norm(q, axis=1, keepdims=True)).T, axis=0)[:3].T
        assert (counts == 3).all() and np.array_equal(ids, want)
    for ix, _ in colls:
        ix.search(q, 3)
    free1, _ = torch.cuda.mem_get_info()
    used = (free0 - free1) / 2**30
    assert used < 2.0, f"{used:.2f} GiB for {n_coll} ten-row collections"
    for ix, _ in colls:
        ix.close()
