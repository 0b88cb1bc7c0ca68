"""Generate tests/golden/memory_store.json: the reference's MemoryVectorStore replayed on a script.

utu/rag/storage/implementations/memory_store.py:163-643 is imported UNMODIFIED (with the exact fake chromadb of
make_golden.py underneath, since chromadb is not installable here) and driven through a fixed scenario: adds with
datetime / list / dict / None metadata, an upsert, searches with the three filter shapes, search_memories,
working memory, get_by_id, the delete family, cleanup, clear, collection bookkeeping.  Every step's inputs and
outputs are recorded; tests replay the same steps on B200MemoryVectorStore.

Usage: python tests/golden/make_memory_store_golden.py     (only where /root/reference exists)
"""

from __future__ import annotations

import asyncio
import importlib
import json
import sys
from datetime import datetime, timedelta
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
sys.path.insert(0, str(HERE.parent.parent))

import make_golden as mg  # noqa: E402
from tests.golden_util import replay_memory_scenario  # noqa: E402

T0 = datetime(2025, 5, 1, 9, 0, 0)


def chunk_specs(d=24, seed=21):
    """Chunk definitions as JSON-able dicts; `created_at` is an ISO string here and a datetime when handed to the store."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((40, d)).astype(np.float32)
    x[9] = x[2]                      # duplicate direction: tie broken by insertion order
    specs = []
    for i in range(40):
        mtype = ["episodic", "procedural", "working", "semantic"][i % 4]
        meta = {"user_id": "u1", "session_id": f"s{i % 3}", "memory_type": mtype, "importance_score": round(0.1 * (i % 10), 1),
                "success_rate": round(0.05 * (i % 20), 2), "created_at": (T0 + timedelta(minutes=7 * ((i * 13) % 40))).isoformat(),
                "avg_latency": float(i), "tool_sequence": [{"tool": "search", "args": {"q": str(i)}}] if i % 5 == 0 else [],
                "metadata": {"note": f"n{i}", "k": i} if i % 6 == 0 else None, "tags": ["a", "b"] if i % 7 == 0 else None,
                "optional": None}
        specs.append({"id": f"mem{i}", "document_id": f"traj{i % 4}", "content": f"memory text {i}", "chunk_index": i % 5,
                      "metadata": meta, "embedding": x[i].tolist()})
    return specs


def scenario(specs):
    """The script, as data: (method, kwargs).  `chunks` are indices into specs; `q` is an index whose embedding is the query."""
    coll = "memory_u1"
    proc = "memory_u1_procedural"
    return [
        ("add_chunks", {"chunks": list(range(0, 30)), "collection_name": coll}),
        ("add_chunks", {"chunks": [i for i in range(40) if i % 4 == 1], "collection_name": proc}),
        ("add_chunks", {"chunks": list(range(30, 36)), "collection_name": None}),            # default collection
        ("count", {"collection_name": coll}), ("count", {"collection_name": proc}), ("count", {"collection_name": None}),
        ("search", {"q": 2, "top_k": 5, "filters": None, "collection_name": coll}),
        ("search", {"q": 7, "top_k": 4, "filters": {"memory_type": "episodic"}, "collection_name": coll}),
        ("search", {"q": 7, "top_k": 4, "filters": {"importance_score": {"$gte": 0.5}}, "collection_name": coll}),
        ("search", {"q": 11, "top_k": 6, "filters": {"$and": [{"session_id": {"$eq": "s1"}}, {"success_rate": {"$lt": 0.6}}]}, "collection_name": coll}),
        ("search", {"q": 3, "top_k": 3, "filters": {"memory_type": {"$regex": "x"}}, "collection_name": coll}),   # engine error -> []
        ("search", {"q": 31, "top_k": 3, "filters": None, "collection_name": None}),
        ("search_memories", {"q": 5, "user_id": "u1", "memory_type": "procedural", "session_id": None, "top_k": 4, "min_importance": 0.0, "include_outdated": False}),
        ("search_memories", {"q": 5, "user_id": "u1", "memory_type": "procedural", "session_id": "s2", "top_k": 4, "min_importance": 0.3, "include_outdated": True}),
        ("search_memories", {"q": 6, "user_id": "u1", "memory_type": None, "session_id": "s0", "top_k": 5, "min_importance": 0.0, "include_outdated": True}),
        ("get_working_memory", {"user_id": "u1", "session_id": "s2", "max_turns": 3}),
        ("get_working_memory", {"user_id": "u1", "session_id": "s0", "max_turns": 10}),
        ("get_by_id", {"chunk_id": "mem6", "collection_name": coll}), ("get_by_id", {"chunk_id": "nope", "collection_name": coll}),
        ("upsert", {"chunk": 6, "new_embedding_from": 4, "importance_score": 0.95, "collection_name": coll}),  # same id, new vector + metadata
        ("count", {"collection_name": coll}),
        ("get_by_id", {"chunk_id": "mem6", "collection_name": coll}),
        ("search", {"q": 4, "top_k": 4, "filters": None, "collection_name": coll}),
        ("search", {"q": 2, "top_k": 4, "filters": None, "collection_name": coll}),
        ("delete", {"chunk_ids": ["mem2", "mem3", "ghost"], "collection_name": coll}),
        ("search", {"q": 2, "top_k": 4, "filters": None, "collection_name": coll}),
        ("delete_by_document_id", {"document_id": "traj1", "collection_name": coll}),
        ("delete_by_document_id", {"document_id": "traj9", "collection_name": coll}),
        ("delete_by_metadata", {"metadata_filter": {"session_id": "s0"}, "collection_name": coll}),
        ("delete_by_metadata", {"metadata_filter": {"session_id": "s1", "memory_type": "semantic"}, "collection_name": coll}),
        ("count", {"collection_name": coll}),
        ("cleanup_outdated_memories", {"user_id": "u1", "success_rate_threshold": 0.3}),
        ("count", {"collection_name": proc}),
        ("search", {"q": 5, "top_k": 20, "filters": None, "collection_name": proc}),
        ("clear", {"collection_name": None}), ("count", {"collection_name": None}),
        ("get_collection_name", {"user_id": "bob", "memory_type": None}), ("get_collection_name", {"user_id": "bob", "memory_type": "semantic"}),
        ("delete_collection", {"collection_name": proc}),
        ("count", {"collection_name": proc}),
    ]


def run():
    mg._install_fakes()
    ref = mg._load_reference()
    ms = importlib.import_module("utu.rag.storage.implementations.memory_store")
    Chunk, Cfg = ref["base"].Chunk, ref["config"].VectorStoreConfig
    store = ms.MemoryVectorStore(config=Cfg(collection_name="agent_memory", persist_directory="/tmp/unused_fake_dir"))
    specs = chunk_specs()
    steps = scenario(specs)
    outputs = asyncio.run(replay_memory_scenario(store, Chunk, specs, [[n, k] for n, k in steps]))
    data = {"specs": specs, "steps": [[n, k] for n, k in steps], "outputs": outputs}
    (HERE / "memory_store.json").write_text(json.dumps(data, ensure_ascii=False, separators=(",", ":")))
    print("wrote", HERE / "memory_store.json", len(steps), "steps")
    for (n, k), o in zip(steps, outputs):
        print(f"  {n:28s}", (len(o) if isinstance(o, list) else o))


if __name__ == "__main__":
    run()
