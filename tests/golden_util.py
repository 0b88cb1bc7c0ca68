import json
from pathlib import Path

import numpy as np

from youtu_rag_b200.base import Chunk

GOLDEN = json.loads((Path(__file__).parent / "golden" / "reference_glue.json").read_text())


class GoldenReranker:
    """Deterministic stand-in for the reference's HTTP rerankers (utu/rag/rerankers): orders hits by a hash of
    (query, chunk id), gives them scores 0.9, 0.8, … and re-numbers the ranks.  Used on BOTH sides — by
    make_golden.py under the reference's VectorRetriever and by the tests under this repo's."""

    async def rerank(self, query, results, top_k=None):
        import hashlib

        def key(r):
            return hashlib.md5(f"{query}|{r.chunk.id}".encode()).hexdigest()

        out = sorted(results, key=key)
        out = out[:top_k] if top_k else out
        return [type(r)(chunk=r.chunk, score=round(0.9 - 0.1 * i, 6), rank=i + 1) for i, r in enumerate(out)]


def golden_chunks():
    x = np.asarray(GOLDEN["corpus"]["embeddings"], np.float32)
    metas = GOLDEN["corpus"]["metadatas"]
    return [Chunk(id=f"doc{i // 8}_chunk_{i % 8}", document_id=f"doc{i // 8}", content=f"text {i}", chunk_index=i % 8,
                  metadata={**metas[i], "none_field": None}, embedding=x[i].tolist()) for i in range(len(metas))]


def compare_with_golden(results, want, tol, tie_eps=5e-7):
    """results: list[(Chunk, score)]; want: golden result dicts.  Ids must agree except inside groups of
    scores tied within tie_eps (the reference's engine breaks such ties by insertion order)."""
    assert len(results) == len(want)
    got_ids = [c.id for c, _ in results]
    got_scores = np.array([s for _, s in results], np.float64)
    want_scores = np.array([w["score"] for w in want], np.float64)
    np.testing.assert_allclose(got_scores, want_scores, rtol=tol, atol=tol)
    for i, (g, w) in enumerate(zip(got_ids, want)):
        if g != w["id"]:
            # accept only if the golden item sits in a tie group that also contains ours
            group = [x["id"] for x in want if abs(x["score"] - w["score"]) <= max(tie_eps, tol)]
            assert g in group or i == len(want) - 1, f"pos {i}: {g} vs {w['id']}"
    for (c, _), w in zip(results, want):
        if "metadata" in w and c.id == w["id"]:
            assert c.document_id == w["document_id"] and c.chunk_index == w["chunk_index"]
            assert c.content == w["content"] and c.metadata == w["metadata"]


# ----------------------------------------------------------------------------- memory-store scenario (a6)
def _to_chunk(Chunk, s):
    from datetime import datetime

    meta = dict(s["metadata"])
    meta["created_at"] = datetime.fromisoformat(meta["created_at"])          # exercised: datetime -> isoformat on add
    return Chunk(id=s["id"], document_id=s["document_id"], content=s["content"], chunk_index=s["chunk_index"],
                 metadata=meta, embedding=s["embedding"])


def _hit_json(pairs):
    return [{"id": c.id, "document_id": c.document_id, "chunk_index": c.chunk_index, "content": c.content,
             "metadata": c.metadata, "score": float(s)} for c, s in pairs]


def _chunk_json(c):
    return None if c is None else {"id": c.id, "document_id": c.document_id, "chunk_index": c.chunk_index, "content": c.content,
                                   "metadata": c.metadata, "embedding": [float(v) for v in c.embedding] if c.embedding is not None else None}


async def replay_memory_scenario(store, Chunk, specs, steps):
    """Runs the scripted steps of tests/golden/memory_store.json on any store with MemoryVectorStore's interface
    (the reference's, in the generator; B200MemoryVectorStore, in the tests); returns the JSON-able outputs."""
    out = []
    for name, kw in steps:
        kw = dict(kw)
        if name == "add_chunks":
            await store.add_chunks([_to_chunk(Chunk, specs[i]) for i in kw["chunks"]], collection_name=kw["collection_name"])
            out.append(None)
        elif name == "upsert":
            s = json.loads(json.dumps(specs[kw["chunk"]]))
            s["embedding"] = specs[kw["new_embedding_from"]]["embedding"]
            s["metadata"]["importance_score"] = kw["importance_score"]
            await store.add_chunks([_to_chunk(Chunk, s)], collection_name=kw["collection_name"])
            out.append(None)
        elif name in ("search", "search_memories"):
            q = specs[kw.pop("q")]["embedding"]
            out.append(_hit_json(await getattr(store, name)(query_embedding=q, **kw)))
        elif name == "get_working_memory":
            out.append([_chunk_json(c) for c in await store.get_working_memory(**kw)])
        elif name == "get_by_id":
            out.append(_chunk_json(await store.get_by_id(**kw)))
        elif name == "get_collection_name":
            out.append(store.get_collection_name(**kw))
        elif name == "delete_collection":
            out.append(store.delete_collection(**kw))
        else:
            out.append(await getattr(store, name)(**kw))
    return out


def check_memory_outputs(golden, got, tol, emb_atol):
    """Step-by-step comparison with the reference's outputs.  Search hits: ids in order (exact ties may swap),
    scores within tol, document_id / chunk_index / content / parsed metadata equal.  Returned embeddings: the
    reference hands back the vector as given, this backend the stored (unit-norm) one — compared after normalising."""
    assert len(got) == len(golden["outputs"])
    for (name, kw), want, have in zip(golden["steps"], golden["outputs"], got):
        where = f"{name} {kw}"
        if name in ("search", "search_memories"):
            class _C:  # adapter for compare_with_golden
                def __init__(self, h):
                    self.id, self.document_id, self.chunk_index, self.content, self.metadata = (
                        h["id"], h["document_id"], h["chunk_index"], h["content"], h["metadata"])
            assert len(have) == len(want), where
            compare_with_golden([(_C(h), h["score"]) for h in have], want, tol=tol)
        elif name in ("get_working_memory", "get_by_id"):
            w_list, h_list = (want, have) if name == "get_working_memory" else ([want], [have])
            assert len(w_list) == len(h_list), where
            for w, h in zip(w_list, h_list):
                if w is None or h is None:
                    assert w is None and h is None, where
                    continue
                assert {k: h[k] for k in ("id", "document_id", "chunk_index", "content", "metadata")} == \
                       {k: w[k] for k in ("id", "document_id", "chunk_index", "content", "metadata")}, where
                e = np.asarray(w["embedding"], np.float64)
                np.testing.assert_allclose(np.asarray(h["embedding"], np.float64), e / np.linalg.norm(e), atol=emb_atol, err_msg=where)
        else:
            assert have == want, where
