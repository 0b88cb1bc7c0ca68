"""B200MemoryVectorStore — agent-memory collections on the same engine.

Mirrors MemoryVectorStore (utu/rag/storage/implementations/memory_store.py:163-643): one collection
per user / memory type (`memory_<user>[_<type>]`, :209-223), always cosine (:237), upsert on add with
datetime / list / dict metadata serialised (:256-283), `search` that swallows engine errors into []
(:318-326) and JSON-decodes `tool_sequence` / `metadata` (:353-375), `search_memories` building the
session / type / importance / success-rate filter (:377-424), working memory by `where` (:426-477).
Every collection is a `B200VectorStore`, so the scoring, filtering and top-k run on the GPU.

`rank_memories` / `rank_skills` restate the toolkit's re-scoring of search hits
(utu/tools/memory_toolkit.py:906-931 and :1008-1034) for callers that route through this store.
"""

from __future__ import annotations

import json
import logging
from datetime import datetime
from typing import Any

from pathlib import Path

from .base import BaseVectorStore, Chunk
from .config import VectorStoreConfig
from .persist import CollectionDir
from .store import B200VectorStore

logger = logging.getLogger(__name__)

_JSON_FIELDS = ("tool_sequence", "metadata")


class B200MemoryVectorStore(BaseVectorStore):
    def __init__(self, config: VectorStoreConfig | None = None, persist_directory: str | None = None):
        self.config = config
        self._persist_directory = persist_directory or (config.persist_directory if config else "./data/memory")
        self._params = dict(config.index_params) if config else {}
        # the reference's memory store is a chromadb.PersistentClient (memory_store.py:183-199): memories survive
        # the process.  Persistence is therefore ON unless index_params says otherwise.
        self._params.setdefault("persist", True)
        self._collections: dict[str, B200VectorStore] = {}
        self._default_collection_name = config.collection_name if config else "agent_memory"

    # ------------------------------------------------------------------ collections
    def get_collection_name(self, user_id: str, memory_type: str | None = None) -> str:
        base = f"memory_{user_id}"
        return f"{base}_{memory_type}" if memory_type else base

    def get_or_create_collection(self, collection_name: str | None = None) -> B200VectorStore:
        name = collection_name or self._default_collection_name
        if name not in self._collections:
            cfg = VectorStoreConfig(backend="b200", collection_name=name, persist_directory=self._persist_directory,
                                    distance_metric="cosine", index_params=self._params)
            self._collections[name] = B200VectorStore(cfg)
        return self._collections[name]

    def list_collections(self) -> list[str]:
        """client.list_collections() (memory_store.py:617-624): what is on disk, opened in this process or not."""
        names = set(self._collections)
        root = Path(self._persist_directory)
        if self._params.get("persist") and root.is_dir():
            names.update(p.name[:-len(".b200")] for p in root.iterdir()
                         if p.is_dir() and p.name.endswith(".b200") and (p / "manifest.json").exists())
        return sorted(names)

    def delete_collection(self, collection_name: str | None = None) -> bool:
        """client.delete_collection(name) (memory_store.py:626-643): works on the persistent state, so a collection
        written by an earlier process is removed from disk even though this process never opened it."""
        name = collection_name or self._default_collection_name
        try:
            store = self._collections.pop(name, None)
            if store is not None:
                store.clear_sync()
                store.close()
            if self._params.get("persist"):
                CollectionDir(self._persist_directory, name).remove()
            return True
        except Exception as e:  # noqa: BLE001
            logger.warning("Failed to delete collection %s: %s", name, e)
            return False

    # ------------------------------------------------------------------ (de)serialisation
    @staticmethod
    def _serialize(meta: dict[str, Any] | None) -> dict[str, Any]:
        out: dict[str, Any] = {}
        for k, v in (meta or {}).items():
            if v is None:
                continue
            if isinstance(v, datetime):
                out[k] = v.isoformat()
            elif isinstance(v, (list, dict)):
                out[k] = json.dumps(v, ensure_ascii=False)
            else:
                out[k] = v
        return out

    @staticmethod
    def _deserialize_metadata(metadata: dict[str, Any]) -> dict[str, Any]:
        out = {}
        for k, v in metadata.items():
            if k in _JSON_FIELDS and isinstance(v, str):
                try:
                    out[k] = json.loads(v)
                except json.JSONDecodeError:
                    out[k] = v
            else:
                out[k] = v
        return out

    def _parse(self, chunk: Chunk) -> Chunk:
        meta = self._deserialize_metadata(chunk.metadata or {})
        return Chunk(id=chunk.id, document_id=meta.get("document_id", ""), content=chunk.content,
                     chunk_index=meta.get("chunk_index", 0), metadata=meta, embedding=chunk.embedding)

    # ------------------------------------------------------------------ BaseVectorStore (+ collection_name)
    async def add_chunks(self, chunks: list[Chunk], collection_name: str | None = None) -> None:
        if not chunks:
            return
        store = self.get_or_create_collection(collection_name)
        await store.upsert_chunks([Chunk(id=c.id, document_id=c.document_id, content=c.content, chunk_index=c.chunk_index,
                                         metadata=self._serialize(c.metadata), embedding=c.embedding) for c in chunks])

    async def search(self, query_embedding: list[float], top_k: int = 5, filters: dict[str, Any] | None = None,
                     collection_name: str | None = None) -> list[tuple[Chunk, float]]:
        store = self.get_or_create_collection(collection_name)
        try:
            results = await store.search(query_embedding, top_k=top_k, filters=filters)
        except Exception as e:  # noqa: BLE001 - memory_store.py:324-326
            logger.warning("Memory search failed: %s", e)
            return []
        return [(self._parse(c), s) for c, s in results]

    async def search_memories(self, query_embedding: list[float], user_id: str, memory_type: str | None = None,
                              session_id: str | None = None, top_k: int = 10, min_importance: float = 0.0,
                              include_outdated: bool = False) -> list[tuple[Chunk, float]]:
        conditions = []
        if session_id:
            conditions.append({"session_id": {"$eq": session_id}})
        if memory_type:
            conditions.append({"memory_type": {"$eq": memory_type}})
        if min_importance > 0:
            conditions.append({"importance_score": {"$gte": min_importance}})
        if not include_outdated:
            conditions.append({"success_rate": {"$gte": 0.2}})
        filters = None
        if len(conditions) == 1:
            filters = conditions[0]
        elif len(conditions) > 1:
            filters = {"$and": conditions}
        return await self.search(query_embedding, top_k=top_k, filters=filters,
                                 collection_name=self.get_collection_name(user_id, memory_type))

    async def get_working_memory(self, user_id: str, session_id: str, max_turns: int = 10) -> list[Chunk]:
        try:
            store = self.get_or_create_collection(self.get_collection_name(user_id))
            chunks = await store.get_where({"$and": [{"session_id": {"$eq": session_id}},
                                                     {"memory_type": {"$eq": "working"}}]}, include_embeddings=True)
        except Exception as e:  # noqa: BLE001
            logger.warning("Failed to get working memory: %s", e)
            return []
        chunks = [self._parse(c) for c in chunks]
        chunks.sort(key=lambda x: x.metadata.get("created_at", ""))
        return chunks[-max_turns:]

    async def delete(self, chunk_ids: list[str], collection_name: str | None = None) -> None:
        if chunk_ids:
            await self.get_or_create_collection(collection_name).delete(chunk_ids)

    async def delete_by_document_id(self, document_id: str, collection_name: str | None = None) -> int:
        return await self.get_or_create_collection(collection_name).delete_by_document_id(document_id)

    async def delete_by_metadata(self, metadata_filter: dict[str, Any], collection_name: str | None = None) -> int:
        return await self.get_or_create_collection(collection_name).delete_by_metadata(metadata_filter)

    async def cleanup_outdated_memories(self, user_id: str, success_rate_threshold: float = 0.2) -> int:
        return await self.delete_by_metadata({"success_rate": {"$lt": success_rate_threshold}},
                                             collection_name=self.get_collection_name(user_id, "procedural"))

    async def get_by_id(self, chunk_id: str, collection_name: str | None = None) -> Chunk | None:
        c = await self.get_or_create_collection(collection_name).get_by_id(chunk_id)
        return self._parse(c) if c is not None else None

    async def count(self, collection_name: str | None = None) -> int:
        return await self.get_or_create_collection(collection_name).count()

    async def clear(self, collection_name: str | None = None) -> None:
        await self.get_or_create_collection(collection_name).clear()


# ---------------------------------------------------------------------- re-scoring (row a10)
def recency_score(created_at: datetime, now: datetime | None = None) -> float:
    """Exponential decay with a 24 h half-life (memory_toolkit.py:928-931)."""
    age_hours = ((now or datetime.now()) - created_at).total_seconds() / 3600
    return 0.5 ** (age_hours / 24)


def _created(meta: dict[str, Any]) -> datetime:
    v = meta.get("created_at")
    if isinstance(v, datetime):
        return v
    try:
        return datetime.fromisoformat(v)
    except (TypeError, ValueError):
        return datetime.now()


def rank_memories(results: list[tuple[Chunk, float]], now: datetime | None = None) -> list[tuple[Chunk, float, float]]:
    """0.5·similarity + 0.3·importance + 0.2·recency, best first (memory_toolkit.py:916-925).
    Returns (chunk, similarity, relevance)."""
    out = []
    for chunk, sim in results:
        m = chunk.metadata or {}
        rel = 0.5 * sim + 0.3 * float(m.get("importance_score", 0.5)) + 0.2 * recency_score(_created(m), now)
        out.append((chunk, sim, rel))
    out.sort(key=lambda t: t[2], reverse=True)
    return out


def rank_skills(results: list[tuple[Chunk, float]], top_k: int, min_success_rate: float = 0.0,
                tool_filter: list[str] | None = None, now: datetime | None = None) -> list[tuple[Chunk, float, float]]:
    """Skills: drop low success rates / non-matching tool tags, then 0.4·similarity + 0.3·importance +
    0.2·success_rate + 0.1·recency (memory_toolkit.py:1008-1034); callers fetch 2·top_k hits first (:998)."""
    out = []
    for chunk, sim in results:
        m = chunk.metadata or {}
        # SkillMemory.success_rate = success_count / (success_count + failure_count), 1.0 when unused
        # (memory_toolkit.py:208-219); importance defaults to 0.7 for skills (:210)
        ok, bad = int(m.get("success_count", 1)), int(m.get("failure_count", 0))
        sr = float(m["success_rate"]) if "success_rate" in m else (ok / (ok + bad) if ok + bad > 0 else 1.0)
        if sr < min_success_rate:
            continue
        tags = m.get("tags", [])
        if isinstance(tags, str):
            try:
                tags = json.loads(tags)
            except json.JSONDecodeError:
                tags = [tags]
        if tool_filter and not any(t in tags for t in tool_filter):
            continue
        rel = 0.4 * sim + 0.3 * float(m.get("importance_score", 0.7)) + 0.2 * sr + 0.1 * recency_score(_created(m), now)
        out.append((chunk, sim, rel))
    out.sort(key=lambda t: t[2], reverse=True)
    return out[:top_k]
