"""B200VectorStore — the drop-in `BaseVectorStore` whose search runs on hand-written sm_100a CUDA.

Interface parity target: ChromaVectorStore (utu/rag/storage/implementations/chroma_store.py), the
store every agent path constructs today.  Method by method:

  add_chunks            chroma_store.py:64-88    flattened metadata rows, embeddings → device (K5)
  search                chroma_store.py:90-148   filters normalisation :104-116, exact top-k instead of
                                                 HNSW, score = 1 - distance :132-135, Chunk shaping :137-146
  delete                chroma_store.py:150-160  tombstones
  delete_by_document_id chroma_store.py:162-183
  delete_by_metadata    chroma_store.py:185-222  (multi-key dict → $and :196-203, errors → 0 :220-222)
  get_by_id             chroma_store.py:224-247
  count / clear         chroma_store.py:249-272

The Python side keeps only what has no arithmetic: row → (chunk id, text, metadata dict) tables and
the columnar metadata mirror; all scoring, filtering and selection happen behind include/yrb200.h.
There is no CPU search path: without libyrb200.so or a B200 the constructor/first add raises.
"""

from __future__ import annotations

import logging
from typing import Any

import numpy as np

from . import native
from .base import BaseVectorStore, Chunk
from .config import VectorStoreConfig
from .metadata import MetadataTable
from .persist import CollectionDir, validate_collection_name
from .where import compile_where, normalize_filters

logger = logging.getLogger(__name__)

_METRIC_NAMES = ("cosine", "euclidean", "dot")


def _threshold_f32(t: float | None) -> float | None:
    """The smallest float32 >= t: scores are float32, so `score >= t` (t a Python double, base_retriever.py:71) and
    `score >= this` select exactly the same hits."""
    if t is None:
        return None
    t32 = np.float32(t)
    if float(t32) < float(t):
        t32 = np.nextafter(t32, np.float32(np.inf))
    return float(t32)


class B200VectorStore(BaseVectorStore):
    """Exact dense retrieval over a device-resident chunk-embedding matrix."""

    supports_score_threshold = True   # search / search_batch take score_threshold (applied inside the scan)

    def __init__(self, config: VectorStoreConfig):
        self.config = config
        validate_collection_name(config.collection_name)   # Chroma's create_collection rejects the same names
        p = dict(getattr(config, "index_params", None) or {})
        self._dtype = p.get("storage_dtype", "bf16")
        if self._dtype not in native.DTYPES:
            raise ValueError(f"index_params.storage_dtype must be one of {sorted(native.DTYPES)}, got {self._dtype!r}")
        self._device = int(p.get("device", 0))
        # index_params.devices = [0, 1, …]: the collection is row-sharded over these GPUs of the box inside this
        # process (native.ShardedIndex); absent or a single entry → one GPU
        devs = p.get("devices")
        self._devices = [int(d) for d in devs] if devs else None
        self._block_rows = int(p.get("shard_block_rows", 0))
        self._reserve = int(p.get("reserve_rows", 0))
        self._include_embeddings = bool(p.get("include_embeddings", False))
        # unknown metric names fall back to cosine like chroma_store.py:52
        metric = config.distance_metric if config.distance_metric in _METRIC_NAMES else "cosine"
        self._metric = metric
        native.lib()  # fail now, loudly, if the CUDA extension was not built
        self._index: native.Index | None = None
        self._ids: list[str | None] = []
        self._documents: list[str | None] = []
        self._metadatas: list[dict[str, Any] | None] = []
        self._row_of: dict[str, int] = {}
        self._meta = MetadataTable()
        self._deleted: set[int] = set()   # tombstoned row numbers (rows are stable across reloads)
        # persistence is opt-in (index_params.persist): the reference's Chroma client always persists
        # (chroma_store.py:41-44); here an in-memory collection is the default so tests leave no files
        self._dir = CollectionDir(config.persist_directory, config.collection_name) if p.get("persist") else None
        if self._dir is not None and self._dir.exists():
            self._load()
        logger.info("Initialized B200 vector store (collection %s, metric %s, storage %s, device %d)",
                    config.collection_name, metric, self._dtype, self._device)

    # ------------------------------------------------------------------ helpers
    def _ensure_index(self, dim: int) -> native.Index:
        if self._index is None:
            if self._devices and len(self._devices) > 1:
                self._index = native.ShardedIndex(dim, self._metric, self._dtype, self._devices, self._reserve, self._block_rows)
            else:
                dev = self._devices[0] if self._devices else self._device
                self._index = native.Index(dim, self._metric, self._dtype, dev, self._reserve)
        elif self._index.dim != dim:
            raise ValueError(f"Embedding dimension {dim} does not match collection dimensionality {self._index.dim}")
        return self._index

    def _make_chunk(self, row: int, embedding: list[float] | None) -> Chunk:
        meta = dict(self._metadatas[row] or {})
        return Chunk(id=self._ids[row], document_id=meta.get("document_id", ""), content=self._documents[row],
                     chunk_index=meta.get("chunk_index", 0), metadata=meta, embedding=embedding)

    def _compile(self, filters: dict[str, Any] | None):
        where = normalize_filters(filters)
        compiled, cols = compile_where(where, self._meta)
        if cols:
            self._meta.sync(self._index, cols)
        return compiled

    def _rows_matching(self, where: dict[str, Any]) -> list[int]:
        """collection.get(where=…) stand-in: rows (live) passing a where clause, via K4."""
        if self._index is None or self._index.rows == 0:
            return []
        compiled, cols = compile_where(where, self._meta)
        if cols:
            self._meta.sync(self._index, cols)
        words, n = self._index.where_mask(compiled)
        if n == 0:
            return []
        bits = np.unpackbits(words.view(np.uint8), bitorder="little")[: self._index.rows]
        return np.flatnonzero(bits).tolist()

    def _load(self) -> None:
        m = self._dir.manifest()
        if (m["metric"], m["dtype"]) != (self._metric, self._dtype):
            raise ValueError(f"collection on disk is {m['metric']}/{m['dtype']}, config asks {self._metric}/{self._dtype}")
        index = self._ensure_index(m["dim"])
        superseded: list[int] = []
        for rows, sqnorm, recs in self._dir.segments():
            base = index.rows
            index.append_raw(rows, sqnorm)
            for i, r in enumerate(recs):
                if r["id"] in self._row_of:       # upserted: the later segment wins even if the tombstone list
                    superseded.append(self._row_of[r["id"]])   # was not rewritten before a crash
                self._row_of[r["id"]] = base + i
            self._ids.extend(r["id"] for r in recs)
            self._documents.extend(r["document"] for r in recs)
            self._metadatas.extend(r["metadata"] for r in recs)
            self._meta.append([r["metadata"] for r in recs])
        gone = sorted({r for r in self._dir.deleted() if 0 <= r < index.rows} | set(superseded))
        if gone:
            index.set_live(gone, False)
            for r in gone:
                if self._row_of.get(self._ids[r]) == r:   # not re-added by a later segment
                    del self._row_of[self._ids[r]]
                self._ids[r] = self._documents[r] = self._metadatas[r] = None
            self._deleted.update(gone)
        logger.info("Loaded %d chunks (%d deleted) from %s", len(self._row_of), len(gone), self._dir.path)

    # ------------------------------------------------------------------ BaseVectorStore
    async def add_chunks(self, chunks: list[Chunk]) -> None:
        if not chunks:
            return
        seen: set[str] = set()
        for c in chunks:
            if c.id in seen:
                raise ValueError(f"Expected IDs to be unique, found duplicates of: {c.id}")
            seen.add(c.id)
        fresh = [c for c in chunks if c.id not in self._row_of]
        if len(fresh) != len(chunks):
            # Chroma's add() leaves existing ids untouched
            logger.warning("Add of %d existing chunk ids ignored", len(chunks) - len(fresh))
        self._append(fresh)

    def _append(self, fresh: list[Chunk]) -> list[int]:
        """Append chunks as new rows (ids may already exist: the caller tombstones the rows they replace).
        Everything that can be rejected is checked before the first mutation; the on-disk segment is written
        before the host tables change, and a failed write rolls the device append back, so memory and disk
        never diverge.  Returns the rows the ids pointed at before (upsert)."""
        if not fresh:
            return []
        for c in fresh:
            if c.embedding is None:
                raise ValueError(f"Chunk {c.id} has no embedding")
        emb = np.asarray([c.embedding for c in fresh], dtype=np.float32)
        if emb.ndim != 2:
            raise ValueError("Expected every embedding to have the same dimension")
        if not np.isfinite(emb).all():
            raise ValueError("Embeddings contain NaN or infinite values")
        metas = [MetadataTable.validate({"document_id": c.document_id, "chunk_index": c.chunk_index,
                                         **{k: v for k, v in (c.metadata or {}).items() if v is not None}})
                 for c in fresh]
        index = self._ensure_index(emb.shape[1])
        base = index.rows
        assert base == len(self._ids) == self._meta.rows
        index.append(emb)
        if self._dir is not None:
            try:
                if not self._dir.exists():
                    self._dir.create(index.dim, self._metric, self._dtype, index.info()["ld"])
                raw, sq = index.read_raw(base, len(fresh))   # exactly what the device holds
                self._dir.append_segment(raw, sq, [c.id for c in fresh], [c.content for c in fresh], metas)
            except BaseException:
                index.truncate(base)   # the rows never became visible to a search
                raise
        replaced = [self._row_of[c.id] for c in fresh if c.id in self._row_of]
        for i, c in enumerate(fresh):
            self._row_of[c.id] = base + i
        self._ids.extend(c.id for c in fresh)
        self._documents.extend(c.content for c in fresh)
        self._metadatas.extend(metas)
        self._meta.append(metas)
        logger.info("Added %d chunks to B200 index", len(fresh))
        return replaced

    def _tombstone(self, rows: list[int]) -> None:
        if not rows:
            return
        self._index.set_live(rows, False)
        for r in rows:
            self._ids[r] = self._documents[r] = self._metadatas[r] = None
        if self._dir is not None:
            self._deleted.update(rows)
            self._dir.write_deleted(list(self._deleted))

    async def upsert_chunks(self, chunks: list[Chunk]) -> None:
        """collection.upsert (memory_store.py:278-283): an existing id is replaced — the new embedding / text /
        metadata are appended (and persisted) first, then the old row is tombstoned, so a failure in between
        leaves the record present; a reload treats an id seen again in a later segment the same way."""
        if not chunks:
            return
        last = {c.id: c for c in chunks}            # within one call the last occurrence of an id wins
        self._tombstone(self._append(list(last.values())))

    async def get_where(self, where: dict[str, Any] | None, include_embeddings: bool = False) -> list[Chunk]:
        """collection.get(where=…) (memory_store.py:442-451): every live chunk passing the filter, in row order."""
        if self._index is None:
            return []
        if where is None:
            rows = sorted(self._row_of.values())
        else:
            rows = self._rows_matching(where)
        embs = self._index.read_rows(rows).tolist() if (include_embeddings and rows) else [None] * len(rows)
        return [self._make_chunk(r, e) for r, e in zip(rows, embs)]

    async def search(self, query_embedding: list[float], top_k: int = 5,
                     filters: dict[str, Any] | None = None, score_threshold: float | None = None) -> list[tuple[Chunk, float]]:
        out = await self.search_batch([query_embedding], top_k=top_k, filters=filters, score_threshold=score_threshold)
        return out[0]

    async def search_batch(self, query_embeddings, top_k: int = 5,
                           filters: dict[str, Any] | list[dict[str, Any] | None] | None = None,
                           score_threshold: float | None = None) -> list[list[tuple[Chunk, float]]]:
        """Q queries in one device pass (the batched entry point VectorRetriever.batch_retrieve uses;
        the reference loops single searches, base_retriever.py:95-99).  `filters` is one filter for all
        queries or a list with one filter (or None) per query — the shape of the text2sql value-linking
        loop (unified_schemalink_valuelink.py:289-303: same vector, one metadata filter per column).
        `score_threshold` (an extension the reference's store does not have): only hits with score >= threshold are
        returned — the retriever's rule (base_retriever.py:71) applied inside the scan instead of after it."""
        q = np.asarray(query_embeddings, dtype=np.float32)
        if q.ndim == 1:
            q = q[None, :]
        if top_k < 1:
            raise ValueError(f"Expected n_results to be a positive integer, got {top_k}")
        if not np.isfinite(q).all():
            raise ValueError("Query embedding contains NaN or infinite values")
        per_query = isinstance(filters, (list, tuple))
        if per_query and len(filters) != q.shape[0]:
            raise ValueError(f"expected {q.shape[0]} filters (one per query), got {len(filters)}")
        if self._index is None or self._index.counts()[1] == 0:
            for f in (filters if per_query else [filters]):  # still validate like Chroma would
                compile_where(normalize_filters(f), self._meta)
            return [[] for _ in range(q.shape[0])]
        if q.shape[1] != self._index.dim:
            raise ValueError(
                f"Embedding dimension {q.shape[1]} does not match collection dimensionality {self._index.dim}")
        if per_query:
            cache: dict[str, Any] = {}
            compiled_list = []
            for f in filters:
                key = repr(f)
                if key not in cache:
                    cache[key] = self._compile(f)
                compiled_list.append(cache[key])  # identical filters share one program (evaluated once)
            ids, scores, counts = self._index.search(q, int(top_k), wheres=compiled_list, min_score=_threshold_f32(score_threshold))
        else:
            compiled = self._compile(filters)
            ids, scores, counts = self._index.search(q, int(top_k), where=compiled, min_score=_threshold_f32(score_threshold))
        results = []
        for j in range(q.shape[0]):
            n = int(counts[j])
            rows = ids[j, :n]
            embs = self._index.read_rows(rows).tolist() if (self._include_embeddings and n) else [None] * n
            results.append([(self._make_chunk(int(rows[i]), embs[i]), float(scores[j, i])) for i in range(n)])
        return results

    async def delete(self, chunk_ids: list[str]) -> None:
        if not chunk_ids:
            return
        rows = [self._row_of.pop(cid) for cid in dict.fromkeys(chunk_ids) if cid in self._row_of]
        self._tombstone(rows)
        logger.info("Deleted %d chunks from B200 index", len(rows))

    async def delete_by_document_id(self, document_id: str) -> int:
        rows = self._rows_matching({"document_id": document_id})
        if not rows:
            logger.info("No chunks found for document_id: %s", document_id)
            return 0
        await self.delete([self._ids[r] for r in rows])
        return len(rows)

    async def delete_by_metadata(self, metadata_filter: dict[str, Any]) -> int:
        if len(metadata_filter) > 1:
            where = {"$and": [{k: v} for k, v in metadata_filter.items()]}
        else:
            where = metadata_filter
        try:
            rows = self._rows_matching(where)
            if not rows:
                return 0
            await self.delete([self._ids[r] for r in rows])
            return len(rows)
        except Exception as e:  # noqa: BLE001 - chroma_store.py:220-222 logs and returns 0
            logger.error("Failed to delete by metadata %s: %s", metadata_filter, e)
            return 0

    async def get_by_id(self, chunk_id: str) -> Chunk | None:
        row = self._row_of.get(chunk_id)
        if row is None:
            return None
        return self._make_chunk(row, self._index.read_rows([row])[0].tolist())

    async def count(self) -> int:
        return len(self._row_of)

    async def clear(self) -> None:
        self.clear_sync()

    def clear_sync(self) -> None:
        """Body of clear(); callable from synchronous code that already runs inside an event loop
        (MemoryVectorStore.delete_collection is a plain method, memory_store.py:626-643)."""
        if self._index is not None:
            self._index.clear()
        self._ids, self._documents, self._metadatas, self._row_of = [], [], [], {}
        self._meta.clear()
        self._deleted.clear()
        if self._dir is not None:
            self._dir.remove()
        logger.info("Cleared B200 collection: %s", self.config.collection_name)

    # ------------------------------------------------------------------ extras
    def delete_collection(self) -> None:
        """Delete the collection permanently, its on-disk directory included (chroma_store.py:331-398; probed with
        hasattr and called synchronously by KnowledgeCleanupManager, cleanup_manager.py:651-652)."""
        self.clear_sync()
        self.close()

    @staticmethod
    def cleanup_orphaned_directories(persist_directory: str) -> dict:
        """Remove collection directories a crashed writer left without a manifest (the counterpart of
        chroma_store.py:274-329, which removes segment directories Chroma's index no longer lists).
        Returns {"deleted_count": int, "deleted_dirs": list} like the reference."""
        import shutil
        from pathlib import Path

        root = Path(persist_directory)
        deleted: list[str] = []
        if root.exists():
            for item in sorted(root.iterdir()):
                if item.is_dir() and item.name.endswith(".b200") and not (item / "manifest.json").exists():
                    shutil.rmtree(item)
                    deleted.append(item.name)
        return {"deleted_count": len(deleted), "deleted_dirs": deleted}

    def close(self) -> None:
        if self._index is not None:
            self._index.close()
            self._index = None

    @property
    def index(self) -> native.Index | None:
        return self._index
