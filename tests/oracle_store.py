"""Test-only BaseVectorStore built on the CPU oracle (never imported by the product)."""

from __future__ import annotations

import numpy as np

from oracle import exact_search as ox
from oracle import where_eval as ow
from youtu_rag_b200.base import BaseVectorStore, Chunk


class OracleStore(BaseVectorStore):
    def __init__(self, metric="cosine", dtype="f32"):
        self.metric, self.dtype = metric, dtype
        self.ids, self.docs, self.metas, self.rows = [], [], [], None

    async def add_chunks(self, chunks):
        emb = np.asarray([c.embedding for c in chunks], np.float32)
        rows = ox.prepare(emb, self.metric, self.dtype)
        self.rows = rows if self.rows is None else np.concatenate([self.rows, rows])
        for c in chunks:
            self.ids.append(c.id); self.docs.append(c.content)
            self.metas.append({"document_id": c.document_id, "chunk_index": c.chunk_index,
                               **{k: v for k, v in (c.metadata or {}).items() if v is not None}})

    async def search(self, query_embedding, top_k=5, filters=None):
        where = ow.normalize_filters(filters)
        mask = ow.eval_where(where, self.metas)
        q = ox.prepare(np.asarray(query_embedding, np.float32), self.metric, self.dtype)[0]
        ids, scores = ox.exact_topk(self.rows, q, top_k, self.metric, mask)
        return [(Chunk(id=self.ids[i], document_id=self.metas[i].get("document_id", ""), content=self.docs[i],
                       chunk_index=self.metas[i].get("chunk_index", 0), metadata=dict(self.metas[i]), embedding=None),
                 float(s)) for i, s in zip(ids.tolist(), scores.tolist())]

    async def delete(self, chunk_ids): raise NotImplementedError
    async def delete_by_document_id(self, document_id): raise NotImplementedError
    async def get_by_id(self, chunk_id): raise NotImplementedError
    async def count(self): return len(self.ids)
    async def clear(self): raise NotImplementedError
