"""VectorRetriever / HybridRetriever mirroring utu/rag/knowledge_retrieval/base_retriever.py.

Same call contract as the reference's retriever (top_k override, `filters` and
`similarity_threshold` kwargs, threshold rule, 1-based rank, optional rerank over 2×top_k, final
slice — base_retriever.py:42-80).  The one behavioural change is inside `batch_retrieve`
(base_retriever.py:82-99): queries are still embedded one by one with `embed_query` (the
embedders use a different prompt for queries than for documents, service_embedder.py:153-166),
but the Q vector searches become ONE `search_batch` call when the store offers it, so a batch
reaches the tcgen05 kernel instead of Q sequential scans.  Results are identical to the loop.
"""

from __future__ import annotations

import logging

from .base import BaseEmbedder, BaseReranker, BaseRetriever, BaseVectorStore, RetrievalResult
from .config import RetrieverConfig

logger = logging.getLogger(__name__)


class VectorRetriever(BaseRetriever):
    def __init__(self, vector_store: BaseVectorStore, embedder: BaseEmbedder, config: RetrieverConfig | None = None,
                 reranker: BaseReranker | None = None):
        self.vector_store = vector_store
        self.embedder = embedder
        self.config = config or RetrieverConfig()
        # the reference builds its reranker from RerankerFactory (an HTTP client, out of scope);
        # here it is injected, and enable_reranking without one is an error rather than a silent skip
        self.reranker = reranker
        if self.config.enable_reranking and reranker is None:
            raise ValueError("enable_reranking=True needs a reranker instance (RerankerFactory lives in utu.rag.rerankers)")

    def _post(self, results, similarity_threshold: float) -> list[RetrievalResult]:
        out = []
        for i, (chunk, score) in enumerate(results):
            if similarity_threshold <= 0.0 or score >= similarity_threshold:
                out.append(RetrievalResult(chunk=chunk, score=score, rank=i + 1))
        return out

    def _device_threshold(self, threshold: float) -> dict:
        """f4: a store that can apply the similarity threshold inside its scan gets it (the hits it drops are exactly the
        ones `_post` would drop: the list is sorted, the threshold cuts its tail, ranks of the kept hits do not move)."""
        if threshold > 0.0 and getattr(self.vector_store, "supports_score_threshold", False):
            return {"score_threshold": threshold}
        return {}

    async def retrieve(self, query: str, top_k: int | None = None, **kwargs) -> list[RetrievalResult]:
        top_k = top_k or self.config.top_k
        filters = kwargs.get("filters")
        threshold = kwargs.get("similarity_threshold", self.config.similarity_threshold)
        query_embedding = await self.embedder.embed_query(query)
        results = await self.vector_store.search(
            query_embedding=query_embedding, top_k=top_k * 2 if self.reranker else top_k, filters=filters, **self._device_threshold(threshold))
        retrieval_results = self._post(results, threshold)
        if self.reranker and retrieval_results:
            retrieval_results = await self.reranker.rerank(query=query, results=retrieval_results, top_k=top_k)
        return retrieval_results[:top_k]

    async def batch_retrieve(self, queries: list[str], top_k: int | None = None, **kwargs) -> list[list[RetrievalResult]]:
        if not queries:
            return []
        if not hasattr(self.vector_store, "search_batch"):
            return [await self.retrieve(query=q, top_k=top_k, **kwargs) for q in queries]
        top_k = top_k or self.config.top_k
        filters = kwargs.get("filters")
        threshold = kwargs.get("similarity_threshold", self.config.similarity_threshold)
        embeddings = [await self.embedder.embed_query(q) for q in queries]
        per_query = await self.vector_store.search_batch(
            embeddings, top_k=top_k * 2 if self.reranker else top_k, filters=filters, **self._device_threshold(threshold))
        out = []
        for query, results in zip(queries, per_query):
            rr = self._post(results, threshold)
            if self.reranker and rr:
                rr = await self.reranker.rerank(query=query, results=rr, top_k=top_k)
            out.append(rr[:top_k])
        return out


class HybridRetriever(BaseRetriever):
    """Delegates to VectorRetriever exactly like the reference stub (base_retriever.py:102-154)."""

    def __init__(self, vector_store: BaseVectorStore, embedder: BaseEmbedder, config: RetrieverConfig | None = None,
                 reranker: BaseReranker | None = None):
        self.vector_retriever = VectorRetriever(vector_store, embedder, config, reranker)
        self.config = config or RetrieverConfig()

    async def retrieve(self, query: str, top_k: int | None = None, **kwargs) -> list[RetrievalResult]:
        return await self.vector_retriever.retrieve(query=query, top_k=top_k, **kwargs)

    async def batch_retrieve(self, queries: list[str], top_k: int | None = None, **kwargs) -> list[list[RetrievalResult]]:
        return await self.vector_retriever.batch_retrieve(queries=queries, top_k=top_k, **kwargs)
