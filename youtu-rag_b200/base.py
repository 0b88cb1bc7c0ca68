"""Types and abstract interfaces of the retrieval path, mirroring utu/rag/base.py.

When the reference package is importable (the backend is installed next to Youtu-RAG), its own
classes are re-exported so `isinstance(store, utu.rag.base.BaseVectorStore)` holds; otherwise the
same shapes are declared here: `Chunk` (utu/rag/base.py:26-39), `RetrievalResult` (:42-51),
`Document` (:12-23), `BaseEmbedder` (:113-124), `BaseReranker` (:127-147), `BaseRetriever`
(:171-184) and the drop-in boundary `BaseVectorStore` (:187-232).
"""

from __future__ import annotations

from abc import ABC, abstractmethod
from dataclasses import dataclass
from typing import Any

try:  # pragma: no cover - only when running inside the reference tree
    from utu.rag.base import (  # type: ignore
        BaseEmbedder, BaseReranker, BaseRetriever, BaseVectorStore, Chunk, Document, RetrievalResult)
    USING_REFERENCE_TYPES = True
except Exception:  # noqa: BLE001 - `import utu` asserts env vars and pulls optional deps
    USING_REFERENCE_TYPES = False

    @dataclass
    class Document:
        id: str
        content: str
        metadata: dict[str, Any] | None = None
        embedding: list[float] | None = None

    @dataclass
    class Chunk:
        id: str
        document_id: str
        content: str
        chunk_index: int
        metadata: dict[str, Any] | None = None
        embedding: list[float] | None = None

    @dataclass
    class RetrievalResult:
        chunk: Chunk
        score: float
        rank: int | None = None

    class BaseEmbedder(ABC):
        @abstractmethod
        async def embed_texts(self, texts: list[str]) -> list[list[float]]: ...

        @abstractmethod
        async def embed_query(self, query: str) -> list[float]: ...

    class BaseReranker(ABC):
        @abstractmethod
        async def rerank(self, query: str, results: list[RetrievalResult],
                         top_k: int | None = None) -> list[RetrievalResult]: ...

    class BaseRetriever(ABC):
        @abstractmethod
        async def retrieve(self, query: str, top_k: int = 5, **kwargs) -> list[RetrievalResult]: ...

        @abstractmethod
        async def batch_retrieve(self, queries: list[str], top_k: int = 5, **kwargs) -> list[list[RetrievalResult]]: ...

    class BaseVectorStore(ABC):
        @abstractmethod
        async def add_chunks(self, chunks: list[Chunk]) -> None: ...

        @abstractmethod
        async def search(self, query_embedding: list[float], top_k: int = 5,
                         filters: dict[str, Any] | None = None) -> list[tuple[Chunk, float]]: ...

        @abstractmethod
        async def delete(self, chunk_ids: list[str]) -> None: ...

        @abstractmethod
        async def delete_by_document_id(self, document_id: str) -> int: ...

        @abstractmethod
        async def get_by_id(self, chunk_id: str) -> Chunk | None: ...

        @abstractmethod
        async def count(self) -> int: ...

        @abstractmethod
        async def clear(self) -> None: ...
