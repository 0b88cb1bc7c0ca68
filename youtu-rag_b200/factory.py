"""VectorStoreFactory mirroring utu/rag/storage/base_storage.py:12-52 with the "b200" branch."""

from __future__ import annotations

from .base import BaseVectorStore
from .config import VectorStoreConfig


class VectorStoreFactory:
    @staticmethod
    def create(config: VectorStoreConfig) -> BaseVectorStore:
        backend = config.backend.lower()
        if backend == "b200":
            from .store import B200VectorStore

            return B200VectorStore(config=config)
        if backend == "chroma":
            try:
                from utu.rag.storage.implementations.chroma_store import ChromaVectorStore  # type: ignore
            except Exception as e:  # noqa: BLE001
                raise ImportError("backend 'chroma' needs the reference package (utu) and chromadb installed") from e
            return ChromaVectorStore(config=config)
        raise ValueError(f"Unsupported vector store backend: {backend}")

    @staticmethod
    def list_backends() -> list[str]:
        return ["b200", "chroma"]
