"""Config objects of the retrieval path, mirroring utu/rag/config.py.

`VectorStoreConfig` (utu/rag/config.py:52-65) with the backend Literal widened to include
"b200" — the one-line change INTEGRATION.md asks of the reference — and `RetrieverConfig`
(:42-49).  B200-specific knobs ride in the reference's otherwise unused `index_params` dict
(utu/rag/config.py:65): {"storage_dtype": "bf16"|"f32", "device": int, "reserve_rows": int,
"include_embeddings": bool}.
"""

from __future__ import annotations

from typing import Any, Literal

from pydantic import BaseModel, Field


class RetrieverConfig(BaseModel):
    top_k: int = Field(default=5, ge=1)
    similarity_threshold: float = Field(default=0.7, ge=0.0, le=1.0)
    enable_reranking: bool = False
    reranker_model: str | None = None
    reranker_top_k: int = Field(default=3, ge=1, le=50)


class VectorStoreConfig(BaseModel):
    backend: Literal["chroma", "b200"] = "b200"
    collection_name: str = "knowledge_base"
    persist_directory: str = "./data/vector_store"
    host: str | None = None
    port: int | None = None
    api_key: str | None = None
    distance_metric: Literal["cosine", "euclidean", "dot"] = "cosine"
    index_type: str | None = None
    index_params: dict[str, Any] = Field(default_factory=dict)
