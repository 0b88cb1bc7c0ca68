"""Importable name of the package whose sources live in ``youtu-rag_b200/``.

The directory name required by the project layout contains a hyphen and cannot be written in an
``import`` statement, so this stub extends its own ``__path__`` with that directory: every
submodule (``youtu_rag_b200.store`` …) is the file ``youtu-rag_b200/<name>.py`` loaded once.
"""

import os as _os

__path__.append(_os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "youtu-rag_b200"))

from .base import (  # noqa: E402
    BaseEmbedder,
    BaseReranker,
    BaseRetriever,
    BaseVectorStore,
    Chunk,
    Document,
    RetrievalResult,
)
from .config import RetrieverConfig, VectorStoreConfig  # noqa: E402
from .factory import VectorStoreFactory  # noqa: E402
from .memory_store import B200MemoryVectorStore, rank_memories, rank_skills  # noqa: E402
from .postprocess import ContextAssembler, dedup_by_file, merge_results  # noqa: E402
from .retriever import HybridRetriever, VectorRetriever  # noqa: E402
from .store import B200VectorStore  # noqa: E402

__all__ = [
    "B200VectorStore", "B200MemoryVectorStore", "rank_memories", "rank_skills", "VectorStoreFactory", "VectorRetriever", "HybridRetriever", "ContextAssembler", "dedup_by_file", "merge_results", "VectorStoreConfig",
    "RetrieverConfig", "BaseVectorStore", "BaseRetriever", "BaseEmbedder", "BaseReranker", "Chunk", "Document",
    "RetrievalResult",
]
