"""Host-side post-processing of search hits — the cheap steps that follow the path.

Mirrors, for callers that switch to this backend:
  * ContextAssembler (utu/rag/knowledge_retrieval/context_assembler.py:11-153): hits → markdown / plain /
    JSON context under a character budget;
  * the per-file de-duplication of kb_file_search (utu/rag/rag_tools/kb_search_toolkit.py:544-568);
  * merge_retrieval_results' de-duplication by chunk id + sort (meta_retrieval_toolkit.py:620-653).
They stay on the host on purpose: a search now takes ~0.3 ms, and these touch k ≤ ~100 small objects.
"""

from __future__ import annotations

import json
from typing import Any, Callable

from .base import RetrievalResult

_HIDDEN_METADATA = ("chunk_index", "total_chunks")
_FILE_ENTRY_EXCLUDED = ("index_type", "chunk_index", "_derived_files_etags")


class ContextAssembler:
    def __init__(self, max_context_length: int = 4000):
        self.max_context_length = max_context_length

    def assemble(self, results: list[RetrievalResult], include_metadata: bool = True, format_style: str = "markdown") -> str:
        if not results:
            return ""
        if format_style == "markdown":
            return "\n\n---\n\n".join(self._take(results, lambda i, r: "\n\n".join(
                [f"## Context {i} (Relevance: {r.score:.2f})"]
                + ([f"**Metadata:** {self._format_metadata(r.chunk.metadata)}"] if include_metadata and r.chunk.metadata else [])
                + [r.chunk.content])))
        if format_style == "plain":
            return "\n\n".join(self._take(results, lambda i, r: "\n".join(
                [f"Context {i}:"]
                + ([f"Metadata: {self._format_metadata(r.chunk.metadata)}"] if include_metadata and r.chunk.metadata else [])
                + [r.chunk.content])))
        if format_style == "json":
            items: list[dict[str, Any]] = []

            def render(_i: int, r: RetrievalResult) -> str:
                item = {"content": r.chunk.content, "score": r.score, "rank": r.rank}
                if include_metadata and r.chunk.metadata:
                    item["metadata"] = r.chunk.metadata
                items.append(item)
                return json.dumps(item, ensure_ascii=False)

            kept = self._take(results, render)
            return json.dumps(items[: len(kept)], ensure_ascii=False, indent=2)
        raise ValueError(f"Unknown format style: {format_style}")

    def _take(self, results: list[RetrievalResult], render: Callable[[int, RetrievalResult], str]) -> list[str]:
        """Sections in order until the next one would exceed the character budget (separators are not counted)."""
        out, used = [], 0
        for i, r in enumerate(results, 1):
            section = render(i, r)
            if used + len(section) > self.max_context_length:
                break
            out.append(section)
            used += len(section)
        return out

    @staticmethod
    def _format_metadata(metadata: dict[str, Any]) -> str:
        return ", ".join(f"{k}={v}" for k, v in metadata.items() if k not in _HIDDEN_METADATA)


def dedup_by_file(results: list[RetrievalResult], include_summary: bool = False) -> list[dict[str, Any]]:
    """First (best-ranked) hit per file, files ordered by relevance — kb_search_toolkit.py:544-568."""
    files: dict[str, dict[str, Any]] = {}
    for r in results:
        meta = r.chunk.metadata or {}
        name = meta.get("source", r.chunk.document_id)
        if name in files:
            continue
        entry = {"file_name": name, "relevance_score": r.score, "chunk_id": r.chunk.id, "content": r.chunk.content}
        if include_summary:
            entry["summary"] = meta.get("summary", "")
        entry["metadata"] = {k: v for k, v in meta.items() if k not in _FILE_ENTRY_EXCLUDED}
        files[name] = entry
    return sorted(files.values(), key=lambda e: e["relevance_score"], reverse=True)


def merge_results(results: list[RetrievalResult]) -> list[RetrievalResult]:
    """Several searches' hits → one list: one entry per chunk id (a later hit replaces an earlier one but
    keeps its place), then a stable sort by score, best first — meta_retrieval_toolkit.py:623-626."""
    by_id: dict[str, RetrievalResult] = {}
    for r in results:
        by_id[r.chunk.id] = r
    return sorted(by_id.values(), key=lambda r: r.score, reverse=True)
