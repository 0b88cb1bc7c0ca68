"""Generate tests/golden/postprocess.json: the reference's result post-processing and the
index-type filter composition of file-level search (§8 a8).

`KBSearchToolkit.kb_file_search` (utu/rag/rag_tools/kb_search_toolkit.py:446-680, embedding-only branch:
filters :526-535, per-file dedup :543-568, top-k shaping :646-657) and `MetaRetrievalToolkit.merge_retrieval_results`
(meta_retrieval_toolkit.py:620-653) are taken UNMODIFIED from the reference (source text via ast; the modules do
not import here: agents SDK, hydra) and run with a stub `self` whose retriever returns preset hits.

Usage: python tests/golden/make_postprocess_golden.py     (only where /root/reference exists)
"""

from __future__ import annotations

import asyncio
import json
import logging
import sys
from pathlib import Path
from typing import Any, Optional

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
sys.path.insert(0, str(HERE.parent.parent))

from make_filter_golden import REF, method_source  # noqa: E402
from youtu_rag_b200 import Chunk, RetrievalResult  # noqa: E402  (same fields as utu/rag/base.py:26-51)


def load(path, cls, name, extra=None):
    ns = {"Optional": Optional, "Any": Any, "logger": logging.getLogger("ref"), "json": json, "Chunk": Chunk,
          "RetrievalResult": RetrievalResult}
    ns.update(extra or {})
    exec(compile(method_source(path, cls, name), str(path), "exec", dont_inherit=True), ns)  # noqa: S102 - unmodified
    return ns[name]


def hits():
    out = []
    files = ["a.pdf", "b.pdf", "a.pdf", "c.pdf", None, "b.pdf", "d.pdf", None, "c.pdf", "e.pdf"]
    for i, f in enumerate(files):
        meta = {"index_type": "index_summary", "chunk_index": i, "summary": f"summary {i}", "authors": ["Li", "Wang"][i % 2], "year": 2020 + i}
        if f is not None:
            meta["source"] = f
        if i % 4 == 1:
            meta["_derived_files_etags"] = "x"
            del meta["summary"]
        out.append(RetrievalResult(chunk=Chunk(id=f"c{i}", document_id=f"doc{i % 3}", content=f"内容 {i}", chunk_index=i, metadata=meta),
                                   score=[0.91, 0.88, 0.93, 0.5, 0.77, 0.95, 0.3, 0.61, 0.52, 0.12][i], rank=i + 1))
    return out


def as_json(r):
    return {"id": r.chunk.id, "document_id": r.chunk.document_id, "content": r.chunk.content, "chunk_index": r.chunk.chunk_index,
            "metadata": r.chunk.metadata, "score": r.score, "rank": r.rank}


def run():
    kb_path = REF / "utu/rag/rag_tools/kb_search_toolkit.py"
    mr_path = REF / "utu/rag/rag_tools/meta_retrieval_toolkit.py"
    build = load(kb_path, "KBSearchToolkit", "_build_metadata_filters")
    file_search = load(kb_path, "KBSearchToolkit", "kb_file_search")
    merge = load(mr_path, "MetaRetrievalToolkit", "merge_retrieval_results")

    class FakeRetriever:
        def __init__(self, results, log):
            self.results, self.log = results, log

        async def retrieve(self, query, filters=None):
            self.log.append(filters)
            return list(self.results)

    class KBSelf:
        file_search_top_k, recall_multiplier, reranker_config = 4, 3, {}

        def __init__(self, results):
            self.results, self.seen_filters, self.seen_top_k = results, [], []

        def _build_metadata_filters(self, mf=None):
            return build(self, mf)

        async def _create_retriever(self, kb_id, top_k):
            self.seen_top_k.append(top_k)
            return FakeRetriever(self.results, self.seen_filters)

    out = {"hits": [as_json(r) for r in hits()], "file_search": [], "merge": []}
    for kw in ({"top_k": None, "metadata_filters": None, "include_summary": True},
               {"top_k": 2, "metadata_filters": {"authors": "Li"}, "include_summary": False},
               {"top_k": 10, "metadata_filters": {"year": {"$gte": 2022}, "source": {"$in": ["a.pdf", "c.pdf"]}}, "include_summary": True}):
        s = KBSelf(hits())
        text = asyncio.run(file_search(s, kb_id=1, query="q", auto_rerank=False, **kw))
        out["file_search"].append({"kwargs": kw, "retriever_top_k": s.seen_top_k[0], "filters": s.seen_filters[0], "output": json.loads(text)})

    class MRSelf:
        kb_id, query, top_k, filters = 1, "q", 5, {"source": "a.pdf"}

    h = hits()
    dup = RetrievalResult(chunk=h[2].chunk, score=0.4, rank=9)          # same chunk id found again by another search, lower score
    tie = RetrievalResult(chunk=Chunk(id="c_tie", document_id="doc9", content="tie", chunk_index=0, metadata={"source": "z.pdf"}), score=0.88, rank=2)
    for valid in (h, h[:5] + [dup] + h[5:] + [tie], [], [h[3], h[0], dup, h[1]]):
        s = MRSelf()
        s.valid_results = list(valid)
        text = asyncio.run(merge(s))
        out["merge"].append({"input": [as_json(r) for r in valid], "output": json.loads(text)})
    (HERE / "postprocess.json").write_text(json.dumps(out, ensure_ascii=False, separators=(",", ":")))
    print("wrote", HERE / "postprocess.json", [len(c["output"]["files"]) for c in out["file_search"]], [c["output"]["total_results"] for c in out["merge"]])


if __name__ == "__main__":
    run()
