"""Generate tests/golden/filter_producers.json: the `where` dicts the REFERENCE's toolkits emit.

The producer methods are taken UNMODIFIED (source text, via ast) from
  /root/reference/utu/rag/rag_tools/kb_search_toolkit.py       KBSearchToolkit._build_metadata_filters      :63-96
  /root/reference/utu/rag/rag_tools/meta_retrieval_toolkit.py  MetaRetrievalToolkit._build_metadata_filters :102-186
                                                               MetaRetrievalToolkit._build_time_range_filter :188-255
  /root/reference/utu/rag/storage/implementations/memory_store.py  MemoryVectorStore.search_memories         :377-424
  /root/reference/utu/rag/knowledge_retrieval/chroma_retrical_text2sql.py  CourseSearcher.search             :148-194
      (called once per text column with the same question by unified_schemalink_valuelink.py:289-303)
and executed here on representative arguments (the modules themselves do not import in this container: agents SDK,
hydra, chromadb are missing, so the functions are exec'd in a bare namespace; `search_memories` is run with a stub
`self` whose `search` records the filters it is handed).  The fixture pins the dicts; their evaluation over the
sample metadata is recorded from the golden fake engine's own matcher (make_golden.chroma_where) (Chroma itself is not installable: "unpinned").

Usage: python tests/golden/make_filter_golden.py     (only where /root/reference exists)
"""

from __future__ import annotations

import ast
import asyncio
import json
import logging
import sys
import textwrap
from pathlib import Path
from typing import Any, Optional

import numpy as np

REF = Path("/root/reference")
HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))

sys.path.insert(0, str(HERE))
from make_golden import chroma_where  # noqa: E402  (the fake engine's own matcher: independent of oracle/where_eval.py)


def method_source(path: Path, cls: str, name: str) -> str:
    src = path.read_text()
    for node in ast.walk(ast.parse(src)):
        if isinstance(node, ast.ClassDef) and node.name == cls:
            for f in node.body:
                if isinstance(f, (ast.FunctionDef, ast.AsyncFunctionDef)) and f.name == name:
                    return textwrap.dedent("".join(src.splitlines(keepends=True)[f.lineno - 1:f.end_lineno]))
    raise KeyError(f"{cls}.{name} not found in {path}")


def load(path: Path, cls: str, name: str):
    from typing import Dict, List

    ns = {"Optional": Optional, "Any": Any, "List": List, "Dict": Dict, "logger": logging.getLogger("ref"),
          "Chunk": object, "MemoryType": str}   # names that only appear in signatures
    exec(compile(method_source(path, cls, name), str(path), "exec", dont_inherit=True), ns)  # noqa: S102 - unmodified
    return ns[name]


def sample_metadata(n=240, seed=3):
    """Rows shaped like the reference's chunk metadata (processors.py:393,454,628; metadata_extractor.py:179-188)
    and memory records (memory_store.py:255-283)."""
    rng = np.random.default_rng(seed)
    metas = []
    for i in range(n):
        m = {"document_id": f"doc{i % 7}", "chunk_index": int(i % 11), "source": f"file{i % 5}.pdf",
             "index_type": ["index_content", "index_summary", "index_element"][i % 3], "char_length": int(rng.integers(50, 900))}
        if i % 4:
            lo = 1_735_689_600 + int(rng.integers(0, 300)) * 86_400           # 2025-01-01 + days
            m["创建时间_min_stamp"], m["创建时间_max_stamp"] = lo, lo + int(rng.integers(0, 40)) * 86_400
        if i % 3 == 0:
            m["author"], m["year"] = ["张三", "John"][i % 2], int(2018 + i % 8)
        if i % 2:
            m.update(type=["column_value", "table_schema"][i % 4 == 3], table_name=f"t{i % 3}", column_name=f"col{(i // 7) % 4}")
        if i % 5 == 0:
            m.update(session_id=f"s{i % 2}", memory_type=["episodic", "procedural"][(i // 10) % 2],
                     importance_score=float(rng.integers(0, 11)) / 10, success_rate=float(rng.integers(0, 11)) / 10)
        metas.append(m)
    return metas


def run():
    kb = load(REF / "utu/rag/rag_tools/kb_search_toolkit.py", "KBSearchToolkit", "_build_metadata_filters")
    meta_f = load(REF / "utu/rag/rag_tools/meta_retrieval_toolkit.py", "MetaRetrievalToolkit", "_build_metadata_filters")
    time_f = load(REF / "utu/rag/rag_tools/meta_retrieval_toolkit.py", "MetaRetrievalToolkit", "_build_time_range_filter")
    mem_f = load(REF / "utu/rag/storage/implementations/memory_store.py", "MemoryVectorStore", "search_memories")

    class MetaSelf:  # the methods only use self to reach each other
        _build_time_range_filter = lambda self, v: time_f(self, v)  # noqa: E731

    class MemSelf:
        def __init__(self):
            self.seen = None

        def get_collection_name(self, user_id, memory_type):
            return f"{user_id}_{memory_type}"

        async def search(self, query_embedding, top_k, filters, collection_name):
            self.seen = filters
            return []

    d0, d1, d2, d3 = 1_738_368_000, 1_743_379_200, 1_751_328_000, 1_759_190_400   # 2025-02-01, 03-31, 07-01, 09-30 (UTC)
    cases = []

    def add(producer, args, where):
        cases.append({"producer": producer, "args": args, "where": where})

    for mf in (None, {}, {"source": "file1.pdf"}, {"source": {"$in": ["file1.pdf", "file3.pdf"]}},
               {"source": "file2.pdf", "index_type": "index_summary"}, {"char_length": {"$gte": 400}, "document_id": {"$ne": "doc3"}},
               {"index_type": {"$nin": ["index_element"]}, "chunk_index": {"$lt": 4}, "source": {"$in": ["file0.pdf", "file4.pdf"]}}):
        add("kb_search._build_metadata_filters", {"metadata_filters": mf}, kb(None, mf))
    for fids, mf in ((None, None), (["file1.pdf"], None), (["file1.pdf", "file2.pdf", "file4.pdf"], None),
                     (None, {"author": "张三", "year": {"$gte": 2020}}),
                     (["file0.pdf"], {"time_ranges": [{"field": "创建时间", "ranges": [d0, d1]}]}),
                     (None, {"time_ranges": [{"field": "创建时间", "ranges": [d0, d1]}, {"field": "创建时间", "ranges": [d2, d3]}]}),
                     (["file1.pdf", "file3.pdf"], {"author": "John", "time_ranges": [{"field": "创建时间", "ranges": [d2, d3]},
                                                                                   {"field": "创建时间", "ranges": []},
                                                                                   {"field": "创建时间", "ranges": [d0]}]}),
                     (None, {"time_ranges": {"field": "创建时间", "ranges": [[d0, d1]]}})):   # dict form: rejected by the producer
        add("meta_retrieval._build_metadata_filters", {"file_ids": fids, "metadata_filters": mf}, meta_f(MetaSelf(), fids, mf))
    for kw in ({"session_id": "s1", "memory_type": "episodic", "min_importance": 0.0, "include_outdated": True},
               {"session_id": None, "memory_type": "procedural", "min_importance": 0.5, "include_outdated": False},
               {"session_id": "s0", "memory_type": None, "min_importance": 0.3, "include_outdated": False},
               {"session_id": None, "memory_type": None, "min_importance": 0.0, "include_outdated": True}):
        s = MemSelf()
        asyncio.run(mem_f(s, user_id="u", query_embedding=[0.0], top_k=5, **kw))
        add("memory_store.search_memories", kw, s.seen)

    course = load(REF / "utu/rag/knowledge_retrieval/chroma_retrical_text2sql.py", "CourseSearcher", "search")

    class CourseSelf:
        def __init__(self):
            self._embedding_cache, self.seen = {}, None
            outer = self

            class Emb:
                async def embed_query(self, q):
                    return [1.0, 0.0]

            class VS:
                async def search(self, query_embedding, top_k, filters):
                    outer.seen = filters
                    return []

            self.embedder, self.vector_store = Emb(), VS()

    for conds in (None, [{"type": "table_schema"}],
                  [{"type": "column_value"}, {"table_name": "t1"}, {"column_name": "col2"}],      # the value-link loop's shape
                  [{"type": "column_value"}, {"table_name": "t0"}, {"column_name": "col0"}],
                  [{"type": "column_value"}, {"table_name": {"$in": ["t0", "t2"]}}]):
        s = CourseSelf()
        asyncio.run(course(s, query="q", top_k=3, filter_conditions=conds))
        add("text2sql.CourseSearcher.search", {"filter_conditions": conds}, s.seen)

    metas = sample_metadata()
    for c in cases:
        c["rows"] = None if c["where"] is None else np.flatnonzero(chroma_where(c["where"], metas)).tolist()
    out = {"metadatas": metas, "cases": cases}
    (HERE / "filter_producers.json").write_text(json.dumps(out, ensure_ascii=False, separators=(",", ":")))
    print("wrote", HERE / "filter_producers.json", len(cases), "cases;",
          "non-trivial:", sum(1 for c in cases if c["rows"] and 0 < len(c["rows"]) < len(metas)))


if __name__ == "__main__":
    run()
