"""Generate tests/golden/memory_rescoring.json: the reference's memory / skill re-scoring.

`VectorMemoryToolkit.search_memories` (utu/tools/memory_toolkit.py:865-925), `search_skills` (:955-1035),
`_calculate_recency_score` (:927-930) and the pydantic models they use (`ToolCall`, `SkillMemory`, `MemoryNode`,
`MemorySearchResult`, `SkillSearchResult`, :166-466) are taken UNMODIFIED from the reference (source text via ast) and
executed over a fake collection that returns preset hits.  The module itself does not import here (agents SDK,
chromadb), hence the extraction.  `datetime.now()` inside the toolkit methods is pinned to NOW so the recency term
is reproducible.  The fixture records, per call: the `where` and `n_results` handed to `collection.query`, and the
ordered (id, distance, relevance_score) the reference returns.

Usage: python tests/golden/make_memory_golden.py     (only where /root/reference exists)
"""

from __future__ import annotations

import ast
import asyncio
import json
import logging
import sys
import textwrap
import uuid
from datetime import datetime, timedelta
from pathlib import Path
from typing import Any, Literal

import numpy as np
from pydantic import BaseModel, Field

REF = Path("/root/reference/utu/tools/memory_toolkit.py")
HERE = Path(__file__).resolve().parent
NOW = datetime(2025, 6, 1, 12, 0, 0)


class PinnedDatetime(datetime):
    @classmethod
    def now(cls, tz=None):
        return NOW


def sources():
    src = REF.read_text()
    lines = src.splitlines(keepends=True)
    tree = ast.parse(src)
    seg = lambda n: textwrap.dedent("".join(lines[n.lineno - 1:n.end_lineno]))  # noqa: E731
    models, methods = [], {}
    for node in tree.body:
        if isinstance(node, ast.Assign) and any(isinstance(t, ast.Name) and t.id == "MemoryType" for t in node.targets):
            models.append(seg(node))
        if isinstance(node, ast.ClassDef) and node.name in ("ToolCall", "SkillMemory", "MemoryNode", "MemorySearchResult", "SkillSearchResult"):
            models.append(seg(node))
        if isinstance(node, ast.ClassDef) and node.name == "VectorMemoryToolkit":
            for f in node.body:
                if isinstance(f, (ast.FunctionDef, ast.AsyncFunctionDef)) and f.name in ("search_memories", "search_skills", "_calculate_recency_score"):
                    methods[f.name] = seg(f)
    return models, methods


class FakeCollection:
    def __init__(self, hits):
        self.hits, self.calls = hits, []

    def query(self, query_embeddings, n_results, where=None, include=()):
        self.calls.append({"where": where, "n_results": n_results})
        h = self.hits[:n_results]
        return {"ids": [[x["id"] for x in h]], "documents": [[x["document"] for x in h]],
                "metadatas": [[dict(x["metadata"]) for x in h]], "distances": [[x["distance"] for x in h]]}


def run():
    models, methods = sources()
    base = {"BaseModel": BaseModel, "Field": Field, "Any": Any, "Literal": Literal, "json": json, "uuid": uuid,
            "logger": logging.getLogger("ref")}
    ns_models = dict(base, datetime=datetime)
    for m in models:
        exec(compile(m, str(REF), "exec", dont_inherit=True), ns_models)  # noqa: S102 - reference source, unmodified
    ns_methods = dict(ns_models, datetime=PinnedDatetime)
    for m in methods.values():
        exec(compile(m, str(REF), "exec", dont_inherit=True), ns_methods)  # noqa: S102

    rng = np.random.default_rng(11)

    def mem_hit(i):
        meta = {"user_id": "u", "session_id": f"s{i % 2}", "memory_type": "episodic", "importance_score": float(rng.integers(0, 11)) / 10,
                "created_at": (NOW - timedelta(hours=float(rng.integers(0, 200)))).isoformat(), "success_rate": float(rng.integers(2, 11)) / 10,
                "avg_latency": 12.5, "tool_sequence": "[]"}
        if i % 3 == 0:
            del meta["importance_score"]                    # model default 0.5
        return {"id": f"m{i}", "document": f"memory {i}", "metadata": meta, "distance": 0.05 + 0.04 * i + float(rng.random()) * 0.02}

    def skill_hit(i):
        meta = {"skill_name": f"skill{i}", "tool_sequence": json.dumps([{"tool": "search", "args": {"q": "x"}}]),
                "tags": json.dumps([["search", "sql"], ["python"], ["search"], []][i % 4]), "trigger_patterns": "[]", "example_qa": "{}",
                "success_count": int(rng.integers(0, 6)), "failure_count": int(rng.integers(0, 6)),
                "importance_score": float(rng.integers(3, 11)) / 10,
                "created_at": (NOW - timedelta(hours=float(rng.integers(0, 100)))).isoformat()}
        if i % 5 == 0:
            del meta["importance_score"]                    # model default 0.7
        if i == 7:
            meta["success_count"], meta["failure_count"] = 0, 0   # success_rate 1.0 when unused
        return {"id": f"k{i}", "document": f"Skill: does thing {i}\nsteps", "metadata": meta, "distance": 0.1 + 0.03 * i}

    class Self:
        default_user_id = "u"

        def __init__(self, hits):
            self.coll = FakeCollection(hits)

        async def _get_embedding(self, q):
            return [0.0, 1.0]

        def _get_collection_name(self, user_id, memory_type):
            return f"{user_id}_{memory_type}"

        def _get_skill_collection_name(self, user_id):
            return f"{user_id}_skills"

        def _get_or_create_collection(self, name):
            return self.coll

        def _calculate_recency_score(self, created_at):
            return ns_methods["_calculate_recency_score"](self, created_at)

    out = {"now": NOW.isoformat(), "memories": [], "skills": []}
    mem_hits = [mem_hit(i) for i in range(12)]
    for kw in ({"top_k": 10}, {"top_k": 5, "memory_type": "episodic", "session_id": "s1"},
               {"top_k": 12, "min_importance": 0.4, "include_outdated": True}, {"top_k": 3, "memory_type": "procedural", "include_outdated": True}):
        s = Self(mem_hits)
        res = asyncio.run(ns_methods["search_memories"](s, "q", **kw))
        out["memories"].append({"kwargs": kw, "query": s.coll.calls[0], "hits": mem_hits[:kw["top_k"]],
                                "results": [{"id": r.memory.id, "distance": r.score, "relevance": r.relevance_score} for r in res]})
    skill_hits = [skill_hit(i) for i in range(14)]
    for kw in ({"top_k": 5}, {"top_k": 4, "tool_filter": "search"}, {"top_k": 7, "tool_filter": ["python", "sql"], "min_success_rate": 0.0},
               {"top_k": 3, "min_success_rate": 0.6}):
        s = Self(skill_hits)
        res = asyncio.run(ns_methods["search_skills"](s, "q", **kw))
        out["skills"].append({"kwargs": kw, "query": s.coll.calls[0], "hits": skill_hits[:2 * kw["top_k"]],
                              "results": [{"id": r.skill.id, "distance": r.score, "relevance": r.relevance_score} for r in res]})
    (HERE / "memory_rescoring.json").write_text(json.dumps(out, ensure_ascii=False, separators=(",", ":")))
    print("wrote", HERE / "memory_rescoring.json", [len(c["results"]) for c in out["memories"]], [len(c["results"]) for c in out["skills"]])


if __name__ == "__main__":
    run()
