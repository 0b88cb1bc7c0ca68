"""Generate tests/golden/*.json by running the REFERENCE's own Python on this container.

Runs only where /root/reference exists (never on the GPU box).  What it pins, and what it cannot:

* The reference's vector-store glue is executed UNMODIFIED, loaded from
  /root/reference/utu/rag/{base,config}.py, storage/implementations/{chroma,faiss}_store.py and
  knowledge_retrieval/base_retriever.py: filters normalisation (chroma_store.py:104-116), the
  `where` handed to the engine, `score = 1 - distance` (:132-135), Chunk shaping (:137-146),
  FAISS's normalise/score/post-filter logic (faiss_store.py:143-199) and the retriever's
  threshold / rank / slice rules (base_retriever.py:53-80).
* The engines underneath (chromadb 1.3.4 HNSW, faiss-cpu 1.12.0) are NOT installable here, so
  `chromadb` and `faiss` are replaced by minimal exact fakes defined in this file (with their OWN `where` matcher,
  `chroma_where` below — not the oracle's: the fixtures are independent of oracle/where_eval.py).  The numbers in
  the fixtures therefore pin the reference's Python semantics, not the engines' arithmetic:
  parity stays "unpinned" for the latter.

Usage: python tests/golden/make_golden.py   (rewrites the JSON fixtures next to this file)
"""

from __future__ import annotations

import asyncio
import importlib
import importlib.util
import json
import sys
import types
from pathlib import Path

import numpy as np

REF = Path("/root/reference")
HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))



# ----------------------------------------------------------------------------- the fake engine's own `where` matcher
# Written here from Chroma's documented `where` grammar, deliberately sharing NO code with oracle/where_eval.py
# (fixtures produced through the oracle's matcher cannot disagree with the oracle).  One metadata dict at
# a time, recursive, types compared by Python type name.  The decisions it restates are DESIGN.md §5's.
_W_LOGICAL = ("$and", "$or")
_W_COMPARE = ("$gt", "$gte", "$lt", "$lte")
_W_ALL = _W_COMPARE + ("$eq", "$ne", "$in", "$nin")


def _w_kind(v):
    return "bool" if isinstance(v, bool) else "int" if isinstance(v, int) else "float" if isinstance(v, float) else \
        "str" if isinstance(v, str) else None


def _w_check(where):
    """Raise ValueError where chromadb's validate_where would."""
    if not isinstance(where, dict) or len(where) != 1:
        raise ValueError(f"Expected where to have exactly one operator, got {where}")
    (key, val), = where.items()
    if not isinstance(key, str):
        raise ValueError(f"Expected where key to be a str, got {key}")
    if key in _W_LOGICAL:
        if not isinstance(val, list) or len(val) < 2:
            raise ValueError(f"Expected where value for {key} to be a list with at least two where expressions, got {val}")
        for child in val:
            _w_check(child)
        return
    if key.startswith("$"):
        raise ValueError(f"Expected where to have a field name or one of {_W_LOGICAL}, got {key}")
    if not isinstance(val, dict):
        if _w_kind(val) is None:
            raise ValueError(f"Expected where value to be a str, int, float, bool or operator expression, got {val}")
        return
    if len(val) != 1:
        raise ValueError(f"Expected operator expression to have exactly one operator, got {val}")
    (op, operand), = val.items()
    if op not in _W_ALL:
        raise ValueError(f"Expected where operator to be one of {_W_ALL}, got {op}")
    if op in _W_COMPARE:
        if _w_kind(operand) not in ("int", "float"):
            raise ValueError(f"Expected operand value to be an int or a float for operator {op}, got {operand}")
    elif op in ("$in", "$nin"):
        if not isinstance(operand, list) or not operand:
            raise ValueError(f"Expected where operand value to be a non-empty list, got {operand}")
        kinds = {_w_kind(x) for x in operand}
        if None in kinds or len(kinds) != 1:
            raise ValueError(f"Expected where operand value to be a list of one of str, int, float or bool, got {operand}")
    elif _w_kind(operand) is None:
        raise ValueError(f"Expected where operand value to be a str, int, float or bool, got {operand}")


def _w_row(where, meta):
    (key, val), = where.items()
    if key == "$and":
        return all(_w_row(c, meta) for c in val)
    if key == "$or":
        return any(_w_row(c, meta) for c in val)
    op, operand = next(iter(val.items())) if isinstance(val, dict) else ("$eq", val)
    probe = operand[0] if op in ("$in", "$nin") else operand
    stored = meta.get(key)
    typed = key in meta and _w_kind(stored) == _w_kind(probe)      # a value of another type is invisible to the operand
    if op == "$eq":
        return typed and stored == operand
    if op == "$ne":
        return not (typed and stored == operand)
    if op == "$in":
        return typed and stored in operand
    if op == "$nin":
        return not (typed and stored in operand)
    if not typed:
        return False
    return {"$gt": stored > operand, "$gte": stored >= operand, "$lt": stored < operand, "$lte": stored <= operand}[op]


def chroma_where(where, metas):
    """bool[len(metas)]: rows a chromadb `where` selects (None selects all)."""
    if where is None:
        return np.ones(len(metas), bool)
    _w_check(where)
    return np.fromiter((_w_row(where, m) for m in metas), bool, count=len(metas))


# ----------------------------------------------------------------------------- fake engines
class _FakeCollection:
    """Exact stand-in for a chromadb Collection: fp32 storage, distances as hnswlib defines them
    (cosine: 1 - cos, l2: squared L2, ip: 1 - dot), ascending distance, ties by insertion order."""

    def __init__(self, name, metadata):
        self.name, self.space = name, (metadata or {}).get("hnsw:space", "l2")
        self.ids, self.emb, self.docs, self.metas = [], [], [], []
        self.last_where = "unset"

    def add(self, ids, embeddings, documents, metadatas):
        for i, e, d, m in zip(ids, embeddings, documents, metadatas):
            if i in self.ids:
                continue
            self.ids.append(i); self.emb.append(np.asarray(e, np.float32)); self.docs.append(d); self.metas.append(dict(m))

    def upsert(self, ids, embeddings, documents, metadatas):
        for i, e, d, m in zip(ids, embeddings, documents, metadatas):
            if i in self.ids:
                j = self.ids.index(i)
                self.emb[j], self.docs[j], self.metas[j] = np.asarray(e, np.float32), d, dict(m)
            else:
                self.add([i], [e], [d], [m])

    def count(self):
        return len(self.ids)

    def _dist(self, q):
        x = np.stack(self.emb).astype(np.float64)
        q = np.asarray(q, np.float64)
        if self.space == "cosine":
            return 1.0 - (x @ q) / (np.linalg.norm(x, axis=1) * np.linalg.norm(q))
        if self.space == "ip":
            return 1.0 - x @ q
        return ((x - q[None]) ** 2).sum(1)

    def query(self, query_embeddings, n_results, where=None, include=()):
        self.last_where = where
        out = {"ids": [], "documents": [], "metadatas": [], "embeddings": [], "distances": []}
        for q in query_embeddings:
            keep = chroma_where(where, self.metas) if self.ids else np.zeros(0, bool)
            idx = np.flatnonzero(keep)
            if idx.size:
                d = self._dist(q)[idx]
                o = idx[np.lexsort((idx, d))][:n_results]
                dd = self._dist(q)[o]
            else:
                o, dd = np.zeros(0, np.int64), np.zeros(0)
            out["ids"].append([self.ids[i] for i in o])
            out["documents"].append([self.docs[i] for i in o])
            out["metadatas"].append([dict(self.metas[i]) for i in o])
            out["embeddings"].append([self.emb[i].tolist() for i in o])
            out["distances"].append([float(np.float32(x)) for x in dd])
        return out

    def get(self, ids=None, where=None, include=()):
        if ids is not None:
            sel = [self.ids.index(i) for i in ids if i in self.ids]
        else:
            sel = np.flatnonzero(chroma_where(where, self.metas)).tolist() if self.ids else []
        return {"ids": [self.ids[i] for i in sel], "documents": [self.docs[i] for i in sel],
                "metadatas": [dict(self.metas[i]) for i in sel], "embeddings": [self.emb[i].tolist() for i in sel]}

    def delete(self, ids=None, where=None):
        sel = set(self.get(ids=ids, where=where)["ids"])
        keep = [i for i, cid in enumerate(self.ids) if cid not in sel]
        self.ids = [self.ids[i] for i in keep]; self.emb = [self.emb[i] for i in keep]
        self.docs = [self.docs[i] for i in keep]; self.metas = [self.metas[i] for i in keep]


class _FakeClient:
    def __init__(self, path=None, settings=None):
        self.cols = {}

    def get_or_create_collection(self, name, metadata=None):
        return self.cols.setdefault(name, _FakeCollection(name, metadata))

    def list_collections(self):
        return list(self.cols.values())

    def delete_collection(self, name):
        self.cols.pop(name, None)


class _FakeFlat:
    def __init__(self, d, ip):
        self.d, self.ip, self.x = d, ip, np.zeros((0, d), np.float32)

    @property
    def ntotal(self):
        return self.x.shape[0]

    def add(self, x):
        self.x = np.concatenate([self.x, np.asarray(x, np.float32)])

    def search(self, q, k):
        q = np.asarray(q, np.float32)
        if self.ip:
            s = (self.x.astype(np.float64) @ q[0].astype(np.float64))
            o = np.lexsort((np.arange(s.size), -s))[:k]
        else:
            s = ((self.x.astype(np.float64) - q[0].astype(np.float64)[None]) ** 2).sum(1)
            o = np.lexsort((np.arange(s.size), s))[:k]
        return s[o][None].astype(np.float32), o[None].astype(np.int64)


def _install_fakes():
    chroma = types.ModuleType("chromadb")
    chroma.PersistentClient = _FakeClient
    chroma.Client = _FakeClient
    chroma.Collection = _FakeCollection
    cfg = types.ModuleType("chromadb.config")
    cfg.Settings = lambda **kw: kw
    chroma.config = cfg
    sys.modules["chromadb"], sys.modules["chromadb.config"] = chroma, cfg
    faiss = types.ModuleType("faiss")
    faiss.IndexFlatIP = lambda d: _FakeFlat(d, True)
    faiss.IndexFlatL2 = lambda d: _FakeFlat(d, False)

    def normalize_L2(x):
        n = np.sqrt((x.astype(np.float64) ** 2).sum(1))
        n[n == 0] = 1.0
        x /= n[:, None].astype(np.float32)

    faiss.normalize_L2 = normalize_L2
    faiss.write_index = lambda index, path: Path(path).write_bytes(b"fake")
    faiss.read_index = lambda path: (_ for _ in ()).throw(RuntimeError("not supported by the fake"))
    sys.modules["faiss"] = faiss


def _load_reference():
    """Import the reference's path modules without executing utu/__init__.py (env asserts, hydra…)."""
    def pkg(name, path):
        m = types.ModuleType(name)
        m.__path__ = [str(path)]
        sys.modules[name] = m
    pkg("utu", REF / "utu"); pkg("utu.db", REF / "utu/db"); pkg("utu.config", REF / "utu/config")
    pkg("utu.rag", REF / "utu/rag"); pkg("utu.rag.storage", REF / "utu/rag/storage")
    pkg("utu.rag.storage.implementations", REF / "utu/rag/storage/implementations")
    pkg("utu.rag.knowledge_retrieval", REF / "utu/rag/knowledge_retrieval")
    pkg("utu.rag.rerankers", REF / "utu/rag/rerankers")
    fac = types.ModuleType("utu.rag.rerankers.factory")
    from tests.golden_util import GoldenReranker   # only reached with enable_reranking=True (base_retriever.py:34-39)

    fac.RerankerFactory = type("RerankerFactory", (), {"create": staticmethod(lambda **kw: GoldenReranker())})
    sys.modules["utu.rag.rerankers.factory"] = fac
    mods = {}
    for name in ("utu.rag.base", "utu.rag.config", "utu.rag.storage.implementations.chroma_store",
                 "utu.rag.storage.implementations.faiss_store", "utu.rag.knowledge_retrieval.base_retriever"):
        mods[name.rsplit(".", 1)[1]] = importlib.import_module(name)
    return mods


# ----------------------------------------------------------------------------- corpus
def corpus(n=64, d=16, seed=7):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, d)).astype(np.float32)
    x[5] = x[3]           # exact duplicate → tie broken by insertion order / id
    x[40] = 2.5 * x[12]   # same direction, larger norm: ties under cosine, not under dot / l2
    metas = []
    for i in range(n):
        m = {"source": f"file{i % 4}.pdf", "index_type": ["index_content", "index_summary"][i % 2],
             "t_min_stamp": 1_700_000_000 + 1000 * i, "t_max_stamp": 1_700_000_000 + 1000 * i + 500,
             "importance_score": round(0.05 * (i % 20), 2)}
        if i % 3 == 0:
            m["year"] = 2020 + (i % 5)
        if i % 7 == 0:
            m["flag"] = bool(i % 2)
        if i % 5 == 0:
            m["mixed"] = i if i % 10 == 0 else float(i)
        metas.append(m)
    return x, metas


FILTERS = [
    None,
    {"source": "file1.pdf"},
    {"source": "file1.pdf", "index_type": "index_content"},          # multi-key: reference does NOT and them
    {"source": {"$in": ["file0.pdf", "file3.pdf"]}},
    {"$and": [{"source": {"$eq": "file2.pdf"}}, {"index_type": {"$eq": "index_content"}}]},
    {"$or": [{"$and": [{"t_min_stamp": {"$lte": 1_700_010_000}}, {"t_max_stamp": {"$gte": 1_700_005_000}}]},
             {"$and": [{"t_min_stamp": {"$lte": 1_700_050_000}}, {"t_max_stamp": {"$gte": 1_700_045_000}}]}]},
    {"importance_score": {"$gte": 0.5}},
    {"year": {"$ne": 2021}},
    {"year": {"$nin": [2020, 2022]}},
    {"flag": True},
    {"mixed": {"$gt": 10}},
    {"mixed": {"$gt": 10.0}},
    {"missing_field": "x"},
    {"$and": [{"source": "file1.pdf"}]},                               # Chroma rejects 1-element $and
    {"source": {"$regex": "file.*"}},                                  # not a where operator
    {"year": {"$gt": "2020"}},
]


def run():
    _install_fakes()
    ref = _load_reference()
    Chunk = ref["base"].Chunk
    x, metas = corpus()
    queries = np.random.default_rng(11).standard_normal((4, x.shape[1])).astype(np.float32)
    queries[1] = x[3]
    out = {"corpus": {"embeddings": x.tolist(), "metadatas": metas}, "queries": queries.tolist(), "chroma": [],
           "faiss": [], "retriever": []}

    def chunks():
        return [Chunk(id=f"doc{i // 8}_chunk_{i % 8}", document_id=f"doc{i // 8}", content=f"text {i}", chunk_index=i % 8,
                      metadata={**metas[i], "none_field": None}, embedding=x[i].tolist()) for i in range(len(metas))]

    import tempfile
    tmp_faiss = tempfile.mkdtemp(prefix="golden_faiss_")

    async def go():
        for metric in ("cosine", "dot", "euclidean"):
            cfg = ref["config"].VectorStoreConfig(collection_name=f"g_{metric}", persist_directory="/tmp/_golden",
                                                  distance_metric=metric)
            store = ref["chroma_store"].ChromaVectorStore(cfg)
            await store.add_chunks(chunks())
            for qi, q in enumerate(queries.tolist()):
                for fi, flt in enumerate(FILTERS):
                    for k in (1, 5, 70):
                        if k == 70 and fi not in (0, 3):
                            continue
                        rec = {"metric": metric, "query": qi, "filter": fi, "filters": flt, "top_k": k}
                        try:
                            res = await store.search(q, top_k=k, filters=flt)
                            rec["where_passed_to_engine"] = store.collection.last_where
                            full = (k == 1)  # Chunk shaping is pinned on the k=1 cases, ids+scores everywhere
                            rec["results"] = [
                                ({"id": c.id, "document_id": c.document_id, "chunk_index": c.chunk_index,
                                  "content": c.content, "metadata": c.metadata, "score": s} if full
                                 else {"id": c.id, "score": s}) for c, s in res]
                        except Exception as e:  # noqa: BLE001
                            rec["error"] = type(e).__name__
                        out["chroma"].append(rec)
            if metric == "cosine":
                # mutation semantics
                mut = {"count0": await store.count()}
                mut["deleted_doc2"] = await store.delete_by_document_id("doc2")
                mut["deleted_meta"] = await store.delete_by_metadata({"source": "file1.pdf", "index_type": "index_summary"})
                await store.delete(["doc0_chunk_0", "nope"])
                mut["count1"] = await store.count()
                g = await store.get_by_id("doc0_chunk_2")
                mut["get"] = {"id": g.id, "document_id": g.document_id, "chunk_index": g.chunk_index, "metadata": g.metadata}
                mut["get_missing"] = await store.get_by_id("doc0_chunk_0")
                res = await store.search(queries[0].tolist(), top_k=5)
                mut["search_after"] = [{"id": c.id, "score": s} for c, s in res]
                await store.clear()
                mut["count2"] = await store.count()
                out["chroma_mutations"] = mut

        for metric in ("cosine", "dot", "euclidean"):
            cfg = ref["config"].VectorStoreConfig(collection_name=f"f_{metric}", persist_directory=tmp_faiss,
                                                  distance_metric=metric)
            fs = ref["faiss_store"].FAISSVectorStore(cfg)
            await fs.clear()
            await fs.add_chunks(chunks())
            for qi, q in enumerate(queries.tolist()):
                for flt in (None, {"source": "file1.pdf"}):
                    res = await fs.search(q, top_k=5, filters=flt)
                    out["faiss"].append({"metric": metric, "query": qi, "filters": flt, "top_k": 5,
                                         "results": [{"id": c.id, "score": s} for c, s in res]})

        class Emb(ref["base"].BaseEmbedder):
            async def embed_texts(self, texts):
                return [queries[int(t)].tolist() for t in texts]

            async def embed_query(self, query):
                return queries[int(query)].tolist()

        cfg = ref["config"].VectorStoreConfig(collection_name="r", persist_directory="/tmp/_golden", distance_metric="cosine")
        store = ref["chroma_store"].ChromaVectorStore(cfg)
        await store.add_chunks(chunks())
        for thr in (0.0, 0.3, 0.7):
            rc = ref["config"].RetrieverConfig(top_k=4, similarity_threshold=thr)
            r = ref["base_retriever"].VectorRetriever(store, Emb(), rc)
            for kw in ({}, {"filters": {"source": "file1.pdf"}}, {"similarity_threshold": 0.1}):
                single = await r.retrieve("1", **kw)
                batch = await r.batch_retrieve(["0", "1", "2"], top_k=3, **kw)
                out["retriever"].append({
                    "config_threshold": thr, "kwargs": kw,
                    "single": [{"id": x.chunk.id, "score": x.score, "rank": x.rank} for x in single],
                    "batch": [[{"id": x.chunk.id, "score": x.score, "rank": x.rank} for x in b] for b in batch]})
        # with a reranker: search 2k, threshold on the 2k hits, rerank to k, slice (base_retriever.py:61-80)
        out["retriever_rerank"] = []
        for thr in (0.0, 0.2):
            rc = ref["config"].RetrieverConfig(top_k=4, similarity_threshold=thr, enable_reranking=True)
            r = ref["base_retriever"].VectorRetriever(store, Emb(), rc)
            for kw in ({}, {"filters": {"source": {"$in": ["file0.pdf", "file1.pdf"]}}}):
                single = await r.retrieve("2", **kw)
                batch = await r.batch_retrieve(["0", "1"], top_k=3, **kw)
                out["retriever_rerank"].append({
                    "config_threshold": thr, "kwargs": kw,
                    "single": [{"id": x.chunk.id, "score": x.score, "rank": x.rank} for x in single],
                    "batch": [[{"id": x.chunk.id, "score": x.score, "rank": x.rank} for x in b] for b in batch]})

    asyncio.run(go())
    # ContextAssembler (context_assembler.py) over retriever-shaped results, run unmodified
    pkg_name = "utu.rag.knowledge_retrieval.context_assembler"
    ca = importlib.import_module(pkg_name)
    RR, Ch = ref["base"].RetrievalResult, ref["base"].Chunk
    hits = [RR(chunk=Ch(id=f"c{i}", document_id=f"d{i % 3}", content=f"内容 text {i} " * (i + 1), chunk_index=i,
                        metadata=({"source": f"file{i % 2}.pdf", "chunk_index": i, "total_chunks": 9, "page": i} if i % 4 else None)),
               score=1.0 - 0.07 * i, rank=i + 1) for i in range(8)]
    out["assembler"] = {"hits": [{"id": h.chunk.id, "document_id": h.chunk.document_id, "content": h.chunk.content,
                                  "chunk_index": h.chunk.chunk_index,
                                  "metadata_items": list(h.chunk.metadata.items()) if h.chunk.metadata else None,  # key order matters
                                  "score": h.score,
                                  "rank": h.rank} for h in hits], "cases": []}
    for style in ("markdown", "plain", "json"):
        for inc in (True, False):
            for budget in (4000, 300, 10):
                out["assembler"]["cases"].append({"style": style, "include_metadata": inc, "max_len": budget,
                                                  "text": ca.ContextAssembler(budget).assemble(hits, inc, style)})
    (HERE / "reference_glue.json").write_text(json.dumps(out, sort_keys=True, ensure_ascii=False, separators=(",", ":")))
    print("wrote", HERE / "reference_glue.json", len(out["chroma"]), "chroma cases,", len(out["faiss"]), "faiss cases")


if __name__ == "__main__":
    run()
