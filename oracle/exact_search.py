"""CPU oracle for the dense-retrieval hot path (TEST INFRASTRUCTURE — not product code).

PARITY UNPINNED: the reference holds no golden vector / known-answer test for this
path and its engines (chromadb==1.3.4, faiss-cpu==1.12.0; pinned in
/root/reference/uv.lock:764-765,1286-1287) are not installable here, so this file restates
their *published* exact-search semantics and is anchored on the reference's call sites:

  * utu/rag/storage/implementations/chroma_store.py:46-59   metric names cosine/l2/ip
  * utu/rag/storage/implementations/chroma_store.py:118-135 query → score = 1 - distance
  * utu/rag/storage/implementations/faiss_store.py:98-110   cosine = normalize_L2 + IndexFlatIP
  * utu/rag/storage/implementations/faiss_store.py:143-199  search, post-filter, score map

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.  The product (youtu-rag_b200/) never does.

Numeric contract restated here:
  rows / queries under `cosine` are L2-normalised with an fp64 sum of squares and an fp64
  divide, rounded fp64→fp32→storage dtype (bf16 round-to-nearest-even, or fp32);
  scores are accumulated in fp64 over the *stored-dtype-rounded* operands;
  ordering is (score desc, row id asc); a filter mask is a PRE-filter (Chroma semantics).
"""

from __future__ import annotations

import numpy as np

METRICS = ("cosine", "dot", "euclidean")
DTYPES = ("bf16", "f32")


# --------------------------------------------------------------------------- rounding
def bf16_round(x: np.ndarray) -> np.ndarray:
    """fp32 → bf16 (round-to-nearest-even) → fp32, elementwise."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    lsb = (u >> np.uint64(16)) & np.uint64(1)
    r = ((u + np.uint64(0x7FFF) + lsb) & np.uint64(0xFFFF0000)).astype(np.uint32)
    out = r.view(np.float32).copy()
    nan = np.isnan(x)
    if nan.any():
        out[nan] = np.float32(np.nan)
    return out.reshape(x.shape)


def bf16_bits(x: np.ndarray) -> np.ndarray:
    """The uint16 bit patterns of bf16_round(x)."""
    return (bf16_round(x).view(np.uint32) >> np.uint32(16)).astype(np.uint16)


def bf16_bits_to_f32(b: np.ndarray) -> np.ndarray:
    return (np.ascontiguousarray(b, dtype=np.uint16).astype(np.uint32) << np.uint32(16)).view(np.float32)


def round_to_storage(x: np.ndarray, dtype: str) -> np.ndarray:
    if dtype == "bf16":
        return bf16_round(x)
    if dtype == "f32":
        return np.ascontiguousarray(x, dtype=np.float32)
    raise ValueError(f"unknown storage dtype {dtype!r}")


# --------------------------------------------------------------------------- ingest
def l2_normalize(x: np.ndarray) -> np.ndarray:
    """faiss.normalize_L2 restated (faiss_store.py:107-108,148-149): x / ||x||, zero rows untouched.
    Sum of squares and divide in fp64, result rounded to fp32 (our reproducibility pin)."""
    x = np.atleast_2d(np.asarray(x, dtype=np.float32))
    xd = x.astype(np.float64)
    ss = np.einsum("ij,ij->i", xd, xd)
    nrm = np.sqrt(ss)
    nrm[ss == 0.0] = 1.0
    return (xd / nrm[:, None]).astype(np.float32)


def prepare(x: np.ndarray, metric: str, dtype: str) -> np.ndarray:
    """What the store keeps for rows (and what the kernels see for queries), as fp32 values."""
    if metric not in METRICS:
        metric = "cosine"  # chroma_store.py:52 `distance_metric_map.get(..., "cosine")`
    x = np.atleast_2d(np.asarray(x, dtype=np.float32))
    if metric == "cosine":
        x = l2_normalize(x)
    return round_to_storage(x, dtype)


# --------------------------------------------------------------------------- scores
def scores_f64(rows: np.ndarray, q: np.ndarray, metric: str) -> np.ndarray:
    """score = 1 - distance for every metric (chroma_store.py:132-135):
    cosine → cos-sim (operands already unit-norm), ip → dot, l2 → 1 - ||a-b||^2."""
    rows = np.asarray(rows, dtype=np.float32)
    q = np.asarray(q, dtype=np.float32).reshape(-1)
    if metric == "euclidean":
        d = rows.astype(np.float64) - q.astype(np.float64)[None, :]
        return 1.0 - np.einsum("ij,ij->i", d, d)
    return rows.astype(np.float64) @ q.astype(np.float64)


def order_desc_id_asc(scores: np.ndarray, ids: np.ndarray) -> np.ndarray:
    """Indices that sort by (score desc, id asc)."""
    return np.lexsort((ids, -scores))


def exact_topk(rows, q, k: int, metric: str = "cosine", mask: np.ndarray | None = None,
               block: int = 262144):
    """Exact pre-filtered top-k of ONE prepared query against prepared rows.

    rows, q: outputs of `prepare` (already rounded to the storage dtype).
    mask: optional bool[N]; rows with False are invisible (pre-filter, chroma `where`).
    Returns (ids int64[<=k], scores float64[<=k]) ordered (score desc, id asc)."""
    rows = np.asarray(rows, dtype=np.float32)
    n = rows.shape[0]
    best_s = np.empty(0, dtype=np.float64)
    best_i = np.empty(0, dtype=np.int64)
    for b0 in range(0, n, block):
        b1 = min(n, b0 + block)
        s = scores_f64(rows[b0:b1], q, metric)
        ids = np.arange(b0, b1, dtype=np.int64)
        if mask is not None:
            keep = np.asarray(mask[b0:b1], dtype=bool)
            s, ids = s[keep], ids[keep]
        s = np.concatenate([best_s, s])
        ids = np.concatenate([best_i, ids])
        if s.shape[0] > k:
            # keep everything tied with the k-th score so the id tie-break stays exact
            kth = np.partition(s, s.shape[0] - k)[s.shape[0] - k]
            keep = s >= kth
            s, ids = s[keep], ids[keep]
        o = order_desc_id_asc(s, ids)[:k]
        best_s, best_i = s[o], ids[o]
    return best_i, best_s


def exact_topk_batch(rows, queries, k: int, metric: str = "cosine", mask=None):
    """Loop of exact_topk (base_retriever.py:95-99 is a loop of single searches)."""
    queries = np.atleast_2d(np.asarray(queries, dtype=np.float32))
    out = []
    for j in range(queries.shape[0]):
        m = None
        if mask is not None:
            m = mask if np.ndim(mask) == 1 else mask[j]
        out.append(exact_topk(rows, queries[j], k, metric, m))
    return out


# --------------------------------------------------------------------------- masks
def pack_mask(mask: np.ndarray) -> np.ndarray:
    """bool[N] → uint32[ceil(N/32)], bit i of word w = row 32*w+i (LSB first)."""
    mask = np.asarray(mask, dtype=bool)
    n = mask.shape[0]
    pad = (-n) % 32
    if pad:
        mask = np.concatenate([mask, np.zeros(pad, dtype=bool)])
    return np.packbits(mask, bitorder="little").view(np.uint32).copy()


def unpack_mask(words: np.ndarray, n: int) -> np.ndarray:
    b = np.unpackbits(np.ascontiguousarray(words, dtype=np.uint32).view(np.uint8), bitorder="little")
    return b[:n].astype(bool)


# --------------------------------------------------------------------------- sort keys
def score_key_u32(s: np.ndarray) -> np.ndarray:
    """Monotone map fp32 → uint32 used by the kernels' 64-bit selection keys."""
    u = np.asarray(s, dtype=np.float32).view(np.uint32)
    neg = (u >> np.uint32(31)).astype(bool)
    return np.where(neg, ~u, u | np.uint32(0x80000000)).astype(np.uint32)


# --------------------------------------------------------------------------- CPU baseline
def faiss_flat_search(rows_f32: np.ndarray, q_f32: np.ndarray, k: int, metric: str = "cosine",
                      keep=None):
    """Restatement of FAISSVectorStore.search (faiss_store.py:143-199) — the reference's only
    exact variant — in fp32 BLAS: IndexFlatIP on normalised rows for cosine (score = IP),
    IndexFlatL2 otherwise (score = 1/(1+d)).  `keep`: optional predicate row→bool, applied as
    a POST-filter over top_k*10 (faiss_store.py:151-152,169-176).  Used as bench.py's CPU
    baseline (kind "port").  rows_f32 must already be normalised for cosine (as the store does
    on add, faiss_store.py:107-108)."""
    n = rows_f32.shape[0]
    if n == 0:
        return np.empty(0, np.int64), np.empty(0, np.float32)
    q = np.asarray(q_f32, dtype=np.float32).reshape(1, -1)
    if metric == "cosine":
        nrm = np.float32(np.sqrt((q * q).sum()))
        if nrm > 0:
            q = q / nrm
        d = rows_f32 @ q[0]                       # IndexFlatIP: larger is better
        search_k = min(k * 10 if keep is not None else k, n)
        part = np.argpartition(-d, search_k - 1)[:search_k]
        part = part[np.lexsort((part, -d[part]))]
        sim = d[part]
    else:
        diff_sq = (rows_f32 * rows_f32).sum(1) - 2.0 * (rows_f32 @ q[0]) + (q * q).sum()
        search_k = min(k * 10 if keep is not None else k, n)
        part = np.argpartition(diff_sq, search_k - 1)[:search_k]
        part = part[np.lexsort((part, diff_sq[part]))]
        sim = 1.0 / (1.0 + diff_sq[part])
    if keep is not None:
        sel = np.fromiter((bool(keep(int(i))) for i in part), dtype=bool, count=part.shape[0])
        part, sim = part[sel], sim[sel]
    return part[:k].astype(np.int64), sim[:k].astype(np.float32)
