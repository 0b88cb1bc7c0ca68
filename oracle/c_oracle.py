"""ctypes access to oracle/_build/liboracle.so (the plain-C restatement; TEST INFRASTRUCTURE only)."""

from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB = None


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        so = _HERE / "_build" / "liboracle.so"
        if not so.exists() or so.stat().st_mtime < (_HERE / "exact_search.c").stat().st_mtime:
            subprocess.run(["make", "-C", str(_HERE)], check=True, capture_output=True)
        L = C.CDLL(str(so))
        L.oracle_topk.restype = C.c_int
        L.oracle_topk.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                  C.c_void_p, C.c_void_p]
        L.oracle_normalize.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_void_p]
        L.oracle_bf16_rne.restype = C.c_uint16
        L.oracle_bf16_rne.argtypes = [C.c_float]
        _LIB = L
    return _LIB


def topk(stored: np.ndarray, q_prepared: np.ndarray, k: int, metric: str, mask_words: np.ndarray | None = None):
    """stored: uint16 [n, ld] (bf16 bits) or float32 [n, ld]; q_prepared: float32 [dim]."""
    stored = np.ascontiguousarray(stored)
    dtype = 0 if stored.dtype == np.uint16 else 1
    q = np.ascontiguousarray(q_prepared, dtype=np.float32)
    n, ld = stored.shape
    ids = np.empty(k, np.int64)
    scores = np.empty(k, np.float64)
    m = None if mask_words is None else np.ascontiguousarray(mask_words, dtype=np.uint32)
    have = lib().oracle_topk(stored.ctypes.data, dtype, n, q.shape[0], ld, q.ctypes.data,
                             {"cosine": 0, "dot": 1, "euclidean": 2}[metric], None if m is None else m.ctypes.data, k,
                             ids.ctypes.data, scores.ctypes.data)
    return ids[:have], scores[:have]


def normalize(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty_like(x)
    lib().oracle_normalize(x.ctypes.data, x.shape[0], x.shape[1], out.ctypes.data)
    return out


def bf16_bits(x: np.ndarray) -> np.ndarray:
    f = lib().oracle_bf16_rne
    return np.fromiter((f(float(v)) for v in np.asarray(x, np.float32).ravel()), dtype=np.uint16).reshape(np.shape(x))
