"""ctypes binding of include/yrb200.h (libyrb200.so, built in-tree by build.py).

There is deliberately no fallback: if the shared library is missing, or no sm_100 device is
visible when an index is created, the caller gets an exception — never a CPU path.
"""

from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("YRB200_LIB", _HERE / "libyrb200.so"))

YRB_OK = 0
ERR_NAMES = {-1: "INVALID", -2: "CUDA", -3: "NOMEM", -4: "UNSUPPORTED", -5: "NODEVICE"}
METRICS = {"cosine": 0, "dot": 1, "euclidean": 2}
DTYPES = {"bf16": 0, "f32": 1}
COL_I64, COL_F64, COL_CODE, COL_BOOL = 0, 1, 2, 3
OPS = {"$eq": 0, "$ne": 1, "$gt": 2, "$gte": 3, "$lt": 4, "$lte": 5, "$in": 6, "$nin": 7}
TOK_AND, TOK_OR, TOK_NOT = -1, -2, -3
PATH_AUTO, PATH_K1, PATH_K2, PATH_K6, PATH_K2_PAIR = 0, 1, 2, 3, 4
FUSED_K_MAX = 128
WHERE_MAX_LEAVES, WHERE_MAX_OPERANDS, WHERE_MAX_TOKENS = 64, 256, 160

# every symbol include/yrb200.h declares (tests check the library exports all of them)
EXPORTS = (
    "yrb_abi_version", "yrb_last_error", "yrb_device_count", "yrb_index_create", "yrb_index_destroy",
    "yrb_index_reserve", "yrb_index_count", "yrb_index_info", "yrb_index_append_host_f32",
    "yrb_index_append_device_f32", "yrb_index_read_rows", "yrb_index_read_raw", "yrb_index_append_raw", "yrb_index_set_live", "yrb_index_clear", "yrb_index_truncate",
    "yrb_index_column_write", "yrb_index_where", "yrb_index_search", "yrb_index_search_ex", "yrb_index_search_multi", "yrb_index_search_device",
    "yrb_index_search_device_ids",
    "yrb_merge_topk_device", "yrb_exchange_handle_bytes", "yrb_exchange_last_error", "yrb_exchange_create",
    "yrb_exchange_connect", "yrb_exchange_merge", "yrb_exchange_search", "yrb_exchange_destroy", "yrb_index_set_path", "yrb_index_set_reserved_sms", "yrb_index_stats", "yrb_index_profile",
    "yrb_index_profile_read", "yrb_index_cache_stats",
    "yrb_shard_locate", "yrb_shard_global", "yrb_shard_rows", "yrb_sharded_create", "yrb_sharded_destroy", "yrb_sharded_count", "yrb_sharded_info", "yrb_sharded_shard",
    "yrb_sharded_append_host_f32", "yrb_sharded_append_device_f32", "yrb_sharded_read_rows", "yrb_sharded_read_raw",
    "yrb_sharded_append_raw", "yrb_sharded_set_live", "yrb_sharded_truncate", "yrb_sharded_clear",
    "yrb_sharded_column_write", "yrb_sharded_where", "yrb_sharded_search", "yrb_sharded_search_multi", "yrb_sharded_search_ex", "yrb_sharded_stats",
)


class NativeError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"yrb200 {ERR_NAMES.get(code, code)}: {message}")
        self.code = code


class WhereLeaf(C.Structure):
    _fields_ = [("col", C.c_int32), ("op", C.c_int32), ("operand_begin", C.c_int32), ("operand_count", C.c_int32)]


class Where(C.Structure):
    _fields_ = [
        ("leaves", C.POINTER(WhereLeaf)), ("n_leaves", C.c_int32),
        ("operands", C.POINTER(C.c_int64)), ("n_operands", C.c_int32),
        ("postfix", C.POINTER(C.c_int32)), ("n_postfix", C.c_int32),
    ]


class SearchOpts(C.Structure):
    _fields_ = [("min_score", C.c_float), ("reserved", C.c_int32 * 7)]


_lib = None


def lib() -> C.CDLL:
    """Load libyrb200.so (once).  Raises if it has not been built — there is no Python fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(
            f"{LIB_PATH} not found: build the CUDA extension first (python youtu-rag_b200/build.py or "
            "__graft_entry__.build()). The B200 backend has no CPU fallback.")
    L = C.CDLL(str(LIB_PATH))
    vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int
    L.yrb_abi_version.restype = i32
    L.yrb_last_error.restype = C.c_char_p
    L.yrb_device_count.argtypes = [C.POINTER(i32)]
    L.yrb_index_create.argtypes = [C.POINTER(vp), i32, i32, i32, i32, i64]
    L.yrb_index_destroy.argtypes = [vp]
    L.yrb_index_reserve.argtypes = [vp, i64]
    L.yrb_index_count.argtypes = [vp, C.POINTER(i64), C.POINTER(i64)]
    L.yrb_index_info.argtypes = [vp] + [C.POINTER(i32)] * 5 + [C.POINTER(i64)]
    L.yrb_index_append_host_f32.argtypes = [vp, vp, i64]
    L.yrb_index_append_device_f32.argtypes = [vp, vp, i64, vp]
    L.yrb_index_read_rows.argtypes = [vp, vp, i64, vp]
    L.yrb_index_read_raw.argtypes = [vp, i64, i64, vp, vp]
    L.yrb_index_append_raw.argtypes = [vp, vp, vp, i64]
    L.yrb_index_set_live.argtypes = [vp, vp, i64, i32]
    L.yrb_index_clear.argtypes = [vp]
    L.yrb_index_truncate.argtypes = [vp, i64]
    L.yrb_index_column_write.argtypes = [vp, i32, i32, i64, i64, vp, vp]
    L.yrb_index_where.argtypes = [vp, C.POINTER(Where), vp, C.POINTER(i64)]
    L.yrb_index_search.argtypes = [vp, vp, i32, i32, C.POINTER(Where), vp, vp, vp, vp]
    L.yrb_index_search_ex.argtypes = [vp, vp, i32, i32, C.POINTER(Where), C.POINTER(C.POINTER(Where)), vp, C.POINTER(SearchOpts), vp, vp, vp]
    L.yrb_sharded_search_ex.argtypes = [vp, vp, i32, i32, C.POINTER(Where), C.POINTER(C.POINTER(Where)), vp, C.POINTER(SearchOpts), vp, vp, vp]
    L.yrb_index_search_multi.argtypes = [vp, vp, i32, i32, C.POINTER(C.POINTER(Where)), vp, vp, vp]
    L.yrb_index_search_device.argtypes = [vp, vp, i32, i32, vp, vp, vp]
    L.yrb_index_search_device_ids.argtypes = [vp, vp, i32, i32, vp, vp, vp, vp, vp]
    L.yrb_merge_topk_device.argtypes = [i32, vp, i32, i32, i32, vp, vp, vp, vp, vp]
    L.yrb_exchange_last_error.restype = C.c_char_p
    L.yrb_exchange_create.argtypes = [C.POINTER(vp), i32, i32, i32, i32, i32, vp]
    L.yrb_exchange_connect.argtypes = [vp, vp]
    L.yrb_exchange_merge.argtypes = [vp, vp, i32, i32, vp, vp, vp, vp, vp]
    L.yrb_exchange_search.argtypes = [vp, vp, vp, i32, i32, vp, vp, vp, vp, vp]
    L.yrb_exchange_destroy.argtypes = [vp]
    L.yrb_index_set_path.argtypes = [vp, i32]
    L.yrb_index_set_reserved_sms.argtypes = [vp, i32]
    L.yrb_index_stats.argtypes = [vp, C.POINTER(i64)]
    L.yrb_index_profile.argtypes = [vp, i32]
    L.yrb_index_profile_read.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(i64)]
    L.yrb_index_cache_stats.argtypes = [vp, C.POINTER(i64), C.POINTER(i64)]
    L.yrb_shard_locate.argtypes = [i32, i32, i64, C.POINTER(i32), C.POINTER(i64)]
    L.yrb_shard_global.argtypes = [i32, i32, i32, i64, C.POINTER(i64)]
    L.yrb_shard_rows.argtypes = [i32, i32, i64, i32, C.POINTER(i64)]
    L.yrb_sharded_create.argtypes = [C.POINTER(vp), C.POINTER(i32), i32, i32, i32, i32, i64, i32]
    L.yrb_sharded_destroy.argtypes = [vp]
    L.yrb_sharded_count.argtypes = [vp, C.POINTER(i64), C.POINTER(i64)]
    L.yrb_sharded_info.argtypes = [vp] + [C.POINTER(i32)] * 6 + [C.POINTER(i64)]
    L.yrb_sharded_shard.argtypes = [vp, i32, C.POINTER(vp), C.POINTER(i32)]
    L.yrb_sharded_append_host_f32.argtypes = [vp, vp, i64]
    L.yrb_sharded_append_device_f32.argtypes = [vp, vp, i64, i32]
    L.yrb_sharded_read_rows.argtypes = [vp, vp, i64, vp]
    L.yrb_sharded_read_raw.argtypes = [vp, i64, i64, vp, vp]
    L.yrb_sharded_append_raw.argtypes = [vp, vp, vp, i64]
    L.yrb_sharded_set_live.argtypes = [vp, vp, i64, i32]
    L.yrb_sharded_truncate.argtypes = [vp, i64]
    L.yrb_sharded_clear.argtypes = [vp]
    L.yrb_sharded_column_write.argtypes = [vp, i32, i32, i64, i64, vp, vp]
    L.yrb_sharded_where.argtypes = [vp, C.POINTER(Where), vp, C.POINTER(i64)]
    L.yrb_sharded_search.argtypes = [vp, vp, i32, i32, C.POINTER(Where), vp, vp, vp, vp]
    L.yrb_sharded_search_multi.argtypes = [vp, vp, i32, i32, C.POINTER(C.POINTER(Where)), vp, vp, vp]
    L.yrb_sharded_stats.argtypes = [vp, C.POINTER(i64), C.POINTER(i64)]
    for name in EXPORTS:
        fn = getattr(L, name)
        if name not in ("yrb_last_error", "yrb_exchange_last_error"):
            fn.restype = i32
    if L.yrb_abi_version() != 1:
        raise RuntimeError(f"libyrb200.so ABI {L.yrb_abi_version()} != 1")
    _lib = L
    return L


def _ck(rc: int) -> None:
    if rc != YRB_OK:
        raise NativeError(rc, (lib().yrb_last_error() or b"").decode("utf-8", "replace"))


def device_count() -> int:
    n = C.c_int(0)
    _ck(lib().yrb_device_count(C.byref(n)))
    return n.value


class CompiledWhere:
    """Owns the ctypes arrays of one `yrb_where` program."""

    def __init__(self, leaves: list[tuple[int, int, int, int]], operands: list[int], postfix: list[int]):
        if len(leaves) > WHERE_MAX_LEAVES or len(operands) > WHERE_MAX_OPERANDS or len(postfix) > WHERE_MAX_TOKENS:
            raise ValueError(
                f"where clause too large for the device evaluator (leaves {len(leaves)}/{WHERE_MAX_LEAVES}, "
                f"operands {len(operands)}/{WHERE_MAX_OPERANDS}, tokens {len(postfix)}/{WHERE_MAX_TOKENS})")
        self._leaves = (WhereLeaf * max(1, len(leaves)))(*[WhereLeaf(*l) for l in leaves])
        self._operands = (C.c_int64 * max(1, len(operands)))(*operands)
        self._postfix = (C.c_int32 * max(1, len(postfix)))(*postfix)
        self.struct = Where(self._leaves, len(leaves), self._operands, len(operands), self._postfix, len(postfix))
        self.leaves, self.operands, self.postfix = leaves, operands, postfix


class Index:
    """One collection's rows on one GPU (opaque `yrb_index*`)."""

    def __init__(self, dim: int, metric: str = "cosine", dtype: str = "bf16", device: int = 0,
                 reserve_rows: int = 0):
        self._h = C.c_void_p()
        self.dim, self.metric, self.dtype, self.device = int(dim), metric, dtype, int(device)
        _ck(lib().yrb_index_create(C.byref(self._h), device, dim, METRICS[metric], DTYPES[dtype], reserve_rows))

    def close(self) -> None:
        if getattr(self, "_h", None) and self._h.value:
            try:
                _lib.yrb_index_destroy(self._h)
            except Exception:  # noqa: BLE001 - interpreter shutdown: module globals may already be gone
                pass
            self._h = None

    __del__ = close

    # ------------------------------------------------------------ residency
    def reserve(self, rows: int) -> None:
        _ck(lib().yrb_index_reserve(self._h, rows))

    def counts(self) -> tuple[int, int]:
        r, l = C.c_int64(), C.c_int64()
        _ck(lib().yrb_index_count(self._h, C.byref(r), C.byref(l)))
        return r.value, l.value

    @property
    def rows(self) -> int:
        return self.counts()[0]

    def info(self) -> dict:
        v = [C.c_int() for _ in range(5)]
        cap = C.c_int64()
        _ck(lib().yrb_index_info(self._h, *[C.byref(x) for x in v], C.byref(cap)))
        return dict(dim=v[0].value, ld=v[1].value, metric=v[2].value, dtype=v[3].value, device=v[4].value,
                    capacity=cap.value)

    def append(self, rows: np.ndarray) -> None:
        rows = np.ascontiguousarray(rows, dtype=np.float32)
        if rows.ndim != 2 or rows.shape[1] != self.dim:
            raise ValueError(f"expected rows of shape [n, {self.dim}], got {rows.shape}")
        _ck(lib().yrb_index_append_host_f32(self._h, rows.ctypes.data, rows.shape[0]))

    def append_device(self, dev_ptr: int, n: int, stream: int = 0) -> None:
        _ck(lib().yrb_index_append_device_f32(self._h, dev_ptr, n, stream))

    def read_rows(self, row_ids) -> np.ndarray:
        ids = np.ascontiguousarray(row_ids, dtype=np.int64)
        out = np.empty((ids.shape[0], self.dim), dtype=np.float32)
        _ck(lib().yrb_index_read_rows(self._h, ids.ctypes.data, ids.shape[0], out.ctypes.data))
        return out

    def read_raw(self, row_begin: int, n: int) -> tuple[np.ndarray, np.ndarray]:
        """Rows exactly as stored: (uint16 [n, ld] bf16 bits or float32 [n, ld], float32 [n] squared norms)."""
        ld = self.info()["ld"]
        rows = np.empty((n, ld), dtype=np.uint16 if self.dtype == "bf16" else np.float32)
        sq = np.empty(n, dtype=np.float32)
        _ck(lib().yrb_index_read_raw(self._h, row_begin, n, rows.ctypes.data, sq.ctypes.data))
        return rows, sq

    def append_raw(self, rows: np.ndarray, sqnorm: np.ndarray) -> None:
        ld = self.info()["ld"]
        rows = np.ascontiguousarray(rows, dtype=np.uint16 if self.dtype == "bf16" else np.float32)
        sqnorm = np.ascontiguousarray(sqnorm, dtype=np.float32)
        if rows.ndim != 2 or rows.shape[1] != ld or sqnorm.shape[0] != rows.shape[0]:
            raise ValueError(f"expected raw rows [n, {ld}] and n squared norms")
        _ck(lib().yrb_index_append_raw(self._h, rows.ctypes.data, sqnorm.ctypes.data, rows.shape[0]))

    def set_live(self, row_ids, live: bool) -> None:
        ids = np.ascontiguousarray(row_ids, dtype=np.int64)
        _ck(lib().yrb_index_set_live(self._h, ids.ctypes.data, ids.shape[0], int(bool(live))))

    def clear(self) -> None:
        _ck(lib().yrb_index_clear(self._h))

    def truncate(self, rows: int) -> None:
        _ck(lib().yrb_index_truncate(self._h, rows))

    # ------------------------------------------------------------ filter
    def column_write(self, col: int, col_type: int, row_begin: int, values: np.ndarray, present: np.ndarray) -> None:
        dt = {COL_I64: np.int64, COL_F64: np.float64, COL_CODE: np.int32, COL_BOOL: np.uint8}[col_type]
        values = np.ascontiguousarray(values, dtype=dt)
        present = np.ascontiguousarray(present, dtype=np.uint8)
        assert values.shape[0] == present.shape[0]
        _ck(lib().yrb_index_column_write(self._h, col, col_type, row_begin, values.shape[0], values.ctypes.data,
                                         present.ctypes.data))

    def where_mask(self, where: CompiledWhere | None) -> tuple[np.ndarray, int]:
        rows = self.rows
        out = np.zeros((rows + 31) // 32, dtype=np.uint32)
        n = C.c_int64()
        _ck(lib().yrb_index_where(self._h, C.byref(where.struct) if where else None, out.ctypes.data, C.byref(n)))
        return out, n.value

    # ------------------------------------------------------------ search
    _SEARCH_EX = "yrb_index_search_ex"

    def search(self, queries: np.ndarray, k: int, where: CompiledWhere | None = None,
               mask: np.ndarray | None = None, wheres: list | None = None, min_score: float | None = None):
        """Host buffers in / out.  Returns (ids int64[nq,k] (-1 pad), scores f32[nq,k], counts int32[nq]).
        `where`: one compiled filter shared by all queries; `wheres`: one (or None) per query; `min_score`: keep only
        hits with score >= min_score (applied inside the scan, base_retriever.py:71)."""
        q = np.ascontiguousarray(queries, dtype=np.float32)
        if q.ndim == 1:
            q = q[None, :]
        if q.ndim != 2 or q.shape[1] != self.dim:
            raise ValueError(f"expected queries of shape [nq, {self.dim}], got {q.shape}")
        nq = q.shape[0]
        ids = np.empty((nq, k), dtype=np.int64)
        scores = np.empty((nq, k), dtype=np.float32)
        counts = np.empty(nq, dtype=np.int32)
        if wheres is not None:
            if where is not None or mask is not None:
                raise ValueError("pass either a shared filter (where/mask) or per-query filters (wheres)")
            if len(wheres) != nq:
                raise ValueError(f"expected {nq} per-query filters, got {len(wheres)}")
        arr = None
        if wheres is not None:
            arr = (C.POINTER(Where) * nq)(*[C.pointer(w.struct) if w is not None else C.POINTER(Where)() for w in wheres])
        mptr = None
        if mask is not None:
            mask = np.ascontiguousarray(mask, dtype=np.uint32)
            if mask.shape[0] < (self.rows + 31) // 32:
                raise ValueError("mask has fewer than ceil(rows/32) words")
            mptr = mask.ctypes.data
        opts = SearchOpts(float("-inf") if min_score is None else float(min_score))
        _ck(getattr(lib(), self._SEARCH_EX)(self._h, q.ctypes.data, nq, k, C.byref(where.struct) if where else None, arr, mptr,
                                            C.byref(opts), ids.ctypes.data, scores.ctypes.data, counts.ctypes.data))
        return ids, scores, counts

    def search_device(self, dev_queries: int, nq: int, k: int, dev_mask: int, dev_out_keys: int, stream: int = 0):
        _ck(lib().yrb_index_search_device(self._h, dev_queries, nq, k, dev_mask or None, dev_out_keys, stream or None))

    def search_device_ids(self, dev_queries: int, nq: int, k: int, dev_mask: int, dev_ids: int, dev_scores: int,
                          dev_counts: int, stream: int = 0):
        _ck(lib().yrb_index_search_device_ids(self._h, dev_queries, nq, k, dev_mask or None, dev_ids, dev_scores,
                                              dev_counts, stream or None))

    def set_path(self, path: int) -> None:
        _ck(lib().yrb_index_set_path(self._h, path))

    def profile(self, enable: bool) -> None:
        _ck(lib().yrb_index_profile(self._h, int(enable)))

    def profile_read(self) -> tuple[float, int]:
        """(summed ms, launches) of the dominant kernel since the last read (CUDA events)."""
        ms, n = C.c_double(), C.c_int64()
        _ck(lib().yrb_index_profile_read(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def set_reserved_sms(self, n: int) -> None:
        _ck(lib().yrb_index_set_reserved_sms(self._h, n))

    def launches(self) -> int:
        n = C.c_int64()
        _ck(lib().yrb_index_stats(self._h, C.byref(n)))
        return n.value

    def cache_stats(self) -> tuple[int, int]:
        """(filter-mask hits, compaction hits) of the per-filter caches."""
        a, b = C.c_int64(), C.c_int64()
        _ck(lib().yrb_index_cache_stats(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value


class ShardedIndex(Index):
    """One collection row-sharded over several GPUs of the box, inside this process (opaque `yrb_sharded*`): same
    methods as `Index`, global row ids, one worker thread per device, cross-GPU merge fused into the kernels
    (csrc/sharded.cu, csrc/xshard.cuh).  `devices` may name a GPU twice (several shards on one GPU — tests)."""

    def __init__(self, dim: int, metric: str = "cosine", dtype: str = "bf16", devices=(0,), reserve_rows: int = 0,
                 block_rows: int = 0):
        self._h = C.c_void_p()
        self.devices = [int(d) for d in devices]
        self.dim, self.metric, self.dtype, self.device = int(dim), metric, dtype, self.devices[0]
        arr = (C.c_int * len(self.devices))(*self.devices)
        _ck(lib().yrb_sharded_create(C.byref(self._h), arr, len(self.devices), dim, METRICS[metric], DTYPES[dtype],
                                     reserve_rows, block_rows))

    def close(self) -> None:
        if getattr(self, "_h", None) and self._h.value:
            try:
                _lib.yrb_sharded_destroy(self._h)
            except Exception:  # noqa: BLE001
                pass
            self._h = None

    __del__ = close

    def reserve(self, rows: int) -> None:  # shards grow on demand
        pass

    def counts(self) -> tuple[int, int]:
        r, l = C.c_int64(), C.c_int64()
        _ck(lib().yrb_sharded_count(self._h, C.byref(r), C.byref(l)))
        return r.value, l.value

    def info(self) -> dict:
        v = [C.c_int() for _ in range(6)]
        cap = C.c_int64()
        _ck(lib().yrb_sharded_info(self._h, *[C.byref(x) for x in v], C.byref(cap)))
        return dict(dim=v[0].value, ld=v[1].value, metric=v[2].value, dtype=v[3].value, device=self.device,
                    n_devices=v[4].value, block_rows=v[5].value, capacity=cap.value)

    def shard(self, s: int) -> tuple[int, int]:
        """(raw yrb_index* of shard s, its device) — for profiling / path forcing only."""
        h, d = C.c_void_p(), C.c_int()
        _ck(lib().yrb_sharded_shard(self._h, s, C.byref(h), C.byref(d)))
        return h.value, d.value

    def append(self, rows: np.ndarray) -> None:
        rows = np.ascontiguousarray(rows, dtype=np.float32)
        if rows.ndim != 2 or rows.shape[1] != self.dim:
            raise ValueError(f"expected rows of shape [n, {self.dim}], got {rows.shape}")
        _ck(lib().yrb_sharded_append_host_f32(self._h, rows.ctypes.data, rows.shape[0]))

    def append_device(self, dev_ptr: int, n: int, stream: int = 0, src_device: int | None = None) -> None:
        _ck(lib().yrb_sharded_append_device_f32(self._h, dev_ptr, n, self.device if src_device is None else src_device))

    def read_rows(self, row_ids) -> np.ndarray:
        ids = np.ascontiguousarray(row_ids, dtype=np.int64)
        out = np.empty((ids.shape[0], self.dim), dtype=np.float32)
        _ck(lib().yrb_sharded_read_rows(self._h, ids.ctypes.data, ids.shape[0], out.ctypes.data))
        return out

    def read_raw(self, row_begin: int, n: int) -> tuple[np.ndarray, np.ndarray]:
        ld = self.info()["ld"]
        rows = np.empty((n, ld), dtype=np.uint16 if self.dtype == "bf16" else np.float32)
        sq = np.empty(n, dtype=np.float32)
        _ck(lib().yrb_sharded_read_raw(self._h, row_begin, n, rows.ctypes.data, sq.ctypes.data))
        return rows, sq

    def append_raw(self, rows: np.ndarray, sqnorm: np.ndarray) -> None:
        ld = self.info()["ld"]
        rows = np.ascontiguousarray(rows, dtype=np.uint16 if self.dtype == "bf16" else np.float32)
        sqnorm = np.ascontiguousarray(sqnorm, dtype=np.float32)
        if rows.ndim != 2 or rows.shape[1] != ld or sqnorm.shape[0] != rows.shape[0]:
            raise ValueError(f"expected raw rows [n, {ld}] and n squared norms")
        _ck(lib().yrb_sharded_append_raw(self._h, rows.ctypes.data, sqnorm.ctypes.data, rows.shape[0]))

    def set_live(self, row_ids, live: bool) -> None:
        ids = np.ascontiguousarray(row_ids, dtype=np.int64)
        _ck(lib().yrb_sharded_set_live(self._h, ids.ctypes.data, ids.shape[0], int(bool(live))))

    def clear(self) -> None:
        _ck(lib().yrb_sharded_clear(self._h))

    def truncate(self, rows: int) -> None:
        _ck(lib().yrb_sharded_truncate(self._h, rows))

    def column_write(self, col: int, col_type: int, row_begin: int, values: np.ndarray, present: np.ndarray) -> None:
        dt = {COL_I64: np.int64, COL_F64: np.float64, COL_CODE: np.int32, COL_BOOL: np.uint8}[col_type]
        values = np.ascontiguousarray(values, dtype=dt)
        present = np.ascontiguousarray(present, dtype=np.uint8)
        assert values.shape[0] == present.shape[0]
        _ck(lib().yrb_sharded_column_write(self._h, col, col_type, row_begin, values.shape[0], values.ctypes.data,
                                           present.ctypes.data))

    def where_mask(self, where: CompiledWhere | None) -> tuple[np.ndarray, int]:
        out = np.zeros((self.rows + 31) // 32, dtype=np.uint32)
        n = C.c_int64()
        _ck(lib().yrb_sharded_where(self._h, C.byref(where.struct) if where else None, out.ctypes.data, C.byref(n)))
        return out, n.value

    _SEARCH_EX = "yrb_sharded_search_ex"

    def _each_shard(self, fn_name: str, *args) -> None:
        for s in range(len(self.devices)):
            _ck(getattr(lib(), fn_name)(C.c_void_p(self.shard(s)[0]), *args))

    def set_path(self, path: int) -> None:
        self._each_shard("yrb_index_set_path", path)

    def profile(self, enable: bool) -> None:
        self._each_shard("yrb_index_profile", int(enable))

    def set_reserved_sms(self, n: int) -> None:
        self._each_shard("yrb_index_set_reserved_sms", n)

    def launches(self) -> int:
        n, m = C.c_int64(), C.c_int64()
        _ck(lib().yrb_sharded_stats(self._h, C.byref(n), C.byref(m)))
        return n.value

    def search_device(self, *a, **k):
        raise NativeError(-4, "a sharded index answers through host buffers (the merged result lands in pinned host memory)")

    search_device_ids = search_device


def shard_locate(n_shards: int, block_rows: int, global_row: int) -> tuple[int, int]:
    s, l = C.c_int(), C.c_int64()
    _ck(lib().yrb_shard_locate(n_shards, block_rows, global_row, C.byref(s), C.byref(l)))
    return s.value, l.value


def shard_global(n_shards: int, block_rows: int, shard: int, local_row: int) -> int:
    g = C.c_int64()
    _ck(lib().yrb_shard_global(n_shards, block_rows, shard, local_row, C.byref(g)))
    return g.value


def shard_rows(n_shards: int, block_rows: int, total_rows: int, shard: int) -> int:
    r = C.c_int64()
    _ck(lib().yrb_shard_rows(n_shards, block_rows, total_rows, shard, C.byref(r)))
    return r.value


def merge_topk_device(device: int, dev_keys: int, parts: int, nq: int, k: int, dev_row_base: int, dev_ids: int,
                      dev_scores: int, dev_counts: int, stream: int = 0) -> None:
    _ck(lib().yrb_merge_topk_device(device, dev_keys, parts, nq, k, dev_row_base, dev_ids, dev_scores,
                                    dev_counts or None, stream or None))


class Exchange:
    """K7: the merge collective over NVLink peer memory (opaque `yrb_exchange*`)."""

    def __init__(self, device: int, world: int, rank: int, nq_cap: int, k_cap: int):
        self._h = C.c_void_p()
        self.handle_bytes = lib().yrb_exchange_handle_bytes()
        buf = (C.c_ubyte * self.handle_bytes)()
        self._ckx(lib().yrb_exchange_create(C.byref(self._h), device, world, rank, nq_cap, k_cap, buf))
        self.handles = bytes(buf)
        self.world, self.nq_cap, self.k_cap = world, nq_cap, k_cap

    @staticmethod
    def _ckx(rc: int) -> None:
        if rc != YRB_OK:
            raise NativeError(rc, (lib().yrb_exchange_last_error() or b"").decode("utf-8", "replace"))

    def connect(self, all_handles: bytes) -> None:
        assert len(all_handles) == self.world * self.handle_bytes
        buf = (C.c_ubyte * len(all_handles)).from_buffer_copy(all_handles)
        self._ckx(lib().yrb_exchange_connect(self._h, buf))

    def merge(self, dev_local_keys: int, nq: int, k: int, dev_row_base: int, dev_ids: int, dev_scores: int,
              dev_counts: int, stream: int = 0) -> None:
        self._ckx(lib().yrb_exchange_merge(self._h, dev_local_keys, nq, k, dev_row_base, dev_ids, dev_scores,
                                           dev_counts or None, stream or None))

    def search(self, index: "Index", queries: np.ndarray, k: int, dev_mask: int, dev_row_base: int):
        """This rank's whole sharded search through host buffers, enqueued from C (yrb_exchange_search)."""
        q = np.ascontiguousarray(queries, dtype=np.float32)
        if q.ndim == 1:
            q = q[None, :]
        nq = q.shape[0]
        ids = np.empty((nq, k), dtype=np.int64)
        scores = np.empty((nq, k), dtype=np.float32)
        counts = np.empty(nq, dtype=np.int32)
        _ck(lib().yrb_exchange_search(self._h, index._h, q.ctypes.data, nq, k, dev_mask or None, dev_row_base,
                                      ids.ctypes.data, scores.ctypes.data, counts.ctypes.data))
        return ids, scores, counts

    def close(self) -> None:
        if getattr(self, "_h", None) and self._h.value:
            try:
                _lib.yrb_exchange_destroy(self._h)
            except Exception:  # noqa: BLE001
                pass
            self._h = None

    __del__ = close


def decode_keys(keys: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """uint64 selection keys → (rows int64, scores float32); key 0 → (-1, -inf).  Mirrors
    yrb_key_row / yrb_key_score in include/yrb200.h (host-side bookkeeping, not a compute path)."""
    keys = np.asarray(keys, dtype=np.uint64)
    m = (keys >> np.uint64(32)).astype(np.uint32)
    u = np.where(m & np.uint32(0x80000000), m & np.uint32(0x7FFFFFFF), ~m).astype(np.uint32)
    scores = u.view(np.float32).copy()
    rows = (~(keys & np.uint64(0xFFFFFFFF)).astype(np.uint32)).astype(np.int64)
    empty = keys == 0
    rows[empty] = -1
    scores[empty] = -np.inf
    return rows, scores
