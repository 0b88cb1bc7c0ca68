"""Test-only stand-in for youtu_rag_b200.native.Index, backed by the CPU oracle.

Lets the HOST logic of the product (store.py, memory_store.py, retriever.py: id/row bookkeeping, metadata
columns and their upload order, filter compilation, tombstones, upsert, result shaping) run in the `-m "not gpu"`
suite.  It is never imported by the product — tests monkeypatch `native.Index` with it — and it is not a search
path: the arithmetic is the oracle's (oracle/exact_search.py), the filter programs are executed by a small
interpreter over the columns the store uploads through `column_write`, exactly the data K4 would see.
"""

from __future__ import annotations

import struct

import numpy as np

from oracle import exact_search as ox
from youtu_rag_b200 import native


class FakeIndex:
    def __init__(self, dim: int, metric: str = "cosine", dtype: str = "bf16", device: int = 0, reserve_rows: int = 0):
        self.dim, self.metric, self.dtype, self.device = int(dim), metric, dtype, int(device)
        self._rows = np.zeros((0, self.dim), np.float32)      # stored values (normalised / rounded), as read_rows returns them
        self._live = np.zeros(0, bool)
        self._cols: dict[int, tuple[int, np.ndarray, np.ndarray]] = {}
        self.closed = False

    # ------------------------------------------------------------ residency
    def close(self) -> None:
        self.closed = True

    def reserve(self, rows: int) -> None:
        pass

    def counts(self) -> tuple[int, int]:
        return self._rows.shape[0], int(self._live.sum())

    @property
    def rows(self) -> int:
        return self._rows.shape[0]

    def info(self) -> dict:
        per = 64 if self.dtype == "bf16" else 32
        return dict(dim=self.dim, ld=(self.dim + per - 1) // per * per, metric=native.METRICS[self.metric],
                    dtype=native.DTYPES[self.dtype], device=self.device, capacity=self.rows)

    def append(self, rows: np.ndarray) -> None:
        rows = np.ascontiguousarray(rows, dtype=np.float32)
        if rows.ndim != 2 or rows.shape[1] != self.dim:
            raise ValueError(f"expected rows of shape [n, {self.dim}], got {rows.shape}")
        self._rows = np.concatenate([self._rows, ox.prepare(rows, self.metric, self.dtype)])
        self._live = np.concatenate([self._live, np.ones(rows.shape[0], bool)])

    def read_rows(self, row_ids) -> np.ndarray:
        return self._rows[np.asarray(row_ids, np.int64)].copy()

    def read_raw(self, row_begin: int, n: int):
        """Rows as the device would hold them: [n, ld] bf16 bit patterns (uint16) or fp32, zero padded, + squared norms."""
        ld = self.info()["ld"]
        vals = np.zeros((n, ld), np.float32)
        vals[:, :self.dim] = self._rows[row_begin:row_begin + n]
        sq = (vals.astype(np.float32) ** 2).sum(1, dtype=np.float32)
        return (ox.bf16_bits(vals) if self.dtype == "bf16" else vals), sq

    def append_raw(self, rows: np.ndarray, sqnorm: np.ndarray) -> None:
        ld = self.info()["ld"]
        rows = np.asarray(rows)
        if rows.ndim != 2 or rows.shape[1] != ld or np.asarray(sqnorm).shape[0] != rows.shape[0]:
            raise ValueError(f"expected raw rows [n, {ld}] and n squared norms")
        vals = ox.bf16_bits_to_f32(rows.astype(np.uint16)) if self.dtype == "bf16" else rows.astype(np.float32)
        self._rows = np.concatenate([self._rows, vals[:, :self.dim]])
        self._live = np.concatenate([self._live, np.ones(rows.shape[0], bool)])

    def set_live(self, row_ids, live: bool) -> None:
        self._live[np.asarray(row_ids, np.int64)] = bool(live)

    def clear(self) -> None:
        self._rows, self._live, self._cols = np.zeros((0, self.dim), np.float32), np.zeros(0, bool), {}

    def truncate(self, rows: int) -> None:
        if not 0 <= rows <= self.rows:
            raise native.NativeError(-1, f"truncate to {rows} rows: index holds {self.rows}")
        self._rows, self._live = self._rows[:rows], self._live[:rows]
        self._cols = {c: (t, v[:rows], p[:rows]) for c, (t, v, p) in self._cols.items()}

    # ------------------------------------------------------------ filter
    def column_write(self, col: int, col_type: int, row_begin: int, values: np.ndarray, present: np.ndarray) -> None:
        dt = {native.COL_I64: np.int64, native.COL_F64: np.float64, native.COL_CODE: np.int32, native.COL_BOOL: np.uint8}[col_type]
        values, present = np.asarray(values, dt), np.asarray(present, np.uint8)
        assert values.shape[0] == present.shape[0]
        old = self._cols.get(col)
        n = max(row_begin + values.shape[0], old[1].shape[0] if old else 0)
        v, p = np.zeros(n, dt), np.zeros(n, np.uint8)
        if old:
            assert old[0] == col_type
            v[:old[1].shape[0]], p[:old[2].shape[0]] = old[1], old[2]
        v[row_begin:row_begin + values.shape[0]], p[row_begin:row_begin + values.shape[0]] = values, present
        self._cols[col] = (col_type, v, p)

    def _eval(self, prog) -> np.ndarray:
        """Rows passing a compiled yrb_where program (same semantics as tests/test_where_host.py::interpret, but over
        the uploaded columns; a row past the end of a column has no value)."""
        n = self.rows
        stack = []
        for tok in prog.postfix:
            if tok >= 0:
                col_id, op, ob, oc = prog.leaves[tok]
                hit = np.zeros(n, bool)
                if col_id >= 0:
                    assert col_id in self._cols, "a filter names a column the store did not upload"
                    ctype, vals, pres = self._cols[col_id]
                    v, p = np.zeros(n, vals.dtype), np.zeros(n, bool)
                    m = min(n, vals.shape[0])
                    v[:m], p[:m] = vals[:m], pres[:m].astype(bool)
                    raw = prog.operands[ob:ob + oc]
                    opnds = [struct.unpack("<d", struct.pack("<q", r))[0] for r in raw] if ctype == native.COL_F64 else list(raw)
                    o = opnds[0]
                    cmp = {0: v == o, 1: v == o, 2: v > o, 3: v >= o, 4: v < o, 5: v <= o}.get(op)
                    if cmp is None:
                        cmp = np.isin(v, opnds)
                    hit = p & cmp
                if op in (1, 7):
                    hit = ~hit
                stack.append(hit)
            elif tok == native.TOK_NOT:
                stack.append(~stack.pop())
            else:
                b, a = stack.pop(), stack.pop()
                stack.append(a & b if tok == native.TOK_AND else a | b)
        assert len(stack) == 1
        return stack[0]

    def _mask(self, where, mask) -> np.ndarray:
        m = self._live.copy()
        if where is not None:
            m &= self._eval(where)
        if mask is not None:
            m &= ox.unpack_mask(np.asarray(mask, np.uint32), self.rows)
        return m

    def where_mask(self, where):
        m = self._mask(where, None)
        return ox.pack_mask(m), int(m.sum())

    # ------------------------------------------------------------ search
    def search(self, queries: np.ndarray, k: int, where=None, mask=None, wheres=None, min_score=None):
        q = np.ascontiguousarray(queries, dtype=np.float32)
        if q.ndim == 1:
            q = q[None, :]
        if q.ndim != 2 or q.shape[1] != self.dim:
            raise ValueError(f"expected queries of shape [nq, {self.dim}], got {q.shape}")
        if k < 1:
            raise native.NativeError(-1, f"k must be >= 1 (got {k})")
        nq = q.shape[0]
        ids = np.full((nq, k), -1, np.int64)
        scores = np.full((nq, k), -np.inf, np.float32)
        counts = np.zeros(nq, np.int32)
        for j in range(nq):
            m = self._mask(wheres[j] if wheres is not None else where, mask)
            qi, si = ox.exact_topk(self._rows, ox.prepare(q[j], self.metric, self.dtype)[0], k, self.metric, m)
            if min_score is not None:
                keep = si.astype(np.float32) >= np.float32(min_score)
                qi, si = qi[keep], si[keep]
            counts[j] = qi.shape[0]
            ids[j, :qi.shape[0]], scores[j, :qi.shape[0]] = qi, si
        return ids, scores, counts


class FakeShardedIndex(FakeIndex):
    """Stand-in for native.ShardedIndex: the store sees the same interface and GLOBAL row ids whether a collection lives
    on one GPU or is dealt over several, so the host logic is exercised with one fake; what the constructor was handed
    is kept for the tests to look at."""

    def __init__(self, dim: int, metric: str = "cosine", dtype: str = "bf16", devices=(0,), reserve_rows: int = 0, block_rows: int = 0):
        super().__init__(dim, metric, dtype, int(list(devices)[0]), reserve_rows)
        self.devices, self.block_rows = [int(d) for d in devices], block_rows

    def info(self) -> dict:
        return {**super().info(), "n_devices": len(self.devices), "block_rows": self.block_rows or 16384}
