"""Columnar metadata of one collection: the host side of the K4 filter evaluator.

Chroma keeps chunk metadata in sqlite and pre-filters there (collection.query(where=…),
utu/rag/storage/implementations/chroma_store.py:118-120).  Here every (field, value-type) pair
becomes one typed column — int64 / float64 / dictionary-coded str / bool, each with a presence
bit — that is mirrored to the GPU the first time a filter references it, so a `where` turns into
the row bitmask with one kernel (csrc/k4_where.cu) instead of a host scan.

Field value types written by the reference's ingest path: str (`source`,
`file_type`, `index_type`, `document_id`, …), int (`chunk_index`, `*_min_stamp/_max_stamp`),
float (`importance_score`, `success_rate`).
"""

from __future__ import annotations

from typing import Any

import numpy as np

from . import native

_NP = {native.COL_I64: np.int64, native.COL_F64: np.float64, native.COL_CODE: np.int32, native.COL_BOOL: np.uint8}
_I64_MIN, _I64_MAX = -(1 << 63), (1 << 63) - 1


def type_class(v: Any) -> int | None:
    """Chroma's typed metadata columns: bool before int (bool is an int subclass in Python)."""
    if isinstance(v, (bool, np.bool_)):
        return native.COL_BOOL
    if isinstance(v, (int, np.integer)):
        return native.COL_I64
    if isinstance(v, (float, np.floating)):
        return native.COL_F64
    if isinstance(v, str):
        return native.COL_CODE
    return None


class _Column:
    __slots__ = ("col_id", "col_type", "values", "present", "n", "synced")

    def __init__(self, col_id: int, col_type: int):
        self.col_id, self.col_type = col_id, col_type
        self.values = np.zeros(1024, dtype=_NP[col_type])
        self.present = np.zeros(1024, dtype=np.uint8)
        self.n = 0        # rows covered so far (rows >= n are absent)
        self.synced = 0   # rows already mirrored to the device

    def extend_to(self, rows: int) -> None:
        if rows > self.values.shape[0]:
            cap = max(rows, self.values.shape[0] * 2)
            self.values = np.concatenate([self.values, np.zeros(cap - self.values.shape[0], dtype=self.values.dtype)])
            self.present = np.concatenate([self.present, np.zeros(cap - self.present.shape[0], dtype=np.uint8)])
        self.n = max(self.n, rows)


class MetadataTable:
    def __init__(self) -> None:
        self.columns: dict[tuple[str, int], _Column] = {}
        self.dictionaries: dict[str, dict[str, int]] = {}
        self.rows = 0

    @staticmethod
    def validate(meta: dict[str, Any]) -> dict[str, Any]:
        """Chroma rejects metadata values that are not str / int / float / bool.  Returns the dict with numpy
        scalars (np.int64, np.float32, np.bool_ …) coerced to plain Python values, so that what is kept on the
        host, written to disk (json) and returned in Chunk.metadata is what Chroma would hand back."""
        out: dict[str, Any] = {}
        for k, v in meta.items():
            if not isinstance(k, str):
                raise ValueError(f"Expected metadata key to be a str, got {k!r}")
            t = type_class(v)
            if t is None:
                raise ValueError(f"Expected metadata value to be a str, int, float or bool, got {v!r} for key {k!r}")
            if t == native.COL_I64 and not (_I64_MIN <= int(v) <= _I64_MAX):
                raise ValueError(f"metadata int {v} for key {k!r} does not fit in int64")
            out[k] = (bool(v) if t == native.COL_BOOL else int(v) if t == native.COL_I64
                      else float(v) if t == native.COL_F64 else str(v))
        return out

    def append(self, metadatas: list[dict[str, Any]]) -> None:
        """Rows [self.rows, self.rows+len) get these metadata dicts (already validated)."""
        base = self.rows
        for i, meta in enumerate(metadatas):
            r = base + i
            for k, v in meta.items():
                t = type_class(v)
                col = self.columns.get((k, t))
                if col is None:
                    col = self.columns[(k, t)] = _Column(len(self.columns), t)
                col.extend_to(r + 1)
                if t == native.COL_CODE:
                    d = self.dictionaries.setdefault(k, {})
                    v = d.setdefault(v, len(d))
                col.values[r] = v
                col.present[r] = 1
        self.rows = base + len(metadatas)

    def column(self, field: str, col_type: int) -> _Column | None:
        return self.columns.get((field, col_type))

    def code_of(self, field: str, s: str) -> int:
        """Dictionary code of a string operand; -2 (never stored) when the string was never seen."""
        return self.dictionaries.get(field, {}).get(s, -2)

    def sync(self, index: native.Index, cols: list[_Column]) -> None:
        """Mirror the not-yet-uploaded tail of the referenced columns to the device."""
        for col in cols:
            col.extend_to(self.rows)
            if col.synced < self.rows:
                a, b = col.synced, self.rows
                index.column_write(col.col_id, col.col_type, a, col.values[a:b], col.present[a:b])
                col.synced = b

    def clear(self) -> None:
        self.columns.clear()
        self.dictionaries.clear()
        self.rows = 0
