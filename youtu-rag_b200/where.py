"""Chroma `where` clauses → device filter programs (host side of kernel K4).

Follows the reference's own handling of `filters` in ChromaVectorStore.search
(utu/rag/storage/implementations/chroma_store.py:104-116) and the validation Chroma applies to a
`where` before running it (chromadb==1.3.4 `validate_where`; un-vendored, restated from its
published behaviour — the reference's producers are written against it: single-element `$and` is
avoided at kb_search_toolkit.py:92-96, multi-key dicts are wrapped at chroma_store.py:196-203).

The compiled form is a postfix program over typed leaves (include/yrb200.h `yrb_where`).
"""

from __future__ import annotations

import struct
from typing import Any

from . import native
from .metadata import MetadataTable, type_class

_COMPARE = ("$gt", "$gte", "$lt", "$lte")
_LEAF_OPS = ("$gt", "$gte", "$lt", "$lte", "$ne", "$eq", "$in", "$nin")


def normalize_filters(filters: dict[str, Any] | None) -> dict[str, Any] | None:
    """The three-way rule of chroma_store.py:104-116."""
    if not filters:
        return None
    keys = list(filters.keys())
    if any(isinstance(k, str) and k.startswith("$") for k in keys):
        return filters
    for v in filters.values():
        if isinstance(v, dict) and any(isinstance(k, str) and k.startswith("$") for k in v):
            return filters
    return {k: {"$eq": v} for k, v in filters.items()}


def _check_operand_scalar(op: str, operand: Any) -> None:
    if type_class(operand) is None:
        raise ValueError(f"Expected where operand value to be a str, int, float or bool, got {operand!r}")
    if op in _COMPARE and type_class(operand) not in (native.COL_I64, native.COL_F64):
        raise ValueError(f"Expected operand value to be an int or a float for operator {op}, got {operand!r}")


def validate_where(where: Any) -> None:
    """Raise ValueError exactly where Chroma's validate_where would."""
    if not isinstance(where, dict):
        raise ValueError(f"Expected where to be a dict, got {where!r}")
    if len(where) != 1:
        raise ValueError(f"Expected where to have exactly one operator, got {where!r}")
    key, value = next(iter(where.items()))
    if not isinstance(key, str):
        raise ValueError(f"Expected where key to be a str, got {key!r}")
    if key in ("$and", "$or"):
        if not isinstance(value, list):
            raise ValueError(f"Expected where value for {key} to be a list of where expressions, got {value!r}")
        if len(value) <= 1:
            raise ValueError(
                f"Expected where value for {key} to be a list with at least two where expressions, got {value!r}")
        for sub in value:
            validate_where(sub)
        return
    if key.startswith("$"):
        raise ValueError(f"Expected where key to be a metadata field, $and or $or, got {key!r}")
    if isinstance(value, dict):
        if len(value) != 1:
            raise ValueError(f"Expected operator expression to have exactly one operator, got {value!r}")
        op, operand = next(iter(value.items()))
        if op not in _LEAF_OPS:
            raise ValueError(f"Expected where operator to be one of {', '.join(_LEAF_OPS)}, got {op!r}")
        if op in ("$in", "$nin"):
            if not isinstance(operand, list):
                raise ValueError(f"Expected operand value to be a list for operator {op}, got {operand!r}")
            classes = {type_class(x) for x in operand}
            if not operand or None in classes or len(classes) != 1:
                raise ValueError(
                    f"Expected where operand value to be a non-empty list, and all values to be of the same type "
                    f"got {operand!r}")
        else:
            _check_operand_scalar(op, operand)
    elif type_class(value) is None:
        raise ValueError(f"Expected where value to be a str, int, float, bool or operator expression, got {value!r}")


def _encode(field: str, operand: Any, table: MetadataTable) -> int:
    t = type_class(operand)
    if t == native.COL_BOOL:
        return int(bool(operand))
    if t == native.COL_I64:
        v = int(operand)
        if not (-(1 << 63) <= v < (1 << 63)):
            raise ValueError(f"where operand {operand} does not fit in int64")
        return v
    if t == native.COL_F64:
        return struct.unpack("<q", struct.pack("<d", float(operand)))[0]
    return table.code_of(field, operand)


def compile_where(where: dict[str, Any] | None, table: MetadataTable):
    """validated where → (CompiledWhere, referenced columns).  None → (None, [])."""
    if where is None:
        return None, []
    validate_where(where)
    leaves: list[tuple[int, int, int, int]] = []
    operands: list[int] = []
    postfix: list[int] = []
    used = []

    def leaf(field: str, op: str, operand: Any) -> None:
        vals = operand if isinstance(operand, list) else [operand]
        t = type_class(vals[0])
        col = table.column(field, t)
        col_id = -1
        if col is not None:
            col_id = col.col_id
            if col not in used:
                used.append(col)
        begin = len(operands)
        operands.extend(_encode(field, v, table) for v in vals)
        leaves.append((col_id, native.OPS[op], begin, len(vals)))
        postfix.append(len(leaves) - 1)

    def walk(node: dict[str, Any]) -> None:
        key, value = next(iter(node.items()))
        if key in ("$and", "$or"):
            walk(value[0])
            for sub in value[1:]:
                walk(sub)
                postfix.append(native.TOK_AND if key == "$and" else native.TOK_OR)
        elif isinstance(value, dict):
            op, operand = next(iter(value.items()))
            leaf(key, op, operand)
        else:
            leaf(key, "$eq", value)

    walk(where)
    return native.CompiledWhere(leaves, operands, postfix), used
