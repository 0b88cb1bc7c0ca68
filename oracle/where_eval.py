"""CPU oracle for the metadata filter (`where`) language (TEST INFRASTRUCTURE — not product code).

PARITY UNPINNED: nothing in /root/reference tests filter semantics and chromadb==1.3.4
(uv.lock:764-765) is not installable here.  This restates (a) the reference's own filter
normalisation, chroma_store.py:104-116, and (b) the published behaviour of Chroma's
`validate_where` + metadata `where` matching, as driven by the reference's filter producers
(kb_search_toolkit.py:63-96, meta_retrieval_toolkit.py:102-255, memory_store.py:403-417,
chroma_retrical_text2sql.py:169-174).  Row-at-a-time pure Python over raw metadata dicts —
deliberately NOT sharing code or data layout with the product's columnar/GPU evaluator.

Semantics pinned by this build:
  * values are typed: str / bool / int / float; an operand only matches stored values of
    its own type (Chroma keeps string/int/float/bool in separate typed columns).
  * `$eq` false on a missing key; `$ne` ≡ NOT `$eq`, `$nin` ≡ NOT `$in` (so both are TRUE
    on a missing key or a value of another type).
  * `$gt/$gte/$lt/$lte` accept int or float operands only.
  * `$and` / `$or` need a list of ≥ 2 expressions; a dict must hold exactly one key.
"""

from __future__ import annotations

from typing import Any

import numpy as np

_CMP = ("$gt", "$gte", "$lt", "$lte")
_OPS = _CMP + ("$ne", "$eq", "$in", "$nin")


def normalize_filters(filters: dict | None) -> dict | None:
    """chroma_store.py:104-116: pass through when any top-level `$op` or nested operator dict
    is present, else wrap every value as {"$eq": v} (multi-key dicts are NOT and-ed)."""
    if not filters:
        return None
    if any(k.startswith("$") for k in filters.keys()):
        return filters
    if any(isinstance(v, dict) and any(k.startswith("$") for k in v.keys()) for v in filters.values()):
        return filters
    return {k: {"$eq": v} for k, v in filters.items()}


def validate_where(where: Any) -> None:
    """Chroma `validate_where` restated; raises ValueError like Chroma does."""
    if not isinstance(where, dict):
        raise ValueError(f"Expected where to be a dict, got {where}")
    if len(where) != 1:
        raise ValueError(f"Expected where to have exactly one operator, got {where}")
    for key, value in where.items():
        if not isinstance(key, str):
            raise ValueError(f"Expected where key to be a str, got {key}")
        if key in ("$and", "$or"):
            if not isinstance(value, list):
                raise ValueError(f"Expected where value for {key} to be a list of where expressions, got {value}")
            if len(value) <= 1:
                raise ValueError(
                    f"Expected where value for {key} to be a list with at least two where expressions, got {value}")
            for w in value:
                validate_where(w)
            continue
        if key.startswith("$"):
            raise ValueError(f"Expected where key to be a metadata field, $and or $or, got {key}")
        if isinstance(value, dict):
            if len(value) != 1:
                raise ValueError(f"Expected operator expression to have exactly one operator, got {value}")
            for op, operand in value.items():
                if op not in _OPS:
                    raise ValueError(f"Expected where operator to be one of {', '.join(_OPS)}, got {op}")
                if op in _CMP:
                    if isinstance(operand, bool) or not isinstance(operand, (int, float)):
                        raise ValueError(
                            f"Expected operand value to be an int or a float for operator {op}, got {operand}")
                elif op in ("$in", "$nin"):
                    if not isinstance(operand, list):
                        raise ValueError(f"Expected operand value to be a list for operator {op}, got {operand}")
                    if len(operand) == 0 or not all(_tclass(x) is not None for x in operand) or \
                            len({_tclass(x) for x in operand}) != 1:
                        raise ValueError(
                            "Expected where operand value to be a non-empty list, and all values to be "
                            f"of the same type got {operand}")
                else:
                    if _tclass(operand) is None:
                        raise ValueError(f"Expected where operand value to be a str, int, float or bool, got {operand}")
        elif _tclass(value) is None:
            raise ValueError(f"Expected where value to be a str, int, float, bool or operator expression, got {value}")


def _tclass(v: Any) -> str | None:
    if isinstance(v, bool):
        return "bool"
    if isinstance(v, (int, np.integer)):
        return "int"
    if isinstance(v, (float, np.floating)):
        return "float"
    if isinstance(v, str):
        return "str"
    return None


def _eq(stored: Any, operand: Any) -> bool:
    return _tclass(stored) == _tclass(operand) and stored == operand


def _leaf(meta: dict, field: str, op: str, operand: Any) -> bool:
    has = field in meta and meta[field] is not None
    stored = meta.get(field)
    if op == "$eq":
        return has and _eq(stored, operand)
    if op == "$ne":
        return not (has and _eq(stored, operand))
    if op == "$in":
        return has and any(_eq(stored, x) for x in operand)
    if op == "$nin":
        return not (has and any(_eq(stored, x) for x in operand))
    if not has or _tclass(stored) != _tclass(operand):
        return False
    if op == "$gt":
        return stored > operand
    if op == "$gte":
        return stored >= operand
    if op == "$lt":
        return stored < operand
    if op == "$lte":
        return stored <= operand
    raise ValueError(op)


def match(where: dict, meta: dict) -> bool:
    (key, value), = where.items()
    if key == "$and":
        return all(match(w, meta) for w in value)
    if key == "$or":
        return any(match(w, meta) for w in value)
    if isinstance(value, dict):
        (op, operand), = value.items()
        return _leaf(meta, key, op, operand)
    return _leaf(meta, key, "$eq", value)


def eval_where(where: dict | None, metadatas: list[dict]) -> np.ndarray:
    """bool[N]: which rows a validated `where` passes (None → all)."""
    if where is None:
        return np.ones(len(metadatas), dtype=bool)
    validate_where(where)
    return np.fromiter((match(where, m or {}) for m in metadatas), dtype=bool, count=len(metadatas))
