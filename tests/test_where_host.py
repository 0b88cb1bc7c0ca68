"""Host side of the metadata filter: normalisation, validation and compilation (CPU only).

The compiled postfix program is executed here by a small test-only interpreter over the columnar
table and compared with the row-at-a-time oracle; the CUDA evaluator (K4) is compared with the same
oracle in tests/test_gpu_where.py.
"""

import json
import struct
from pathlib import Path

import numpy as np
import pytest
from hypothesis import given, settings
from hypothesis import strategies as st

from oracle import where_eval as ow
from youtu_rag_b200 import native
from youtu_rag_b200.metadata import MetadataTable
from youtu_rag_b200.where import compile_where, normalize_filters, validate_where

GOLDEN = json.loads((Path(__file__).parent / "golden" / "reference_glue.json").read_text())


def interpret(prog, table: MetadataTable) -> np.ndarray:
    """Reference interpreter of a yrb_where program (mirrors the semantics K4 must have)."""
    n = table.rows
    cols = {c.col_id: c for c in table.columns.values()}
    stack = []
    for tok in prog.postfix:
        if tok >= 0:
            col_id, op, ob, oc = prog.leaves[tok]
            hit = np.zeros(n, bool)
            if col_id >= 0:
                c = cols[col_id]
                c.extend_to(n)
                vals, pres = c.values[:n], c.present[:n].astype(bool)
                raw = prog.operands[ob:ob + oc]
                if c.col_type == native.COL_F64:
                    opnds = [struct.unpack("<d", struct.pack("<q", r))[0] for r in raw]
                else:
                    opnds = list(raw)
                o = opnds[0]
                cmp = {0: vals == o, 1: vals == o, 2: vals > o, 3: vals >= o, 4: vals < o, 5: vals <= o}.get(op)
                if cmp is None:
                    cmp = np.isin(vals, opnds)
                hit = pres & cmp
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


def test_normalize_filters_matches_reference_golden():
    seen = 0
    for rec in GOLDEN["chroma"]:
        if "where_passed_to_engine" in rec:
            assert normalize_filters(rec["filters"]) == rec["where_passed_to_engine"], rec["filters"]
            assert ow.normalize_filters(rec["filters"]) == rec["where_passed_to_engine"]
            seen += 1
    assert seen > 100


def test_validation_errors_match_reference_golden():
    for rec in GOLDEN["chroma"]:
        where = normalize_filters(rec["filters"])
        if rec.get("error") == "ValueError":
            with pytest.raises(ValueError):
                validate_where(where)
        elif where is not None:
            validate_where(where)


def _table(metas):
    t = MetadataTable()
    for m in metas:
        MetadataTable.validate(m)
    t.append(metas)
    return t


def test_compiled_program_on_golden_corpus():
    metas = GOLDEN["corpus"]["metadatas"]
    t = _table(metas)
    for flt in {json.dumps(r["filters"], sort_keys=True) for r in GOLDEN["chroma"] if "error" not in r}:
        where = normalize_filters(json.loads(flt))
        prog, _ = compile_where(where, t)
        want = ow.eval_where(where, metas)
        got = interpret(prog, t) if prog else np.ones(len(metas), bool)
        assert np.array_equal(got, want), flt


# ---- property test: random metadata + random where trees
_fields = st.sampled_from(["a", "b", "c", "d"])
_scalars = st.one_of(st.integers(-3, 3), st.floats(-3, 3, allow_nan=False).map(lambda x: round(x * 2) / 2),
                     st.sampled_from(["x", "y", "z"]), st.booleans())
_meta = st.dictionaries(_fields, _scalars, max_size=4)


def _same_type_lists():
    return st.one_of(st.lists(st.integers(-3, 3), min_size=1, max_size=4),
                     st.lists(st.sampled_from([-1.5, 0.0, 0.5, 2.0]), min_size=1, max_size=4),
                     st.lists(st.sampled_from(["x", "y", "q"]), min_size=1, max_size=4),
                     st.lists(st.booleans(), min_size=1, max_size=2))


_leaf = st.one_of(
    st.tuples(_fields, _scalars).map(lambda t: {t[0]: t[1]}),
    st.tuples(_fields, st.sampled_from(["$eq", "$ne"]), _scalars).map(lambda t: {t[0]: {t[1]: t[2]}}),
    st.tuples(_fields, st.sampled_from(["$gt", "$gte", "$lt", "$lte"]),
              st.one_of(st.integers(-3, 3), st.sampled_from([-1.5, 0.0, 0.5, 2.0]))).map(lambda t: {t[0]: {t[1]: t[2]}}),
    st.tuples(_fields, st.sampled_from(["$in", "$nin"]), _same_type_lists()).map(lambda t: {t[0]: {t[1]: t[2]}}),
)
_tree = st.recursive(_leaf, lambda c: st.tuples(st.sampled_from(["$and", "$or"]), st.lists(c, min_size=2, max_size=3))
                     .map(lambda t: {t[0]: t[1]}), max_leaves=8)


@settings(max_examples=300, deadline=None)
@given(metas=st.lists(_meta, min_size=1, max_size=40), where=_tree)
def test_compiled_program_equals_oracle(metas, where):
    t = _table(metas)
    prog, _ = compile_where(where, t)
    assert np.array_equal(interpret(prog, t), ow.eval_where(where, metas))


_junk = st.recursive(
    st.one_of(st.none(), st.integers(-2, 2), st.text(max_size=3), st.booleans(), st.floats(allow_nan=False, width=16)),
    lambda c: st.one_of(st.lists(c, max_size=3),
                        st.dictionaries(st.sampled_from(["$and", "$or", "$eq", "$in", "$gt", "$bad", "f", "g"]), c, max_size=2)),
    max_leaves=6)


@settings(max_examples=500, deadline=None)
@given(w=_junk)
def test_product_and_oracle_validators_agree(w):
    def ok(fn):
        try:
            fn(w)
            return True
        except ValueError:
            return False
    assert ok(validate_where) == ok(ow.validate_where)


def test_metadata_rejects_unsupported_values():
    with pytest.raises(ValueError):
        MetadataTable.validate({"a": [1, 2]})
    with pytest.raises(ValueError):
        MetadataTable.validate({"a": 1 << 70})
    with pytest.raises(ValueError):
        compile_where({"a": {"$eq": 1 << 70}}, MetadataTable())
