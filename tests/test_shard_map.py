"""The block-cyclic row map of a sharded collection (csrc/sharded.cu, csrc/xshard.cuh) through its stateless C entries —
no GPU needed: bijection, balance, monotone local order (what keeps ties breaking on the GLOBAL row id), agreement with
a three-line restatement."""

import numpy as np
import pytest
from hypothesis import given, settings
from hypothesis import strategies as st

from youtu_rag_b200 import native


def model(g, n, b):
    blk = g // b
    return blk % n, (blk // n) * b + g % b


@settings(max_examples=200, deadline=None)
@given(n=st.integers(1, 8), shift=st.integers(6, 20), g=st.integers(0, 2**40))
def test_locate_matches_the_model_and_inverts(n, shift, g):
    b = 1 << shift
    s, l = native.shard_locate(n, b, g)
    assert (s, l) == model(g, n, b)
    if l < 2**32 - 1:                                   # a shard's row id is 32 bits of the selection key
        assert native.shard_global(n, b, s, l) == g


@pytest.mark.parametrize("n,b,total", [(1, 64, 1000), (3, 64, 1000), (8, 16384, 10_000_000), (8, 64, 63), (5, 128, 128 * 5 * 7), (2, 1024, 0)])
def test_shards_partition_the_rows_in_order(n, b, total):
    counts = [native.shard_rows(n, b, total, s) for s in range(n)]
    assert sum(counts) == total
    if total >= n * b:
        assert max(counts) - min(counts) <= b           # balanced to within one block
    step = max(1, total // 5000)
    seen = {s: -1 for s in range(n)}
    for g in range(0, total, step):
        s, l = native.shard_locate(n, b, g)
        assert l < counts[s] and l > seen[s]             # inside the shard, and local order follows global order
        seen[s] = l
    # every shard's rows are exactly 0 .. count-1: the last global row of each shard lands on count-1
    last = {}
    for g in range(max(0, total - n * b - b), total):
        s, l = native.shard_locate(n, b, g)
        last[s] = max(last.get(s, -1), l)
    for s, l in last.items():
        assert l == counts[s] - 1


def test_bad_arguments_are_rejected():
    for args in ((0, 64, 1), (9, 64, 1), (2, 63, 1), (2, 96, 1), (2, 64, -1)):
        with pytest.raises(native.NativeError):
            native.shard_locate(*args)
    with pytest.raises(native.NativeError):
        native.shard_global(2, 64, 2, 0)
    with pytest.raises(native.NativeError):
        native.shard_rows(2, 64, 10, 5)
