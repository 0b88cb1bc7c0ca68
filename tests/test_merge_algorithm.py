"""The pruning + ranking merge of K1 (csrc/k1_gemv_topk.cu::merge_sorted_lists), restated step for step in numpy and
fuzzed against a plain sort.  This checks the ALGORITHM — the bound (depth, m), which lists are copied, that a
row's survivors are a prefix, that ranks are distinct output slots — over list counts and k the GPU tests do not
reach (1..256 lists, k 1..32, short and empty lists).  The CUDA code itself is checked by the -m gpu parity tests."""

import numpy as np
import pytest


def merge_sorted_lists(lists: np.ndarray, k: int):
    n_lists = lists.shape[0]
    depth = (k + n_lists - 1) // n_lists
    m = (k + depth - 1) // depth
    copy = depth == 1
    heads = lists[:, depth - 1]
    T, hot = np.uint64(0), [-1] * 32
    for i in range(n_lists):
        h = heads[i]
        if h == 0:
            continue
        c = int((heads > h).sum())
        if c < m:
            if c == m - 1:
                T = h
            if copy:
                assert hot[c] == -1, "two lists with the same head rank"
                hot[c] = i
    if copy:
        M = np.zeros((k, k), np.uint64)
        for r in range(k):
            if hot[r] >= 0:
                M[r] = lists[hot[r], :k]
    else:
        M = lists[:, :k]
    assert M.shape[0] <= 32, "the prefix scan over rows is one warp wide"
    live = (M != 0) & (M >= T)
    cnt = live.sum(1)
    for r in range(M.shape[0]):
        assert live[r, :cnt[r]].all() and not live[r, cnt[r]:].any(), "survivors of a sorted row are a prefix"
    out = np.zeros(k, np.uint64)
    filled = np.zeros(k, bool)
    ns = int(cnt.sum())
    for r in range(M.shape[0]):
        for e in range(cnt[r]):
            key = M[r, e]
            c = int((M > key).sum())
            if c < k:
                assert not filled[c], "two keys with the same rank"
                out[c], filled[c] = key, True
    assert filled[:min(ns, k)].all() and not filled[min(ns, k):].any()
    return out, min(ns, k)


def random_lists(rng, n_lists, k, fill):
    total = n_lists * k
    keys = rng.choice(np.arange(1, 50 * total + 50, dtype=np.uint64), size=total, replace=False)   # unique, non-zero
    lists = np.zeros((n_lists, k), np.uint64)
    pos = 0
    for i in range(n_lists):
        n = int(rng.integers(0, k + 1)) if fill == "ragged" else (k if fill == "full" else int(rng.random() < 0.15) * int(rng.integers(1, k + 1)))
        lists[i, :n] = np.sort(keys[pos:pos + n])[::-1]
        pos += n
    return lists


@pytest.mark.parametrize("fill", ["full", "ragged", "sparse"])
def test_merge_matches_a_plain_sort(fill):
    rng = np.random.default_rng({"full": 1, "ragged": 2, "sparse": 3}[fill])
    shapes = [(16, k) for k in range(1, 33)] + [(148, k) for k in (1, 2, 10, 31, 32)] + [(146, 10), (256, 32), (1, 1), (1, 32), (2, 32), (31, 32), (33, 32)]
    shapes += [(int(rng.integers(1, 257)), int(rng.integers(1, 33))) for _ in range(60)]
    for n_lists, k in shapes:
        for _ in range(3):
            lists = random_lists(rng, n_lists, k, fill)
            out, n_out = merge_sorted_lists(lists, k)
            flat = np.sort(lists[lists != 0])[::-1][:k]
            want = np.zeros(k, np.uint64)
            want[:flat.shape[0]] = flat
            assert n_out == flat.shape[0] and np.array_equal(out, want), (n_lists, k, fill)


def test_merge_when_one_list_holds_everything():
    """Adversarial: the k best keys all sit in one list, every other head is smaller than its k-th entry."""
    k, n_lists = 32, 148
    lists = np.zeros((n_lists, k), np.uint64)
    lists[77] = np.arange(10_000, 10_000 - k, -1, dtype=np.uint64)
    for i in range(n_lists):
        if i != 77:
            lists[i] = np.arange(100 * i + k, 100 * i, -1, dtype=np.uint64) % np.uint64(9000) + np.uint64(1)
            lists[i] = np.sort(np.unique(lists[i]))[::-1][:k] if np.unique(lists[i]).shape[0] == k else lists[i]
    # make every key unique
    seen, nxt = set(), 20_000
    for i in range(n_lists):
        for e in range(k):
            if int(lists[i, e]) in seen:
                lists[i, e] = 0
            seen.add(int(lists[i, e]))
        nz = np.sort(lists[i][lists[i] != 0])[::-1]
        lists[i] = 0
        lists[i, :nz.shape[0]] = nz
    out, n_out = merge_sorted_lists(lists, k)
    flat = np.sort(lists[lists != 0])[::-1][:k]
    assert n_out == k and np.array_equal(out, flat)


def test_dense_staging_numbers_every_candidate_once():
    """K3's dense staging (csrc/select.cuh): an exclusive prefix over the (clamped) segment lengths numbers the
    candidates; candidate i belongs to the LAST segment whose offset is <= i, found by an 8-step binary search over
    the 257 offsets (empty segments share offsets with their successor)."""
    rng = np.random.default_rng(5)
    for _ in range(300):
        n_seg = int(rng.integers(1, 257))
        cnt = np.zeros(256, np.int64)
        cnt[:n_seg] = rng.choice([0, 0, 1, 2, 37, 256], size=n_seg)
        off = np.zeros(257, np.int64)
        off[1:] = np.cumsum(cnt)
        total = int(off[256])
        seen = np.zeros((256, 257), bool)
        for i in rng.permutation(total)[:400] if total > 400 else range(total):
            lo, hi = 0, 255
            for _step in range(8):
                mid = (lo + hi + 1) >> 1
                if off[mid] <= i:
                    lo = mid
                else:
                    hi = mid - 1
            assert lo < n_seg and off[lo] <= i < off[lo + 1] and 0 <= i - off[lo] < cnt[lo]
            assert not seen[lo, i - off[lo]]
            seen[lo, i - off[lo]] = True
