"""One collection over several shards inside one process (csrc/sharded.cu, csrc/xshard.cuh) against
the oracle and against the single-GPU index.  `devices=[0, 0, 0]` puts three shards on one GPU, so the block-cyclic
row map, the per-shard filters and the cross-shard merge folded into the finishing kernels all run on the driver's
one-GPU box; the same cases run over distinct GPUs when the box has them."""

import asyncio

import numpy as np
import pytest

from oracle import exact_search as ox
from tests.golden_util import GOLDEN, compare_with_golden, golden_chunks
from tests.helpers import check_topk, unit_rows
from youtu_rag_b200 import B200VectorStore, VectorStoreConfig, native
from youtu_rag_b200.where import compile_where, normalize_filters

pytestmark = pytest.mark.gpu


def _device_sets():
    import torch

    sets = [[0, 0, 0]]
    n = torch.cuda.device_count()
    if n >= 2:
        sets.append([0, 1])
    if n >= 8:
        sets.append(list(range(8)))
    return sets


DEVICE_SETS = None


def device_sets():
    global DEVICE_SETS
    if DEVICE_SETS is None:
        DEVICE_SETS = _device_sets()
    return DEVICE_SETS


def run(c):
    return asyncio.run(c)


def model_locate(g, n_shards, block):
    """The row map restated: global row -> (shard, local row)."""
    b = g // block
    return b % n_shards, (b // n_shards) * block + g % block


@pytest.mark.parametrize("dtype", ["bf16", "f32"])
def test_row_map_residency_roundtrip(dtype):
    """append / read_rows / read_raw / append_raw / set_live / truncate / count in GLOBAL row order, with pieces that
    start and end inside blocks; every shard ends up with exactly the rows the arithmetic map assigns to it."""
    n, d, block = 1000, 96, 64
    x = unit_rows(n, d, 3)
    for devs in device_sets():
        ix = native.ShardedIndex(d, "cosine", dtype, devs, block_rows=block)
        for a, b in ((0, 1), (1, 70), (70, 500), (500, 1000)):      # ragged appends
            ix.append(x[a:b])
        assert ix.counts() == (n, n) and ix.info()["block_rows"] == block and ix.info()["n_devices"] == len(devs)
        want = ox.prepare(x, "cosine", dtype)
        ids = np.random.default_rng(0).permutation(n)[:300]
        np.testing.assert_array_equal(ix.read_rows(ids), want[ids])
        np.testing.assert_array_equal(ix.read_rows(np.arange(n)), want)
        # per-shard contents follow the model
        single = native.Index(d, "cosine", dtype, devs[0], n)
        single.append(x)
        for s in range(len(devs)):
            h, _ = ix.shard(s)
            rows_s = sum(1 for g in range(n) if model_locate(g, len(devs), block)[0] == s)
            r, l = native.C.c_int64(), native.C.c_int64()
            native._ck(native.lib().yrb_index_count(native.C.c_void_p(h), native.C.byref(r), native.C.byref(l)))
            assert r.value == rows_s
        raw, sq = ix.read_raw(37, 555)
        raw1, sq1 = single.read_raw(37, 555)
        np.testing.assert_array_equal(raw, raw1)
        np.testing.assert_array_equal(sq, sq1)
        # raw reload into a fresh sharded index is bit-identical
        jx = native.ShardedIndex(d, "cosine", dtype, devs, block_rows=block)
        r_all, s_all = ix.read_raw(0, n)
        jx.append_raw(r_all[:130], s_all[:130])
        jx.append_raw(r_all[130:], s_all[130:])
        np.testing.assert_array_equal(jx.read_rows(np.arange(n)), want)
        # tombstones + truncate
        dead = [0, 63, 64, 65, 999, 500]
        ix.set_live(dead, False)
        assert ix.counts() == (n, n - len(dead))
        ix.truncate(700)
        assert ix.counts() == (700, 700 - 5)
        ix.append(x[700:800])
        assert ix.counts() == (800, 800 - 5)
        np.testing.assert_array_equal(ix.read_rows(np.arange(800)), want[:800])
        ix.clear()
        assert ix.counts() == (0, 0)
        for o in (ix, jx, single):
            o.close()


def test_device_side_appends_cross_gpus():
    """Rows generated on one GPU and appended with append_device: pieces for other shards travel over NVLink and
    must be ingested only after they have landed (bench.py's sharded-store leg builds its corpus this way)."""
    import torch

    n, d, block = 200_000, 256, 1024
    for devs in device_sets():
        ix = native.ShardedIndex(d, "cosine", "bf16", devs, block_rows=block)
        single = native.Index(d, "cosine", "bf16", devs[0], n)
        gen = torch.Generator(device=f"cuda:{devs[0]}")
        for b in range(0, n, 50_000):
            gen.manual_seed(b)
            blk = torch.randn(50_000, d, device=f"cuda:{devs[0]}", generator=gen)
            torch.cuda.synchronize()
            ix.append_device(blk.data_ptr(), 50_000, src_device=devs[0])
            single.append_device(blk.data_ptr(), 50_000)
        a, sa = ix.read_raw(0, n)
        b_, sb = single.read_raw(0, n)
        np.testing.assert_array_equal(a, b_)
        np.testing.assert_array_equal(sa, sb)
        ix.close()
        single.close()


def _check(ix, rows_stored, q, k, metric, dtype, oracle_mask=None, **kw):
    ids, scores, counts = ix.search(q, k, **kw)
    qq = np.atleast_2d(q)
    for j in range(qq.shape[0]):
        n = int(counts[j])
        check_topk(ids[j, :n], scores[j, :n], rows_stored, ox.prepare(qq[j], metric, dtype)[0], k, metric, dtype, oracle_mask)
        assert (ids[j, n:] == -1).all()


@pytest.mark.parametrize("metric,dtype", [("cosine", "bf16"), ("dot", "bf16"), ("euclidean", "bf16"), ("cosine", "f32"), ("euclidean", "f32")])
def test_sharded_search_matches_oracle(metric, dtype):
    """Every kernel family behind the cross-shard finish: K1 fused (k <= 32), K1 with the block selection (k = 100),
    K2 one-CTA and pair (+ chunks), K1Q (fp32 batches), K6 (k > 128), with a host bitmask, tombstones, k > rows and
    shards that hold no rows."""
    n, d, block = 5000, 128, 256
    x = unit_rows(n, d, 11) * (1.7 if metric == "dot" else 1.0)   # (euclidean scores of longer rows leave the 2e-6 tie band of check_topk)
    rng = np.random.default_rng(5)
    q1 = rng.standard_normal(d).astype(np.float32)
    qb = rng.standard_normal((20, d)).astype(np.float32)
    qbig = rng.standard_normal((300, d)).astype(np.float32)
    for devs in device_sets():
        ix = native.ShardedIndex(d, metric, dtype, devs, block_rows=block)
        ix.append(x)
        stored = ix.read_rows(np.arange(n))
        _check(ix, stored, q1, 10, metric, dtype)                       # K1, fused rank merge
        _check(ix, stored, q1, 100, metric, dtype)                      # K1, block selection in the last CTA
        _check(ix, stored, qb, 10, metric, dtype)                       # K2 (bf16) / K1Q (f32)
        _check(ix, stored, qb, 100, metric, dtype)                      # K2 / K1 loop
        if dtype == "bf16":
            _check(ix, stored, qbig, 10, metric, dtype)                 # pair kernel + a second chunk on the one-CTA kernel
        _check(ix, stored, q1, 200, metric, dtype)                      # K6 + stand-alone finish
        m = rng.random(n) < 0.1
        _check(ix, stored, q1, 10, metric, dtype, m, mask=ox.pack_mask(m))
        _check(ix, stored, qb, 10, metric, dtype, m, mask=ox.pack_mask(m))
        dead = rng.permutation(n)[:400]
        ix.set_live(dead, False)
        live = np.ones(n, bool)
        live[dead] = False
        _check(ix, stored, q1, 10, metric, dtype, live)
        _check(ix, stored, qb, 10, metric, dtype, live & m, mask=ox.pack_mask(m))
        ix.close()
        # a collection smaller than one block: only shard 0 holds rows; k larger than the collection
        small = native.ShardedIndex(d, metric, dtype, devs, block_rows=block)
        small.append(x[:40])
        _check(small, small.read_rows(np.arange(40)), q1, 64, metric, dtype)
        _check(small, small.read_rows(np.arange(40)), qb, 64, metric, dtype)
        small.close()


def test_sharded_equals_single_gpu_bitwise():
    """Same rows, same queries: the sharded index returns exactly the ids AND score bits of the one-GPU index (the
    per-row arithmetic does not depend on where a row lives; ties break on the GLOBAL row id)."""
    n, d = 20000, 1024
    x = unit_rows(n, d, 21)
    x[7000] = x[123]
    x[15000] = x[123]            # exact ties across shards
    rng = np.random.default_rng(9)
    single = native.Index(d, "cosine", "bf16", 0, n)
    single.append(x)
    for devs in device_sets():
        ix = native.ShardedIndex(d, "cosine", "bf16", devs, block_rows=1024)
        ix.append(x)
        for nq, k in ((1, 10), (1, 128), (16, 10), (256, 100)):
            q = rng.standard_normal((nq, d)).astype(np.float32)
            q[0] = x[123]
            a, b = ix.search(q, k), single.search(q, k)
            np.testing.assert_array_equal(a[0], b[0])
            np.testing.assert_array_equal(a[1].view(np.uint32), b[1].view(np.uint32))
            np.testing.assert_array_equal(a[2], b[2])
        ix.close()
    single.close()


def _store(devs, metric="cosine", dtype="f32", **extra):
    cfg = VectorStoreConfig(backend="b200", collection_name="col_sharded", distance_metric=metric,
                            index_params={"storage_dtype": dtype, "devices": devs, "shard_block_rows": 64, **extra})
    return B200VectorStore(cfg)


@pytest.mark.parametrize("metric", ["cosine", "dot", "euclidean"])
def test_sharded_store_reproduces_reference_glue(metric):
    """The golden suite of the drop-in store (outputs of the reference's unmodified ChromaVectorStore glue) with
    index_params.devices set: where → K4 per shard, global ids, mutations routed to shards."""
    for devs in device_sets():
        s = _store(devs, metric)
        run(s.add_chunks(golden_chunks()))
        assert isinstance(s.index, native.ShardedIndex)
        for rec in GOLDEN["chroma"]:
            if rec["metric"] != metric:
                continue
            q = GOLDEN["queries"][rec["query"]]
            if "error" in rec:
                with pytest.raises(ValueError):
                    run(s.search(q, rec["top_k"], rec["filters"]))
            else:
                compare_with_golden(run(s.search(q, rec["top_k"], rec["filters"])), rec["results"], tol=1e-5)
        s.close()


def test_sharded_store_mutations_and_batches():
    for devs in device_sets():
        s = _store(devs)
        run(s.add_chunks(golden_chunks()))
        mut = GOLDEN["chroma_mutations"]
        assert run(s.count()) == mut["count0"]
        assert run(s.delete_by_document_id("doc2")) == mut["deleted_doc2"]
        assert run(s.delete_by_metadata({"source": "file1.pdf", "index_type": "index_summary"})) == mut["deleted_meta"]
        run(s.delete(["doc0_chunk_0", "nope"]))
        assert run(s.count()) == mut["count1"]
        g = run(s.get_by_id("doc0_chunk_2"))
        assert {"id": g.id, "document_id": g.document_id, "chunk_index": g.chunk_index, "metadata": g.metadata} == mut["get"]
        compare_with_golden(run(s.search(GOLDEN["queries"][0], top_k=5)), mut["search_after"], tol=1e-5)
        # a batch with one filter per query equals the single searches
        filters = [None, {"source": "file0.pdf"}, {"chunk_index": {"$gte": 2}}]
        qs = GOLDEN["queries"][:3]
        batch = run(s.search_batch(qs, top_k=4, filters=filters))
        for q, f, got in zip(qs, filters, batch):
            want = run(s.search(q, 4, f))
            assert [(c.id, sc) for c, sc in got] == [(c.id, sc) for c, sc in want]
        run(s.clear())
        assert run(s.count()) == 0 and run(s.search(GOLDEN["queries"][0], top_k=5)) == []
        s.close()


def test_sharded_store_persistence_roundtrip(tmp_path):
    """A collection written by a sharded store reloads bit-identically — also into a store with a different device
    list (the on-disk form is in global row order)."""
    ch = golden_chunks()
    for devs in device_sets():
        cfg = VectorStoreConfig(collection_name="col_ps", persist_directory=str(tmp_path / f"d{len(devs)}"), distance_metric="cosine",
                                index_params={"persist": True, "devices": devs, "shard_block_rows": 64})
        a = B200VectorStore(cfg)
        run(a.add_chunks(ch[:40]))
        run(a.add_chunks(ch[40:]))
        run(a.delete_by_document_id("doc2"))
        want = [run(a.search(q, 7, {"source": {"$in": ["file0.pdf", "file1.pdf"]}})) for q in GOLDEN["queries"]]
        a.close()
        cfg1 = VectorStoreConfig(collection_name="col_ps", persist_directory=str(tmp_path / f"d{len(devs)}"), distance_metric="cosine",
                                 index_params={"persist": True})
        for c in (cfg, cfg1):
            b = B200VectorStore(c)
            for q, w in zip(GOLDEN["queries"], want):
                got = run(b.search(q, 7, {"source": {"$in": ["file0.pdf", "file1.pdf"]}}))
                assert [(x.id, s, x.metadata) for x, s in got] == [(x.id, s, x.metadata) for x, s in w]
            b.close()


def test_back_to_back_searches_of_changing_shape_and_two_threads():
    """The ticket / gather protocol of the cross-shard finish over many searches whose batch size and k change from one
    to the next (the per-query tickets must be back at zero every time), also issued from two threads at once (the
    entry point serialises them): every answer equals the one-GPU index bit for bit."""
    import threading

    n, d = 40_000, 128
    x = unit_rows(n, d, 51)
    single = native.Index(d, "cosine", "bf16", 0, n)
    single.append(x)
    rng = np.random.default_rng(6)
    shapes = [(1, 10), (3, 100), (300, 5), (1, 128), (17, 32), (2, 200), (64, 10)]
    qs = {s: rng.standard_normal((s[0], d)).astype(np.float32) for s in shapes}
    want = {s: single.search(qs[s], s[1]) for s in shapes}
    for devs in device_sets():
        ix = native.ShardedIndex(d, "cosine", "bf16", devs, block_rows=512)
        ix.append(x)

        def run_many(seed, errors):
            order = np.random.default_rng(seed).integers(0, len(shapes), 60)
            for o in order:
                s = shapes[o]
                got = ix.search(qs[s], s[1])
                if not (np.array_equal(got[0], want[s][0]) and np.array_equal(got[1].view(np.uint32), want[s][1].view(np.uint32))
                        and np.array_equal(got[2], want[s][2])):
                    errors.append(s)

        errs: list = []
        run_many(0, errs)
        threads = [threading.Thread(target=run_many, args=(t + 1, errs)) for t in range(2)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        assert not errs, errs
        ix.close()
    single.close()


def test_entry_points_leave_the_callers_device_alone():
    """The library switches devices internally (a sharded collection walks over several); a caller sharing the thread
    with torch must find its current device unchanged after every call."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs to tell devices apart")
    torch.cuda.set_device(1)
    x = unit_rows(3000, 64, 1)
    ix = native.ShardedIndex(64, "cosine", "bf16", [0, 1], block_rows=256)
    ix.append(x)
    one = native.Index(64, "cosine", "bf16", 0, 3000)
    one.append(x)
    for index in (ix, one):
        index.search(x[:3], 5)
        index.set_live([1, 2], False)
        index.read_rows([0, 700])
        assert torch.cuda.current_device() == 1
    ix.close()
    one.close()
    assert torch.cuda.current_device() == 1
    torch.cuda.set_device(0)
