"""Row-sharded search across the GPUs of one box: one process per GPU (torch.distributed / NCCL).

Rank r holds the contiguous rows [row_base[r], row_base[r+1]) of the collection.  A search is
(1) the local scan + top-k on every rank (K1/K2), (2) ONE all-gather of the nq·k selection keys
(8 bytes each) over NVLink, (3) kernel K3 (`yrb_merge_topk_device`) merging the world·k candidates
per query with the global (score desc, id asc) rule — on every rank, so any rank can answer.
The reference has no counterpart (a single Chroma collection in one process,
utu/rag/storage/implementations/chroma_store.py:41-59); SURVEY.md §8e defines this exchange.

PyTorch here is plumbing only: tensor ownership of the exchange buffers, the current CUDA stream
and the process group.  With the `gloo` backend (CPU tests) the exchange is exercised on host
tensors and the merge runs through `merge_host`, a test-only stand-in for K3.
"""

from __future__ import annotations

import os

import numpy as np
import torch
import torch.distributed as dist

from . import native


def shard_bounds(n_rows: int, world: int) -> list[int]:
    """row_base[0..world]: contiguous shards, the first n_rows % world ranks get one extra row."""
    base, rem = divmod(n_rows, world)
    out = [0]
    for r in range(world):
        out.append(out[-1] + base + (1 if r < rem else 0))
    return out


class ShardedSearcher:
    """Search over a row-sharded collection.  Two CUDA streams: the local scan runs on `stream`, the
    exchange (all-gather + K3 merge) on `comm_stream`, with `depth` result slots, so when searches
    are issued back to back the exchange of search i overlaps the scan of search i+1 (independent
    queries; the latency of one search is unchanged).  With world == 1 there is no exchange and the
    scan kernel writes ids/scores itself (one launch per single-query search)."""

    def __init__(self, index: native.Index | None, row_base: list[int], group=None, depth: int = 2,
                 exchange: str = "p2p", nq_cap: int = 1024, k_cap: int = 128):
        self.index = index
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        assert len(row_base) == self.world + 1
        self.row_base = list(row_base)
        self.depth = max(1, depth)
        self._turn = 0
        self._cuda = index is not None
        if self._cuda:
            self.device = torch.device("cuda", index.device)
            self._base_dev = torch.tensor(self.row_base[:-1], dtype=torch.int64, device=self.device)
            # real (non-NULL) streams: the C ABI treats a NULL stream as "the index's own stream"
            self.stream = torch.cuda.Stream(self.device)
            self.comm_stream = torch.cuda.Stream(self.device, priority=-1) if self.world > 1 else self.stream
            if self.world > 1:
                # room for the exchange kernel beside the persistent scan (YRB_RESERVED_SMS overrides, for experiments)
                index.set_reserved_sms(int(os.environ.get("YRB_RESERVED_SMS", "2")))
        self._bufs = {}
        self._hq = {}
        # exchange: "p2p" = K7, one kernel storing into peers' buffers over NVLink (CUDA IPC); "nccl" = all-gather
        # + merge kernel.  p2p needs peer access between the GPUs; if it cannot be set up on EVERY rank, all ranks
        # use NCCL (both are GPU paths).
        self.exchange = None
        self.exchange_kind = "none" if self.world == 1 else "nccl"
        if self._cuda and self.world > 1 and exchange == "p2p":
            k_cap = min(k_cap, 2048 // self.world)
            ex, ok = None, 1
            try:
                ex = native.Exchange(index.device, self.world, self.rank, nq_cap, k_cap)
                handles = [None] * self.world
                dist.all_gather_object(handles, ex.handles, group=group)
                ex.connect(b"".join(handles))
            except Exception as e:  # noqa: BLE001
                ok = 0
                self._p2p_error = repr(e)
            flag = torch.tensor([ok], device=self.device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
            if int(flag.item()) == 1:
                self.exchange, self.exchange_kind = ex, "p2p"
            elif ex is not None:
                ex.close()

    def _buffers(self, nq: int, k: int):
        key = (nq, k)
        if key not in self._bufs:
            dev = self.device
            slots = []
            for _ in range(self.depth):
                # ids | scores | counts in ONE device buffer (and one pinned mirror): a host search ends with one D2H copy
                packed = torch.empty(nq * k * 12 + nq * 4, dtype=torch.uint8, device=dev)
                slots.append(dict(
                    local=torch.zeros(nq * k, dtype=torch.int64, device=dev),
                    gathered=torch.zeros(self.world * nq * k, dtype=torch.int64, device=dev),
                    packed=packed, host=torch.empty(packed.shape, dtype=torch.uint8, pin_memory=True),
                    ids=packed[: nq * k * 8].view(torch.int64),
                    scores=packed[nq * k * 8: nq * k * 12].view(torch.float32),
                    counts=packed[nq * k * 12:].view(torch.int32),
                    scanned=torch.cuda.Event(), merged=torch.cuda.Event()))
            self._bufs[key] = slots
        return self._bufs[key]

    def search_device(self, dev_queries: torch.Tensor, k: int, dev_mask: torch.Tensor | None = None):
        """queries fp32 [nq, dim] on this rank's GPU (identical on all ranks).  Returns device tensors
        (ids [nq,k] global row ids, scores [nq,k], counts [nq]) that are complete once `comm_stream`
        reaches this point (inputs must be ready before the call).  The returned tensors are reused
        `depth` searches later."""
        nq = dev_queries.shape[0]
        slots = self._buffers(nq, k)
        b = slots[self._turn % self.depth]
        self._turn += 1
        mptr = dev_mask.data_ptr() if dev_mask is not None else 0
        if self.world == 1:
            self.index.search_device_ids(dev_queries.data_ptr(), nq, k, mptr, b["ids"].data_ptr(), b["scores"].data_ptr(),
                                         b["counts"].data_ptr(), self.stream.cuda_stream)
            return b["ids"].view(nq, k), b["scores"].view(nq, k), b["counts"]
        # the slot's previous exchange must have consumed `local` before the scan overwrites it
        self.stream.wait_event(b["merged"])
        self.index.search_device(dev_queries.data_ptr(), nq, k, mptr, b["local"].data_ptr(), self.stream.cuda_stream)
        b["scanned"].record(self.stream)
        with torch.cuda.stream(self.comm_stream):
            self.comm_stream.wait_event(b["scanned"])
            if self.exchange is not None and nq <= self.exchange.nq_cap and k <= self.exchange.k_cap:
                self.exchange.merge(b["local"].data_ptr(), nq, k, self._base_dev.data_ptr(), b["ids"].data_ptr(),
                                    b["scores"].data_ptr(), b["counts"].data_ptr(), self.comm_stream.cuda_stream)
            else:
                dist.all_gather_into_tensor(b["gathered"], b["local"], group=self.group)
                native.merge_topk_device(self.index.device, b["gathered"].data_ptr(), self.world, nq, k,
                                         self._base_dev.data_ptr(), b["ids"].data_ptr(), b["scores"].data_ptr(),
                                         b["counts"].data_ptr(), self.comm_stream.cuda_stream)
            b["merged"].record(self.comm_stream)
        return b["ids"].view(nq, k), b["scores"].view(nq, k), b["counts"]

    def synchronize(self) -> None:
        self.stream.synchronize()
        self.comm_stream.synchronize()

    def search(self, queries: np.ndarray, k: int, dev_mask: torch.Tensor | None = None):
        """Host in / host out: pinned staging both ways, one H2D and one D2H copy."""
        q = np.ascontiguousarray(queries, dtype=np.float32)
        if q.ndim == 1:
            q = q[None, :]
        nq = q.shape[0]
        if self.exchange is not None and nq <= self.exchange.nq_cap and k <= self.exchange.k_cap:
            # the whole search is enqueued from C on one stream (yrb_exchange_search): no torch calls per search
            self.synchronize()   # earlier search_device() calls share the index's scratch
            return self.exchange.search(self.index, q, k, dev_mask.data_ptr() if dev_mask is not None else 0,
                                        self._base_dev.data_ptr())
        slots = self._buffers(nq, k)
        stage = self._hq.get(nq)
        if stage is None:
            stage = self._hq[nq] = (torch.empty(q.shape, dtype=torch.float32, pin_memory=True),
                                    torch.empty(q.shape, dtype=torch.float32, device=self.device))
        stage[0].numpy()[...] = q
        with torch.cuda.stream(self.stream):
            stage[1].copy_(stage[0], non_blocking=True)
        b = slots[self._turn % self.depth]
        self.search_device(stage[1], k, dev_mask)
        with torch.cuda.stream(self.comm_stream):
            b["host"].copy_(b["packed"], non_blocking=True)
        self.comm_stream.synchronize()
        h = b["host"].numpy()
        ids = h[: nq * k * 8].view(np.int64).reshape(nq, k).copy()
        scores = h[nq * k * 8: nq * k * 12].view(np.float32).reshape(nq, k).copy()
        counts = h[nq * k * 12:].view(np.int32).copy()
        return ids, scores, counts


# ------------------------------------------------------------------ host-side exchange (gloo tests)
def exchange_host(local_keys: np.ndarray, group=None) -> np.ndarray:
    """all-gather of one rank's [nq, k] uint64 keys over any backend → [world, nq, k]."""
    world = dist.get_world_size(group)
    t = torch.from_numpy(np.ascontiguousarray(local_keys).view(np.int64))
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t, group=group)
    return np.stack([o.numpy().view(np.uint64) for o in out])


def merge_host(gathered: np.ndarray, row_base: list[int], k: int):
    """What K3 computes, on the host: [world, nq, k] keys → global ids/scores ordered (score desc, id asc).
    Bookkeeping for CPU tests of the sharding logic; the product path is yrb_merge_topk_device."""
    world, nq, _ = gathered.shape
    ids = np.full((nq, k), -1, np.int64)
    scores = np.full((nq, k), -np.inf, np.float32)
    counts = np.zeros(nq, np.int32)
    for q in range(nq):
        cand = []
        for p in range(world):
            rows, sc = native.decode_keys(gathered[p, q])
            ok = rows >= 0
            cand += [(-float(s), int(r) + row_base[p], (gathered[p, q][i] >> np.uint64(32))) for i, (r, s) in
                     enumerate(zip(rows, sc)) if ok[i]]
        cand.sort(key=lambda t: (-int(t[2]), t[1]))
        cand = cand[:k]
        counts[q] = len(cand)
        for j, (ns, gid, _) in enumerate(cand):
            ids[q, j], scores[q, j] = gid, -ns
    return ids, scores, counts
