#!/usr/bin/env python
"""bench.py — BASELINE.json's metric (queries/sec + p50 latency of exact top-k) on synthetic data.

    python bench.py --gpus N --steps K --warmup W [--workload NAME] [--impl reference]

A "step" is one search: one pass of the hot path over the resident corpus for one batch of queries.

Default workload (every N) = the north-star target: **10M x 1024 bf16 corpus, single query, top-10, cosine**
(`t10m`; 20.5 GB, fits one B200).  With N > 1 (torchrun, one rank per GPU) the SAME corpus is row-sharded over the
ranks ("strong" scaling: 5M / 2.5M / 1.25M-row shards): local scan + top-k, then the top-k merge collective (K7: one
kernel over NVLink peer memory; `--exchange nccl` = all-gather + merge kernel).

The same JSON line carries, under `also`, short runs of the other BASELINE.json configurations on the same box:
  c4     10M x 1024 + 10 % metadata mask, single query (configs[3]; every N)
  t10mb  10M x 1024, 256-query batches, top-100 (the batched scaling case of north_star; every N)
  t10mq  10M x 1024, 1024-query batches, top-10 (C5's batch shape on the resident corpus: ONE K2 launch sweeps four
         256-query chunks per row tile, the corpus is read once; every N)
  c3     1M x 1024, 256-query batch, top-100 (configs[2]; N = 1 only — a one-GPU configuration)
  sharded_store  (N > 1, rank 0) the drop-in store's in-process form: ONE process driving all N GPUs through
         `yrb_sharded_search` (host buffers in, merged result in pinned host memory out)
each with its own `ms_per_step` / `roofline`.  `--workload NAME` runs one named workload alone (no `also`).

`value`     whole-job queries/s with queries already resident in HBM (device-timed, max over ranks)
`e2e`       same metric through the host-buffer call (H2D query + D2H result inside the timed region)
`roofline`  algorithmic bytes (or flops) of the dominant kernel / its CUDA-event duration (events recorded inside the
            library on the launching stream around that kernel only)
`parity`    step-0 results (ids + scores) of every workload against the sliced oracle: every rank scores ITS rows on
            the CPU (fp32 BLAS shortlist → exact fp64 re-score with oracle/exact_search.py), rank 0 merges the
            shortlists with the (score desc, id asc) rule and compares; `recall_at_k` beside it
`cpu_baseline`  the oracle port of the reference's exact CPU path (FAISSVectorStore semantics, numpy fp32 BLAS) timed
            on this box's host cores, reported beside — not a target.
`--impl reference` prints that CPU arm as the main line (the reference's engines, chromadb / faiss-cpu, are not
installable offline; see DESIGN.md §6 — when `faiss` imports, the arm runs faiss.IndexFlatIP itself and says so).
"""

from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time
from pathlib import Path

if "reference" in sys.argv and int(os.environ.get("RANK", "0")) == 0:
    # torchrun pins OMP_NUM_THREADS=1 for multi-rank launches; the reference arm is a CPU measurement on rank 0
    # alone and must use every host core it can, as it does when launched without torchrun (BLAS reads these
    # variables when numpy is imported, hence before the import)
    _cores = str(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_v] = _cores

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

BLOCK_ROWS = 125_000  # corpus is generated in blocks; block b uses seed b (reproducible per shard)
WORKLOADS = {
    #        rows        dim  nq   k   mask selectivity
    "c2": (1_000_000, 1024, 1, 10, None),
    "t10m": (10_000_000, 1024, 1, 10, None),         # north-star target: single query over 10M x 1024
    "c3": (1_000_000, 1024, 256, 100, None),
    "c4": (10_000_000, 1024, 1, 10, 0.10),
    "c4b": (10_000_000, 1024, 256, 10, 0.10),
    "c5": (100_000_000, 768, 1024, 10, None),
    "c5s": (12_500_000, 768, 1024, 10, None),        # one C5 shard (12.5M x 768 per GPU of the 8-GPU config)
    "tiny": (250_000, 1024, 1, 10, None),
    "t10mb": (10_000_000, 1024, 256, 100, None),     # 256-query batches over the 10M corpus (batched scaling case)
    "c3s": (125_000, 1024, 256, 100, None),          # one C3 shard of an 8-GPU run, for the fixed-cost breakdown
    "t10mbs": (1_250_000, 1024, 256, 100, None),     # one t10mb shard of an 8-GPU run
    "t10mq": (10_000_000, 1024, 1024, 10, None),     # 1024-query batches over the 10M corpus (C5's batch shape: one K2 launch, one corpus pass)
}
DEFAULT_WORKLOAD = "t10m"
N_QUERY_SETS = 64
PARITY_QUERIES = 4   # queries of a batch checked against the sliced oracle
ALSO_KEYS = ("value", "unit", "ms_per_step", "steps", "latency_ms", "e2e", "roofline", "gpu_launches", "config")


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


def host_queries(dim: int, nq: int) -> np.ndarray:
    rng = np.random.default_rng(1)
    q = rng.standard_normal((N_QUERY_SETS, nq, dim)).astype(np.float32)
    q /= np.linalg.norm(q, axis=2, keepdims=True)
    return q


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons of one GPU through NVML, back to back (no sleep: NVML's own call time
    sets the rate, a few kHz), from before the warm-up until after the timed region."""

    def __init__(self, device: int):
        super().__init__(daemon=True)
        self.device, self.samples, self.reasons, self.max_mhz = device, [], set(), None
        self._halt = threading.Event()
        self._mark = 0

    def run(self):
        try:
            import pynvml as nv

            nv.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(visible.split(",")[self.device]) if visible and visible.split(",")[0].isdigit() else self.device
            h = nv.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {
                getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
                getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
                getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
                getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
                getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
            }
            get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or getattr(
                nv, "nvmlDeviceGetCurrentClocksThrottleReasons")
            n = 0
            while not self._halt.is_set():
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                n += 1
                if n % 8 == 0:
                    r = get_reasons(h)
                    for bit, name in names.items():
                        if r & bit:
                            self.reasons.add(name)
                time.sleep(0)
        except Exception as e:  # noqa: BLE001
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def mark(self):
        """The timed region starts here: `sm_mhz` is the median of the samples taken from now on."""
        self._mark = len(self.samples)

    def finish(self, seconds: float | None = None) -> dict:
        self._halt.set()
        self.join(timeout=2)
        timed = self.samples[self._mark:] or self.samples
        out = {"sm_mhz": statistics.median(timed) if timed else None, "sm_max_mhz": self.max_mhz,
               "reasons": sorted(self.reasons), "samples": len(timed), "samples_incl_warmup": len(self.samples)}
        if seconds and timed:
            out["sample_hz"] = round(len(timed) / seconds)
        return out


# ----------------------------------------------------------------------------- CPU arm
def cpu_corpus(rows: int, dim: int) -> np.ndarray:
    from oracle import exact_search as ox

    out = np.empty((rows, dim), np.float32)
    for b in range((rows + BLOCK_ROWS - 1) // BLOCK_ROWS):
        a, e = b * BLOCK_ROWS, min(rows, (b + 1) * BLOCK_ROWS)
        out[a:e] = ox.l2_normalize(np.random.default_rng(b).standard_normal((e - a, dim), dtype=np.float32))
    return out


def _faiss_or_none():
    try:
        import faiss  # noqa: F401  (absent from this image; used when a box has the reference's pinned wheel)

        return faiss
    except Exception:  # noqa: BLE001
        return None


def cpu_arm(corpus: np.ndarray, queries: np.ndarray, k: int, steps: int, warmup: int, budget_s: float = 25.0):
    """FAISSVectorStore.search semantics (faiss_store.py:143-199): fp32 scan + partition + ordered top-k, looped over
    the step's queries like base_retriever.py:95-99.  With `faiss` importable the scan IS faiss.IndexFlatIP (kind
    "reference"); otherwise the oracle's restatement (kind "port").  Returns (qps, per-step seconds, kind)."""
    from oracle import exact_search as ox

    faiss = _faiss_or_none()
    if faiss is not None:
        index = faiss.IndexFlatIP(corpus.shape[1])
        index.add(corpus)

        def one(q):
            qq = np.ascontiguousarray(q[None, :], np.float32)
            faiss.normalize_L2(qq)
            index.search(qq, k)
        kind = "reference"
    else:
        def one(q):
            ox.faiss_flat_search(corpus, q, k, "cosine")
        kind = "port"

    def step(i):
        for q in queries[i % queries.shape[0]]:
            one(q)

    for i in range(warmup):
        step(i)
    times = []
    t_all = time.perf_counter()
    for i in range(steps):
        t0 = time.perf_counter()
        step(i)
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_all > budget_s:
            break
    nq = queries.shape[1]
    return nq * len(times) / sum(times), times, kind


def blas_threads() -> int:
    try:
        from threadpoolctl import threadpool_info

        return max((p.get("num_threads", 1) for p in threadpool_info() if p.get("user_api") == "blas"), default=1)
    except Exception:  # noqa: BLE001
        return os.cpu_count() or 1


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    name = args.workload or DEFAULT_WORKLOAD
    rows, dim, nq, k, sel = WORKLOADS[name]
    # bounded sample of the workload: at most 1M rows of the corpus and at most 4 queries per step
    s_rows, s_nq = min(rows, 1_000_000), min(nq, 4)
    corpus = cpu_corpus(s_rows, dim)
    q = host_queries(dim, nq)[:, :s_nq]
    qps, times, kind = cpu_arm(corpus, q, k, args.steps, args.warmup, budget_s=120.0)
    # scale to the full workload: a step scans `rows` rows for `nq` queries; the scan is linear in both
    scale = (s_rows / rows)
    value = qps * scale
    sample = f"{s_rows}x{dim} fp32 rows, {s_nq} of {nq} queries per step, {len(times)} steps; scaled x{scale:g} to {rows} rows"
    line = {
        "impl": "reference", "metric": "queries_per_sec", "value": value, "unit": "queries/s", "n_gpus": args.gpus,
        "steps": len(times), "warmup": args.warmup, "ms_per_step": 1e3 * nq / value, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(name, args.gpus),
        "cpu_baseline": {"value": value, "unit": "queries/s", "cores": blas_threads(), "kind": kind, "sample": sample,
                         "p50_ms": 1e3 * statistics.median(times) / s_nq / scale},
        "e2e": {"value": value, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": ("faiss.IndexFlatIP (the engine of the reference's FAISSVectorStore) on the host cores" if kind == "reference" else
                 "reference engines (chromadb 1.3.4 HNSW / faiss-cpu 1.12.0) are not installable offline; this is the "
                 "oracle port of FAISSVectorStore.search (exact, numpy fp32 BLAS) on the host cores"),
    }
    print(json.dumps(line))


def workload_config(name: str, gpus: int) -> dict:
    rows, dim, nq, k, sel = WORKLOADS[name]
    return {"workload": f"{name}: {rows}x{dim} bf16 corpus, {nq}-query batch, top-{k}, cosine"
                        + (f", {int(sel * 100)}% metadata mask" if sel else ""),
            "rows": rows, "dim": dim, "queries_per_step": nq, "k": k, "mask_selectivity": sel,
            "sharding": f"rows/{gpus}" if gpus > 1 else "none",
            "l2_policy": "corpus shard larger than the 126 MB L2; 64 distinct query sets cycled"}


# ----------------------------------------------------------------------------- GPU arm
class Env:
    """Process-group plumbing of one rank."""

    def __init__(self, gpus: int):
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != gpus:
            raise SystemExit(f"--gpus {gpus} but WORLD_SIZE={self.world}: launch with torch.distributed.run --nproc-per-node {gpus}")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        self.cpu_group = None
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
            self.cpu_group = dist.new_group(backend="gloo")

    def host_wait(self):
        """Rendezvous that leaves every GPU idle: an NCCL barrier parks a spinning kernel on the waiting ranks' GPUs,
        which would time-slice against the one process that drives all GPUs in the sharded-store leg."""
        self.torch.cuda.synchronize(self.dev)
        if self.world > 1:
            self.dist.barrier(group=self.cpu_group)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def max_over_ranks(self, x: float) -> float:
        t = self.torch.tensor([x], device=self.dev, dtype=self.torch.float64)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def gather_objects(self, obj):
        if self.world == 1:
            return [obj]
        out = [None] * self.world if self.rank == 0 else None
        self.dist.gather_object(obj, out, dst=0)
        return out


class Corpus:
    """This rank's contiguous slice of a synthetic corpus, resident in a `native.Index`."""

    def __init__(self, env: Env, rows: int, dim: int, path: int = 0):
        from youtu_rag_b200 import native

        torch = env.torch
        self.env, self.rows, self.dim = env, rows, dim
        n_blocks = (rows + BLOCK_ROWS - 1) // BLOCK_ROWS
        if n_blocks % env.world:
            raise SystemExit(f"{n_blocks} corpus blocks do not divide over {env.world} ranks")
        per = n_blocks // env.world
        self.bounds = [min(rows, b * per * BLOCK_ROWS) for b in range(env.world + 1)]
        self.bounds[-1] = rows
        self.my_blocks = range(env.rank * per, (env.rank + 1) * per)
        self.base = self.bounds[env.rank]
        self.n_local = self.bounds[env.rank + 1] - self.base
        self.index = native.Index(dim, "cosine", "bf16", env.local, self.n_local)
        gen = torch.Generator(device=env.dev)
        for b in self.my_blocks:
            n_b = min(BLOCK_ROWS, rows - b * BLOCK_ROWS)
            gen.manual_seed(b)
            blk = torch.randn(n_b, dim, device=env.dev, generator=gen)
            torch.cuda.synchronize(env.dev)
            self.index.append_device(blk.data_ptr(), n_b)
            del blk
        assert self.index.rows == self.n_local
        if path:
            self.index.set_path(path)
        self._masks = {}

    def mask(self, sel: float):
        """(device words, local bool array on the host, passing rows) of the Bernoulli(sel) metadata mask."""
        if sel not in self._masks:
            torch, env = self.env.torch, self.env
            gen = torch.Generator(device=env.dev)
            bits = []
            for b in self.my_blocks:
                n_b = min(BLOCK_ROWS, self.rows - b * BLOCK_ROWS)
                gen.manual_seed(2_000_000 + b)
                bits.append(torch.rand(n_b, device=env.dev, generator=gen) < sel)
            bits = torch.cat(bits)
            host = bits.cpu().numpy()
            pad = (-bits.numel()) % 64
            w = torch.cat([bits, torch.zeros(pad, dtype=torch.bool, device=env.dev)]).view(-1, 32).to(torch.int64)
            words = (w << torch.arange(32, device=env.dev, dtype=torch.int64)).sum(1).to(torch.int32)
            self._masks[sel] = (words, host, int(host.sum()))
        return self._masks[sel]


def measure(env: Env, corpus: Corpus, searcher, name: str, steps: int, warmup: int, args, with_cpu: bool) -> tuple[dict, dict]:
    """Times one workload on an already resident corpus.  Returns (result fields, step-0 host results for parity)."""
    torch = env.torch
    rows, dim, nq, k, sel = WORKLOADS[name]
    index, world, n_local = corpus.index, env.world, corpus.n_local
    dev_mask, host_mask, n_pass = corpus.mask(sel) if sel else (None, None, n_local)
    hq = host_queries(dim, nq)
    dq = torch.from_numpy(hq).to(env.dev)
    st = searcher.stream  # every kernel of a step is launched on this stream

    def step_device(i):
        return searcher.search_device(dq[i % N_QUERY_SETS], k, dev_mask)

    def step_host(i):
        if world == 1 and dev_mask is None:
            return index.search(hq[i % N_QUERY_SETS], k)
        return searcher.search(hq[i % N_QUERY_SETS], k, dev_mask)

    # ---------------- step 0 through the host path: the result the parity check looks at
    first = step_host(0)
    sampler = ClockSampler(env.local)
    sampler.start()
    # ---------------- value: device-resident inputs, K steps timed with CUDA events
    for i in range(warmup):
        step_device(i)
    env.barrier()
    index.profile(True)
    # device-side rendezvous: one untimed step whose exchange waits for every peer ON THE DEVICE, and the scan
    # stream waits for it — the timed region then starts within microseconds on all ranks (a host barrier alone
    # leaves the ranks' launch skew inside a short timed region)
    step_device(warmup)
    st.wait_stream(searcher.comm_stream)
    index.profile_read()
    l0 = index.launches()
    # steps are issued back to back: scan on `st`, exchange + merge on the searcher's comm stream, so the
    # exchange of step i overlaps the scan of step i+1 (independent queries).  The timed region starts on the
    # scan stream and ends when the LAST step's merged result is complete on the comm stream.
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.mark()
    t_host0 = time.perf_counter()
    e0.record(st)
    for i in range(steps):
        step_device(warmup + 1 + i)
    e1.record(searcher.comm_stream)
    env.barrier()
    clocks = sampler.finish(time.perf_counter() - t_host0)
    ms_total = e0.elapsed_time(e1)
    launches = index.launches() - l0 + (steps if world > 1 else 0)  # + the exchange/merge kernel per step
    kern_ms, kern_n = index.profile_read()
    index.profile(False)
    ms_total = env.max_over_ranks(ms_total)
    ms_step = ms_total / steps
    value = nq * 1e3 / ms_step

    # ---------------- latency: per-step synchronised (p50 / p99)
    lat = []
    for i in range(min(steps, 500)):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(env.dev)
        a.record(st)
        step_device(i)
        b.record(searcher.comm_stream)
        b.synchronize()
        lat.append(a.elapsed_time(b))
    lat.sort()

    # ---------------- e2e: host buffers (H2D query, D2H ids+scores inside)
    env.barrier()
    e2e_steps = min(steps, 500)
    for i in range(max(3, warmup // 4)):
        step_host(i)
    env.barrier()
    host_lat = []
    t0 = time.perf_counter()
    t_prev = t0
    for i in range(e2e_steps):
        step_host(i)
        t_now = time.perf_counter()
        host_lat.append(t_now - t_prev)  # host call -> ids/scores on the host (SURVEY.md §8d's latency)
        t_prev = t_now
    torch.cuda.synchronize(env.dev)
    e2e_s = env.max_over_ranks(time.perf_counter() - t0)
    host_lat.sort()
    e2e_value = nq * e2e_steps / e2e_s

    # ---------------- roofline of the dominant kernel
    hbm, tf_burst, tf_sust, peak_src = peaks()
    frac_rows = (n_pass / n_local) if sel else 1.0
    kern_avg_ms = kern_ms / max(1, kern_n)
    launches_per_step = kern_n / steps
    if nq < 2 or args.path == 1:
        # K1: one launch scans the shard once for one query
        algo_bytes = frac_rows * n_local * dim * 2 + (n_local / 8 if sel else 0) + dim * 4 + (148 if world == 1 else 146) * k * 8
        roof = {"bound": "hbm", "kernel": "k1_scan_topk", "achieved": algo_bytes / (kern_avg_ms * 1e-3) / 1e9,
                "peak": hbm, "unit": "GB/s", "peak_source": peak_src, "bytes_per_launch": algo_bytes,
                "avg_launch_ms": kern_avg_ms, "launches_per_step": launches_per_step}
        roof.update(traffic_of("k1", algo_bytes, sel))
    else:
        # the timed launch is K2's GEMM over every tile of the shard for the first (up to) 1024 queries, sampling included
        # (a shared mask passing <= 25 % of the rows is compacted first by K8: the GEMM then covers the passing rows)
        rows_b = n_pass if (sel and k <= n_pass <= n_local // 4) else n_local
        # one launch takes up to 1024 queries (four 256-query chunks per row tile: the corpus is read once)
        flops = 2.0 * min(nq, 1024) * rows_b * dim
        # MEASURED_PEAKS.json holds two cuBLAS figures: the burst one for a kernel timed alone, the sustained one for a
        # kernel inside a long step.  A launch of a millisecond or more, issued back to back, runs at the sustained
        # clocks (sw_power_cap): it is held against the sustained figure; both are in the line.
        sustained = kern_avg_ms >= 1.0
        roof = {"bound": "tensor", "kernel": "k2_gemm_topk_pair" if nq > 128 else "k2_gemm_topk",
                "achieved": flops / (kern_avg_ms * 1e-3) / 1e12,
                "peak": tf_sust if sustained else tf_burst, "peak_kind": "sustained" if sustained else "burst",
                "unit": "TFLOP/s", "peak_source": peak_src, "peak_burst": tf_burst, "peak_sustained": tf_sust,
                "frac_of_burst": flops / (kern_avg_ms * 1e-3) / 1e12 / tf_burst, "flops_per_launch": flops,
                "avg_launch_ms": kern_avg_ms, "launches_per_step": launches_per_step,
                "hbm_gbs_same_kernel": (rows_b * dim * 2) / (kern_avg_ms * 1e-3) / 1e9,
                "rows_in_gemm": rows_b,
                "note": "dense GEMM over rows_in_gemm rows (all rows, or the K8-compacted passing rows of a selective shared mask)"}
        # the multi-chunk launch (more than 256 queries) has its own capture: its claim is that the rows still come from
        # HBM once while four chunks of queries use them
        roof.update(traffic_of("k2mc" if nq > 256 else "k2", rows_b * dim * 2.0, sel))
    roof["frac"] = roof["achieved"] / roof["peak"]

    # ---------------- CPU baseline beside it (rank 0, N=1 only; bounded sample)
    cpu = None
    if with_cpu and world == 1 and env.rank == 0 and not args.no_cpu:
        s_rows = min(n_local, 1_000_000)
        sample_rows = index.read_rows(np.arange(s_rows))          # the same stored rows, decoded to fp32
        s_nq = min(nq, 4)
        qps, times, kind = cpu_arm(sample_rows, hq[:, :s_nq], k, 200, 10, budget_s=20.0)
        scale = s_rows / n_local
        cpu = {"value": qps * scale, "unit": "queries/s", "cores": blas_threads(), "kind": kind,
               "sample": f"{s_rows}x{dim} fp32 rows (read back from the index), {s_nq} of {nq} queries per step, "
                         f"{len(times)} steps after 10 warm-ups; scaled x{scale:g} to the {n_local}-row workload",
               "p50_ms": 1e3 * statistics.median(times) / s_nq / scale, "host_cpus": os.cpu_count()}
        del sample_rows

    res = {
        "metric": "queries_per_sec", "value": value, "unit": "queries/s", "n_gpus": world, "steps": steps,
        "warmup": warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": workload_config(name, world),
        "latency_ms": {"p50": lat[len(lat) // 2], "p99": lat[min(len(lat) - 1, int(len(lat) * 0.99))],
                       "min": lat[0], "n": len(lat)},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "queries/s", "h2d_bytes_per_step": nq * dim * 4,
                "d2h_bytes_per_step": nq * k * 12 + nq * 4, "ms_per_step": 1e3 * e2e_s / e2e_steps,
                "p50_ms": 1e3 * host_lat[len(host_lat) // 2], "p99_ms": 1e3 * host_lat[min(len(host_lat) - 1, int(len(host_lat) * 0.99))],
                "steps": e2e_steps, "api": "yrb_index_search (C ABI, host buffers)" if world == 1 and dev_mask is None
                else "ShardedSearcher.search (host buffers)"},
        "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu, "exchange": searcher.exchange_kind,
    }
    probe = {"name": name, "k": k, "queries": hq[0][:PARITY_QUERIES], "mask": host_mask,
             "ids": first[0][:PARITY_QUERIES], "scores": first[1][:PARITY_QUERIES], "counts": first[2][:PARITY_QUERIES]}
    return res, probe


def traffic_of(kernel: str, algo_bytes: float, sel) -> dict:
    """`roofline.traffic`: dram bytes per launch.  ncu cannot run inside the timed bench, so the figure is the
    measured dram bytes PER ALGORITHMIC BYTE of the same kernel from the committed `ncu --set full` capture
    (profiles/traffic.json: {"k1": {"ratio": …, "source": …}, …}) applied to this run's algorithmic bytes — or null."""
    tp = ROOT / "profiles" / "traffic.json"
    if tp.exists():
        t = json.loads(tp.read_text()).get(kernel + ("_masked" if sel else ""))
        if isinstance(t, dict) and "ratio" in t:
            return {"traffic": t["ratio"] * algo_bytes, "traffic_source": t.get("source")}
    return {"traffic": None}


# ----------------------------------------------------------------------------- parity (sliced oracle)
def local_shortlists(corpus: Corpus, probes: list[dict]) -> list[dict]:
    """This rank's contribution to the oracle: for every probe query, its best rows of THIS shard — shortlisted with
    an fp32 BLAS scan over the stored rows (read back bit-exactly), then re-scored exactly (fp64 over the stored
    operands, oracle/exact_search.py).  The shortlist holds 4k rows per 250k-row block, far wider than fp32 noise."""
    from oracle import exact_search as ox

    index, out = corpus.index, []
    qp = [np.stack([ox.prepare(q, "cosine", "bf16")[0] for q in p["queries"]]) for p in probes]   # [nqp, dim] each
    qall = np.concatenate(qp).astype(np.float32)
    spans = np.cumsum([0] + [x.shape[0] for x in qp])
    cand = [[[] for _ in range(x.shape[0])] for x in qp]
    blk = 250_000
    for a in range(0, corpus.n_local, blk):
        m = min(blk, corpus.n_local - a)
        raw, _ = index.read_raw(a, m)
        rows = ox.bf16_bits_to_f32(raw[:, :corpus.dim])
        s_all = rows @ qall.T                                   # fp32 shortlist scores [m, total queries]
        for pi, p in enumerate(probes):
            keep = None if p["mask"] is None else p["mask"][a:a + m]
            for j in range(qp[pi].shape[0]):
                s = s_all[:, spans[pi] + j]
                if keep is not None:
                    s = np.where(keep, s, -np.inf)
                w = min(m, 4 * p["k"])
                top = np.argpartition(-s, w - 1)[:w]
                top = top[np.isfinite(s[top])]
                exact = ox.scores_f64(rows[top], qp[pi][j], "cosine")
                cand[pi][j].append((top.astype(np.int64) + a + corpus.base, exact))
        del rows, raw, s_all
    for pi, p in enumerate(probes):
        out.append({"name": p["name"], "lists": [(np.concatenate([c[0] for c in cj]) if cj else np.empty(0, np.int64),
                                                  np.concatenate([c[1] for c in cj]) if cj else np.empty(0))
                                                 for cj in cand[pi]]})
    return out


def parity_report(probes: list[dict], gathered: list[list[dict]]) -> dict:
    """rank 0: merge every rank's shortlists by (score desc, id asc) and compare with the GPU's step-0 results."""
    from oracle import exact_search as ox

    rep = {"ok": True, "oracle": "oracle/exact_search.py (fp64 over the stored bf16 operands), sliced per rank", "workloads": {}}
    for pi, p in enumerate(probes):
        k, bad, worst, rec, n_checked = p["k"], 0, 0.0, [], 0
        for j in range(p["queries"].shape[0]):
            ids = np.concatenate([g[pi]["lists"][j][0] for g in gathered])
            sc = np.concatenate([g[pi]["lists"][j][1] for g in gathered])
            o = ox.order_desc_id_asc(sc, ids)[:k]
            want_ids, want_s = ids[o], sc[o]
            n = int(p["counts"][j])
            got_ids, got_s = np.asarray(p["ids"][j][:n]), np.asarray(p["scores"][j][:n], np.float64)
            n_checked += 1
            rec.append(len(set(got_ids.tolist()) & set(want_ids.tolist())) / max(1, len(want_ids)))
            if n != len(want_ids):
                bad += 1
                continue
            worst = max(worst, float(np.max(np.abs(got_s - want_s) / np.maximum(1e-3, np.abs(want_s)))) if n else 0.0)
            if not np.array_equal(got_ids, want_ids):
                # a different order / choice is accepted only between rows whose oracle scores tie within fp32 noise
                look = dict(zip(ids.tolist(), sc.tolist()))
                for a in np.flatnonzero(got_ids != want_ids):
                    if int(got_ids[a]) not in look or abs(look[int(got_ids[a])] - want_s[a]) > 2e-6:
                        bad += 1
                        break
        w = {"queries_checked": n_checked, "id_mismatches": bad, "max_score_rel_err": worst, "recall_at_k": min(rec) if rec else None,
             "k": k, "tolerance": "ids exact (ties within 2e-6 may swap); scores 1e-3 relative (bf16 storage)"}
        w["ok"] = bad == 0 and worst <= 1e-3
        rep["ok"] = rep["ok"] and w["ok"]
        rep["workloads"][p["name"]] = w
    return rep


# ----------------------------------------------------------------------------- in-process sharded store (rank 0, N > 1)
def sharded_store_leg(env: Env, rows: int, dim: int, k: int, steps: int, ref_first) -> dict:
    """The drop-in store's multi-GPU form: ONE process (this rank) drives all N GPUs through `yrb_sharded_search`
    (csrc/sharded.cu): host buffers in, merged result in pinned host memory out — what an agent process would call."""
    from youtu_rag_b200 import native

    torch = env.torch
    devs = list(range(env.world))
    sx = native.ShardedIndex(dim, "cosine", "bf16", devs, reserve_rows=rows, block_rows=16384)
    gen = torch.Generator(device=env.dev)
    for b in range((rows + BLOCK_ROWS - 1) // BLOCK_ROWS):
        n_b = min(BLOCK_ROWS, rows - b * BLOCK_ROWS)
        gen.manual_seed(b)
        blk = torch.randn(n_b, dim, device=env.dev, generator=gen)
        torch.cuda.synchronize(env.dev)
        sx.append_device(blk.data_ptr(), n_b, src_device=env.local)
        del blk
    hq = host_queries(dim, 1)
    first = sx.search(hq[0], k)
    for i in range(10):
        sx.search(hq[i % N_QUERY_SETS], k)
    lat = []
    l0 = sx.launches()
    t0 = time.perf_counter()
    for i in range(steps):
        a = time.perf_counter()
        sx.search(hq[i % N_QUERY_SETS], k)
        lat.append(time.perf_counter() - a)
    total = time.perf_counter() - t0
    launches = sx.launches() - l0
    lat.sort()
    same = bool(np.array_equal(first[0], ref_first[0]) and np.array_equal(first[1].view(np.uint32), np.asarray(ref_first[1]).view(np.uint32)))
    sx.close()
    return {"api": "yrb_sharded_search (one process, %d GPUs, host buffers)" % env.world, "value": steps / total, "unit": "queries/s",
            "ms_per_step": 1e3 * total / steps, "p50_ms": 1e3 * lat[len(lat) // 2], "p99_ms": 1e3 * lat[min(len(lat) - 1, int(len(lat) * 0.99))],
            "steps": steps, "gpu_launches": int(launches), "launches_per_search": launches / steps,
            "identical_to_torchrun_result": same, "h2d_bytes_per_step": dim * 4 * env.world, "d2h_bytes_per_step": k * 12 + 4}


def run_b200(args):
    from youtu_rag_b200.sharded import ShardedSearcher

    env = Env(args.gpus)
    main_name = args.workload or DEFAULT_WORKLOAD
    rows, dim, nq, k, sel = WORKLOADS[main_name]
    corpus = Corpus(env, rows, dim, args.path)
    searcher = ShardedSearcher(corpus.index, corpus.bounds, exchange=args.exchange)
    main, probe = measure(env, corpus, searcher, main_name, args.steps, args.warmup, args, with_cpu=True)
    probes, also = [probe], {}
    default_run = args.workload is None and not args.no_also
    if default_run:
        # the other BASELINE.json configurations on the same resident corpus (short runs)
        for name, st in (("c4", max(20, args.steps // 2)), ("t10mb", max(5, args.steps // 10)), ("t10mq", max(3, args.steps // 25))):
            try:
                r, p = measure(env, corpus, searcher, name, st, max(3, args.warmup // 4), args, with_cpu=False)
            except Exception as e:  # noqa: BLE001 - a side leg must not take the main line with it (one process: no peers to desynchronise)
                if env.world > 1:
                    raise
                also[name] = {"error": f"{type(e).__name__}: {e}"}
                continue
            also[name] = {x: r[x] for x in ALSO_KEYS}
            probes.append(p)
    rep = None
    if not args.no_parity:
        gathered = env.gather_objects(local_shortlists(corpus, probes))
        if env.rank == 0:
            rep = parity_report(probes, gathered)
    if default_run:
        if env.world == 1:
            # C3 is a one-GPU configuration over its own 1M-row corpus
            c3 = Corpus(env, *WORKLOADS["c3"][:2], args.path)
            s3 = ShardedSearcher(c3.index, c3.bounds, exchange=args.exchange)
            r, p = measure(env, c3, s3, "c3", max(20, args.steps // 2), max(3, args.warmup // 4), args, with_cpu=False)
            also["c3"] = {x: r[x] for x in ALSO_KEYS}
            if rep is not None:
                r3 = parity_report([p], [local_shortlists(c3, [p])])
                rep["workloads"].update(r3["workloads"])
                rep["ok"] = rep["ok"] and r3["ok"]
            c3.index.close()
        else:
            env.host_wait()
            if env.rank == 0:
                try:
                    also["sharded_store"] = sharded_store_leg(env, rows, dim, k, min(args.steps, 300), (probe["ids"], probe["scores"]))
                except Exception as e:  # noqa: BLE001 - report, do not lose the main line
                    also["sharded_store"] = {"error": f"{type(e).__name__}: {e}"}
            env.host_wait()
    if env.rank == 0:
        line = dict(main)
        if also:
            line["also"] = also
        line["parity"] = rep
        if rep is not None:
            line["recall_at_k"] = rep["workloads"][main_name]["recall_at_k"]
        print(json.dumps(line, default=float))
    if env.world > 1:
        env.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS),
                    help=f"run ONE named workload (default: {DEFAULT_WORKLOAD} + the `also` runs)")
    ap.add_argument("--path", type=int, default=0, help="force kernel family: 1 K1, 2 K2, 3 K6")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-also", action="store_true", help="skip the `also` runs")
    ap.add_argument("--no-parity", action="store_true", help="skip the in-bench parity check")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="multi-GPU top-k exchange: p2p = one kernel over NVLink peer memory (K7), nccl = all-gather + merge")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
