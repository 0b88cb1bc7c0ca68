#!/usr/bin/env python
"""bench.py — BASELINE.json's metric (queries/sec + p50 latency of exact top-k) on synthetic data.

    python bench.py --gpus N --steps K --warmup W [--workload c2|t10m|c3|c4|c5] [--impl reference]

A "step" is one search (one pass of the hot path over the resident corpus for one batch of
queries).  Default workload = BASELINE.json configs[1] (C2): 1M x 1024 bf16 corpus, single query,
top-10, cosine.  With N>1 (torchrun, one rank per GPU) the SAME corpus is row-sharded across the
ranks ("strong" scaling): local scan+top-k, one NCCL all-gather of the nq*k keys, K3 merge.

`value`   whole-job queries/s with queries already resident in HBM (device-timed, max over ranks)
`e2e`     same metric through the C-ABI call with HOST buffers (H2D query + D2H result inside)
`roofline`  algorithmic bytes (or flops) of the dominant kernel / its CUDA-event duration
`cpu_baseline`  the oracle port of the reference's exact CPU path (FAISSVectorStore semantics,
            numpy fp32 BLAS) timed on this box's host cores, reported beside — not a target.
`--impl reference` prints that CPU arm as the main line (the reference's engines, chromadb /
faiss-cpu, are not installable offline).
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
    "t10m": (10_000_000, 1024, 1, 10, None),
    "c3": (1_000_000, 1024, 256, 100, None),
    "c4": (10_000_000, 1024, 1, 10, 0.10),
    "c4b": (10_000_000, 1024, 256, 10, 0.10),
    "c5": (100_000_000, 768, 1024, 10, None),
    "tiny": (250_000, 1024, 1, 10, None),
    "t10mb": (10_000_000, 1024, 256, 100, None),   # 256-query batches over the 10M corpus (near-linear scaling case)
    "c5s": (12_500_000, 768, 1024, 10, None),        # one C5 shard (12.5M x 768 per GPU of the 8-GPU config)
    "c3s": (125_000, 1024, 256, 100, None),        # one C3 shard of an 8-GPU run, for the fixed-cost breakdown
}
N_QUERY_SETS = 64


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
    """Samples SM clock + throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, device: int):
        super().__init__(daemon=True)
        self.device, self.samples, self.reasons, self.max_mhz = device, [], set(), None
        self._halt = threading.Event()

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
            while not self._halt.is_set():
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = get_reasons(h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
                time.sleep(0.02)
        except Exception as e:  # noqa: BLE001
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def finish(self) -> dict:
        self._halt.set()
        self.join(timeout=2)
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ----------------------------------------------------------------------------- CPU arm
def cpu_corpus(rows: int, dim: int) -> np.ndarray:
    from oracle import exact_search as ox

    out = np.empty((rows, dim), np.float32)
    for b in range((rows + BLOCK_ROWS - 1) // BLOCK_ROWS):
        a, e = b * BLOCK_ROWS, min(rows, (b + 1) * BLOCK_ROWS)
        out[a:e] = ox.l2_normalize(np.random.default_rng(b).standard_normal((e - a, dim), dtype=np.float32))
    return out


def cpu_arm(corpus: np.ndarray, queries: np.ndarray, k: int, steps: int, warmup: int, budget_s: float = 25.0):
    """FAISSVectorStore.search semantics (faiss_store.py:143-199) as restated in oracle/exact_search.py:
    fp32 BLAS scan + partition + ordered top-k, looped over the step's queries like
    base_retriever.py:95-99.  Returns (qps, per-step seconds list)."""
    from oracle import exact_search as ox

    def step(i):
        for q in queries[i % queries.shape[0]]:
            ox.faiss_flat_search(corpus, q, k, "cosine")

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
    return nq * len(times) / sum(times), times


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
    rows, dim, nq, k, sel = WORKLOADS[args.workload]
    # bounded sample of the workload: at most 1M rows of the corpus and at most 4 queries per step
    s_rows, s_nq = min(rows, 1_000_000), min(nq, 4)
    corpus = cpu_corpus(s_rows, dim)
    q = host_queries(dim, nq)[:, :s_nq]
    qps, times = cpu_arm(corpus, q, k, args.steps, args.warmup, budget_s=120.0)
    # scale to the full workload: a step scans `rows` rows for `nq` queries; the scan is linear in both
    scale = (s_rows / rows)
    value = qps * scale
    sample = f"{s_rows}x{dim} fp32 rows, {s_nq} of {nq} queries per step, {len(times)} steps"
    line = {
        "impl": "reference", "metric": "queries_per_sec", "value": value, "unit": "queries/s", "n_gpus": args.gpus,
        "steps": len(times), "warmup": args.warmup, "ms_per_step": 1e3 * nq / value, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.workload, args.gpus),
        "cpu_baseline": {"value": value, "unit": "queries/s", "cores": blas_threads(), "kind": "port", "sample": sample,
                         "p50_ms": 1e3 * statistics.median(times) / s_nq / scale},
        "e2e": {"value": value, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference engines (chromadb 1.3.4 HNSW / faiss-cpu 1.12.0) are not installable offline; this is the "
                "oracle port of FAISSVectorStore.search (exact, numpy fp32 BLAS) on the host cores",
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
def run_b200(args):
    import torch
    import torch.distributed as dist

    from youtu_rag_b200 import native
    from youtu_rag_b200.sharded import ShardedSearcher

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    rows, dim, nq, k, sel = WORKLOADS[args.workload]
    n_blocks = (rows + BLOCK_ROWS - 1) // BLOCK_ROWS
    if n_blocks % world:
        raise SystemExit(f"{n_blocks} corpus blocks do not divide over {world} ranks")
    bounds = [min(rows, b * (n_blocks // world) * BLOCK_ROWS) for b in range(world + 1)]
    bounds[-1] = rows
    my_blocks = range(rank * (n_blocks // world), (rank + 1) * (n_blocks // world))
    n_local = bounds[rank + 1] - bounds[rank]

    index = native.Index(dim, "cosine", "bf16", local, n_local)
    gen = torch.Generator(device=dev)
    mask_bits = []
    for b in my_blocks:
        n_b = min(BLOCK_ROWS, rows - b * BLOCK_ROWS)
        gen.manual_seed(b)
        blk = torch.randn(n_b, dim, device=dev, generator=gen)
        torch.cuda.synchronize(dev)
        index.append_device(blk.data_ptr(), n_b)
        if sel:
            gen.manual_seed(2_000_000 + b)
            mask_bits.append(torch.rand(n_b, device=dev, generator=gen) < sel)
        del blk
    assert index.rows == n_local
    dev_mask = None
    if sel:
        bits = torch.cat(mask_bits)
        pad = (-bits.numel()) % 64
        bits = torch.cat([bits, torch.zeros(pad, dtype=torch.bool, device=dev)]).view(-1, 32).to(torch.int64)
        words = (bits << torch.arange(32, device=dev, dtype=torch.int64)).sum(1)
        dev_mask = words.to(torch.int32)  # low 32 bits (two's complement wrap is the bit pattern we want)
        n_pass = int(torch.cat(mask_bits).sum().item())
    if args.path:
        index.set_path(args.path)

    hq = host_queries(dim, nq)
    dq = torch.from_numpy(hq).to(dev)
    searcher = ShardedSearcher(index, bounds, exchange=args.exchange)
    st = searcher.stream  # every kernel of a step is launched on this stream

    def step_device(i):
        return searcher.search_device(dq[i % N_QUERY_SETS], k, dev_mask)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---------------- value: device-resident inputs, K steps timed with CUDA events
    for i in range(args.warmup):
        step_device(i)
    barrier()
    index.profile(True)
    index.profile_read()
    l0 = index.launches()
    sampler = ClockSampler(local)
    sampler.start()
    # steps are issued back to back: scan on `st`, exchange + merge on the searcher's comm stream, so the
    # exchange of step i overlaps the scan of step i+1 (independent queries).  The timed region starts on the
    # scan stream and ends when the LAST step's merged result is complete on the comm stream.
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for i in range(args.steps):
        step_device(args.warmup + i)
    e1.record(searcher.comm_stream)
    barrier()
    clocks = sampler.finish()
    ms_total = e0.elapsed_time(e1)
    launches = index.launches() - l0 + (args.steps if world > 1 else 0)  # + the K3 global merge per step
    kern_ms, kern_n = index.profile_read()
    index.profile(False)
    t = torch.tensor([ms_total], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = nq * 1e3 / ms_step

    # ---------------- latency: per-step synchronised (p50 / p99)
    lat = []
    for i in range(min(args.steps, 500)):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(dev)
        a.record(st)
        step_device(i)
        b.record(searcher.comm_stream)
        b.synchronize()
        lat.append(a.elapsed_time(b))
    lat.sort()

    # ---------------- e2e: host buffers through the C ABI (H2D query, D2H ids+scores inside)
    barrier()
    e2e_steps = min(args.steps, 500)

    def step_host(i):
        if world == 1 and dev_mask is None:
            return index.search(hq[i % N_QUERY_SETS], k)
        return searcher.search(hq[i % N_QUERY_SETS], k, dev_mask)

    for i in range(max(3, args.warmup // 4)):
        step_host(i)
    barrier()
    host_lat = []
    t0 = time.perf_counter()
    t_prev = t0
    for i in range(e2e_steps):
        step_host(i)
        t_now = time.perf_counter()
        host_lat.append(t_now - t_prev)  # host call -> ids/scores on the host (SURVEY.md §8d's latency)
        t_prev = t_now
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    host_lat.sort()
    t = torch.tensor([e2e_s], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    e2e_value = nq * e2e_steps / e2e_s

    # ---------------- roofline of the dominant kernel
    hbm, tf_burst, tf_sust, peak_src = peaks()
    frac_rows = (n_pass / n_local) if sel else 1.0
    traffic = None
    tp = ROOT / "profiles" / "traffic.json"
    if tp.exists():
        traffic = json.loads(tp.read_text()).get(args.workload)
    kern_avg_ms = kern_ms / max(1, kern_n)
    launches_per_step = kern_n / args.steps
    if nq < 8 or args.path == 1:
        # K1: one launch scans the shard once for one query
        algo_bytes = frac_rows * n_local * dim * 2 + (n_local / 8 if sel else 0) + dim * 4 + (148 if world == 1 else 146) * k * 8
        roof = {"bound": "hbm", "kernel": "k1_scan_topk", "achieved": algo_bytes / (kern_avg_ms * 1e-3) / 1e9,
                "peak": hbm, "unit": "GB/s", "peak_source": peak_src, "bytes_per_launch": algo_bytes,
                "avg_launch_ms": kern_avg_ms, "launches_per_step": launches_per_step, "traffic": traffic}
    else:
        # the timed launch is K2's main GEMM (every tile of the shard, first 256-query chunk)
        # (a shared mask passing <= 25 % of the rows is compacted first by K8: the GEMM then covers the passing rows)
        rows_b = n_pass if (sel and k <= n_pass <= n_local // 4) else n_local
        flops = 2.0 * min(nq, 256) * rows_b * dim
        roof = {"bound": "tensor", "kernel": "k2_gemm_topk (phase B launch)", "achieved": flops / (kern_avg_ms * 1e-3) / 1e12,
                "peak": tf_burst, "unit": "TFLOP/s", "peak_source": peak_src, "flops_per_launch": flops,
                "avg_launch_ms": kern_avg_ms, "launches_per_step": launches_per_step, "traffic": traffic,
                "hbm_gbs_same_kernel": (rows_b * dim * 2) / (kern_avg_ms * 1e-3) / 1e9,
                "rows_in_gemm": rows_b,
                "note": "dense GEMM over rows_in_gemm rows (all rows, or the K8-compacted passing rows of a selective shared mask)"}
    roof["frac"] = roof["achieved"] / roof["peak"]

    # ---------------- CPU baseline beside it (rank 0, N=1 only; bounded sample)
    cpu = None
    if world == 1 and rank == 0 and not args.no_cpu:
        s_rows = min(n_local, 1_000_000)
        corpus = index.read_rows(np.arange(s_rows))          # the same stored rows, decoded to fp32
        s_nq = min(nq, 4)
        qps, times = cpu_arm(corpus, hq[:, :s_nq], k, 200, 10, budget_s=20.0)
        scale = s_rows / n_local
        cpu = {"value": qps * scale, "unit": "queries/s", "cores": blas_threads(), "kind": "port",
               "sample": f"{s_rows}x{dim} fp32 rows (read back from the index), {s_nq} of {nq} queries per step, "
                         f"{len(times)} steps after 10 warm-ups",
               "p50_ms": 1e3 * statistics.median(times) / s_nq / scale, "host_cpus": os.cpu_count()}
        del corpus

    if rank == 0:
        line = {
            "metric": "queries_per_sec", "value": value, "unit": "queries/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": workload_config(args.workload, world),
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
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--path", type=int, default=0, help="force kernel family: 1 K1, 2 K2, 3 K6")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
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
