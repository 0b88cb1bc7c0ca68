"""Turn gpurun_out/*.ncu-rep and launch-list CSVs into the small text summaries kept under profiles/.

    python profiles/summarize.py launches gpurun_out/launches_c2.csv > profiles/r01_c2_launches.txt
    python profiles/summarize.py full gpurun_out/prof_k1.ncu-rep   > profiles/r01_k1_full.txt
"""
import collections
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "smsp__inst_executed.sum", "l1tex__t_bytes.sum", "sm__cycles_active.avg",
    "smsp__cycles_active.avg", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__shared_mem_per_block_dynamic",
]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    h = rows[0]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        v = v / 1e3 if r[ui] in ("ns", "nsecond") else v
        agg.setdefault(r[ki], []).append(v)
    total = sum(sum(v) for v in agg.values())
    print(f"# {path}: per-kernel device time (us), cold-cache + serialised under ncu: compare SHARES")
    print(f"{'kernel':90s} {'n':>4s} {'avg_us':>10s} {'total_us':>10s} {'share':>7s}")
    for k, v in agg.items():
        print(f"{k[:90]:90s} {len(v):4d} {sum(v) / len(v):10.1f} {sum(v):10.1f} {100 * sum(v) / total:6.1f}%")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h = rows[0]
    ni = h.index("Kernel Name")
    print(f"# {path}: ncu --set full, per captured launch (row 1 = units)")
    for r in rows[2:]:
        print("kernel:", r[ni][:120])
    for k in KEYS:
        if k in h:
            i = h.index(k)
            print(f"{k:75s} {rows[1][i]:>16s} " + " ".join(f"{r[i]:>14s}" for r in rows[2:]))


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
