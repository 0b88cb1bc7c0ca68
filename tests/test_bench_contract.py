"""bench.py's reference arm runs without a GPU: check the JSON line it prints against the bench contract."""

import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_reference_arm_prints_the_contract_line():
    p = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--workload", "tiny", "--steps", "3", "--warmup", "3"],
                       capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, p.stdout
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"] == "queries_per_sec" and j["unit"] == "queries/s"
    assert j["higher_is_better"] is True and j["n_gpus"] == 1 and j["steps"] == 3 and j["warmup"] >= 3
    assert j["value"] > 0 and abs(j["value"] - 1e3 * j["config"]["queries_per_step"] / j["ms_per_step"]) < 1e-6 * j["value"]
    assert j["vs_baseline"] is None and j["data"] == "synthetic" and "workload" in j["config"]
    cb = j["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == j["value"] and cb["sample"]
    assert j["e2e"] == {"value": j["value"], "unit": j["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_traffic_ratios_cover_every_kernel_variant_the_bench_reports():
    """bench.py's `roofline.traffic` = ratio x this run's algorithmic bytes; every ratio comes from a committed
    `ncu --set full` summary and must name it (profiles/traffic.json)."""
    root = Path(__file__).resolve().parent.parent
    t = json.loads((root / "profiles" / "traffic.json").read_text())
    for key in ("k1", "k2", "k2mc"):
        assert 1.0 <= t[key]["ratio"] < 1.1, key            # no re-reads: dram bytes within 10 % of the algorithmic bytes
        src = t[key]["source"].split(":")[0]
        assert (root / src).exists(), src
    sys.path.insert(0, str(root))
    try:
        import bench
    finally:
        sys.path.pop(0)
    assert bench.traffic_of("k2mc", 1e9, None)["traffic"] == t["k2mc"]["ratio"] * 1e9
    assert bench.traffic_of("k2mc", 1e9, True) == {"traffic": None}   # no capture of a masked multi-chunk launch
