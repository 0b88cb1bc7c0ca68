"""Build libyrb200.so in-tree with nvcc for sm_100a (no JIT cache, the .so travels with the repo)."""

from __future__ import annotations

import concurrent.futures as cf
import os
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
OUT = HERE / "libyrb200.so"
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-Wall", "--expt-relaxed-constexpr",
    "-I", str(HERE.parent / "include"),
]


def _stale(target: Path, deps: list[Path]) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    srcs = sorted(CSRC.glob("*.cu"))
    hdrs = sorted(CSRC.glob("*.h")) + sorted(CSRC.glob("*.cuh")) + sorted((HERE.parent / "include").glob("*.h"))
    objdir = HERE / "build"
    objdir.mkdir(exist_ok=True)
    jobs = []
    for s in srcs:
        o = objdir / (s.stem + ".o")
        if force or _stale(o, [s] + hdrs):
            cmd = [NVCC, *FLAGS, "-c", str(s), "-o", str(o)]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            jobs.append(cmd)
    if jobs:
        with cf.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for cmd, res in zip(jobs, ex.map(lambda c: subprocess.run(c, capture_output=True, text=True), jobs)):
                if verbose or res.returncode:
                    sys.stderr.write(" ".join(cmd) + "\n" + res.stdout + res.stderr)
                if res.returncode:
                    raise RuntimeError(f"nvcc failed: {' '.join(cmd)}")
    objs = [objdir / (s.stem + ".o") for s in srcs]
    if force or jobs or _stale(OUT, objs):
        cmd = [NVCC, "-shared", "-o", str(OUT), *map(str, objs), "-gencode", "arch=compute_100a,code=sm_100a",
               "-Xcompiler", "-fPIC"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError("link failed")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
