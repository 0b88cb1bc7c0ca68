"""The C-ABI library loads and exports every symbol include/yrb200.h declares (no GPU needed)."""

import ctypes
import re
from pathlib import Path

import pytest

from youtu_rag_b200 import native

HEADER = (Path(__file__).resolve().parent.parent / "include" / "yrb200.h").read_text()


def test_header_and_binding_list_the_same_symbols():
    declared = set(re.findall(r"YRB_API\s+[\w\s\*]+?\b(yrb_\w+)\s*\(", HEADER))
    assert declared == set(native.EXPORTS)


def test_library_exports_every_declared_symbol():
    L = ctypes.CDLL(str(native.LIB_PATH))
    for name in native.EXPORTS:
        assert hasattr(L, name), name
    assert native.lib().yrb_abi_version() == 1


def test_header_constants_match_binding():
    def const(name):
        return int(re.search(rf"#define {name} \(?(-?\d+)\)?", HEADER).group(1))
    assert const("YRB_FUSED_K_MAX") == native.FUSED_K_MAX
    assert const("YRB_WHERE_MAX_LEAVES") == native.WHERE_MAX_LEAVES
    assert const("YRB_WHERE_MAX_OPERANDS") == native.WHERE_MAX_OPERANDS
    assert const("YRB_WHERE_MAX_TOKENS") == native.WHERE_MAX_TOKENS
    for k, v in native.OPS.items():
        assert const("YRB_OP_" + k[1:].upper()) == v
    assert (const("YRB_TOK_AND"), const("YRB_TOK_OR"), const("YRB_TOK_NOT")) == (-1, -2, -3)
    assert [const(f"YRB_COL_{n}") for n in ("I64", "F64", "CODE", "BOOL")] == [0, 1, 2, 3]
    assert [const(f"YRB_METRIC_{n}") for n in ("COSINE", "DOT", "L2")] == [0, 1, 2]


def test_no_cpu_fallback_without_a_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(native.NativeError) as e:
        native.Index(16)
    assert e.value.code == -5 and "no CPU fallback" in str(e.value)


def test_decode_keys_inverse_of_oracle_key():
    import numpy as np

    from oracle import exact_search as ox
    s = np.random.default_rng(0).standard_normal(1000).astype(np.float32)
    rows = np.arange(1000, dtype=np.uint32)
    keys = (ox.score_key_u32(s).astype(np.uint64) << np.uint64(32)) | (~rows).astype(np.uint64)
    r, sc = native.decode_keys(np.concatenate([keys, np.zeros(1, np.uint64)]))
    assert np.array_equal(r[:-1], rows) and np.array_equal(sc[:-1], s) and r[-1] == -1 and sc[-1] == -np.inf


def _run_c_consumer(tmp_path):
    import subprocess

    root = Path(__file__).resolve().parent.parent
    exe = tmp_path / "abi_smoke"
    cmd = ["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", str(root / "include"),
           str(root / "tests" / "abi_c" / "abi_smoke.c"), "-o", str(exe), str(native.LIB_PATH), f"-Wl,-rpath,{native.LIB_PATH.parent}"]
    b = subprocess.run(cmd, capture_output=True, text=True)
    assert b.returncode == 0, b.stdout + b.stderr
    return subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)


def test_header_is_plain_c_and_a_c_program_links_every_entry(tmp_path):
    """include/yrb200.h through a C11 compiler with -pedantic -Werror, linked against libyrb200.so from C.  Without a
    B200 the program checks that constructors refuse to work (no CPU fallback) and that the stateless row map answers."""
    r = _run_c_consumer(tmp_path)
    assert r.returncode == 0 and r.stdout.startswith("OK"), r.stdout + r.stderr


@pytest.mark.gpu
def test_c_consumer_searches_on_the_device(tmp_path):
    r = _run_c_consumer(tmp_path)
    assert r.returncode == 0 and r.stdout.startswith("OK device"), r.stdout + r.stderr
