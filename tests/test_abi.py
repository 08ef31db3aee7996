"""CPU-only: the C-ABI library builds for sm_100a, loads, and exports every symbol include/p2b.h
declares; the product package does not route through the oracle."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import city_rollup_b200 as m

    m.build()
    return ctypes.CDLL(m.SO_PATH)


def header_symbols():
    src = open(os.path.join(ROOT, "include", "p2b.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(p2b_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(lib):
    syms = header_symbols()
    assert len(syms) >= 40
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/p2b.h but not exported by libp2b.so"


def test_binding_table_matches_header():
    from city_rollup_b200 import _lib

    assert sorted(_lib.SIGNATURES) == header_symbols()


def test_version_and_no_cpu_fallback(lib):
    lib.p2b_version.restype = ctypes.c_int
    assert lib.p2b_version() >= 100
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present: the failure path is exercised on CPU-only hosts")
    import city_rollup_b200 as m

    with pytest.raises(m.P2BError) as e:
        m.Context()
    assert "no CPU fallback" in str(e.value)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "city_rollup_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", "Makefile")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "p2oracle" not in txt and "oracle/" not in txt.replace("oracle/ is test", ""), (dirpath, f)
    for f in ("include/p2b.h",):
        assert "p2oracle" not in open(os.path.join(ROOT, f)).read()


def test_cpp_host_mirror_compiles(lib, tmp_path):
    """the header-only C++ mirror (city_rollup_b200/cpp/plonky2_b200.hpp) and the native job loop built on it
    (tools/prove_bench.cpp, tools/qbench_replay.cpp) compile and link against libp2b.so"""
    import shutil
    import subprocess

    import city_rollup_b200 as m

    if not shutil.which("g++"):
        pytest.skip("no g++")
    for src in ("prove_bench.cpp", "qbench_replay.cpp"):
        out = tmp_path / src.replace(".cpp", "")
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-Wall", "-I", ROOT, os.path.join(ROOT, "tools", src),
                               "-L", os.path.dirname(m.SO_PATH), "-lp2b", "-lpthread", "-o", str(out)])
        assert out.exists()
