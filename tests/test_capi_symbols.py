"""CPU: librg_b200.so loads and exports every symbol include/rg_b200.h declares; no compute without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, has_cuda

import roargraph_b200 as rg


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "rg_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rg_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    if not os.path.exists(rg.LIB_PATH):
        rg.build()
    L = ctypes.CDLL(rg.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/rg_b200.h but not exported"
    assert L.rg_abi_version() == 1


@pytest.mark.skipif(has_cuda(), reason="only meaningful on a machine without a GPU")
def test_no_cpu_fallback_without_gpu():
    base = np.zeros((16, 8), dtype=np.float32)
    with pytest.raises(rg.RgError) as e:
        rg.RoarIndex(base, "ip")
    assert e.value.code == rg.RG_ECUDA
    with pytest.raises(rg.RgError) as e:
        rg.knn(base, base, 1, "ip")
    assert e.value.code == rg.RG_ECUDA


def test_product_never_touches_the_oracle():
    """The product path must not import, link or execute anything under oracle/."""
    bad = []
    for d in ("roargraph_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, d)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")) or f == "Makefile":
                    txt = open(os.path.join(dirpath, f), errors="ignore").read()
                    if re.search(r"oracle|liboracle|libref_roargraph|_ref/", txt):
                        bad.append(os.path.join(dirpath, f))
    assert not bad, bad


def test_integration_snippets_compile_and_link(tmp_path):
    """tests/cpp/integration_calls.cpp repeats the C-ABI calls INTEGRATION.md shows a reference maintainer (same argument
    order and types): it must compile as C++ against include/rg_b200.h and link against the shipped library."""
    import subprocess
    src = os.path.join(ROOT, "tests", "cpp", "integration_calls.cpp")
    out = str(tmp_path / "libintegration_calls.so")
    libdir = os.path.join(ROOT, "roargraph_b200")
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-Werror", "-shared", "-fPIC", "-I", os.path.join(ROOT, "include"), src,
                           "-o", out, "-L", libdir, "-l:librg_b200.so", "-Wl,--no-undefined", "-Wl,-rpath," + libdir])
    assert os.path.getsize(out) > 0
