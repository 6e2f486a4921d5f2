"""CPU-side checks of the boundary: the shared library loads, exports every symbol that
include/metro_b200.h declares, and refuses to run without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest

import metrocpp_b200 as mb
from metrocpp_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "metro_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(metro_[a-z0-9_]+)\s*\(", text)))


def test_header_and_loader_agree():
    assert declared_symbols() == sorted(_lib.SYMBOLS)


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(lib, name), name
    assert b"sm_100a" in _lib.load().metro_version()


def test_library_exports_nothing_undeclared():
    """The product ABI is exactly include/metro_b200.h: no extra extern "C" metro_* entry point (benchmarks live in
    libmetro_b200_bench.so)."""
    import subprocess
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = sorted({ln.split()[-1] for ln in out.splitlines() if re.match(r".* T metro_[a-z0-9_]+$", ln)})
    assert exported == declared_symbols()


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(mb.MetroError) as e:
        mb.MetroEngine(0)
    assert e.value.code == _lib.ERR_CUDA


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "metrocpp_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")) or f == "Makefile":
                assert "oracle" not in open(os.path.join(dirpath, f)).read(), os.path.join(dirpath, f)
