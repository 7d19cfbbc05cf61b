"""CPU: the C-ABI shared library loads and exports every symbol include/ba_gpu.h declares; argument
errors are reported without a GPU; there is no CPU fallback."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from bundleadjustment_benchmarks_b200 import _lib, bal, solver

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "ba_gpu.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(ba_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported():
    L = _lib.lib()
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(L, s), s
    assert sorted(_lib.SYMBOLS) == syms


def test_version_and_error_strings():
    L = _lib.lib()
    assert b"sm_100a" in L.ba_version()
    assert isinstance(L.ba_last_error(), bytes)


def test_no_cpu_fallback_or_argument_errors(tiny):
    import torch
    L = _lib.lib()
    h = C.c_void_p()
    view = np.ascontiguousarray(tiny.view); point = np.ascontiguousarray(tiny.point)
    meas = np.ascontiguousarray(tiny.meas)
    ip, dp = C.POINTER(C.c_int), C.POINTER(C.c_double)
    rc = L.ba_create(C.byref(h), tiny.N, tiny.M, tiny.K, view.ctypes.data_as(ip), point.ctypes.data_as(ip),
                     meas.ctypes.data_as(dp), 0.5, 1, 7, 0)
    assert rc == -1 and b"variant" in L.ba_last_error()
    bad = point.copy(); bad[[0, -1]] = bad[[-1, 0]]  # unsorted
    rc = L.ba_create(C.byref(h), tiny.N, tiny.M, tiny.K, view.ctypes.data_as(ip), bad.ctypes.data_as(ip),
                     meas.ctypes.data_as(dp), 0.5, 1, 1, 0)
    assert rc == -1 and b"sorted" in L.ba_last_error()
    if not torch.cuda.is_available():
        with pytest.raises(solver.BAError, match="no CPU fallback|CUDA"):
            solver.GpuSolver(tiny)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "bundleadjustment_benchmarks_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("oracle/", "").lower() or f == "README.md" or "ba_oracle" not in txt, f
                assert "ba_oracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, f
