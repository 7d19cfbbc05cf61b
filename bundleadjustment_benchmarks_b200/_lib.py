"""ctypes loader of the in-tree CUDA library ``libba_b200.so`` (C ABI: include/ba_gpu.h).

There is no CPU fallback: if the library is missing it is built with nvcc (cross-compiles without a
GPU); if it cannot be loaded, importing a solver raises. Creating a solver without a CUDA device
fails inside ``ba_create``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libba_b200.so")
_SRC = [os.path.join(_HERE, "csrc", n) for n in
        ("ba_gpu.cu", "ba_tile.cuh", "ba_model.cuh", "ba_dense.cuh", "ba_ldlt2.cuh", "ba_split.cuh", "ba_qr.cuh")] + \
       [os.path.join(os.path.dirname(_HERE), "include", "ba_gpu.h")]

# every symbol include/ba_gpu.h declares
SYMBOLS = [
    "ba_last_error", "ba_version", "ba_create", "ba_destroy", "ba_comm_unique_id", "ba_comm_init",
    "ba_bandwidth", "ba_set_bandwidth", "ba_set_state", "ba_get_state", "ba_eval", "ba_linearize",
    "ba_compute", "ba_solve_try", "ba_accept", "ba_reject", "ba_get_dx", "ba_step_streamed", "ba_error_statistics", "ba_split_plan", "ba_get_residuals",
    "ba_get_reduced_system", "ba_keep_reduced_system", "ba_get_jacobian", "ba_launch_count",
    "ba_stage_ms", "ba_set_profiling", "ba_timer_start", "ba_timer_stop", "ba_debug_counters",
    "ba_debug_band_solve", "ba_numeric_status", "ba_set_strict_numeric", "ba_debug_counters_n",
]

_LIB = None


def needs_build() -> bool:
    if not os.path.exists(SO_PATH):
        return True
    t = os.path.getmtime(SO_PATH)
    return any(os.path.exists(s) and os.path.getmtime(s) > t for s in _SRC)


def build(force: bool = False, verbose: bool = False) -> str:
    """nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo ... (see csrc/Makefile)."""
    if force or needs_build():
        cmd = ["make", "-C", os.path.join(_HERE, "csrc")] + (["-B"] if force else [])
        out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if verbose or out.returncode != 0:
            print(out.stdout)
        if out.returncode != 0:
            raise RuntimeError("building libba_b200.so failed")
    return SO_PATH


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    if needs_build():  # missing, or older than any source: never measure a stale binary
        build()
    L = C.CDLL(SO_PATH)  # raises OSError loudly if the CUDA extension is missing/unloadable
    dp, ip, vp = C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_void_p
    L.ba_last_error.restype = C.c_char_p
    L.ba_version.restype = C.c_char_p
    L.ba_create.argtypes = [C.POINTER(vp), C.c_int, C.c_int, C.c_int, ip, ip, dp, C.c_double, C.c_int, C.c_int, C.c_int]
    L.ba_destroy.argtypes = [vp]
    L.ba_comm_unique_id.argtypes = [vp]
    L.ba_comm_init.argtypes = [vp, C.c_int, C.c_int, vp]
    L.ba_bandwidth.argtypes = [vp, ip]
    L.ba_set_bandwidth.argtypes = [vp, C.c_int]
    L.ba_set_state.argtypes = [vp] + [dp] * 6
    L.ba_get_state.argtypes = [vp] + [dp] * 6
    L.ba_eval.argtypes = [vp, dp]
    L.ba_linearize.argtypes = [vp, dp, dp, dp]
    L.ba_compute.argtypes = [vp, C.c_double]
    L.ba_solve_try.argtypes = [vp, dp, dp, dp]
    L.ba_accept.argtypes = [vp]
    L.ba_reject.argtypes = [vp]
    L.ba_get_dx.argtypes = [vp, dp]
    L.ba_error_statistics.argtypes = [vp, C.c_double, C.c_double, dp]
    L.ba_split_plan.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, ip, C.c_int]
    L.ba_step_streamed.argtypes = [vp, dp, dp, dp, dp, dp, dp, C.c_double, dp, dp, dp, dp, dp]
    L.ba_get_residuals.argtypes = [vp, dp]
    L.ba_get_reduced_system.argtypes = [vp, dp, dp]
    L.ba_keep_reduced_system.argtypes = [vp, C.c_int]
    L.ba_get_jacobian.argtypes = [vp, dp, dp]
    L.ba_launch_count.argtypes = [vp, C.POINTER(C.c_longlong)]
    L.ba_stage_ms.argtypes = [vp, dp]
    L.ba_set_profiling.argtypes = [vp, C.c_int]
    L.ba_debug_counters.argtypes = [vp, C.POINTER(C.c_longlong)]
    L.ba_debug_counters_n.argtypes = [vp, C.POINTER(C.c_longlong), C.c_int]
    L.ba_debug_band_solve.argtypes = [vp, C.c_int, C.c_int, dp, dp, dp]
    L.ba_timer_start.argtypes = [vp]
    L.ba_numeric_status.argtypes = [vp, ip]
    L.ba_set_strict_numeric.argtypes = [vp, C.c_int]
    L.ba_timer_stop.argtypes = [vp, dp]
    for s in SYMBOLS:
        if s not in ("ba_last_error", "ba_version"):
            getattr(L, s).restype = C.c_int
    _LIB = L
    return L
