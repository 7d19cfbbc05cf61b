"""Host-side mirror of the reference's functor/solver interface for the LM hot path, over the C ABI.

``GpuSolver`` plays the role of BAFunctor + its QRSolver typedef (reference:
src/Optimization/BAFunctor.h:34-123): ``__call__``-like ``eval`` (operator()), ``linearize`` (df +
JtRes + column norms), ``compute(lambda)`` / ``solve_try()`` (m_solver.compute, matrixQ^T b, right
solver, back-substitution, increment_in_place, functor(xTest)), ``accept`` / ``reject``.
``minimize`` is the LM control flow of src/Eigen_ext/BacktrackLevMarq{QRChol,More,Cholesky}.h with the
same constants; all arithmetic of a trial runs on the GPU.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
import math
from typing import List, Optional

import numpy as np

from . import _lib

QRKIT, QRCHOL, MOREQR, CHOLESKY = 0, 1, 2, 3
VARIANTS = {"QRKIT": QRKIT, "QRCHOL": QRCHOL, "MOREQR": MOREQR, "CHOLESKY": CHOLESKY}
F32, F64 = 0, 1

# Status values of BacktrackLevMarqQRCHolInfo::Status (QRChol.h:39-46)
NOT_STARTED, RUNNING, SUCCESS, EXCEEDED_LAMBDA_MAX, TOO_MANY_FEVALS, MAX_ITERS = -2, -1, 0, 1, 2, 3
STATUS_STR = {SUCCESS: "Success (Energy Flatlined)", EXCEEDED_LAMBDA_MAX: "Success (Exceeded Maximum Lambda)",
              TOO_MANY_FEVALS: "Too Many Function Evaluations", MAX_ITERS: "Maximum Iterations Reached",
              NOT_STARTED: "Not Started", RUNNING: "Running"}


class BAError(RuntimeError):
    pass


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


@dataclasses.dataclass
class Trial:
    iter: int
    accepted: bool
    energy: float
    energy_test: float
    rho: float
    lambda_used: float
    lambda_next: float
    dx_norm: float


_PLAN_LIB = None


def split_plan(n: int, kd: int, mode: int = 1, segments: int = 3) -> dict:
    """Host-side plan of the separator split of the band LDL^T for a reduced system of n rows and half-bandwidth kd
    (ba_split_plan; pure arithmetic, runs without a GPU)."""
    global _PLAN_LIB
    if _PLAN_LIB is None:
        _PLAN_LIB = _lib.lib()
    out = np.zeros(24, dtype=np.int32)
    rc = _PLAN_LIB.ba_split_plan(int(n), int(kd), int(mode), int(segments), out.ctypes.data_as(C.POINTER(C.c_int)), 24)
    if rc != 0:
        raise BAError(f"ba_split_plan failed ({rc})")
    keys = ("ok", "w", "s0", "p1", "n0", "n1", "q0", "q1", "nm0", "nm1", "ntm0", "ntm1", "npE0", "npE1", "segments")
    d = {k: int(v) for k, v in zip(keys, out[:15])}
    d["bounds"] = [int(v) for v in out[15:15 + d["segments"] + 1]]
    return d


class GpuSolver:
    def __init__(self, prob, variant="QRCHOL", precision="f64", tau: float = 0.5, device: int = 0):
        self._L = _lib.lib()
        self._h = C.c_void_p()
        if not prob.is_sorted_by_point():
            raise BAError("observations must be sorted by (point, camera); use BALProblem.sorted_by_point()")
        self.variant = VARIANTS[variant] if isinstance(variant, str) else int(variant)
        self.precision = precision
        self.N, self.M, self.K = prob.N, prob.M, prob.K
        self.n = 3 * self.M + 9 * self.N
        view = np.ascontiguousarray(prob.view, dtype=np.int32)
        point = np.ascontiguousarray(prob.point, dtype=np.int32)
        meas = np.ascontiguousarray(prob.meas, dtype=np.float64)
        self._ck(self._L.ba_create(C.byref(self._h), self.N, self.M, self.K, _ip(view), _ip(point), _dp(meas),
                                   tau, F32 if precision == "f32" else F64, self.variant, device))
        self.set_state(prob.R, prob.T, prob.f, prob.k1, prob.k2, prob.X)

    def _ck(self, rc):
        if rc != 0:
            raise BAError(f"ba_gpu error {rc}: {self._L.ba_last_error().decode()}")

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._L.ba_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # --- state
    def set_state(self, R, T, f, k1, k2, X):
        # no copy when the caller hands in contiguous float64 (e.g. pinned) buffers
        arrs = [np.ascontiguousarray(a, dtype=np.float64).reshape(-1) for a in (R, T, f, k1, k2, X)]
        self._ck(self._L.ba_set_state(self._h, *[_dp(a) for a in arrs]))

    def get_state(self):
        N, M = self.N, self.M
        out = [np.empty(s) for s in (9 * N, 3 * N, N, N, N, 3 * M)]
        self._ck(self._L.ba_get_state(self._h, *[_dp(a) for a in out]))
        return out[0].reshape(N, 3, 3), out[1].reshape(N, 3), out[2], out[3], out[4], out[5].reshape(M, 3)

    # --- functor / solver calls
    def eval(self) -> float:
        e = C.c_double()
        self._ck(self._L.ba_eval(self._h, C.byref(e)))
        return e.value

    def linearize(self, colnorms: bool = True):
        e, a, b = C.c_double(), C.c_double(), C.c_double()
        if colnorms:
            self._ck(self._L.ba_linearize(self._h, C.byref(e), C.byref(a), C.byref(b)))
        else:
            self._ck(self._L.ba_linearize(self._h, C.byref(e), None, None))
        return e.value, a.value, b.value

    def compute(self, lam: float):
        self._ck(self._L.ba_compute(self._h, float(lam)))

    def solve_try(self):
        a, b, c = C.c_double(), C.c_double(), C.c_double()
        self._ck(self._L.ba_solve_try(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value  # |dx|, rho denominator, test energy

    def accept(self):
        self._ck(self._L.ba_accept(self._h))

    def reject(self):
        self._ck(self._L.ba_reject(self._h))

    def numeric_status(self) -> int:
        """0 fine; r > 0: zero/NaN pivot at row r of the reduced system in the last trial; -1: non-finite step."""
        v = C.c_int()
        self._ck(self._L.ba_numeric_status(self._h, C.byref(v)))
        return v.value

    def set_strict_numeric(self, flag=True):
        self._ck(self._L.ba_set_strict_numeric(self._h, int(flag)))

    # --- diagnostics
    def dx(self):
        v = np.empty(self.n)
        self._ck(self._L.ba_get_dx(self._h, _dp(v)))
        return v

    def step_streamed(self, R, T, f, k1, k2, X, lam, dx_out=None):
        """One trial from host state to host step (ba_step_streamed): upload, energy, compute(lam), solve_try and the step
        download in one call with the copies pipelined against the point stage. Pass contiguous (ideally pinned) float64
        arrays; returns (energy, |dx|, rho denominator, test energy) and fills dx_out (3M+9N) when given."""
        arrs = [np.ascontiguousarray(a, dtype=np.float64).reshape(-1) for a in (R, T, f, k1, k2, X)]
        if dx_out is not None:
            assert dx_out.dtype == np.float64 and dx_out.size == self.n and dx_out.flags["C_CONTIGUOUS"]
        e, a, b, c = C.c_double(), C.c_double(), C.c_double(), C.c_double()
        self._ck(self._L.ba_step_streamed(self._h, *[_dp(x) for x in arrs], float(lam), _dp(dx_out) if dx_out is not None else None,
                                          C.byref(e), C.byref(a), C.byref(b), C.byref(c)))
        return e.value, a.value, b.value, c.value

    def error_statistics(self, avg_focal_length: float, inlier_threshold: float):
        """Utils::showErrorStatistics / showObjective at the device state: (mean reprojection error, inlier mean error,
        number of inliers, true objective)."""
        out = np.zeros(4)
        self._ck(self._L.ba_error_statistics(self._h, float(avg_focal_length), float(inlier_threshold), _dp(out)))
        return out[0] / self.K, (out[1] / out[2] if out[2] > 0 else float("nan")), int(out[2]), out[3]

    def step_resident(self, lam, dx_out=None):
        """ba_step_streamed on the state already on the device: (energy, |dx|, rho denominator, test energy), one synchronisation."""
        if dx_out is not None:
            assert dx_out.dtype == np.float64 and dx_out.size == self.n and dx_out.flags["C_CONTIGUOUS"]
        e, a, b, c = C.c_double(), C.c_double(), C.c_double(), C.c_double()
        self._ck(self._L.ba_step_streamed(self._h, None, None, None, None, None, None, float(lam), _dp(dx_out) if dx_out is not None else None,
                                          C.byref(e), C.byref(a), C.byref(b), C.byref(c)))
        return e.value, a.value, b.value, c.value

    def dx_into(self, out):
        """Step download into a caller-owned (ideally pinned) float64 buffer of 3M+9N entries."""
        assert out.dtype == np.float64 and out.size == self.n and out.flags["C_CONTIGUOUS"]
        self._ck(self._L.ba_get_dx(self._h, _dp(out)))
        return out

    def residuals(self):
        r = np.empty(2 * self.K)
        self._ck(self._L.ba_get_residuals(self._h, _dp(r)))
        return r

    def jacobian(self):
        Jc, Jp = np.empty((self.K, 2, 9)), np.empty((self.K, 2, 3))
        self._ck(self._L.ba_get_jacobian(self._h, _dp(Jc), _dp(Jp)))
        return Jc, Jp

    def keep_reduced(self, flag=True):
        self._ck(self._L.ba_keep_reduced_system(self._h, int(flag)))

    def reduced(self):
        n = 9 * self.N
        S, g = np.empty((n, n)), np.empty(n)
        self._ck(self._L.ba_get_reduced_system(self._h, _dp(S), _dp(g)))
        return S, g

    def set_profiling(self, flag=True):
        self._ck(self._L.ba_set_profiling(self._h, int(flag)))

    def stage_ms(self):
        v = np.empty(8)
        self._ck(self._L.ba_stage_ms(self._h, _dp(v)))
        return v

    def debug_counters(self):
        v = (C.c_longlong * 16)()
        self._ck(self._L.ba_debug_counters(self._h, v))
        return list(v)

    def debug_counters_n(self, n=256):
        v = (C.c_longlong * n)()
        self._ck(self._L.ba_debug_counters_n(self._h, v, n))
        return list(v)

    def debug_band_solve(self, S, g, kd):
        S = np.ascontiguousarray(S, dtype=np.float64)
        g = np.ascontiguousarray(g, dtype=np.float64)
        y = np.empty(S.shape[0])
        self._ck(self._L.ba_debug_band_solve(self._h, S.shape[0], int(kd), _dp(S), _dp(g), _dp(y)))
        return y

    def timer_start(self):
        self._ck(self._L.ba_timer_start(self._h))

    def timer_stop(self) -> float:
        v = C.c_double()
        self._ck(self._L.ba_timer_stop(self._h, C.byref(v)))
        return v.value

    def launches(self) -> int:
        v = C.c_longlong()
        self._ck(self._L.ba_launch_count(self._h, C.byref(v)))
        return v.value

    @property
    def bandwidth(self) -> int:
        v = C.c_int()
        self._ck(self._L.ba_bandwidth(self._h, C.byref(v)))
        return v.value

    def set_bandwidth(self, bw: int):
        self._ck(self._L.ba_set_bandwidth(self._h, int(bw)))

    def comm_init(self, rank: int, nranks: int, uid: bytes):
        buf = C.create_string_buffer(uid, 128)
        self._ck(self._L.ba_comm_init(self._h, rank, nranks, buf))

    # --- LM control flow (host), QRChol.h:204-436 / More.h:204-425 / Cholesky.h:190-361
    def minimize(self, max_outer: int = 0, verbose: bool = False, on_trial=None):
        f32 = self.precision == "f32"
        S = np.float32 if f32 else np.float64
        lam_min, lam_max, inc_base, tol_fun = S(1e-10), S(1e10), S(2), S(1e-8)
        lam, lam_inc = S(1e-3), inc_base
        fun_evals, it = 0, 0
        hist = [S(0), S(0)]
        status = RUNNING
        log: List[Trial] = []
        while True:
            it += 1
            if it > 10 ** 6 or (max_outer > 0 and it > max_outer):
                status = MAX_ITERS
                break
            if fun_evals > 10 ** 6:
                status = TOO_MANY_FEVALS
                break
            e0, cn2, cn = self.linearize(colnorms=(it == 1))
            fun_evals += 1
            energy = S(e0)
            if it == 1:  # lambda_0 rule (QRChol.h:279, Cholesky.h:264, More.h:284; QRKIT assumed = QRCHOL)
                lam = S(1e-6 * float(S(cn))) if self.variant == MOREQR else S(1e-12 * float(S(cn2)))
            stop = False
            while True:
                self.compute(float(lam))
                dxn, rho_den, e_test = self.solve_try()
                fun_evals += 1
                e_test = S(e_test)
                if e_test < energy:
                    rho = (energy - e_test) / S(rho_den)
                    mul = S(1.0) - (S(2.0) * rho - S(1.0)) ** 3
                    lam_used = lam
                    lam = lam * max(S(1.0) / S(3.0), S(mul))
                    lam = max(lam, lam_min)
                    log.append(Trial(it, True, float(energy), float(e_test), float(rho), float(lam_used), float(lam), dxn))
                    lam_inc = inc_base
                    energy = e_test
                    hist[it % 2] = energy
                    if on_trial:
                        on_trial(log[-1])
                    break
                else:
                    log.append(Trial(it, False, float(energy), float(e_test), 0.0, float(lam), float(lam), dxn))
                    self.reject()
                    if on_trial:
                        on_trial(log[-1])
                    if lam > lam_max:
                        status = EXCEEDED_LAMBDA_MAX
                        stop = True
                        break
                    lam = S(lam * lam_inc)
                    lam_inc = S(math.pow(float(lam_inc), 1.5))
            if stop:
                break
            if it > 2:
                maxf = max(hist)
                if abs(energy - maxf) < tol_fun * energy:
                    status = SUCCESS  # x NOT committed (quirk Q8)
                    break
            self.accept()
        return status, log
