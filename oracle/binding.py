"""TEST INFRASTRUCTURE — ctypes binding of oracle/libba_oracle.so (CPU restatement, parity unpinned;
see oracle/ba_oracle.hpp). Never imported by the product package."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

QRKIT, QRCHOL, MOREQR, CHOLESKY = 0, 1, 2, 3
VARIANTS = {"QRKIT": 0, "QRCHOL": 1, "MOREQR": 2, "CHOLESKY": 3}


class TrialRecord(C.Structure):
    _fields_ = [("iter", C.c_int), ("accepted", C.c_int), ("energy", C.c_double),
                ("energy_test", C.c_double), ("rho", C.c_double), ("lambda_used", C.c_double),
                ("lambda_next", C.c_double), ("dx_norm", C.c_double)]


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libba_oracle.so")
    srcs = [os.path.join(_HERE, n) for n in ("ba_oracle_capi.cpp", "ba_oracle.hpp")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libba_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        dp = C.POINTER(C.c_double)
        ip = C.POINTER(C.c_int)
        L.bao_create.restype = C.c_void_p
        L.bao_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, ip, ip, dp, C.c_double]
        L.bao_destroy.argtypes = [C.c_void_p]
        L.bao_set_state.argtypes = [C.c_void_p] + [dp] * 6
        L.bao_get_state.argtypes = [C.c_void_p] + [dp] * 6
        L.bao_linearize.argtypes = [C.c_void_p, dp, dp, dp]
        L.bao_residuals.argtypes = [C.c_void_p, dp]
        L.bao_jacobian.argtypes = [C.c_void_p, dp, dp]
        L.bao_jtres.argtypes = [C.c_void_p, dp]
        L.bao_moreqr_outer.argtypes = [C.c_void_p]
        L.bao_step.restype = C.c_int
        L.bao_step.argtypes = [C.c_void_p, C.c_int, C.c_double, dp]
        L.bao_energy_at.restype = C.c_double
        L.bao_energy_at.argtypes = [C.c_void_p, dp]
        L.bao_apply.argtypes = [C.c_void_p, dp]
        L.bao_reduced.argtypes = [C.c_void_p, dp, dp]
        L.bao_set_tall.argtypes = [C.c_void_p, C.c_int]
        L.bao_set_threads.argtypes = [C.c_void_p, C.c_int]
        L.bao_has_openmp.restype = C.c_int
        L.bao_kd.restype = C.c_int
        L.bao_kd.argtypes = [C.c_void_p]
        L.bao_minimize.restype = C.c_int
        L.bao_minimize.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(TrialRecord), C.c_int, ip]
        _LIB = L
    return _LIB


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


class Oracle:
    """CPU oracle on one BALProblem (observations must be sorted by point)."""

    def __init__(self, prob, precision: str = "f64", tau: float = 0.5):
        assert prob.is_sorted_by_point()
        self.N, self.M, self.K = prob.N, prob.M, prob.K
        self.n = 3 * self.M + 9 * self.N
        self._L = lib()
        view = np.ascontiguousarray(prob.view, dtype=np.int32)
        point = np.ascontiguousarray(prob.point, dtype=np.int32)
        meas = np.ascontiguousarray(prob.meas, dtype=np.float64)
        self._h = C.c_void_p(self._L.bao_create(0 if precision == "f32" else 1, self.N, self.M, self.K,
                                                _ip(view), _ip(point), _dp(meas), tau))
        self.set_state(prob.R, prob.T, prob.f, prob.k1, prob.k2, prob.X)

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.bao_destroy(self._h)
            self._h = None

    def set_state(self, R, T, f, k1, k2, X):
        arrs = [np.ascontiguousarray(a, dtype=np.float64).reshape(-1) for a in (R, T, f, k1, k2, X)]
        self._L.bao_set_state(self._h, *[_dp(a) for a in arrs])

    def get_state(self):
        N, M = self.N, self.M
        out = [np.empty(s) for s in (9 * N, 3 * N, N, N, N, 3 * M)]
        self._L.bao_get_state(self._h, *[_dp(a) for a in out])
        return out[0].reshape(N, 3, 3), out[1].reshape(N, 3), out[2], out[3], out[4], out[5].reshape(M, 3)

    def linearize(self):
        e, a, b = C.c_double(), C.c_double(), C.c_double()
        self._L.bao_linearize(self._h, C.byref(e), C.byref(a), C.byref(b))
        return e.value, a.value, b.value

    def residuals(self):
        r = np.empty(2 * self.K)
        self._L.bao_residuals(self._h, _dp(r))
        return r

    def jacobian(self):
        Jc, Jp = np.empty((self.K, 2, 9)), np.empty((self.K, 2, 3))
        self._L.bao_jacobian(self._h, _dp(Jc), _dp(Jp))
        return Jc, Jp

    def jtres(self):
        v = np.empty(self.n)
        self._L.bao_jtres(self._h, _dp(v))
        return v

    def moreqr_outer(self):
        self._L.bao_moreqr_outer(self._h)

    def step(self, variant: int, lam: float):
        dx = np.empty(self.n)
        ok = self._L.bao_step(self._h, variant, lam, _dp(dx))
        return bool(ok), dx

    def energy_at(self, dx):
        dx = np.ascontiguousarray(dx, dtype=np.float64)
        return self._L.bao_energy_at(self._h, _dp(dx))

    def apply(self, dx):
        dx = np.ascontiguousarray(dx, dtype=np.float64)
        self._L.bao_apply(self._h, _dp(dx))

    def reduced(self):
        n = 9 * self.N
        S, g = np.empty((n, n)), np.empty(n)
        self._L.bao_reduced(self._h, _dp(S), _dp(g))
        return S, g

    def set_tall(self, flag: bool):
        self._L.bao_set_tall(self._h, int(flag))

    def has_openmp(self) -> bool:
        return bool(self._L.bao_has_openmp())

    def set_threads(self, n: int):
        """Timing legs only: OpenMP over observations / points and in the reduced solve (default 1, like the reference)."""
        self._L.bao_set_threads(self._h, int(n))

    @property
    def kd(self):
        return self._L.bao_kd(self._h)

    def minimize(self, variant: int, max_outer: int = 0, cap: int = 100000):
        log = (TrialRecord * cap)()
        n = C.c_int()
        st = self._L.bao_minimize(self._h, variant, max_outer, log, cap, C.byref(n))
        return st, [log[i] for i in range(min(n.value, cap))]
