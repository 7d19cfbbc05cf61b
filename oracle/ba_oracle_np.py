"""TEST INFRASTRUCTURE — independent NumPy restatement used to pin the C++ oracle.

Nothing here shares code or derivations with oracle/ba_oracle.hpp: the Jacobian comes from
forward-mode dual numbers pushed through the model equations (reference model: BAFunctor.h:151-178,
DistortionFunction.cpp:14-23, update rule BAFunctor.h:311-338), the LM step from a dense
extended-precision (np.longdouble) Householder least-squares solve of [J; sqrt(lambda) I] dx = -[r; 0].
Pure NumPy; meant for small problems only.
"""
from __future__ import annotations

import numpy as np


class Dual:
    """value (K,), derivative (K, P)."""
    __array_priority__ = 1000

    def __init__(self, v, d):
        self.v, self.d = v, d

    @staticmethod
    def const(v, P):
        v = np.asarray(v)
        return Dual(v, np.zeros(v.shape + (P,), dtype=v.dtype))

    def _lift(self, o):
        return o if isinstance(o, Dual) else Dual.const(np.broadcast_to(np.asarray(o, dtype=self.v.dtype), self.v.shape), self.d.shape[-1])

    def __add__(self, o):
        o = self._lift(o); return Dual(self.v + o.v, self.d + o.d)
    __radd__ = __add__

    def __sub__(self, o):
        o = self._lift(o); return Dual(self.v - o.v, self.d - o.d)

    def __rsub__(self, o):
        o = self._lift(o); return Dual(o.v - self.v, o.d - self.d)

    def __mul__(self, o):
        o = self._lift(o); return Dual(self.v * o.v, self.d * o.v[..., None] + o.d * self.v[..., None])
    __rmul__ = __mul__

    def __truediv__(self, o):
        o = self._lift(o); return Dual(self.v / o.v, (self.d * o.v[..., None] - o.d * self.v[..., None]) / (o.v * o.v)[..., None])

    def __rtruediv__(self, o):
        return self._lift(o) / self

    def sqrt(self):
        s = np.sqrt(self.v); return Dual(s, self.d / (2 * s)[..., None])


def residual_and_jacobian(prob, tau=0.5, dtype=np.float64):
    """Returns e (K,2), Jc (K,2,9), Jp (K,2,3) with the reference's column order (T, omega, f, k1, k2 | X)."""
    K, P = prob.K, 12
    R = prob.R[prob.view].astype(dtype); T = prob.T[prob.view].astype(dtype)
    X = prob.X[prob.point].astype(dtype)
    one = np.ones(K, dtype=dtype)

    def seed(v, idx):
        d = np.zeros((K, P), dtype=dtype); d[:, idx] = 1; return Dual(v.astype(dtype), d)

    Tt = [seed(T[:, i], i) for i in range(3)]
    w = [seed(np.zeros(K), 3 + i) for i in range(3)]
    f = seed(prob.f[prob.view], 6); k1 = seed(prob.k1[prob.view], 7); k2 = seed(prob.k2[prob.view], 8)
    Xp = [seed(X[:, i], 9 + i) for i in range(3)]
    RX = [R[:, i, 0] * Xp[0] + R[:, i, 1] * Xp[1] + R[:, i, 2] * Xp[2] for i in range(3)]
    # first-order left-multiplicative rotation update: (I + [w]_x) R X
    XX = [RX[0] + (w[1] * RX[2] - w[2] * RX[1]) + Tt[0],
          RX[1] + (w[2] * RX[0] - w[0] * RX[2]) + Tt[1],
          RX[2] + (w[0] * RX[1] - w[1] * RX[0]) + Tt[2]]
    xu = [XX[0] / XX[2], XX[1] / XX[2]]
    r2u = xu[0] * xu[0] + xu[1] * xu[1]
    kr = one + k1 * r2u + k2 * (r2u * r2u)
    q = [f * (kr * xu[0]), f * (kr * xu[1])]
    r = [q[0] - prob.meas[:, 0].astype(dtype), q[1] - prob.meas[:, 1].astype(dtype)]
    r2 = r[0] * r[0] + r[1] * r[1]
    tau2 = dtype(tau) * dtype(tau)
    inl = r2.v < tau2
    psi_in = r2 * (2.0 - r2 / tau2) / 4.0
    psi = Dual(np.where(inl, psi_in.v, tau2 / 4), np.where(inl[:, None], psi_in.d, 0))
    s = psi.sqrt() / r2.sqrt()
    e = [r[0] * s, r[1] * s]
    ev = np.stack([e[0].v, e[1].v], axis=1)
    J = np.stack([e[0].d, e[1].d], axis=1)
    return ev, J[:, :, 0:9], J[:, :, 9:12]


def dense_jacobian(prob, Jc, Jp):
    n = 3 * prob.M + 9 * prob.N
    J = np.zeros((2 * prob.K, n), dtype=Jc.dtype)
    for i in range(prob.K):
        c, p = int(prob.view[i]), int(prob.point[i])
        J[2 * i:2 * i + 2, 3 * p:3 * p + 3] = Jp[i]
        J[2 * i:2 * i + 2, 3 * prob.M + 9 * c:3 * prob.M + 9 * c + 9] = Jc[i]
    return J


def householder_lstsq(A, b):
    """min |A x - b| by Householder QR in A's dtype (use np.longdouble for the extended-precision check)."""
    A = A.copy(); b = b.copy()
    m, n = A.shape
    for k in range(n):
        x = A[k:, k]
        nrm = np.sqrt((x * x).sum())
        if nrm == 0:
            continue
        alpha = -nrm if x[0] >= 0 else nrm
        v = x.copy(); v[0] -= alpha
        vv = (v * v).sum()
        if vv == 0:
            continue
        A[k:, k:] -= np.outer(v, (2 / vv) * (v @ A[k:, k:]))
        b[k:] -= v * ((2 / vv) * (v @ b[k:]))
    x = np.zeros(n, dtype=A.dtype)
    for i in range(n - 1, -1, -1):
        x[i] = (b[i] - A[i, i + 1:n] @ x[i + 1:]) / A[i, i]
    return x


def lm_step_extended(prob, lam, tau=0.5):
    """dx = argmin |J dx + r|^2 + lam |dx|^2 in np.longdouble (the unique LM step)."""
    e, Jc, Jp = residual_and_jacobian(prob, tau, dtype=np.longdouble)
    J = dense_jacobian(prob, Jc, Jp)
    n = J.shape[1]
    A = np.vstack([J, np.sqrt(np.longdouble(lam)) * np.eye(n, dtype=np.longdouble)])
    b = np.concatenate([-e.reshape(-1), np.zeros(n, dtype=np.longdouble)])
    return householder_lstsq(A, b).astype(np.float64)


def apply_update(prob, dx):
    """update_params (BAFunctor.h:299-342) on a copy."""
    from bundleadjustment_benchmarks_b200.bal import rodrigues
    out = prob.copy()
    M = prob.M
    d = dx[3 * M:].reshape(prob.N, 9)
    out.T = prob.T + d[:, 0:3]
    out.R = rodrigues(d[:, 3:6]) @ prob.R
    out.f = prob.f + d[:, 6]; out.k1 = prob.k1 + d[:, 7]; out.k2 = prob.k2 + d[:, 8]
    out.X = prob.X + dx[:3 * M].reshape(M, 3)
    return out


def energy(prob, tau=0.5):
    e, _, _ = residual_and_jacobian(prob, tau)
    return float((e * e).sum())
