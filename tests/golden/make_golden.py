"""Regenerates / cross-checks tests/golden/survey_anchors.json with an INDEPENDENT NumPy restatement of the model.

Nothing here touches the C++ oracle, the CUDA library or the reference: residuals are plain NumPy (BAFunctor.h:160-178 as
restated in SURVEY.md App. A), the Jacobian comes from complex-step differentiation of that residual (no hand-written
derivative shared with oracle/ or csrc/), and the first LM step solves the full sparse normal equations
(J^T J + lambda I) dx = -J^T e with SciPy instead of any Schur-complement code. Only the BAL file parser is shared (bal.py).

  python tests/golden/make_golden.py            # cross-check: prints the relative difference to every stored anchor
  python tests/golden/make_golden.py --write    # rewrite the anchors from this script's values

The reference itself cannot be built or run in this environment (its Eigen fork and SuiteSparse are absent) and ships no golden
vectors, so these anchors pin the oracle against an independent derivation, not against reference output.
"""
import json
import os
import sys

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from bundleadjustment_benchmarks_b200 import bal  # noqa: E402

TAU2 = bal.INLIER_THRESHOLD ** 2
PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "survey_anchors.json")


def project(R, T, f, k1, k2, X):
    """f * dist(pi(R X + T)); all arrays per observation, any dtype (complex for the complex step)."""
    XX = np.einsum("kij,kj->ki", R, X) + T
    xu = XX[:, :2] / XX[:, 2:3]
    r2 = (xu * xu).sum(axis=1)
    return (f * (1 + k1 * r2 + k2 * r2 * r2))[:, None] * xu


def residual(R, T, f, k1, k2, X, m):
    """e = r / |r| * sqrt(psi(|r|^2)), psi(r2) = r2 (2 - r2 / tau2) / 4 inside the threshold, tau2 / 4 outside."""
    r = project(R, T, f, k1, k2, X) - m
    r2 = (r * r).sum(axis=1)
    psi = np.where(r2.real < TAU2, r2 * (2 - r2 / TAU2) / 4, TAU2 / 4 + 0 * r2)
    n = np.sqrt(r2)
    n = np.where(n.real > 1e-15, n, 1e-15)
    return r * (np.sqrt(psi) / n)[:, None]


def cross(w):
    C = np.zeros(w.shape[:-1] + (3, 3), dtype=w.dtype)
    C[..., 0, 1], C[..., 0, 2] = -w[..., 2], w[..., 1]
    C[..., 1, 0], C[..., 1, 2] = w[..., 2], -w[..., 0]
    C[..., 2, 0], C[..., 2, 1] = -w[..., 1], w[..., 0]
    return C


def jacobian(p):
    """(K, 2, 9) camera block over (T, omega, f, k1, k2) and (K, 2, 3) point block by the complex step (h = 1e-30)."""
    v, q, h = p.view, p.point, 1e-30
    R, T, f, k1, k2, X, m = p.R[v].astype(complex), p.T[v].astype(complex), p.f[v].astype(complex), p.k1[v].astype(complex), \
        p.k2[v].astype(complex), p.X[q].astype(complex), np.asarray(p.meas).reshape(-1, 2)
    Jc, Jp = np.zeros((p.K, 2, 9)), np.zeros((p.K, 2, 3))
    for a in range(3):
        d = np.zeros((p.K, 3), dtype=complex); d[:, a] = 1j * h
        Jc[:, :, a] = residual(R, T + d, f, k1, k2, X, m).imag / h
        Jc[:, :, 3 + a] = residual(R + cross(d) @ R, T, f, k1, k2, X, m).imag / h     # R <- exp([w]x) R, first order
        Jp[:, :, a] = residual(R, T, f, k1, k2, X + d, m).imag / h
    Jc[:, :, 6] = residual(R, T, f + 1j * h, k1, k2, X, m).imag / h
    Jc[:, :, 7] = residual(R, T, f, k1 + 1j * h, k2, X, m).imag / h
    Jc[:, :, 8] = residual(R, T, f, k1, k2 + 1j * h, X, m).imag / h
    return Jc, Jp


def anchors(name, with_step):
    p = bal.load_named(name)
    v, q = p.view, p.point
    m = np.asarray(p.meas).reshape(-1, 2)
    e = residual(p.R[v], p.T[v], p.f[v], p.k1[v], p.k2[v], p.X[q], m)
    out = {"N": int(p.N), "M": int(p.M), "K": int(p.K), "initial_energy": float((e * e).sum())}
    Jc, Jp = jacobian(p)
    rows = (2 * np.arange(p.K)[:, None, None] + np.arange(2)[None, :, None])
    J = sp.coo_matrix((np.concatenate([Jp.ravel(), Jc.ravel()]),
                       (np.concatenate([np.broadcast_to(rows, Jp.shape).ravel(), np.broadcast_to(rows, Jc.shape).ravel()]),
                        np.concatenate([(3 * q[:, None, None] + np.arange(3)[None, None, :] + 0 * rows).ravel(),
                                        (3 * p.M + 9 * v[:, None, None] + np.arange(9)[None, None, :] + 0 * rows).ravel()]))),
                      shape=(2 * p.K, 3 * p.M + 9 * p.N)).tocsr()
    cn2 = np.asarray(J.multiply(J).sum(axis=0)).ravel().max()
    out["max_colnorm2"], out["max_colnorm"] = float(cn2), float(np.sqrt(cn2))
    err = bal.AVG_FOCAL_LENGTH * np.linalg.norm(project(p.R[v], p.T[v], p.f[v], p.k1[v], p.k2[v], p.X[q]) - m, axis=1)
    out["initial_mean_reproj_px"], out["initial_inliers"] = float(err.mean()), int((err <= bal.INLIER_THRESHOLD).sum())
    if with_step:
        lam = 1e-12 * cn2
        out["lambda0_qrchol"] = float(lam)
        A = (J.T @ J + lam * sp.identity(J.shape[1])).tocsc()
        dx = spla.spsolve(A, -(J.T @ e.ravel()))
        out["iter1_dx_norm"] = float(np.linalg.norm(dx))
        dX, dc = dx[:3 * p.M].reshape(-1, 3), dx[3 * p.M:].reshape(-1, 9)
        R1 = bal.rodrigues(dc[:, 3:6]) @ p.R                       # update_params, BAFunctor.h:311-333
        e1 = residual(R1[v], (p.T + dc[:, :3])[v], (p.f + dc[:, 6])[v], (p.k1 + dc[:, 7])[v], (p.k2 + dc[:, 8])[v], (p.X + dX)[q], m)
        out["iter1_energy_test"] = float((e1 * e1).sum())
    return out


if __name__ == "__main__":
    stored = json.load(open(PATH))
    fresh = {"_comment": stored["_comment"]}
    for name in ("problem-21-11315", "problem-39-18060"):
        fresh[name] = anchors(name, "iter1_dx_norm" in stored.get(name, {}))
        for k, val in fresh[name].items():
            old = stored.get(name, {}).get(k)
            if old is None:
                print(f"{name:18s} {k:24s} {val!r} (new)")
            elif isinstance(val, int):
                print(f"{name:18s} {k:24s} {val} stored {old} {'ok' if val == old else 'DIFFERENT'}")
            else:
                print(f"{name:18s} {k:24s} {val:.12g} stored {old:.12g} rel diff {abs(val - old) / abs(old):.1e}")
    if "--write" in sys.argv:
        json.dump(fresh, open(PATH, "w"), indent=2)
        print("written", PATH)
