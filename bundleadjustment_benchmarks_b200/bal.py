"""BAL ("Bundle Adjustment in the Large") problem files and BAL-shaped synthetic problems.

Mirrors the reference's loader semantics (reference: src/bundle_adjustment_large.cpp:57-107):

* header ``N M K``; K lines ``cam pt x y``; then 9 scalars per camera (omega[3], T[3], f, k1, k2),
  one per line; then 3 scalars per point.
* the focal length is stored NEGATED (``K(0,0) = K(1,1) = -f``, :88-90),
* the rotation is converted once from the angle-axis vector with the reference's Rodrigues rule and
  its hard ``|omega| > 1e-6`` cut-off (src/MathUtils.h:66-82),
* the radial distortion coefficients are pre-scaled: ``k1*f^2``, ``k2*f^4`` (:97-98).

The LM state is therefore (R[3x3], T, f, k1, k2) per camera and X per point, NOT an angle-axis vector.

The solver kernels need observations grouped by point (the reference's row
permutation silently relies on the files being sorted by point). ``BALProblem.sorted_by_point``
performs a stable sort and keeps the permutation.
"""
from __future__ import annotations

import dataclasses
import gzip
import os
from typing import Optional

import numpy as np

SEED = 20261018
INLIER_THRESHOLD = 0.5  # reference: src/bundle_adjustment_large.cpp:36
AVG_FOCAL_LENGTH = 1.0  # reference: src/bundle_adjustment_large.cpp:35


def rodrigues(omega: np.ndarray) -> np.ndarray:
    """Batch Rodrigues, reference rule (src/MathUtils.h:66-82): identity when |omega| <= 1e-6.

    omega: (..., 3) -> (..., 3, 3) row-major rotation matrices.
    """
    omega = np.asarray(omega, dtype=np.float64)
    theta = np.linalg.norm(omega, axis=-1)
    J = np.zeros(omega.shape[:-1] + (3, 3), dtype=np.float64)
    J[..., 0, 1] = -omega[..., 2]
    J[..., 0, 2] = omega[..., 1]
    J[..., 1, 0] = omega[..., 2]
    J[..., 1, 2] = -omega[..., 0]
    J[..., 2, 0] = -omega[..., 1]
    J[..., 2, 1] = omega[..., 0]
    J2 = J @ J
    big = theta > 1e-6
    th = np.where(big, theta, 1.0)
    c1 = np.where(big, np.sin(th) / th, 0.0)
    c2 = np.where(big, (1.0 - np.cos(th)) / (th * th), 0.0)
    R = np.broadcast_to(np.eye(3), J.shape).copy()
    R = R + c1[..., None, None] * J + c2[..., None, None] * J2
    return R


@dataclasses.dataclass
class BALProblem:
    """One bundle-adjustment problem in the reference's in-memory convention.

    view/point: int32[K]; meas: float64[K,2]; R: [N,3,3]; T: [N,3]; f (negative), k1, k2: [N];
    X: [M,3].
    """

    view: np.ndarray
    point: np.ndarray
    meas: np.ndarray
    R: np.ndarray
    T: np.ndarray
    f: np.ndarray
    k1: np.ndarray
    k2: np.ndarray
    X: np.ndarray
    name: str = "unnamed"
    perm: Optional[np.ndarray] = None  # observation permutation applied by sorted_by_point()

    @property
    def N(self) -> int:
        return int(self.R.shape[0])

    @property
    def M(self) -> int:
        return int(self.X.shape[0])

    @property
    def K(self) -> int:
        return int(self.view.shape[0])

    def copy(self) -> "BALProblem":
        return dataclasses.replace(
            self,
            view=self.view.copy(), point=self.point.copy(), meas=self.meas.copy(),
            R=self.R.copy(), T=self.T.copy(), f=self.f.copy(), k1=self.k1.copy(),
            k2=self.k2.copy(), X=self.X.copy(),
            perm=None if self.perm is None else self.perm.copy())

    def is_sorted_by_point(self) -> bool:
        key = self.point.astype(np.int64) * (self.N + 1) + self.view.astype(np.int64)
        return bool(np.all(key[1:] >= key[:-1]))

    def sorted_by_point(self) -> "BALProblem":
        """Stable sort of observations by (point, camera); remembers the permutation."""
        if self.is_sorted_by_point():
            return self
        key = self.point.astype(np.int64) * (self.N + 1) + self.view.astype(np.int64)
        perm = np.argsort(key, kind="stable")
        out = self.copy()
        out.view = np.ascontiguousarray(self.view[perm])
        out.point = np.ascontiguousarray(self.point[perm])
        out.meas = np.ascontiguousarray(self.meas[perm])
        out.perm = perm
        return out

    def point_offsets(self) -> np.ndarray:
        """CSR offsets int32[M+1] of each point's observation run (requires sorted)."""
        counts = np.bincount(self.point, minlength=self.M)
        off = np.zeros(self.M + 1, dtype=np.int64)
        np.cumsum(counts, out=off[1:])
        return off.astype(np.int32)

    def validate(self) -> None:
        if self.K == 0:
            return
        if self.view.min() < 0 or self.view.max() >= self.N:
            raise ValueError("camera index out of range")
        if self.point.min() < 0 or self.point.max() >= self.M:
            raise ValueError("point index out of range")
        counts = np.bincount(self.point, minlength=self.M)
        if counts.min() < 1:
            raise ValueError("every point needs at least one observation (SURVEY.md App. F Q11)")


def read_bal(path: str) -> BALProblem:
    """Parse a BAL text file with the reference's conventions (bundle_adjustment_large.cpp:57-107)."""
    opener = gzip.open if path.endswith(".gz") else open
    with opener(path, "rb") as fh:
        header = fh.readline().split()
        N, M, K = int(header[0]), int(header[1]), int(header[2])
        rest = np.array(fh.read().split(), dtype=np.float64)
    need = 4 * K + 9 * N + 3 * M
    if rest.size < need:
        raise ValueError(f"{path}: expected {need} numbers after the header, found {rest.size}")
    obs = rest[: 4 * K].reshape(K, 4)
    view = obs[:, 0].astype(np.int32)
    point = obs[:, 1].astype(np.int32)
    meas = np.ascontiguousarray(obs[:, 2:4]) / AVG_FOCAL_LENGTH
    cam = rest[4 * K: 4 * K + 9 * N].reshape(N, 9)
    X = np.ascontiguousarray(rest[4 * K + 9 * N: need].reshape(M, 3))
    return from_file_params(view, point, meas, cam, X, name=os.path.basename(path))


def from_file_params(view, point, meas, cam9, X, name="unnamed") -> BALProblem:
    """cam9[N,9] in FILE units (omega, T, f, k1, k2) -> reference in-memory state."""
    cam9 = np.asarray(cam9, dtype=np.float64)
    f_file = cam9[:, 6]
    f2 = f_file * f_file
    prob = BALProblem(
        view=np.ascontiguousarray(view, dtype=np.int32),
        point=np.ascontiguousarray(point, dtype=np.int32),
        meas=np.ascontiguousarray(meas, dtype=np.float64),
        R=rodrigues(cam9[:, 0:3]),
        T=np.ascontiguousarray(cam9[:, 3:6]),
        f=-f_file / AVG_FOCAL_LENGTH,
        k1=cam9[:, 7] * f2,
        k2=cam9[:, 8] * f2 * f2,
        X=np.ascontiguousarray(X, dtype=np.float64),
        name=name,
    )
    prob.validate()
    return prob


def write_bal(path: str, view, point, meas, cam9, X) -> None:
    """Write BAL text in the same shape as data/*.txt (used for CLI tests and stand-in files)."""
    K, N, M = len(view), len(cam9), len(X)
    with open(path, "w") as fh:
        fh.write(f"{N} {M} {K}\n")
        for i in range(K):
            fh.write(f"{int(view[i])} {int(point[i])}     {meas[i, 0]:.6e} {meas[i, 1]:.6e}\n")
        for v in np.asarray(cam9).reshape(-1):
            fh.write(f"{v:.16e}\n")
        for v in np.asarray(X).reshape(-1):
            fh.write(f"{v:.16e}\n")


# ----------------------------------------------------------------------------------------------
# Synthetic BAL-shaped problems (SURVEY.md §8(d) generator spec)
# ----------------------------------------------------------------------------------------------

def _project_file_units(cam9, X, view, point):
    """BAL camera model in file units: p = -P.xy/P.z ; f * (1 + k1 p^2 + k2 p^4) * p."""
    R = rodrigues(cam9[:, 0:3])
    P = np.einsum("kij,kj->ki", R[view], X[point]) + cam9[view, 3:6]
    p = -P[:, 0:2] / P[:, 2:3]
    n2 = np.sum(p * p, axis=1)
    kr = 1.0 + cam9[view, 7] * n2 + cam9[view, 8] * n2 * n2
    return cam9[view, 6][:, None] * kr[:, None] * p, P[:, 2]


def synthetic_file_arrays(N: int, M: int, K: Optional[int] = None, *, mean_obs: float = 5.0,
                          window: int = 30, seed: int = SEED, outlier_frac: float = 0.2,
                          noise_px: float = 0.3):
    """Seeded BAL-shaped problem in FILE units: returns (view, point, meas, cam9, X).

    Cameras sit on a smooth path looking down -z (camera-space depth negative, as in the bundled
    data); a point is seen by n_j = 2 + Poisson(mean_obs - 2) cameras (>= 2 observations each, every
    point index occurs) drawn from a window of +-``window`` cameras around the nearest one, so the
    reduced camera matrix is block-banded for N >> window. If ``K`` is given the counts are nudged
    so that sum n_j == K exactly (stand-ins for the missing bundled files keep their (N, M, K)).
    Observations are sorted by point, then camera.
    """
    rng = np.random.default_rng(seed)
    spacing = 0.25
    wsize = min(N, 2 * window + 1)

    # --- truth ---
    cx = np.arange(N) * spacing
    centers = np.stack([cx, 0.3 * np.sin(cx * 0.7), 0.2 * np.cos(cx * 0.4)], axis=1)
    omega = rng.normal(0.0, 0.03, size=(N, 3))
    Rt = rodrigues(omega)
    Tt = -np.einsum("nij,nj->ni", Rt, centers)
    f = rng.uniform(1300.0, 2500.0, size=N)
    k1 = rng.normal(0.0, 3e-8, size=N)
    k2 = rng.normal(0.0, 1e-14, size=N)
    cam_true = np.concatenate([omega, Tt, f[:, None], k1[:, None], k2[:, None]], axis=1)

    px = np.sort(rng.uniform(0.0, max(N - 1, 1) * spacing, size=M))
    X_true = np.stack([px, rng.uniform(-4.0, 4.0, size=M), rng.uniform(-40.0, -20.0, size=M)],
                      axis=1)

    # --- visibility ---
    lam = max(mean_obs - 2.0, 0.0)
    n = 2 + rng.poisson(lam, size=M)
    n = np.minimum(n, wsize)
    if K is not None:
        if not (2 * M <= K <= wsize * M):
            raise ValueError("K outside [2M, window*M]")
        diff = int(K - n.sum())
        while diff != 0:
            step = 1 if diff > 0 else -1
            ok = np.flatnonzero((n < wsize) if step > 0 else (n > 2))
            take = rng.choice(ok, size=min(abs(diff), ok.size), replace=False)
            n[take] += step
            diff = int(K - n.sum())
    Ktot = int(n.sum())
    c0 = np.rint(px / spacing).astype(np.int64)
    lo = np.clip(c0 - window, 0, N - wsize)

    off = np.zeros(M + 1, dtype=np.int64)
    np.cumsum(n, out=off[1:])
    point = np.repeat(np.arange(M, dtype=np.int64), n)
    view = np.empty(Ktot, dtype=np.int64)
    chunk = 1 << 16
    for s in range(0, M, chunk):
        e = min(M, s + chunk)
        keys = rng.random((e - s, wsize))
        order = np.argsort(keys, axis=1)
        nn = n[s:e]
        mask = np.arange(wsize)[None, :] < nn[:, None]
        sel = np.where(mask, order, wsize)  # pad with sentinel
        sel.sort(axis=1)
        picked = sel[mask]
        view[off[s]:off[e]] = picked + np.repeat(lo[s:e], nn)

    # --- measurements ---
    meas, depth = _project_file_units(cam_true, X_true, view, point)
    assert np.all(depth < 0), "synthetic scene must stay in front of the cameras (negative depth)"
    meas = meas + rng.normal(0.0, noise_px, size=meas.shape)
    bad = rng.random(Ktot) < outlier_frac
    meas[bad] += rng.uniform(-30.0, 30.0, size=(int(bad.sum()), 2))

    # --- initial parameters = truth + perturbation ---
    cam0 = cam_true.copy()
    cam0[:, 0:3] += rng.normal(0.0, 1e-3, size=(N, 3))
    cam0[:, 3:6] += rng.normal(0.0, 1e-2, size=(N, 3))
    cam0[:, 6] *= 1.0 + rng.normal(0.0, 1e-3, size=N)
    X0 = X_true + rng.normal(0.0, 5e-2, size=X_true.shape)
    return view.astype(np.int32), point.astype(np.int32), meas, cam0, X0


def synthetic(N: int, M: int, K: Optional[int] = None, **kw) -> BALProblem:
    view, point, meas, cam9, X = synthetic_file_arrays(N, M, K, **kw)
    return from_file_params(view, point, meas, cam9, X, name=f"synthetic-{N}-{M}-{len(view)}")


# Named configurations of BASELINE.json / BASELINE.md §3. Missing bundled files get seeded stand-ins
# with identical (N, M, K).
STANDINS = {
    "problem-16-22106": (16, 22106, 83718),
    "problem-126-40037": (126, 40037, 148117),
    "problem-257-65132": (257, 65132, 225911),
}


def load_named(name: str, data_dir: Optional[str] = None) -> BALProblem:
    """'problem-21-11315' etc.: the real file if present under data_dir, else a seeded stand-in;
    'synthetic-5m' is BASELINE config 5; 'synthetic-N-M[-K]' builds an arbitrary one."""
    data_dir = data_dir or os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "data")
    for ext in ("-pre.txt", "-pre.txt.gz"):
        p = os.path.join(data_dir, name + ext)
        if os.path.exists(p):
            return read_bal(p)
    if name in STANDINS:
        N, M, K = STANDINS[name]
        prob = synthetic(N, M, K, window=max(30, N))
        prob.name = name + "-standin"
        return prob
    if name == "synthetic-5m":
        return synthetic(1800, 1_000_000, None, mean_obs=5.0, window=30)
    if name.startswith("synthetic-"):
        parts = [int(t) for t in name.split("-")[1:]]
        return synthetic(parts[0], parts[1], parts[2] if len(parts) > 2 else None)
    raise FileNotFoundError(name)
