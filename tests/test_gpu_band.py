"""GPU (-m gpu): the reduced-camera-block solver in isolation (ba_debug_band_solve) against numpy on
random symmetric positive-definite band systems: ragged sizes (n not a multiple of the 32-wide panel),
half-bandwidths below / at / above one panel, dense (kd = n-1), both LDL^T paths (one-cluster kernel and
the cooperative-grid kernel) and the Householder QR path. Replaces SimplicialLDLT::compute/solve and
DenseBlockedThinQR (BacktrackLevMarqQRChol.h:339-341, BAFunctor.h:101,111)."""
import os

import numpy as np
import pytest

from bundleadjustment_benchmarks_b200 import bal, solver

pytestmark = pytest.mark.gpu

CASES = [(18, 17), (33, 5), (64, 31), (100, 32), (189, 188), (351, 98), (700, 44), (1000, 548), (2313, 300)]


def band_spd(n, kd, seed):
    """Random symmetric, strictly diagonally dominant band matrix (half-bandwidth kd) and right-hand side."""
    rng = np.random.default_rng(seed)
    A = np.tril(rng.standard_normal((n, n)), -1)
    i, j = np.indices((n, n), sparse=True)
    A[(i - j) > kd] = 0.0
    A = A + A.T
    A[np.diag_indices(n)] = np.abs(A).sum(axis=1) + rng.uniform(0.5, 2.0, n)
    return A, rng.standard_normal(n)


def _solve(variant, precision, force_grid, env=None):
    prob = bal.synthetic(4, 40, seed=3)
    env = dict(env or {})
    if force_grid:
        env["BA_FORCE_GRID_LDLT"] = "1"
    os.environ.update(env)
    try:
        s = solver.GpuSolver(prob, variant, precision)
    finally:
        for k in env:
            os.environ.pop(k, None)
    return s


@pytest.mark.parametrize("path", ["cluster", "grid", "qr"])
def test_band_solve_matches_numpy(path):
    s = _solve("QRKIT" if path == "qr" else "QRCHOL", "f64", path == "grid")
    for i, (n, kd) in enumerate(CASES):
        if path == "qr" and n > 1000:
            continue
        A, g = band_spd(n, kd, 100 + i)
        y = s.debug_band_solve(A, g, kd)
        ref = np.linalg.solve(A, g)
        err = np.linalg.norm(y - ref) / np.linalg.norm(ref)
        assert err < 1e-12, (path, n, kd, err)
    s.close()


@pytest.mark.parametrize("n,kd", [(2500, 257), (3001, 100), (4100, 548), (1990, 31)])
def test_band_solve_two_sided(n, kd):
    """Two-sided (twisted) factorisation: rows eliminated from both ends by two clusters, middle block by one.
    Ragged sizes (n, n - 2*32*q not multiples of 32), compared with numpy and with the one-sided kernel."""
    A, g = band_spd(n, kd, 300 + n)
    ref = np.linalg.solve(A, g)
    s2 = _solve("QRCHOL", "f64", False, {"BA_LDLT_SPLIT": "0"})
    y2 = s2.debug_band_solve(A, g, kd)
    s2.close()
    s1 = _solve("QRCHOL", "f64", False, {"BA_LDLT_TWOSIDED": "0"})
    y1 = s1.debug_band_solve(A, g, kd)
    s1.close()
    assert np.linalg.norm(y2 - ref) / np.linalg.norm(ref) < 1e-12
    assert np.linalg.norm(y1 - ref) / np.linalg.norm(ref) < 1e-12
    assert np.linalg.norm(y2 - y1) / np.linalg.norm(ref) < 1e-12


@pytest.mark.parametrize("n,kd,segments", [(1990, 31, 3), (3001, 100, 1), (2500, 257, 4), (2433, 64, 2), (5001, 548, 3), (6000, 576, 3)])
def test_band_solve_separator_split(n, kd, segments):
    """Separator split (ba_split.cuh): S = [part 0 | separator | part 1], four elimination chains, the spike of either part
    (k_spike), the separator's Schur complement (k_sep_syrk) and the corrected backward passes, forced on (BA_LDLT_SPLIT=2)
    for 1..18 row tiles per panel, ragged part / middle-block sizes and 1..4 chain segments; against numpy and against the
    two-sided kernel. The same factor then serves a second right-hand side in solve-only mode through the QR variants'
    refinement (test_solve_with_existing_factor)."""
    A, g = band_spd(n, kd, 300 + n)
    ref = np.linalg.solve(A, g)
    s4 = _solve("QRCHOL", "f64", False, {"BA_LDLT_SPLIT": "2", "BA_LDLT_SPLIT_SEGMENTS": str(segments)})
    y4 = s4.debug_band_solve(A, g, kd)
    y4b = s4.debug_band_solve(A, g, kd)
    s4.close()
    s2 = _solve("QRCHOL", "f64", False, {"BA_LDLT_SPLIT": "0"})
    y2 = s2.debug_band_solve(A, g, kd)
    s2.close()
    nr = np.linalg.norm(ref)
    assert np.linalg.norm(y4 - ref) / nr < 1e-12 and np.linalg.norm(y4 - y2) / nr < 1e-12
    assert np.array_equal(y4, y4b)          # fixed summation orders: bit-reproducible


@pytest.mark.parametrize("n,kd", [(1990, 31), (3001, 100), (2500, 257), (4100, 548), (5000, 576)])
def test_band_solve_owner_computes_forward(n, kd):
    """BA_LDLT_V2=1: the owner-computes forward elimination (k_band_ldlt_fwd2, ba_ldlt2.cuh: tiles resident in shared memory,
    DMMA triangular solves, chain on the lightest CTA, DSMEM flags / hand-over) against numpy and against the first-generation
    kernel, for 1..18 row tiles per panel (kd = 31 .. 576), band edges cutting through tiles, ragged n."""
    A, g = band_spd(n, kd, 300 + n)
    ref = np.linalg.solve(A, g)
    s2 = _solve("QRCHOL", "f64", False, {"BA_LDLT_V2": "1", "BA_LDLT_SPLIT": "0"})
    y2 = s2.debug_band_solve(A, g, kd)
    s2.close()
    s1 = _solve("QRCHOL", "f64", False, {"BA_LDLT_SPLIT": "0"})
    y1 = s1.debug_band_solve(A, g, kd)
    s1.close()
    nr = np.linalg.norm(ref)
    assert np.linalg.norm(y2 - ref) / nr < 1e-12 and np.linalg.norm(y2 - y1) / nr < 1e-12


def test_solve_with_existing_factor():
    """QR variants: the refinement correction is solved by forward + backward substitution with the factor of S already in
    place (do_fwd = 2 of k_band_ldlt_cluster, two-sided and one-sided): QRKIT's step on a banded problem (two-sided path) and on
    a dense one must satisfy the normal equations of the trial as well as QRCHOL's does."""
    for prob in (bal.synthetic(400, 8000, window=10, seed=8), bal.synthetic(30, 2000, window=30, seed=9)):
        res = {}
        for variant in ("QRCHOL", "QRKIT"):
            s = solver.GpuSolver(prob, variant)
            e, cn2, _ = s.linearize()
            lam = 1e-9 * cn2
            s.compute(lam)
            dxn, _, et = s.solve_try()
            res[variant] = (s.dx(), et)
            s.close()
        d = np.linalg.norm(res["QRKIT"][0] - res["QRCHOL"][0]) / np.linalg.norm(res["QRCHOL"][0])
        assert d < 1e-8 and abs(res["QRKIT"][1] - res["QRCHOL"][1]) / res["QRCHOL"][1] < 1e-11, (prob.name, d)


def test_band_solve_indefinite_ldlt():
    """SimplicialLDLT semantics: un-pivoted LDL^T also factors symmetric indefinite matrices (D < 0 allowed)."""
    s = _solve("QRCHOL", "f64", False)
    n, kd = 300, 40
    A, g = band_spd(n, kd, 7)
    sgn = np.where(np.arange(n) % 3 == 0, -1.0, 1.0)
    A = A * sgn[:, None] * sgn[None, :]
    A[::5, ::5] *= 1.0
    A = A - 2.0 * np.diag(np.diag(A)) * (np.arange(n) % 7 == 0)
    ref = np.linalg.solve(A, g)
    y = s.debug_band_solve(A, g, kd)
    assert np.linalg.norm(y - ref) / np.linalg.norm(ref) < 1e-10
    s.close()


def test_band_solve_float():
    s = _solve("QRCHOL", "f32", False)
    for i, (n, kd) in enumerate(CASES[:7]):
        A, g = band_spd(n, kd, 200 + i)
        y = s.debug_band_solve(A, g, kd)
        ref = np.linalg.solve(A, g)
        assert np.linalg.norm(y - ref) / np.linalg.norm(ref) < 2e-5, (n, kd)
    s.close()


@pytest.mark.parametrize("n,kd,mode", [(800, 700, ""), (800, 700, "global"), (1134, 1133, ""), (2313, 2312, ""), (2313, 300, ""),
                                       (351, 98, "one_cta"), (1003, 548, "one_cta")])
def test_band_qr_paths(n, kd, mode):
    """Householder QR of the reduced camera block (QRKIT / MOREQR right block, BAFunctor.h:101,111): the tall kernel
    (kd + 8 > 640: streamed compact-WY updates; dense 126- and 257-camera systems), the global-memory fallback behind
    it, the look-ahead kernel on a larger ragged system (last panel narrower than 8 columns), and the single-CTA back
    substitution next to the cluster one."""
    env = {"global": "BA_QR_GLOBAL", "one_cta": "BA_QR_SOLVE_1CTA"}.get(mode)
    if env:
        os.environ[env] = "1"
    try:
        s = _solve("QRKIT", "f64", False)
        A, g = band_spd(n, kd, 500 + n)
        y = s.debug_band_solve(A, g, kd)
    finally:
        if env:
            os.environ.pop(env, None)
    ref = np.linalg.solve(A, g)
    assert np.linalg.norm(y - ref) / np.linalg.norm(ref) < 1e-12, (n, kd)
    s.close()


def test_band_qr_unsymmetric_scaling_and_float():
    """QR does not need definiteness: a symmetric indefinite band system; and the float instantiation."""
    s = _solve("QRKIT", "f64", False)
    n, kd = 420, 60
    A, g = band_spd(n, kd, 11)
    sgn = np.where(np.arange(n) % 4 == 0, -1.0, 1.0)
    A = A * sgn[:, None] * sgn[None, :] - 2.0 * np.diag(np.diag(A)) * (np.arange(n) % 5 == 0)
    y = s.debug_band_solve(A, g, kd)
    ref = np.linalg.solve(A, g)
    assert np.linalg.norm(y - ref) / np.linalg.norm(ref) < 1e-10
    s.close()
    s = _solve("QRKIT", "f32", False)
    for i, (n, kd) in enumerate(CASES[:7]):
        A, g = band_spd(n, kd, 600 + i)
        y = s.debug_band_solve(A, g, kd)
        ref = np.linalg.solve(A, g)
        assert np.linalg.norm(y - ref) / np.linalg.norm(ref) < 5e-5, (n, kd)
    s.close()
