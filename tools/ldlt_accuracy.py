"""Accuracy of the reduced-system solvers on a real (ill-conditioned) reduced camera matrix: problem-21 at small lambda."""
import os, sys
sys.path.insert(0, ".")
import numpy as np
from bundleadjustment_benchmarks_b200 import bal, solver
p = bal.load_named("problem-21-11315")
s = solver.GpuSolver(p, "QRCHOL"); s.keep_reduced(True)
e, cn2, cn = s.linearize()
os.environ["BA_FORCE_GRID_LDLT"] = "1"; sg = solver.GpuSolver(p, "QRCHOL"); os.environ.pop("BA_FORCE_GRID_LDLT")
for lam in (1e-12 * cn2, 1e-4, 1e-7, 1e-9):
    s.compute(lam); s.solve_try(); s.reject()
    S, g = s.reduced()
    n = S.shape[0]
    ref = np.linalg.solve(S.astype(np.longdouble).astype(np.float64), g)  # LAPACK LU with pivoting
    # extended precision refinement of the reference
    Sl, gl = S.astype(np.longdouble), g.astype(np.longdouble)
    x = ref.astype(np.longdouble)
    for _ in range(5):
        r = gl - Sl @ x
        x = x + np.linalg.solve(S, r.astype(np.float64)).astype(np.longdouble)
    ref = x.astype(np.float64)
    yc = s.debug_band_solve(S, g, n - 1)
    yg = sg.debug_band_solve(S, g, n - 1)
    import scipy.linalg as sl
    L, D, perm = sl.ldl(S, lower=True)
    print(f"lam {lam:.3e} cond {np.linalg.cond(S):.2e}  cluster err {np.linalg.norm(yc-ref)/np.linalg.norm(ref):.2e}  grid err {np.linalg.norm(yg-ref)/np.linalg.norm(ref):.2e}  numpy-solve err {np.linalg.norm(np.linalg.solve(S,g)-ref)/np.linalg.norm(ref):.2e}")
