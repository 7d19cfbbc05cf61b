"""Small end-to-end exercise of every kernel path (for compute-sanitizer): all variants and both precisions on a bundled
problem, the long-track paths, band solves through the LDL^T cluster kernel (one- and two-sided), the register and the
tall band QR and both back substitutions."""
import os, sys; sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
from bundleadjustment_benchmarks_b200 import bal, solver
from test_gpu_band import band_spd
from test_gpu_parity import _long_track_problem

p = bal.load_named("problem-21-11315")
for prec in ("f64", "f32"):
    for v in ("QRKIT", "QRCHOL", "MOREQR", "CHOLESKY"):
        s = solver.GpuSolver(p, v, prec)
        e, cn2, cn = s.linearize()
        lam = 1e-6 * cn if v == "MOREQR" else 1e-12 * cn2
        s.compute(lam); r = s.solve_try(); s.accept()
        print(prec, v, e, r, flush=True)
        s.close()
lt = _long_track_problem()
for v in ("QRCHOL", "CHOLESKY"):
    s = solver.GpuSolver(lt, v)
    e, cn2, cn = s.linearize(); s.compute(1e-12 * cn2); print("long tracks", v, s.solve_try(), flush=True); s.close()
small = bal.synthetic(4, 40, seed=3)
for v, cases in (("QRCHOL", [(100, 32), (700, 44), (2500, 257)]), ("QRKIT", [(100, 32), (351, 98), (700, 660), (1300, 1299)])):
    s = solver.GpuSolver(small, v)
    for n, kd in cases:
        A, g = band_spd(n, kd, n)
        y = s.debug_band_solve(A, g, kd)
        print(v, n, kd, np.linalg.norm(y - np.linalg.solve(A, g)) / np.linalg.norm(y), flush=True)
    s.close()
os.environ["BA_QR_SOLVE_1CTA"] = "1"
s = solver.GpuSolver(small, "QRKIT"); A, g = band_spd(351, 98, 5); y = s.debug_band_solve(A, g, 98)
print("1cta", np.linalg.norm(y - np.linalg.solve(A, g)) / np.linalg.norm(y)); s.close()
