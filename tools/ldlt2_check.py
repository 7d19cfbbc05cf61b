"""GPU check of the owner-computes forward elimination (ba_ldlt2.cuh) against the first-generation cluster kernel and
numpy: two-sided band solves on random SPD band systems, then one LM trial of BASELINE config 5 with stage times."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bundleadjustment_benchmarks_b200 import bal, solver

def band_spd(n, kd, seed):
    rng = np.random.default_rng(seed)
    A = np.tril(rng.standard_normal((n, n)), -1)
    i, j = np.indices((n, n), sparse=True)
    A[(i - j) > kd] = 0.0
    A = A + A.T
    A[np.diag_indices(n)] = np.abs(A).sum(axis=1) + rng.uniform(0.5, 2.0, n)
    return A, rng.standard_normal(n)

def mk(v2):
    os.environ["BA_LDLT_V2"] = "1" if v2 else "0"
    s = solver.GpuSolver(bal.synthetic(4, 40, seed=3), "QRCHOL", "f64")
    os.environ.pop("BA_LDLT_V2")
    return s

quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
s2, s1 = mk(True), mk(False)
for n, kd in [(1990, 31), (3001, 100), (2500, 257), (4100, 548), (5000, 576), (6000, 548), (2600, 570)]:
    A, g = band_spd(n, kd, 300 + n)
    ref = np.linalg.solve(A, g)
    y2 = s2.debug_band_solve(A, g, kd); y1 = s1.debug_band_solve(A, g, kd)
    nr = np.linalg.norm(ref)
    print(f"n={n} kd={kd}: v2 vs numpy {np.linalg.norm(y2 - ref) / nr:.2e}  v1 vs numpy {np.linalg.norm(y1 - ref) / nr:.2e}  v2 vs v1 {np.linalg.norm(y2 - y1) / nr:.2e}", flush=True)
s1.close(); s2.close()
if not quick:
    prob = bal.load_named("synthetic-5m")
    res = {}
    for v2 in (True, False):
        os.environ["BA_LDLT_V2"] = "1" if v2 else "0"
        s = solver.GpuSolver(prob, "QRCHOL")
        os.environ.pop("BA_LDLT_V2")
        e, cn2, _ = s.linearize(); lam = 1e-12 * cn2
        for _ in range(3):
            s.compute(lam); out = s.solve_try(); s.reject()
        s.set_profiling(True)
        st = np.zeros(8)
        for _ in range(5):
            s.compute(lam); out = s.solve_try(); s.reject(); st += s.stage_ms()
        st /= 5
        res[v2] = (out, s.dx(), st)
        print("v2" if v2 else "v1", "dx_norm %.15e et %.15e" % (out[0], out[2]), "stages", np.round(st, 3), "info", s.numeric_status(), flush=True)
        s.close()
    d = np.linalg.norm(res[True][1] - res[False][1]) / np.linalg.norm(res[False][1])
    print("synthetic-5m dx v2 vs v1: %.2e ; factor %.3f -> %.3f ms" % (d, res[False][2][3], res[True][2][3]))
