"""Separator split of the band LDL^T (BA_LDLT_SPLIT=1, ba_split.cuh): random SPD band systems against numpy and against
the two-sided kernel, then the factor-stage time on BASELINE config 5 with and without the split."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bundleadjustment_benchmarks_b200 import bal, solver
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_gpu_band import band_spd

small = bal.synthetic(4, 40, seed=3)
def make(split):
    os.environ["BA_LDLT_SPLIT"] = "2" if split else "0"
    s = solver.GpuSolver(small, "QRCHOL")
    os.environ.pop("BA_LDLT_SPLIT", None)
    return s
cases = [(1990, 31), (3001, 100), (2500, 257), (5001, 548), (6000, 576)]
if len(sys.argv) > 1 and sys.argv[1] == "quick": cases = cases[:2]
for n, kd in cases:
    A, g = band_spd(n, kd, 300 + n)
    ref = np.linalg.solve(A, g)
    s = make(True); y = s.debug_band_solve(A, g, kd); s.close()
    s = make(False); y1 = s.debug_band_solve(A, g, kd); s.close()
    nr = np.linalg.norm(ref)
    print(n, kd, "split vs numpy %.2e  two-sided vs numpy %.2e" % (np.linalg.norm(y - ref) / nr, np.linalg.norm(y1 - ref) / nr), flush=True)
    if np.linalg.norm(y - ref) / nr > 1e-10:
        bad = np.abs(y - ref) > 1e-8 * np.abs(ref).max()
        idx = np.nonzero(bad)[0]
        print("  bad rows: %d, first %d last %d" % (len(idx), idx[0], idx[-1]))
if len(sys.argv) > 1 and sys.argv[1] == "noperf": sys.exit(0)
prob = bal.load_named("synthetic-5m")
ref = None
for split in (False, True):
    os.environ["BA_LDLT_SPLIT"] = "1" if split else "0"
    s = solver.GpuSolver(prob, "QRCHOL")
    os.environ.pop("BA_LDLT_SPLIT", None)
    e, cn2, _ = s.linearize(); lam = 1e-12 * cn2
    for _ in range(3):
        s.compute(lam); out = s.solve_try(); s.reject()
    s.set_profiling(True)
    st = np.zeros(8)
    for _ in range(5):
        s.compute(lam); out = s.solve_try(); s.reject(); st += s.stage_ms()
    st /= 5
    dx = s.dx()
    if ref is None: ref = dx
    print("split" if split else "two-sided", "factor %.3f ms" % st[3], "stages", np.round(st, 3), "dx vs two-sided %.1e" % (np.linalg.norm(dx - ref) / np.linalg.norm(ref)), flush=True)
    s.close()
