"""Quick GPU check + timing of the reduced-camera-block solver on synthetic band systems."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
from bundleadjustment_benchmarks_b200 import bal, solver
sys.path.insert(0, "tests")
from test_gpu_band import band_spd, CASES

prob = bal.synthetic(4, 40, seed=3)
s = solver.GpuSolver(prob, "QRCHOL")
for i, (n, kd) in enumerate(CASES):
    A, g = band_spd(n, kd, 100 + i)
    y = s.debug_band_solve(A, g, kd)
    ref = np.linalg.solve(A, g)
    print(f"n={n:5d} kd={kd:4d} rel err {np.linalg.norm(y - ref) / np.linalg.norm(ref):.2e}", flush=True)
