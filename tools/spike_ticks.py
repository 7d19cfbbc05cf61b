import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bundleadjustment_benchmarks_b200 import bal, solver
os.environ["BA_LDLT_SPLIT"] = "1"
prob = bal.load_named("synthetic-5m")
s = solver.GpuSolver(prob, "QRCHOL")
e, cn2, _ = s.linearize(); lam = 1e-12 * cn2
for _ in range(3):
    s.compute(lam); out = s.solve_try(); s.reject()
d = np.array(s.debug_counters_n(256))[128:216].reshape(11, 8)
t0 = d[:, 0].min()
for w in range(11):
    print("warp", w, " ".join("%7d" % (d[w, i] - t0) for i in range(6)))
s.close()
