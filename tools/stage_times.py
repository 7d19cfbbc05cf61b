"""Print device stage times of one LM iteration on synthetic-5m (QRCHOL f64), averaged over a few runs."""
import sys; sys.path.insert(0, ".")
import numpy as np
from bundleadjustment_benchmarks_b200 import bal, solver
p = bal.load_named("synthetic-5m")
s = solver.GpuSolver(p, sys.argv[1] if len(sys.argv) > 1 else "QRCHOL")
e, cn2, cn = s.linearize()
lam = 1e-12 * cn2
for _ in range(3):
    s.compute(lam); r = s.solve_try(); s.reject()
s.set_profiling(True)
acc = np.zeros(8)
for _ in range(10):
    s.compute(lam); r = s.solve_try(); s.reject(); acc += np.array(s.stage_ms())
print("stage_ms", np.round(acc / 10, 4).tolist(), "check", r)
