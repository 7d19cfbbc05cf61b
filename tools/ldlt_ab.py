"""A/B of the first-generation band LDL^T kernel's chain options on BASELINE config 5 (factor stage, ms)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bundleadjustment_benchmarks_b200 import bal, solver
prob = bal.load_named("synthetic-5m")
ref = None
for env in ({}, {"BA_LDLT_W_AFTER": "1"}, {"BA_LDLT_ROWS_AFTER": "1"}, {"BA_LDLT_W_AFTER": "1", "BA_LDLT_ROWS_AFTER": "1"}, {"BA_LDLT_ROWW": "3"}, {"BA_LDLT_ROWW": "3", "BA_LDLT_W_AFTER": "1"}):
    os.environ.update(env)
    s = solver.GpuSolver(prob, "QRCHOL")
    for k in env: os.environ.pop(k)
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
    print(env, "factor %.3f ms" % st[3], "dx vs default %.1e" % (np.linalg.norm(dx - ref) / np.linalg.norm(ref)), flush=True)
    s.close()
