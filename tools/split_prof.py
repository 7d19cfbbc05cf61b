"""Three split-factorisation trials on BASELINE config 5 (for an ncu launch list)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bundleadjustment_benchmarks_b200 import bal, solver
os.environ["BA_LDLT_SPLIT"] = os.environ.get("BA_LDLT_SPLIT", "1")
prob = bal.load_named("synthetic-5m")
s = solver.GpuSolver(prob, "QRCHOL")
e, cn2, _ = s.linearize(); lam = 1e-12 * cn2
for _ in range(3):
    s.compute(lam); out = s.solve_try(); s.reject()
print(out)
s.close()
