"""One linearisation + N LM trials of BASELINE config 5 for ncu (launch list / --set full captures). Usage: prof_step.py VARIANT [steps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bundleadjustment_benchmarks_b200 import bal, solver
variant = sys.argv[1] if len(sys.argv) > 1 else "QRCHOL"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
prob = bal.load_named("synthetic-5m")
s = solver.GpuSolver(prob, variant)
e, cn2, cn = s.linearize()
lam = 1e-6 * cn if variant == "MOREQR" else 1e-12 * cn2
for _ in range(steps):
    s.linearize(colnorms=False); s.compute(lam); out = s.solve_try(); s.reject()
print(variant, "ok", out, "launches", s.launches())
s.close()
