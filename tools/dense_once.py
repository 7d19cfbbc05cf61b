import sys; sys.path.insert(0, ".")
from bundleadjustment_benchmarks_b200 import bal, solver
p = bal.load_named("synthetic-5m")
v = sys.argv[1] if len(sys.argv) > 1 else "QRCHOL"
s = solver.GpuSolver(p, v)
e, cn2, cn = s.linearize()
lam = 1e-6 * cn if v == "MOREQR" else 1e-12 * cn2
for _ in range(2):
    s.compute(lam); print(s.solve_try()); s.reject()
