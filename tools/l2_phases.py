"""Phase cycle counters of k_band_ldlt_fwd2 (build with NVCCFLAGS += -DBA_L2_TICKS): CTA 0 of cluster 0."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bundleadjustment_benchmarks_b200 import bal, solver
prob = bal.load_named("synthetic-5m")
s = solver.GpuSolver(prob, "QRCHOL")
e, cn2, _ = s.linearize(); lam = 1e-12 * cn2
for _ in range(2):
    s.compute(lam); s.solve_try(); s.reject()
s.set_profiling(True)
s.compute(lam); s.solve_try(); s.reject()
print("stages", np.round(s.stage_ms(), 3))
c = s.debug_counters()
names = ["F:stage-wait", "F:mma", "F:stg->regs", "F:factor+publish", "-", "T phase", "U staging", "U compute", "wait S1", "wait S2", "-", "w0 wait S2", "w0 rest (F + wait S1)", "-", "-", "-"]
npan = 253.0
ghz = 1.965
for n, v in zip(names, c):
    if n != "-": print(f"{n:24s} {v / npan / ghz / 1e3:8.3f} us/panel (CTA 0 is diagonal owner / column owner every 4th panel)")
