"""Phase cycle counters of k_band_ldlt_fwd2 (build with NVCCFLAGS += -DBA_L2_TICKS): CTA 0 of cluster 0."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bundleadjustment_benchmarks_b200 import bal, solver
prob = bal.load_named("synthetic-5m")
import os
os.environ["BA_LDLT_V2"] = "1"
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

ev = s.debug_counters_n(16 + 16 * 12)[16:]
print("timeline of iteration 101 (us after the CTA left the cluster barrier); chain CTA = 1 (class r - c = 3), diagonal-tile owner 10, column CTAs = 2, 6, 10, 14")
print("CTA | warp0: (chain) staged / stg->regs / published | warp1: staged+barrier / window done / column solved | warp4: staged(or W done) / window done / column solved | all: barrier passed (warp0, warp1, warp4)")
for rk in range(16):
    v = [x / 1.965e3 for x in ev[rk * 12:(rk + 1) * 12]]
    print(f"{rk:3d} | " + " ".join(f"{x:6.2f}" for x in v[0:3]) + " | " + " ".join(f"{x:6.2f}" for x in v[4:7]) + " | " + " ".join(f"{x:6.2f}" for x in v[8:11]) + f" | {v[3]:6.2f} {v[7]:6.2f} {v[11]:6.2f}")
