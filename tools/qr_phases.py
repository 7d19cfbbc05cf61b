"""Phase split of the band QR (needs a library built with -DBA_QR_TICKS): cycles of thread 0 of the panel CTA and of one update CTA."""
import sys; sys.path.insert(0, ".")
from bundleadjustment_benchmarks_b200 import bal, solver
p = bal.load_named("synthetic-5m")
s = solver.GpuSolver(p, "QRKIT")
e, cn2, cn = s.linearize()
for _ in range(2):
    s.compute(1e-12 * cn2); s.solve_try(); s.reject()
c = s.debug_counters()
names = ["loop top", "reflectors -> smem + barrier", "apply panel to next-panel columns", "factor next panel", "trailing update", "fence + grid barrier", "-", "-"]
fnames = ["panel rows -> smem + barrier", "column pass 1 (dots)", "warp reduce + block barrier", "totals + beta/tau", "t_q + T column", "column pass 2 (update)", "final barrier + write-back", "-"]
npanel = (9 * p.N + 7) // 8
for n, v in zip(names, c[:8]):
    print(f"panel CTA  {n:36s} {v:12d} cycles = {v/1.965e3/npanel:7.2f} us/panel")
for n, v in zip(fnames, c[8:]):
    print(f"  factor:  {n:36s} {v:12d} cycles = {v/1.965e3/npanel:7.2f} us/panel")
