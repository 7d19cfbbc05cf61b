import sys; sys.path.insert(0, ".")
from bundleadjustment_benchmarks_b200 import bal, solver
p = bal.load_named("synthetic-5m")
s = solver.GpuSolver(p, "QRCHOL")
e, cn2, cn = s.linearize()
for _ in range(3):
    s.compute(1e-12 * cn2); s.solve_try(); s.reject()
s.set_profiling(True); s.compute(1e-12 * cn2); s.solve_try(); s.reject(); print("stage_ms", s.stage_ms())
c = s.debug_counters()
names = ["-", "upd CTA: loop top", "upd CTA: (chain skipped)", "upd CTA: block phase", "upd CTA: wait at cluster.sync", "backward(total)", "-", "-",
         "w0: stage + wait + barrier", "w0: column products + barrier", "w0: factor", "w1: trsm (after products barrier)", "w1: write-out", "w0: publish + wait for row warps", "w1: (other)", "w0: outside chain"]
nt = (9 * p.N + 31) // 32
for n, v in zip(names, c):
    print(f"{n:42s} {v:12d} cycles  = {v/1.9e3:9.1f} us total, {v/1.9e3/nt:6.2f} us/panel")

print("backward launch (last kernel): CTA 0 warp 0 chain %.2f us/step, its barrier wait %.2f; CTA 0 warp 1 stage %.2f us/step, its barrier wait %.2f" % tuple(c[i] / 1.965e3 / 253 for i in (8, 9, 11, 12)))
