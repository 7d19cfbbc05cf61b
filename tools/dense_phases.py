import sys; sys.path.insert(0, ".")
from bundleadjustment_benchmarks_b200 import bal, solver
p = bal.load_named("synthetic-5m")
s = solver.GpuSolver(p, "QRCHOL")
e, cn2, cn = s.linearize()
for _ in range(3):
    s.compute(1e-12 * cn2); s.solve_try(); s.reject()
s.set_profiling(True); s.compute(1e-12 * cn2); s.solve_try(); s.reject(); print("stage_ms", s.stage_ms())
c = s.debug_counters()
names = ["phaseB(col)", "sync", "chain(team0 only; t128 idle)", "phaseA(rest)", "sync", "backward(total)", "blk:loop top", "blk:grab+C load issue", "blk:wait+barrier", "blk:decode+stage next", "blk:mma", "blk:store"]
nt = (9 * p.N + 31) // 32
for n, v in zip(names, c):
    print(f"{n:16s} {v:12d} cycles  = {v/1.9e3:9.1f} us total, {v/1.9e3/nt:6.2f} us/panel")
