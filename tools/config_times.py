"""Device stage times of one LM iteration on the smaller BASELINE configs (all variants)."""
import sys; sys.path.insert(0, ".")
import numpy as np
from bundleadjustment_benchmarks_b200 import bal, solver
cfgs = [("problem-16-22106", ("QRKIT",)), ("problem-21-11315", ("CHOLESKY",)), ("problem-39-18060", ("QRCHOL", "MOREQR")),
        ("problem-126-40037", ("QRCHOL", "MOREQR")), ("problem-257-65132", ("QRKIT", "QRCHOL"))]
if len(sys.argv) > 1 and sys.argv[1] == "all":
    cfgs.append(("synthetic-5m", ("QRCHOL", "QRKIT", "MOREQR", "CHOLESKY")))
for name, variants in cfgs:
    p = bal.load_named(name)
    for v in variants:
        s = solver.GpuSolver(p, v)
        e, cn2, cn = s.linearize()
        lam = 1e-6 * cn if v == "MOREQR" else 1e-12 * cn2
        for _ in range(2):
            s.compute(lam); s.solve_try(); s.reject()
        s.set_profiling(True)
        acc = np.zeros(8)
        for _ in range(5):
            s.compute(lam); r = s.solve_try(); s.reject(); acc += np.array(s.stage_ms())
        st = acc / 5
        print(f"{name:20s} {v:9s} total {st.sum():8.3f} ms  point {st[0]:.3f} S {st[1]:.3f} factor {st[3]:.3f} solve {st[4]:.3f} backsub {st[6]:.3f}", flush=True)
        s.close()
