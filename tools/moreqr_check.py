"""MOREQR two-stage scheme on the GPU (More.h:288-348): equivalence with re-factoring the damped blocks per trial,
parity with the oracle's own two-stage step, and what a trial costs in the point stage (BASELINE config 5 unless 'quick')."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bundleadjustment_benchmarks_b200 import bal, solver
from oracle.binding import Oracle

def mk(prob, two):
    os.environ["BA_MOREQR_TWOSTAGE"] = "1" if two else "0"
    s = solver.GpuSolver(prob, "MOREQR")
    os.environ.pop("BA_MOREQR_TWOSTAGE")
    return s

for name in ("problem-21-11315", "problem-39-18060"):
    prob = bal.load_named(name)
    o = Oracle(prob); eo, cn2o, cno = o.linearize(); o.moreqr_outer()
    for lam in (1e-6 * cno, 1e-4, 1e-7):
        ok, dxo = o.step(2, lam); eto = o.energy_at(dxo)
        out = {}
        for two in (True, False):
            s = mk(prob, two); s.linearize(); s.compute(lam); dxn, rd, et = s.solve_try(); out[two] = (dxn, et, s.dx()); s.close()
        r = lambda a, b: abs(a - b) / abs(b)
        print(f"{name} lam={lam:.3e}: two-stage vs oracle: cost {r(out[True][1], eto):.1e} |dx| {r(out[True][0], np.linalg.norm(dxo)):.1e} dx {np.linalg.norm(out[True][2]-dxo)/np.linalg.norm(dxo):.1e}"
              f" | per-trial path vs oracle: cost {r(out[False][1], eto):.1e} dx {np.linalg.norm(out[False][2]-dxo)/np.linalg.norm(dxo):.1e}"
              f" | two-stage vs per-trial dx {np.linalg.norm(out[True][2]-out[False][2])/np.linalg.norm(dxo):.1e}", flush=True)
if len(sys.argv) > 1 and sys.argv[1] == "quick":
    sys.exit(0)
prob = bal.load_named("synthetic-5m")
for two in (True, False):
    s = mk(prob, two)
    e, cn2, cn = s.linearize(); lam = 1e-6 * cn
    s.set_profiling(True)
    rows = []
    for t in range(3):
        s.compute(lam * 2 ** t); s.solve_try(); s.reject(); rows.append(s.stage_ms().copy())
    print("two-stage" if two else "per-trial", "point-factor stage per trial (ms):", [round(r[0], 3) for r in rows], "whole trial:", [round(r.sum(), 2) for r in rows], flush=True)
    s.close()
