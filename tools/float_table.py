"""Scalar = float build (src/BATypeUtils.h:6): one LM trial, GPU f32 path against the oracle run in float (apples to
apples) and against the double oracle, for the four variants at lambda_0, 1e3 lambda_0 and 1e6 lambda_0.
Writes gpurun_out/float_table.json / .md (committed copy: profiles/r02_float_table.md)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bundleadjustment_benchmarks_b200 import bal, solver
from oracle.binding import Oracle

def relv(a, b): return abs(a - b) / abs(b) if np.isfinite(a) and np.isfinite(b) and b != 0 else float("nan")
rows = []
probs = [("small (40 cameras, 3000 points)", bal.synthetic(40, 3000, window=8, seed=2)), ("problem-21-11315", bal.load_named("problem-21-11315")),
         ("problem-39-18060", bal.load_named("problem-39-18060"))]
for pname, prob in probs:
    for variant in ("QRKIT", "QRCHOL", "MOREQR", "CHOLESKY"):
        vid = solver.VARIANTS[variant]
        o64 = Oracle(prob); e64, cn2, cn = o64.linearize()
        o32 = Oracle(prob, precision="f32"); e32, _, _ = o32.linearize()
        if variant == "MOREQR": o64.moreqr_outer(); o32.moreqr_outer()
        lam0 = 1e-6 * cn if variant == "MOREQR" else 1e-12 * cn2
        g = solver.GpuSolver(prob, variant, "f32"); ge, _, _ = g.linearize()
        for mult in (1.0, 1e3, 1e6):
            lam = lam0 * mult
            ok64, dx64 = o64.step(vid, lam); et64 = o64.energy_at(dx64)
            ok32, dx32 = o32.step(vid, lam); et32 = o32.energy_at(dx32) if ok32 else float("nan")
            g.compute(lam); dxn, _, et = g.solve_try(); g.reject()
            rows.append({"problem": pname, "variant": variant, "lambda": lam, "mult": mult, "energy_rel_err_gpu32_vs_o64": relv(ge, e64),
                         "gain64": (e64 - et64) / e64, "gain_o32": (e64 - et32) / e64 if np.isfinite(et32) else float("nan"), "gain_gpu32": (e64 - et) / e64 if np.isfinite(et) else float("nan"),
                         "cost_gpu32_vs_o64": relv(et, et64), "cost_o32_vs_o64": relv(et32, et64), "cost_gpu32_vs_o32": relv(et, et32),
                         "dxn_gpu32_vs_o64": relv(dxn, float(np.linalg.norm(dx64))), "dxn_o32_vs_o64": relv(float(np.linalg.norm(dx32)), float(np.linalg.norm(dx64))),
                         "info": g.numeric_status()})
            print(rows[-1], flush=True)
        g.close()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/float_table.json", "w"), indent=1)
f = lambda v: "nan" if not np.isfinite(v) else f"{v:.1e}"
with open("gpurun_out/float_table.md", "w") as out:
    out.write("| problem | variant | lambda | cost: GPU f32 vs oracle f64 | cost: oracle f32 vs oracle f64 | cost: GPU f32 vs oracle f32 | dx norm: GPU f32 vs f64 | dx norm: oracle f32 vs f64 | relative gain f64 / oracle f32 / GPU f32 |\n|---|---|---|---|---|---|---|---|---|\n")
    for r in rows:
        out.write(f"| {r['problem']} | {r['variant']} | {r['lambda']:.2e} | {f(r['cost_gpu32_vs_o64'])} | {f(r['cost_o32_vs_o64'])} | {f(r['cost_gpu32_vs_o32'])} | {f(r['dxn_gpu32_vs_o64'])} | {f(r['dxn_o32_vs_o64'])} | {f(r['gain64'])} / {f(r['gain_o32'])} / {f(r['gain_gpu32'])} |\n")
