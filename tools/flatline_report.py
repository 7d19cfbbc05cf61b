"""Teacher-forced LM runs to the flat-line exit, GPU against the CPU oracle, WITHOUT assertions: per trial
(lambda, cost disagreement, |dx| disagreement, decisions). Run on the GPU box; writes gpurun_out/flatline_trials.json.
The tolerances of tests/test_gpu_configs.py::test_teacher_forced_to_flatline come from these measurements."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bundleadjustment_benchmarks_b200 import bal, solver
from oracle.binding import Oracle

def relv(a, b): return abs(a - b) / abs(b)

def run(name, variant, max_outer):
    prob = bal.load_named(name); vid = solver.VARIANTS[variant]
    s = solver.GpuSolver(prob, variant); o = Oracle(prob)
    lam, lam_inc, hist, rows, status = None, 2.0, [0.0, 0.0], [], "cap"
    for it in range(1, max_outer + 1):
        e, cn2, cn = s.linearize(colnorms=(it == 1))
        o.set_state(*s.get_state()); eo, _, _ = o.linearize()
        if it == 1: lam = 1e-6 * cn if variant == "MOREQR" else 1e-12 * cn2
        if variant == "MOREQR": o.moreqr_outer()
        stop = False
        while True:
            s.compute(lam); dxn, rho_den, et = s.solve_try()
            ok, dxo = o.step(vid, lam); eto = o.energy_at(dxo)
            rows.append({"it": it, "lam": lam, "e": e, "et": et, "eto": eto, "cost_err": relv(et, eto) if np.isfinite(et) and np.isfinite(eto) else None,
                         "dx_err": relv(dxn, float(np.linalg.norm(dxo))) if np.isfinite(dxn) else None, "acc_gpu": bool(et < e), "acc_oracle": bool(eto < eo),
                         "gain": (e - et) / e, "info": s.numeric_status()})
            if et < e:
                rho = (e - et) / rho_den
                lam = max(lam * max(1.0 / 3.0, 1.0 - (2.0 * rho - 1.0) ** 3), 1e-10); lam_inc = 2.0; e = et; hist[it % 2] = e
                break
            s.reject()
            if lam > 1e10: stop, status = True, "lambda_max"; break
            lam *= lam_inc; lam_inc = lam_inc ** 1.5
        if stop: break
        if it > 2 and abs(e - max(hist)) < 1e-8 * e: status = "flatlined"; break
        s.accept()
    s.close()
    return {"problem": name, "variant": variant, "status": status, "outer": it, "trials": rows}

if __name__ == "__main__":
    out = []
    cases = [("problem-21-11315", v, 400) for v in ("QRCHOL", "CHOLESKY", "QRKIT", "MOREQR")] + [("problem-39-18060", v, 60) for v in ("QRCHOL", "MOREQR")]
    for name, v, mo in cases:
        t0 = time.time(); r = run(name, v, mo); out.append(r)
        errs = [t["cost_err"] for t in r["trials"] if t["cost_err"] is not None]
        print(name, v, r["status"], r["outer"], len(r["trials"]), "worst cost err %.2e" % max(errs), "flips", sum(t["acc_gpu"] != t["acc_oracle"] for t in r["trials"]), "%.0fs" % (time.time() - t0), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/flatline_trials.json", "w"))
