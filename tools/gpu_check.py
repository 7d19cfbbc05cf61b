"""First-contact GPU diagnostic: compares every stage of the CUDA path with the CPU oracle and prints
the differences (no asserts). Usage: python tools/gpu_check.py [tiny|small|p21|all]"""
import sys, time, traceback
import numpy as np
sys.path.insert(0, ".")
from bundleadjustment_benchmarks_b200 import bal, solver
from oracle.binding import Oracle

def rel(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300))

def check(prob, variants=("QRCHOL", "CHOLESKY", "QRKIT", "MOREQR"), precisions=("f64",), lam=None):
    print(f"=== {prob.name}: N={prob.N} M={prob.M} K={prob.K}", flush=True)
    for prec in precisions:
        o = Oracle(prob, precision=prec)
        e0, cn2, cn = o.linearize()
        lam0 = 1e-12 * cn2 if lam is None else lam
        for var in variants:
            try:
                t0 = time.time()
                s = solver.GpuSolver(prob, variant=var, precision=prec)
                s.keep_reduced(True)
                ge, gcn2, gcn = s.linearize()
                print(f"[{prec} {var}] energy gpu {ge:.12e} cpu {e0:.12e} rel {abs(ge-e0)/e0:.2e} | cn2 rel {abs(gcn2-cn2)/cn2:.2e}")
                if var == variants[0]:
                    r = s.residuals(); ro = o.residuals()
                    Jc, Jp = s.jacobian(); Jco, Jpo = o.jacobian()
                    print(f"    residual rel {rel(r, ro):.2e} maxabs {np.abs(r-ro).max():.2e} | Jc rel {rel(Jc, Jco):.2e} Jp rel {rel(Jp, Jpo):.2e}")
                vid = solver.VARIANTS[var]
                if var == "MOREQR":
                    o.moreqr_outer()
                ok, dxo = o.step(vid, lam0)
                So, go = o.reduced()
                eto = o.energy_at(dxo)
                s.compute(lam0)
                dxn, rho_den, et = s.solve_try()
                S, g = s.reduced()
                dx = s.dx()
                jt = o.jtres()
                rho_o = float(dxo @ (lam0 * dxo + jt))
                print(f"    S rel {rel(S, So):.2e} g rel {rel(g, go):.2e} | dx rel {rel(dx, dxo):.2e} cam {rel(dx[3*prob.M:], dxo[3*prob.M:]):.2e} "
                      f"|dx| gpu {dxn:.10e} cpu {np.linalg.norm(dxo):.10e} rel {abs(dxn-np.linalg.norm(dxo))/np.linalg.norm(dxo):.2e}")
                print(f"    E_test gpu {et:.12e} cpu {eto:.12e} rel {abs(et-eto)/eto:.2e} | rho_den gpu {rho_den:.10e} cpu {rho_o:.10e} rel {abs(rho_den-rho_o)/abs(rho_o):.2e} | {time.time()-t0:.1f}s")
                s.close()
            except Exception:
                traceback.print_exc()

which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("tiny", "all"):
    check(bal.synthetic(6, 60, seed=1))
if which in ("small", "all"):
    check(bal.synthetic(40, 3000, window=8, seed=2), precisions=("f64", "f32"))
if which in ("p21", "all"):
    check(bal.load_named("problem-21-11315"))
if which in ("lm", "all"):
    prob = bal.load_named("problem-21-11315")
    for var in ("QRCHOL", "CHOLESKY", "QRKIT", "MOREQR"):
        try:
            t0 = time.time()
            s = solver.GpuSolver(prob, variant=var)
            st, log = s.minimize(max_outer=12)
            tg = time.time() - t0
            o = Oracle(prob); t0 = time.time(); sto, logo = o.minimize(solver.VARIANTS[var], 12); tc = time.time() - t0
            print(f"--- LM {var}: gpu status {st} trials {len(log)} {tg:.2f}s | cpu status {sto} trials {len(logo)} {tc:.2f}s")
            for a, b in zip(log, logo):
                print(f"  it {a.iter:3d} acc {int(a.accepted)}/{b.accepted} E {a.energy_test:.10f}/{b.energy_test:.10f} rel {abs(a.energy_test-b.energy_test)/b.energy_test:.1e} "
                      f"|dx| rel {abs(a.dx_norm-b.dx_norm)/b.dx_norm:.1e} lam {a.lambda_next:.4e}/{b.lambda_next:.4e}")
            s.close()
        except Exception:
            traceback.print_exc()
