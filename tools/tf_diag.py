import sys
sys.path.insert(0, ".")
import numpy as np
from bundleadjustment_benchmarks_b200 import bal, solver
from oracle.binding import Oracle
p21 = bal.load_named("problem-21-11315")
relv = lambda a, b: abs(a - b) / abs(b)
for variant, prec in (("QRCHOL", "f64"), ("CHOLESKY", "f64")):
    vid = solver.VARIANTS[variant]
    s = solver.GpuSolver(p21, variant, prec); o = Oracle(p21)
    lam, lam_inc = None, 2.0
    for it in range(1, 13):
        e, cn2, cn = s.linearize(colnorms=(it == 1))
        o.set_state(*s.get_state()); eo, _, _ = o.linearize()
        if it == 1: lam = 1e-12 * cn2
        while True:
            s.compute(lam); dxn, rho_den, et = s.solve_try()
            ok, dxo = o.step(vid, lam); eto = o.energy_at(dxo)
            print(f"{variant} it {it:2d} lam {lam:.3e} cost rel {relv(et, eto):.2e} |dx| rel {relv(dxn, np.linalg.norm(dxo)):.2e} dxvec rel {np.linalg.norm(s.dx()-dxo)/np.linalg.norm(dxo):.2e} acc {et<e}/{eto<eo}")
            if et < e:
                rho = (e - et) / rho_den
                lam = max(lam * max(1.0 / 3.0, 1.0 - (2.0 * rho - 1.0) ** 3), 1e-10); lam_inc = 2.0; s.accept(); break
            s.reject(); lam *= lam_inc; lam_inc = lam_inc ** 1.5
