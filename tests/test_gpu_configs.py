"""GPU (-m gpu): parity on BASELINE.json's own configurations, through the C ABI against the CPU oracle.

* config 5 (synthetic 1800 cameras / 1M points / ~5M observations, QRCHOL): one full LM trial against the oracle
  (the oracle needs ~11 s for it), not only self-consistency.
* configs 1, 3b, 4 (the stand-ins 16-22106 QRKIT, 126-40037 QRCHOL + MOREQR, 257-65132 QRKIT + QRCHOL): one trial
  each inside a real LM step; these exercise the dense-S paths (one-cluster LDL^T with kd = 9N - 1, the tall band QR).
  For the QR right block of the 257-camera problem the oracle's own dense QR would take ~1 min, so the yardstick there
  is the oracle's reduced system solved by numpy (LAPACK LU) plus the oracle's back-substitution of the QRCHOL step.
* teacher-forced runs to the flat-line exit on the bundled files (QRChol.h:257-428): the oracle recomputes every trial
  from the GPU's (x, lambda); cost, |dx| and the accept decision are compared on every trial until the LM loop stops.
* N > 1: tools/mgpu_check.py under torch.distributed.run when at least two GPUs are visible.
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from bundleadjustment_benchmarks_b200 import bal, solver
from oracle.binding import Oracle

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def relv(a, b):
    return abs(a - b) / abs(b)


def rel(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300))


def _report(name, payload):
    """Measured parity numbers go next to the other GPU artefacts (copied to profiles/ by hand)."""
    d = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(d):
        with open(os.path.join(d, name), "w") as f:
            json.dump(payload, f, indent=1)


def test_full_size_step_against_oracle():
    """BASELINE config 5, QRCHOL f64: energy 1e-12, test energy 1e-9, |dx| 1e-8, rho denominator 1e-8 against the oracle."""
    prob = bal.load_named("synthetic-5m")
    s = solver.GpuSolver(prob, "QRCHOL")
    e, cn2, cn = s.linearize()
    o = Oracle(prob)
    eo, cn2o, cno = o.linearize()
    assert relv(e, eo) < 1e-12 and relv(cn2, cn2o) < 1e-12
    lam = 1e-12 * cn2o
    s.compute(lam)
    dxn, rho_den, et = s.solve_try()
    ok, dxo = o.step(solver.QRCHOL, lam)
    assert ok
    eto = o.energy_at(dxo)
    rho_o = float(dxo @ (lam * dxo + o.jtres()))
    dx = s.dx()
    out = {"energy": relv(e, eo), "energy_test": relv(et, eto), "dx_norm": relv(dxn, np.linalg.norm(dxo)),
           "dx_vector": rel(dx, dxo), "rho_den": relv(rho_den, rho_o), "lambda": lam, "K": prob.K}
    _report("parity_full_size.json", out)
    assert out["energy_test"] < 1e-9, out
    assert out["dx_norm"] < 1e-8, out
    assert out["dx_vector"] < 1e-6, out
    assert out["rho_den"] < 1e-8, out
    s.close()


@pytest.mark.parametrize("name,variant", [("problem-16-22106", "QRKIT"), ("problem-126-40037", "QRCHOL"), ("problem-126-40037", "MOREQR"),
                                          ("problem-257-65132", "QRCHOL"), ("problem-257-65132", "QRKIT")])
def test_baseline_standin_configs(name, variant):
    prob = bal.load_named(name)
    vid = solver.VARIANTS[variant]
    s = solver.GpuSolver(prob, variant)
    s.keep_reduced(True)
    o = Oracle(prob)
    eo, cn2o, cno = o.linearize()
    e, cn2, cn = s.linearize()
    assert relv(e, eo) < 1e-12 and relv(cn2, cn2o) < 1e-12 and relv(cn, cno) < 1e-12
    lam = 1e-6 * cno if variant == "MOREQR" else 1e-12 * cn2o
    slow_qr = variant == "QRKIT" and prob.N > 200     # oracle QR of a dense 2313^2 block: ~1 min; use its S, g instead
    if variant == "MOREQR":
        o.moreqr_outer()
    ok, dxo = o.step(solver.QRCHOL if slow_qr else vid, lam)
    assert ok
    So, go = o.reduced()
    s.compute(lam)
    dxn, rho_den, et = s.solve_try()
    S, g = s.reduced()
    dx = s.dx()
    assert rel(S, So) < 1e-11 and rel(g, go) < 1e-9
    ycam = -np.linalg.solve(So, go)                    # third opinion on the camera part: LAPACK LU on the oracle's system
    M3 = 3 * prob.M
    qr_right = variant in ("QRKIT", "MOREQR")
    out = {"dx_cam_vs_lapack": rel(dx[M3:], ycam), "dx_vs_oracle": rel(dx, dxo), "dx_norm": relv(dxn, np.linalg.norm(dxo)),
           "energy_test": relv(et, o.energy_at(dxo))}
    _report(f"parity_{name}_{variant}.json", out)
    assert out["dx_cam_vs_lapack"] < 1e-6, out
    assert out["dx_vs_oracle"] < (1e-6 if qr_right else 1e-7), out
    assert out["dx_norm"] < (1e-8 if qr_right else 1e-9), out
    assert out["energy_test"] < 1e-9, out
    s.close()


def _cond_eps(lam, c0=7e9):
    """cond(S + lambda I) * eps on the bundled files: ~ c0 / lambda until it saturates near 1e16."""
    return min(c0 / lam, 1e16) * 2.2e-16


@pytest.mark.parametrize("name,variant,max_outer", [("problem-21-11315", "QRCHOL", 400), ("problem-21-11315", "CHOLESKY", 400),
                                                    ("problem-21-11315", "QRKIT", 400), ("problem-21-11315", "MOREQR", 400),
                                                    ("problem-39-18060", "QRCHOL", 60), ("problem-39-18060", "MOREQR", 60)])
def test_teacher_forced_to_flatline(name, variant, max_outer):
    """The GPU runs the reference's LM loop to its own exit ("energy flat-lined", QRChol.h:419-425; problem-39: first 60
    outer iterations); the oracle recomputes every trial from the GPU's (x, lambda). What two correct double-precision
    solvers can agree on is set by cond(S + lambda I) ~ 7e9 / lambda (measured: profiles/r02_flatline_parity.md):
      * lambda >= 1e-5 : cost to 1e-6 (north-star "final cost" bound; measured <= 2e-8), same accept/reject decision;
      * lambda >= 1e-7 : cost within max(1e-9, 10 cond eps) (measured ~ 1 cond eps), same decision;
      * lambda <  1e-7 : cond eps > 1e-5, the reduced system is numerically singular for any solver (SURVEY.md App. E:
        LAPACK LU and QR differ by 3.6 % in test energy at 1e-10): disagreement is reported, not asserted.
    At the end the committed final cost must be reproduced by the oracle's evaluation of the same state to 1e-12."""
    prob = bal.load_named(name)
    vid = solver.VARIANTS[variant]
    s = solver.GpuSolver(prob, variant)
    o = Oracle(prob)
    lam, lam_inc = None, 2.0
    hist = [0.0, 0.0]
    worst = {"lam>=1e-5": 0.0, "lam>=1e-7": 0.0, "all": 0.0}
    trials, flips, flips_low = 0, 0, 0
    status = "cap"
    for it in range(1, max_outer + 1):
        e, cn2, cn = s.linearize(colnorms=(it == 1))
        o.set_state(*s.get_state())
        eo, _, _ = o.linearize()
        assert relv(e, eo) < 1e-12
        if it == 1:
            lam = 1e-6 * cn if variant == "MOREQR" else 1e-12 * cn2
        if variant == "MOREQR":
            o.moreqr_outer()
        stop = False
        while True:
            s.compute(lam)
            dxn, rho_den, et = s.solve_try()
            ok, dxo = o.step(vid, lam)
            eto = o.energy_at(dxo)
            trials += 1
            both_bad = (not np.isfinite(et)) and (not np.isfinite(eto))
            err = relv(et, eto) if np.isfinite(et) and np.isfinite(eto) else (0.0 if both_bad else 1.0)
            worst["all"] = max(worst["all"], err)
            flip = (et < e) != (eto < eo)
            if lam >= 1e-7:
                worst["lam>=1e-7"] = max(worst["lam>=1e-7"], err)
                assert err < max(1e-9, 10.0 * _cond_eps(lam)), (it, lam, err)
                assert not flip, (it, lam, et, eto, e)
                if lam >= 1e-5:
                    worst["lam>=1e-5"] = max(worst["lam>=1e-5"], err)
                    assert err < 1e-6, (it, lam, err)
            flips += int(flip)
            flips_low += int(flip and lam < 1e-7)
            if et < e:
                rho = (e - et) / rho_den
                lam = max(lam * max(1.0 / 3.0, 1.0 - (2.0 * rho - 1.0) ** 3), 1e-10)
                lam_inc = 2.0
                e = et
                hist[it % 2] = e
                break
            s.reject()
            if lam > 1e10:
                stop, status = True, "lambda_max"
                break
            lam *= lam_inc
            lam_inc = lam_inc ** 1.5
        if stop:
            break
        if it > 2 and abs(e - max(hist)) < 1e-8 * e:
            status = "flatlined"
            break
        s.accept()
    final_gpu = s.eval()
    o.set_state(*s.get_state())
    final_oracle, _, _ = o.linearize()
    out = {"problem": name, "variant": variant, "outer_iterations": it, "trials": trials, "status": status, "final_cost": final_gpu,
           "final_cost_rel_err": relv(final_gpu, final_oracle), "worst_trial_cost_rel_err": worst, "decision_flips": flips,
           "decision_flips_below_1e-7": flips_low, "last_lambda": lam}
    _report(f"flatline_{name}_{variant}.json", out)
    assert status in ("flatlined", "lambda_max") or max_outer < 400, out
    assert flips == flips_low, out
    assert out["final_cost_rel_err"] < 1e-12, out
    s.close()


def test_multi_gpu_equivalence():
    """2 ranks over NCCL against 1 GPU: a sharded LM trial reproduces the single-GPU energy, |dx|, rho denominator and
    test energy to rounding (tools/mgpu_check.py asserts the bounds). Skipped on a single-GPU box."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least two GPUs")
    port = 29600 + os.getpid() % 1000
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", "mgpu_check.py"), "problem-39-18060"]
    r = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-4000:]
    assert "QRCHOL x2" in r.stdout and "CHOLESKY x2" in r.stdout
