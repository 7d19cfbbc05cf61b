"""What the right block of QRKIT / MOREQR buys numerically (BAFunctor.h:99-102,109-112: dense QR of the tall J2bot).
Camera step dx_cam of one LM trial on problem-21 at lambda in {lambda_0, 1e-4, 1e-7} from
  (a) GPU QRCHOL (LDL^T of S), (b) GPU QRKIT = LDL^T of S + corrected semi-normal refinement through J2bot (1 and 2 steps),
  (c) GPU QRKIT with the round-1 Householder QR of the square S (BA_QR_HOUSEHOLDER=1), (d) the oracle's QR of S,
  (e) the oracle's reference-faithful QR of the tall J2bot (Oracle.set_tall),
against an EXTENDED-PRECISION yardstick: the tall J2bot assembled and QR-solved in np.longdouble (eps 1e-19) from the
oracle's double Jacobian. Writes gpurun_out/qr_accuracy.md."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bundleadjustment_benchmarks_b200 import bal, solver
from oracle.binding import Oracle
from oracle.ba_oracle_np import householder_lstsq

LD = np.longdouble

def truth_camera_step(prob, Jc, Jp, res, lam):
    """y = argmin |J2bot y - d| in longdouble; returns dx_cam = -y."""
    N, K = prob.N, prob.K
    off = prob.point_offsets()
    sl = np.sqrt(LD(lam))
    rows = []
    rhs = []
    Jc = Jc.astype(LD); Jp = Jp.astype(LD); res = res.astype(LD).reshape(K, 2)
    tall = np.zeros((2 * K + 9 * N, 9 * N), dtype=LD)
    d = np.zeros(2 * K + 9 * N, dtype=LD)
    r0 = 0
    for j in range(prob.M):
        o0, o1 = int(off[j]), int(off[j + 1]); n = o1 - o0
        A = np.zeros((2 * n + 3, 3), dtype=LD)
        A[:2 * n] = Jp[o0:o1].reshape(2 * n, 3)
        A[2 * n:] = sl * np.eye(3, dtype=LD)
        B = np.zeros((2 * n + 3, 9 * n + 1), dtype=LD)
        for i in range(n):
            B[2 * i:2 * i + 2, 9 * i:9 * i + 9] = Jc[o0 + i]
            B[2 * i:2 * i + 2, 9 * n] = res[o0 + i]
        for k in range(3):                       # Householder, no pivoting needed for the yardstick (lambda > 0)
            x = A[k:, k]
            nrm = np.sqrt((x * x).sum())
            alpha = -nrm if x[0] >= 0 else nrm
            v = x.copy(); v[0] -= alpha
            vv = (v * v).sum()
            if vv == 0: continue
            A[k:, k:] -= np.outer(v, (2 / vv) * (v @ A[k:, k:]))
            B[k:] -= np.outer(v, (2 / vv) * (v @ B[k:]))
        m = 2 * n                                # rows 3.. are this point's slice of J2bot and of d
        cams = prob.view[o0:o1]
        for i in range(n):
            tall[r0:r0 + m, 9 * cams[i]:9 * cams[i] + 9] = B[3:, 9 * i:9 * i + 9]
        d[r0:r0 + m] = B[3:, 9 * n]
        r0 += m
    tall[r0:r0 + 9 * N] = sl * np.eye(9 * N, dtype=LD)
    r0 += 9 * N
    y = householder_lstsq(tall[:r0], d[:r0])
    return -(y.astype(np.float64))

def gpu_step(prob, variant, lam, env=None):
    for k, v in (env or {}).items(): os.environ[k] = v
    try:
        s = solver.GpuSolver(prob, variant)
    finally:
        for k in (env or {}): os.environ.pop(k)
    s.linearize(); s.compute(lam); s.solve_try()
    dx = s.dx(); s.close()
    return dx[3 * prob.M:]

prob = bal.load_named("problem-21-11315")
o = Oracle(prob); e, cn2, cn = o.linearize()
Jc, Jp = o.jacobian(); res = o.residuals()
M3 = 3 * prob.M
rel = lambda a, b: float(np.linalg.norm(a - b) / np.linalg.norm(b))
lines = ["| lambda | GPU QRCHOL (LDL^T of S) | GPU QRKIT, CSNE x1 | GPU QRKIT, CSNE x2 | GPU QRKIT, Householder QR of S (r1) | oracle QR of S | oracle QR of tall J2bot |", "|---|---|---|---|---|---|---|"]
for lam in (1e-12 * cn2, 1e-4, 1e-7):
    t0 = time.time(); ref = truth_camera_step(prob, Jc, Jp, res, lam); tt = time.time() - t0
    ok, dqs = o.step(0, lam)
    o.set_tall(True); ok, dqt = o.step(0, lam); o.set_tall(False)
    row = [rel(gpu_step(prob, "QRCHOL", lam), ref), rel(gpu_step(prob, "QRKIT", lam), ref), rel(gpu_step(prob, "QRKIT", lam, {"BA_QR_REFINE": "2"}), ref),
           rel(gpu_step(prob, "QRKIT", lam, {"BA_QR_HOUSEHOLDER": "1"}), ref), rel(dqs[M3:], ref), rel(dqt[M3:], ref)]
    lines.append(f"| {lam:.3e} | " + " | ".join(f"{v:.1e}" for v in row) + " |")
    print(lines[-1], f"(yardstick {tt:.0f} s)", flush=True)
os.makedirs("gpurun_out", exist_ok=True)
open("gpurun_out/qr_accuracy.md", "w").write("\n".join(lines) + "\n")
