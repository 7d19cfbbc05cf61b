#!/usr/bin/env python
"""ms per LM iteration (Jacobian evaluation + per-point block QR + reduced camera solve) on B200.

    python bench.py --gpus N --steps K --warmup W [--workload synthetic-5m] [--variant QRCHOL]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # CPU oracle of the same path (the reference is CPU-only)

A "step" is one outer LM iteration with one lambda trial, exactly the sequence of
BacktrackLevMarqQRChol.h:257-371: residual energy at x, (lambda_0 fixed beforehand), per-point block
QR + Schur accumulation, all-reduce (N>1), factorisation of the reduced camera block, reduced solve,
back-substitution, parameter update and the test-point energy. The trial is then rejected so that
every step is the same work on the same state. Points are sharded across ranks (strong scaling of
one fixed problem); the reduced camera system is summed with ncclAllReduce inside the C ABI.

`value` is device time (CUDA events on the library's own stream, max over ranks) with all inputs
resident in HBM; `e2e` is the same step driven with HOST buffers through the C ABI (state upload +
step download inside the timed region, wall clock, max over ranks).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ms per LM iteration (J eval + block QR + camera solve)"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="synthetic-5m")
    ap.add_argument("--variant", default="QRCHOL", choices=["QRKIT", "QRCHOL", "MOREQR", "CHOLESKY"])
    ap.add_argument("--precision", default="f64", choices=["f32", "f64"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def workload_config(prob, args, nranks):
    return {"workload": f"{args.workload}: {prob.N} cameras / {prob.M} points / {prob.K} observations "
                        f"(BAL-shaped synthetic, seed 20261018)" if args.workload.startswith("synthetic")
            else f"{args.workload}: {prob.N} cameras / {prob.M} points / {prob.K} observations",
            "variant": args.variant, "precision": args.precision,
            "step": "outer LM iteration with one lambda trial (eval + schur + factor + solve + backsub + test energy), rejected",
            "parallelism": f"points sharded over {nranks} GPU(s); all-reduce of the reduced camera system",
            "l2": "inputs larger than L2 (observations + points + band matrix > 126 MB)" if prob.K > 3_000_000
            else "inputs smaller than L2; state re-uploaded only in the e2e leg"}


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.samples, self.proc, self.thread = index, [], None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append([t.strip() for t in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for s in self.samples:
            try:
                sm.append(float(s[0])); mx.append(float(s[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------- CPU oracle legs
def oracle_time_model(prob, variant, budget_s=20.0):
    """Times the CPU oracle (single thread, like the reference) on two bounded prefixes of the workload
    and extrapolates linearly in the number of observations (the reduced-system factorisation is the
    intercept: it is done at full size on every sample). Returns (ms for the full workload, description)."""
    from bundleadjustment_benchmarks_b200 import sharding
    from oracle.binding import VARIANTS, Oracle
    import dataclasses
    vid = VARIANTS[variant]
    off = prob.point_offsets()

    def prefix(npts):
        o1 = int(off[npts])
        return dataclasses.replace(prob, view=prob.view[:o1], point=prob.point[:o1], meas=prob.meas[:o1], X=prob.X[:npts], perm=None)

    def one_iteration(p):
        o = Oracle(p)
        t0 = time.perf_counter()
        e, cn2, cn = o.linearize()
        lam = 1e-6 * cn if variant == "MOREQR" else 1e-12 * cn2
        if variant == "MOREQR":
            o.moreqr_outer()
        ok, dx = o.step(vid, lam)
        o.energy_at(dx)
        return time.perf_counter() - t0

    if prob.K <= 300_000:
        t = one_iteration(prob)
        return t * 1e3, f"whole workload ({prob.K} observations), one LM iteration, 1 thread"
    n1 = min(prob.M, 8_000)
    t1 = one_iteration(prefix(n1))
    k1 = int(off[n1])
    per_obs = max(t1 / k1, 1e-7)
    n2 = int(min(prob.M, max(3 * n1, n1 + (budget_s - t1) / per_obs / (prob.K / prob.M) / 2)))
    n2 = max(n2, n1 + 1)
    t2 = one_iteration(prefix(n2))
    k2 = int(off[n2])
    slope = (t2 - t1) / (k2 - k1)
    icpt = max(t1 - slope * k1, 0.0)
    full = icpt + slope * prob.K
    return full * 1e3, (f"two prefixes of the workload ({k1} and {k2} of {prob.K} observations, all {prob.N} cameras; "
                        f"{t1:.2f} s and {t2:.2f} s), extrapolated linearly in observations; reduced-system "
                        f"factorisation at full size in both; 1 thread")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from bundleadjustment_benchmarks_b200 import bal
    prob = bal.load_named(args.workload)
    ncores = os.cpu_count() or 1
    times, desc = [], ""
    budget = max(4.0, 150.0 / max(args.steps + args.warmup, 1))
    for i in range(args.warmup + args.steps):
        ms, desc = oracle_time_model(prob, args.variant, budget_s=budget)
        if i >= args.warmup:
            times.append(ms)
    ms = float(np.mean(times)) if times else float("nan")
    line = {"impl": "reference", "metric": METRIC, "value": ms, "unit": "ms", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
            "dtype": args.precision, "data": "synthetic" if args.workload.startswith("synthetic") else "bundled BAL file",
            "config": workload_config(prob, args, 1),
            "cpu_baseline": {"value": ms, "unit": "ms", "cores": 1, "kind": "port", "sample": desc,
                             "host_cores_available": ncores,
                             "note": "the reference (C++/Eigen + private Eigen fork + SuiteSparse) cannot be built here; this is the "
                                     "oracle restatement, single-threaded like the reference"},
            "e2e": {"value": ms, "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------- ours
def run_ours(args):
    import torch
    import torch.distributed as dist
    from bundleadjustment_benchmarks_b200 import bal, sharding, solver, _lib
    import ctypes as C

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return float(x)
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    full = bal.load_named(args.workload)
    prob = sharding.shard(full, rank, world)
    s = solver.GpuSolver(prob, args.variant, args.precision, device=local_rank)
    if world > 1:
        uid = [None]
        if rank == 0:
            buf = C.create_string_buffer(128)
            rc = _lib.lib().ba_comm_unique_id(buf)
            assert rc == 0, _lib.lib().ba_last_error()
            uid[0] = buf.raw
        dist.broadcast_object_list(uid, src=0)
        s.set_bandwidth(sharding.global_bandwidth(full))
        s.comm_init(rank, world, uid[0])

    e0, cn2, cn = s.linearize(colnorms=True)
    lam = 1e-6 * cn if args.variant == "MOREQR" else 1e-12 * cn2

    def step():
        s.linearize(colnorms=False)
        s.compute(lam)
        out = s.solve_try()
        s.reject()
        return out

    for _ in range(max(args.warmup, 3)):
        step()
    # ---- device-timed region
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = s.launches()
    barrier()
    s.timer_start()
    for _ in range(args.steps):
        dxn, rho_den, et = step()
    dev_ms = s.timer_stop()
    barrier()
    launches = s.launches() - l0
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = max_over_ranks(dev_ms) / args.steps

    # ---- end-to-end region: PINNED host buffers through the C ABI every step (state upload + step download)
    def pinned(a):
        t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).pin_memory()
        return t, t.numpy()
    keep, (R, T, f, k1, k2, X) = zip(*[pinned(a) for a in (prob.R, prob.T, prob.f, prob.k1, prob.k2, prob.X)])
    dx_t = torch.empty(3 * prob.M + 9 * prob.N, dtype=torch.float64).pin_memory()
    dx_h = dx_t.numpy()
    h2d = sum(a.nbytes for a in (R, T, f, k1, k2, X))
    d2h = dx_h.nbytes + 3 * 8
    e2e_steps = max(3, min(args.steps, 10))

    def e2e_step():
        s.set_state(R, T, f, k1, k2, X)
        out = step()
        s.dx_into(dx_h)
        return out

    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / e2e_steps

    # ---- per-stage device times (CUDA events on the library stream) for the roofline of the dominant kernel
    s.set_profiling(True)
    stage = np.zeros(8)
    nprof = 5
    for _ in range(nprof):
        step()
        stage += s.stage_ms()
    stage /= nprof
    s.set_profiling(False)
    names = ["k_point_factor", "k_schur_gather", "all_reduce", "factor", "reduced_solve", "k_cam_update", "k_backsub_eval", "reductions"]
    sz = 4 if args.precision == "f32" else 8
    N, M, K = prob.N, prob.M, prob.K
    bw = s.bandwidth
    kd = min(9 * N - 1, 9 * bw + 8)
    band_bytes = 9 * N * (kd + 1) * sz
    pairs = int(sum(n * (n - 1) // 2 for n in np.bincount(prob.point, minlength=M)))  # off-diagonal pair list entries
    rec = 28 * sz  # bytes of one per-observation record, P or D (csrc/ba_tile.cuh)
    alg_bytes = {
        # observations in (indices, measurement), points + cameras in, P and D records + point records out
        "k_point_factor": K * (12 + 2 * sz) + 3 * M * sz + 16 * N * sz + 2 * K * rec + 16 * M * sz,
        # every P record once for the pair sums (re-reads are served by L2), every D record once for the
        # diagonal blocks, the pair list, S band + g + gJ out
        "k_schur_gather": 2 * K * rec + 8 * pairs + band_bytes + 2 * 9 * N * sz,
        "k_backsub_eval": K * (12 + 2 * sz) + K * rec + 16 * M * sz + 9 * M * sz + 2 * 16 * N * sz + 9 * N * sz,
        "factor": 2 * band_bytes,
        "reduced_solve": band_bytes + 3 * 9 * N * sz,
    }
    dom = max(alg_bytes, key=lambda k: stage[names.index(k)])
    dom_ms = float(stage[names.index(dom)])
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(f"{args.workload}:{args.variant}:{args.precision}", {}).get(dom)
        except Exception:
            traffic = None
    kernel_name = {"factor": "k_band_ldlt_cluster" if args.variant in ("QRCHOL", "CHOLESKY") else "k_band_qr",
                   "reduced_solve": "k_band_ldlt_cluster (solve folded in)" if args.variant in ("QRCHOL", "CHOLESKY") else "k_band_qr_backsolve",
                   "k_schur_gather": "k_schur_diag + k_schur_gather"}.get(dom, dom)
    n_red = 9 * N
    if dom == "factor":
        # the band factorisation is an FP64 contraction (n kd^2 flops for LDL^T, 4x for Householder QR of the
        # square band) on the FP64 tensor pipe (DMMA); peak = 64 FMA/clk/SM measured with tools/ubench
        # (DMMA m8n8k4 and DFMA both) x 148 SMs x max SM clock
        flops = float(n_red) * kd * kd * (1.0 if args.variant in ("QRCHOL", "CHOLESKY") else 4.0)
        fp64_peak = 64 * 2 * 148 * 1.965e9 / 1e12
        achieved = flops / (dom_ms * 1e-3) / 1e12 if dom_ms > 0 else 0.0
        roofline = {"bound": "tensor", "kernel": kernel_name, "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                    "frac": achieved / fp64_peak, "traffic": traffic,
                    "peak_source": "FP64 DMMA/DFMA issue rate measured with tools/ubench (64 FMA/clk/SM) x 148 SMs x 1.965 GHz; "
                                   "the kernel runs on ONE 16-CTA cluster (latency-bound panel chain), see DESIGN.md",
                    "algorithmic_flops": flops, "kernel_ms": dom_ms}
    else:
        achieved = alg_bytes[dom] / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
        roofline = {"bound": "hbm", "kernel": kernel_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                    "algorithmic_bytes": alg_bytes[dom], "kernel_ms": dom_ms}
    roofline["stages_ms"] = {n: float(v) for n, v in zip(names, stage)}
    roofline["hbm_frac_by_kernel"] = {k: (alg_bytes[k] / (stage[names.index(k)] * 1e-3) / 1e9 / peak) if stage[names.index(k)] > 0 else None
                                      for k in ("k_point_factor", "k_schur_gather", "k_backsub_eval")}
    roofline["hbm_peak"] = {"value": peak, "unit": "GB/s", "source": peak_src}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:  # reported at N = 1 only (the scaling runs carry null)
        ms, desc = oracle_time_model(full, args.variant, budget_s=20.0)
        cpu_baseline = {"value": ms, "unit": "ms", "cores": 1, "kind": "port", "sample": desc,
                        "host_cores_available": os.cpu_count()}
    if rank == 0:
        line = {"metric": METRIC, "value": ms_per_step, "unit": "ms", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": False, "scaling": "strong",
                "vs_baseline": None, "dtype": args.precision,
                "data": "synthetic" if args.workload.startswith("synthetic") else "bundled BAL file",
                "config": workload_config(full, args, world), "clocks": clocks,
                "e2e": {"value": e2e_ms, "unit": "ms", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                        "steps": e2e_steps},
                "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu_baseline,
                "check": {"energy": e0, "energy_test": et, "dx_norm": dxn, "lambda": lam}}
        print(json.dumps(line), flush=True)
    s.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
