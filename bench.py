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
step download inside the timed region, wall clock, max over ranks; one ba_step_streamed call per step, which pipelines the
copies against the point stage and returns bit-identical results). Also in the line: `roofline` (dominant stage against the
measured FP64 / HBM peaks, ncu DRAM traffic, point stage against SURVEY 8(d)'s algorithmic bytes), `variants` (QRKIT next to
QRCHOL), `parity_probe` (a sharded trial against the CPU oracle at this N, outside the timed regions), `clocks` (nvidia-smi samples
inside the timed regions), `cpu_baseline` (N = 1: the oracle on the whole workload, all cores and single-threaded).
`--impl reference`: the oracle on all host cores, whole workload, every step one complete LM iteration (rank 0 only).
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
    ap.add_argument("--no-parity-probe", action="store_true")
    ap.add_argument("--no-other-variant", action="store_true")
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
                                          "-lms", "20", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.perf_counter(), [t.strip() for t in line.split(",")]))

    def wait_first(self, timeout=3.0):
        t0 = time.perf_counter()
        while not self.samples and time.perf_counter() - t0 < timeout:
            time.sleep(0.02)

    def stop(self, t_begin=None, t_end=None):
        """Median SM clock and throttle reasons of the samples taken inside [t_begin, t_end] (the timed regions)."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        inside = [v for (t, v) in self.samples if (t_begin is None or t >= t_begin) and (t_end is None or t <= t_end + 0.05)]
        for s in (inside or [v for (_, v) in self.samples]):
            try:
                sm.append(float(s[0])); mx.append(float(s[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "samples_inside_timed_regions": len(inside)}


# --------------------------------------------------------------------------------- CPU oracle legs
def oracle_iteration_timer(prob, variant, threads):
    """Returns a callable that runs ONE LM iteration (linearise at x, lambda_0 trial, test energy) of the CPU oracle on
    `prob` and returns its wall time in seconds. threads = 1 is the reference's own execution model (single thread);
    threads > 1 runs the per-observation / per-point loops and the reduced solve with OpenMP."""
    from oracle.binding import VARIANTS, Oracle
    vid = VARIANTS[variant]
    o = Oracle(prob)
    o.set_threads(threads)

    def one_iteration():
        t0 = time.perf_counter()
        e, cn2, cn = o.linearize()
        lam = 1e-6 * cn if variant == "MOREQR" else 1e-12 * cn2
        if variant == "MOREQR":
            o.moreqr_outer()
        ok, dx = o.step(vid, lam)
        o.energy_at(dx)
        return time.perf_counter() - t0

    return one_iteration, o


def prefix_problem(prob, nobs_target):
    """First points of the workload (all cameras) with about nobs_target observations: the bounded sample used only
    when a whole-workload CPU iteration would not fit the run."""
    import dataclasses
    off = prob.point_offsets()
    npts = int(np.searchsorted(off, nobs_target))
    npts = max(1, min(prob.M, npts))
    o1 = int(off[npts])
    return dataclasses.replace(prob, view=prob.view[:o1], point=prob.point[:o1], meas=prob.meas[:o1], X=prob.X[:npts], perm=None)


def cpu_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def run_reference(args):
    """The reference is CPU-only and cannot be built here: this arm times the oracle restatement of the same
    path on the host cores, with all the threads it can use, on the WHOLE workload, `steps` times after `warmup`
    (every step is one complete LM iteration: nothing is extrapolated). Only if one iteration is so slow that the run
    would exceed ~5 minutes is a prefix of the workload timed instead and scaled by the observation count (said in `sample`)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from bundleadjustment_benchmarks_b200 import bal
    prob = bal.load_named(args.workload)
    ncores = cpu_threads()
    it_full, o = oracle_iteration_timer(prob, args.variant, ncores)
    if not o.has_openmp():
        ncores = 1
    t_first = it_full()                                   # untimed: pages the problem in, tells how long a step is
    total = max(args.steps + args.warmup, 1)
    scale, sample = 1.0, f"whole workload ({prob.K} observations), every step one complete LM iteration, {ncores} thread(s) (OpenMP over observations / points + reduced solve)"
    timer = it_full
    if t_first * total > 300.0:
        frac = max(0.02, 300.0 / (t_first * total))
        sub = prefix_problem(prob, int(prob.K * frac))
        timer, _o2 = oracle_iteration_timer(sub, args.variant, ncores)
        scale = prob.K / sub.K
        sample = (f"prefix of the workload ({sub.K} of {prob.K} observations, all {prob.N} cameras; the reduced solve at full size), "
                  f"scaled by {scale:.2f}; a whole-workload iteration took {t_first:.1f} s; {ncores} thread(s)")
    times = []
    for i in range(args.warmup + args.steps):
        t = timer()
        if i >= args.warmup:
            times.append(t * scale)
    ms = float(np.mean(times)) * 1e3 if times else float("nan")
    line = {"impl": "reference", "metric": METRIC, "value": ms, "unit": "ms", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
            "dtype": args.precision, "data": "synthetic" if args.workload.startswith("synthetic") else "bundled BAL file",
            "config": workload_config(prob, args, args.gpus),
            "cpu_baseline": {"value": ms, "unit": "ms", "cores": ncores, "kind": "port", "sample": sample,
                             "host_cores_available": os.cpu_count(), "ms_median": float(np.median(times)) * 1e3 if times else None,
                             "note": "the reference (C++/Eigen + private Eigen fork + SuiteSparse) cannot be built here; this is the "
                                     "oracle restatement of the same LM iteration; the reference itself is single-threaded"},
            "e2e": {"value": ms, "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def cpu_baseline_legs(prob, variant):
    """cpu_baseline of our arm (rank 0, N = 1): whole workload, all cores (median of 3 after one warm-up) and one
    single-threaded iteration (the reference's execution model), ~20-30 s of CPU work at BASELINE config 5."""
    ncores = cpu_threads()
    it_mt, o = oracle_iteration_timer(prob, variant, ncores)
    if not o.has_openmp():
        ncores = 1
    it_mt()
    mt = sorted(it_mt() for _ in range(3))[1]
    del it_mt, o
    it_st, _o = oracle_iteration_timer(prob, variant, 1)
    st = it_st()
    return {"value": mt * 1e3, "unit": "ms", "cores": ncores, "kind": "port",
            "sample": f"whole workload ({prob.K} observations), one complete LM iteration; all-core leg: median of 3 after one warm-up; "
                      f"single-thread leg: one iteration",
            "single_thread": {"value": st * 1e3, "unit": "ms", "cores": 1},
            "host_cores_available": os.cpu_count()}


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
        # energy + compute(lam) + solve_try on the device-resident state in one call (one host synchronisation per step);
        # bit-identical to ba_linearize + ba_compute + ba_solve_try (tests/test_gpu_parity.py)
        out = s.step_resident(lam)
        s.reject()
        return out[1:]

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        sampler.wait_first()
    for _ in range(max(args.warmup, 3)):
        step()
    # ---- device-timed region
    l0 = s.launches()
    barrier()
    t_region0 = time.perf_counter()
    s.timer_start()
    for _ in range(args.steps):
        dxn, rho_den, et = step()
    dev_ms = s.timer_stop()
    barrier()
    launches = s.launches() - l0
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
        # one call: state upload + energy + compute + solve_try + step download, copies pipelined against the point stage
        # (ba_step_streamed; bit-identical to set_state / linearize / compute / solve_try / get_dx called one by one)
        out = s.step_streamed(R, T, f, k1, k2, X, lam, dx_h)
        s.reject()
        return out[1:]

    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / e2e_steps
    clocks = sampler.stop(t_region0, time.perf_counter()) if rank == 0 else None

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
    kernel_name = {"factor": "k_band_ldlt_cluster (band LDL^T stage: separator split = 7 launches of it + k_spike / k_sep_syrk beside them)",
                   "reduced_solve": "k_band_ldlt_cluster (solve folded in)" if args.variant in ("QRCHOL", "CHOLESKY") else "k_csne_point + k_csne_cam + k_band_ldlt_cluster (refinement solve)",
                   "k_schur_gather": "k_schur_diag + k_schur_gather"}.get(dom, dom)
    n_red = 9 * N
    if dom == "factor":
        # the band factorisation is an FP64 contraction (n kd^2 flops for LDL^T, 4x for Householder QR of the
        # square band) on the FP64 tensor pipe (DMMA); peak = 64 FMA/clk/SM measured with tools/ubench
        # (DMMA m8n8k4 and DFMA both) x 148 SMs x max SM clock
        flops = float(n_red) * kd * kd   # LDL^T of the band (all variants: the QR variants refine its solution through J2bot)
        fpath = os.path.join(ROOT, "profiles", "fp64_peak.json")
        if os.path.exists(fpath):
            fp = json.load(open(fpath))
            fp64_peak = float(fp["fp64_tflops"])
            fp_src = f"measured (profiles/fp64_peak.json: {fp.get('how', '')}; {fp.get('gpu_name', '')}, {fp.get('when', '')})"
        else:
            sms = torch.cuda.get_device_properties(local_rank).multi_processor_count
            fp64_peak = 64 * 2 * sms * 1.965e9 / 1e12
            fp_src = "fallback: 64 FMA/clk/SM (tools/ubench) x SM count x 1.965 GHz"
        achieved = flops / (dom_ms * 1e-3) / 1e12 if dom_ms > 0 else 0.0
        roofline = {"bound": "tensor", "kernel": kernel_name, "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                    "frac": achieved / fp64_peak, "traffic": traffic, "peak_source": fp_src,
                    "algorithmic_flops": flops, "kernel_ms": dom_ms}
    else:
        achieved = alg_bytes[dom] / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
        roofline = {"bound": "hbm", "kernel": kernel_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                    "algorithmic_bytes": alg_bytes[dom], "kernel_ms": dom_ms}
    roofline["stages_ms"] = {n: float(v) for n, v in zip(names, stage)}
    # Point stage (Jacobian + block QR + Schur accumulation + back-substitution) against the HBM roofline by SURVEY.md
    # 8(d)'s ALGORITHMIC bytes per trial: "recompute" form (nothing stored between the factor and the back-substitution)
    # and "store" form (R_j, c_j, R12_j written once and read once). achieved_bw_frac_by_kernel divides the
    # implementation's own record traffic by the time instead: achieved bandwidth, not a roofline fraction.
    pt_ms = float(stage[names.index("k_point_factor")] + stage[names.index("k_schur_gather")] + stage[names.index("k_backsub_eval")])
    b_recompute = K * (8 + 2 * sz) + 3 * M * sz + 15 * N * sz + 3 * M * sz + 9 * N * sz + band_bytes
    b_store = b_recompute + 2 * (9 * M * sz + 27 * K * sz)
    roofline["point_stage"] = {"ms": pt_ms, "algorithmic_bytes_recompute_form": int(b_recompute), "algorithmic_bytes_store_form": int(b_store),
                               "hbm_frac_recompute_form": (b_recompute / (pt_ms * 1e-3) / 1e9 / peak) if pt_ms > 0 else None,
                               "hbm_frac_store_form": (b_store / (pt_ms * 1e-3) / 1e9 / peak) if pt_ms > 0 else None}
    roofline["achieved_bw_frac_by_kernel"] = {k: (alg_bytes[k] / (stage[names.index(k)] * 1e-3) / 1e9 / peak) if stage[names.index(k)] > 0 else None
                                              for k in ("k_point_factor", "k_schur_gather", "k_backsub_eval")}
    roofline["hbm_peak"] = {"value": peak, "unit": "GB/s", "source": peak_src}

    # ---- parity probe at this N: one LM trial on a 300-camera / ~300k-observation problem of the same synthetic family,
    # sharded exactly like the benchmark, against the CPU oracle (rank 0). Outside every timed region; the run fails if
    # the sharded GPU trial and the oracle disagree (energy 1e-12, test energy 1e-9, |dx| 1e-8).
    probe = None
    if not args.no_parity_probe:
        pfull = bal.synthetic(300, 60000, window=30, seed=20261019)
        ps = solver.GpuSolver(sharding.shard(pfull, rank, world), args.variant, args.precision, device=local_rank)
        if world > 1:
            uid2 = [None]
            if rank == 0:
                buf2 = C.create_string_buffer(128)
                assert _lib.lib().ba_comm_unique_id(buf2) == 0, _lib.lib().ba_last_error()
                uid2[0] = buf2.raw
            dist.broadcast_object_list(uid2, src=0)
            ps.set_bandwidth(sharding.global_bandwidth(pfull))
            ps.comm_init(rank, world, uid2[0])
        pe, pcn2, pcn = ps.linearize(colnorms=True)
        plam = 1e-6 * pcn if args.variant == "MOREQR" else 1e-12 * pcn2
        ps.compute(plam)
        pdxn, _, pet = ps.solve_try()
        ps.close()
        if rank == 0:
            from oracle.binding import VARIANTS as OV, Oracle
            po = Oracle(pfull)
            po.set_threads(cpu_threads())
            oe, ocn2, ocn = po.linearize()
            if args.variant == "MOREQR":
                po.moreqr_outer()
            ok, odx = po.step(OV[args.variant], plam)
            oet = po.energy_at(odx)
            r = lambda a, b: abs(a - b) / abs(b)
            probe = {"problem": f"synthetic 300 cameras / 60000 points / {pfull.K} observations, {world} rank(s)",
                     "energy_rel_err": r(pe, oe), "energy_test_rel_err": r(pet, oet), "dx_norm_rel_err": r(pdxn, float(np.linalg.norm(odx)))}
            tol = (1e-5, 1e-3, 1e-2) if args.precision == "f32" else (1e-12, 1e-9, 1e-8)
            probe["ok"] = bool(ok and probe["energy_rel_err"] < tol[0] and probe["energy_test_rel_err"] < tol[1] and probe["dx_norm_rel_err"] < tol[2])
            if not probe["ok"]:
                raise SystemExit(f"parity probe failed at {world} rank(s): {probe}")

    # ---- the other variant BASELINE config 5 names (QRKIT next to the QRCHOL headline, or vice versa): same step, device time
    other = None
    if args.workload == "synthetic-5m" and args.variant in ("QRCHOL", "QRKIT") and not args.no_other_variant:
        ov = "QRKIT" if args.variant == "QRCHOL" else "QRCHOL"
        s2 = solver.GpuSolver(prob, ov, args.precision, device=local_rank)
        if world > 1:
            uid3 = [None]
            if rank == 0:
                buf3 = C.create_string_buffer(128)
                assert _lib.lib().ba_comm_unique_id(buf3) == 0, _lib.lib().ba_last_error()
                uid3[0] = buf3.raw
            dist.broadcast_object_list(uid3, src=0)
            s2.set_bandwidth(sharding.global_bandwidth(full))
            s2.comm_init(rank, world, uid3[0])
        s2.linearize(colnorms=True)

        def step2():
            s2.linearize(colnorms=False)
            s2.compute(lam)
            out = s2.solve_try()
            s2.reject()
            return out

        for _ in range(3):
            step2()
        nst = max(3, min(args.steps, 10))
        barrier()
        s2.timer_start()
        for _ in range(nst):
            step2()
        ms2 = max_over_ranks(s2.timer_stop()) / nst
        barrier()
        s2.set_profiling(True)
        step2()
        st2 = s2.stage_ms()
        other = {"variant": ov, "ms_per_step": ms2, "steps": nst, "stages_ms": {n_: float(v) for n_, v in zip(names, st2)}}
        s2.close()

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:  # reported at N = 1 only (the scaling runs carry null)
        cpu_baseline = cpu_baseline_legs(full, args.variant)
    if rank == 0:
        line = {"metric": METRIC, "value": ms_per_step, "unit": "ms", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": False, "scaling": "strong",
                "vs_baseline": None, "dtype": args.precision,
                "data": "synthetic" if args.workload.startswith("synthetic") else "bundled BAL file",
                "config": workload_config(full, args, world), "clocks": clocks,
                "e2e": {"value": e2e_ms, "unit": "ms", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                        "steps": e2e_steps},
                "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu_baseline,
                "check": {"energy": e0, "energy_test": et, "dx_norm": dxn, "lambda": lam}, "parity_probe": probe,
                "variants": other}
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
