"""N-GPU vs 1-GPU equivalence of one LM trial (run under torchrun): the sharded step must reproduce the
single-GPU energy, |dx| and rho denominator to rounding."""
import ctypes as C, os, sys
sys.path.insert(0, ".")
import numpy as np, torch, torch.distributed as dist
from bundleadjustment_benchmarks_b200 import bal, sharding, solver, _lib
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", lr))
name = sys.argv[1] if len(sys.argv) > 1 else "problem-39-18060"
full = bal.load_named(name)
for variant in ("QRCHOL", "CHOLESKY"):
    s = solver.GpuSolver(sharding.shard(full, rank, world), variant, device=lr)
    uid = [None]
    if rank == 0:
        buf = C.create_string_buffer(128); assert _lib.lib().ba_comm_unique_id(buf) == 0; uid[0] = buf.raw
    dist.broadcast_object_list(uid, src=0)
    s.set_bandwidth(sharding.global_bandwidth(full)); s.comm_init(rank, world, uid[0])
    e, cn2, cn = s.linearize()
    lam = 1e-12 * cn2
    s.compute(lam); out = s.solve_try()
    if rank == 0:
        s1 = solver.GpuSolver(full, variant, device=lr)
        e1, cn21, _ = s1.linearize(); s1.compute(1e-12 * cn21); o1 = s1.solve_try()
        rel = lambda a, b: abs(a - b) / abs(b)
        print(f"{name} {variant} x{world}: energy {rel(e, e1):.1e} cn2 {rel(cn2, cn21):.1e} |dx| {rel(out[0], o1[0]):.1e} rho_den {rel(out[1], o1[1]):.1e} e_test {rel(out[2], o1[2]):.1e}", flush=True)
        assert rel(e, e1) < 1e-12 and rel(out[2], o1[2]) < 1e-9 and rel(out[0], o1[0]) < 1e-8 and rel(out[1], o1[1]) < 1e-8
        s1.close()
    s.close(); dist.barrier()
dist.destroy_process_group()
