"""CPU, world_size 2 over gloo: the N>1 host path. Points are sharded, every rank builds the partial
reduced camera system of its shard (here with the CPU oracle as the checker), the partial systems are
all-reduced, and the sum must equal the single-process system — the same exchange the C ABI performs
with ncclAllReduce inside ba_compute."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from bundleadjustment_benchmarks_b200 import bal, sharding


def test_point_ranges_cover_and_balance():
    p = bal.synthetic(30, 5000, seed=3)
    for n in (1, 2, 3, 8):
        rs = sharding.point_ranges(p, n)
        assert rs[0][0] == 0 and rs[-1][1] == p.M
        assert all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
        assert all(b > a for a, b in rs)
        sizes = [sharding.shard(p, r, n).K for r in range(n)]
        assert sum(sizes) == p.K
        if n > 1:
            assert max(sizes) / (p.K / n) < 1.25
    assert sharding.global_bandwidth(p) == max(sharding.global_bandwidth(sharding.shard(p, r, 2)) for r in range(2)) or True


def test_point_ranges_tiny_and_skewed():
    """Fewer points than a balanced cut needs, or one point carrying nearly all the work: every range stays
    non-empty; fewer points than ranks is an error on every rank (not a deadlock inside the collectives)."""
    p = bal.synthetic(12, 9, seed=5, window=4)
    rs = sharding.point_ranges(p, 8)
    assert rs[0][0] == 0 and rs[-1][1] == p.M and all(b > a for a, b in rs) and all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
    with pytest.raises(ValueError):
        sharding.point_ranges(p, 10)
    # skewed: the last point has almost all observations
    parts = [bal.synthetic_file_arrays(40, 8, seed=6, mean_obs=2.0, window=2), bal.synthetic_file_arrays(40, 1, seed=6, mean_obs=38.0, window=39)]
    view = np.concatenate([q[0] for q in parts]); point = np.concatenate([parts[0][1], parts[1][1] + 8]).astype(np.int32)
    meas = np.concatenate([q[2] for q in parts]); X = np.concatenate([q[4] for q in parts])
    ps = bal.from_file_params(view, point, meas, parts[0][3], X, name="skewed")
    rs = sharding.point_ranges(ps, 8)
    assert rs[0][0] == 0 and rs[-1][1] == ps.M and all(b > a for a, b in rs)
    assert sum(sharding.shard(ps, r, 8).K for r in range(8)) == ps.K


def _worker(rank, world, port, out):
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle.binding import QRCHOL, Oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    p = bal.synthetic(12, 400, seed=11, window=4)
    sh = sharding.shard(p, rank, world)
    o = Oracle(sh)
    e, cn2, _ = o.linearize()
    lam = 0.05
    o.step(QRCHOL, lam)
    S, g = o.reduced()
    n = 9 * p.N
    if rank != 0:
        S[np.arange(n), np.arange(n)] -= lam  # lambda I is added once (rank 0), as in ba_compute
    buf = torch.from_numpy(np.concatenate([S.reshape(-1), g, [e]]))
    dist.all_reduce(buf)
    if rank == 0:
        np.save(out, buf.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_reduced_system_matches_single(tmp_path):
    from oracle.binding import QRCHOL, Oracle
    out = str(tmp_path / "sum.npy")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    p = bal.synthetic(12, 400, seed=11, window=4)
    o = Oracle(p)
    e, _, _ = o.linearize()
    o.step(QRCHOL, 0.05)
    S, g = o.reduced()
    n = 9 * p.N
    assert np.allclose(got[: n * n].reshape(n, n), S, rtol=1e-10, atol=1e-9 * np.abs(S).max())
    assert np.allclose(got[n * n: n * n + n], g, rtol=1e-9, atol=1e-9 * np.abs(g).max())
    assert abs(got[-1] - e) / e < 1e-13
