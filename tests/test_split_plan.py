"""CPU (-m "not gpu"): host-side plan of the separator split of the band LDL^T (ba_split_plan, csrc/ba_split.cuh): pure
arithmetic behind the C ABI, so the invariants the kernels rely on are checked without a GPU."""
import numpy as np
import pytest

from bundleadjustment_benchmarks_b200 import solver

NB = 32


def _check(n, kd, pl):
    bt = -(-kd // NB)
    w, s0, p1 = pl["w"], pl["s0"], pl["p1"]
    assert w >= kd + 1 and w % 2 == 0 and p1 == s0 + w and p1 % 2 == 0          # parts do not touch; part 1 starts 16-byte aligned
    assert pl["n0"] == s0 and pl["n1"] == n - p1 and abs(pl["n0"] - pl["n1"]) <= 3
    assert -(-(w - 1) // NB) <= 144                                               # the separator block fits the cluster kernel
    for p in (0, 1):
        npart, q, nm, ntm, npE = pl[f"n{p}"], pl[f"q{p}"], pl[f"nm{p}"], pl[f"ntm{p}"], pl[f"npE{p}"]
        assert q >= bt + 2                                                        # chains longer than the band is wide
        assert nm == npart - 2 * q * NB and kd + 1 <= nm < kd + 1 + 2 * NB        # the chains' last updates stay inside the middle block
        assert ntm == -(-nm // NB) and npE == q + ntm
        assert npE * NB < npart + NB                                              # the spike never reads past the part's rows
    b = pl["bounds"]
    assert b[0] == 0 and b[-1] == pl["q0"] and all(x < y for x, y in zip(b, b[1:]))
    if pl["segments"] >= 3:
        assert b[-1] - b[-2] <= pl["q0"] // 4 + 1                                 # last segment: a quarter of the chain


def test_plan_of_the_baseline_configuration():
    """BASELINE config 5: n = 16200, kd = 548 -> separator of 550 rows, four chains of 113 panels, middle blocks of 19 panels."""
    pl = solver.split_plan(16200, 548, 1, 3)
    assert pl["ok"] == 1
    _check(16200, 548, pl)
    assert (pl["w"], pl["q0"], pl["q1"], pl["ntm0"], pl["ntm1"]) == (550, 113, 113, 19, 19)
    assert pl["bounds"] == [0, 42, 85, 113]


@pytest.mark.parametrize("mode", [1, 2])
def test_plan_invariants_over_sizes(mode):
    rng = np.random.default_rng(17)
    used = 0
    for _ in range(1500):
        kd = int(rng.integers(1, 600))
        n = int(rng.integers(kd + 1, 40000))
        for seg in (1, 2, 3, 4):
            pl = solver.split_plan(n, kd, mode, seg)
            if pl["ok"]:
                used += 1
                _check(n, kd, pl)
                bt = -(-kd // NB)
                if mode == 1:
                    assert min(pl["q0"], pl["q1"]) >= 2 * bt + 8                  # auto mode: only when the chains pay for the extra stages
    assert used > 400


def test_plan_refuses_what_it_cannot_do():
    assert solver.split_plan(16200, 548, 0)["ok"] == 0                            # switched off
    assert solver.split_plan(2313, 2312, 2)["ok"] == 0                            # dense system: nothing to split
    assert solver.split_plan(4000, 548, 2)["ok"] == 0                             # chains shorter than the band is wide
    assert solver.split_plan(20000, 700, 2)["ok"] == 0                            # more than 18 row tiles per panel: spike kernel's ring
    assert solver.split_plan(5001, 548, 2)["ok"] == 1 and solver.split_plan(5001, 548, 1)["ok"] == 0
    with pytest.raises(solver.BAError):
        solver.split_plan(0, 5)
