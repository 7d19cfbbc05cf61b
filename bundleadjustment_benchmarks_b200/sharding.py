"""Point sharding across the GPUs of one box.

Points (and, because observations are sorted by point, contiguous observation ranges) are split
into ``nranks`` contiguous ranges balanced by sum_j n_j(n_j+1)/2 (the Schur-accumulation work), the
camera parameters are replicated. Every rank then owns a self-contained BALProblem with LOCAL point
indices; the only exchange per LM trial is the all-reduce of the reduced camera system and three
scalars (done inside the C ABI with NCCL).
"""
from __future__ import annotations

import dataclasses
from typing import List, Tuple

import numpy as np

from .bal import BALProblem


def point_ranges(prob: BALProblem, nranks: int) -> List[Tuple[int, int]]:
    if nranks < 1 or prob.M < nranks:
        raise ValueError(f"cannot shard {prob.M} points over {nranks} ranks: every rank needs at least one point")
    counts = np.bincount(prob.point, minlength=prob.M).astype(np.float64)
    work = counts * (counts + 1.0) / 2.0 + 4.0 * counts  # pair blocks + per-observation Jacobians
    cum = np.cumsum(work)
    total = cum[-1]
    cuts = [0]
    for r in range(1, nranks):
        cuts.append(int(np.searchsorted(cum, total * r / nranks)))
    cuts.append(prob.M)
    # every rank needs at least one point (an empty shard cannot create a handle and the other ranks would wait for it
    # inside the collectives): keep the cuts strictly increasing from the left, then leave room on the right
    for i in range(1, nranks):
        cuts[i] = max(cuts[i], cuts[i - 1] + 1)
    for i in range(nranks - 1, 0, -1):
        cuts[i] = min(cuts[i], cuts[i + 1] - 1)
    return [(cuts[r], cuts[r + 1]) for r in range(nranks)]


def shard(prob: BALProblem, rank: int, nranks: int) -> BALProblem:
    """Rank's slice: its points + their observations (local point indices), all cameras."""
    assert prob.is_sorted_by_point()
    if nranks == 1:
        return prob
    p0, p1 = point_ranges(prob, nranks)[rank]
    off = prob.point_offsets()
    o0, o1 = int(off[p0]), int(off[p1])
    return dataclasses.replace(
        prob,
        view=np.ascontiguousarray(prob.view[o0:o1]),
        point=np.ascontiguousarray(prob.point[o0:o1] - p0),
        meas=np.ascontiguousarray(prob.meas[o0:o1]),
        X=np.ascontiguousarray(prob.X[p0:p1]),
        name=f"{prob.name}[{rank}/{nranks}]",
        perm=None,
    )


def global_bandwidth(prob: BALProblem) -> int:
    """Block half-bandwidth of the reduced camera matrix: max over points of (max cam - min cam)."""
    off = prob.point_offsets()
    first = prob.view[off[:-1]]
    last = prob.view[off[1:] - 1]
    return int((last - first).max())
