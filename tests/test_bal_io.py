"""BAL reader / synthetic generator (host logic; reference: src/bundle_adjustment_large.cpp:57-107)."""
import json
import os

import numpy as np
import pytest

from bundleadjustment_benchmarks_b200 import bal

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "survey_anchors.json")))


def test_bundled_headers(p21, p39):
    assert (p21.N, p21.M, p21.K) == (21, 11315, 36455)
    assert (p39.N, p39.M, p39.K) == (39, 18060, 63551)
    assert p21.is_sorted_by_point() and p39.is_sorted_by_point()


def test_loader_conventions(p21):
    # focal negated (bundle_adjustment_large.cpp:88-90), distortion pre-scaled (:97-98)
    assert np.all(p21.f < 0)
    R = p21.R
    assert np.allclose(np.einsum("nij,nkj->nik", R, R), np.eye(3)[None], atol=1e-12)
    n = np.bincount(p21.point, minlength=p21.M)
    assert n.min() == 2


def test_rodrigues_cutoff():
    # hard |omega| <= 1e-6 -> identity (src/MathUtils.h:74, quirk Q2)
    assert np.array_equal(bal.rodrigues(np.array([5e-7, 0, 0])), np.eye(3))
    R = bal.rodrigues(np.array([0.0, 0.0, np.pi / 2]))
    assert np.allclose(R, [[0, -1, 0], [1, 0, 0], [0, 0, 1]], atol=1e-15)


def test_roundtrip_write_read(tmp_path):
    view, point, meas, cam9, X = bal.synthetic_file_arrays(5, 40, seed=3)
    p = tmp_path / "s.txt"
    bal.write_bal(str(p), view, point, meas, cam9, X)
    a = bal.read_bal(str(p))
    b = bal.from_file_params(view, point, meas, cam9, X)
    assert np.array_equal(a.view, b.view) and np.array_equal(a.point, b.point)
    assert np.allclose(a.meas, b.meas, rtol=1e-6) and np.allclose(a.X, b.X, rtol=1e-15)
    assert np.allclose(a.R, b.R, rtol=0, atol=1e-15)


def test_synthetic_shape_and_determinism():
    a = bal.synthetic(50, 2000, seed=7)
    b = bal.synthetic(50, 2000, seed=7)
    assert np.array_equal(a.view, b.view) and np.array_equal(a.meas, b.meas)
    assert a.is_sorted_by_point()
    n = np.bincount(a.point, minlength=a.M)
    assert n.min() >= 2
    key = a.point.astype(np.int64) * 1000 + a.view
    assert len(np.unique(key)) == a.K  # no duplicate (point, camera) pairs
    # window => block-banded reduced camera matrix
    off = a.point_offsets()
    span = max(a.view[off[j + 1] - 1] - a.view[off[j]] for j in range(a.M))
    assert span <= 60


@pytest.mark.parametrize("name", list(bal.STANDINS))
def test_standins_match_missing_files(name):
    N, M, K = bal.STANDINS[name]
    p = bal.load_named(name)
    assert (p.N, p.M, p.K) == (N, M, K)


def test_sort_by_point_permutation(tiny):
    rng = np.random.default_rng(0)
    perm = rng.permutation(tiny.K)
    shuffled = tiny.copy()
    shuffled.view, shuffled.point, shuffled.meas = tiny.view[perm], tiny.point[perm], tiny.meas[perm]
    assert not shuffled.is_sorted_by_point()
    s = shuffled.sorted_by_point()
    assert np.array_equal(s.view, tiny.view) and np.array_equal(s.meas, tiny.meas)
    assert np.array_equal(perm[s.perm], np.arange(tiny.K))
