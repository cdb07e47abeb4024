"""Size-independent properties of the projection the oracle must have (and the GPU path is held to the same
ones at full BASELINE size in tests/test_e2e_gpu.py / test_projection_gpu.py): they follow from the
reference's arithmetic (src/utils/mv_utils.py:91-127, 30-37, 173-201), not from any particular input.

  * point order does not matter (scatter-max and per-axis min / max are order independent);
  * repeating points changes nothing (max is idempotent);
  * scaling a cluster by a power of two is exact in fp32, and (p - centre) / range cancels it bit for bit;
  * clusters of a packed batch do not see each other (ragged offsets, any split of the batch);
  * every image has background 1.0, its deepest smoothed pixel at exactly 0.0, values within [0, 1];
  * the views are independent: view v of a V-view table equals the same rotation applied alone.
"""
import numpy as np
import pytest

from oracle import pipeline as opipe
from oracle import projection as op
from vilgod_b200 import synthetic


def _project(points, rot, fused):
    off = np.array([0, len(points)], np.int32)
    dens, u8 = op.project_batch(points, off, rot, fused=fused)
    return dens[0], u8[0]


@pytest.fixture(scope="module")
def clusters():
    rng = np.random.default_rng(20240807)
    pts, off = synthetic.make_clusters(6, n_min=30, n_max=900, rng=rng)
    return [pts[off[c]:off[c + 1]] for c in range(len(off) - 1)]


def test_point_order_and_duplicates_do_not_matter(clusters):
    rot = op.view_rot_mats(6)
    rng = np.random.default_rng(1)
    for p in clusters:
        fused = 9 * len(p) >= 400
        d0, u0 = _project(p, rot, fused)
        perm = rng.permutation(len(p))
        d1, u1 = _project(p[perm], rot, fused)
        assert np.array_equal(d0, d1) and np.array_equal(u0, u1)
        # every point twice (the bmm rule is kept: it depends on N in the reference)
        d2, u2 = _project(np.concatenate([p, p[perm]]), rot, fused)
        assert np.array_equal(d0, d2) and np.array_equal(u0, u2)


@pytest.mark.parametrize("k", [-7, 3, 11])
def test_power_of_two_scale_cancels_exactly(clusters, k):
    rot = op.view_rot_mats(4)
    for p in clusters[:3]:
        fused = 9 * len(p) >= 400
        d0, u0 = _project(p, rot, fused)
        d1, u1 = _project((p * np.float32(2.0 ** k)).astype(np.float32), rot, fused)
        assert np.array_equal(d0, d1) and np.array_equal(u0, u1)


def test_clusters_of_a_batch_are_independent(clusters):
    V = 10
    pts = np.concatenate(clusters)
    off = np.zeros(len(clusters) + 1, np.int32)
    off[1:] = np.cumsum([len(c) for c in clusters])
    dens_all, u8_all = opipe.project(pts, off, V, want_dens=True)
    for c, p in enumerate(clusters):
        dens_c, u8_c = opipe.project(p, np.array([0, len(p)], np.int32), V, want_dens=True)
        assert np.array_equal(dens_all[c], dens_c[0]) and np.array_equal(u8_all[c], u8_c[0])
    # an empty cluster between two others is rejected by the oracle like a degenerate one, not skipped silently
    with pytest.raises(ValueError):
        opipe.project(pts, np.insert(off, 2, off[2]), V)


def test_image_range_background_and_deepest_pixel(clusters):
    rot = op.view_rot_mats(10)
    for p in clusters:
        dens, u8 = _project(p, rot, 9 * len(p) >= 400)
        assert dens.min() == 0.0 and dens.max() == 1.0          # 1 - x / max(x): the maximum maps to exactly 0
        assert np.all((dens >= 0.0) & (dens <= 1.0))
        # the grid keeps cells 1 .. R-2 and obj_ratio = 0.8 leaves a margin: the image border is background
        assert np.all(dens[:, 0, :] == 1.0) and np.all(dens[:, :, 0] == 1.0)
        assert np.all(dens[:, -1, :] == 1.0) and np.all(dens[:, :, -1] == 1.0)
        assert u8.min() == 0 and u8.max() in (254, 255)         # floor(x * 255) with torch-CPU's bilinear rounding
        assert (dens == 1.0).mean() > 0.4                        # most of a depth image is background


def test_views_are_independent(clusters):
    rot10 = op.view_rot_mats(10)
    for p in clusters[:3]:
        fused = 9 * len(p) >= 400
        d10, u10 = _project(p, rot10, fused)
        for v in (0, 3, 9):
            d1, u1 = _project(p, rot10[v:v + 1], fused)
            assert np.array_equal(d10[v], d1[0]) and np.array_equal(u10[v], u1[0])
    # every view is a rotation: orthonormal to fp32 accuracy, determinant +1
    for V in (4, 6, 10):
        r = op.view_rot_mats(V).astype(np.float64)
        assert np.abs(r @ r.transpose(0, 2, 1) - np.eye(3)).max() < 1e-6
        assert np.allclose(np.linalg.det(r), 1.0, atol=1e-6)
