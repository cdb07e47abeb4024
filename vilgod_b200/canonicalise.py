"""Host-side cluster canonicalisation (SURVEY.md section 8 row a0 / "next" row f1).

Mirrors, with the same dtypes at every step, what ``ZeroShotDetector.classification`` does to a
cluster before it reaches the projection (reference ``src/vilgod/zero_shot_detector.py:391-394``):

  ``apply_transform``                      reference ``src/utils/pointcloud_utils.py:21-46``
  ``transform_cluster_points_to_origin``   reference ``src/utils/pointcloud_utils.py:390-412``

plus a vectorised variant over a packed ragged batch (one call per frame instead of one per
cluster) that produces the same values.  The arithmetic is float64 inside, like the reference, so
the result is ``.float()``-ed once at the very end.
"""
from __future__ import annotations

import numpy as np
from scipy.spatial.transform import Rotation as R


def apply_transform(pts, transformation):
    """pts [N,>=3] -> same dtype/shape with xyz replaced by (T @ [x y z 1]^T)[:3]
    (mode='left', box=False of the reference)."""
    if len(pts) == 0:
        return pts
    out = np.array(pts, copy=True)
    homog = np.hstack((out[:, :3], np.ones((len(out), 1))))
    out[..., :3] = np.einsum('ij,kj->ki', transformation, homog)[..., :3]
    return out


def _image_axes_matrix():
    rot = np.eye(4)
    rot[:3, :3] = R.from_euler('x', np.pi).as_matrix() @ R.from_euler('z', np.pi / 2.).as_matrix()
    return rot


def transform_cluster_points_to_origin(points):
    """One cluster [N,3] (ego frame) -> float64 [N,3] in image coordinates: subtract the xy
    median, rotate by -atan2(cy, cx) about z, shift x by -1, reorder to (z, y, x), apply
    Rx(pi) @ Rz(pi/2)."""
    pts = np.array(points, copy=True)
    center = np.median(pts[..., :3], axis=0)
    angle = np.arctan2(center[1], center[0])
    pts[..., :2] -= center[:2]
    pts = R.from_euler('z', -angle).apply(pts)
    pts[..., 0] -= 1
    pts = np.stack([pts[:, 2], pts[:, 1], pts[:, 0]], axis=1)
    return apply_transform(pts, _image_axes_matrix())


def canonicalise_cluster(points, transform_to_ego=None):
    """Exactly the three host lines of zero_shot_detector.py:391-394 -> float32 [N,3]."""
    pts = np.asarray(points)[..., :3]
    if transform_to_ego is not None:
        pts = apply_transform(pts, transform_to_ego)
    return transform_cluster_points_to_origin(pts).astype(np.float32)


def _segment_median(values, seg_ids, offsets):
    """Per-segment median of a 1-D array with numpy.median's arithmetic (mean of the two middle
    order statistics, in the array's dtype)."""
    order = np.lexsort((values, seg_ids))
    sv = values[order]
    n = np.diff(offsets)
    lo = offsets[:-1] + (n - 1) // 2
    hi = offsets[:-1] + n // 2
    a, b = sv[lo], sv[hi]
    med = ((a + b) / values.dtype.type(2)).astype(values.dtype)
    return np.where(lo == hi, a, med)


def canonicalise_packed(points, offsets, transform_to_ego=None):
    """Packed ragged batch: points [sum N, 3], offsets [C+1] -> float32 [sum N, 3], cluster by
    cluster identical to ``canonicalise_cluster``."""
    pts = np.asarray(points)[..., :3]
    offsets = np.asarray(offsets, dtype=np.int64)
    if transform_to_ego is not None:
        pts = apply_transform(pts, transform_to_ego)
    else:
        pts = np.array(pts, copy=True)
    n = np.diff(offsets)
    if np.any(n < 0):
        raise ValueError("offsets must be non-decreasing")
    if np.any(n == 0):
        # clusters without points contribute nothing to the packed array: canonicalise the others
        # (the projection reports VG_EDEGENERATE for the empty ones)
        keep = np.flatnonzero(n > 0)
        if keep.size == 0:
            return np.zeros((0, 3), dtype=np.float32)
        offsets = np.concatenate([offsets[:-1][keep][:1], offsets[1:][keep]])
        n = np.diff(offsets)
    seg = np.repeat(np.arange(len(n)), n)
    cx = _segment_median(np.ascontiguousarray(pts[:, 0]), seg, offsets)
    cy = _segment_median(np.ascontiguousarray(pts[:, 1]), seg, offsets)
    angle = np.arctan2(cy, cx)
    pts[:, 0] -= cx[seg]
    pts[:, 1] -= cy[seg]
    mats = R.from_euler('z', (-angle.astype(np.float64))[:, None]).as_matrix()   # [C,3,3]
    p64 = np.einsum('nij,nj->ni', mats[seg], pts.astype(np.float64))
    p64[:, 0] -= 1
    p64 = np.stack([p64[:, 2], p64[:, 1], p64[:, 0]], axis=1)
    return apply_transform(p64, _image_axes_matrix()).astype(np.float32)
