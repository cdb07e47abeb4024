"""Synthetic Waymo / Argoverse-2 shaped cluster batches.

There is no dataset in the build or bench environment, so every test and benchmark input comes
from here: box-surface point clusters of four object archetypes, log-uniform point counts, placed
at 5-75 m and run through the host canonicalisation (``canonicalise.py``) so the projection sees
the coordinates the reference's classification loop would feed it.
"""
from __future__ import annotations

import numpy as np

from . import canonicalise

DEFAULT_SEED = 20240807
ARCHETYPES = np.array([[4.6, 1.9, 1.7],     # vehicle
                       [0.7, 0.7, 1.75],    # pedestrian
                       [1.8, 0.7, 1.7],     # cyclist
                       [0.4, 0.4, 3.0]])    # pole / background


def make_clusters_raw(num_clusters, n_min=10, n_max=2048, seed=DEFAULT_SEED, rng=None):
    """-> (points f32 [sum N,3] in the ego frame, offsets i32 [C+1], archetype i32 [C])."""
    rng = rng or np.random.default_rng(seed)
    C = int(num_clusters)
    kind = rng.integers(0, len(ARCHETYPES), size=C)
    n = np.exp(rng.uniform(np.log(n_min), np.log(n_max + 1), size=C)).astype(np.int64)
    n = np.clip(n, n_min, n_max)
    offsets = np.zeros(C + 1, dtype=np.int64)
    np.cumsum(n, out=offsets[1:])
    total = int(offsets[-1])
    seg = np.repeat(np.arange(C), n)
    p = rng.uniform(-0.5, 0.5, size=(total, 3))
    axis = rng.integers(0, 3, size=total)
    side = rng.integers(0, 2, size=total) - 0.5
    p[np.arange(total), axis] = side
    p *= ARCHETYPES[kind][seg]
    p += rng.normal(0.0, 0.02, size=(total, 3))
    heading = rng.uniform(-np.pi, np.pi, size=C)
    ch, sh = np.cos(heading)[seg], np.sin(heading)[seg]
    x = ch * p[:, 0] - sh * p[:, 1]
    y = sh * p[:, 0] + ch * p[:, 1]
    rng_m = rng.uniform(5.0, 75.0, size=C)
    az = rng.uniform(-np.pi, np.pi, size=C)
    x += (rng_m * np.cos(az))[seg]
    y += (rng_m * np.sin(az))[seg]
    z = p[:, 2] + (ARCHETYPES[kind][:, 2] / 2)[seg]
    pts = np.stack([x, y, z], axis=1).astype(np.float32)
    return pts, offsets.astype(np.int32), kind.astype(np.int32)


def make_clusters(num_clusters, n_min=10, n_max=2048, seed=DEFAULT_SEED, rng=None):
    """Canonicalised clusters ready for the projection: (points f32 [sum N,3], offsets i32 [C+1])."""
    pts, offsets, _ = make_clusters_raw(num_clusters, n_min, n_max, seed, rng)
    return canonicalise.canonicalise_packed(pts, offsets), offsets


def make_frames(num_frames, clusters_per_frame=300, n_min=10, n_max=2048, seed=DEFAULT_SEED):
    """Waymo-shaped batch: list of (points, offsets) per frame, C_f ~ Poisson(clusters_per_frame)."""
    rng = np.random.default_rng(seed)
    frames = []
    for _ in range(num_frames):
        c = max(1, int(rng.poisson(clusters_per_frame)))
        frames.append(make_clusters(c, n_min, n_max, rng=rng))
    return frames


def concat_frames(frames):
    """list of (points, offsets) -> one packed batch + frame boundaries (in clusters)."""
    pts = np.concatenate([f[0] for f in frames])
    bounds = [0]
    offs = [np.zeros(1, dtype=np.int64)]
    base = 0
    for p, o in frames:
        offs.append(o[1:].astype(np.int64) + base)
        base += int(o[-1])
        bounds.append(bounds[-1] + len(o) - 1)
    return pts, np.concatenate(offs).astype(np.int32), np.asarray(bounds, dtype=np.int64)


def make_sequence_raw(num_frames, clusters_per_frame=150, n_min=10, n_max=2048, seed=DEFAULT_SEED):
    """One fixed sequence in the form the classification loop receives it (zero_shot_detector.py:365,
    389-393): per frame the RAW fp32 cluster points in the sensor frame, their packed offsets and the
    frame's 4x4 transform_to_ego -- canonicalisation still to be done.  Deterministic in (seed, frame)."""
    frames = []
    for f in range(num_frames):
        rng = np.random.default_rng([seed, 7919, f])
        c = max(1, int(rng.poisson(clusters_per_frame)))
        pts, off, _ = make_clusters_raw(c, n_min, n_max, rng=rng)
        a = rng.uniform(-np.pi, np.pi)
        T = np.eye(4)
        T[:2, :2] = [[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]]
        T[:3, 3] = rng.uniform(-2.0, 2.0, size=3)
        frames.append((pts, off, T))
    return frames
