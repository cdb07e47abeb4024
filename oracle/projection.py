"""TEST INFRASTRUCTURE ONLY -- Python face of the C projection oracle (oracle/projection_oracle.c).

Restates, for checking purposes, the reference's
  * view table and rotation matrices   src/utils/mv_utils.py:40-88,134-166
  * Gaussian smoothing weights         src/utils/mv_utils.py:204-220
and wraps the C functions that restate points2grid / GridToImage / the upsample+uint8 glue.
The product (vilgod_b200/) must never import this module.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None

R_DEFAULT, D_DEFAULT, S_DEFAULT = 112, 8, 224
OBJ_RATIO, DEPTH_BIAS = 0.8, 0.2


def build(force=False):
    src = os.path.join(_HERE, "projection_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = ctypes.CDLL(_LIB_PATH)
        fp = ctypes.POINTER(ctypes.c_float)
        i64p = ctypes.POINTER(ctypes.c_int64)
        i32p = ctypes.POINTER(ctypes.c_int32)
        u8p = ctypes.POINTER(ctypes.c_uint8)
        L.vgo_rotate.argtypes = [fp, ctypes.c_int64, fp, ctypes.c_int, fp]
        L.vgo_points2grid.argtypes = [fp, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                                      ctypes.c_double, ctypes.c_double, fp, i64p, fp]
        L.vgo_densify.argtypes = [fp, ctypes.c_int, ctypes.c_int, fp, fp, fp, fp]
        L.vgo_upsample_u8.argtypes = [fp, ctypes.c_int, ctypes.c_int, fp, u8p]
        L.vgo_project_batch.argtypes = [fp, i32p, ctypes.c_int, fp, ctypes.c_int, ctypes.c_int,
                                        ctypes.c_int, ctypes.c_int, ctypes.c_double,
                                        ctypes.c_double, fp, ctypes.c_int, fp, u8p]
        L.vgo_set_div_mode.argtypes = [ctypes.c_int]
        L.vgo_set_div_mode.restype = None
        for f in (L.vgo_rotate, L.vgo_points2grid, L.vgo_densify, L.vgo_upsample_u8,
                  L.vgo_project_batch):
            f.restype = ctypes.c_int
        _lib = L
    return _lib


def set_div_mode(mode):
    """0: true division by (1 + depth_bias) (torch-CPU, the pinned reference behaviour);
    1: multiplication by the fp32 reciprocal (torch-CUDA's scalar division).  Process-wide."""
    lib().vgo_set_div_mode(int(mode))


def _f(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float)) if a is not None else None


# ------------------------------------------------------------------------------------------------
# init-time tables
# ------------------------------------------------------------------------------------------------
def view_angles(num_views):
    """(x, y, z) euler angles of the 4 live views, the 6-view set (two commented rows restored) and
    the 10-view PointCLIPv2 set -- src/utils/mv_utils.py:134-153."""
    pi = np.pi
    if num_views in (1, 4, 6):
        v = [[0, 0, 0], [-pi / 10, 0, 0], [0, pi / 30, 0], [0, -pi / 30, 0],
             [-pi / 10, pi / 30, 0], [-pi / 10, -pi / 30, 0]]
        return np.asarray(v[:num_views], dtype=np.float64)
    if num_views == 10:
        return np.asarray(
            [[k * pi / 4, 0, pi / 2] for k in (1, 3, 5, 7)]
            + [[k * pi / 2, 0, pi / 2] for k in (0, 1, 2, 3)]
            + [[0, -pi / 2, pi / 2], [0, pi / 2, pi / 2]], dtype=np.float64)
    raise ValueError(f"no view table for {num_views} views")


def view_rot_mats(num_views):
    """rot_mat = (Rx @ Ry @ Rz)^T in fp32 torch arithmetic -- mv_utils.py:40-88 and :165-166."""
    import torch

    ang = torch.tensor(view_angles(num_views)).float()
    x, y, z = ang[:, 0], ang[:, 1], ang[:, 2]
    zero, one = torch.zeros_like(z), torch.ones_like(z)
    cz, sz, cy, sy, cx, sx = (torch.cos(z), torch.sin(z), torch.cos(y), torch.sin(y),
                              torch.cos(x), torch.sin(x))
    zmat = torch.stack([cz, -sz, zero, sz, cz, zero, zero, zero, one], 1).reshape(-1, 3, 3)
    ymat = torch.stack([cy, zero, sy, zero, one, zero, -sy, zero, cy], 1).reshape(-1, 3, 3)
    xmat = torch.stack([one, zero, zero, zero, cx, -sx, zero, sx, cx], 1).reshape(-1, 3, 3)
    rot = (xmat @ ymat @ zmat).transpose(1, 2).contiguous()
    return rot.numpy().astype(np.float32)


def gaussian_weights(ksize=3, sigma=3.0, zsigma=1.0):
    """Conv3d weight [3,3] -- mv_utils.py:204-220 with kernel_size (1,3,3): 2-D kernel normalised,
    multiplied by the depth-1 z-kernel exp(0)=1 and normalised again (as torch fp32 ops)."""
    import torch

    center = ksize // 2
    xs = np.arange(ksize, dtype=np.float32) - center
    k1 = np.exp(-(xs ** 2) / (2 * sigma ** 2))
    k2 = torch.from_numpy(k1[..., None] @ k1[None, ...])
    k2 = k2 / k2.sum()
    zs = np.arange(1, dtype=np.float32) - 0
    zk = np.exp(-(zs ** 2) / (2 * zsigma ** 2))
    k3 = np.repeat(k2[None, :, :], 1, axis=0) * zk[:, None, None]
    k3 = k3 / torch.sum(k3)
    return torch.Tensor(k3).numpy().astype(np.float32).reshape(3, 3)


# ------------------------------------------------------------------------------------------------
# stages
# ------------------------------------------------------------------------------------------------
def rotate(points, rot, fused=True):
    p = np.ascontiguousarray(points, dtype=np.float32)
    r = np.ascontiguousarray(rot, dtype=np.float32).reshape(9)
    out = np.empty_like(p)
    lib().vgo_rotate(_f(p), p.shape[0], _f(r), int(fused), _f(out))
    return out


def points2grid(q, R=R_DEFAULT, D=D_DEFAULT, obj_ratio=OBJ_RATIO, depth_bias=DEPTH_BIAS,
                return_cells=False):
    q = np.ascontiguousarray(q, dtype=np.float32)
    grid = np.empty((D, R, R), dtype=np.float32)
    cells = np.empty(q.shape[0], dtype=np.int64)
    vals = np.empty(q.shape[0], dtype=np.float32)
    rc = lib().vgo_points2grid(_f(q), q.shape[0], R, D, obj_ratio, depth_bias, _f(grid),
                               cells.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), _f(vals))
    if rc != 0:
        raise ValueError(f"vgo_points2grid rc={rc} (degenerate cluster)")
    return (grid, cells, vals) if return_cells else grid


def densify(grid, gauss=None, return_stages=False):
    D, R, _ = grid.shape
    g = np.ascontiguousarray(grid, dtype=np.float32)
    w = np.ascontiguousarray(gaussian_weights() if gauss is None else gauss, dtype=np.float32)
    Q = R - 2
    img = np.empty((Q, Q), dtype=np.float32)
    pooled = np.empty((D, Q, Q), dtype=np.float32) if return_stages else None
    conv = np.empty((D, Q, Q), dtype=np.float32) if return_stages else None
    rc = lib().vgo_densify(_f(g), R, D, _f(w), _f(img), _f(pooled), _f(conv))
    assert rc == 0
    return (img, pooled, conv) if return_stages else img


def upsample_u8(img, S=S_DEFAULT, return_float=False):
    a = np.ascontiguousarray(img, dtype=np.float32)
    Q = a.shape[0]
    up = np.empty((S, S), dtype=np.float32) if return_float else None
    u8 = np.empty((S, S), dtype=np.uint8)
    lib().vgo_upsample_u8(_f(a), Q, S, _f(up), u8.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)))
    return (u8, up) if return_float else u8


def project_batch(points, offsets, rot, R=R_DEFAULT, D=D_DEFAULT, S=S_DEFAULT,
                  obj_ratio=OBJ_RATIO, depth_bias=DEPTH_BIAS, gauss=None, fused=True,
                  want_dens=True, want_u8=True):
    """points [sum N, 3] f32, offsets [C+1] i32 -> dens [C,V,R-2,R-2] f32, u8 [C,V,S,S]."""
    p = np.ascontiguousarray(points, dtype=np.float32)
    o = np.ascontiguousarray(offsets, dtype=np.int32)
    r = np.ascontiguousarray(rot, dtype=np.float32)
    V, C, Q = r.shape[0], o.shape[0] - 1, R - 2
    w = np.ascontiguousarray(gaussian_weights() if gauss is None else gauss, dtype=np.float32)
    dens = np.empty((C, V, Q, Q), dtype=np.float32) if want_dens else None
    u8 = np.empty((C, V, S, S), dtype=np.uint8) if want_u8 else None
    rc = lib().vgo_project_batch(
        _f(p), o.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), C, _f(r), V, R, D, S, obj_ratio,
        depth_bias, _f(w), int(fused), _f(dens),
        u8.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)) if u8 is not None else None)
    if rc != 0:
        raise ValueError(f"vgo_project_batch rc={rc}")
    return dens, u8


def project_batch_threaded(points, offsets, rot, threads=None, **kw):
    """Same as project_batch, clusters split over host threads (ctypes drops the GIL)."""
    from concurrent.futures import ThreadPoolExecutor

    o = np.asarray(offsets, dtype=np.int64)
    C = o.shape[0] - 1
    threads = threads or os.cpu_count() or 1
    bounds = np.linspace(0, C, min(threads, max(C, 1)) + 1).astype(int)
    p = np.ascontiguousarray(points, dtype=np.float32)

    def work(i):
        a, b = bounds[i], bounds[i + 1]
        if a == b:
            return None, None
        sub = (o[a:b + 1] - o[a]).astype(np.int32)
        return project_batch(p[o[a]:o[b]], sub, rot, **kw)

    with ThreadPoolExecutor(len(bounds) - 1) as ex:
        parts = [r for r in ex.map(work, range(len(bounds) - 1)) if r[0] is not None or r[1] is not None]
    dens = np.concatenate([d for d, _ in parts]) if parts and parts[0][0] is not None else None
    u8 = np.concatenate([u for _, u in parts]) if parts and parts[0][1] is not None else None
    return dens, u8
