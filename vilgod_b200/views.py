"""Init-time tables of the projection: view rotations and the Gaussian smoothing weights.

Host-side mirror of ``RealisticProjection.__init__`` (reference ``src/utils/mv_utils.py:133-171``),
``euler2mat`` (``:40-88``) and ``get3DGaussianKernel`` (``:204-220``).  They are evaluated with the
same fp32 torch / numpy operations as the reference so the tables are bit-identical
(tests/test_host_side.py checks them against tests/golden/tables.npz).
"""
from __future__ import annotations

import numpy as np
import torch

# (x, y, z) euler angles in units of pi.  4 = the reference's live view list, 6 = the same list
# with its two commented rows restored, 10 = the commented PointCLIPv2 block (mv_utils.py:134-153).
_VIEW_SETS = {
    4: [(0, 0, 0), (-1 / 10, 0, 0), (0, 1 / 30, 0), (0, -1 / 30, 0)],
    6: [(0, 0, 0), (-1 / 10, 0, 0), (0, 1 / 30, 0), (0, -1 / 30, 0),
        (-1 / 10, 1 / 30, 0), (-1 / 10, -1 / 30, 0)],
    10: [(1 / 4, 0, 1 / 2), (3 / 4, 0, 1 / 2), (5 / 4, 0, 1 / 2), (7 / 4, 0, 1 / 2),
         (0, 0, 1 / 2), (1 / 2, 0, 1 / 2), (1, 0, 1 / 2), (3 / 2, 0, 1 / 2),
         (0, -1 / 2, 1 / 2), (0, 1 / 2, 1 / 2)],
}


def view_angles(num_views: int) -> np.ndarray:
    if num_views not in _VIEW_SETS:
        raise ValueError(f"no built-in view set with {num_views} views; pass rot_mat explicitly")
    # written as k * np.pi / d in the reference; (k/d) * pi differs in the last f64 bit for some
    # entries, so rebuild the exact expressions
    out = []
    for row in _VIEW_SETS[num_views]:
        out.append([_angle(v) for v in row])
    return np.asarray(out, dtype=np.float64)


def _angle(frac):
    from fractions import Fraction
    f = Fraction(frac).limit_denominator(60)
    if f == 0:
        return 0.0
    sign = -1.0 if f < 0 else 1.0
    num, den = abs(f.numerator), f.denominator
    # reference spellings: np.pi / d, -np.pi / d, k * np.pi / d
    val = (num * np.pi / den) if num != 1 else (np.pi / den)
    return sign * val


def euler_to_rot_mats(angles_xyz) -> torch.Tensor:
    """[V,3] angles -> rot_mat [V,3,3] = (Rx @ Ry @ Rz)^T, fp32, as euler2mat(...).transpose(1,2)."""
    a = torch.as_tensor(np.asarray(angles_xyz)).float()
    ax, ay, az = a[:, 0], a[:, 1], a[:, 2]
    o, z = torch.ones_like(ax), torch.zeros_like(ax)
    cx, sx, cy, sy, cz, sz = (torch.cos(ax), torch.sin(ax), torch.cos(ay), torch.sin(ay),
                              torch.cos(az), torch.sin(az))
    rx = torch.stack([o, z, z, z, cx, -sx, z, sx, cx], dim=1).reshape(-1, 3, 3)
    ry = torch.stack([cy, z, sy, z, o, z, -sy, z, cy], dim=1).reshape(-1, 3, 3)
    rz = torch.stack([cz, -sz, z, sz, cz, z, z, z, o], dim=1).reshape(-1, 3, 3)
    return (rx @ ry @ rz).transpose(1, 2).contiguous()


def view_rot_mats(num_views: int) -> torch.Tensor:
    return euler_to_rot_mats(view_angles(num_views))


def gaussian_weights(ksize: int = 3, sigma: float = 3.0, zsigma: float = 1.0) -> torch.Tensor:
    """[3,3] fp32 weight of the (1,3,3) smoothing convolution: normalised 2-D Gaussian times the
    depth-1 z kernel (exp(0) = 1), normalised again -- same op sequence as the reference."""
    xs = np.arange(ksize, dtype=np.float32) - ksize // 2
    g = np.exp(-(xs ** 2) / (2 * sigma ** 2))
    k2 = torch.from_numpy(g[:, None] @ g[None, :])
    k2 = k2 / k2.sum()
    zk = np.exp(-((np.arange(1, dtype=np.float32) - 0) ** 2) / (2 * zsigma ** 2))
    k3 = np.repeat(k2[None], 1, axis=0) * zk[:, None, None]
    k3 = k3 / torch.sum(k3)
    return torch.as_tensor(k3, dtype=torch.float32).reshape(ksize, ksize).contiguous()
