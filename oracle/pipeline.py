"""TEST INFRASTRUCTURE ONLY -- the whole hot path on the CPU, composed from the oracle stages.

Follows the loop body of src/vilgod/zero_shot_detector.py:389-416: per-cluster projection
(oracle/projection.py), upsample + uint8, CLIP preprocess + ViT-B/16 + scoring (oracle/vit.py),
24 -> 4 mapping and the per-cluster view vote (oracle/vote.py).  Used by tests (as the checker),
by __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs.
"""
from __future__ import annotations

import numpy as np
import torch

from . import projection as op
from . import vit as ovit
from . import vote as ovote


def torch_cpu_rotate_fused(n_points):
    """torch-CPU bmm rule the reference's rotation runs into: 9N < 400 -> naive unfused loop."""
    return 9 * int(n_points) >= 400


def project(points, offsets, num_views, threads=None, rotate_rule="torch_cpu", want_dens=False):
    """-> (dens [C,V,110,110] or None, u8 [C,V,224,224])."""
    rot = op.view_rot_mats(num_views)
    offsets = np.asarray(offsets, dtype=np.int64)
    n = np.diff(offsets)
    if rotate_rule != "torch_cpu":
        return op.project_batch_threaded(points, offsets, rot, threads=threads,
                                         fused=(rotate_rule == "fused"), want_dens=want_dens)
    C, V = len(n), num_views
    u8 = np.empty((C, V, op.S_DEFAULT, op.S_DEFAULT), np.uint8)
    dens = np.empty((C, V, op.R_DEFAULT - 2, op.R_DEFAULT - 2), np.float32) if want_dens else None
    for fused in (False, True):
        sel = np.flatnonzero((9 * n >= 400) == fused)
        if len(sel) == 0:
            continue
        pts = np.concatenate([points[offsets[c]:offsets[c + 1]] for c in sel])
        off = np.zeros(len(sel) + 1, np.int32)
        off[1:] = np.cumsum(n[sel])
        d, u = op.project_batch_threaded(pts, off, rot, threads=threads, fused=fused,
                                         want_dens=want_dens)
        u8[sel] = u
        if want_dens:
            dens[sel] = d
    return dens, u8


@torch.no_grad()
def classify(points, offsets, num_views, weights, text_features, threads=None, batch=50):
    """-> dict(u8, probs [C,V,24], logits, feats, top1 [C,V], voted_name [C], voted_score [C])."""
    _, u8 = project(points, offsets, num_views, threads=threads)
    C, V = u8.shape[:2]
    feats = ovit.encode_u8(weights, u8.reshape(C * V, *u8.shape[2:]), batch=batch)
    probs, logits, fn = ovit.score(feats, text_features)
    probs_np = probs.numpy()
    top1, names, scores = ovit.top1_labels(probs_np, ovote.CLASS_LIST)
    mapped = np.asarray(ovote.MAPPED)[top1].reshape(C, V)
    vname, vscore = ovote.vote(mapped, scores.reshape(C, V))
    return dict(u8=u8, probs=probs_np.reshape(C, V, -1), logits=logits.numpy().reshape(C, V, -1),
                feats=fn.numpy().reshape(C, V, -1), top1=top1.reshape(C, V),
                scores=scores.reshape(C, V).astype(np.float32), voted_name=vname,
                voted_score=vscore)
