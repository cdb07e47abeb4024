"""Frame sharding across ranks.

Frames are independent through projection, ViT and vote, so the path shards by frame with no
data-path collective: rank r of W takes frames f = r (mod W).  Only the final per-cluster labels
travel: one gather of (frame id, cluster index, class id, score) to rank 0 at the end.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def frames_of_rank(num_frames: int, rank: int, world: int):
    return list(range(rank, num_frames, world))


def gather_labels(frame_ids, cluster_index, voted_class, voted_score, dst: int = 0):
    """Each rank passes 1-D tensors of equal length (its clusters).  Rank ``dst`` receives the
    concatenation sorted by (frame id, cluster index) as numpy arrays; other ranks get None.
    Works with gloo (CPU tensors) and nccl (CUDA tensors)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        out = [t.detach().cpu().numpy() for t in (frame_ids, cluster_index, voted_class, voted_score)]
        order = np.lexsort((out[1], out[0]))
        return tuple(o[order] for o in out)
    world, rank = dist.get_world_size(), dist.get_rank()
    dev = voted_class.device
    n = torch.tensor([voted_class.numel()], dtype=torch.int64, device=dev)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n)
    nmax = int(max(int(c) for c in counts))
    payload = torch.zeros((nmax, 4), dtype=torch.float64, device=dev)
    k = voted_class.numel()
    payload[:k, 0] = frame_ids.to(dev, torch.float64)
    payload[:k, 1] = cluster_index.to(dev, torch.float64)
    payload[:k, 2] = voted_class.to(dev, torch.float64)
    payload[:k, 3] = voted_score.to(dev, torch.float64)
    bufs = [torch.zeros_like(payload) for _ in range(world)]
    dist.all_gather(bufs, payload)
    if rank != dst:
        return None
    rows = torch.cat([b[:int(c)] for b, c in zip(bufs, counts)]).cpu().numpy()
    order = np.lexsort((rows[:, 1], rows[:, 0]))
    rows = rows[order]
    return (rows[:, 0].astype(np.int64), rows[:, 1].astype(np.int64), rows[:, 2].astype(np.int32),
            rows[:, 3].astype(np.float32))
