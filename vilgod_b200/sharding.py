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


class FrameQueue:
    """Dynamic frame assignment for one sequence: every rank pulls the next unprocessed frame from a
    shared counter instead of owning f = r (mod W) up front, so a slower board (the 1000 W cap settles
    the eight boards of a box at different clocks) simply takes fewer frames and no rank waits for it.
    The counter lives in the process group's key-value store (one atomic add per pull, served by
    rank 0's TCPStore over localhost): no collective, nothing on the GPU.  Without a process group the
    queue degenerates to a local counter.  The first pull of rank r returns frame r, which keeps the
    start of the sequence identical to the static mod-W assignment (zero_shot_detector.py:365
    iterates the frames in order)."""

    def __init__(self, num_frames: int, name: str = "frames", store=None, rank: int = 0, world: int = 1):
        self.n, self.rank, self.world = int(num_frames), rank, world
        self.key = f"vilgod_b200/{name}"
        self.store = store
        self.first = True
        self.local = 0
        if store is None and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            self.store = dist.distributed_c10d._get_default_store()
            self.rank, self.world = dist.get_rank(), dist.get_world_size()

    def next(self):
        """-> next frame index for this rank, or None when the sequence is exhausted."""
        if self.store is None:
            f, self.local = self.local, self.local + 1
            return f if f < self.n else None
        if self.first:                      # frames 0 .. W-1 are pre-assigned, the counter starts at W
            self.first = False
            return self.rank if self.rank < self.n else None
        f = self.store.add(self.key, 1) - 1 + self.world
        return f if f < self.n else None


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
