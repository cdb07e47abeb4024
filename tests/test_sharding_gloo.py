"""CPU, world_size 2 over gloo: frame sharding and the final label gather."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vilgod_b200 import sharding
    frames = sharding.frames_of_rank(7, rank, world)
    fid, cidx, cls, sc = [], [], [], []
    for f in frames:
        n = 3 + f                       # ragged: different cluster counts per frame
        fid += [f] * n
        cidx += list(range(n))
        cls += [(f + i) % 4 for i in range(n)]
        sc += [0.01 * f + 0.001 * i for i in range(n)]
    res = sharding.gather_labels(torch.tensor(fid), torch.tensor(cidx), torch.tensor(cls, dtype=torch.int32),
                                 torch.tensor(sc, dtype=torch.float32))
    if rank == 0:
        q.put([r.tolist() for r in res])
    else:
        assert res is None
    dist.destroy_process_group()


def test_frames_partition_exactly():
    from vilgod_b200 import sharding
    for world in (1, 2, 4, 8):
        got = sorted(sum((sharding.frames_of_rank(64, r, world) for r in range(world)), []))
        assert got == list(range(64))


def test_gather_labels_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    fid, cidx, cls, sc = res
    exp_f, exp_c = [], []
    for f in range(7):
        exp_f += [f] * (3 + f)
        exp_c += list(range(3 + f))
    assert fid == exp_f and cidx == exp_c
    assert cls == [(f + i) % 4 for f, i in zip(exp_f, exp_c)]
    assert np.allclose(sc, [0.01 * f + 0.001 * i for f, i in zip(exp_f, exp_c)], atol=1e-7)


def _queue_worker(rank, world, port, q):
    import time
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vilgod_b200 import sharding
    fq = sharding.FrameQueue(23, name="seq_test")
    got = []
    while True:
        f = fq.next()
        if f is None:
            break
        got.append(f)
        time.sleep(0.002 if rank == 0 else 0.02)       # rank 1 is the slow board
    fid = torch.tensor(got, dtype=torch.int64)
    res = sharding.gather_labels(fid, torch.zeros_like(fid), torch.zeros(len(got), dtype=torch.int32),
                                 torch.full((len(got),), float(rank)))
    if rank == 0:
        q.put([r.tolist() for r in res])
    dist.barrier()
    dist.destroy_process_group()


def test_dynamic_frame_queue_world2_gloo():
    """Every frame of the sequence is taken exactly once, frames 0 .. W-1 go to their mod-W ranks, and
    the faster rank ends up with more frames (the queue balances by progress, not by a fixed stride)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_queue_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    fid, _, _, owner = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert fid == list(range(23))
    assert owner[0] == 0.0 and owner[1] == 1.0
    assert owner.count(0.0) > owner.count(1.0)


def test_frame_queue_without_process_group_is_a_local_counter():
    from vilgod_b200 import sharding
    fq = sharding.FrameQueue(5)
    assert [fq.next() for _ in range(7)] == [0, 1, 2, 3, 4, None, None]


def test_raw_sequence_generator_is_deterministic():
    from vilgod_b200 import synthetic
    a = synthetic.make_sequence_raw(3, clusters_per_frame=20, seed=5)
    b = synthetic.make_sequence_raw(3, clusters_per_frame=20, seed=5)
    for (p1, o1, t1), (p2, o2, t2) in zip(a, b):
        assert np.array_equal(p1, p2) and np.array_equal(o1, o2) and np.array_equal(t1, t2)
        assert p1.dtype == np.float32 and o1[-1] == len(p1) and t1.shape == (4, 4)
