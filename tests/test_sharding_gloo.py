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
