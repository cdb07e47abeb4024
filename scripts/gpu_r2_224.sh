#!/bin/bash
# Round 2: R = 224 fast kernel, 512 vs 1024 threads per CTA (Waymo mix, 10 views) + parity of both.
for v in 1 2; do
VG_PROJ_VARIANT=$v timeout 200 python - <<'PY'
import os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
from vilgod_b200 import synthetic
from vilgod_b200.engine import Engine, _ptr, _stream
V = 10
eng = Engine(num_views=V, resolution=224)
pts, off = synthetic.make_clusters(3000, n_min=10, n_max=2048, seed=3)
C = len(off) - 1
d_p, d_o = torch.from_numpy(pts).cuda(), torch.from_numpy(off).cuda()
tiles = torch.empty((C * V, 196, 256), dtype=eng.op_torch_dtype, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def run(): eng._check(eng.lib.vg_project(eng._h, _ptr(d_p), _ptr(d_o), C, _ptr(tiles), None, None, None, _stream()))
for _ in range(3): run()
ts = []
for _ in range(7):
    flush.zero_(); s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); run(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
ms = float(np.median(ts)); by = 12.0 * int(off[-1]) + C * V * 100352.0
print("R=224 variant %s: %.3f ms, %.4f us/image, %.3f of 6451 GB/s, checksum %d" % (os.environ["VG_PROJ_VARIANT"], ms, 1e3 * ms / (C * V), by / ms / 1e6 / 6451.2, int(tiles.view(torch.int16).long().sum())))
eng.close()
PY
done
