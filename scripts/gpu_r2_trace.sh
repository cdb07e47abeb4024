#!/bin/bash
# Round 2: phase timeline (clock64) of the fast projection kernel on the Waymo mix.
VG_LIB_PATH=vilgod_b200/lib/libvilgod_b200_trace.so VG_PROJ_TRACE=1 timeout 120 python - <<'PY'
import os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
from vilgod_b200 import synthetic
from vilgod_b200.engine import Engine, _ptr, _stream
V = 10
eng = Engine(num_views=V)
pts, off = synthetic.make_clusters(3000, n_min=10, n_max=2048, seed=3)
C = len(off) - 1
d_p, d_o = torch.from_numpy(pts).cuda(), torch.from_numpy(off).cuda()
tiles = torch.empty((C * V, 196, 256), dtype=eng.op_torch_dtype, device="cuda")
for _ in range(2):
    eng._check(eng.lib.vg_project(eng._h, _ptr(d_p), _ptr(d_o), C, _ptr(tiles), None, None, None, _stream()))
    torch.cuda.synchronize()
eng.close()
PY
