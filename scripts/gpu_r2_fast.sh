#!/bin/bash
# Round 2: fast projection kernel -- parity tests, then the Waymo-mix timing of every variant.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_projection_gpu.py -q -m gpu -x 2>&1 | grep -v Warning | grep -E "^E |passed|failed|Error" | head -30
for v in 0 1 2 3; do
VG_PROJ_VARIANT=$v timeout 120 python - <<'PY'
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
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def run(): eng._check(eng.lib.vg_project(eng._h, _ptr(d_p), _ptr(d_o), C, _ptr(tiles), None, None, None, _stream()))
for _ in range(3): run()
ts = []
for _ in range(7):
    flush.zero_(); s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); run(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
ms = float(np.median(ts)); by = 12.0 * int(off[-1]) + C * V * 100352.0
print("variant %s waymo mix V=%d: %.3f ms, %.4f us/image, %.1f GB/s, %.3f of 6451 GB/s" % (os.environ["VG_PROJ_VARIANT"], V, ms, 1e3 * ms / (C * V), by / ms / 1e6, by / ms / 1e6 / 6451.2))
eng.close()
PY
done
