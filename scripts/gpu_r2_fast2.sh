#!/bin/bash
# Round 2 (third session): emit-table variants of the fast projection kernel -- parity tests, then the
# Waymo-mix timing of variants 1 .. 4 at R = 112 and 1 / 4 at R = 224.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_projection_gpu.py -q -m gpu -x 2>&1 | grep -v Warning | grep -E "^E |passed|failed|Error" | head -30
timeout 300 python - <<'PY' 2>&1 | tee gpurun_out/r02_fast2_variants.txt
import os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
from vilgod_b200 import synthetic
from vilgod_b200.engine import Engine, _ptr, _stream
V = 10
pts, off = synthetic.make_clusters(3000, n_min=10, n_max=2048, seed=3)
C = len(off) - 1
d_p, d_o = torch.from_numpy(pts).cuda(), torch.from_numpy(off).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
by = 12.0 * int(off[-1]) + C * V * 100352.0
for R, variants, rounds in ((112, "131313", 7), (224, "1313", 5)):
    for v in variants:
        os.environ["VG_PROJ_VARIANT"] = v
        eng = Engine(num_views=V, resolution=R)
        tiles = torch.empty((C * V, 196, 256), dtype=eng.op_torch_dtype, device="cuda")
        def run(): eng._check(eng.lib.vg_project(eng._h, _ptr(d_p), _ptr(d_o), C, _ptr(tiles), None, None, None, _stream()))
        for _ in range(3): run()
        ts = []
        for _ in range(rounds):
            flush.zero_(); s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); run(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
        ms = float(np.median(ts))
        print("R=%d variant %s waymo mix V=%d: %.3f ms (min %.3f), %.4f us/image, %.1f GB/s, %.3f of 6451 GB/s, checksum %d" % (
            R, v, V, ms, min(ts), 1e3 * ms / (C * V), by / ms / 1e6, by / ms / 1e6 / 6451.2, int(tiles.view(torch.int16).long().sum())))
        eng.close(); del tiles
PY
