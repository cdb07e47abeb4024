"""BASELINE.json configs[3]: projection-only sweep (fixed N per cluster, 6 / 10 views, R = 112 / 224)
against the measured HBM peak.  Algorithmic bytes per cluster = 12 N + V * 224*224*2.
Usage: python scripts/bench_projection.py [quick]"""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vilgod_b200 import synthetic
from vilgod_b200.engine import Engine

peaks = json.load(open("MEASURED_PEAKS.json")) if os.path.exists("MEASURED_PEAKS.json") else {"hbm_gbs": 6650.0}
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
rows = []
quick = "quick" in sys.argv or "one" in sys.argv
for V, R in ((10, 112),) if quick else ((6, 112), (10, 112), (6, 224), (10, 224)):
    eng = Engine(num_views=V, resolution=R)
    for N in (("waymo",) if "one" in sys.argv else ("waymo", 256, 1024, 4096) if quick else (256, 1024, 4096, 16384, 65536, "waymo")):
        if N == "waymo":
            pts, off = synthetic.make_clusters(3000, n_min=10, n_max=2048, seed=3)
        else:
            C = max(int(1.0e9 / (V * 100352)), 64) if N <= 4096 else 512
            C = min(C, 2000)
            pts, off = synthetic.make_clusters(C, n_min=N, n_max=N, seed=N)
        C = len(off) - 1
        d_p, d_o = torch.from_numpy(pts).cuda(), torch.from_numpy(off).cuda()
        tiles = torch.empty((C * V, 196, 256), dtype=eng.op_torch_dtype, device="cuda")
        import ctypes as Cc
        from vilgod_b200.engine import _ptr, _stream
        def run():
            eng._check(eng.lib.vg_project(eng._h, _ptr(d_p), _ptr(d_o), C, _ptr(tiles), None, None, None, _stream()))
        for _ in range(3):
            run()
        ts = []
        for _ in range(5):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); run(); e.record(); torch.cuda.synchronize()
            ts.append(s.elapsed_time(e))
        ms = float(np.median(ts))
        bytes_alg = 12.0 * int(off[-1]) + C * V * 100352.0
        gbs = bytes_alg / (ms * 1e-3) / 1e9
        rows.append(dict(views=V, resolution=R, points_per_cluster=N, clusters=C, images=C * V, ms=ms,
                         us_per_image=1e3 * ms / (C * V), algorithmic_GBs=gbs,
                         frac_of_measured_hbm_peak=gbs / peaks["hbm_gbs"]))
        print(json.dumps(rows[-1]), flush=True)
    eng.close()
