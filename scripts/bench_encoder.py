"""BASELINE.json configs[4]: CLIP encoder only -- ViT-B/16 on 1k..16k depth images + prompt scoring,
against the measured bf16 tensor-core peak (35.127 GFLOP per image)."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vilgod_b200 import synthetic, weights
from vilgod_b200.engine import Engine
peaks = json.load(open("MEASURED_PEAKS.json")) if os.path.exists("MEASURED_PEAKS.json") else {"bf16_tflops_sustained": 1400.0}
V = 6
eng = Engine(num_views=V)
eng.load_vit_weights(weights.random_init_visual_state_dict(1234))
eng.set_text_features(weights.synthetic_text_features(24))
pts, off = synthetic.make_clusters(128, seed=synthetic.DEFAULT_SEED)          # cfg1's projection, tiled
base = eng.project(pts, off)["tiles"]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for B in (1024, 2048, 4096, 8192, 16384):
    tiles = base.repeat((B + base.shape[0] - 1) // base.shape[0], 1, 1)[:B].contiguous()
    for _ in range(2):
        eng.encode_score(tiles, want_feats=False)
    ts = []
    for _ in range(3):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); eng.encode_score(tiles, want_feats=False); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ms = float(np.median(ts))
    tf = 35.127e9 * B / (ms * 1e-3) / 1e12
    print(json.dumps(dict(images=B, ms=ms, images_per_s=B / (ms * 1e-3), algorithmic_tflops=tf,
                          frac_of_measured_sustained_bf16_peak=tf / peaks["bf16_tflops_sustained"])), flush=True)
