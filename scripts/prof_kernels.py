"""Small driver for ncu: one projection launch (1 frame) and one encoder pass (1024 images)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vilgod_b200 import synthetic, weights
from vilgod_b200.engine import Engine
V = int(os.environ.get("VIEWS", "10"))
nimg = int(os.environ.get("IMAGES", "1024"))
eng = Engine(num_views=V)
eng.load_vit_weights(weights.random_init_visual_state_dict(1234))
eng.set_text_features(weights.synthetic_text_features(24))
pts, off = synthetic.make_clusters(max(nimg // V, 1) * 3, seed=5)
out = eng.project(pts, off)
torch.cuda.synchronize()
tiles = out["tiles"][:nimg].contiguous()
r = eng.encode_score(tiles)
torch.cuda.synchronize()
print("done", tiles.shape, eng.launch_count)
