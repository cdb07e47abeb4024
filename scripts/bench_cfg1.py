"""BASELINE.json configs[0]: one synthetic Waymo-shaped frame (128 clusters <= 2048 points, 6 views,
24 prompts) end to end from host memory -- the case the reference itself was timed on in the build
container (tests/golden/e2e.npz: ref_seconds, 8 host cores)."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vilgod_b200 import weights
from vilgod_b200.engine import Engine
g = np.load("tests/golden/e2e.npz"); t = np.load("tests/golden/tables.npz")
for dt in ("bf16", "f16"):
    eng = Engine(num_views=6, operand_dtype=dt)
    eng.load_vit_weights(weights.random_init_visual_state_dict(1234))
    eng.set_text_features(t["text_features"])
    hp = torch.from_numpy(g["points"]).pin_memory(); ho = torch.from_numpy(g["offsets"]).pin_memory()
    out = eng.alloc_outputs(128)
    def step():
        dp = hp.cuda(non_blocking=True); do = ho.cuda(non_blocking=True)
        eng.classify(dp, do, out=out)
        return out["voted_class"].cpu(), out["voted_score"].cpu(), out["top1"].cpu()
    for _ in range(3): step()
    ts = []
    for _ in range(10):
        torch.cuda.synchronize(); t0 = time.perf_counter(); step(); ts.append(time.perf_counter() - t0)
    ms = 1e3 * float(np.median(ts))
    print(json.dumps(dict(config="cfg1: 1 frame, 128 clusters, 6 views", operand_dtype=dt, ms_per_frame=ms,
                          clusters_per_s=128 / (ms * 1e-3), reference_cpu_seconds=float(g["ref_seconds"]),
                          reference_cpu_cores=int(g["ref_cores"]),
                          speedup_vs_reference_cpu=float(g["ref_seconds"]) / (ms * 1e-3))), flush=True)
    eng.close()
