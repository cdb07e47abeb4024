"""GPU diagnostic: error statistics of the bf16 tower against the reference golden logits."""
import numpy as np, torch, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vilgod_b200 import weights
from vilgod_b200.engine import Engine
g = np.load("tests/golden/e2e.npz"); t = np.load("tests/golden/tables.npz")
e = Engine(num_views=6)
e.load_vit_weights(weights.random_init_visual_state_dict(1234))
e.set_text_features(t["text_features"])
out = e.classify(g["points"], g["offsets"])
res_l = None
C, V = 128, 6
tiles = e.project(g["points"], g["offsets"])["tiles"]
r = e.encode_score(tiles, want_logits=True)
lg = r["logits"].cpu().numpy(); ref = g["logits"]
d = lg - ref
print("logit err raw max %.4f mean %.4f | centred max %.4f mean %.4f" % (np.abs(d).max(), np.abs(d).mean(), np.abs(d - d.mean(1, keepdims=True)).max(), np.abs(d - d.mean(1, keepdims=True)).mean()))
srt = np.sort(ref, 1); m = srt[:, -1] - srt[:, -2]
print("ref margin median %.4f p10 %.4f min %.5f" % (np.median(m), np.percentile(m, 10), m.min()))
top1 = lg.argmax(1); rt = ref.argmax(1)
for thr in (0, 0.02, 0.04, 0.06, 0.1):
    sel = m > thr
    print("margin > %.2f: n=%d agree=%.4f" % (thr, sel.sum(), (top1[sel] == rt[sel]).mean()))
f = r["feats"].cpu().numpy(); fr = g["feats"].astype(np.float32)
print("feat cos min %.6f" % (f * fr).sum(1).min())
pr = torch.from_numpy(ref).softmax(-1).numpy()
print("prob err max %.5f" % np.abs(r["probs"].cpu().numpy() - pr).max())
