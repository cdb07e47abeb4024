"""GPU end to end: BASELINE.json configs[0] (one synthetic Waymo-shaped frame, 128 clusters, 6
views) through vg_classify against the golden run of the UNMODIFIED reference on the same inputs
and random-init weights, plus a well-separated-prompt variant where top-1 must agree >= 99.5 %."""
import numpy as np
import pytest
import torch

from oracle import pipeline as opipe
from oracle import vit as ovit
from oracle import vote as ovote

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=["bf16", "f16"])
def engine6(request):
    from vilgod_b200 import weights
    from vilgod_b200.engine import Engine
    e = Engine(num_views=6, operand_dtype=request.param)
    e.load_vit_weights(weights.random_init_visual_state_dict(1234))
    yield e
    e.close()


def test_cfg1_frame_against_reference_golden(golden, engine6):
    g, t = golden["e2e"], golden["tables"]
    e = engine6
    e.set_text_features(t["text_features"])
    out = e.classify(g["points"], g["offsets"])
    torch.cuda.synchronize()
    assert int(out["status"].abs().sum()) == 0
    C, V = 128, 6
    probs = out["probs"].cpu().numpy().reshape(C * V, 24)
    ref_logits = g["logits"]
    ref_probs = torch.from_numpy(ref_logits).softmax(dim=-1).numpy()
    # stated bf16 tolerance: probabilities within CLIP's own test tolerance (atol 0.01)
    assert np.abs(probs - ref_probs).max() <= 0.01
    feats = out["feats"].cpu().numpy().reshape(C * V, 512)
    cos = (feats * g["feats"].astype(np.float32)).sum(axis=1)
    assert cos.min() >= 0.9995
    # top-1: raw agreement is reported, margin-aware agreement is asserted.  With a random-init
    # text tower all 24 prompts collapse (median top1-top2 margin ~0.04 logit units), so only
    # images whose reference margin exceeds the stated logit tolerance are decidable.
    top1 = out["top1"].cpu().numpy().reshape(-1)
    ref_top1 = ref_logits.argmax(axis=1)
    srt = np.sort(ref_logits, axis=1)
    margin = srt[:, -1] - srt[:, -2]
    clear = margin > 0.06
    raw = (top1 == ref_top1).mean()
    aware = (top1[clear] == ref_top1[clear]).mean() if clear.any() else 1.0
    print(f"cfg1 [{e.operand_dtype}] top-1 agreement raw {raw:.4f}, margin-aware {aware:.4f} on "
          f"{clear.sum()} images; max |dprob| {np.abs(probs - ref_probs).max():.5f}")
    assert aware >= 0.995
    assert raw >= (0.80 if e.operand_dtype == "bf16" else 0.99)


def test_cfg1_frame_well_separated_prompts(golden, engine6):
    """Same frame and weights, prompt embeddings replaced by well-separated unit vectors (the
    prompt embeddings are an input of the path): margins >> bf16 noise, so >= 99.5 % top-1 and the
    4-class voted labels must match the fp32 oracle."""
    from vilgod_b200 import weights
    g = golden["e2e"]
    e = engine6
    text = weights.synthetic_text_features(24)
    e.set_text_features(text)
    C, V = 48, 6
    off = g["offsets"][:C + 1]
    pts = g["points"][:off[-1]]
    out = e.classify(pts, off)
    ref = opipe.classify(pts, off, V, ovit.make_visual_weights(1234), text.numpy())
    top1 = out["top1"].cpu().numpy().reshape(-1)
    ref_top1 = ref["top1"].reshape(-1)
    srt = np.sort(ref["logits"].reshape(-1, 24), axis=1)
    margin = srt[:, -1] - srt[:, -2]
    # random-init image embeddings are nearly input independent (mean cosine 0.97 between images),
    # so even with separated prompts some images sit on a decision boundary; an image is decidable
    # when the oracle's own top1-top2 margin exceeds twice the stated centred logit tolerance.
    clear = margin > 0.12
    raw = (top1 == ref_top1).mean()
    aware = (top1[clear] == ref_top1[clear]).mean()
    print(f"separated prompts [{e.operand_dtype}]: top-1 agreement raw {raw:.4f}, margin-aware {aware:.4f} on "
          f"{clear.sum()}/{len(clear)} images, median margin {np.median(margin):.3f}")
    assert clear.mean() > 0.5 and aware >= 0.995
    assert raw >= 0.90
    assert np.abs(out["probs"].cpu().numpy() - ref["probs"]).max() <= 0.02
    names = np.asarray(e.mapped_names)[out["voted_class"].cpu().numpy()]
    # the 4-class voted label the loop consumes: exact wherever every view of the cluster is
    # decidable, reported over all clusters
    cl = clear.reshape(C, V).all(axis=1)
    same = names == ref["voted_name"]
    print(f"voted labels: raw agreement {same.mean():.4f}, {cl.sum()} fully decidable clusters")
    assert same[cl].all()
    assert same.mean() >= 0.85
    # propagate_labels thresholds the voted score at 0.5 / 0.35 / 0.3
    # (zero_shot_detector.py:775-795) -- the only place small probability differences can change
    # pseudo-labels.  Same decision wherever the labels agree and the oracle score is not within the
    # stated probability tolerance of the threshold.
    gs, os_ = out["voted_score"].cpu().numpy(), ref["voted_score"]
    for thr in (0.5, 0.35, 0.3):
        decid = same & (np.abs(os_ - thr) > 0.02)
        assert ((gs >= thr) == (os_ >= thr))[decid].all(), thr
    assert np.abs(out["voted_score"].cpu().numpy()[same] - ref["voted_score"][same]).max() <= 0.02


def test_classify_frame_mirror_and_chunking(golden, engine6):
    """classify_frame (host glue of classification()) on raw clusters, and vg_classify's internal
    chunking: a small workspace must give the same bits as one big chunk."""
    from vilgod_b200 import synthetic
    from vilgod_b200.reference_api import classify_frame
    e = engine6
    e.set_text_features(golden["tables"]["text_features"])
    raw, off, _ = synthetic.make_clusters_raw(20, n_min=10, n_max=500, seed=9)
    clusters = [raw[off[c]:off[c + 1]] for c in range(20)]
    res = classify_frame(e, clusters, transform_to_ego=np.eye(4))
    assert res["class_names"].shape == (20, 6) and res["class_scores"].dtype == np.float32
    assert set(res["voted_names"]) <= set(e.mapped_names)
    from vilgod_b200 import canonicalise
    packed = canonicalise.canonicalise_packed(raw, off, np.eye(4))
    full = e.classify(packed, off)
    probs_full = full["probs"].clone()
    import ctypes as C
    need = int(e.lib.vg_workspace_bytes(e._h, 12))          # room for 2 clusters x 6 views
    e._ws = torch.empty(need, dtype=torch.uint8, device="cuda")
    ws_small = e._ws
    orig = e.workspace
    e.workspace = lambda n: ws_small
    try:
        small = e.classify(packed, off)
    finally:
        e.workspace = orig
        e._ws = None
    assert torch.equal(small["probs"], probs_full)
    assert torch.equal(small["voted_class"], full["voted_class"])


def test_classify_is_cuda_graph_capturable(golden, engine6):
    """no hidden allocation or synchronisation after the weights are loaded,
    so one frame of vg_classify can be captured once and replayed on new points (same shapes)."""
    from vilgod_b200 import canonicalise, synthetic
    e = engine6
    e.set_text_features(golden["tables"]["text_features"])
    raw, off, _ = synthetic.make_clusters_raw(16, n_min=10, n_max=400, seed=21)
    packed = canonicalise.canonicalise_packed(raw, off, np.eye(4))
    p_static = torch.as_tensor(packed, dtype=torch.float32).cuda()
    o_static = torch.as_tensor(off, dtype=torch.int32).cuda()
    eager = e.classify(p_static, o_static)
    torch.cuda.synchronize()
    eager = {k: v.clone() for k, v in eager.items() if v is not None}
    out = e.alloc_outputs(16)
    e.workspace(16 * 6)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        e.classify(p_static, o_static, out=out)          # warm-up on the capture stream
    torch.cuda.current_stream().wait_stream(side)
    launches0 = e.launch_count
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        e.classify(p_static, o_static, out=out)
    captured = e.launch_count - launches0
    assert captured > 0
    for k in out:
        if out[k] is not None:
            out[k].zero_()
    graph.replay()
    torch.cuda.synchronize()
    assert e.launch_count - launches0 == captured          # replay goes through no library call
    for k, v in eager.items():
        assert torch.equal(out[k], v), k
    # replay on new points of the same packed shape: a permutation of the points inside each cluster
    # leaves every result unchanged (scatter-max is order independent) but exercises fresh inputs
    perm = np.concatenate([off[c] + np.random.default_rng(c).permutation(off[c + 1] - off[c])
                           for c in range(16)])
    p_static.copy_(torch.as_tensor(packed[perm], dtype=torch.float32))
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(out["top1"], eager["top1"])
    assert torch.equal(out["probs"], eager["probs"])
