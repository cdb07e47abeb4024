"""CPU: host-side mirrors of the reference interface against the golden vectors."""
import numpy as np
import pytest
import torch

from vilgod_b200 import canonicalise, synthetic, views, voting, weights


def test_view_tables_bit_exact(golden):
    g = golden["tables"]
    for V in (4, 6, 10):
        assert np.array_equal(views.view_rot_mats(V).numpy(), g[f"rot{V}"])
    assert np.array_equal(views.gaussian_weights().numpy(), g["gauss"])


def test_canonicalise_matches_reference_bitwise(golden):
    g = golden["canon"]
    raw, off, T = g["raw"], g["offsets"], g["transform"]
    packed = canonicalise.canonicalise_packed(raw, off, T)
    assert np.array_equal(packed, g["canon_f32"])
    for c in range(len(off) - 1):
        one = canonicalise.canonicalise_cluster(raw[off[c]:off[c + 1]], T)
        assert np.array_equal(one, g["canon_f32"][off[c]:off[c + 1]])


@pytest.mark.parametrize("V", [4, 6, 10])
def test_host_vote_matches_reference(golden, V):
    from vilgod_b200.engine import CLASS_LIST, CLASS_MAPPING
    g = golden["vote"]
    mapped_names = sorted(set(CLASS_MAPPING.values()))
    ids = np.asarray([mapped_names.index(CLASS_MAPPING[c]) for c in CLASS_LIST])[g[f"idx{V}"]]
    vid, vs = voting.vote(ids, g[f"scores{V}"], len(mapped_names))
    assert np.array_equal(np.asarray(mapped_names)[vid], g[f"voted_name{V}"])
    assert np.array_equal(vs, g[f"voted_score{V}"])


def test_update_object_classes_writes_reference_entries(golden):
    g = golden["vote"]
    from vilgod_b200.engine import CLASS_LIST, CLASS_MAPPING

    class Det:
        def __init__(self):
            self.store = {}
            self.depth_image = None

        def add_object_entry(self, entry, key, data):
            self.store.setdefault(entry, {})[key] = data

    idx, scores = g["idx4"][:20], g["scores4"][:20]
    detailed = np.asarray(CLASS_LIST)[idx]
    names = np.vectorize(CLASS_MAPPING.get)(detailed)
    dets = [Det() for _ in range(25)]
    upd = [True] * 10 + [False] * 5 + [True] * 10
    voting.update_object_classes(dets, names, detailed, scores, upd, key="k")
    j = 0
    for d, u in zip(dets, upd):
        if not u:
            assert d.store == {}
            continue
        assert d.store["object_class"]["k"] == g["voted_name4"][j]
        assert d.store["object_class_score"]["k"] == g["voted_score4"][j]
        assert np.array_equal(d.store["object_class_predictions_detailed"]["k"], detailed[j])
        assert d.store["object_class_predictions_score"]["k"].dtype == np.float32
        j += 1


def test_product_weights_equal_reference_checkpoint(golden):
    from oracle.make_golden import weights_fingerprint
    g = golden["vit"]
    sha, sums = weights_fingerprint(weights.random_init_visual_state_dict(1234))
    assert sha == str(g["plain_weights_sha256"])
    sha, _ = weights_fingerprint(weights.perturb_layernorms(weights.random_init_visual_state_dict(1234)))
    assert sha == str(g["ln_weights_sha256"])


def test_synthetic_generator_is_deterministic_and_shaped(golden):
    p1, o1 = synthetic.make_clusters(128, seed=synthetic.DEFAULT_SEED)
    g = golden["e2e"]
    assert np.array_equal(o1, g["offsets"]) and np.array_equal(p1, g["points"])
    n = np.diff(o1)
    assert n.min() >= 10 and n.max() <= 2048
    frames = synthetic.make_frames(3, clusters_per_frame=50, seed=1)
    pts, off, bounds = synthetic.concat_frames(frames)
    assert off[-1] == len(pts) and bounds[-1] == len(off) - 1
    assert np.array_equal(pts[off[bounds[1]]:off[bounds[1] + 1]],
                          frames[1][0][frames[1][1][0]:frames[1][1][1]])


def test_canonicalise_passes_empty_clusters_through():
    """A cluster without points contributes nothing to the packed batch (the projection then
    reports VG_EDEGENERATE for it); the reference guards only empty FRAMES
    (zero_shot_detector.py:403)."""
    raw, off, _ = synthetic.make_clusters_raw(5, n_min=10, n_max=50, seed=2)
    ref = canonicalise.canonicalise_packed(raw, off)
    off2 = np.array([0, 0, off[1], off[2], off[2], off[3], off[4], off[5], off[5]])
    assert np.array_equal(canonicalise.canonicalise_packed(raw, off2), ref)
    assert canonicalise.canonicalise_packed(np.zeros((0, 3)), np.array([0, 0])).shape == (0, 3)
    with pytest.raises(ValueError):
        canonicalise.canonicalise_packed(raw, np.array([0, 5, 3]))


def test_classify_frame_on_a_frame_without_clusters():
    """No engine call is made for an empty frame, so this runs without a GPU."""
    from vilgod_b200.reference_api import classify_frame

    class _NoEngine:
        num_views = 6

    res = classify_frame(_NoEngine(), [], np.eye(4))
    assert res["class_names"].shape == (0, 6) and res["class_scores"].dtype == np.float32
    assert res["voted_names"].shape == (0,) and res["depth_images"] == []


def test_top_k_labels_follow_the_reference_tail():
    """Same lists as clip_utils.py:49-63 builds from a probability matrix, for k = 1 and k = 3."""
    from vilgod_b200.engine import CLASS_LIST
    from vilgod_b200.reference_api import top_k_labels
    rng = np.random.default_rng(4)
    logits = rng.normal(size=(7, 24)).astype(np.float32)
    probs = torch.from_numpy(logits).softmax(dim=-1).numpy()
    ids = dict(enumerate(CLASS_LIST))
    for k in (1, 3):
        names, scores = top_k_labels(probs, k, ids)
        assert len(names) == len(scores) == 7 * k
        for i in range(7):
            order = np.argsort(-probs[i])[:k]
            assert names[i * k:(i + 1) * k] == [CLASS_LIST[j] for j in order]
            assert np.array_equal(np.asarray(scores[i * k:(i + 1) * k]), probs[i][order])
            assert scores[i * k].dtype == np.float32


def test_prompt_encoding_through_the_callers_clip_package(golden, tmp_path):
    """ClipWrapper(clip_cfg, model_path, device) builds text_features itself, through the caller's
    `clip` (clip_utils.py:19-26).  With the reference tree mounted, the helper must reproduce the
    reference wrapper's text features and hand over the visual weights the reference model holds."""
    from oracle import ref_harness as rh
    if not rh.available():
        pytest.skip("reference tree not mounted")
    rh.install_shims()
    from vilgod_b200.reference_api import encode_prompts_with_clip
    rh.make_random_checkpoint(str(tmp_path / "ViT-B-16.pt"), seed=1234)
    orig = torch.jit.load

    def _no_jit(*a, **k):      # clip.load's JIT attempt consumes the file handle on torch 2.11
        raise RuntimeError("not a JIT archive (shim)")

    torch.jit.load = _no_jit
    try:
        model, preprocess, tok, tf = encode_prompts_with_clip(rh.clip_cfg(), str(tmp_path), "cpu")
    finally:
        torch.jit.load = orig
    g = golden["tables"]
    assert np.array_equal(tok[0].numpy(), g["text_tokens0"])
    assert np.array_equal(tf.detach().float().numpy(), g["text_features"])
    from oracle.make_golden import weights_fingerprint
    sha, _ = weights_fingerprint({k: v.float() for k, v in model.visual.state_dict().items()})
    assert sha == str(golden["vit"]["plain_weights_sha256"])


def test_fast_projection_kernel_closed_forms():
    """The fast projection kernel (csrc/projection.cu: projection_fast_kernel) derives the output rows /
    8-pixel column groups that can differ from background in closed form instead of scanning the bilinear
    table: it relies on (1) the align_corners source index of torch's area_pixel_compute_source_index,
    computed in fp32 as floor(fl(fl((Q-1)/(S-1)) * o)), being equal to the integer floor(o (Q-1) / (S-1)) for
    every output index (projection_init re-checks this on the device table and falls back to the general
    kernel otherwise), and (2) first_with(t) = ceil(t (S-1) / (Q-1)) being the first output index whose
    source index reaches t.  Both are checked here exhaustively, for R = 112 and R = 224, against the
    definition the general kernel uses (a row is active when one of its two source rows lies in [lo, hi])."""
    f = np.float32
    S, NG = 224, 28
    for R in (112, 224):
        Q = R - 2
        scale = f(f(Q - 1) / f(S - 1))
        i0 = np.array([min(int(np.floor(f(scale * f(o)))), Q - 1) for o in range(S)])
        assert np.array_equal(i0, (np.arange(S) * (Q - 1)) // (S - 1))

        def first_with(t):
            return 0 if t <= 0 else (t * (S - 1) + (Q - 2)) // (Q - 1)

        y1 = i0 + (i0 < Q - 1)
        g = np.arange(NG)
        xa, xb = i0[8 * g], i0[8 * g + 7] + (i0[8 * g + 7] < Q - 1)
        for lo in range(Q):
            for hi in range(lo, Q, 3 if R == 224 else 1):
                rows = np.flatnonzero(~((y1 < lo) | (i0 > hi)))
                assert (rows.min(), rows.max()) == (first_with(lo - 1), min(first_with(hi + 1) - 1, S - 1))
                cols = np.flatnonzero(~((xb < lo) | (xa > hi)))
                assert (cols.min(), cols.max()) == (first_with(lo - 1) >> 3,
                                                    min(min(first_with(hi + 1) - 1, S - 1) >> 3, NG - 1))


def test_fast_projection_kernel_emit_source_window():
    """The fast kernel writes the normalised image back into its bounding-box buffer (buffer row = y + 1 - ulo,
    float column = x + 4 (2 - (vlo >> 2))) and fills with 1.0 only what the bilinear emit can read beside it:
    rows 0 and nr + 1, nr + 2, and the strips 0, 1, ns + 2, ns + 3 of the image rows.
    The emit reads source rows i0[oy], i0[oy] + 1 of the active output rows and source columns i0[ox],
    i0[ox] + 1 of the active 8-pixel groups (unclamped at the last row / column, where the second weight is
    exactly 0): checked exhaustively that this window is rows 0 .. nr + 1 x strips 0 .. ns + 3."""
    S, NG = 224, 28
    for R in (112, 224):
        Q = R - 2
        i0 = (np.arange(S) * (Q - 1)) // (S - 1)

        def first_with(t):
            return 0 if t <= 0 else (t * (S - 1) + (Q - 2)) // (Q - 1)

        lo_seen, hi_seen = 99, -1
        for lo in range(Q):
            for hi in range(lo, Q):
                # columns: lo, hi = vlo, vhi
                ns = (hi >> 2) - (lo >> 2) + 1
                g_lo, g_hi = first_with(lo - 1) >> 3, min(min(first_with(hi + 1) - 1, S - 1) >> 3, NG - 1)
                ocol = 4 * (2 - (lo >> 2))
                c_first, c_last = i0[8 * g_lo] + ocol, i0[8 * g_hi + 7] + 1 + ocol
                assert 0 <= c_first and (c_last >> 2) <= ns + 3
                lo_seen, hi_seen = min(lo_seen, c_first >> 2), max(hi_seen, (c_last >> 2) - ns)
                # rows: lo, hi = ulo, uhi
                nr = hi - lo + 1
                oy_lo, oy_hi = first_with(lo - 1), min(first_with(hi + 1) - 1, S - 1)
                r_first, r_last = i0[oy_lo] + 1 - lo, i0[oy_hi] + 1 + 1 - lo
                assert 0 <= r_first and r_last <= nr + 1
        assert (lo_seen, hi_seen) == (0, 3)       # the four margin strips are all needed
        # the four output columns of an emit thread start at source columns s + (0, d1, d2, d3): the patterns
        # the run-load form of the horizontal interpolation selects from
        pats = {tuple(int(i0[o + j] - i0[o]) for j in range(1, 4)) for o in range(0, S, 4)}
        assert pats == ({(0, 0, 1), (0, 1, 1), (1, 1, 1), (1, 1, 2)} if R == 112 else {(0, 1, 2), (1, 2, 3)})


def test_fast_projection_kernel_region_bound():
    """projection_init sizes the fast kernel's shared-memory layout from obj_ratio: at the reference's 0.8
    every occupied grid cell has X, Y in [ceil(R/2 (1-0.8)), ceil(R/2 (1+0.8))] = [12, 101] (R = 112) /
    [23, 202] (R = 224), so the touched region (4 cells below, 2 above the occupied range) never exceeds
    96 rows x 24 float4 strips / 186 x 48 and no image of a valid cluster is handed over for its geometry.
    Checked on the oracle's points2grid over random, anisotropic and adversarial clusters and all ten views."""
    from oracle import projection as op
    from vilgod_b200 import synthetic
    rng = np.random.default_rng(5)
    clusters = []
    pts, off = synthetic.make_clusters(60, n_min=2, n_max=600, rng=rng)
    clusters += [pts[off[c]:off[c + 1]] for c in range(len(off) - 1)]
    for scale in ((1, 1, 1), (50, 0.01, 0.01), (0.01, 50, 1), (1e-3, 1e-3, 1e3), (1e4, 1e4, 1e4)):
        clusters.append((rng.uniform(-1, 1, size=(300, 3)) * np.asarray(scale)).astype(np.float32))
    cube = np.stack(np.meshgrid(*[np.array([-1.0, 1.0])] * 3, indexing="ij"), -1).reshape(-1, 3)   # the corners
    clusters.append(cube.astype(np.float32))
    clusters.append((cube * np.array([3.0, 1.0, 0.2]) + 1e5).astype(np.float32))
    rot = op.view_rot_mats(10)
    for R, (lo, hi), (max_rows, max_strips) in ((112, (12, 101), (96, 24)), (224, (23, 202), (186, 48))):
        Q = R - 2
        for p in clusters:
            for v in range(10):
                q = op.rotate(p, rot[v], fused=9 * len(p) >= 400)
                try:
                    grid = op.points2grid(q, R=R)
                except ValueError:
                    continue                                  # degenerate (no extent): flagged, not projected
                _, ys, xs = np.nonzero(grid)
                assert lo <= ys.min() and ys.max() <= hi and lo <= xs.min() and xs.max() <= hi
                ulo, uhi = max(ys.min() - 4, 0), min(ys.max() + 2, Q - 1)
                vlo, vhi = max(xs.min() - 4, 0), min(xs.max() + 2, Q - 1)
                assert uhi - ulo + 1 <= max_rows and (vhi >> 2) - (vlo >> 2) + 1 <= max_strips
