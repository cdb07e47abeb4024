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
