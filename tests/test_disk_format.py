"""the label write-back and the on-disk (pickle) format of the path's
outputs.  Runs the reference's own Detection / LidarFrame.update_object_classes when the reference
tree is present (the build container); the host mirror must leave byte-identical state behind."""
import pickle

import numpy as np
import pytest

from oracle import ref_harness as rh
from vilgod_b200 import voting
from vilgod_b200.engine import CLASS_LIST, CLASS_MAPPING

pytestmark = pytest.mark.skipif(not rh.available(), reason="reference tree not mounted")


def _detections(n):
    rh.install_shims()
    from src.dataclass.objects import Detection
    rng = np.random.default_rng(0)
    return [Detection(i, rng.normal(size=(20, 3)).astype(np.float32), np.arange(20)) for i in range(n)]


def _inputs(C, V, seed=3):
    rng = np.random.default_rng(seed)
    idx = rng.integers(0, 24, size=(C, V))
    idx[: C // 2] = rng.integers(0, 24, size=(C // 2, 1))          # unanimous votes too
    detailed = np.asarray(CLASS_LIST)[idx]
    names = np.vectorize(CLASS_MAPPING.get)(detailed)
    scores = rng.uniform(0.05, 0.9, size=(C, V)).astype(np.float32)
    return names, detailed, scores


class _Frame:
    def __init__(self, dets):
        self.detections = dets


@pytest.mark.parametrize("V", [4, 6, 10])
def test_write_back_is_byte_identical_to_the_reference(V):
    rh.install_shims()
    from src.vilgod import lidar_frame
    C = 40
    names, detailed, scores = _inputs(C, V)
    update = [True] * 15 + [False] * 5 + [True] * 25
    ref_dets, our_dets = _detections(C + 5), _detections(C + 5)
    lidar_frame.LidarFrame.update_object_classes(_Frame(ref_dets), names, detailed, scores, update,
                                                 key="clip_a_point_representation_of_a",
                                                 aggregation="voting")
    voting.update_object_classes(our_dets, names, detailed, scores, update,
                                 key="clip_a_point_representation_of_a", aggregation="voting")
    for r, o in zip(ref_dets, our_dets):
        sr, so = r.serialize, o.serialize
        assert sorted(sr) == sorted(so)
        for k in sr:
            if isinstance(sr[k], dict):
                assert sorted(sr[k]) == sorted(so[k])
                for kk in sr[k]:
                    a, b = np.asarray(sr[k][kk]), np.asarray(so[k][kk])
                    assert a.dtype.kind == b.dtype.kind and a.shape == b.shape, (k, a.dtype, b.dtype)
                    assert np.array_equal(a, b), k
        # resume logic keys off these dicts (zero_shot_detector.py:345-363)
        assert (r.object_class is None) == (o.object_class is None)
    # whole-sequence pickle round trip (sync_lidar_frames, zero_shot_detector.py:105-123)
    blob = pickle.dumps([d.serialize for d in our_dets])
    back = pickle.loads(blob)
    fresh = _detections(C + 5)
    for d, data in zip(fresh, back):
        d.sync_detection(data)
    for a, b in zip(fresh, ref_dets):
        if b.object_class is None:
            assert a.object_class is None
            continue
        key = "clip_a_point_representation_of_a"
        assert a.object_class[key] == b.object_class[key]
        assert np.float32(a.object_class_score[key]) == np.float32(b.object_class_score[key])
        assert np.array_equal(a.object_class_predictions_detailed[key], b.object_class_predictions_detailed[key])
        assert a.object_class_predictions_score[key].dtype == np.float32


def test_gpu_style_vote_hand_off_keeps_the_format():
    """classify_frame hands (voted names, voted scores) from the GPU; the stored entries must have
    the same types as the host vote's."""
    C, V = 12, 6
    names, detailed, scores = _inputs(C, V, seed=9)
    mapped = sorted(set(CLASS_MAPPING.values()))
    ids = np.vectorize(mapped.index)(names)
    vid, vs = voting.vote(ids, scores, len(mapped))
    dets = _detections(C)
    voting.update_object_classes(dets, names, detailed, scores, [True] * C, key="k",
                                 voted=(np.asarray(mapped)[vid], vs))
    ref = _detections(C)
    rh.install_shims()
    from src.vilgod import lidar_frame
    lidar_frame.LidarFrame.update_object_classes(_Frame(ref), names, detailed, scores, [True] * C, key="k")
    for a, b in zip(dets, ref):
        assert a.object_class["k"] == b.object_class["k"]
        assert np.float32(a.object_class_score["k"]) == np.float32(b.object_class_score["k"])
