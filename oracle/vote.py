"""TEST INFRASTRUCTURE ONLY -- restatement of the per-cluster view vote.

Follows src/vilgod/lidar_frame.py:269-285 (LidarFrame.update_object_classes, aggregation='voting')
and the 24 -> 4 class mapping at src/vilgod/zero_shot_detector.py:412-415 /
tools/configs/preprocessor/waymo.yaml:116-140.  Pinned by tests/golden/vote.npz.
"""
from __future__ import annotations

import numpy as np

CLASS_LIST = ['car', 'truck', 'bus', 'van', 'minivan', 'pickup truck', 'school bus', 'fire truck',
              'ambulance', 'pedestrian', 'human body', 'human', 'cyclist', 'rider', 'bicycle',
              'bike', 'traffic light', 'traffic sign', 'fence', 'pole', 'clutter', 'tree', 'house',
              'wall']
MAPPED = ['Vehicle'] * 9 + ['Pedestrian'] * 3 + ['Cyclist'] * 4 + ['Background'] * 8


def vote_one(names, scores):
    """names: [V] str (mapped class per view), scores: [V] f32 -> (name, score).

    np.unique sorts alphabetically; a unique maximum count wins with the mean score of its views;
    on a tie every name competes by mean score with strict '>' (first alphabetically on equality),
    starting from max_score = 0."""
    names = np.asarray(names)
    scores = np.asarray(scores)
    uniq, counts = np.unique(names, return_counts=True)
    if np.sum(counts[np.argmax(counts)] == counts) > 1:
        best, best_score = None, 0
        for n in uniq:
            s = np.mean(scores[names == n])
            if s > best_score:
                best, best_score = n, s
        return best, best_score
    n = uniq[np.argmax(counts)]
    return n, np.mean(scores[names == n])


def vote(names_cv, scores_cv):
    out = [vote_one(n, s) for n, s in zip(names_cv, scores_cv)]
    return np.array([o[0] for o in out]), np.array([o[1] for o in out], dtype=np.float32)
