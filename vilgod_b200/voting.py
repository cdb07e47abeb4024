"""Host-side view vote and label write-back.

``vote`` is a vectorised numpy mirror of the per-detection loop in
``LidarFrame.update_object_classes`` (reference ``src/vilgod/lidar_frame.py:260-291``);
``update_object_classes`` writes the same five dictionary entries the reference stores on each
``Detection`` so the pickle format (``src/dataclass/objects.py:87-103``) round-trips.
"""
from __future__ import annotations

import numpy as np


def vote(mapped_ids: np.ndarray, scores: np.ndarray, num_classes: int):
    """mapped_ids [C,V] int (class ids in alphabetical order of the class names), scores [C,V] f32
    -> (voted id [C], voted score [C] f32).  Majority vote; on a tie of the top count every class
    present competes by mean score (strict '>', alphabetical order, starting from 0)."""
    mapped_ids = np.asarray(mapped_ids)
    scores = np.asarray(scores, dtype=np.float32)
    C, V = mapped_ids.shape
    voted = np.full(C, -1, dtype=np.int32)
    vscore = np.zeros(C, dtype=np.float32)
    onehot = mapped_ids[:, :, None] == np.arange(num_classes)[None, None, :]
    counts = onehot.sum(axis=1)
    top = counts.max(axis=1)
    tie = (counts == top[:, None]).sum(axis=1) > 1
    for c in range(C):
        present = np.flatnonzero(counts[c])
        means = {k: np.mean(scores[c][mapped_ids[c] == k]) for k in present}
        if tie[c]:
            best, best_s = -1, np.float32(0)
            for k in present:
                if means[k] > best_s:
                    best, best_s = k, means[k]
            voted[c], vscore[c] = best, best_s
        else:
            k = int(np.argmax(counts[c]))
            voted[c], vscore[c] = k, means[k]
    return voted, vscore


def update_object_classes(detections, class_names, class_names_detailed, class_scores,
                          cluster_update_list, key="class_key", aggregation="voting",
                          depth_images=None, voted=None):
    """Same signature and side effects as LidarFrame.update_object_classes, operating on any
    objects with ``add_object_entry``.  ``voted`` = (names [n], scores [n]) lets the caller pass the
    GPU vote instead of recomputing it on the host."""
    if aggregation != "voting":
        raise NotImplementedError
    idx = 0
    if voted is None:
        uniq = sorted(set(np.asarray(class_names).reshape(-1).tolist()))
        ids = np.vectorize(uniq.index)(np.asarray(class_names))
        vid, vsc = vote(ids, np.asarray(class_scores), len(uniq))
        voted = (np.asarray(uniq, dtype=object)[vid], vsc)
    for d_idx, det in enumerate(detections):
        if not cluster_update_list[d_idx]:
            continue
        det.add_object_entry("object_class_predictions", key, class_names[idx])
        det.add_object_entry("object_class_predictions_detailed", key, class_names_detailed[idx])
        det.add_object_entry("object_class_predictions_score", key, class_scores[idx])
        det.add_object_entry("object_class", key, voted[0][idx])
        det.add_object_entry("object_class_score", key, voted[1][idx])
        if depth_images is not None:
            det.depth_image = depth_images[idx]
        idx += 1
