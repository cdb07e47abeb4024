"""Drop-in mirrors of the two reference operators on the hot path, plus the fused frame call.

Same names, argument meaning and error behaviour as the reference:

  RealisticProjection(cfg).get_img(points)          reference src/utils/mv_utils.py:130-201
  ClipWrapper(clip_cfg, model_path, device)         reference src/utils/clip_utils.py:10-63
      .predict_clip_labels(list_of_PIL_images)
  classify_frame(...) / classify_lidar_frame(...)   loop body of ZeroShotDetector.classification,
                                                    reference src/vilgod/zero_shot_detector.py:389-417

Everything numerical on the hot path runs in libvilgod_b200.so on the GPU.  The CLIP *text* tower
and tokenizer run once at start-up in the reference (clip_utils.py:23-26) and are not part of the
hot path: ``ClipWrapper`` runs them once through the caller's own ``clip`` package exactly like the
reference does (or takes the cached ``text_features [P,512]`` as an argument).
"""
from __future__ import annotations

from pathlib import Path
from typing import Optional

import numpy as np
import torch

from . import canonicalise, views
from .engine import CLASS_LIST, CLASS_MAPPING, Engine, u8_to_tiles


def _cfg_get(cfg, name, default=None):
    if isinstance(cfg, dict):
        return cfg.get(name, default)
    return getattr(cfg, name, default)


class RealisticProjection:
    """For creating images from PC based on the view information (GPU, fused)."""

    def __init__(self, lidar_image_projection_cfg, num_views: int = 4, engine: Optional[Engine] = None):
        cfg = lidar_image_projection_cfg
        self.resolution = _cfg_get(cfg, "resolution", 112)
        self.depth = _cfg_get(cfg, "depth", 8)
        self.obj_ratio = _cfg_get(cfg, "obj_ratio", 0.8)
        self.depth_bias = _cfg_get(cfg, "depth_bias", 0.2)
        self.num_views = num_views
        self.rot_mat = views.view_rot_mats(num_views)
        self.translation = torch.tensor([[-0.5, -0.5, 0.0]] * num_views).float().unsqueeze(1)
        gk = _cfg_get(cfg, "gaussian_kernel", None)
        sigma = _cfg_get(gk, "sigma", 3) if gk is not None else 3
        zsigma = _cfg_get(gk, "zsigma", 1) if gk is not None else 1
        self.grid2image = None   # fused into the kernel; kept as an attribute for compatibility
        self.engine = engine or Engine(num_views=num_views, rot_mat=self.rot_mat,
                                       resolution=self.resolution, depth=self.depth,
                                       obj_ratio=self.obj_ratio, depth_bias=self.depth_bias,
                                       gauss=views.gaussian_weights(3, sigma, zsigma))
        self.rot_mat = self.rot_mat.to(self.engine.device)
        self.translation = self.translation.to(self.engine.device)

    def get_img(self, points: torch.Tensor) -> torch.Tensor:
        """points [b, N, 3] fp32 (cuda) -> [b * V, 3, R-2, R-2] fp32, view index fastest, in the
        reference's orientation (first image axis = x, mv_utils.py:125)."""
        if points.ndim != 3 or points.shape[-1] != 3:
            raise ValueError("points must be [b, N, 3]")
        b, n, _ = points.shape
        offsets = torch.arange(b + 1, dtype=torch.int32) * n
        out = self.engine.project(points.reshape(b * n, 3), offsets, want_tiles=False,
                                  want_densified=True)
        if int((out["status"] != 0).sum()) != 0:
            raise ValueError("degenerate cluster (no extent): the reference yields NaN here")
        img = out["densified"].transpose(1, 2)
        return img[:, None].repeat(1, 3, 1, 1)


def encode_prompts_with_clip(clip_cfg, model_path, device):
    """What ``ClipWrapper.__init__`` of the reference does once per process (clip_utils.py:19-26),
    with the CALLER's own ``clip`` package (third_party/CLIP of the reference tree): load the
    checkpoint, tokenize the prompt ensemble, run the text tower, L2-normalise.  Returns
    (model, preprocess, text_tokenized, text_features).  The text tower is not part of the hot path
    and is not re-implemented here."""
    import clip   # the reference's dependency; this package never ships its own copy
    model, preprocess = clip.load(Path(model_path) / _cfg_get(clip_cfg, "model_name"), device=device)
    class_list = list(_cfg_get(clip_cfg, "class_list", CLASS_LIST))
    template = _cfg_get(clip_cfg, "prompt_template", "a point representation of a {}")
    text_prompt_list = [template.format(x) for x in class_list]
    text_tokenized = clip.tokenize(text_prompt_list).to(device)
    with torch.no_grad():
        text_features = model.encode_text(text_tokenized)
        text_features /= text_features.norm(dim=-1, keepdim=True)
    return model, preprocess, text_tokenized, text_features


def top_k_labels(probs, top_k, id_to_class_dict):
    """The host tail of predict_clip_labels (clip_utils.py:49-63): per image argpartition top-k,
    then sorted best-first; flat image-major lists of names and np.float32 scores."""
    cls_result_list, score_result_list = [], []
    for idx in range(probs.shape[0]):
        img_score = probs[idx, :]
        top_k_indices = np.argpartition(img_score, -top_k)[-top_k:]
        top_k_scores = img_score[top_k_indices]
        sort_ind = np.argsort(-top_k_scores)
        score_result_list.extend(top_k_scores[sort_ind])
        cls_result_list.extend(id_to_class_dict[int(top_k_indices[i])] for i in sort_ind)
    return cls_result_list, score_result_list


class ClipWrapper:
    """Zero-shot scoring of depth images with the CLIP ViT-B/16 visual tower on the B200.

    ``ClipWrapper(clip_cfg, model_path, device)`` works unmodified from the reference's call site
    (tools/preprocess_data.py:48): the checkpoint is loaded and the prompts are encoded once with
    the caller's ``clip`` package, then the visual tower moves into the library.  ``text_features``
    / ``visual_state_dict`` let a caller that already holds them skip that start-up step."""

    def __init__(self, clip_cfg, model_path, device=None, text_features=None,
                 visual_state_dict=None, engine: Optional[Engine] = None, num_views: int = 4):
        assert model_path is not None or visual_state_dict is not None, 'model_path is None'
        if device is None:
            device = 'cuda'
        if not str(device).startswith('cuda'):
            raise RuntimeError("vilgod_b200.ClipWrapper runs on the GPU only (no CPU fallback)")
        self.device = device
        self.top_k = _cfg_get(clip_cfg, "top_k", 1)
        self.split_size = _cfg_get(clip_cfg, "split_size", 50)
        self.template = _cfg_get(clip_cfg, "prompt_template", "a point representation of a {}")
        class_list = list(_cfg_get(clip_cfg, "class_list", CLASS_LIST))
        class_mapping = dict(_cfg_get(clip_cfg, "class_mapping", CLASS_MAPPING))
        self.id_to_class_dict = {i: c for i, c in enumerate(class_list)}
        if not 1 <= int(self.top_k) <= len(class_list):
            raise ValueError(f"top_k must be in [1, {len(class_list)}]")
        self.model = self.preprocess = self.text_tokenized = None
        if text_features is None:
            self.model, self.preprocess, self.text_tokenized, text_features = \
                encode_prompts_with_clip(clip_cfg, model_path, device)
            if visual_state_dict is None:
                visual_state_dict = {k: v.float() for k, v in self.model.visual.state_dict().items()}
        if visual_state_dict is None:
            visual_state_dict = self._load_visual(Path(model_path) / _cfg_get(clip_cfg, "model_name"))
        self.engine = engine or Engine(num_views=num_views)
        self.engine.load_vit_weights(visual_state_dict)
        self.engine.set_text_features(torch.as_tensor(text_features).float(), class_list, class_mapping)
        self.text_features = torch.as_tensor(text_features)

    @staticmethod
    def _load_visual(path):
        try:
            sd = torch.jit.load(str(path), map_location="cpu").state_dict()
        except RuntimeError:
            sd = torch.load(str(path), map_location="cpu")
        vis = {k[len("visual."):]: v.float() for k, v in sd.items() if k.startswith("visual.")}
        # build_model() pushes conv / linear / attention / proj weights through fp16
        for k in list(vis):
            if k.endswith(("in_proj_weight", "in_proj_bias", "proj")) or \
                    ((k.endswith(".weight") or k.endswith(".bias")) and ".ln_" not in k
                     and not k.startswith("ln_")):
                vis[k] = vis[k].half().float()
        return vis

    def predict_clip_labels(self, images):
        """list of 224x224 PIL images (3 identical channels, as the reference's depth images are)
        -> (class names, scores): n * top_k entries, image-major, best first (clip_utils.py:34-63)."""
        if len(images) == 0:
            return [], []
        arr = np.stack([np.asarray(im) for im in images])
        if arr.ndim == 4:
            if not (np.array_equal(arr[..., 0], arr[..., 1]) and np.array_equal(arr[..., 0], arr[..., 2])):
                raise ValueError("depth images must have three identical channels")
            arr = arr[..., 0]
        if arr.shape[1:] != (224, 224) or arr.dtype != np.uint8:
            raise ValueError("expected uint8 224x224 images")
        u8 = torch.from_numpy(np.ascontiguousarray(arr)).to(self.engine.device)
        res = self.engine.encode_score(u8_to_tiles(u8, self.engine.op_torch_dtype), want_feats=False)
        probs = res["probs"].cpu().numpy()
        if self.top_k == 1:        # the fused arg-max (first maximum, like argpartition on distinct scores)
            top1 = res["top1"].cpu().numpy()
            return ([self.id_to_class_dict[int(i)] for i in top1],
                    [probs[i, top1[i]] for i in range(len(top1))])
        return top_k_labels(probs, int(self.top_k), self.id_to_class_dict)


def _empty_frame(engine, V, top_k):
    z = np.zeros((0, V * top_k))
    return dict(class_names=z.astype(str), class_names_detailed=z.astype(str),
                class_scores=z.astype(np.float32), voted_names=np.zeros((0,), dtype=str),
                voted_scores=np.zeros((0,), np.float32), status=np.zeros((0,), np.int32),
                depth_images=[])


def classify_frame(engine: Engine, clusters, transform_to_ego=None, key=None, gpu_canonicalise=False,
                   top_k: int = 1, want_depth_images: bool = False):
    """One frame of ZeroShotDetector.classification on the GPU.

    clusters: list of [N_i, >=3] arrays (``det.cluster_points``) in the reference frame.
    Returns the arrays the reference hands to ``update_object_classes``
    (zero_shot_detector.py:412-417): mapped names [C, V*top_k], detailed names, scores f32, the
    first view's depth image of every cluster as PIL images (``depth_images``, on request), plus the
    GPU vote (names [C], scores [C]; top_k = 1) and the per-cluster status.  A frame without clusters
    returns empty arrays (the reference skips it, zero_shot_detector.py:403); a cluster without
    points comes back with status VG_EDEGENERATE (the reference would produce NaN)."""
    V = engine.num_views
    if len(clusters) == 0:
        return _empty_frame(engine, V, top_k)
    pts = [np.asarray(c)[..., :3].reshape(-1, 3) for c in clusters]
    offsets = np.zeros(len(pts) + 1, dtype=np.int32)
    offsets[1:] = np.cumsum([len(p) for p in pts])
    if gpu_canonicalise:     # no host loop at all (fp32 cluster points expected)
        packed, _ = engine.canonicalise(np.concatenate(pts).astype(np.float32), offsets, transform_to_ego)
    else:
        packed = canonicalise.canonicalise_packed(np.concatenate(pts), offsets, transform_to_ego)
    out = engine.classify(packed, offsets, want_feats=False, want_depth_u8=want_depth_images)
    top1 = out["top1"].cpu().numpy()
    probs = out["probs"].cpu().numpy()
    class_list = np.asarray(engine.class_list)
    mapped_of_prompt = np.asarray(engine.mapped_names)[engine.class_map]
    if top_k == 1:
        scores = np.take_along_axis(probs, top1[..., None].astype(np.int64), axis=2)[..., 0]
        detailed = class_list[top1]
        mapped = mapped_of_prompt[top1]
        voted_names = np.asarray(engine.mapped_names)[out["voted_class"].cpu().numpy()]
        voted_scores = out["voted_score"].cpu().numpy()
    else:
        C = len(pts)
        names, sc = top_k_labels(probs.reshape(C * V, -1), top_k, dict(enumerate(engine.class_list)))
        detailed = np.stack(names).reshape(C, -1)
        scores = np.stack(sc).reshape(C, -1)
        to_mapped = dict(zip(engine.class_list, mapped_of_prompt))
        mapped = np.vectorize(to_mapped.get)(detailed)
        voted_names = voted_scores = None      # the GPU vote is per top-1; the host vote takes over
    depth_images = None
    if want_depth_images:
        from PIL import Image
        u8 = out["depth_u8"].cpu().numpy()
        depth_images = [Image.fromarray(np.repeat(a[..., None], 3, axis=2)) for a in u8]
    return dict(class_names=mapped, class_names_detailed=detailed,
                class_scores=scores.astype(np.float32), voted_names=voted_names,
                voted_scores=voted_scores, status=out["status"].cpu().numpy(),
                depth_images=depth_images)


def classify_lidar_frame(engine: Engine, lidar_frame, detections, cluster_update_list, key_,
                         aggregation='voting', classify_gt=False, classified_detections=False,
                         top_k: int = 1, gpu_canonicalise=False):
    """Drop-in for the per-frame body of ``ZeroShotDetector.classification``
    (zero_shot_detector.py:389-417): the same per-detection filter and ``cluster_update_list``
    bookkeeping as the reference, one fused GPU call instead of the per-cluster loop, and the same
    ``update_object_classes`` call including ``depth_images`` (the first view of every cluster).
    Returns the number of clusters classified (the reference's ``length``)."""
    clusters = []
    for d_idx, det in enumerate(detections):
        if (det.gt and classify_gt) or (not det.gt and not classified_detections):
            clusters.append(det.cluster_points[..., :3])
            cluster_update_list[d_idx] &= True
        else:
            cluster_update_list[d_idx] &= False
    length = len(clusters)
    if length > 0:
        res = classify_frame(engine, clusters, lidar_frame.transform_to_ego, top_k=top_k,
                             want_depth_images=True, gpu_canonicalise=gpu_canonicalise)
        lidar_frame.update_object_classes(res["class_names"], res["class_names_detailed"],
                                          res["class_scores"], cluster_update_list, key=key_,
                                          aggregation=aggregation, depth_images=res["depth_images"])
    return length
