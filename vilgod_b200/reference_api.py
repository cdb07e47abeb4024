"""Drop-in mirrors of the two reference operators on the hot path, plus the fused frame call.

Same names, argument meaning and error behaviour as the reference:

  RealisticProjection(cfg).get_img(points)          reference src/utils/mv_utils.py:130-201
  ClipWrapper(clip_cfg, model_path, device)         reference src/utils/clip_utils.py:10-63
      .predict_clip_labels(list_of_PIL_images)
  classify_frame(...)                               loop body of ZeroShotDetector.classification,
                                                    reference src/vilgod/zero_shot_detector.py:389-416

Everything numerical runs in libvilgod_b200.so on the GPU.  The CLIP *text* tower and tokenizer run
once at start-up in the reference (clip_utils.py:23-26) and are not part of the hot path: their
output, the cached ``text_features [P,512]``, is an input here.
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


class ClipWrapper:
    """Zero-shot scoring of depth images with the CLIP ViT-B/16 visual tower on the B200."""

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
        if self.top_k != 1:
            raise NotImplementedError("the fused head returns top-1 (reference config: top_k = 1)")
        if visual_state_dict is None:
            visual_state_dict = self._load_visual(Path(model_path) / _cfg_get(clip_cfg, "model_name"))
        if text_features is None:
            raise ValueError(
                "text_features [P,512] required: encode the prompts once with the reference's own "
                "text tower (clip_utils.py:23-26) and pass the cached, L2-normalised tensor")
        self.engine = engine or Engine(num_views=num_views)
        self.engine.load_vit_weights(visual_state_dict)
        self.engine.set_text_features(text_features, class_list, class_mapping)
        self.text_features = torch.as_tensor(text_features)
        self.model = None
        self.preprocess = None   # folded into the patch-embedding weights

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
        -> (class names, scores), image-major, like the reference with top_k = 1."""
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
        top1 = res["top1"].cpu().numpy()
        names = [self.id_to_class_dict[int(i)] for i in top1]
        scores = [probs[i, top1[i]] for i in range(len(top1))]
        return names, scores


def classify_frame(engine: Engine, clusters, transform_to_ego=None, key=None, gpu_canonicalise=False):
    """One frame of ZeroShotDetector.classification on the GPU.

    clusters: list of [N_i, >=3] arrays (``det.cluster_points``) in the reference frame.
    Returns the arrays the reference hands to ``update_object_classes``
    (zero_shot_detector.py:412-416): mapped names [C,V], detailed names [C,V], scores [C,V] f32,
    plus the GPU vote (names [C], scores [C])."""
    pts = [np.asarray(c)[..., :3] for c in clusters]
    offsets = np.zeros(len(pts) + 1, dtype=np.int32)
    offsets[1:] = np.cumsum([len(p) for p in pts])
    if gpu_canonicalise:     # no host loop at all (fp32 cluster points expected)
        packed, _ = engine.canonicalise(np.concatenate(pts).astype(np.float32), offsets, transform_to_ego)
    else:
        packed = canonicalise.canonicalise_packed(np.concatenate(pts), offsets, transform_to_ego)
    out = engine.classify(packed, offsets, want_feats=False)
    top1 = out["top1"].cpu().numpy()
    probs = out["probs"].cpu().numpy()
    scores = np.take_along_axis(probs, top1[..., None].astype(np.int64), axis=2)[..., 0]
    detailed = np.asarray(engine.class_list)[top1]
    mapped = np.asarray(engine.mapped_names)[engine.class_map[top1]]
    vc = out["voted_class"].cpu().numpy()
    voted_names = np.asarray(engine.mapped_names)[vc]
    return dict(class_names=mapped, class_names_detailed=detailed,
                class_scores=scores.astype(np.float32), voted_names=voted_names,
                voted_scores=out["voted_score"].cpu().numpy(), status=out["status"].cpu().numpy())
