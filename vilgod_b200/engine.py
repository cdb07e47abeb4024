"""Host side of the B200 hot path: owns a VgHandle and moves torch tensors across the C ABI.

PyTorch is plumbing here (device memory, streams); every computation on the path happens inside
libvilgod_b200.so.  ``Engine.classify`` is the fused surface of SURVEY.md section 8b
(``classify_clusters``): canonicalised packed clusters in, per-view probabilities / top-1 /
embeddings and the per-cluster vote out.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Sequence

import numpy as np
import torch

from . import _lib, views
from ._lib import VilgodError

# tools/configs/preprocessor/waymo.yaml:110-140
CLASS_LIST = ['car', 'truck', 'bus', 'van', 'minivan', 'pickup truck', 'school bus', 'fire truck',
              'ambulance', 'pedestrian', 'human body', 'human', 'cyclist', 'rider', 'bicycle',
              'bike', 'traffic light', 'traffic sign', 'fence', 'pole', 'clutter', 'tree', 'house',
              'wall']
CLASS_MAPPING = {**{k: 'Vehicle' for k in CLASS_LIST[0:9]},
                 **{k: 'Pedestrian' for k in CLASS_LIST[9:12]},
                 **{k: 'Cyclist' for k in CLASS_LIST[12:16]},
                 **{k: 'Background' for k in CLASS_LIST[16:24]}}

_LAYER_KEYS = [("ln_1_weight", "ln_1.weight"), ("ln_1_bias", "ln_1.bias"),
               ("attn_in_proj_weight", "attn.in_proj_weight"),
               ("attn_in_proj_bias", "attn.in_proj_bias"),
               ("attn_out_proj_weight", "attn.out_proj.weight"),
               ("attn_out_proj_bias", "attn.out_proj.bias"),
               ("ln_2_weight", "ln_2.weight"), ("ln_2_bias", "ln_2.bias"),
               ("mlp_c_fc_weight", "mlp.c_fc.weight"), ("mlp_c_fc_bias", "mlp.c_fc.bias"),
               ("mlp_c_proj_weight", "mlp.c_proj.weight"), ("mlp_c_proj_bias", "mlp.c_proj.bias")]
_TOP_KEYS = [("conv1_weight", "conv1.weight"), ("class_embedding", "class_embedding"),
             ("positional_embedding", "positional_embedding"), ("ln_pre_weight", "ln_pre.weight"),
             ("ln_pre_bias", "ln_pre.bias"), ("ln_post_weight", "ln_post.weight"),
             ("ln_post_bias", "ln_post.bias"), ("proj", "proj")]
_SHAPES = {"conv1.weight": (768, 3, 16, 16), "class_embedding": (768,),
           "positional_embedding": (197, 768), "proj": (768, 512),
           "attn.in_proj_weight": (2304, 768), "attn.in_proj_bias": (2304,),
           "attn.out_proj.weight": (768, 768), "mlp.c_fc.weight": (3072, 768),
           "mlp.c_fc.bias": (3072,), "mlp.c_proj.weight": (768, 3072)}


def _ptr(t: Optional[torch.Tensor]):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class Engine:
    """One handle on the current CUDA device."""

    def __init__(self, num_views: int = 4, rot_mat=None, resolution: int = 112, depth: int = 8,
                 image_size: int = 224, obj_ratio: float = 0.8, depth_bias: float = 0.2,
                 gauss=None, logit_scale: float = 100.0, rotate_mode: int = _lib.VG_ROTATE_TORCH_CPU,
                 device=None, operand_dtype: str = "f16", div_mode: int = _lib.VG_DIV_TRUE,
                 pool_kernel: int = 5, pool_pad: int = 1):
        if not torch.cuda.is_available():
            raise RuntimeError("vilgod_b200 needs an sm_100 CUDA device; there is no CPU fallback")
        self.lib = _lib.load(operand_dtype)
        self.operand_dtype = operand_dtype
        self.op_torch_dtype = torch.float16 if operand_dtype == "f16" else torch.bfloat16
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None \
            else torch.device(device)
        rot = views.view_rot_mats(num_views) if rot_mat is None else torch.as_tensor(rot_mat).float()
        rot = rot.detach().cpu().contiguous()
        if rot.shape != (num_views, 3, 3):
            raise ValueError(f"rot_mat must be [{num_views},3,3]")
        g = views.gaussian_weights() if gauss is None else torch.as_tensor(gauss).float()
        cfg = _lib.VgConfig()
        cfg.abi_version = _lib.VG_ABI_VERSION
        cfg.resolution, cfg.depth, cfg.image_size = resolution, depth, image_size
        cfg.num_views, cfg.rotate_mode = num_views, rotate_mode
        cfg.div_mode, cfg.pool_kernel, cfg.pool_pad = div_mode, pool_kernel, pool_pad
        cfg.obj_ratio, cfg.depth_bias, cfg.logit_scale = obj_ratio, depth_bias, logit_scale
        flat = rot.reshape(num_views, 9).numpy()
        for v in range(num_views):
            for k in range(9):
                cfg.rot[v][k] = float(flat[v, k])
        gf = g.reshape(9).cpu().numpy()
        for k in range(9):
            cfg.gauss[k] = float(gf[k])
        self.num_views, self.resolution, self.image_size = num_views, resolution, image_size
        self.rot_mat = rot
        self.gauss = g.reshape(3, 3)
        self.num_prompts = 0
        self.class_list: Sequence[str] = ()
        self.mapped_names: Sequence[str] = ()
        self._ws = None
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            rc = self.lib.vg_create(C.byref(cfg), C.byref(self._h))
        if rc != _lib.VG_OK:
            raise VilgodError(rc, "vg_create failed (needs an sm_100 device, R in {112, 224}, D=8, "
                                  "S=224, max-pool 5/1)")

    # ------------------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self.lib.vg_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != _lib.VG_OK:
            raise VilgodError(rc, self.lib.vg_last_error(self._h).decode(errors="replace"))

    @property
    def launch_count(self) -> int:
        return int(self.lib.vg_launch_count(self._h))

    # ------------------------------------------------------------------------------------------
    def load_vit_weights(self, state_dict: Dict[str, torch.Tensor]):
        """``clip_model.visual.state_dict()`` of the reference (names as in SURVEY.md appendix B;
        a ``visual.`` prefix is accepted).  Tensors are staged as fp32 on the device; the library
        makes its own bf16 / folded copies, so nothing is retained from ``state_dict``."""
        sd = {(k[len("visual."):] if k.startswith("visual.") else k): v for k, v in state_dict.items()}
        keep = []

        def dev(name):
            if name not in sd:
                raise KeyError(f"missing ViT parameter {name!r}")
            t = sd[name].detach().to(device=self.device, dtype=torch.float32).contiguous()
            base = name.split("resblocks.")[-1].split(".", 1)[-1] if "resblocks." in name else name
            want = _SHAPES.get(base)
            if want is not None and tuple(t.shape) != want:
                raise ValueError(f"{name}: expected shape {want}, got {tuple(t.shape)}")
            keep.append(t)
            return t.data_ptr()

        w = _lib.VgVitWeights()
        for field, name in _TOP_KEYS:
            setattr(w, field, dev(name))
        for i in range(_lib.VG_VIT_LAYERS):
            for field, name in _LAYER_KEYS:
                setattr(w.layers[i], field, dev(f"transformer.resblocks.{i}.{name}"))
        with torch.cuda.device(self.device):
            self._check(self.lib.vg_load_vit_weights(self._h, C.byref(w), _stream()))
        del keep

    def set_text_features(self, text_features, class_list=CLASS_LIST, class_mapping=CLASS_MAPPING):
        """Cached, L2-normalised prompt embeddings [P,512] (reference clip_utils.py:23-26) and the
        prompt -> mapped class table.  Mapped class ids follow the alphabetical order of the mapped
        names, which is the order numpy.unique gives the reference's vote."""
        t = torch.as_tensor(text_features).detach().to(device=self.device, dtype=torch.float32).contiguous()
        P = t.shape[0]
        if t.shape != (P, 512) or len(class_list) != P:
            raise ValueError("text_features must be [P,512] with one class name per row")
        self.class_list = list(class_list)
        self.mapped_names = sorted(set(class_mapping[c] for c in class_list))
        ids = np.asarray([self.mapped_names.index(class_mapping[c]) for c in class_list], np.int32)
        self.class_map = ids
        with torch.cuda.device(self.device):
            self._check(self.lib.vg_set_text_features(
                self._h, _ptr(t), P, ids.ctypes.data_as(C.POINTER(C.c_int32)),
                len(self.mapped_names), _stream()))
        self.num_prompts = P

    # ------------------------------------------------------------------------------------------
    def workspace(self, max_images: int) -> torch.Tensor:
        need = int(self.lib.vg_workspace_bytes(self._h, int(max_images)))
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._ws

    def _packed(self, points, offsets):
        p = torch.as_tensor(points)
        p = p.to(device=self.device, dtype=torch.float32).contiguous()
        o = torch.as_tensor(offsets).to(device=self.device, dtype=torch.int32).contiguous()
        if p.ndim != 2 or p.shape[1] != 3 or o.ndim != 1 or o.numel() < 1:
            raise ValueError("points must be [sum N, 3] and offsets [C+1]")
        host = offsets if isinstance(offsets, np.ndarray) else (
            offsets.numpy() if isinstance(offsets, torch.Tensor) and offsets.device.type == "cpu" else None)
        if host is not None and host.size:      # cheap when the offsets are on the host anyway
            if host[0] < 0 or host[-1] > p.shape[0] or np.any(np.diff(host) < 0):
                raise ValueError("offsets must be non-decreasing within [0, len(points)]")
        return p, o, o.numel() - 1

    def canonicalise(self, points_raw, offsets, transform_to_ego=None):
        """GPU version of ``canonicalise.canonicalise_packed``: raw fp32 cluster points [sum N,3] (+ the
        frame's 4x4 ego transform) -> canonicalised fp32 points on the device."""
        p, o, Cn = self._packed(points_raw, offsets)
        out = torch.empty_like(p)
        status = torch.empty((Cn,), dtype=torch.int32, device=self.device)
        T = None
        if transform_to_ego is not None:
            T = torch.as_tensor(np.asarray(transform_to_ego, dtype=np.float64)).to(self.device).contiguous()
            if T.shape != (4, 4):
                raise ValueError("transform_to_ego must be 4x4")
        with torch.cuda.device(self.device):
            self._check(self.lib.vg_canonicalise(self._h, _ptr(p), _ptr(o), Cn, _ptr(T), _ptr(out),
                                                 _ptr(status), _stream()))
        return out, status

    def project(self, points, offsets, want_tiles=True, want_u8=False, want_grid=False,
                want_densified=False):
        """Packed clusters -> dict with 'tiles' bf16 [B,196,256], 'u8' [B,224,224], 'status' [C],
        and the debug taps 'grid' [B,8,112,112] / 'densified' [B,110,110] on request."""
        p, o, Cn = self._packed(points, offsets)
        B, R, S = Cn * self.num_views, self.resolution, self.image_size
        out = {}
        tiles = torch.empty((B, 196, 256), dtype=self.op_torch_dtype, device=self.device) if want_tiles else None
        u8 = torch.empty((B, S, S), dtype=torch.uint8, device=self.device) if want_u8 else None
        status = torch.empty((Cn,), dtype=torch.int32, device=self.device)
        dbg = _lib.VgProjectDebug()
        grid = torch.empty((B, 8, R, R), dtype=torch.float32, device=self.device) if want_grid else None
        dens = torch.empty((B, R - 2, R - 2), dtype=torch.float32, device=self.device) if want_densified else None
        dbg.d_grid = grid.data_ptr() if grid is not None else None
        dbg.d_densified = dens.data_ptr() if dens is not None else None
        with torch.cuda.device(self.device):
            self._check(self.lib.vg_project(self._h, _ptr(p), _ptr(o), Cn, _ptr(tiles), _ptr(u8),
                                            _ptr(status), C.byref(dbg), _stream()))
        out.update(tiles=tiles, u8=u8, status=status, grid=grid, densified=dens)
        return out

    def encode_score(self, tiles: torch.Tensor, want_feats=True, want_logits=False,
                     stop_after_layer: Optional[int] = None):
        """bf16 tiles [B,196,256] -> dict(probs [B,P], top1 [B], feats [B,512], logits [B,P]).
        ``stop_after_layer`` (-2: after ln_pre, k: after resblock k) returns only 'x' [B,197,768]."""
        if tiles.dtype != self.op_torch_dtype or tiles.ndim != 3 or tiles.shape[1:] != (196, 256):
            raise ValueError(f"tiles must be {self.op_torch_dtype} [B,196,256]")
        tiles = tiles.to(self.device).contiguous()
        B, P = tiles.shape[0], self.num_prompts
        ws = self.workspace(B)
        dbg = None
        x = None
        if stop_after_layer is not None:
            x = torch.empty((B, 197, 768), dtype=torch.float32, device=self.device)
            dbg = _lib.VgVitDebug(int(stop_after_layer), x.data_ptr())
        probs = torch.empty((B, P), dtype=torch.float32, device=self.device)
        top1 = torch.empty((B,), dtype=torch.int32, device=self.device)
        feats = torch.empty((B, 512), dtype=torch.float32, device=self.device) if want_feats else None
        logits = torch.empty((B, P), dtype=torch.float32, device=self.device) if want_logits else None
        with torch.cuda.device(self.device):
            self._check(self.lib.vg_encode_score(
                self._h, _ptr(tiles), B, _ptr(probs), _ptr(top1), _ptr(feats), _ptr(logits),
                _ptr(ws), ws.numel(), C.byref(dbg) if dbg is not None else None, _stream()))
        if stop_after_layer is not None:
            return dict(x=x)
        return dict(probs=probs, top1=top1, feats=feats, logits=logits)

    def vote(self, probs: torch.Tensor, top1: torch.Tensor):
        Cn = probs.shape[0] // self.num_views
        vc = torch.empty((Cn,), dtype=torch.int32, device=self.device)
        vs = torch.empty((Cn,), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            self._check(self.lib.vg_vote(self._h, _ptr(probs.contiguous()), _ptr(top1.contiguous()),
                                         Cn, _ptr(vc), _ptr(vs), _stream()))
        return vc, vs

    def classify(self, points, offsets, want_feats=True, out=None, want_depth_u8=False):
        """The fused surface: device (or host) packed clusters -> device results.
        Returns dict(probs [C,V,P], top1 [C,V], feats [C,V,512], voted_class [C], voted_score [C],
        status [C]) and, with ``want_depth_u8``, depth_u8 [C,224,224]: the first view's image of
        every cluster (what the reference keeps as ``det.depth_image``).
        Under CUDA-graph capture keep ``out`` and the workspace alive (and call ``workspace()`` with
        the largest batch first): a later, larger batch reallocates the workspace and a graph that
        captured the old one must not be replayed."""
        p, o, Cn = self._packed(points, offsets)
        V, P = self.num_views, self.num_prompts
        ws = self.workspace(Cn * V)
        if out is None:
            out = self.alloc_outputs(Cn, want_feats, want_depth_u8)
        else:
            for k, shape in (("probs", (Cn, V, P)), ("top1", (Cn, V)), ("voted_class", (Cn,)),
                             ("voted_score", (Cn,)), ("status", (Cn,)), ("feats", (Cn, V, 512)),
                             ("depth_u8", (Cn, self.image_size, self.image_size))):
                t = out.get(k)
                if t is None and k in ("feats", "depth_u8"):
                    continue
                if t is None or tuple(t.shape) != shape or not t.is_contiguous() or t.device != self.device:
                    raise ValueError(f"out[{k!r}] must be a contiguous {shape} tensor on {self.device}")
        with torch.cuda.device(self.device):
            self._check(self.lib.vg_classify(
                self._h, _ptr(p), _ptr(o), Cn, _ptr(out["probs"]), _ptr(out["top1"]),
                _ptr(out.get("feats")), _ptr(out["voted_class"]), _ptr(out["voted_score"]),
                _ptr(out["status"]), _ptr(out.get("depth_u8")), _ptr(ws), ws.numel(), _stream()))
        return out

    def alloc_outputs(self, Cn, want_feats=True, want_depth_u8=False):
        V, P, d = self.num_views, self.num_prompts, self.device
        return dict(
            depth_u8=torch.empty((Cn, self.image_size, self.image_size), dtype=torch.uint8, device=d)
            if want_depth_u8 else None,
            probs=torch.empty((Cn, V, P), dtype=torch.float32, device=d),
            top1=torch.empty((Cn, V), dtype=torch.int32, device=d),
            feats=torch.empty((Cn, V, 512), dtype=torch.float32, device=d) if want_feats else None,
            voted_class=torch.empty((Cn,), dtype=torch.int32, device=d),
            voted_score=torch.empty((Cn,), dtype=torch.float32, device=d),
            status=torch.empty((Cn,), dtype=torch.int32, device=d))

    def profile_begin(self):
        self._check(self.lib.vg_profile_begin(self._h))

    def profile_end(self):
        """-> {kernel class: dict(ms, launches, work)} from CUDA events around every launch."""
        kt = _lib.VgKernelTimes()
        self._check(self.lib.vg_profile_end(self._h, C.byref(kt)))
        return {n: dict(ms=kt.ms[i], launches=int(kt.launches[i]), work=kt.work[i])
                for i, n in enumerate(_lib.VG_K_NAMES)}

    # kernel-level hooks used by the tests -----------------------------------------------------
    def test_gemm(self, a, w, bias, epilogue, out=None):
        M, K = a.shape
        N = w.shape[0]
        if out is None:
            out = torch.empty((M, N), device=self.device,
                              dtype=torch.float32 if epilogue == _lib.VG_EPI_BIAS_RESID_F32 else self.op_torch_dtype)
        with torch.cuda.device(self.device):
            self._check(self.lib.vg_test_gemm(self._h, _ptr(a), _ptr(w), _ptr(bias), M, N, K,
                                              epilogue, _ptr(out), _stream()))
        return out

    def test_gemm_lnf(self, a, w, bias, epilogue, stats, colsum=None, x_inout=None):
        """The LayerNorm-folded GEMM instantiations of the tower.  bf16-output epilogues: returns out
        [M,N]; residual epilogue: ``x_inout`` fp32 [M,768] is updated in place, returns
        (x, xb [M,768] operand-typed copy) and fills ``stats`` [M,3,2]."""
        M, K = a.shape
        N = w.shape[0]
        resid = epilogue == _lib.VG_EPI_BIAS_RESID_F32
        out = x_inout if resid else torch.empty((M, N), device=self.device, dtype=self.op_torch_dtype)
        xb = torch.empty((M, N), device=self.device, dtype=self.op_torch_dtype) if resid else None
        with torch.cuda.device(self.device):
            self._check(self.lib.vg_test_gemm_lnf(self._h, _ptr(a), _ptr(w), _ptr(bias), _ptr(colsum),
                                                  _ptr(stats), _ptr(xb), M, N, K, epilogue, _ptr(out),
                                                  _stream()))
        return (out, xb) if resid else out

    def test_gemm_patch(self, tiles, w, table, x):
        """Patch embedding on the production kernel: rows 1..196 of x [B,197,768] fp32 are written."""
        with torch.cuda.device(self.device):
            self._check(self.lib.vg_test_gemm_patch(self._h, _ptr(tiles), _ptr(w), _ptr(table),
                                                    tiles.shape[0], _ptr(x), _stream()))
        return x

    def test_attention(self, qkv):
        B = qkv.shape[0]
        out = torch.empty((B, 197, 768), dtype=self.op_torch_dtype, device=self.device)
        with torch.cuda.device(self.device):
            self._check(self.lib.vg_test_attention(self._h, _ptr(qkv), B, _ptr(out), _stream()))
        return out

    def test_layernorm(self, x, w, b):
        y = torch.empty(x.shape, dtype=self.op_torch_dtype, device=self.device)
        with torch.cuda.device(self.device):
            self._check(self.lib.vg_test_layernorm(self._h, _ptr(x), _ptr(w), _ptr(b),
                                                   x.shape[0], _ptr(y), _stream()))
        return y


def u8_to_tiles(u8: torch.Tensor, dtype=torch.float16) -> torch.Tensor:
    """uint8 [B,224,224] -> patch-major tiles [B,196,256] in the engine's operand dtype (layout
    plumbing for callers that already hold images, e.g. ClipWrapper.predict_clip_labels)."""
    B = u8.shape[0]
    t = u8.reshape(B, 14, 16, 14, 16).permute(0, 1, 3, 2, 4).reshape(B, 196, 256)
    return t.to(dtype).contiguous()
