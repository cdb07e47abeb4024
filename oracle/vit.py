"""TEST INFRASTRUCTURE ONLY -- fp32 torch restatement (oracle) of the CLIP ViT-B/16 visual tower,
the CLIP preprocess and the zero-shot scoring used by ViLGOD.  The product never imports this.

Parity status: pinned against outputs of the reference itself (tests/golden/vit_*.npz, made by
oracle/make_golden.py from the unmodified third_party/CLIP/clip/model.py + src/utils/clip_utils.py).

Restated reference code:
  make_visual_weights   third_party/CLIP/clip/model.py:206-221 (constructor / RNG order),
                        :375-396 (convert_weights: fp16 rounding), :399-436 (build_model)
  preprocess_u8         third_party/CLIP/clip/clip.py:79-86 as called at src/utils/clip_utils.py:35
  vit_forward           third_party/CLIP/clip/model.py:223-240, :171-192, :157-168
  score                 src/utils/clip_utils.py:41-43
  top1_labels           src/utils/clip_utils.py:51-61 (top_k = 1)
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

WIDTH, LAYERS, HEADS, PATCH, RES, EMBED = 768, 12, 12, 16, 224, 512
TOKENS = (RES // PATCH) ** 2 + 1
CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


def make_visual_weights(seed=1234, fp16_round=True):
    """Random-init visual tower with the reference's parameter names.

    CLIP.__init__ builds ``self.visual`` before anything else (model.py:263-281), so the visual
    parameters depend only on the seed and on the order in which VisionTransformer.__init__
    constructs its sub-modules; constructing the same torch.nn modules in the same order consumes
    the RNG identically.  ``fp16_round`` reproduces convert_weights (conv / linear / MHA / proj go
    through fp16; LayerNorm, class and positional embeddings stay fp32) followed by ``.float()``
    (clip.py:140-141)."""
    from torch import nn

    torch.manual_seed(seed)
    w = {}
    conv1 = nn.Conv2d(3, WIDTH, PATCH, PATCH, bias=False)
    scale = WIDTH ** -0.5
    w["conv1.weight"] = conv1.weight.detach()
    w["class_embedding"] = scale * torch.randn(WIDTH)
    w["positional_embedding"] = scale * torch.randn(TOKENS, WIDTH)
    w["ln_pre.weight"], w["ln_pre.bias"] = torch.ones(WIDTH), torch.zeros(WIDTH)
    half = set(["conv1.weight", "proj"])
    for i in range(LAYERS):
        p = f"transformer.resblocks.{i}."
        attn = nn.MultiheadAttention(WIDTH, HEADS)
        c_fc = nn.Linear(WIDTH, WIDTH * 4)
        c_proj = nn.Linear(WIDTH * 4, WIDTH)
        w[p + "attn.in_proj_weight"] = attn.in_proj_weight.detach()
        w[p + "attn.in_proj_bias"] = attn.in_proj_bias.detach()
        w[p + "attn.out_proj.weight"] = attn.out_proj.weight.detach()
        w[p + "attn.out_proj.bias"] = attn.out_proj.bias.detach()
        w[p + "ln_1.weight"], w[p + "ln_1.bias"] = torch.ones(WIDTH), torch.zeros(WIDTH)
        w[p + "mlp.c_fc.weight"], w[p + "mlp.c_fc.bias"] = c_fc.weight.detach(), c_fc.bias.detach()
        w[p + "mlp.c_proj.weight"] = c_proj.weight.detach()
        w[p + "mlp.c_proj.bias"] = c_proj.bias.detach()
        w[p + "ln_2.weight"], w[p + "ln_2.bias"] = torch.ones(WIDTH), torch.zeros(WIDTH)
        half.update(p + s for s in ("attn.in_proj_weight", "attn.in_proj_bias",
                                    "attn.out_proj.weight", "attn.out_proj.bias",
                                    "mlp.c_fc.weight", "mlp.c_fc.bias", "mlp.c_proj.weight",
                                    "mlp.c_proj.bias"))
    w["ln_post.weight"], w["ln_post.bias"] = torch.ones(WIDTH), torch.zeros(WIDTH)
    w["proj"] = scale * torch.randn(WIDTH, EMBED)
    out = {}
    for k, v in w.items():
        v = v.detach().clone().float()
        if fp16_round and k in half:
            v = v.half().float()
        out[k] = v.contiguous()
    return out


def perturb_layernorms(weights, seed=7, amount=0.2):
    """Test helper: random-init LayerNorms are identity (weight 1, bias 0), which would hide
    gamma/beta indexing bugs.  Returns a copy with perturbed LN parameters."""
    g = torch.Generator().manual_seed(seed)
    out = dict(weights)
    for k in weights:
        if ".ln_" in k or k.startswith("ln_"):
            if k.endswith("weight"):
                out[k] = (1.0 + amount * torch.randn(WIDTH, generator=g)).float()
            else:
                out[k] = (amount * torch.randn(WIDTH, generator=g)).float()
    return out


def preprocess_u8(u8):
    """uint8 [B,S,S] single-channel depth image (the three PIL channels are identical) ->
    fp32 [B,3,S,S]:  ToTensor (/255) then Normalize((x-mean)/std)  -- clip.py:79-86."""
    x = torch.as_tensor(np.asarray(u8)).to(torch.float32).div(255)
    x = x[:, None].repeat(1, 3, 1, 1)
    mean = torch.tensor(CLIP_MEAN, dtype=torch.float32).view(1, 3, 1, 1)
    std = torch.tensor(CLIP_STD, dtype=torch.float32).view(1, 3, 1, 1)
    return (x - mean) / std


def _ln(x, w, b):
    return F.layer_norm(x.float(), (x.shape[-1],), w, b, 1e-5)


def _block(x, w, p, stages=None):
    """x [B,L,W] (batch first; the reference's LND permutes are pure layout)."""
    B, L, W = x.shape
    hd = W // HEADS
    y = _ln(x, w[p + "ln_1.weight"], w[p + "ln_1.bias"])
    qkv = F.linear(y, w[p + "attn.in_proj_weight"], w[p + "attn.in_proj_bias"])
    q, k, v = qkv.split(W, dim=-1)
    q = q.view(B, L, HEADS, hd).transpose(1, 2) * (1.0 / math.sqrt(hd))
    k = k.view(B, L, HEADS, hd).transpose(1, 2)
    v = v.view(B, L, HEADS, hd).transpose(1, 2)
    a = torch.softmax(q @ k.transpose(-1, -2), dim=-1) @ v
    a = a.transpose(1, 2).reshape(B, L, W)
    x = x + F.linear(a, w[p + "attn.out_proj.weight"], w[p + "attn.out_proj.bias"])
    y = _ln(x, w[p + "ln_2.weight"], w[p + "ln_2.bias"])
    h = F.linear(y, w[p + "mlp.c_fc.weight"], w[p + "mlp.c_fc.bias"])
    h = h * torch.sigmoid(1.702 * h)
    x = x + F.linear(h, w[p + "mlp.c_proj.weight"], w[p + "mlp.c_proj.bias"])
    return x


@torch.no_grad()
def vit_forward(w, x, return_stages=False):
    """x fp32 [B,3,224,224] (already preprocessed) -> image features [B,512] (not normalised)."""
    stages = {}
    x = F.conv2d(x, w["conv1.weight"], stride=PATCH)
    x = x.reshape(x.shape[0], x.shape[1], -1).permute(0, 2, 1)
    cls = w["class_embedding"].expand(x.shape[0], 1, -1)
    x = torch.cat([cls, x], dim=1) + w["positional_embedding"]
    stages["embed"] = x
    x = _ln(x, w["ln_pre.weight"], w["ln_pre.bias"])
    stages["ln_pre"] = x
    for i in range(LAYERS):
        x = _block(x, w, f"transformer.resblocks.{i}.")
        if return_stages and i in (0, LAYERS - 1):
            stages[f"block{i}"] = x
    x = _ln(x[:, 0, :], w["ln_post.weight"], w["ln_post.bias"])
    f = x @ w["proj"]
    return (f, stages) if return_stages else f


@torch.no_grad()
def encode_u8(w, u8, batch=50):
    """u8 [B,S,S] -> features [B,512]; mini-batches of ``split_size`` like clip_utils.py:37."""
    out = []
    for i in range(0, len(u8), batch):
        out.append(vit_forward(w, preprocess_u8(u8[i:i + batch])))
    return torch.cat(out)


@torch.no_grad()
def score(feats, text_features):
    """clip_utils.py:41-43: f /= |f|; softmax(100 f T^T).  Returns (probs, logits, f_normed)."""
    f = torch.as_tensor(feats).float()
    t = torch.as_tensor(text_features).float()
    f = f / f.norm(dim=-1, keepdim=True)
    logits = 100.0 * f @ t.T
    return logits.softmax(dim=-1), logits, f


def top1_labels(probs, class_list):
    """clip_utils.py:48-61 with top_k = 1: index of the largest probability, its name and score."""
    p = np.asarray(probs)
    idx = np.array([int(np.argpartition(r, -1)[-1:][0]) for r in p], dtype=np.int64)
    return idx, [class_list[i] for i in idx], p[np.arange(len(p)), idx]
