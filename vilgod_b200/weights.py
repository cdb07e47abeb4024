"""Synthetic ViT-B/16 checkpoints (there is no network for real CLIP weights).

``random_init_visual_state_dict`` reproduces, tensor for tensor, what the reference gets from
``build_model(CLIP(512,224,12,768,16,77,49408,512,8,12).state_dict())`` under a fixed seed
(third_party/CLIP/clip/model.py:206-221, 243-281 constructor order; :375-396 fp16 rounding):
the visual tower is constructed first, so its parameters depend only on the seed and on the order
in which the torch.nn modules are created.
"""
from __future__ import annotations

import torch
from torch import nn

WIDTH, LAYERS, HEADS, PATCH, TOKENS, EMBED = 768, 12, 12, 16, 197, 512


def random_init_visual_state_dict(seed: int = 1234, fp16_round: bool = True):
    torch.manual_seed(seed)
    sd = {}
    fp16_keys = {"conv1.weight", "proj"}
    sd["conv1.weight"] = nn.Conv2d(3, WIDTH, PATCH, PATCH, bias=False).weight
    scale = WIDTH ** -0.5
    sd["class_embedding"] = scale * torch.randn(WIDTH)
    sd["positional_embedding"] = scale * torch.randn(TOKENS, WIDTH)
    ones, zeros = torch.ones(WIDTH), torch.zeros(WIDTH)
    sd["ln_pre.weight"], sd["ln_pre.bias"] = ones, zeros
    for i in range(LAYERS):
        pre = f"transformer.resblocks.{i}."
        mha = nn.MultiheadAttention(WIDTH, HEADS)
        fc, pj = nn.Linear(WIDTH, 4 * WIDTH), nn.Linear(4 * WIDTH, WIDTH)
        block = {"attn.in_proj_weight": mha.in_proj_weight, "attn.in_proj_bias": mha.in_proj_bias,
                 "attn.out_proj.weight": mha.out_proj.weight, "attn.out_proj.bias": mha.out_proj.bias,
                 "mlp.c_fc.weight": fc.weight, "mlp.c_fc.bias": fc.bias,
                 "mlp.c_proj.weight": pj.weight, "mlp.c_proj.bias": pj.bias}
        for k, v in block.items():
            sd[pre + k] = v
            fp16_keys.add(pre + k)
        for ln in ("ln_1", "ln_2"):
            sd[pre + ln + ".weight"], sd[pre + ln + ".bias"] = ones, zeros
    sd["ln_post.weight"], sd["ln_post.bias"] = ones, zeros
    sd["proj"] = scale * torch.randn(WIDTH, EMBED)
    out = {}
    for k, v in sd.items():
        v = v.detach().clone().float()
        if fp16_round and k in fp16_keys:
            v = v.half().float()
        out[k] = v.contiguous()
    return out


def perturb_layernorms(sd, seed: int = 7, amount: float = 0.2):
    """Copy with non-trivial LayerNorm gains/biases (random init has gamma=1, beta=0)."""
    g = torch.Generator().manual_seed(seed)
    out = dict(sd)
    for k in sd:
        if ".ln_" in k or k.startswith("ln_"):
            if k.endswith("weight"):
                out[k] = (1.0 + amount * torch.randn(WIDTH, generator=g)).float()
            else:
                out[k] = (amount * torch.randn(WIDTH, generator=g)).float()
    return out


def synthetic_text_features(num_prompts: int = 24, seed: int = 99):
    """Well-separated unit prompt embeddings (random directions in R^512).  A random-init CLIP text
    tower collapses all prompts onto nearly one direction (logit margins ~0.04);
    these keep the zero-shot margins far above bf16 noise."""
    g = torch.Generator().manual_seed(seed)
    t = torch.randn(num_prompts, EMBED, generator=g)
    return (t / t.norm(dim=-1, keepdim=True)).float()
