"""fp32 restatement of the CLIP ViT image tower as the VACNIC scripts run it.  TEST INFRASTRUCTURE ONLY (same rules as
oracle/model.py: only tests/, __graft_entry__.smoke() and bench.py's baseline legs may import this).

What is restated:
  * `extract_clip_img_feat` -- TRAIN:220-240 (identical at TRAINVIS:196-216): conv1 -> flatten -> prepend class token ->
    + positional embedding -> ln_pre -> transformer -> ln_post on the CLS token (`x_cls`, what feeds the ClipCap prefix MLP,
    TRAIN:236 / MFULL:1274) and on the patch tokens (`x`), both returned as float32;
  * `clip.model.VisionTransformer` / `ResidualAttentionBlock` / `QuickGELU` of OpenAI CLIP (`clip==1.0`, vacnic.yml:223) --
    a THIRD-PARTY dependency that is not vendored under /root/reference and not installed here (no network), so its
    published architecture is restated: pre-LN blocks x = x + MHA(ln_1(x)); x = x + c_proj(QuickGELU(c_fc(ln_2(x)))),
    nn.MultiheadAttention with packed in_proj ([q; k; v]), heads = width // 64, QuickGELU(x) = x * sigmoid(1.702 x),
    LayerNorm eps 1e-5, conv1 without bias (kernel = stride = patch).

Parity status: UNPINNED against the real `clip` package (absent).  Pinned as far as this container allows:
tests/test_clip_oracle_cpu.py rebuilds the tower from torch's own nn.Conv2d / nn.MultiheadAttention / nn.LayerNorm modules in
the published layout and requires this functional restatement to match it to 1e-5, with the state_dict names of
`clip_model.visual` (so a real checkpoint's `visual.*` entries load by name)."""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


def vit_cfg(width=768, layers=12, patch=16, image=224):
    return dict(width=width, layers=layers, heads=width // 64, patch=patch, image=image, tokens=(image // patch) ** 2 + 1)


def param_shapes(cfg) -> Dict[str, tuple]:
    """state_dict of `clip_model.visual` (OpenAI naming), without the final `proj` (unused by extract_clip_img_feat)."""
    w, p, n = cfg["width"], cfg["patch"], cfg["tokens"]
    sh = {"conv1.weight": (w, 3, p, p), "class_embedding": (w,), "positional_embedding": (n, w),
          "ln_pre.weight": (w,), "ln_pre.bias": (w,), "ln_post.weight": (w,), "ln_post.bias": (w,)}
    for i in range(cfg["layers"]):
        b = f"transformer.resblocks.{i}."
        sh.update({b + "attn.in_proj_weight": (3 * w, w), b + "attn.in_proj_bias": (3 * w,), b + "attn.out_proj.weight": (w, w),
                   b + "attn.out_proj.bias": (w,), b + "ln_1.weight": (w,), b + "ln_1.bias": (w,), b + "ln_2.weight": (w,),
                   b + "ln_2.bias": (w,), b + "mlp.c_fc.weight": (4 * w, w), b + "mlp.c_fc.bias": (4 * w,),
                   b + "mlp.c_proj.weight": (w, 4 * w), b + "mlp.c_proj.bias": (w,)})
    return sh


def random_state_dict(cfg, seed: int) -> SD:
    """CLIP's own initialisation scales (clip.model.CLIP.initialize_parameters / VisionTransformer.__init__)."""
    g = torch.Generator().manual_seed(seed)
    w, L = cfg["width"], cfg["layers"]
    scale = w ** -0.5
    proj_std, attn_std, fc_std = (w ** -0.5) * ((2 * L) ** -0.5), w ** -0.5, (2 * w) ** -0.5
    sd = {}
    for k, shp in param_shapes(cfg).items():
        if "ln_" in k:
            sd[k] = torch.ones(shp) + 0.05 * torch.randn(shp, generator=g) if k.endswith("weight") else 0.05 * torch.randn(shp, generator=g)
        elif k.endswith("bias"):
            sd[k] = 0.02 * torch.randn(shp, generator=g)
        elif k in ("class_embedding", "positional_embedding"):
            sd[k] = scale * torch.randn(shp, generator=g)
        elif k == "conv1.weight":
            sd[k] = torch.randn(shp, generator=g) * (3 * cfg["patch"] ** 2) ** -0.5
        elif "in_proj_weight" in k:
            sd[k] = attn_std * torch.randn(shp, generator=g)
        elif "out_proj" in k or "c_proj" in k:
            sd[k] = proj_std * torch.randn(shp, generator=g)
        else:
            sd[k] = fc_std * torch.randn(shp, generator=g)
    return sd


def quick_gelu(x):
    return x * torch.sigmoid(1.702 * x)


def resblock(sd: SD, p: str, heads: int, x):
    """ResidualAttentionBlock.forward on [B, N, w] (the reference permutes to LND for nn.MultiheadAttention; same math)."""
    B, N, w = x.shape
    hd = w // heads
    h = F.layer_norm(x, (w,), sd[p + "ln_1.weight"], sd[p + "ln_1.bias"], 1e-5)
    qkv = F.linear(h, sd[p + "attn.in_proj_weight"], sd[p + "attn.in_proj_bias"])
    q, k, v = (t.reshape(B, N, heads, hd).transpose(1, 2) for t in qkv.chunk(3, dim=-1))
    att = torch.softmax((q * hd ** -0.5) @ k.transpose(-1, -2), dim=-1) @ v
    x = x + F.linear(att.transpose(1, 2).reshape(B, N, w), sd[p + "attn.out_proj.weight"], sd[p + "attn.out_proj.bias"])
    h = F.layer_norm(x, (w,), sd[p + "ln_2.weight"], sd[p + "ln_2.bias"], 1e-5)
    h = F.linear(quick_gelu(F.linear(h, sd[p + "mlp.c_fc.weight"], sd[p + "mlp.c_fc.bias"])), sd[p + "mlp.c_proj.weight"],
                 sd[p + "mlp.c_proj.bias"])
    return x + h


def extract_clip_img_feat(sd: SD, cfg, images):
    """TRAIN:220-240.  images [B, 3, H, W] -> (x [B, tokens-1, w], x_cls [B, w]), float32."""
    w = cfg["width"]
    x = F.conv2d(images, sd["conv1.weight"], stride=cfg["patch"])                    # TRAIN:225
    x = x.reshape(x.shape[0], x.shape[1], -1).permute(0, 2, 1)                        # TRAIN:226-227
    cls = sd["class_embedding"].to(x.dtype) + torch.zeros(x.shape[0], 1, w, dtype=x.dtype, device=x.device)
    x = torch.cat([cls, x], dim=1) + sd["positional_embedding"].to(x.dtype)           # TRAIN:228-229
    x = F.layer_norm(x, (w,), sd["ln_pre.weight"], sd["ln_pre.bias"], 1e-5)           # TRAIN:230
    for i in range(cfg["layers"]):                                                    # TRAIN:232-234
        x = resblock(sd, f"transformer.resblocks.{i}.", cfg["heads"], x)
    x_cls = F.layer_norm(x[:, 0, :], (w,), sd["ln_post.weight"], sd["ln_post.bias"], 1e-5).float()   # TRAIN:236-237
    x = F.layer_norm(x[:, 1:, :], (w,), sd["ln_post.weight"], sd["ln_post.bias"], 1e-5).float()      # TRAIN:238-239
    return x, x_cls
