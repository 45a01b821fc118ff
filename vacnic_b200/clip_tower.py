"""Frozen CLIP ViT image tower on the sm_100a kernels: `extract_clip_img_feat(clip_model, x)` of
the training / inference scripts (TRAIN:220-240, TRAINVIS:196-216, INFER) -- the step immediately upstream of the ClipCap
prefix MLP, run every step under no_grad.

`ClipVisionTower` is built from (or loaded with) the state_dict of `clip_model.visual` (OpenAI CLIP naming: conv1,
class_embedding, positional_embedding, ln_pre, transformer.resblocks.N.{ln_1, attn.in_proj_*, attn.out_proj, ln_2,
mlp.c_fc, mlp.c_proj}, ln_post); `tower(images)` returns `(x, x_cls)` exactly like the script function.  The reference
runs the tower in fp16 on the GPU (clip.load(..., device="cuda")); here: bf16 GEMM / attention operands with fp32
accumulation, the pre-LN residual stream and every LayerNorm in fp32.

Kernels: conv1 = `vacnic_vit_patchify` (fused fp32->bf16 im2col) + the tcgen05 GEMM; class token + positional embedding +
ln_pre = `vacnic_vit_embed_ln`; each block = packed [q;k;v] GEMM -> fused tcgen05 attention (197 tokens, no mask) ->
out-proj GEMM -> `vacnic_add_layernorm_fwd` (stream += branch, emits ln_2 of the new stream) -> c_fc GEMM with QuickGELU in
the epilogue -> c_proj GEMM -> add + the next block's ln_1 (or ln_post).  No torch arithmetic, no CPU fallback."""
from __future__ import annotations

from typing import Dict

import torch

from . import kernels as K
from .blocks import _heads, sdpa_fwd


class ClipVisionTower:
    def __init__(self, state_dict: Dict[str, torch.Tensor], device="cuda", patch: int = None):
        sd = {k[len("visual."):] if k.startswith("visual.") else k: v for k, v in state_dict.items()}
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise K._l.VacnicError("ClipVisionTower runs on CUDA only (no CPU fallback)")
        conv = sd["conv1.weight"]
        self.width, self.patch = conv.shape[0], conv.shape[-1] if patch is None else patch
        if self.width % 256 != 0:
            raise ValueError("ViT width must be a multiple of 256 (ViT-B: 768)")
        self.heads = self.width // 64
        self.tokens = sd["positional_embedding"].shape[0]
        self.layers = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("transformer.resblocks."))
        bf = lambda t: t.detach().to(self.device, torch.bfloat16).contiguous()   # noqa: E731  (GEMM operands)
        f32 = lambda t: t.detach().to(self.device, torch.float32).contiguous()   # noqa: E731  (LN params, biases, embeddings)
        self.conv_w = bf(conv.reshape(self.width, -1))
        self.cls, self.pos = f32(sd["class_embedding"]), f32(sd["positional_embedding"])
        self.ln_pre = (f32(sd["ln_pre.weight"]), f32(sd["ln_pre.bias"]))
        self.ln_post = (f32(sd["ln_post.weight"]), f32(sd["ln_post.bias"]))
        self.blocks = []
        for i in range(self.layers):
            p = f"transformer.resblocks.{i}."
            self.blocks.append(dict(
                ln_1=(f32(sd[p + "ln_1.weight"]), f32(sd[p + "ln_1.bias"])), ln_2=(f32(sd[p + "ln_2.weight"]), f32(sd[p + "ln_2.bias"])),
                w_in=bf(sd[p + "attn.in_proj_weight"]), b_in=f32(sd[p + "attn.in_proj_bias"]),
                w_out=bf(sd[p + "attn.out_proj.weight"]), b_out=f32(sd[p + "attn.out_proj.bias"]),
                w_fc=bf(sd[p + "mlp.c_fc.weight"]), b_fc=f32(sd[p + "mlp.c_fc.bias"]),
                w_proj=bf(sd[p + "mlp.c_proj.weight"]), b_proj=f32(sd[p + "mlp.c_proj.bias"])))

    @torch.no_grad()
    def __call__(self, images: torch.Tensor):
        """images fp32 [B, 3, H, W] (already through clip_preprocess) -> (x fp32 [B, tokens-1, w], x_cls fp32 [B, w])."""
        if not images.is_cuda:
            raise K._l.VacnicError("ClipVisionTower needs CUDA tensors (no CPU fallback)")
        B, C, Hh, Ww = images.shape
        w, H, N = self.width, self.heads, self.tokens
        if (Hh // self.patch) * (Ww // self.patch) + 1 != N:
            raise ValueError(f"image {Hh}x{Ww} with patch {self.patch} does not give {N - 1} patch tokens")
        patches = K.vit_patchify(images.contiguous().float(), self.patch)              # conv1 operand (TRAIN:225)
        tok = K.gemm(patches, self.conv_w)                                             # [B*(N-1), w]
        s = K.vit_embed_ln(tok, self.cls, self.pos, *self.ln_pre, batch=B, tokens=N)   # TRAIN:228-230, fp32 stream
        s2 = torch.empty_like(s)
        y, _, _ = K.add_layernorm_fwd(None, None, *self.blocks[0]["ln_1"], res32=s.view(B * N, w), want_stats=False)
        for i, b in enumerate(self.blocks):
            qkv = K.gemm(y, b["w_in"], bias=b["b_in"])                                 # [B*N, 3w] = [q | k | v]
            q4, k4, v4 = (_heads(qkv, B, N, H, c * w, 64) for c in range(3))
            O, _ = sdpa_fwd(q4, k4, v4, None, False, want_stats=False)
            a = K.gemm(O.view(B * N, w), b["w_out"], bias=b["b_out"])
            y, _, _ = K.add_layernorm_fwd(a, None, *b["ln_2"], res32=s.view(B * N, w), sum32_out=s2.view(B * N, w), want_stats=False)
            h = K.gemm(y, b["w_fc"], bias=b["b_fc"], act=K.ACT_QUICKGELU)
            m = K.gemm(h, b["w_proj"], bias=b["b_proj"])
            last = i + 1 == len(self.blocks)
            nxt = self.ln_post if last else self.blocks[i + 1]["ln_1"]
            out = K.add_layernorm_fwd(m, None, *nxt, res32=s2.view(B * N, w), sum32_out=s.view(B * N, w), want_stats=False,
                                      want_y32=last)
            y = out[0]
        post = out[3].view(B, N, w)                                                    # ln_post of every token, fp32
        return post[:, 1:, :].contiguous(), post[:, 0, :].contiguous()                 # TRAIN:236-239


def extract_clip_img_feat(tower: ClipVisionTower, x: torch.Tensor):
    """Drop-in for the script-level function (TRAIN:220): pass a ClipVisionTower where the script passes clip_model."""
    return tower(x)
