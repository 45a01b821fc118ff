"""Importable loss ops with the math of the reference's script-level loss block (TRAIN:284-363), backed
by the fused sm_100a kernels.  These are what a maintainer swaps into a copy of the training script
(INTEGRATION.md §1); `trainer.TrainStep` uses the same autograd functions.

    masked_mean_pool(h, mask)                        pool()                       TRAIN:178-182
    colam_margin_loss(h, h_guide, tgt_ids, margin)   CoLaM hinge on the diagonal  TRAIN:292-309, 820
    names_embed(model, names_ids)                    get_embedding_ner            TRAIN:112-133
    secla_loss(face, names)                          BatchSoftmax                 TRAIN:631-660, 326-330
    token_ce(logits, tgt_ids, pad_id)                CrossEntropyLoss(ignore_idx) TRAIN:287, 816
"""
from __future__ import annotations

import torch

from . import blocks as Bk
from . import kernels as K


def colam_margin_loss(h: torch.Tensor, h_guide: torch.Tensor, tgt_ids: torch.Tensor, margin: float = 1.0,
                      pad_id: int = 1) -> torch.Tensor:
    """mean_b max(0, margin - cos(pool(h)_b, pool(h_guide)_b)); gradient flows into `h` only."""
    return Bk.ColamFn.apply(h.to(torch.bfloat16), h_guide.detach().to(torch.bfloat16), tgt_ids, float(margin), pad_id)[0]


def masked_mean_pool(h: torch.Tensor, tgt_ids: torch.Tensor, pad_id: int = 1) -> torch.Tensor:
    """pool(h, mask = tgt_ids != pad): fp32 [B, d] (nan -> 1.0 for all-pad rows, TRAIN:181)."""
    hb = h.detach().to(torch.bfloat16).contiguous()
    _, pa, _, _ = K.colam_fwd(hb, hb, tgt_ids.contiguous(), 1.0, pad_id)
    return pa


def names_embed(model, names_ids: torch.Tensor) -> torch.Tensor:
    """fp32 [B, N, d] = mean_t LN_ner(E_ner[ids] + Pos_ner[t + 2]) (no grad, pads included in the mean)."""
    enc = model.model.encoder
    st = model.store
    return K.names_embed(names_ids.contiguous(), st.w16(enc.embed_tokens_ner.weight), st.w16(enc.embed_positions_ner.weight),
                         enc.ln_emb_ner.g, enc.ln_emb_ner.b)


def secla_loss(face: torch.Tensor, names: torch.Tensor) -> torch.Tensor:
    """CE_rows(mean_n max_f n·f) + CE_rows(mean_f max_n f·n) against the diagonal; gradient into `face`."""
    return Bk.SeclaFn.apply(face.to(torch.bfloat16), names.detach().float())[0]


class _TokenCeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, targets, pad_id):
        V = logits.shape[-1]
        l2 = logits.reshape(-1, V)
        if l2.dtype != torch.float32 or l2.stride(1) != 1:
            l2 = l2.float().contiguous()
        t = targets.reshape(-1).contiguous()
        out, lse, _ = K.ce_fwd(l2, V, t, ignore_index=pad_id)
        ctx.saved, ctx.pad, ctx.shape = (l2, lse, t, out), pad_id, logits.shape
        return out[0]

    @staticmethod
    def backward(ctx, dloss):
        l2, lse, t, out = ctx.saved
        V = l2.shape[1]
        dl = torch.empty(l2.shape, dtype=torch.bfloat16, device=l2.device)
        K.ce_bwd(l2, V, lse, t, out, dloss.reshape(1).contiguous().float(), 1.0, dl, ignore_index=ctx.pad)
        return dl.view(ctx.shape).float(), None, None


def token_ce(logits: torch.Tensor, tgt_ids: torch.Tensor, pad_id: int = 1) -> torch.Tensor:
    return _TokenCeFn.apply(logits, tgt_ids, pad_id)
