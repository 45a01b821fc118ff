"""Packed (varlen) article rows: device-side batch assembly that replaces the collate's right padding (DNYT:957-972,
`create_src_mask_bart` TRAIN:255-271) on the hot path -- SURVEY.md 8(f) rank 3.

The reference pads every article to the longest one and masks the padding inside attention; a quarter of all GEMM /
LayerNorm rows and attention tiles of a NYTimes800k-shaped batch is padding (lengths U{L/2..L}).  Here the article tokens
of a batch are stored back to back (`ids`, `pos_ids`), rounded up to a multiple of `ROW_BUCKET` rows so that a captured
CUDA graph can be re-used for every batch of the same bucket; the attention kernels get per-sequence row ranges
(`kernels.Packed`) instead of a padding mask.  Results on the valid tokens -- encoder memory, logits, losses and every
parameter gradient -- are those of the padded computation: pad rows never influence valid rows in the reference either
(masked keys get probability exactly 0) and their output gradient is exactly zero.

Row padding of the bucket (`tail` rows, token id = pad): appended as QUERY-ONLY rows to the last sequence -- they attend to
that sequence's keys (finite values), nobody attends to them, so their gradients are exactly zero as well.

`pack_articles` is host-side index arithmetic on the CPU batch (like the collate it replaces) and is covered by CPU tests.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

ROW_BUCKET = 512  # packed row counts are rounded up to this (one captured step graph per bucket)


def pack_articles(article_ids: torch.Tensor, pad_id: int = 1, bucket: int = ROW_BUCKET) -> Dict[str, torch.Tensor]:
    """article_ids int64 [B, L], right-padded with `pad_id` (DNYT:957-972) -> host tensors:
         ids [M] int64 (packed tokens, tail = pad_id), pos [M] int32 (position inside the article, tail = 0),
         start [B] / len [B] int32 (key ranges), qlen [B] int32 (query ranges: len, + tail on the last sequence)."""
    if article_ids.is_cuda:
        raise ValueError("pack_articles works on the host batch (it replaces the collate's padding)")
    B, L = article_ids.shape
    valid = article_ids != pad_id
    lens = valid.sum(1)
    # the collate pads on the right only: every row must be `len` valid tokens followed by padding
    if not bool((valid == (torch.arange(L)[None, :] < lens[:, None])).all()):
        raise ValueError("pack_articles: articles must be right-padded (valid tokens first), as the reference collate produces them")
    if int(lens.min()) < 1:
        raise ValueError("pack_articles: empty article")
    M = int(lens.sum())
    M_pad = (M + bucket - 1) // bucket * bucket
    ids = torch.full((M_pad,), pad_id, dtype=torch.int64)
    pos = torch.zeros(M_pad, dtype=torch.int32)
    ids[:M] = article_ids[valid]
    pos[:M] = torch.arange(L, dtype=torch.int32)[None, :].expand(B, L)[valid]
    start = torch.zeros(B, dtype=torch.int32)
    start[1:] = torch.cumsum(lens, 0)[:-1].to(torch.int32)
    ln = lens.to(torch.int32)
    qlen = ln.clone()
    qlen[-1] += M_pad - M
    return {"ids": ids, "pos": pos, "start": start, "len": ln, "qlen": qlen}


def unpack_rows(packed: torch.Tensor, start: torch.Tensor, lens: torch.Tensor, L: int, fill: float = 0.0) -> torch.Tensor:
    """[M, d] packed rows -> [B, L, d] right-padded (for API consumers / tests; never on the training path)."""
    B = start.numel()
    out = packed.new_full((B, L, packed.shape[-1]), fill)
    st, ln = start.tolist(), lens.tolist()
    for b in range(B):
        out[b, :ln[b]] = packed[st[b]:st[b] + ln[b]]
    return out


class ArticlePack:
    """Device-side view of a packed batch for one model forward: geometry objects for the three attention patterns that
    touch article rows (kernels.Packed) -- built once per forward from the static device tensors of the step."""

    def __init__(self, dev_pack: Dict[str, torch.Tensor], B: int, L: int, prefix_len: int, dec_len: int):
        from . import kernels as K
        from .blocks import PackedMask
        ids, pos, start, ln, qlen = (dev_pack[k] for k in ("ids", "pos", "start", "len", "qlen"))
        self.ids, self.pos, self.start, self.len, self.qlen = ids, pos, start, ln, qlen
        self.rows = ids.numel()
        self.B, self.L = B, L
        dev = ids.device
        tail_max = min(self.rows, ROW_BUCKET)  # rows that may be bucket padding: the last ROW_BUCKET rows at most
        # encoder self-attention: queries [start, +qlen), keys [start, +len); dK / dV of the bucket tail must be zeroed
        self.self_mask = PackedMask(K.Packed(start, qlen, start, ln, max_q=min(self.rows, L + ROW_BUCKET), max_k=L), k_tail=tail_max)
        # prefix cross-attention (encoder): packed queries over the regular [B, prefix_len] key block of each sample
        self.prefix_len = prefix_len
        if prefix_len > 0:
            pk = torch.arange(B, dtype=torch.int32, device=dev) * prefix_len
            pl = torch.full((B,), prefix_len, dtype=torch.int32, device=dev)
            self.prefix_mask = PackedMask(K.Packed(start, qlen, pk, pl, max_q=min(self.rows, L + ROW_BUCKET), max_k=prefix_len), k_tail=0)
        else:
            self.prefix_mask = None
        # decoder cross-attention: regular [B, dec_len] queries over the packed encoder memory
        dq = torch.arange(B, dtype=torch.int32, device=dev) * dec_len
        dl = torch.full((B,), dec_len, dtype=torch.int32, device=dev)
        self.cross_mask = PackedMask(K.Packed(dq, dl, start, ln, max_q=dec_len, max_k=L), k_tail=tail_max)
