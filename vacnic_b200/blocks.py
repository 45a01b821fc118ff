"""Block-level autograd functions of the VACNIC encoder / decoder, built only from the C-ABI kernels.

Granularity = one residual block of the reference (`x -> LN(x + dropout(f(x)))`), so that the
gradient fan-in of the residual stream is fused into kernel epilogues (LayerNorm-backward writes the
residual gradient, the data-gradient GEMM accumulates onto it) instead of being summed by autograd.
Parameter gradients are written straight into the flat fp32 gradient buffer of the ParamStore
(`Lin.gw / Lin.gb`), never returned through autograd.

Reference: BartAttention.forward MFULL:454-565, BartEncoderLayer.forward MFULL:618-762,
BartDecoderLayer.forward MFULL:793-890, embeddings MFULL:1243-1260, 1555-1563.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import kernels as K
from .store import Lin, ParamStore


class Runtime:
    """Per-model execution context shared by all blocks."""

    def __init__(self, store: ParamStore, p_drop: float = 0.0, seed: int = 0, p_attn: float = 0.0, p_act: float = 0.0):
        self.store = store
        self.p_drop = p_drop      # config.dropout: embeddings, attention / FFN outputs (MFULL:705, 721, 742, 1249 ...)
        self.p_attn = p_attn      # config.attention_dropout: on the probabilities, after the softmax (MFULL:546)
        self.p_act = p_act        # config.activation_dropout: after the FFN activation (MFULL:649, 660, 684, 740, 874)
        self.training = False
        self.rng = K.Rng(store.device, seed)
        self._salt = 0
        self.grad_hook = None
        # residual stream in fp32: every LayerNorm also emits its output in fp32 and the next block adds its branch to THAT
        # (what the reference does under torch.autocast, where LayerNorm runs and returns fp32); the GEMMs consume the
        # bf16 copy.  Forward only -- the backward pass keeps using the bf16 copy to rebuild x_hat.
        self.res_fp32 = True
        # requires-grad root every block of a forward pass hangs off (store.anchor, or its ParamTouchFn image under DDP)
        self.fwd_anchor = store.anchor
        # optional second stream for the prefix side of the encoder (set by the owner of the step, who also joins it after
        # the backward pass and then clears `keepalive`: the tensors that crossed the two streams in this step)
        self.side_stream = None
        self.keepalive = []

    def new_salt(self) -> int:
        self._salt += 1
        return self._salt

    @property
    def drop(self) -> float:
        return self.p_drop if self.training else 0.0

    @property
    def attn_drop(self) -> float:
        return self.p_attn if self.training else 0.0

    @property
    def act_drop(self) -> float:
        return self.p_act if self.training else 0.0


class LN:
    """fp32 LayerNorm parameter views + gradient views + dropout salt of one call site."""
    __slots__ = ("g", "b", "gg", "gb", "salt")

    def __init__(self, store: ParamStore, mod, salt: int):
        self.g, self.b = store.f32(mod.weight).view(-1), store.f32(mod.bias).view(-1)
        self.gg = None if store.frozen else store.g32(mod.weight).view(-1)
        self.gb = None if store.frozen else store.g32(mod.bias).view(-1)
        self.salt = salt


def _ceil8(n: int) -> int:
    return (n + 7) // 8 * 8


def wgrad_k_slices(rows: int, n_out: int, k_in: int, sms: int = 148) -> int:
    """How many slices of the token dimension a weight gradient gw[n_out, k_in] = dy[rows, n_out]^T x[rows, k_in] is
    cut into.  A long reduction (rows = B*L) with few 256 x 256 output tiles (the 1024 x 1024 projections: 16 tiles for
    74 SM pairs) leaves most of the GPU idle; the slices run as the batch dimension of ONE CTA-pair GEMM into fp32
    partials that `vacnic_sum_partials` adds in slice order (deterministic).  1 = no slicing."""
    if rows < 8192 or n_out < 256 or k_in < 256:
        return 1
    tiles = ((n_out + 255) // 256) * ((k_in + 255) // 256)
    pairs = sms // 2
    if 2 * tiles > pairs:
        return 1
    s = min(pairs // tiles, 8)
    while s > 1 and (rows % s != 0 or rows // s < 2048 or (rows // s) % 8 != 0):
        s -= 1
    return s


def _wgrad(rt: Runtime, lin: Lin, dy2d: torch.Tensor, x2d: torch.Tensor, bias_from: Optional[torch.Tensor] = None):
    """gw (+)= dy^T x ; gb += colsum(dy)."""
    rows, n_out, k_in = dy2d.shape[0], dy2d.shape[1], x2d.shape[1]
    s = wgrad_k_slices(rows, n_out, k_in, K.sm_count(dy2d.device))
    if s > 1:
        part = torch.empty(s, n_out, k_in, dtype=torch.float32, device=dy2d.device)
        K.gemm(dy2d.view(s, rows // s, n_out), x2d.view(s, rows // s, k_in), out=part, a_mn=True, b_mn=True)
        K.sum_partials(part.view(s, n_out * k_in), lin.gw, accumulate=rt.store.touch(lin.key))
    else:
        K.gemm(dy2d, x2d, out=lin.gw, a_mn=True, b_mn=True, accumulate=rt.store.touch(lin.key))
    if lin.gb is not None and bias_from is not None:
        K.colsum_into(bias_from, lin.gb)


# --------------------------------------------------------------------------------------------------
# scaled dot-product attention core: fused tcgen05 kernels (scores and probabilities never reach HBM)
# --------------------------------------------------------------------------------------------------
class KeyMask:
    """uint8 key-padding mask [B, Sk] (1 = attend) plus the per-row count of leading keys that can be
    attended at all (keys beyond it are skipped by the attention kernels)."""
    __slots__ = ("mask", "len")

    def __init__(self, mask_like: torch.Tensor):
        self.mask = mask_like.to(torch.uint8).contiguous()
        self.len = K.mask_key_len(self.mask)


class PackedMask:
    """Takes the place of a KeyMask when the rows of a batch are PACKED (varlen): `geo` (kernels.Packed) says which query
    and key rows each sequence owns; `k_tail` = number of trailing rows of the key-side buffers that belong to no
    sequence's key range (the row padding of the packed batch) -- their dK / dV are never written by the kernels and must
    read as zero for the weight-gradient GEMMs."""
    __slots__ = ("geo", "k_tail")

    def __init__(self, geo, k_tail: int = 0):
        self.geo, self.k_tail = geo, int(k_tail)


def sdpa_fwd(q4, k4, v4, key_mask, causal, want_stats=True, drop=(0.0, None, 0)):
    """q4 [B,H,Sq,hd], k4/v4 [B,H,Sk,hd] (strided views).  Returns (O [B,Sq,H*hd], stats [B,H,Sq,2]).
    `drop` = (p, Rng, salt) of the attention dropout (p = 0: none)."""
    dk = dict(p_drop=drop[0], rng=drop[1], salt=drop[2])
    if isinstance(key_mask, PackedMask):
        return K.attn_fwd(q4, k4, v4, causal=causal, want_stats=want_stats, packed=key_mask.geo, **dk)
    if key_mask is not None and tuple(key_mask.mask.shape) != (q4.shape[0], k4.shape[2]):
        raise ValueError(f"Attention mask should be of size {(q4.shape[0], 1, q4.shape[2], k4.shape[2])}, but is "
                         f"{(key_mask.mask.shape[0], 1, q4.shape[2], key_mask.mask.shape[1])}")  # MFULL:516-520
    return K.attn_fwd(q4, k4, v4, key_mask.mask if key_mask is not None else None,
                      key_mask.len if key_mask is not None else None, causal, want_stats=want_stats, **dk)


def sdpa_bwd(dO, O, stats, q4, k4, v4, key_mask, causal, dq4, dk4, dv4, drop=(0.0, None, 0)):
    dk = dict(p_drop=drop[0], rng=drop[1], salt=drop[2])
    if isinstance(key_mask, PackedMask):
        return K.attn_bwd(dO, O, stats, q4, k4, v4, dq4, dk4, dv4, causal=causal, packed=key_mask.geo, **dk)
    K.attn_bwd(dO, O, stats, q4, k4, v4, dq4, dk4, dv4, key_mask.mask if key_mask is not None else None,
               key_mask.len if key_mask is not None else None, causal, **dk)


def _heads(t2d: torch.Tensor, B: int, S: int, H: int, col0: int, hd: int) -> torch.Tensor:
    """[B*S, ld] buffer -> [B,H,S,hd] view of columns [col0, col0 + H*hd)."""
    ld = t2d.stride(0)
    return t2d.as_strided((B, H, S, hd), (S * ld, hd, ld, 1), t2d.storage_offset() + col0)


class AttnBlockFn(torch.autograd.Function):
    """y = LN(x + dropout(out_proj(softmax(q k^T * hd^-0.5 + mask) v)))   (MFULL:697-707, 711-723,
    669-679, 831-841, 856-866).  Self-attention: kv_src is None and one fused [k;v;q] GEMM projects x.
    Cross-attention: q from x; k/v either from `kv_src` through `lin_kv`, or precomputed (`kv_pre`,
    the hoisted decoder cross-K/V GEMM) with gradients written to `dkv_pre`."""

    @staticmethod
    def forward(ctx, x, x32, kv_src, kv_pre, rt: Runtime, lin_qkv: Lin, lin_q: Lin, lin_kv: Lin, lin_o: Lin, ln: LN,
                H: int, key_mask, causal: bool, use_dropout: bool, kv_col0: int, dkv_all, return_dkv_all: bool):
        """x: bf16 [B, Sq, d]; x32: the same state in fp32 (or None).  Returns (y bf16, y32 fp32 or None)."""
        B, Sq, d = x.shape
        hd = d // H
        x2 = x.view(B * Sq, d)
        # packed (varlen) rows: the attention kernels see ONE [rows, H*64] matrix per operand + per-sequence row ranges
        packed = isinstance(key_mask, PackedMask)
        Bq, Rq = (1, B * Sq) if packed else (B, Sq)
        if lin_qkv is not None:  # self attention, fused [k; v; q] projection (MFULL:444-446 order)
            qkv = K.gemm(x2, lin_qkv.w16, bias=lin_qkv.b32)
            Bk, Sk = Bq, Rq
            k4, v4, q4 = (_heads(qkv, Bq, Rq, H, c * d, hd) for c in range(3))
            kvp = None
        else:
            qkv = K.gemm(x2, lin_q.w16, bias=lin_q.b32)
            q4 = _heads(qkv, Bq, Rq, H, 0, hd)
            if kv_pre is not None:  # [Bk, Sk, n_layers*2d]: this layer's k|v live at columns [kv_col0, +2d)
                Bk, Sk = kv_pre.shape[0], kv_pre.shape[1]
                kvp = kv_pre.view(Bk * Sk, -1)[:, kv_col0:kv_col0 + 2 * d]
            else:
                Bk, Sk = kv_src.shape[0], kv_src.shape[1]
                kvp = K.gemm(kv_src.view(Bk * Sk, d), lin_kv.w16, bias=lin_kv.b32)
            if packed:
                Bk, Sk = 1, Bk * Sk
            k4, v4 = _heads(kvp, Bk, Sk, H, 0, hd), _heads(kvp, Bk, Sk, H, d, hd)
        adrop = (rt.attn_drop, rt.rng, ln.salt + 4096)  # attention dropout on the probabilities (MFULL:546)
        O, stats = sdpa_fwd(q4, k4, v4, key_mask, causal, want_stats=not rt.store.frozen, drop=adrop)
        a = K.gemm(O.view(B * Sq, d), lin_o.w16, bias=lin_o.b32)
        p = rt.drop if use_dropout else 0.0
        y32 = None
        if rt.res_fp32:
            y, mean, rstd, y32 = K.add_layernorm_fwd(a, x2, ln.g, ln.b, p_drop=p, rng=rt.rng, salt=ln.salt, want_y32=True,
                                                     res32=None if x32 is None else x32.view(B * Sq, d))
            y32 = y32.view(B, Sq, d)
            ctx.mark_non_differentiable(y32)
        else:
            y, mean, rstd = K.add_layernorm_fwd(a, x2, ln.g, ln.b, p_drop=p, rng=rt.rng, salt=ln.salt)
        ctx.rt, ctx.lins, ctx.ln, ctx.H, ctx.p = rt, (lin_qkv, lin_q, lin_kv, lin_o), ln, H, p
        ctx.saved = (x2, kv_src, qkv, kvp, stats, O, a, mean, rstd, q4, k4, v4)
        ctx.mask = (key_mask, causal)
        ctx.adrop = adrop
        hoisted = kv_pre is not None and dkv_all is not None
        ctx.dkv_pre = dkv_all.view(Bk * Sk, -1)[:, kv_col0:kv_col0 + 2 * d] if hoisted else None
        ctx.dkv_all = dkv_all if (hoisted and return_dkv_all) else None
        ctx.shape = (B, Sq, d, Sk, Bq, Rq, Bk)
        return y.view(B, Sq, d), y32

    @staticmethod
    def backward(ctx, dy, _dy32=None):
        rt, ln, H = ctx.rt, ctx.ln, ctx.H
        lin_qkv, lin_q, lin_kv, lin_o = ctx.lins
        x2, kv_src, qkv, kvp, stats, O, a, mean, rstd, q4, k4, v4 = ctx.saved
        key_mask, causal = ctx.mask
        B, Sq, d, Sk, Bq, Rq, Bk = ctx.shape
        hd = d // H
        k_tail = key_mask.k_tail if isinstance(key_mask, PackedMask) else 0
        dy = dy.contiguous()
        dsum, da = K.add_layernorm_bwd(dy.view(B * Sq, d), a, x2, ln.g, mean, rstd, ln.gg, ln.gb, dbias=lin_o.gb,
                                       want_dx=True, p_drop=ctx.p, rng=rt.rng, salt=ln.salt)
        O2 = O.view(B * Sq, d)
        _wgrad(rt, lin_o, da, O2)
        dO = K.gemm(da, lin_o.w16, b_mn=True).view(Bq, Rq, d)
        dkv_src = None
        if lin_qkv is not None:
            dqkv = torch.empty_like(qkv)
            if k_tail:  # key rows no sequence owns: the kernels never write their dK / dV (their dQ is written afterwards)
                dqkv[-k_tail:].zero_()
            dk4, dv4, dq4 = (_heads(dqkv, Bq, Rq, H, c * d, hd) for c in range(3))
            sdpa_bwd(dO, O, stats, q4, k4, v4, key_mask, causal, dq4, dk4, dv4, drop=ctx.adrop)
            _wgrad(rt, lin_qkv, dqkv, x2, bias_from=dqkv)
            K.gemm(dqkv, lin_qkv.w16, out=dsum, b_mn=True, accumulate=True)  # dx = dsum + dqkv W
        else:
            dq = torch.empty_like(qkv)
            dq4 = _heads(dq, Bq, Rq, H, 0, hd)
            if ctx.dkv_pre is not None:
                dkvp = ctx.dkv_pre  # (its unowned rows were zeroed where the shared buffer was allocated)
            else:
                dkvp = torch.empty_like(kvp)
                if k_tail:
                    dkvp[-k_tail:].zero_()
            dk4, dv4 = _heads(dkvp, Bk, Sk, H, 0, hd), _heads(dkvp, Bk, Sk, H, d, hd)
            sdpa_bwd(dO, O, stats, q4, k4, v4, key_mask, causal, dq4, dk4, dv4, drop=ctx.adrop)
            _wgrad(rt, lin_q, dq, x2, bias_from=dq)
            K.gemm(dq, lin_q.w16, out=dsum, b_mn=True, accumulate=True)
            if ctx.dkv_pre is None:
                kv2 = kv_src.reshape(Bk * Sk, d)
                _wgrad(rt, lin_kv, dkvp, kv2, bias_from=dkvp)
                if ctx.needs_input_grad[2]:  # kv_src (inputs: x, x32, kv_src, kv_pre, ...)
                    dkv_src = K.gemm(dkvp, lin_kv.w16, b_mn=True).view(kv_src.shape)
        ctx.saved = None
        # hoisted cross K/V: every layer wrote its slice of the shared gradient buffer in place; the
        # block that runs last in the backward pass (layer 0) hands the whole buffer to autograd.
        return (dsum.view(B, Sq, d), None, dkv_src, ctx.dkv_all) + (None,) * 13


class MlpBlockFn(torch.autograd.Function):
    """z = act(x W1^T + b1) W2^T + b2, then y = LN(x + dropout(z)) when `ln` is given (the FFN blocks,
    MFULL:647-653, 658-664, 738-744, 868-878) or y = z (the ClipCap prefix MLP, MFULL:111-123)."""

    @staticmethod
    def forward(ctx, x, x32, anchor, rt: Runtime, lin1: Lin, lin2: Lin, act: int, ln: Optional[LN]):
        """Returns (y bf16, y32 fp32 or None); x32 = fp32 copy of x (residual stream) or None."""
        shp = x.shape
        d_in = shp[-1]
        x2 = x.reshape(-1, d_in)
        rows = x2.shape[0]
        aux = torch.empty(rows, lin1.out_f, dtype=torch.bfloat16, device=x.device) if act == K.ACT_GELU else None
        h = K.gemm(x2, lin1.w16, bias=lin1.b32, act=act, aux_out=aux)
        p_act = rt.act_drop if ln is not None else 0.0  # activation dropout of the FFN blocks (not of the ClipCap prefix MLP)
        if p_act > 0.0:
            K.dropout_inplace(h, p_act, rt.rng, ln.salt + 8192)
        z = K.gemm(h, lin2.w16, bias=lin2.b32)
        y32 = None
        if ln is not None:
            p = rt.drop
            if rt.res_fp32:
                y, mean, rstd, y32 = K.add_layernorm_fwd(z, x2, ln.g, ln.b, p_drop=p, rng=rt.rng, salt=ln.salt, want_y32=True,
                                                         res32=None if x32 is None else x32.reshape(-1, d_in))
                y32 = y32.view(shp)
                ctx.mark_non_differentiable(y32)
            else:
                y, mean, rstd = K.add_layernorm_fwd(z, x2, ln.g, ln.b, p_drop=p, rng=rt.rng, salt=ln.salt)
            ctx.lnstate = (z, mean, rstd, p)
            out = y.view(shp)
        else:
            ctx.lnstate = None
            out = z.view(shp[:-1] + (lin2.out_f,))
        ctx.rt, ctx.lin1, ctx.lin2, ctx.act, ctx.ln, ctx.p_act = rt, lin1, lin2, act, ln, p_act
        ctx.saved = (x2, aux, h)
        ctx.in_shape = shp
        return out, y32

    @staticmethod
    def backward(ctx, dy, _dy32=None):
        rt, lin1, lin2, act, ln = ctx.rt, ctx.lin1, ctx.lin2, ctx.act, ctx.ln
        x2, aux, h = ctx.saved
        rows = x2.shape[0]
        if ln is not None:
            z, mean, rstd, p = ctx.lnstate
            dsum, dz = K.add_layernorm_bwd(dy.contiguous(), z, x2, ln.g, mean, rstd, ln.gg, ln.gb, dbias=lin2.gb,
                                           want_dx=True, p_drop=p, rng=rt.rng, salt=ln.salt)
        else:
            dz = dy.contiguous().view(rows, lin2.out_f)
            dsum = None
            if lin2.gb is not None:
                K.colsum_into(dz, lin2.gb)
        _wgrad(rt, lin2, dz, h)
        # dh * act'(.) fused in the data-gradient GEMM epilogue (GELU' from the pre-activation, tanh' from the output)
        dpre = K.gemm(dz, lin2.w16, b_mn=True, dact=act, aux_in=aux if act == K.ACT_GELU else h)
        if ctx.p_act > 0.0:  # the mask of the forward pass (elementwise, commutes with act')
            K.dropout_inplace(dpre, ctx.p_act, rt.rng, ln.salt + 8192)
        _wgrad(rt, lin1, dpre, x2, bias_from=dpre)
        if dsum is not None:
            K.gemm(dpre, lin1.w16, out=dsum, b_mn=True, accumulate=True)
            dx = dsum.view(ctx.in_shape)
        elif ctx.needs_input_grad[0]:
            dx = K.gemm(dpre, lin1.w16, b_mn=True).view(ctx.in_shape)
        else:
            dx = None
        ctx.saved = None
        return (dx,) + (None,) * 7


class LinearFn(torch.autograd.Function):
    """y = x W^T + b (visual_map MFULL:1277-1278, face _linear_1 MFULL:1269, hoisted decoder cross K/V,
    LM head MFULL:1997 with fp32 output)."""

    @staticmethod
    def forward(ctx, x, anchor, rt: Runtime, lin: Lin, out_dtype, need_dx: bool, out_buf, dy_buf):
        shp = x.shape
        x2 = x.reshape(-1, shp[-1])
        if out_buf is not None:
            y = K.gemm(x2, lin.w16, out=out_buf, bias=lin.b32)
        else:
            y = K.gemm(x2, lin.w16, bias=lin.b32, out_dtype=out_dtype)
        ctx.rt, ctx.lin, ctx.x2, ctx.need_dx, ctx.shp, ctx.dy_buf = rt, lin, x2, need_dx, shp, dy_buf
        return y.view(shp[:-1] + (y.shape[-1],))

    @staticmethod
    def backward(ctx, dy):
        rt, lin, x2 = ctx.rt, ctx.lin, ctx.x2
        if ctx.dy_buf is not None:
            dy2 = ctx.dy_buf  # gradients were written in place by the consumers (hoisted cross K/V)
        else:
            dy2 = dy.reshape(x2.shape[0], -1)
            if dy2.dtype != torch.bfloat16:
                # e.g. the fp32 gradient of a script-level torch loss on the logits (TRAIN:287): bf16, TMA-friendly pitch
                dy2 = K.cast_rows_bf16(dy2.float() if dy2.dtype != torch.float32 else (dy2 if dy2.stride(-1) == 1 else dy2.contiguous()))
            elif dy2.stride(-1) != 1 or (dy2.stride(0) * 2) % 16 != 0:
                dy2 = K.pad_rows(dy2.contiguous(), (dy2.shape[1] + 7) // 8 * 8)[:, :dy2.shape[1]]
        _wgrad(rt, lin, dy2, x2, bias_from=dy2 if lin.gb is not None else None)
        dx = K.gemm(dy2, lin.w16, b_mn=True).view(ctx.shp) if (ctx.need_dx and ctx.needs_input_grad[0]) else None
        return (dx,) + (None,) * 7


class EmbedFn(torch.autograd.Function):
    """y = dropout(LN(tok[ids] + pos[t + offset]))  (MFULL:1243-1249, 1254-1260, 1555-1563)."""

    @staticmethod
    def forward(ctx, anchor, ids, rt: Runtime, tok_p, pos_p, ln: LN, pos_offset: int, pad_id: int, pos_ids=None):
        """Returns (y bf16, y32 fp32 or None).  `pos_ids` (int32, one per id): explicit positions for packed rows."""
        st = rt.store
        tok16, pos16 = st.w16(tok_p), st.w16(pos_p)
        p = rt.drop
        y32 = None
        if rt.res_fp32:
            y, mean, rstd, y32 = K.embed_ln_fwd(ids, tok16, pos16, ln.g, ln.b, pos_offset=pos_offset, p_drop=p, rng=rt.rng,
                                                salt=ln.salt, want_y32=True, pos_ids=pos_ids)
            ctx.mark_non_differentiable(y32)
        else:
            y, mean, rstd = K.embed_ln_fwd(ids, tok16, pos16, ln.g, ln.b, pos_offset=pos_offset, p_drop=p, rng=rt.rng,
                                           salt=ln.salt, pos_ids=pos_ids)
        ctx.rt, ctx.ln, ctx.args = rt, ln, (ids, tok_p, pos_p, tok16, pos16, mean, rstd, pos_offset, pad_id, p, pos_ids)
        return y, y32

    @staticmethod
    def backward(ctx, dy, _dy32=None):
        rt, ln = ctx.rt, ctx.ln
        ids, tok_p, pos_p, tok16, pos16, mean, rstd, pos_offset, pad_id, p, pos_ids = ctx.args
        st = rt.store
        K.embed_ln_bwd(dy.contiguous(), ids, tok16, pos16, ln.g, mean, rstd, st.g32(tok_p), st.g32(pos_p), ln.gg, ln.gb,
                       pos_offset=pos_offset, pad_id=pad_id, p_drop=p, rng=rt.rng, salt=ln.salt, pos_ids=pos_ids)
        return (None,) * 9


class NerMapFn(torch.autograd.Function):
    """prefix = LN(reshape(gelu(reshape(ner) W_up^T + b) W_down^T + b))  (MFULL:682-688): [B,E,d] -> [B,G,d].

    The reference's reshape(B, d, E) is a memory REINTERPRETATION of the contiguous [B, E, d] buffer, so the two
    linears are plain GEMMs over R = B*d rows of E = 80 contiguous elements (K = 80, N = 80 / 20): they run on the
    tcgen05 GEMM (TMA zero-fills the ragged K / N edges).  Only the 20-wide tensors (40-byte rows) need a re-pitch
    before TMA can read them back in the backward pass."""
    SPLIT = 64  # weight gradients: R rows are reduced in SPLIT independent slabs, then summed (deterministic)

    @staticmethod
    def forward(ctx, ner, rt: Runtime, up: Lin, down: Lin, ln: LN):
        B, E, d = ner.shape
        G = down.out_f
        x_rows = ner.reshape(B * d, E)
        z1 = torch.empty(B * d, up.out_f, dtype=torch.bfloat16, device=ner.device)
        y1 = K.gemm(x_rows, up.w16, bias=up.b32, act=K.ACT_GELU, aux_out=z1)
        p_act = rt.act_drop  # MFULL:684
        if p_act > 0.0:
            K.dropout_inplace(y1, p_act, rt.rng, ln.salt + 8192)
        z2 = K.gemm(y1, down.w16, bias=down.b32)
        z2r = z2.view(B * G, d)
        # nn.functional.dropout(hidden_states_ner_prefix, p=self.dropout, training=self.training) sits between
        # ner_map_down and the reshape + ner_map_layer_norm (MFULL:685-687); elementwise, so it commutes with the reshape
        p = rt.drop
        y, mean, rstd = K.add_layernorm_fwd(z2r, None, ln.g, ln.b, p_drop=p, rng=rt.rng, salt=ln.salt)
        ctx.rt, ctx.up, ctx.down, ctx.ln, ctx.p, ctx.p_act = rt, up, down, ln, p, p_act
        ctx.saved = (x_rows, z1, y1, z2r, mean, rstd)
        ctx.dims = (B, E, d, G)
        return y.view(B, G, d)

    @staticmethod
    def backward(ctx, dy):
        rt, up, down, ln = ctx.rt, ctx.up, ctx.down, ctx.ln
        x_rows, z1, y1, z2r, mean, rstd = ctx.saved
        B, E, d, G = ctx.dims
        R, U = B * d, up.out_f
        _, dz2 = K.add_layernorm_bwd(dy.contiguous().view(B * G, d), z2r, None, ln.g, mean, rstd, ln.gg, ln.gb,
                                     want_dx=True, p_drop=ctx.p, rng=rt.rng, salt=ln.salt)  # dz2 = dropout-masked branch gradient
        Gp = (G + 7) // 8 * 8
        dz2p = K.pad_rows(dz2.view(R, G), Gp)[:, :G]  # 16-byte row pitch for TMA
        # dz1 = (dz2 W_down) * gelu'(z1) ; dx = dz1 W_up
        dz1 = K.gemm(dz2p, down.w16, b_mn=True, dact=K.ACT_GELU, aux_in=z1)
        if ctx.p_act > 0.0:
            K.dropout_inplace(dz1, ctx.p_act, rt.rng, ln.salt + 8192)
        dx = K.gemm(dz1, up.w16, b_mn=True)
        # weight gradients: [S, R/S, .]^T [S, R/S, .] partial products, then a fixed-order sum over the S slabs
        S = NerMapFn.SPLIT if R % NerMapFn.SPLIT == 0 and (R // NerMapFn.SPLIT) % 8 == 0 else 1
        part_up = torch.empty(S, U, E, dtype=torch.float32, device=dy.device)
        K.gemm(dz1.view(S, R // S, U), x_rows.view(S, R // S, E), out=part_up, a_mn=True, b_mn=True)
        K.sum_partials(part_up.view(S, U * E), up.gw, accumulate=rt.store.touch(up.key))
        part_dn = torch.empty(S, G, U, dtype=torch.float32, device=dy.device)
        K.gemm(dz2p.unflatten(0, (S, R // S)), y1.view(S, R // S, U), out=part_dn, a_mn=True, b_mn=True)
        K.sum_partials(part_dn.view(S, G * U), down.gw, accumulate=rt.store.touch(down.key))
        K.colsum_into(dz1, up.gb)
        K.colsum_into(dz2p, down.gb)
        ctx.saved = None
        return dx.view(B, E, d), None, None, None, None


class Concat2Fn(torch.autograd.Function):
    """torch.cat((a, b), dim=1) for [B, Sa, d] / [B, Sb, d] (MFULL:668, 691): pure data movement."""

    @staticmethod
    def forward(ctx, a, b):
        ctx.sa = a.shape[1]
        out = torch.empty(a.shape[0], a.shape[1] + b.shape[1], a.shape[2], dtype=a.dtype, device=a.device)
        K.concat_rows(a, b, out)
        return out

    @staticmethod
    def backward(ctx, d):
        return d[:, :ctx.sa], d[:, ctx.sa:]


class FanoutFn(torch.autograd.Function):
    """Identity with n outputs whose gradients are summed by the vacnic_add_bf16 kernel."""

    @staticmethod
    def forward(ctx, x, n: int):
        ctx.set_materialize_grads(False)
        return tuple(x.view_as(x) for _ in range(n))

    @staticmethod
    def backward(ctx, *grads):
        gs = [g.contiguous() for g in grads if g is not None]
        if not gs:
            return None, None
        acc = gs[0]
        i = 1
        while i < len(gs):
            if i + 1 < len(gs):
                acc = K.add_bf16(acc, gs[i], gs[i + 1])
                i += 2
            else:
                acc = K.add_bf16(acc, gs[i])
                i += 1
        return acc, None


class ParamTouchFn(torch.autograd.Function):
    """Identity on the anchor that makes every parameter an autograd ancestor of the forward outputs.

    Parameter gradients are written by kernels straight into the flat gradient buffer (`p.grad` are views of it), so
    autograd itself never delivers a gradient to a parameter and the hooks `DistributedDataParallel` (TRAIN:86-87,
    TRAINVIS:84) hangs on each parameter's AccumulateGrad node would never fire.  Every block takes the anchor as an
    input, hence this node is the LAST one of the backward pass: when it runs all kernels that write gradients have been
    enqueued.  It returns no gradient for the parameters (nothing is added to `p.grad`), but their AccumulateGrad nodes
    still execute, DDP's hooks run, find the finished local gradient in `p.grad`, bucket it, all-reduce it and copy the
    average back IN PLACE (the views survive).  `find_unused_parameters=True` also works: all parameters are reachable."""

    @staticmethod
    def forward(ctx, anchor, store, *params):
        ctx.n, ctx.store = len(params), store
        return anchor.view_as(anchor)

    @staticmethod
    def backward(ctx, g):
        ctx.store.finish_backward()  # end of the backward pass of the plain loop: unused GEMM weights read as zero
        return (None,) * (2 + ctx.n)


class GradMarkFn(torch.autograd.Function):
    """Identity whose backward calls `hook(tag)`: it sits on the residual stream at a layer boundary, so when its
    backward runs every parameter gradient of the layers after the boundary is final and their data-parallel
    all-reduce can start while the rest of the backward pass continues (vacnic_b200.dp / trainer)."""

    @staticmethod
    def forward(ctx, x, hook, tag):
        ctx.hook, ctx.tag = hook, tag
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        ctx.hook(ctx.tag)
        return g, None, None


def grad_mark(x, rt: Runtime, tag):
    hook = getattr(rt, "grad_hook", None)
    if hook is None or not x.requires_grad:
        return x
    return GradMarkFn.apply(x, hook, tag)


def fanout(x, n):
    if n == 1 or not x.requires_grad:
        return (x,) * n
    return FanoutFn.apply(x, n)


# --------------------------------------------------------------------------------------------------
# heads and losses
# --------------------------------------------------------------------------------------------------
def _pad8(n: int) -> int:
    return (n + 7) // 8 * 8


class LmHeadCeFn(torch.autograd.Function):
    """logits = h W_lm^T + final_logits_bias (MFULL:1997) fused with CrossEntropyLoss(ignore_index=pad)
    (TRAIN:287, 816).  Returns (loss[1], logits fp32 [B,T,V]); only the loss is differentiable here."""

    @staticmethod
    def forward(ctx, h, rt: Runtime, lin: Lin, targets, pad_id: int):
        ctx.set_materialize_grads(False)
        B, T, d = h.shape
        V = lin.out_f
        ldp = _pad8(V)
        h2 = h.reshape(B * T, d)
        buf = torch.empty(B * T, ldp, dtype=torch.float32, device=h.device)
        logits = K.gemm(h2, lin.w16, out=buf[:, :V], bias=lin.b32)
        tflat = targets.reshape(-1).contiguous()
        out, lse, _ = K.ce_fwd(logits, V, tflat, ignore_index=pad_id)
        ctx.rt, ctx.lin, ctx.saved, ctx.pad = rt, lin, (h2, logits, lse, tflat, out), pad_id
        ctx.shape = (B, T, d, V, ldp)
        lg = logits.view(B, T, V) if ldp == V else buf.view(B, T, ldp)[..., :V]
        ctx.mark_non_differentiable(lg)
        return out[:1], lg

    @staticmethod
    def backward(ctx, dloss, _dlogits):
        rt, lin = ctx.rt, ctx.lin
        h2, logits, lse, tflat, out = ctx.saved
        B, T, d, V, ldp = ctx.shape
        dl = torch.empty(B * T, ldp, dtype=torch.bfloat16, device=h2.device)
        K.ce_bwd(logits, V, lse, tflat, out, dloss.contiguous().float(), 1.0, dl, ignore_index=ctx.pad)
        _wgrad(rt, lin, dl[:, :V], h2)
        dh = K.gemm(dl[:, :V], lin.w16, b_mn=True)
        ctx.saved = None
        return dh.view(B, T, d), None, None, None, None


class ColamFn(torch.autograd.Function):
    """CoLaM margin loss (TRAIN:292-309) between the model's and the frozen guide's last decoder states."""

    @staticmethod
    def forward(ctx, h, h_guide, tgt_ids, margin: float, pad_id: int):
        loss, pa, pb, stats = K.colam_fwd(h.contiguous(), h_guide.contiguous(), tgt_ids.contiguous(), margin, pad_id)
        ctx.saved = (pa, pb, stats, tgt_ids)
        ctx.shape, ctx.pad = h.shape, pad_id
        return loss

    @staticmethod
    def backward(ctx, dloss):
        pa, pb, stats, tgt_ids = ctx.saved
        dh = torch.empty(ctx.shape, dtype=torch.bfloat16, device=pa.device)
        K.colam_bwd(pa, pb, stats, tgt_ids, dloss.contiguous().float(), 1.0, dh, pad_id=ctx.pad)
        return dh, None, None, None, None


class SeclaFn(torch.autograd.Function):
    """SECLA BatchSoftmax (TRAIN:631-660): gradient flows into the face states only (names are no_grad)."""

    @staticmethod
    def forward(ctx, face, names):
        loss, ws = K.secla_fwd(names.contiguous(), face.contiguous())
        ctx.saved = (ws, names)
        ctx.shape = face.shape
        return loss

    @staticmethod
    def backward(ctx, dloss):
        ws, names = ctx.saved
        dface = torch.empty(ctx.shape, dtype=torch.bfloat16, device=names.device)
        K.secla_bwd(ws, names, dloss.contiguous().float(), 1.0, dface)
        return dface, None
