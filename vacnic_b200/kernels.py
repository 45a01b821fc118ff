"""Tensor-level wrappers over the C ABI.  Every function enqueues hand-written sm_100a kernels on
torch's current CUDA stream; none of them falls back to PyTorch arithmetic."""
from __future__ import annotations

import ctypes as C

import torch

from . import lib as _l
from .lib import (ACT_GELU, ACT_NONE, ACT_QUICKGELU, ACT_TANH, DT_BF16, DT_F32, MASK_CAUSAL, MASK_KEYPAD,  # noqa: F401
                  MASK_NONE, AttnDesc, GemmDesc, check, lib, ptr, stream_ptr)


_WORKSPACES = {}  # (device index, stream slot) -> zero-initialised split-K scratch of vacnic_gemm.  GEMMs of one device are
# issued on one stream at a time (eager stream, capture stream and graph replays never overlap) -- except the prefix side
# stream of the encoder (BartEncoder.forward), which runs beside the main stream and therefore owns a scratch of its own
SPLIT_K_BYTES = 64 << 20
_SIDE_STREAMS = set()  # cudaStream_t handles registered by register_side_stream


def register_side_stream(stream: "torch.cuda.Stream"):
    """Give `stream` its own split-K scratch (allocated and zeroed here, outside any capture)."""
    _SIDE_STREAMS.add(stream.cuda_stream)
    key = (stream.device.index, stream.cuda_stream)
    if key not in _WORKSPACES:
        _WORKSPACES[key] = torch.zeros(SPLIT_K_BYTES, dtype=torch.uint8, device=stream.device)


def _gemm_workspace(device):
    cur = torch.cuda.current_stream(device).cuda_stream
    key = (device.index, cur if cur in _SIDE_STREAMS else 0)
    ws = _WORKSPACES.get(key)
    if ws is None:
        if torch.cuda.is_current_stream_capturing():
            return None  # never allocate + memset inside a capture: this launch simply does not split
        ws = torch.zeros(SPLIT_K_BYTES, dtype=torch.uint8, device=device)
        _WORKSPACES[key] = ws
    return ws


_SM_COUNT: dict = {}


def sm_count(device) -> int:
    n = _SM_COUNT.get(device.index)
    if n is None:
        n = _SM_COUNT[device.index] = torch.cuda.get_device_properties(device).multi_processor_count
    return n


def uses_pair_kernel(M: int, N: int, batch: int, tile_n: int = 0, sms: int = 148) -> bool:
    """Mirror of choose_pair_tile_n (csrc/gemm_sm100.cu): does vacnic_gemm route this problem to the CTA-pair kernel?"""
    if tile_n in (1128, 1256):
        return True
    if tile_n != 0 or M < 256:
        return False
    num_m = (M + 255) // 256
    return any(N >= bn and batch * num_m * ((N + bn - 1) // bn) >= sms // 2 for bn in (256, 128))


PROFILE = None  # set to a list to collect (flops, start_event, end_event, shape) per GEMM launch


def _require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _l.VacnicError("vacnic_b200 kernels need CUDA tensors: there is no CPU fallback")


def _as4(t: torch.Tensor) -> torch.Tensor:
    while t.dim() < 4:
        t = t.unsqueeze(0)
    if t.dim() != 4:
        raise ValueError(f"expected <=4 dims, got {t.dim()}")
    return t


def gemm(a: torch.Tensor, b: torch.Tensor, out: torch.Tensor | None = None, *, a_mn: bool = False,
         b_mn: bool = False, bias: torch.Tensor | None = None, alpha: float = 1.0, act: int = ACT_NONE,
         aux_out: torch.Tensor | None = None, aux_in: torch.Tensor | None = None, dact: int = ACT_NONE,
         accumulate: bool = False, out_dtype: torch.dtype = torch.bfloat16, tile_n: int = 0,
         head_major: tuple | None = None, split_k: bool = False) -> torch.Tensor:
    """out[..., m, n] = epilogue(sum_k A[..., m, k] * B[..., n, k]) for up to two batch dims.

    a: [..., M, K] (or [..., K, M] when a_mn), b: [..., N, K] (or [..., K, N] when b_mn); bf16, innermost
    stride 1, arbitrary outer strides (multiples of 8 elements).  Batch dims must match exactly."""
    _require_cuda(a, b, out, bias, aux_out, aux_in)
    if a.dtype != torch.bfloat16 or b.dtype != torch.bfloat16:
        raise ValueError("gemm operands must be bf16")
    a4, b4 = _as4(a), _as4(b)
    if a4.stride(3) != 1 or b4.stride(3) != 1:
        raise ValueError("gemm operands need innermost stride 1")
    if a_mn:
        K, M = a4.shape[2], a4.shape[3]
    else:
        M, K = a4.shape[2], a4.shape[3]
    if b_mn:
        Kb, N = b4.shape[2], b4.shape[3]
    else:
        N, Kb = b4.shape[2], b4.shape[3]
    if K != Kb:
        raise ValueError(f"gemm reduction dims differ: {K} vs {Kb}")
    if a4.shape[:2] != b4.shape[:2]:
        raise ValueError(f"gemm batch dims differ: {tuple(a4.shape[:2])} vs {tuple(b4.shape[:2])}")
    nb1, nb0 = a4.shape[0], a4.shape[1]
    if out is None:
        out = torch.empty(tuple(a.shape[:-2]) + (M, N), dtype=out_dtype, device=a.device)
    if head_major is not None:  # scattered (head-major) output: `out` only supplies base pointer, dtype and capacity
        if aux_out is not None or aux_in is not None or out.numel() < nb1 * nb0 * M * N:
            raise ValueError("gemm head_major output: no aux tensors, and out must hold batch*M*N elements")
        o4 = torch.empty_strided((nb1, nb0, M, N), (0, 0, 0, 1), dtype=out.dtype, device="meta")  # shape carrier only
    else:
        o4 = _as4(out)
    if tuple(o4.shape) != (nb1, nb0, M, N) or o4.stride(3) != 1:
        raise ValueError(f"gemm out has shape {tuple(out.shape)}, expected batch+({M},{N}) with innermost stride 1")
    for aux in (aux_out, aux_in):
        if aux is not None:
            x4 = _as4(aux)
            if x4.dtype != torch.bfloat16 or tuple(x4.shape) != tuple(o4.shape) or x4.stride() != o4.stride():
                raise ValueError("gemm aux tensors must be bf16 with the shape and strides of out")
    if bias is not None and (bias.dtype != torch.float32 or bias.numel() != N or not bias.is_contiguous()):
        raise ValueError("gemm bias must be contiguous fp32 of length N")
    d = GemmDesc()
    d.M, d.N, d.K, d.batch0, d.batch1 = M, N, K, nb0, nb1
    d.a, d.lda, d.a_sb0, d.a_sb1 = a4.data_ptr(), a4.stride(2), a4.stride(1), a4.stride(0)
    d.b, d.ldb, d.b_sb0, d.b_sb1 = b4.data_ptr(), b4.stride(2), b4.stride(1), b4.stride(0)
    d.c, d.ldc, d.c_sb0, d.c_sb1 = out.data_ptr(), o4.stride(2), o4.stride(1), o4.stride(0)
    d.bias, d.aux_out, d.aux_in = ptr(bias), ptr(aux_out), ptr(aux_in)
    d.alpha = alpha
    d.a_mn_major, d.b_mn_major = int(a_mn), int(b_mn)
    d.c_dtype = DT_F32 if out.dtype == torch.float32 else DT_BF16
    if out.dtype not in (torch.float32, torch.bfloat16):
        raise ValueError("gemm out must be bf16 or fp32")
    d.act, d.dact, d.accumulate, d.tile_n = act, dact, int(accumulate), tile_n
    # split-K: measured on B200 the partial-tile round trip + the serial last-arriver reduction cost more than they save for
    # the K <= 4096 problems of this model (decode fc2: 23 -> 60 us), but pay off for weight gradients whose reduction runs
    # over all B*L rows with only a few output tiles (1024x1024 with K = 16384: 75 -> ~35 us).  Hence: automatic for the
    # wgrad form (both operands MN-major) with K >= 8192, explicit (`split_k=True`) otherwise.
    auto_split = a_mn and b_mn and K >= 8192 and tile_n == 0
    ws = _gemm_workspace(a.device) if ((split_k or auto_split) and tile_n == 0) else None
    if ws is not None:
        d.workspace, d.workspace_bytes = ws.data_ptr(), ws.numel()
        d.split_k_min_blocks = 4 if split_k else 128
    if head_major is not None:
        # `out` only supplies the base pointer: element (b0, m, n) goes to b0*sb0 + m*ldc + (n // 64)*chunk + n % 64
        d.ldc, d.c_sb0, d.c_sb1, d.c_chunk_stride = head_major
    if PROFILE is not None:  # bench.py roofline pass: CUDA events around every launch, on the launching stream
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        check(lib().vacnic_gemm(C.byref(d), stream_ptr()), "vacnic_gemm")
        e1.record()
        PROFILE.append((2.0 * M * N * K * nb0 * nb1, e0, e1, (M, N, K, nb0 * nb1), uses_pair_kernel(M, N, nb0 * nb1, tile_n)))
        return out
    check(lib().vacnic_gemm(C.byref(d), stream_ptr()), "vacnic_gemm")
    return out


# ------------------------------------------------------------------------------------------------
# LayerNorm family, embeddings, softmax, reductions, optimizer, losses (thin checks + C-ABI call)
# ------------------------------------------------------------------------------------------------
LN_EPS = 1e-5  # nn.LayerNorm default (MFULL:577)


def _c(t, dtype, what):
    if t is None:
        return None
    if not t.is_cuda:
        raise _l.VacnicError(f"{what}: CUDA tensor required (no CPU fallback)")
    if t.dtype != dtype or not t.is_contiguous():
        raise ValueError(f"{what}: expected contiguous {dtype}, got {t.dtype} contiguous={t.is_contiguous()}")
    return t


class Rng:
    """Device-side dropout counter: one uint64 advanced once per training step (graph-capturable)."""

    def __init__(self, device, seed: int = 0):
        self.state = torch.tensor([seed * 0x9E3779B97F4A7C15 % (1 << 63)], dtype=torch.int64, device=device)

    def advance(self):
        check(lib().vacnic_rng_advance(self.state.data_ptr(), stream_ptr()), "vacnic_rng_advance")


def dropout_inplace(x, p_drop, rng, salt):
    """x = dropout(x) in place (contiguous bf16); the same (rng, salt) on the gradient applies the same mask."""
    _c(x, torch.bfloat16, "dropout operand")
    check(lib().vacnic_dropout_inplace(ptr(x), x.numel(), float(p_drop), rng.state.data_ptr(), int(salt), stream_ptr()),
          "vacnic_dropout_inplace")
    return x


def add_layernorm_fwd(x, res, gamma, beta, out=None, rows_per_group=0, group_stride=0, p_drop=0.0, rng=None, salt=0,
                      want_stats=True, res32=None, want_y32=False, y32_out=None, sum32_out=None):
    """y = LN(res + dropout(x)); returns (y, mean, rstd).  `out` may be a [rows, d] view whose groups of
    `rows_per_group` rows are `group_stride` elements apart (slice of the prefix/NER concat buffer).
    `want_stats=False` (inference) skips the per-row statistics the backward pass needs.
    `res32` (fp32 residual, replaces `res`) / `want_y32` (also return the output in fp32 as a 4th value): the fp32
    residual stream."""
    x_ptr = ptr(x)                       # x may be None (null pointer): plain LayerNorm of the fp32 stream `res32`
    src = x if x is not None else res32
    d = src.shape[-1]
    rows = src.numel() // d
    _c(x, torch.bfloat16, "x"); _c(res, torch.bfloat16, "res"); _c(gamma, torch.float32, "gamma"); _c(beta, torch.float32, "beta")
    _c(res32, torch.float32, "res32"); _c(sum32_out, torch.float32, "sum32_out")
    if out is None:
        out = torch.empty(src.shape, dtype=torch.bfloat16, device=src.device)
    y32 = y32_out if y32_out is not None else (torch.empty(src.shape, dtype=torch.float32, device=src.device) if want_y32 else None)
    mean = torch.empty(rows, dtype=torch.float32, device=src.device) if want_stats else None
    rstd = torch.empty(rows, dtype=torch.float32, device=src.device) if want_stats else None
    check(lib().vacnic_add_layernorm_fwd(x_ptr, ptr(res), ptr(gamma), ptr(beta), ptr(out), ptr(mean), ptr(rstd), rows, d,
                                         rows_per_group, group_stride, LN_EPS, p_drop,
                                         rng.state.data_ptr() if (rng is not None and p_drop > 0) else 0, salt,
                                         ptr(res32), ptr(y32), ptr(sum32_out), stream_ptr()), "vacnic_add_layernorm_fwd")
    if want_y32:
        return out, mean, rstd, y32
    return out, mean, rstd


def add_layernorm_bwd(dy, x, res, gamma, mean, rstd, dgamma, dbeta, dbias=None, dsum=None, want_dx=False,
                      rows_per_group=0, group_stride=0, p_drop=0.0, rng=None, salt=0, accumulate_dsum=False):
    """Returns (dsum, dx).  dx is dsum itself when no dropout is active."""
    d = x.shape[-1]
    rows = x.numel() // d
    if dsum is None:
        dsum = torch.empty_like(x)
    dx = dsum
    if p_drop > 0 and want_dx:
        dx = torch.empty_like(x)
    check(lib().vacnic_add_layernorm_bwd(ptr(dy), ptr(x), ptr(res), ptr(gamma), ptr(mean), ptr(rstd), ptr(dsum), ptr(dx),
                                         ptr(dgamma), ptr(dbeta), ptr(dbias), rows, d, rows_per_group, group_stride, p_drop,
                                         rng.state.data_ptr() if (rng is not None and p_drop > 0) else 0, salt,
                                         int(accumulate_dsum), stream_ptr()), "vacnic_add_layernorm_bwd")
    return dsum, dx


def embed_ln_fwd(ids, tok, pos, gamma, beta, pos_offset=2, p_drop=0.0, rng=None, salt=0, want_y32=False, pos_ids=None):
    _c(ids, torch.int64, "ids"); _c(tok, torch.bfloat16, "tok"); _c(pos, torch.bfloat16, "pos"); _c(pos_ids, torch.int32, "pos_ids")
    d = tok.shape[1]
    rows, seq = ids.numel(), ids.shape[-1]
    y = torch.empty(tuple(ids.shape) + (d,), dtype=torch.bfloat16, device=ids.device)
    mean = torch.empty(rows, dtype=torch.float32, device=ids.device)
    rstd = torch.empty(rows, dtype=torch.float32, device=ids.device)
    y32 = torch.empty(y.shape, dtype=torch.float32, device=ids.device) if want_y32 else None
    check(lib().vacnic_embed_ln_fwd(ptr(ids), ptr(tok), ptr(pos), ptr(gamma), ptr(beta), ptr(y), ptr(mean), ptr(rstd), rows,
                                    seq, pos_offset, d, LN_EPS, p_drop,
                                    rng.state.data_ptr() if (rng is not None and p_drop > 0) else 0, salt, ptr(y32),
                                    ptr(pos_ids), stream_ptr()),
          "vacnic_embed_ln_fwd")
    if want_y32:
        return y, mean, rstd, y32
    return y, mean, rstd


def embed_ln_bwd(dy, ids, tok, pos, gamma, mean, rstd, dtok, dpos, dgamma, dbeta, pos_offset=2, pad_id=1, p_drop=0.0,
                 rng=None, salt=0, pos_ids=None):
    d = tok.shape[1]
    check(lib().vacnic_embed_ln_bwd(ptr(dy), ptr(ids), ptr(tok), ptr(pos), ptr(gamma), ptr(mean), ptr(rstd), ptr(dtok),
                                    ptr(dpos), ptr(dgamma), ptr(dbeta), ids.numel(), ids.shape[-1], pos_offset, d, pad_id,
                                    p_drop, rng.state.data_ptr() if (rng is not None and p_drop > 0) else 0, salt,
                                    ptr(pos_ids), stream_ptr()), "vacnic_embed_ln_bwd")


def vit_patchify(images, patch):
    """fp32 [B, C, H, W] -> bf16 [B * (H/p) * (W/p), C*p*p] (conv1 of the CLIP ViT as a GEMM operand)."""
    _c(images, torch.float32, "images")
    B, Cc, H, W = images.shape
    out = torch.empty(B * (H // patch) * (W // patch), Cc * patch * patch, dtype=torch.bfloat16, device=images.device)
    check(lib().vacnic_vit_patchify(ptr(images), ptr(out), B, Cc, H, W, patch, stream_ptr()), "vacnic_vit_patchify")
    return out


def vit_embed_ln(tok, cls, pos, gamma, beta, batch, tokens):
    """class token + positional embedding + ln_pre -> fp32 [batch, tokens, d] (start of the ViT residual stream)."""
    _c(tok, torch.bfloat16, "tok"); _c(cls, torch.float32, "cls"); _c(pos, torch.float32, "pos")
    d = tok.shape[-1]
    y32 = torch.empty(batch, tokens, d, dtype=torch.float32, device=tok.device)
    check(lib().vacnic_vit_embed_ln(ptr(tok), ptr(cls), ptr(pos), ptr(gamma), ptr(beta), ptr(y32), batch, tokens, d, LN_EPS,
                                    stream_ptr()), "vacnic_vit_embed_ln")
    return y32


def unpack_rows(src2d, start, lens, L, out=None):
    """bf16 packed rows [rows, d] -> right-padded [B, L, d] (pad rows zero)."""
    _c(src2d, torch.bfloat16, "packed rows"); _c(start, torch.int32, "start"); _c(lens, torch.int32, "len")
    B, d = start.numel(), src2d.shape[-1]
    if out is None:
        out = torch.empty(B, L, d, dtype=torch.bfloat16, device=src2d.device)
    check(lib().vacnic_unpack_rows(ptr(src2d), ptr(start), ptr(lens), ptr(out), B, L, d, stream_ptr()), "vacnic_unpack_rows")
    return out


def names_embed(ids3, tok, pos, gamma, beta):
    """get_embedding_ner (TRAIN:112-133): ids [B,N,len] -> fp32 [B,N,d]."""
    _c(ids3, torch.int64, "names_ids")
    B, N, ln = ids3.shape
    d = tok.shape[1]
    out = torch.empty(B, N, d, dtype=torch.float32, device=ids3.device)
    check(lib().vacnic_names_embed(ptr(ids3), ptr(tok), ptr(pos), ptr(gamma), ptr(beta), ptr(out), B * N, ln, d, LN_EPS,
                                   stream_ptr()), "vacnic_names_embed")
    return out


def softmax_fwd(scores, key_mask, Sk, causal=False, past=0, out=None):
    """scores fp32 [B,H,Sq,ld] -> probs bf16 [B,H,Sq,ld] (pad columns zero)."""
    _c(scores, torch.float32, "scores"); _c(key_mask, torch.uint8, "key_mask")
    B, H, Sq, ld = scores.shape
    if out is None:
        out = torch.empty(scores.shape, dtype=torch.bfloat16, device=scores.device)
    check(lib().vacnic_softmax_fwd(ptr(scores), ptr(out), ptr(key_mask), B, H, Sq, Sk, ld, int(causal), past, stream_ptr()),
          "vacnic_softmax_fwd")
    return out


def softmax_bwd(probs, dprobs, Sk, out=None):
    _c(probs, torch.bfloat16, "probs"); _c(dprobs, torch.float32, "dprobs")
    ld = probs.shape[-1]
    if out is None:
        out = torch.empty_like(probs)
    check(lib().vacnic_softmax_bwd(ptr(probs), ptr(dprobs), ptr(out), probs.numel() // ld, Sk, ld, stream_ptr()),
          "vacnic_softmax_bwd")
    return out


def colsum_into(x2d, out):
    """out[c] += sum_r x2d[r, c]   (bias gradients)."""
    if x2d.dtype != torch.bfloat16 or x2d.stride(-1) != 1 or x2d.dim() != 2:
        raise ValueError("colsum: expected 2-D bf16 with innermost stride 1")
    check(lib().vacnic_colsum(ptr(x2d), ptr(out), x2d.shape[0], x2d.shape[1], x2d.stride(0), stream_ptr()), "vacnic_colsum")


def cast_bf16(src, dst):
    check(lib().vacnic_cast_f32_bf16(ptr(src), ptr(dst), src.numel(), stream_ptr()), "vacnic_cast_f32_bf16")


def add_bf16(a, b, c=None, out=None):
    if out is None:
        out = torch.empty_like(a)
    for t in (a, b, c, out):
        _c(t, torch.bfloat16, "add_bf16 operand")
    check(lib().vacnic_add_bf16(ptr(a), ptr(b), ptr(c), ptr(out), a.numel(), stream_ptr()), "vacnic_add_bf16")
    return out


def pad_rows(src2d, ld_dst):
    """bf16 [rows, n] contiguous -> [rows, ld_dst] zero padded (TMA needs a 16-byte row pitch)."""
    rows, n = src2d.shape
    dst = torch.empty(rows, ld_dst, dtype=torch.bfloat16, device=src2d.device)
    check(lib().vacnic_pad_rows(ptr(src2d), ptr(dst), rows, n, ld_dst, stream_ptr()), "vacnic_pad_rows")
    return dst


def cast_rows_bf16(src2d):
    """fp32 [rows, n] (row stride arbitrary, innermost 1) -> bf16 [rows, n] view of a buffer with a 16-byte row pitch."""
    rows, n = src2d.shape
    if src2d.dtype != torch.float32 or src2d.stride(1) != 1:
        raise ValueError("cast_rows_bf16: fp32 rows with innermost stride 1 expected")
    ld = (n + 7) // 8 * 8
    dst = torch.empty(rows, ld, dtype=torch.bfloat16, device=src2d.device)
    check(lib().vacnic_cast_rows_f32_bf16(ptr(src2d), ptr(dst), rows, n, src2d.stride(0), ld, stream_ptr()),
          "vacnic_cast_rows_f32_bf16")
    return dst[:, :n]


def sum_partials(parts2d, dst, accumulate):
    """dst (+)= parts2d.sum(0); fp32."""
    check(lib().vacnic_sum_partials(ptr(parts2d), ptr(dst), parts2d.shape[0], dst.numel(), int(accumulate), stream_ptr()),
          "vacnic_sum_partials")


def adamw(p, g, m, v, p16, hyper):
    check(lib().vacnic_adamw(ptr(p), ptr(g), ptr(m), ptr(v), ptr(p16), p.numel(), ptr(hyper), stream_ptr()), "vacnic_adamw")


def optim_schedule(step_dev, hyper, base_lr, beta1, beta2, eps, weight_decay, warmup_steps, total_steps, grad_scale):
    """Advance the device-side step counter (int64[1]) and write this update's {lr, betas, eps, wd, bias corrections,
    grad_scale} into `hyper` (fp32[8]) -- TRAIN:102, 371-374."""
    _c(step_dev, torch.int64, "step counter"); _c(hyper, torch.float32, "hyper")
    check(lib().vacnic_optim_schedule(ptr(step_dev), ptr(hyper), float(base_lr), float(beta1), float(beta2), float(eps),
                                      float(weight_decay), int(warmup_steps), int(total_steps), float(grad_scale), stream_ptr()),
          "vacnic_optim_schedule")


def dp_adamw_shard(grad_ptrs, shadow_ptrs, grad_mc, shadow_mc, world, rank, master, m, v, hyper, begin, count, max_blocks=0):
    """Fused reduce-scatter + AdamW + all-gather over peer memory for this rank's shard [begin, begin+count) of the flat
    buffers.  grad_ptrs / shadow_ptrs: ctypes arrays (c_uint64 * world) of every rank's buffer address."""
    check(lib().vacnic_dp_adamw_shard(grad_ptrs, shadow_ptrs, int(grad_mc), int(shadow_mc), world, rank, ptr(master), ptr(m),
                                      ptr(v), ptr(hyper), int(begin), int(count), int(max_blocks), stream_ptr()),
          "vacnic_dp_adamw_shard")


def clip_grad_scale(g, max_norm, base_scale, scratch, scale_out, norm_out=None):
    """scale_out[0] = base_scale * min(1, max_norm / (|base_scale| * ||g|| + 1e-6))  (clip_grad_norm_ folded into AdamW)."""
    check(lib().vacnic_clip_grad_scale(ptr(g), g.numel(), float(max_norm), float(base_scale), ptr(scratch), ptr(scale_out),
                                       ptr(norm_out), stream_ptr()), "vacnic_clip_grad_scale")


def ce_fwd(logits2d, V, targets, ignore_index=1):
    """logits fp32 [rows, ld] (ld >= V). Returns (out[2] = {mean loss, count}, lse, row_loss)."""
    _c(targets, torch.int64, "targets")
    if logits2d.dtype != torch.float32 or logits2d.stride(1) != 1:
        raise ValueError("ce_fwd: logits must be fp32 with innermost stride 1")
    rows, ld = logits2d.shape[0], logits2d.stride(0)
    dev = logits2d.device
    lse = torch.empty(rows, dtype=torch.float32, device=dev)
    row_loss = torch.empty(rows, dtype=torch.float32, device=dev)
    out = torch.empty(2, dtype=torch.float32, device=dev)
    check(lib().vacnic_ce_fwd(ptr(logits2d), ptr(targets), ptr(lse), ptr(row_loss), ptr(out), rows, V, ld, ignore_index,
                              stream_ptr()), "vacnic_ce_fwd")
    return out, lse, row_loss


def ce_bwd(logits2d, V, lse, targets, stats, gscale, coef, dlogits, ignore_index=1):
    rows, ld = logits2d.shape[0], logits2d.stride(0)
    check(lib().vacnic_ce_bwd(ptr(logits2d), ptr(lse), ptr(targets), ptr(stats), ptr(gscale), coef, ptr(dlogits), rows, V, ld,
                              ignore_index, stream_ptr()), "vacnic_ce_bwd")
    return dlogits


def colam_fwd(h, hg, tgt_ids, margin, pad_id=1):
    _c(h, torch.bfloat16, "h"); _c(hg, torch.bfloat16, "h_guide"); _c(tgt_ids, torch.int64, "tgt_ids")
    B, T, d = h.shape
    dev = h.device
    pa = torch.empty(B, d, dtype=torch.float32, device=dev)
    pb = torch.empty(B, d, dtype=torch.float32, device=dev)
    stats = torch.empty(B, 8, dtype=torch.float32, device=dev)
    loss = torch.empty(1, dtype=torch.float32, device=dev)
    check(lib().vacnic_colam_fwd(ptr(h), ptr(hg), ptr(tgt_ids), ptr(pa), ptr(pb), ptr(stats), ptr(loss), B, T, d, pad_id,
                                 margin, stream_ptr()), "vacnic_colam_fwd")
    return loss, pa, pb, stats


def colam_bwd(pa, pb, stats, tgt_ids, gscale, coef, dh, accumulate=False, pad_id=1):
    B, T, d = dh.shape
    check(lib().vacnic_colam_bwd(ptr(pa), ptr(pb), ptr(stats), ptr(tgt_ids), ptr(gscale), coef, ptr(dh), B, T, d, pad_id,
                                 int(accumulate), stream_ptr()), "vacnic_colam_bwd")
    return dh


def secla_fwd(names, face):
    _c(names, torch.float32, "names"); _c(face, torch.bfloat16, "face")
    B, N, d = names.shape
    F = face.shape[1]
    ws = torch.empty(int(lib().vacnic_secla_workspace_bytes(B, N, F)) // 4, dtype=torch.float32, device=face.device)
    loss = torch.empty(1, dtype=torch.float32, device=face.device)
    check(lib().vacnic_secla_fwd(ptr(names), ptr(face), ptr(ws), ptr(loss), B, N, F, d, stream_ptr()), "vacnic_secla_fwd")
    return loss, ws


def secla_bwd(ws, names, gscale, coef, dface, accumulate=False):
    B, N, d = names.shape
    F = dface.shape[1]
    check(lib().vacnic_secla_bwd(ptr(ws), ptr(names), ptr(gscale), coef, ptr(dface), B, N, F, d, int(accumulate), stream_ptr()),
          "vacnic_secla_bwd")
    return dface


def concat_rows(a, b, out):
    """out[B, Sa+Sb, d] = cat(a [B,Sa,d], b [B,Sb,d]) along dim 1 (bf16, contiguous)."""
    a, b = a.contiguous(), b.contiguous()
    for t in (a, b, out):
        _c(t, torch.bfloat16, "concat operand")
    check(lib().vacnic_concat_rows(ptr(a), ptr(b), ptr(out), a.shape[0], a.shape[1], b.shape[1], a.shape[2], stream_ptr()),
          "vacnic_concat_rows")
    return out


# ------------------------------------------------------------------------------------------------
# cached decoding / device-side search (csrc/decode.cu)
# ------------------------------------------------------------------------------------------------
def decode_embed_ln(seq, cur_len, tok, pos, gamma, beta, y, maxT, pos_offset=2, pingpong=False, y32=None):
    R, d = y.shape
    check(lib().vacnic_decode_embed_ln(ptr(seq), ptr(cur_len), ptr(tok), ptr(pos), ptr(gamma), ptr(beta), ptr(y), R, maxT, d,
                                       pos_offset, int(pingpong), LN_EPS, ptr(y32), stream_ptr()), "vacnic_decode_embed_ln")
    return y


def decode_self_attn(qkv, kcache, vcache, anc, cur_len, out, H, maxT):
    R, d = out.shape
    check(lib().vacnic_decode_self_attn(ptr(qkv), ptr(kcache), ptr(vcache), ptr(anc), ptr(cur_len), ptr(out), R, H, d // H,
                                        maxT, stream_ptr()), "vacnic_decode_self_attn")
    return out


def decode_cross_attn(q, k4, v4, key_mask, key_len, out, nq):
    """q [captions*nq, d]; k4 / v4 [captions, H, L, 64] views (any caption / head / row strides, innermost 1)."""
    captions, H, L, hd = k4.shape
    if k4.stride() != v4.stride() or k4.stride(3) != 1:
        raise ValueError("decode_cross_attn: k and v views must share strides, innermost stride 1")
    check(lib().vacnic_decode_cross_attn(ptr(q), q.stride(0), k4.data_ptr(), v4.data_ptr(), k4.stride(2), k4.stride(1),
                                         k4.stride(0), ptr(key_mask), ptr(key_len), ptr(out), out.stride(0), captions, nq, L,
                                         H, hd, stream_ptr()), "vacnic_decode_cross_attn")
    return out


def mask_key_len(mask_u8):
    _c(mask_u8, torch.uint8, "key mask")
    B, L = mask_u8.shape
    out = torch.empty(B, dtype=torch.int32, device=mask_u8.device)
    check(lib().vacnic_mask_key_len(ptr(mask_u8), ptr(out), B, L, stream_ptr()), "vacnic_mask_key_len")
    return out


def decode_topk(logits2d, V, K_, top_lp, top_idx):
    check(lib().vacnic_decode_topk(ptr(logits2d), logits2d.stride(0), logits2d.shape[0], V, K_, ptr(top_lp), ptr(top_idx),
                                   stream_ptr()), "vacnic_decode_topk")


def beam_step(top_lp, top_idx, st, captions, beams, maxT, max_len, eos, V, length_penalty):
    check(lib().vacnic_beam_step(ptr(top_lp), ptr(top_idx), ptr(st["run_seq"]), ptr(st["run_anc"]), ptr(st["run_score"]),
                                 ptr(st["fin_seq"]), ptr(st["fin_score"]), ptr(st["fin_len"]), ptr(st["fin_flag"]),
                                 ptr(st["unsat"]), ptr(st["flags"]), ptr(st["cur_len"]), captions, beams, maxT, max_len, eos,
                                 V, length_penalty, stream_ptr()), "vacnic_beam_step")


def greedy_step(top_idx, seq, unfinished, flags, cur_len, rows, maxT, max_len, eos, pad):
    check(lib().vacnic_greedy_step(ptr(top_idx), ptr(seq), ptr(unfinished), ptr(flags), ptr(cur_len), rows, maxT, max_len, eos,
                                   pad, stream_ptr()), "vacnic_greedy_step")


def advance_len(cur_len):
    check(lib().vacnic_advance_len(ptr(cur_len), stream_ptr()), "vacnic_advance_len")


# ------------------------------------------------------------------------------------------------
# fused attention (csrc/attn_sm100.cu)
# ------------------------------------------------------------------------------------------------
class Packed:
    """Varlen geometry of one attention call (device int32 arrays of length n_seq + host-known bounds): sequence b owns
    query rows [q_start[b], +q_len[b]) and key rows [k_start[b], +k_len[b]) of the packed row buffers."""
    __slots__ = ("q_start", "q_len", "k_start", "k_len", "n_seq", "max_q", "max_k")

    def __init__(self, q_start, q_len, k_start, k_len, max_q: int, max_k: int):
        for t in (q_start, q_len, k_start, k_len):
            _c(t, torch.int32, "packed attention geometry")
        if not (q_start.numel() == q_len.numel() == k_start.numel() == k_len.numel()):
            raise ValueError("packed attention geometry: the four arrays must have one entry per sequence")
        self.q_start, self.q_len, self.k_start, self.k_len = q_start, q_len, k_start, k_len
        self.n_seq, self.max_q, self.max_k = q_start.numel(), int(max_q), int(max_k)


def _attn_desc(q4, k4, v4, key_mask, key_len, causal, packed: "Packed | None" = None, p_drop: float = 0.0, rng=None, salt: int = 0):
    B, H, Sq, hd = q4.shape
    Sk = k4.shape[2]
    for t in (q4, k4, v4):
        if t.dtype != torch.bfloat16 or t.stride(3) != 1:
            raise ValueError("attention operands must be bf16 [B,H,S,hd] views with innermost stride 1")
    _require_cuda(q4, k4, v4, key_mask, key_len)
    if key_mask is not None and (key_mask.dtype != torch.uint8 or tuple(key_mask.shape) != (B, Sk) or not key_mask.is_contiguous()):
        raise ValueError(f"Attention mask should be of size {(B, 1, Sq, Sk)}: expected a contiguous uint8 key mask [{B}, {Sk}]")
    d = AttnDesc()
    d.B, d.H, d.Sq, d.Sk, d.head_dim, d.causal = B, H, Sq, Sk, hd, int(causal)
    d.q, d.ldq, d.q_sh, d.q_sb = q4.data_ptr(), q4.stride(2), q4.stride(1), q4.stride(0)
    d.k, d.ldk, d.k_sh, d.k_sb = k4.data_ptr(), k4.stride(2), k4.stride(1), k4.stride(0)
    d.v, d.ldv, d.v_sh, d.v_sb = v4.data_ptr(), v4.stride(2), v4.stride(1), v4.stride(0)
    d.key_mask, d.key_len = ptr(key_mask), ptr(key_len)
    if packed is not None:
        # q4 / k4 / v4 are [1, H, total rows, 64] views of the packed buffers; B / Sq / Sk become sequence count and maxima
        if B != 1 or key_mask is not None or key_len is not None:
            raise ValueError("packed attention: operands must be [1, H, rows, 64] views and take no key mask")
        d.B, d.Sq, d.Sk = packed.n_seq, packed.max_q, packed.max_k
        d.total_q, d.total_k = Sq, Sk
        d.q_start, d.q_len, d.k_start, d.k_len = (ptr(t) for t in (packed.q_start, packed.q_len, packed.k_start, packed.k_len))
    if p_drop > 0.0:  # attention dropout (config.attention_dropout, MFULL:546); the same key rebuilds the mask in attn_bwd
        if rng is None:
            raise ValueError("attention dropout needs the model's Rng")
        d.p_drop, d.rng_state, d.salt = float(p_drop), rng.state.data_ptr(), int(salt)
    return d


def attn_fwd(q4, k4, v4, key_mask=None, key_len=None, causal=False, want_stats=True, packed=None, p_drop=0.0, rng=None, salt=0):
    """Fused softmax(q k^T * hd^-0.5 + mask) v.  q4 [B,H,Sq,64], k4/v4 [B,H,Sk,64] strided views.
    Returns (O bf16 [B,Sq,H*64], stats fp32 [B,H,Sq,2] or None).  `packed` (kernels.Packed): varlen mode -- q4/k4/v4 are
    [1,H,rows,64] views of packed row buffers, O is [1,total_q,H*64], stats [1,H,total_q,2]."""
    B, H, Sq, hd = q4.shape
    d = _attn_desc(q4, k4, v4, key_mask, key_len, causal, packed, p_drop, rng, salt)
    out = torch.empty(B, Sq, H * hd, dtype=torch.bfloat16, device=q4.device)
    stats = torch.empty(B, H, Sq, 2, dtype=torch.float32, device=q4.device) if want_stats else None
    d.out, d.ldo, d.o_sb, d.stats = out.data_ptr(), out.stride(1), out.stride(0), ptr(stats)
    check(lib().vacnic_attn_fwd(C.byref(d), stream_ptr()), "vacnic_attn_fwd")
    return out, stats


def attn_bwd(dO, O, stats, q4, k4, v4, dq4, dk4, dv4, key_mask=None, key_len=None, causal=False, packed=None, p_drop=0.0,
             rng=None, salt=0):
    """Gradients of attn_fwd.  dO / O bf16 [B,Sq,H*64] (row stride = stride(1), innermost 1); dq4/dk4/dv4 are
    [B,H,S,64] strided views that receive the results (dq, dk include the hd^-0.5 factor)."""
    B, H, Sq, hd = q4.shape
    d = _attn_desc(q4, k4, v4, key_mask, key_len, causal, packed, p_drop, rng, salt)
    for t in (dO, O):
        if t.dtype != torch.bfloat16 or t.stride(2) != 1 or tuple(t.shape) != (B, Sq, H * hd):
            raise ValueError("attn_bwd: dO / O must be bf16 [B,Sq,H*hd] with innermost stride 1")
    delta = torch.empty(B, H, Sq, dtype=torch.float32, device=q4.device)
    d.out, d.ldo, d.o_sb, d.stats = O.data_ptr(), O.stride(1), O.stride(0), ptr(stats)
    d.dout, d.lddo, d.do_sb = dO.data_ptr(), dO.stride(1), dO.stride(0)
    d.dq, d.lddq, d.dq_sh, d.dq_sb = dq4.data_ptr(), dq4.stride(2), dq4.stride(1), dq4.stride(0)
    d.dk, d.lddk, d.dk_sh, d.dk_sb = dk4.data_ptr(), dk4.stride(2), dk4.stride(1), dk4.stride(0)
    d.dv, d.lddv, d.dv_sh, d.dv_sb = dv4.data_ptr(), dv4.stride(2), dv4.stride(1), dv4.stride(0)
    d.delta = delta.data_ptr()
    check(lib().vacnic_attn_bwd(C.byref(d), stream_ptr()), "vacnic_attn_bwd")
