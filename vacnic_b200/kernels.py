"""Tensor-level wrappers over the C ABI.  Every function enqueues hand-written sm_100a kernels on
torch's current CUDA stream; none of them falls back to PyTorch arithmetic."""
from __future__ import annotations

import ctypes as C

import torch

from . import lib as _l
from .lib import (ACT_GELU, ACT_NONE, ACT_TANH, DT_BF16, DT_F32, MASK_CAUSAL, MASK_KEYPAD,  # noqa: F401
                  MASK_NONE, GemmDesc, check, lib, ptr, stream_ptr)


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
         accumulate: bool = False, out_dtype: torch.dtype = torch.bfloat16, tile_n: int = 0) -> torch.Tensor:
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
    d.c, d.ldc, d.c_sb0, d.c_sb1 = o4.data_ptr(), o4.stride(2), o4.stride(1), o4.stride(0)
    d.bias, d.aux_out, d.aux_in = ptr(bias), ptr(aux_out), ptr(aux_in)
    d.alpha = alpha
    d.a_mn_major, d.b_mn_major = int(a_mn), int(b_mn)
    d.c_dtype = DT_F32 if out.dtype == torch.float32 else DT_BF16
    if out.dtype not in (torch.float32, torch.bfloat16):
        raise ValueError("gemm out must be bf16 or fp32")
    d.act, d.dact, d.accumulate, d.tile_n = act, dact, int(accumulate), tile_n
    check(lib().vacnic_gemm(C.byref(d), stream_ptr()), "vacnic_gemm")
    return out
