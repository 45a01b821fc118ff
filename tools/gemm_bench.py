"""Microbenchmark of vacnic_gemm on the GEMM shapes of one BART-large VACNIC training step (B=16, L=1024).
Prints TFLOP/s per shape (CUDA events, L2 flushed by rotating through buffers larger than L2)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vacnic_b200 import kernels as K  # noqa: E402

dev = torch.device("cuda:0")
bf = torch.bfloat16


def bench(name, make, flops, reps=20):
    sets = [make() for _ in range(3)]
    for s in sets:
        s()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        sets[i % 3]()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{name:58s} {ms * 1e3:9.1f} us  {flops / 1e12 / (ms / 1e3):8.1f} TF/s", flush=True)


def lin_fwd(M, N, Kd, act=K.ACT_NONE, tile_n=0):
    def make():
        a = torch.randn(M, Kd, device=dev, dtype=bf)
        w = torch.randn(N, Kd, device=dev, dtype=bf)
        b = torch.randn(N, device=dev)
        o = torch.empty(M, N, device=dev, dtype=bf)
        return lambda: K.gemm(a, w, out=o, bias=b, act=act, tile_n=tile_n)
    return make


def lin_dgrad(M, N, Kd, tile_n=0):  # dx[M,K] = dy[M,N] W[N,K]
    def make():
        dy = torch.randn(M, N, device=dev, dtype=bf)
        w = torch.randn(N, Kd, device=dev, dtype=bf)
        o = torch.empty(M, Kd, device=dev, dtype=bf)
        return lambda: K.gemm(dy, w, out=o, b_mn=True, tile_n=tile_n)
    return make


def lin_wgrad(M, N, Kd, tile_n=0):  # gw[N,K] = dy[M,N]^T x[M,K]
    def make():
        dy = torch.randn(M, N, device=dev, dtype=bf)
        x = torch.randn(M, Kd, device=dev, dtype=bf)
        o = torch.empty(N, Kd, device=dev, dtype=torch.float32)
        return lambda: K.gemm(dy, x, out=o, a_mn=True, b_mn=True, tile_n=tile_n)
    return make


def qk(B, H, S, hd=64):
    def make():
        qkv = torch.randn(B * S, 3 * H * hd, device=dev, dtype=bf)
        ld = qkv.stride(0)
        q4 = qkv.as_strided((B, H, S, hd), (S * ld, hd, ld, 1), 2 * H * hd)
        k4 = qkv.as_strided((B, H, S, hd), (S * ld, hd, ld, 1), 0)
        s = torch.empty(B, H, S, S, device=dev, dtype=torch.float32)
        return lambda: K.gemm(q4, k4, out=s, alpha=0.125)
    return make


def pv(B, H, S, hd=64):
    def make():
        qkv = torch.randn(B * S, 3 * H * hd, device=dev, dtype=bf)
        ld = qkv.stride(0)
        v4 = qkv.as_strided((B, H, S, hd), (S * ld, hd, ld, 1), H * hd)
        p = torch.randn(B, H, S, S, device=dev, dtype=bf)
        o = torch.empty(B, S, H, hd, device=dev, dtype=bf)
        return lambda: K.gemm(p, v4, out=o.permute(0, 2, 1, 3), b_mn=True)
    return make


if __name__ == "__main__":
    M = 16 * 1024
    only = sys.argv[1] if len(sys.argv) > 1 else ""
    cases = [
        ("fwd qkv   16384x3072x1024", lin_fwd(M, 3072, 1024), 2 * M * 3072 * 1024),
        ("fwd out   16384x1024x1024", lin_fwd(M, 1024, 1024), 2 * M * 1024 * 1024),
        ("fwd fc1   16384x4096x1024 gelu", lin_fwd(M, 4096, 1024, K.ACT_GELU), 2 * M * 4096 * 1024),
        ("fwd fc2   16384x1024x4096", lin_fwd(M, 1024, 4096), 2 * M * 1024 * 4096),
        ("fwd fc1 tile128", lin_fwd(M, 4096, 1024, K.ACT_NONE, 128), 2 * M * 4096 * 1024),
        ("fwd crosskv 16384x24576x1024", lin_fwd(M, 24576, 1024), 2 * M * 24576 * 1024),
        ("dgrad fc2 16384x(1024)->4096", lin_dgrad(M, 1024, 4096), 2 * M * 1024 * 4096),
        ("dgrad fc1 16384x(4096)->1024", lin_dgrad(M, 4096, 1024), 2 * M * 1024 * 4096),
        ("dgrad qkv 16384x(3072)->1024", lin_dgrad(M, 3072, 1024), 2 * M * 3072 * 1024),
        ("wgrad fc1 4096x1024 k16384", lin_wgrad(M, 4096, 1024), 2 * M * 1024 * 4096),
        ("wgrad fc2 1024x4096 k16384", lin_wgrad(M, 1024, 4096), 2 * M * 1024 * 4096),
        ("wgrad out 1024x1024 k16384", lin_wgrad(M, 1024, 1024), 2 * M * 1024 * 1024),
        ("QK^T 16x16 heads 1024x1024x64 fp32 out", qk(16, 16, 1024), 2 * 256 * 1024 * 1024 * 64),
        ("PV   16x16 heads 1024x64x1024", pv(16, 16, 1024), 2 * 256 * 1024 * 1024 * 64),
        ("decode qkv 256x3072x1024", lin_fwd(256, 3072, 1024), 2 * 256 * 3072 * 1024),
        ("decode out 256x1024x1024", lin_fwd(256, 1024, 1024), 2 * 256 * 1024 * 1024),
        ("decode fc1 256x4096x1024", lin_fwd(256, 4096, 1024), 2 * 256 * 4096 * 1024),
        ("decode fc2 256x1024x4096", lin_fwd(256, 1024, 4096), 2 * 256 * 1024 * 4096),
        ("decode lm  256x50267x1024", lin_fwd(256, 50264, 1024), 2 * 256 * 50264 * 1024),
    ]
    for name, mk, fl in cases:
        if only and only not in name:
            continue
        bench(name, mk, fl)
    # torch (cuBLAS) reference points for the same shapes
    if not only:
        for (m, n, k) in ((M, 4096, 1024), (M, 1024, 4096), (M, 3072, 1024)):
            a = torch.randn(m, k, device=dev, dtype=bf); w = torch.randn(n, k, device=dev, dtype=bf)
            bench(f"cuBLAS {m}x{n}x{k}", lambda: (lambda: torch.matmul(a, w.t())), 2 * m * n * k)
