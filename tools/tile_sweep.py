"""Tile choice of vacnic_gemm for the 1024-row GEMMs (decode step: 256 captions x 4 beams; training decoder: 16 x 64 tokens):
every tile option on every shape, 40 launches back to back inside one CUDA graph (the way the step graphs run them), time
per launch and max |difference| against the automatic choice."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vacnic_b200 import kernels as K  # noqa: E402

dev = torch.device("cuda:0")
bf = torch.bfloat16
torch.manual_seed(0)
N_CHAIN = 40
OPTS = [("auto", 0), ("single 64", 64), ("single 128", 128), ("single 256", 256), ("pair 128", 1128), ("pair 256", 1256)]


def run(name, M, N, Kd, b_mn=False, a_mn=False, act=K.ACT_NONE, fp32=False):
    if a_mn:   # weight gradient: out[N, Kd] = dy[M, N]^T x[M, Kd], reduction over the M rows
        a = torch.randn(M, N, device=dev, dtype=bf) * 0.05
        b = torch.randn(M, Kd, device=dev, dtype=bf) * 0.05
        outs = (N, Kd)
        kw = dict(a_mn=True, b_mn=True)
        bias = None
    elif b_mn:  # data gradient: out[M, Kd] = dy[M, N] W[N, Kd]
        a = torch.randn(M, N, device=dev, dtype=bf) * 0.05
        b = torch.randn(N, Kd, device=dev, dtype=bf) * 0.05
        outs = (M, Kd)
        kw = dict(b_mn=True)
        bias = None
    else:
        a = torch.randn(M, Kd, device=dev, dtype=bf) * 0.05
        b = torch.randn(N, Kd, device=dev, dtype=bf) * 0.05
        outs = (M, N)
        kw = dict(act=act)
        bias = torch.randn(N, device=dev)
    ref = None
    res = []
    for label, tn in OPTS:
        o = torch.zeros(outs, device=dev, dtype=torch.float32 if (fp32 or a_mn) else bf)
        try:
            K.gemm(a, b, out=o, bias=bias, tile_n=tn, **kw)
            torch.cuda.synchronize()
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                K.gemm(a, b, out=o, bias=bias, tile_n=tn, **kw)
            torch.cuda.current_stream().wait_stream(s)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for _ in range(N_CHAIN):
                    K.gemm(a, b, out=o, bias=bias, tile_n=tn, **kw)
            g.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                g.replay()
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / (5 * N_CHAIN)
            if ref is None:
                ref = o.float().clone()
            err = (o.float() - ref).abs().max().item()
            res.append(f"{label} {us:6.1f} us (d {err:.1e})")
        except Exception as e:  # noqa: BLE001
            res.append(f"{label} n/a ({type(e).__name__})")
    print(f"{name:34s} " + " | ".join(res), flush=True)


for M in (1024, 512):
    run(f"fwd  {M}x3072x1024", M, 3072, 1024)
    run(f"fwd  {M}x1024x1024", M, 1024, 1024)
    run(f"fwd  {M}x4096x1024 gelu", M, 4096, 1024, act=K.ACT_GELU)
    run(f"fwd  {M}x1024x4096", M, 1024, 4096)
run("fwd  1024x50267x1024 fp32", 1024, 50267, 1024, fp32=True)
run("dgrad 1024x(1024)->1024", 1024, 1024, 1024, b_mn=True)
run("dgrad 1024x(4096)->1024", 1024, 4096, 1024, b_mn=True)
run("dgrad 1024x(1024)->4096", 1024, 1024, 4096, b_mn=True)
run("dgrad 1024x(3072)->1024", 1024, 3072, 1024, b_mn=True)
run("wgrad 1024x1024 k1024", 1024, 1024, 1024, a_mn=True)
run("wgrad 4096x1024 k1024", 1024, 4096, 1024, a_mn=True)
run("wgrad 1024x4096 k1024", 1024, 1024, 4096, a_mn=True)
run("wgrad 3072x1024 k1024", 1024, 3072, 1024, a_mn=True)
run("fwd  1280x1024x1024 (ner rows)", 1280, 1024, 1024)
run("fwd  320x4096x1024 gelu (img)", 320, 4096, 1024, act=K.ACT_GELU)
run("fwd  320x1024x4096 (img)", 320, 1024, 4096)
