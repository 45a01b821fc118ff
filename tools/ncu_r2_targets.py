"""One launch each of the kernels whose round-2 ncu numbers DESIGN.md / bench.py cite (run under
`ncu --set full --clock-control none -k regex:<name>`):
  gemm2_sm100_kernel   fc1 + bias + GELU, M = 16384 (the representative launch of bench.py's roofline.traffic) and M = 12800
                       (a typical packed-row count)
  add_layernorm_fwd    residual + LayerNorm with the fp32 residual stream in and out (12800 x 1024)
  add_layernorm_bwd    its backward
  adamw_kernel         912.7 M parameters"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vacnic_b200 import kernels as K  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
for M in (16384, 12800):
    x = torch.randn(M, 1024, device=dev).bfloat16()
    w = (torch.randn(4096, 1024, device=dev) * 0.02).bfloat16()
    b = torch.zeros(4096, device=dev)
    o = torch.empty(M, 4096, device=dev, dtype=torch.bfloat16)
    aux = torch.empty_like(o)
    for _ in range(3):
        K.gemm(x, w, out=o, bias=b, act=K.ACT_GELU, aux_out=aux)
rows, d = 12800, 1024
xb = torch.randn(rows, d, device=dev).bfloat16()
res32 = torch.randn(rows, d, device=dev)
res16 = res32.bfloat16()
g, be = torch.ones(d, device=dev), torch.zeros(d, device=dev)
rng = K.Rng(dev, 1)
for _ in range(3):
    y, mean, rstd, y32 = K.add_layernorm_fwd(xb, res16, g, be, p_drop=0.1, rng=rng, salt=3, res32=res32, want_y32=True)
dy = torch.randn(rows, d, device=dev).bfloat16()
dg, db, dbias = (torch.zeros(d, device=dev) for _ in range(3))
for _ in range(3):
    K.add_layernorm_bwd(dy, xb, res16, g, mean, rstd, dg, db, dbias=dbias, want_dx=True, p_drop=0.1, rng=rng, salt=3)
n = 912_700_000 // 64 * 64
p, gr, m, v = (torch.zeros(n, device=dev) for _ in range(4))
p16 = torch.zeros(n, device=dev, dtype=torch.bfloat16)
hyper = torch.tensor([3e-5, 0.9, 0.999, 1e-8, 0.01, 0.1, 0.001, 1.0], device=dev)
for _ in range(2):
    K.adamw(p, gr, m, v, p16, hyper)
torch.cuda.synchronize()
print("done")
