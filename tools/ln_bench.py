"""Residual + LayerNorm forward / backward (csrc/norm.cu) timed alone with CUDA events over buffer sets larger than L2:
achieved HBM GB/s against the algorithmic bytes (fwd: x 2 + res32 4 in, y 2 + y32 4 out; bwd: dy 2 + x 2 + res 2 in,
dsum 2 + dx 2 out per element)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vacnic_b200 import kernels as K  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
peak = 6527.8
if os.path.exists("MEASURED_PEAKS.json"):
    peak = json.load(open("MEASURED_PEAKS.json")).get("hbm_gbs", peak)
d = 1024
for rows in (12800, 16384, 1024):
    nset = max(2, int(400e6 // (rows * d * 14)))  # rotate over > 2 x L2 of distinct buffers
    sets = []
    for _ in range(nset):
        xb = torch.randn(rows, d, device=dev).bfloat16()
        res32 = torch.randn(rows, d, device=dev)
        sets.append((xb, res32, res32.bfloat16(), torch.randn(rows, d, device=dev).bfloat16()))
    g, be = torch.ones(d, device=dev), torch.zeros(d, device=dev)
    dg, db, dbias = (torch.zeros(d, device=dev) for _ in range(3))
    rng = K.Rng(dev, 1)
    for p in (0.1, 0.0):
        outs = [K.add_layernorm_fwd(s[0], s[2], g, be, p_drop=p, rng=rng, salt=3, res32=s[1], want_y32=True) for s in sets]
        torch.cuda.synchronize()
        reps = 10
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            for s in sets:
                K.add_layernorm_fwd(s[0], s[2], g, be, p_drop=p, rng=rng, salt=3, res32=s[1], want_y32=True)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / (reps * nset)
        gb = rows * d * 12 / 1e9
        print(f"fwd rows {rows} p_drop {p}: {us:.1f} us  {gb / (us * 1e-6):.0f} GB/s  {gb / (us * 1e-6) / peak:.2f} of peak")
        e0.record()
        for _ in range(reps):
            for s, o in zip(sets, outs):
                K.add_layernorm_bwd(s[3], s[0], s[2], g, o[1], o[2], dg, db, dbias=dbias, want_dx=True, p_drop=p, rng=rng, salt=3)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / (reps * nset)
        gb = rows * d * 10 / 1e9
        print(f"bwd rows {rows} p_drop {p}: {us:.1f} us  {gb / (us * 1e-6):.0f} GB/s  {gb / (us * 1e-6) / peak:.2f} of peak")
