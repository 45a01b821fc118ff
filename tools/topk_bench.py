"""Times `vacnic_decode_topk` alone at the decode step's shape (1024 rows = 256 captions x 4 beams, V = 50267, K = 8).
The logits buffer (206 MB) is larger than L2, so every launch reads the rows from HBM.
    python tools/topk_bench.py [--rows 1024] [--iters 50]        (under ncu: --iters 2)"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from vacnic_b200 import kernels as K  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=1024)
ap.add_argument("--vocab", type=int, default=50267)
ap.add_argument("--k", type=int, default=8)
ap.add_argument("--iters", type=int, default=50)
a = ap.parse_args()
dev = torch.device("cuda:0")
ld = (a.vocab + 7) // 8 * 8
bufs = [torch.randn(a.rows, ld, device=dev) * 4 for _ in range(2)]
lp = torch.empty(a.rows, a.k, device=dev)
ix = torch.empty(a.rows, a.k, dtype=torch.int32, device=dev)
for i in range(3):
    K.decode_topk(bufs[i & 1][:, :a.vocab], a.vocab, a.k, lp, ix)
torch.cuda.synchronize()
ref_lp, ref_ix = torch.topk(torch.log_softmax(bufs[0][:, :a.vocab], -1), a.k)
assert bool((ix.long() == ref_ix).all()) and torch.allclose(lp, ref_lp, atol=2e-5, rtol=1e-5)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
ev[0].record()
for i in range(a.iters):
    K.decode_topk(bufs[i & 1][:, :a.vocab], a.vocab, a.k, lp, ix)
ev[1].record()
torch.cuda.synchronize()
us = ev[0].elapsed_time(ev[1]) * 1e3 / a.iters
gb = a.rows * a.vocab * 4 / 1e9
print(f"decode_topk rows={a.rows} V={a.vocab} K={a.k}: {us:.1f} us/launch, {gb / (us * 1e-6):.0f} GB/s of the {gb * 1e3:.0f} MB it must read")
