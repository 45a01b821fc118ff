"""Fused attention kernels at the BART-large encoder self-attention shape (B=16, H=16, S=1024, hd=64):
time per kernel (CUDA events) and achieved TFLOP/s (4*S*S*64 FLOP per head forward)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vacnic_b200 import kernels as K  # noqa: E402

dev = torch.device("cuda:0")
B, H, S, hd = 16, 16, int(os.environ.get("S", 1024)), 64
reps = int(os.environ.get("REPS", 30))
d = H * hd
torch.manual_seed(0)
qkv = torch.randn(B * S, 3 * d, device=dev).bfloat16()
ld = qkv.stride(0)
heads = lambda t, c0: t.as_strided((B, H, S, hd), (S * ld, hd, ld, 1), c0)
q4, k4, v4 = heads(qkv, 2 * d), heads(qkv, 0), heads(qkv, d)
mask = torch.ones(B, S, dtype=torch.uint8, device=dev)
g = torch.Generator().manual_seed(1)
for b in range(1, B):
    mask[b, int(torch.randint(S // 2, S + 1, (1,), generator=g)):] = 0
kl = K.mask_key_len(mask)
dqkv = torch.empty_like(qkv)
dq4, dk4, dv4 = heads(dqkv, 2 * d), heads(dqkv, 0), heads(dqkv, d)
dO = torch.randn(B, S, d, device=dev).bfloat16()
out, stats = K.attn_fwd(q4, k4, v4, mask, kl, False)
K.attn_bwd(dO, out, stats, q4, k4, v4, dq4, dk4, dv4, mask, kl, False)
torch.cuda.synchronize()
flops = 4.0 * float((kl.float() * S).sum()) * hd * H  # algorithmic: only unmasked keys


def timeit(fn):
    """best of 3 rounds of `reps` back-to-back launches after 5 warm-up launches (CUDA events)"""
    for _ in range(5):
        fn()
    best = float("inf")
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / reps)
    return best


if os.environ.get("PACKED"):
    # the same work through the packed (varlen) interface: one [B*S, .] row buffer, per-sequence row ranges, no mask bytes
    q1, k1, v1 = (t.as_strided((1, H, B * S, hd), (B * S * ld, hd, ld, 1), t.storage_offset()) for t in (q4, k4, v4))
    dq1, dk1, dv1 = (t.as_strided((1, H, B * S, hd), (B * S * ld, hd, ld, 1), t.storage_offset()) for t in (dq4, dk4, dv4))
    start = (torch.arange(B, device=dev, dtype=torch.int32) * S)
    geo = K.Packed(start, kl.to(torch.int32), start, kl.to(torch.int32), S, S)
    outp, statsp = K.attn_fwd(q1, k1, v1, packed=geo)
    dO1 = dO.view(1, B * S, d)
    flops = 4.0 * float((kl.float() * kl.float()).sum()) * hd * H
    t_f = timeit(lambda: K.attn_fwd(q1, k1, v1, packed=geo))
    t_b = timeit(lambda: K.attn_bwd(dO1, outp, statsp, q1, k1, v1, dq1, dk1, dv1, packed=geo))
    print(f"PACKED S={S} fwd {t_f * 1e3:.1f} us ({flops / 1e12 / (t_f / 1e3):.0f} TF/s on len^2 work)   bwd(dq+dkv) {t_b * 1e3:.1f} us "
          f"({2.5 * flops / 1e12 / (t_b / 1e3):.0f} TF/s)")
    sys.exit(0)
t_f = timeit(lambda: K.attn_fwd(q4, k4, v4, mask, kl, False))
t_b = timeit(lambda: K.attn_bwd(dO, out, stats, q4, k4, v4, dq4, dk4, dv4, mask, kl, False))
t_f2 = timeit(lambda: K.attn_fwd(q4, k4, v4, mask, kl, False))
print(f"S={S} fwd {t_f * 1e3:.1f} / {t_f2 * 1e3:.1f} us ({flops / 1e12 / (min(t_f, t_f2) / 1e3):.0f} TF/s)   bwd(dq+dkv) {t_b * 1e3:.1f} us "
      f"({2.5 * flops / 1e12 / (t_b / 1e3):.0f} TF/s)")
