"""Decode-side kernels alone at the bench shape (64 captions x 4 beams, L=1024, BART-large):
cross-attention bandwidth and the M=256 GEMMs replayed from a CUDA graph (no host overhead)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vacnic_b200 import kernels as K  # noqa: E402

dev = torch.device("cuda:0")
C, nb, L, H, d, f = 64, 4, 1024, 16, 1024, 4096
R = C * nb
torch.manual_seed(0)
nl = 12
kv = torch.randn(nl, C, 2, H, L, 64, device=dev).bfloat16()  # head-major, as the generator lays it out
q = torch.randn(R, d, device=dev).bfloat16()
out = torch.empty(R, d, device=dev, dtype=torch.bfloat16)
mask = torch.ones(C, L, dtype=torch.uint8, device=dev)
g = torch.Generator().manual_seed(1)
for c in range(1, C):
    mask[c, int(torch.randint(L // 2, L + 1, (1,), generator=g)):] = 0
kl = K.mask_key_len(mask)
bytes_per = float(kl.sum()) * 2 * d * 2


def graph_time(fn, reps=5):
    fn(); torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        fn()
    gr.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        gr.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def cross_all():
    for l in range(nl):
        K.decode_cross_attn(q, kv[l, :, 0], kv[l, :, 1], mask, kl, out, nb)


t = graph_time(cross_all) / nl
print(f"decode_cross_attn: {t * 1e3:.1f} us/launch, {bytes_per / 1e9 / (t / 1e3):.0f} GB/s")

x = torch.randn(R, d, device=dev).bfloat16()
hbuf = torch.randn(R, f, device=dev).bfloat16()
ws = {n: [torch.randn(s, device=dev).bfloat16() for _ in range(nl)] for n, s in
      (("qkv", (3 * d, d)), ("o", (d, d)), ("fc1", (f, d)), ("fc2", (d, f)))}
outs = {"qkv": torch.empty(R, 3 * d, device=dev, dtype=torch.bfloat16), "o": torch.empty(R, d, device=dev, dtype=torch.bfloat16),
        "fc1": torch.empty(R, f, device=dev, dtype=torch.bfloat16), "fc2": torch.empty(R, d, device=dev, dtype=torch.bfloat16)}
bias = {n: torch.zeros(w[0].shape[0], device=dev) for n, w in ws.items()}
for n in ws:
    a = hbuf if n == "fc2" else x

    def run(n=n, a=a):
        for l in range(nl):
            K.gemm(a, ws[n][l], out=outs[n], bias=bias[n])
    t = graph_time(run) / nl
    wb = ws[n][0].numel() * 2
    print(f"decode gemm {n:4s} M={R} N={ws[n][0].shape[0]} K={ws[n][0].shape[1]}: {t * 1e3:.1f} us, weights {wb / 1e9 / (t / 1e3):.0f} GB/s")

# the fused tcgen05 attention kernel used for the decode cross-attention (Sq = beams)
q4 = q.view(C, nb, H, 64).permute(0, 2, 1, 3)
for l in range(1):
    o, _ = K.attn_fwd(q4, kv[l, :, 0], kv[l, :, 1], mask, kl, False, want_stats=False)
K.decode_cross_attn(q, kv[0, :, 0], kv[0, :, 1], mask, kl, out, nb)
print("fused vs streaming max diff", (o.view(R, d).float() - out.float()).abs().max().item())


def fused_all():
    for l in range(nl):
        K.attn_fwd(q4, kv[l, :, 0], kv[l, :, 1], mask, kl, False, want_stats=False)


t = graph_time(fused_all) / nl
print(f"attn_fwd as decode cross-attention: {t * 1e3:.1f} us/launch, {bytes_per / 1e9 / (t / 1e3):.0f} GB/s (algorithmic bytes)")
