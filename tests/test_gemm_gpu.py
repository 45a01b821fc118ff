"""Parity of the tcgen05 GEMM family (vacnic_gemm) against fp32 torch.matmul on the same bf16 inputs.

Covers the four operand-major combinations used by forward / dgrad / wgrad / attention, ragged
M/N/K edges (TMA zero fill + predicated epilogue), every tile width, batched strided operands in
the [B,S,H,hd] attention layout, and each epilogue option."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _mk(shape, dev, scale=1.0, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(torch.bfloat16).to(dev)


def _ref(a, b, a_mn, b_mn):
    A = a.float().transpose(-1, -2) if a_mn else a.float()
    B = b.float().transpose(-1, -2) if b_mn else b.float()
    return A @ B.transpose(-1, -2)


@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, False), (True, True)])
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (256, 512, 256), (300, 200, 136), (16, 7680, 768), (80, 20, 80)])
@pytest.mark.parametrize("tile_n", [0, 64, 128, 256])
def test_gemm_majors_and_edges(cuda_device, a_mn, b_mn, M, N, K, tile_n):
    from vacnic_b200 import kernels as k
    # MN-major operands need their contiguous (M or N) extent to be a multiple of 8 for TMA strides
    if (a_mn and M % 8) or (b_mn and N % 8):
        pytest.skip("TMA row pitch must be a multiple of 16 bytes")
    a = _mk((K, M) if a_mn else (M, K), cuda_device, seed=1)
    b = _mk((K, N) if b_mn else (N, K), cuda_device, seed=2)
    out = k.gemm(a, b, a_mn=a_mn, b_mn=b_mn, out_dtype=torch.float32, tile_n=tile_n)
    torch.cuda.synchronize()
    ref = _ref(a, b, a_mn, b_mn)
    err = (out - ref).abs().max().item()
    assert err <= 2e-3 * max(1.0, ref.abs().max().item()), f"max abs err {err}"


def test_gemm_large_k_and_bf16_out(cuda_device):
    from vacnic_b200 import kernels as k
    a = _mk((1024, 4096), cuda_device, seed=3)
    b = _mk((1024, 4096), cuda_device, seed=4)
    out = k.gemm(a, b)
    torch.cuda.synchronize()
    ref = _ref(a, b, False, False)
    rel = ((out.float() - ref).abs().max() / ref.abs().max()).item()
    assert out.dtype == torch.bfloat16 and rel < 1e-2, rel


def test_gemm_epilogues(cuda_device):
    from vacnic_b200 import kernels as k
    M, N, K = 200, 328, 192
    a = _mk((M, K), cuda_device, 0.2, seed=5)
    b = _mk((N, K), cuda_device, 0.2, seed=6)
    bias = torch.randn(N, device=cuda_device)
    z_ref = (_ref(a, b, False, False) + bias) * 0.5
    # bias + alpha + gelu with pre-activation side output
    z = torch.empty(M, N, dtype=torch.bfloat16, device=cuda_device)
    y = k.gemm(a, b, bias=bias, alpha=0.5, act=k.ACT_GELU, aux_out=z, out_dtype=torch.float32)
    torch.cuda.synchronize()
    assert (z.float() - z_ref).abs().max().item() < 2e-2
    assert (y - torch.nn.functional.gelu(z_ref)).abs().max().item() < 5e-3
    # tanh
    y = k.gemm(a, b, bias=bias, act=k.ACT_TANH, out_dtype=torch.float32)
    assert (y - torch.tanh(_ref(a, b, False, False) + bias)).abs().max().item() < 5e-3
    # activation gradient epilogues: out = (A B^T) * act'(aux)
    zz = _mk((M, N), cuda_device, seed=7)
    y = k.gemm(a, b, aux_in=zz, dact=k.ACT_GELU, out_dtype=torch.float32)
    zf = zz.float().requires_grad_(True)
    torch.nn.functional.gelu(zf).sum().backward()
    assert (y - _ref(a, b, False, False) * zf.grad).abs().max().item() < 5e-3
    hh = torch.tanh(zz.float()).to(torch.bfloat16)
    y = k.gemm(a, b, aux_in=hh, dact=k.ACT_TANH, out_dtype=torch.float32)
    assert (y - _ref(a, b, False, False) * (1 - hh.float() ** 2)).abs().max().item() < 5e-3
    # accumulate into fp32 (wgrad) and bf16
    acc = torch.randn(M, N, device=cuda_device)
    want = acc + _ref(a, b, False, False)
    k.gemm(a, b, out=acc, accumulate=True)
    assert (acc - want).abs().max().item() < 5e-3


def test_gemm_attention_layouts(cuda_device):
    """QK^T, PV and the wgrad-like dV = P^T dO on [B,S,H,hd] strided views, two batch dims."""
    from vacnic_b200 import kernels as k
    B, H, Sq, Sk, hd = 3, 4, 200, 136, 64
    qkv = _mk((B, Sq, 3 * H * hd), cuda_device, 0.5, seed=8)
    kv = _mk((B, Sk, 2 * H * hd), cuda_device, 0.5, seed=9)
    q = qkv[..., : H * hd].view(B, Sq, H, hd).permute(0, 2, 1, 3)          # [B,H,Sq,hd] strided
    kk = kv[..., : H * hd].view(B, Sk, H, hd).permute(0, 2, 1, 3)
    v = kv[..., H * hd:].view(B, Sk, H, hd).permute(0, 2, 1, 3)
    s = k.gemm(q, kk, out_dtype=torch.float32)                              # [B,H,Sq,Sk]
    s_ref = q.float() @ kk.float().transpose(-1, -2)
    assert (s - s_ref).abs().max().item() < 2e-3 * s_ref.abs().max().item()
    p = torch.softmax(s_ref, -1).to(torch.bfloat16)
    o = torch.empty(B, Sq, H, hd, dtype=torch.bfloat16, device=cuda_device)
    k.gemm(p, v, out=o.permute(0, 2, 1, 3), b_mn=True)                      # O = P V, V is [Sk,hd]
    o_ref = (p.float() @ v.float()).permute(0, 2, 1, 3)
    assert (o.float() - o_ref).abs().max().item() < 1e-2
    do = _mk((B, Sq, H, hd), cuda_device, seed=10).permute(0, 2, 1, 3)
    dv = k.gemm(p, do, a_mn=True, b_mn=True, out_dtype=torch.float32)        # dV = P^T dO  [B,H,Sk,hd]
    dv_ref = p.float().transpose(-1, -2) @ do.float()
    assert (dv - dv_ref).abs().max().item() < 2e-3 * dv_ref.abs().max().item()


def test_gemm_rejects_cpu_tensors():
    from vacnic_b200 import kernels as k
    from vacnic_b200.lib import VacnicError
    a = torch.zeros(8, 8, dtype=torch.bfloat16)
    with pytest.raises(VacnicError):
        k.gemm(a, a)


@pytest.mark.gpu
def test_gemm_head_major_output_and_broadcast_weight(cuda_device):
    """Batched GEMM with one weight matrix broadcast over the batch (zero batch stride) whose output columns are
    scattered head-major: [batch][64-column group][rows][64] — the decode-time cross K/V cache layout."""
    from vacnic_b200 import kernels as K
    torch.manual_seed(1)
    C, L, d, H = 3, 200, 256, 8
    h = torch.randn(C, L, d, device=cuda_device).bfloat16()
    w = (torch.randn(2 * H * 64, d, device=cuda_device) * 0.1).bfloat16()
    bias = torch.randn(2 * H * 64, device=cuda_device)
    out = torch.zeros(C, 2, H, L, 64, device=cuda_device, dtype=torch.bfloat16)
    K.gemm(h, w.unsqueeze(0).expand(C, -1, -1), out=out, bias=bias, head_major=(64, 2 * H * 64 * L, 0, L * 64))
    ref = (h.float() @ w.float().t() + bias).view(C, L, 2, H, 64).permute(0, 2, 3, 1, 4)
    assert (out.float() - ref).abs().max().item() <= 3e-2


# ---------------------------------------------------------------------------------------------- CTA-pair kernel
@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, False), (True, True)])
@pytest.mark.parametrize("M,N,K", [(256, 256, 64), (512, 512, 320), (1000, 392, 136), (2048, 1024, 1024), (264, 128, 72)])
@pytest.mark.parametrize("tile_n", [1128, 1256])
def test_gemm_cta_pair_kernel(cuda_device, a_mn, b_mn, M, N, K, tile_n):
    """tcgen05.mma.cta_group::2 path (256 x BN tiles over a 2-CTA cluster), forced through tile_n = 1000 + BN."""
    from vacnic_b200 import kernels as k
    a = _mk((K, M) if a_mn else (M, K), cuda_device, seed=11)
    b = _mk((K, N) if b_mn else (N, K), cuda_device, seed=12)
    bias = torch.randn(N, device=cuda_device)
    out = k.gemm(a, b, a_mn=a_mn, b_mn=b_mn, bias=bias, out_dtype=torch.float32, tile_n=tile_n)
    torch.cuda.synchronize()
    ref = _ref(a, b, a_mn, b_mn) + bias
    err = (out - ref).abs().max().item()
    assert err <= 2e-3 * max(1.0, ref.abs().max().item()), f"max abs err {err}"


def test_gemm_cta_pair_batched_and_many_tiles(cuda_device):
    """More tiles than SM pairs (persistent loop, accumulator double buffering) and batched strided operands."""
    from vacnic_b200 import kernels as k
    a = _mk((3, 1280, 256), cuda_device, 0.3, seed=13)
    b = _mk((3, 2304, 256), cuda_device, 0.3, seed=14)
    for tn in (0, 1256, 1128):
        out = k.gemm(a, b, out_dtype=torch.bfloat16, tile_n=tn)
        ref = a.float() @ b.float().transpose(-1, -2)
        rel = ((out.float() - ref).abs().max() / ref.abs().max()).item()
        assert rel < 1e-2, (tn, rel)
    # the auto heuristic must agree with the single-CTA kernel on a large GELU forward (epilogue shared)
    x = _mk((4096, 1024), cuda_device, 0.5, seed=15)
    w = _mk((4096, 1024), cuda_device, 0.05, seed=16)
    bias = torch.randn(4096, device=cuda_device)
    y_pair = k.gemm(x, w, bias=bias, act=k.ACT_GELU, tile_n=1256)
    y_one = k.gemm(x, w, bias=bias, act=k.ACT_GELU, tile_n=256)
    assert torch.equal(y_pair, y_one)


@pytest.mark.parametrize("M,N,K,a_mn,b_mn", [(256, 1024, 4096, False, False), (64, 3072, 1024, False, False),
                                              (320, 4096, 1024, False, False), (1024, 1024, 1024, False, True),
                                              (200, 328, 1000, False, False), (1024, 1024, 1280, True, True)])
def test_gemm_split_k_is_deterministic_and_matches(cuda_device, M, N, K, a_mn, b_mn):
    """Skinny problems (few tiles, long K) can opt into split-K (workspace given); the fixed-order reduction
    must reproduce the unsplit kernel up to fp32 summation order and be bit-reproducible run to run."""
    from vacnic_b200 import kernels as k
    a = _mk((K, M) if a_mn else (M, K), cuda_device, 0.3, seed=21)
    b = _mk((K, N) if b_mn else (N, K), cuda_device, 0.3, seed=22)
    bias = torch.randn(N, device=cuda_device)
    ref = _ref(a, b, a_mn, b_mn) + bias
    outs = [k.gemm(a, b, a_mn=a_mn, b_mn=b_mn, bias=bias, out_dtype=torch.float32, split_k=True) for _ in range(3)]
    unsplit = k.gemm(a, b, a_mn=a_mn, b_mn=b_mn, bias=bias, out_dtype=torch.float32, tile_n=64)
    torch.cuda.synchronize()
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[1], outs[2])
    scale = max(1.0, ref.abs().max().item())
    assert (outs[0] - ref).abs().max().item() <= 2e-3 * scale
    assert (outs[0] - unsplit).abs().max().item() <= 1e-4 * scale
    # gelu + bf16 output through the split path
    y = k.gemm(a, b, a_mn=a_mn, b_mn=b_mn, bias=bias, act=k.ACT_GELU, split_k=True)
    yref = torch.nn.functional.gelu(ref)
    assert (y.float() - yref).abs().max().item() <= 2e-2 * max(1.0, yref.abs().max().item())


@pytest.mark.parametrize("tile_n", [1128, 1256])
@pytest.mark.parametrize("M,N,K", [(1000, 392, 136), (264, 416, 72), (776, 1024, 256)])
def test_gemm_cta_pair_bf16_epilogues(cuda_device, M, N, K, tile_n):
    """bf16 results of the CTA-pair kernel leave through the warp-cooperative coalesced store: ragged M (partial warps),
    ragged N (scalar tail chunk), pre-activation side output, activation gradient, accumulate (thread-owns-row path)."""
    from vacnic_b200 import kernels as k
    a = _mk((M, K), cuda_device, 0.3, seed=31)
    b = _mk((N, K), cuda_device, 0.3, seed=32)
    bias = torch.randn(N, device=cuda_device)
    pre = (a.float() @ b.float().t() + bias) * 0.5   # bias is added before alpha
    # canary-framed outputs: nothing outside [M, N] may be written
    big = torch.full((M + 8, N + 8), 7.0, device=cuda_device, dtype=torch.bfloat16)
    zbig = torch.full((M + 8, N + 8), 7.0, device=cuda_device, dtype=torch.bfloat16)
    y, z = big[:M, :N], zbig[:M, :N]
    k.gemm(a, b, out=y, bias=bias, alpha=0.5, act=k.ACT_GELU, aux_out=z, tile_n=tile_n)
    torch.cuda.synchronize()
    assert (z.float() - pre).abs().max().item() <= 1e-2 * max(1.0, pre.abs().max().item())
    assert (y.float() - torch.nn.functional.gelu(pre)).abs().max().item() <= 1e-2 * max(1.0, pre.abs().max().item())
    assert (big[M:] == 7).all() and (big[:, N:] == 7).all() and (zbig[M:] == 7).all() and (zbig[:, N:] == 7).all()
    # same numbers as the single-CTA kernel (shared arithmetic, different store path)
    y1 = k.gemm(a, b, bias=bias, alpha=0.5, act=k.ACT_GELU, tile_n=128)
    assert torch.equal(y, y1)
    # activation gradient + tanh
    dz = k.gemm(a, b, aux_in=z.contiguous(), dact=k.ACT_GELU, tile_n=tile_n)
    zf = z.float().requires_grad_(True)
    torch.nn.functional.gelu(zf).sum().backward()
    ref = (a.float() @ b.float().t()) * zf.grad
    assert (dz.float() - ref).abs().max().item() <= 1e-2 * max(1.0, ref.abs().max().item())
    t = k.gemm(a, b, bias=bias, act=k.ACT_TANH, tile_n=tile_n)
    tref = torch.tanh(a.float() @ b.float().t() + bias)
    assert (t.float() - tref).abs().max().item() <= 1e-2
    # accumulate into bf16
    acc = torch.ones(M, N, device=cuda_device, dtype=torch.bfloat16)
    k.gemm(a, b, out=acc, accumulate=True, tile_n=tile_n)
    ref = a.float() @ b.float().t() + 1.0
    assert (acc.float() - ref).abs().max().item() <= 1e-2 * max(1.0, ref.abs().max().item())


def test_gemm_cta_pair_head_major(cuda_device):
    """Head-major scattered output (decode cross-K/V layout) through the CTA-pair kernel's coalesced store."""
    from vacnic_b200 import kernels as K
    torch.manual_seed(2)
    C, L, d, H = 2, 520, 256, 8
    h = torch.randn(C, L, d, device=cuda_device).bfloat16()
    w = (torch.randn(2 * H * 64, d, device=cuda_device) * 0.1).bfloat16()
    bias = torch.randn(2 * H * 64, device=cuda_device)
    ref = (h.float() @ w.float().t() + bias).view(C, L, 2, H, 64).permute(0, 2, 3, 1, 4)
    for tn in (1256, 1128):
        out = torch.zeros(C, 2, H, L, 64, device=cuda_device, dtype=torch.bfloat16)
        K.gemm(h, w.unsqueeze(0).expand(C, -1, -1), out=out, bias=bias, head_major=(64, 2 * H * 64 * L, 0, L * 64), tile_n=tn)
        assert (out.float() - ref).abs().max().item() <= 3e-2, tn


def test_wgrad_k_slices_batched_matches(cuda_device):
    """Weight gradient of a 1024 x 1024 projection over 16384 tokens: K-sliced batched CTA-pair GEMM + ordered partial
    sum (blocks._wgrad) against fp32 torch; bit-reproducible; accumulates into an existing gradient."""
    from vacnic_b200 import blocks, kernels as k
    rows, n_out, k_in = 16384, 1024, 1024
    s = blocks.wgrad_k_slices(rows, n_out, k_in, k.sm_count(cuda_device))
    assert s == 4
    dy = _mk((rows, 3 * n_out), cuda_device, 0.1, seed=41)[:, n_out:2 * n_out]   # a column slice: row pitch 3072
    x = _mk((rows, k_in), cuda_device, 0.1, seed=42)
    ref = dy.float().t() @ x.float()
    outs = []
    for _ in range(2):
        part = torch.empty(s, n_out, k_in, dtype=torch.float32, device=cuda_device)
        k.gemm(dy.view(s, rows // s, n_out), x.view(s, rows // s, k_in), out=part, a_mn=True, b_mn=True)
        gw = torch.ones(n_out, k_in, device=cuda_device)
        k.sum_partials(part.view(s, -1), gw, accumulate=True)
        outs.append(gw)
    torch.cuda.synchronize()
    assert torch.equal(outs[0], outs[1])
    assert (outs[0] - 1.0 - ref).abs().max().item() <= 2e-3 * ref.abs().max().item()
