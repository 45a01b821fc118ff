"""Parity of the bandwidth-class kernels (LayerNorm family, embeddings, softmax, reductions, AdamW,
CE / CoLaM / SECLA) against fp32 torch on the same inputs.  bf16 outputs: tolerance = bf16 rounding of
the result (2^-8 relative) plus accumulated input rounding, stated per assert."""
import pytest
import torch
import torch.nn.functional as F

from oracle import model as OM

pytestmark = pytest.mark.gpu


def rnd(shape, dev, scale=1.0, seed=0, dtype=torch.bfloat16):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(dtype).to(dev)


@pytest.mark.parametrize("d", [768, 1024])
@pytest.mark.parametrize("rows", [1, 37, 4096])
@pytest.mark.parametrize("with_res", [True, False])
def test_add_layernorm_fwd_bwd(cuda_device, d, rows, with_res):
    from vacnic_b200 import kernels as k
    x = rnd((rows, d), cuda_device, 1.0, 1)
    res = rnd((rows, d), cuda_device, 1.0, 2) if with_res else None
    gamma = 1 + 0.1 * rnd((d,), cuda_device, 1.0, 3, torch.float32)
    beta = 0.1 * rnd((d,), cuda_device, 1.0, 4, torch.float32)
    y, mean, rstd = k.add_layernorm_fwd(x, res, gamma, beta)
    xs = (x.float() + (res.float() if with_res else 0)).requires_grad_(True)
    gref, bref = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    yr = F.layer_norm(xs, (d,), gref, bref, 1e-5)
    assert (y.float() - yr).abs().max().item() <= 2 ** -7 * max(1.0, yr.abs().max().item())
    dy = rnd((rows, d), cuda_device, 1.0, 5)
    yr.backward(dy.float())
    dgamma = torch.zeros(d, device=cuda_device)
    dbeta = torch.zeros(d, device=cuda_device)
    dbias = torch.zeros(d, device=cuda_device)
    dsum, dx = k.add_layernorm_bwd(dy, x, res, gamma, mean, rstd, dgamma, dbeta, dbias)
    torch.cuda.synchronize()
    assert dx is dsum
    assert (dsum.float() - xs.grad).abs().max().item() <= 2 ** -6 * max(1.0, xs.grad.abs().max().item())
    tol = 2e-2 * max(1.0, rows ** 0.5)
    assert (dgamma - gref.grad).abs().max().item() <= tol
    assert (dbeta - bref.grad).abs().max().item() <= tol
    assert (dbias - dsum.float().sum(0)).abs().max().item() <= tol


def test_add_layernorm_strided_output_and_dropout(cuda_device):
    from vacnic_b200 import kernels as k
    B, P, G, d = 3, 5, 7, 1024
    x = rnd((B * P, d), cuda_device, 1.0, 1)
    res = rnd((B * P, d), cuda_device, 1.0, 2)
    gamma = torch.ones(d, device=cuda_device)
    beta = torch.zeros(d, device=cuda_device)
    cat = torch.zeros(B, P + G, d, dtype=torch.bfloat16, device=cuda_device)
    y, mean, rstd = k.add_layernorm_fwd(x, res, gamma, beta, out=cat, rows_per_group=P, group_stride=(P + G) * d)
    ref = F.layer_norm(x.float() + res.float(), (d,))
    assert (cat[:, :P].float().reshape(B * P, d) - ref).abs().max().item() < 2 ** -6
    assert cat[:, P:].abs().max().item() == 0
    # backward reading dy out of the same strided slice
    dcat = rnd((B, P + G, d), cuda_device, 1.0, 6)
    dgamma = torch.zeros(d, device=cuda_device); dbeta = torch.zeros(d, device=cuda_device)
    dsum, _ = k.add_layernorm_bwd(dcat, x, res, gamma, mean, rstd, dgamma, dbeta, rows_per_group=P, group_stride=(P + G) * d)
    xs = (x.float() + res.float()).requires_grad_(True)
    F.layer_norm(xs, (d,)).backward(dcat[:, :P].float().reshape(B * P, d))
    assert (dsum.float() - xs.grad).abs().max().item() <= 2 ** -5
    # dropout: mask is reproducible between forward and backward and has the right keep rate
    rng = k.Rng(cuda_device, seed=3)
    p = 0.1
    big = rnd((4096, d), cuda_device, 1.0, 7)
    zero_res = torch.zeros_like(big)
    # LN is scale invariant per row, so recover the mask from a second call with p = 0 instead:
    y1, m1, r1 = k.add_layernorm_fwd(big, zero_res, gamma, beta, p_drop=p, rng=rng, salt=11)
    y2, _, _ = k.add_layernorm_fwd(big, zero_res, gamma, beta, p_drop=p, rng=rng, salt=11)
    assert torch.equal(y1, y2)
    y3, _, _ = k.add_layernorm_fwd(big, zero_res, gamma, beta, p_drop=p, rng=rng, salt=12)
    assert not torch.equal(y1, y3)
    dy = torch.ones_like(big)
    dg = torch.zeros(d, device=cuda_device); db = torch.zeros(d, device=cuda_device)
    dsum, dx = k.add_layernorm_bwd(dy * 0 + rnd((4096, d), cuda_device, 1.0, 8), big, zero_res, gamma, m1, r1, dg, db,
                                   want_dx=True, p_drop=p, rng=rng, salt=11)
    dropped = (dx == 0) & (dsum != 0)
    rate = dropped.float().mean().item()
    assert abs(rate - p) < 0.01, rate
    kept = ~dropped
    assert (dx[kept].float() - dsum[kept].float() / (1 - p)).abs().max().item() <= 2 ** -6 * dsum.abs().max().item()


@pytest.mark.parametrize("d", [768, 1024])
def test_embed_ln_fwd_bwd(cuda_device, d):
    from vacnic_b200 import kernels as k
    V, S, B = 300, 24, 5
    tok32 = rnd((V, d), cuda_device, 0.02, 1, torch.float32)
    pos32 = rnd((S + 10, d), cuda_device, 0.02, 2, torch.float32)
    tok, pos = tok32.to(torch.bfloat16), pos32.to(torch.bfloat16)
    gamma = 1 + 0.1 * rnd((d,), cuda_device, 1.0, 3, torch.float32)
    beta = 0.1 * rnd((d,), cuda_device, 1.0, 4, torch.float32)
    ids = torch.randint(0, V, (B, S), generator=torch.Generator().manual_seed(5)).to(cuda_device)
    ids[:, -3:] = 1
    y, mean, rstd = k.embed_ln_fwd(ids, tok, pos, gamma, beta, pos_offset=5)
    tr, pr = tok.float().requires_grad_(True), pos.float().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    h = F.embedding(ids, tr, padding_idx=1) + pr[torch.arange(S, device=cuda_device) + 5]
    yr = F.layer_norm(h, (d,), gr, br, 1e-5)
    assert (y.float() - yr).abs().max().item() <= 2 ** -6 * max(1.0, yr.abs().max().item())
    dy = rnd((B, S, d), cuda_device, 1.0, 6)
    yr.backward(dy.float())
    dtok = torch.zeros(V, d, device=cuda_device); dpos = torch.zeros(S + 10, d, device=cuda_device)
    dg = torch.zeros(d, device=cuda_device); db = torch.zeros(d, device=cuda_device)
    k.embed_ln_bwd(dy, ids, tok, pos, gamma, mean, rstd, dtok, dpos, dg, db, pos_offset=5, pad_id=1)
    torch.cuda.synchronize()
    for got, want in ((dtok, tr.grad), (dpos, pr.grad), (dg, gr.grad), (db, br.grad)):
        assert (got - want).abs().max().item() <= 2e-2 * max(1.0, want.abs().max().item())
    assert dtok[1].abs().max().item() == 0  # padding_idx row receives no gradient


def test_names_embed(cuda_device):
    from vacnic_b200 import kernels as k
    d, V = 768, 400
    tok = rnd((V, d), cuda_device, 0.02, 1); pos = rnd((40, d), cuda_device, 0.02, 2)
    gamma = 1 + 0.1 * rnd((d,), cuda_device, 1.0, 3, torch.float32)
    beta = 0.1 * rnd((d,), cuda_device, 1.0, 4, torch.float32)
    ids = torch.randint(0, V, (3, 5, 8), generator=torch.Generator().manual_seed(5)).to(cuda_device)
    out = k.names_embed(ids, tok, pos, gamma, beta)
    sd = {"e.embed_tokens_ner.weight": tok.float(), "e.embed_positions_ner.weight": pos.float(),
          "e.layernorm_embedding_ner.weight": gamma, "e.layernorm_embedding_ner.bias": beta}
    ref = OM.names_embedding(sd, ids, prefix="e.")
    assert (out - ref).abs().max().item() <= 1e-4


@pytest.mark.parametrize("Sq,Sk,causal,masked", [(80, 84, False, True), (64, 64, True, False), (33, 30, False, False),
                                                 (128, 1024, False, True), (7, 12, True, True)])
def test_softmax_fwd_bwd(cuda_device, Sq, Sk, causal, masked):
    from vacnic_b200 import kernels as k
    B, H = 2, 3
    ld = (Sk + 7) // 8 * 8
    s = torch.zeros(B, H, Sq, ld, device=cuda_device)
    s[..., :Sk] = rnd((B, H, Sq, Sk), cuda_device, 2.0, 1, torch.float32)
    km = None
    add = torch.zeros(B, 1, Sq, Sk, device=cuda_device)
    if masked:
        km = torch.ones(B, Sk, dtype=torch.uint8, device=cuda_device)
        km[0, Sk // 2:] = 0
        km[1, -1] = 0
        add = add + OM.expand_mask(km.long(), torch.float32, Sq)
    if causal:
        past = Sk - Sq if Sk >= Sq else 0
        add = add + OM.causal_mask(Sq, torch.float32, cuda_device, past)[..., :Sk]
    else:
        past = 0
    p = k.softmax_fwd(s, km, Sk, causal=causal, past=past)
    sr = s[..., :Sk].clone().requires_grad_(True)
    pr = torch.softmax(sr + add, dim=-1)
    assert (p[..., :Sk].float() - pr).abs().max().item() <= 2 ** -8
    assert p[..., Sk:].abs().max().item() == 0 if ld > Sk else True
    dp = torch.zeros(B, H, Sq, ld, device=cuda_device)
    dp[..., :Sk] = rnd((B, H, Sq, Sk), cuda_device, 1.0, 2, torch.float32)
    p.float()[..., :Sk].detach()
    # reference gradient evaluated at the bf16-rounded probabilities the kernel saw
    pb = p[..., :Sk].float()
    want = pb * (dp[..., :Sk] - (pb * dp[..., :Sk]).sum(-1, keepdim=True))
    ds = k.softmax_bwd(p, dp, Sk)
    assert (ds[..., :Sk].float() - want).abs().max().item() <= 2 ** -7 * max(1.0, want.abs().max().item())


def test_colsum_cast_add_adamw(cuda_device):
    from vacnic_b200 import kernels as k
    x = rnd((1000, 4096 + 24), cuda_device, 1.0, 1)[:, :4096 + 3]  # ragged width, strided rows
    out = torch.ones(4099, device=cuda_device)
    k.colsum_into(x, out)
    assert (out - 1 - x.float().sum(0)).abs().max().item() <= 1e-2
    src = rnd((100003,), cuda_device, 1.0, 2, torch.float32)
    dst = torch.empty(100003, dtype=torch.bfloat16, device=cuda_device)
    k.cast_bf16(src, dst)
    assert torch.equal(dst, src.to(torch.bfloat16))
    a, b, c = (rnd((5000,), cuda_device, 1.0, s) for s in (3, 4, 5))
    assert (k.add_bf16(a, b).float() - (a.float() + b.float())).abs().max().item() <= 2 ** -6
    assert (k.add_bf16(a, b, c).float() - (a.float() + b.float() + c.float())).abs().max().item() <= 2 ** -5
    # AdamW against torch.optim.AdamW over three steps
    n = 10007
    p0 = rnd((n,), cuda_device, 1.0, 6, torch.float32)
    pt = torch.nn.Parameter(p0.clone())
    opt = torch.optim.AdamW([pt], lr=3e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01)
    p = p0.clone(); m = torch.zeros_like(p); v = torch.zeros_like(p)
    p16 = torch.empty(n, dtype=torch.bfloat16, device=cuda_device)
    for step in range(1, 4):
        g = rnd((n,), cuda_device, 1.0, 10 + step, torch.float32)
        pt.grad = g.clone()
        opt.step()
        hyper = torch.tensor([3e-3, 0.9, 0.999, 1e-8, 0.01, 1 - 0.9 ** step, 1 - 0.999 ** step, 1.0], device=cuda_device)
        k.adamw(p, g, m, v, p16, hyper)
    assert (p - pt.data).abs().max().item() <= 1e-5
    assert torch.equal(p16, p.to(torch.bfloat16))


def test_ce_fwd_bwd(cuda_device):
    from vacnic_b200 import kernels as k
    rows, V, ld = 50, 50267, 50272
    buf = torch.zeros(rows, ld, device=cuda_device)
    buf[:, :V] = rnd((rows, V), cuda_device, 3.0, 1, torch.float32)
    tgt = torch.randint(0, V, (rows,), generator=torch.Generator().manual_seed(2)).to(cuda_device)
    tgt[::7] = 1
    out, lse, row_loss = k.ce_fwd(buf[:, :V], V, tgt, ignore_index=1)
    lr = buf[:, :V].clone().requires_grad_(True)
    ref = F.cross_entropy(lr, tgt, ignore_index=1)
    assert abs(out[0].item() - ref.item()) <= 1e-5 * abs(ref.item())
    assert out[1].item() == (tgt != 1).sum().item()
    (ref * 0.5).backward()
    dl = torch.empty(rows, ld, dtype=torch.bfloat16, device=cuda_device)
    gs = torch.tensor([0.5], device=cuda_device)
    k.ce_bwd(buf[:, :V], V, lse, tgt, out, gs, 1.0, dl, ignore_index=1)
    assert (dl[:, :V].float() - lr.grad).abs().max().item() <= 2 ** -8 * lr.grad.abs().max().item() + 1e-9
    assert dl[:, V:].abs().max().item() == 0


@pytest.mark.parametrize("margin", [1.0, 0.3])
def test_colam(cuda_device, margin):
    from vacnic_b200 import kernels as k
    B, T, d = 6, 20, 1024
    h = rnd((B, T, d), cuda_device, 1.0, 1)
    hg = (0.7 * h.float() + 0.7 * rnd((B, T, d), cuda_device, 1.0, 2).float()).to(torch.bfloat16)
    tgt = torch.randint(3, 100, (B, T), generator=torch.Generator().manual_seed(3)).to(cuda_device)
    tgt[0, 5:] = 1
    tgt[3, 12:] = 1
    loss, pa, pb, stats = k.colam_fwd(h, hg, tgt, margin)
    hr = h.float().requires_grad_(True)
    ref = OM.colam_loss(hr, hg.float(), tgt, margin)
    assert abs(loss.item() - ref.item()) <= 1e-5
    (ref * 0.5).backward()
    dh = torch.empty_like(h)
    k.colam_bwd(pa, pb, stats, tgt, torch.tensor([1.0], device=cuda_device), 0.5, dh)
    assert (dh.float() - hr.grad).abs().max().item() <= 2 ** -7 * hr.grad.abs().max().item() + 1e-9


def test_secla(cuda_device):
    from vacnic_b200 import kernels as k
    B, N, Fc, d = 16, 8, 4, 1024
    names = rnd((B, N, d), cuda_device, 0.5, 1, torch.float32)
    face = rnd((B, Fc, d), cuda_device, 0.5, 2)
    loss, ws = k.secla_fwd(names, face)
    fr = face.float().requires_grad_(True)
    ref = OM.secla_loss(fr, names)
    assert abs(loss.item() - ref.item()) <= 1e-4 * abs(ref.item())
    ref.backward()
    dface = torch.empty_like(face)
    k.secla_bwd(ws, names, None, 1.0, dface)
    assert (dface.float() - fr.grad).abs().max().item() <= 2 ** -7 * fr.grad.abs().max().item() + 1e-9


# ---------------------------------------------------------------------------------------------- fused attention
def _ref_attention(q4, k4, v4, key_mask, causal):
    B, H, Sq, hd = q4.shape
    Sk = k4.shape[2]
    s = (q4.float() @ k4.float().transpose(-1, -2)) * hd ** -0.5
    neg = torch.finfo(torch.float32).min
    if key_mask is not None:
        s = s + (key_mask[:, None, None, :] == 0).float() * neg
    if causal:
        s = s + torch.triu(torch.ones(Sq, Sk, device=s.device), 1) * neg
    p = torch.softmax(s, -1)
    return (p @ v4.float()).permute(0, 2, 1, 3).reshape(B, Sq, H * hd), p


ATTN_CASES = [  # B, H, Sq, Sk, causal, masked
    (2, 12, 48, 48, False, True), (2, 16, 1024, 1024, False, True), (2, 16, 1024, 40, False, False),
    (3, 12, 80, 84, False, True), (2, 16, 64, 64, True, True), (2, 16, 64, 1024, False, True),
    (1, 12, 300, 200, False, True), (2, 12, 130, 130, True, False), (1, 16, 40, 512, False, True),
]


@pytest.mark.parametrize("B,H,Sq,Sk,causal,masked", ATTN_CASES)
def test_attn_fwd_matches_torch(cuda_device, B, H, Sq, Sk, causal, masked):
    from vacnic_b200 import kernels as K
    torch.manual_seed(B * 1000 + Sq + Sk)
    dev = cuda_device
    d = H * 64
    # q/k/v as strided head views of fused projection buffers, as the model uses them
    qkv = (torch.randn(B * Sq, 3 * d, device=dev) * 1.5).bfloat16()
    kvb = (torch.randn(B * Sk, 2 * d, device=dev) * 1.5).bfloat16()

    def heads(t, S, col0):
        ld = t.stride(0)
        return t.as_strided((B, H, S, 64), (S * ld, 64, ld, 1), t.storage_offset() + col0)

    q4, k4, v4 = heads(qkv, Sq, 2 * d), heads(kvb, Sk, 0), heads(kvb, Sk, d)
    key_mask = None
    if masked:
        key_mask = torch.ones(B, Sk, dtype=torch.uint8, device=dev)
        for b in range(B):
            key_mask[b, torch.randint(max(1, Sk // 2), Sk + 1, (1,)).item():] = 0
        if Sk > 16:
            key_mask[0, 5] = 0
    ref, _ = _ref_attention(q4, k4, v4, key_mask, causal)
    for use_len in (False, True):
        kl = K.mask_key_len(key_mask) if (use_len and key_mask is not None) else None
        out, stats = K.attn_fwd(q4, k4, v4, key_mask, kl, causal)
        err = (out.float() - ref).abs().max().item()
        assert err <= 3e-2, (err, use_len)
        assert (out.float() - ref).abs().mean().item() <= 2e-3


def test_attn_fwd_fully_masked_row_is_uniform(cuda_device):
    from vacnic_b200 import kernels as K
    torch.manual_seed(0)
    B, H, S = 2, 12, 70
    q = torch.randn(B, H, S, 64, device=cuda_device).bfloat16()
    k = torch.randn(B, H, S, 64, device=cuda_device).bfloat16()
    v = torch.randn(B, H, S, 64, device=cuda_device).bfloat16()
    mask = torch.ones(B, S, dtype=torch.uint8, device=cuda_device)
    mask[1] = 0  # the reference adds finfo.min everywhere -> uniform attention over all keys (MFULL:387-398)
    ref, _ = _ref_attention(q, k, v, mask, False)
    out, _ = K.attn_fwd(q, k, v, mask, K.mask_key_len(mask), False)
    assert (out.float() - ref).abs().max().item() <= 3e-2


@pytest.mark.parametrize("B,H,Sq,Sk,causal,masked", ATTN_CASES)
def test_attn_bwd_matches_torch_autograd(cuda_device, B, H, Sq, Sk, causal, masked):
    from vacnic_b200 import kernels as K
    torch.manual_seed(B * 77 + Sq + 3 * Sk)
    dev = cuda_device
    d = H * 64
    qkv = (torch.randn(B * Sq, 3 * d, device=dev) * 1.2).bfloat16()
    kvb = (torch.randn(B * Sk, 2 * d, device=dev) * 1.2).bfloat16()

    def heads(t, S, col0):
        ld = t.stride(0)
        return t.as_strided((B, H, S, 64), (S * ld, 64, ld, 1), t.storage_offset() + col0)

    q4, k4, v4 = heads(qkv, Sq, 2 * d), heads(kvb, Sk, 0), heads(kvb, Sk, d)
    key_mask = None
    if masked:
        key_mask = torch.ones(B, Sk, dtype=torch.uint8, device=dev)
        for b in range(B):
            key_mask[b, torch.randint(max(1, Sk // 2), Sk + 1, (1,)).item():] = 0
        if Sk > 16:
            key_mask[0, 5] = 0
    kl = K.mask_key_len(key_mask) if key_mask is not None else None
    out, stats = K.attn_fwd(q4, k4, v4, key_mask, kl, causal)
    dO = torch.randn(B, Sq, d, device=dev).bfloat16()
    # reference gradients through fp32 autograd
    qf, kf, vf = (t.float().clone().requires_grad_(True) for t in (q4, k4, v4))
    ref, _ = _ref_attention(qf, kf, vf, key_mask, causal)
    ref.backward(dO.float())
    dqkv = torch.full_like(qkv, float("nan"))
    dkvb = torch.full_like(kvb, float("nan"))
    dq4, dk4, dv4 = heads(dqkv, Sq, 2 * d), heads(dkvb, Sk, 0), heads(dkvb, Sk, d)
    K.attn_bwd(dO, out, stats, q4, k4, v4, dq4, dk4, dv4, key_mask, kl, causal)
    for name, got, want in (("dq", dq4, qf.grad), ("dk", dk4, kf.grad), ("dv", dv4, vf.grad)):
        got = got.float()
        assert torch.isfinite(got).all(), name
        scale = want.abs().max().item() + 1e-6
        err = (got - want).abs().max().item()
        assert err <= 3e-2 * scale + 2e-3, (name, err, scale)
        cos = torch.nn.functional.cosine_similarity(got.flatten(), want.flatten(), dim=0).item()
        assert cos >= 0.999, (name, cos)


def test_ner_map_block_applies_the_reference_dropout(cuda_device):
    """MFULL:685-687: dropout sits between ner_map_down and the reshape + ner_map_layer_norm.  With res = None the
    LayerNorm kernel is y = LN(dropout(z)): the mask the backward pass rebuilds is the one the forward pass applied,
    the drop rate is p, and NerMapFn uses it in training mode only."""
    from vacnic_b200 import blocks as Bk, kernels as k, spec
    from vacnic_b200.modeling import VacnicBart
    d, rows, p = 1024, 2048, 0.25
    gamma = torch.ones(d, device=cuda_device); beta = torch.zeros(d, device=cuda_device)
    rng = k.Rng(cuda_device, seed=5)
    z = torch.ones(rows, d, dtype=torch.bfloat16, device=cuda_device)        # dropout(1) in {0, 1/(1-p)}: LN output sign = mask
    y, mean, rstd = k.add_layernorm_fwd(z, None, gamma, beta, p_drop=p, rng=rng, salt=3)
    kept_fwd = y.float() > 0
    assert abs((~kept_fwd).float().mean().item() - p) < 0.01
    # the LayerNorm kernels draw TWO decisions from one counter hash (16-bit fields): the two elements of a pair, neighbouring
    # pairs, and the same element under another salt / another step must be independent Bernoulli(1 - p) draws
    kf = kept_fwd.float()
    q = 1 - p
    assert abs((kf[:, 0::2] * kf[:, 1::2]).mean().item() - q * q) < 0.01
    assert abs((kf[:, 1:-1:2] * kf[:, 2::2]).mean().item() - q * q) < 0.01
    y_salt, _, _ = k.add_layernorm_fwd(z, None, gamma, beta, p_drop=p, rng=rng, salt=4)
    assert abs((kf * (y_salt.float() > 0).float()).mean().item() - q * q) < 0.01
    rng2 = k.Rng(cuda_device, seed=6)
    y_seed, _, _ = k.add_layernorm_fwd(z, None, gamma, beta, p_drop=p, rng=rng2, salt=3)
    assert abs((kf * (y_seed.float() > 0).float()).mean().item() - q * q) < 0.01
    assert abs(kf.mean(0).std().item() - (p * q / rows) ** 0.5) < 0.3 * (p * q / rows) ** 0.5   # per-column keep rates: binomial spread
    dg = torch.zeros(d, device=cuda_device); db = torch.zeros(d, device=cuda_device)
    dy = rnd((rows, d), cuda_device, 1.0, 9)
    _, dx = k.add_layernorm_bwd(dy, z, None, gamma, mean, rstd, dg, db, want_dx=True, p_drop=p, rng=rng, salt=3)
    assert torch.equal(dx != 0, kept_fwd) or ((dx != 0) ^ kept_fwd).float().mean().item() < 1e-4   # (a kept gradient may round to 0)
    # block level: training mode drops, eval mode does not
    cfg = spec.VacnicConfig(d_model=768, heads=12, ffn=1024, enc_layers=1, dec_layers=1, prompt_size=4, max_pos=128)
    m = VacnicBart(cfg, device=cuda_device, p_drop=0.5, seed=1)
    layer = m.model.encoder.layers[0]
    ner = rnd((2, 80, 768), cuda_device, 1.0, 4)
    outs = {}
    for mode in (False, True):
        m.rt.training = mode
        with torch.no_grad():
            outs[mode] = Bk.NerMapFn.apply(ner, m.rt, layer.lin_nup, layer.lin_ndown, layer.ln_nmap).float()
    m.rt.training = False
    with torch.no_grad():
        again = Bk.NerMapFn.apply(ner, m.rt, layer.lin_nup, layer.lin_ndown, layer.ln_nmap).float()
    assert torch.equal(outs[False], again)
    assert (outs[True] - outs[False]).abs().mean().item() > 0.1


def _keep_mask(seed, salt, rowid, key, p):
    """Python restatement of ptx.cuh keep_elem / attn_sm100.cu keep_attn (uint32 arithmetic on int64 tensors)."""
    M = 0xFFFFFFFF

    def hash32(x):
        x = x ^ (x >> 16); x = (x * 0x85EBCA6B) & M; x = x ^ (x >> 13); x = (x * 0xC2B2AE35) & M; x = x ^ (x >> 16)
        return x
    idx = (rowid << 16) | key
    h = hash32(((idx & M) * 0x9E3779B1 + seed) & M)
    h = hash32(h ^ (((idx >> 32) + salt * 0x7F4A7C15) & M))
    thr = min(int(p * 4294967296.0), M)
    return h >= thr


@pytest.mark.parametrize("B,H,Sq,Sk,causal,masked", [(2, 3, 150, 200, False, True), (2, 2, 70, 70, True, False), (1, 4, 260, 130, False, False)])
def test_attention_dropout_fwd_bwd(cuda_device, B, H, Sq, Sk, causal, masked):
    """config.attention_dropout (MFULL:546): dropout on the probabilities AFTER the softmax.  The counter-based mask is
    restated in Python, so forward and both backward kernels are checked against torch autograd with the SAME mask."""
    from vacnic_b200 import kernels as K
    dev = cuda_device
    torch.manual_seed(Sq + Sk)
    p, salt = 0.3, 77
    rng = K.Rng(dev, seed=9)
    q = (torch.randn(B, H, Sq, 64, device=dev) * 1.2).bfloat16()
    k = (torch.randn(B, H, Sk, 64, device=dev) * 1.2).bfloat16()
    v = (torch.randn(B, H, Sk, 64, device=dev) * 1.2).bfloat16()
    key_mask = None
    if masked:
        key_mask = torch.ones(B, Sk, dtype=torch.uint8, device=dev)
        key_mask[0, Sk - 37:] = 0
        key_mask[1, 5] = 0
    kl = K.mask_key_len(key_mask) if key_mask is not None else None
    out, stats = K.attn_fwd(q, k, v, key_mask, kl, causal, p_drop=p, rng=rng, salt=salt)
    out0, _ = K.attn_fwd(q, k, v, key_mask, kl, causal)
    assert (out.float() - out0.float()).abs().mean().item() > 1e-2          # dropout really changes the result
    seed = int(rng.state.item()) & 0xFFFFFFFF
    rowid = ((torch.arange(B, device=dev)[:, None, None] * H + torch.arange(H, device=dev)[None, :, None]) * Sq
             + torch.arange(Sq, device=dev)[None, None, :])[..., None].to(torch.int64)
    keep = _keep_mask(seed, salt, rowid, torch.arange(Sk, device=dev, dtype=torch.int64)[None, None, None, :], p)
    assert abs(keep.float().mean().item() - (1 - p)) < 0.01
    qf, kf, vf = (t.float().clone().requires_grad_(True) for t in (q, k, v))
    s = qf @ kf.transpose(-1, -2) * 0.125
    neg = torch.finfo(torch.float32).min
    if key_mask is not None:
        s = s.masked_fill(key_mask[:, None, None, :] == 0, neg)
    if causal:
        s = s.masked_fill(torch.ones(Sq, Sk, dtype=torch.bool, device=dev).triu(1), neg)
    pr = torch.softmax(s, -1) * keep / (1 - p)
    ref = (pr @ vf).transpose(1, 2).reshape(B, Sq, H * 64)
    err = (out.float() - ref).abs()
    assert err.max().item() <= 4e-2 and err.mean().item() <= 3e-3, (err.max().item(), err.mean().item())
    dO = torch.randn(B, Sq, H * 64, device=dev).bfloat16()
    ref.backward(dO.float())
    dq, dk, dv = (torch.full_like(t, float("nan")) for t in (q, k, v))
    K.attn_bwd(dO, out, stats, q, k, v, dq, dk, dv, key_mask, kl, causal, p_drop=p, rng=rng, salt=salt)
    for name, got, want in (("dq", dq, qf.grad), ("dk", dk, kf.grad), ("dv", dv, vf.grad)):
        got = got.float()
        assert torch.isfinite(got).all(), name
        scale = want.abs().max().item() + 1e-6
        e = (got - want).abs().max().item()
        cos = torch.nn.functional.cosine_similarity(got.flatten(), want.flatten(), dim=0).item()
        assert e <= 4e-2 * scale + 3e-3 and cos >= 0.998, (name, e, scale, cos)


def test_activation_dropout_inplace_and_ffn_block(cuda_device):
    """config.activation_dropout (MFULL:649, 660, 684, 740, 874): the in-place kernel keeps 1 - p of the elements scaled by
    1 / (1 - p), the same key reproduces the same mask (backward), and the FFN block uses it in training mode only."""
    from vacnic_b200 import kernels as K, spec
    from vacnic_b200.modeling import VacnicBart
    dev = cuda_device
    rng = K.Rng(dev, seed=4)
    x = torch.ones(1000, 1031, dtype=torch.bfloat16, device=dev)[:, :1024].contiguous()
    a = K.dropout_inplace(x.clone(), 0.25, rng, 5)
    b = K.dropout_inplace(x.clone(), 0.25, rng, 5)
    c = K.dropout_inplace(x.clone(), 0.25, rng, 6)
    assert torch.equal(a, b) and not torch.equal(a, c)
    assert abs((a == 0).float().mean().item() - 0.25) < 0.01 and set(a.float().unique().tolist()) <= {0.0, float(torch.tensor(1 / 0.75).bfloat16())}
    cfg = spec.VacnicConfig(d_model=768, heads=12, ffn=1024, enc_layers=1, dec_layers=1, prompt_size=4, max_pos=128)
    outs = {}
    for p_act in (0.0, 0.5):
        m = VacnicBart(cfg, device=dev, p_drop=0.0, seed=1, p_act=p_act, p_attn=0.0)
        from vacnic_b200 import synthetic
        batch = synthetic.to_device(synthetic.make_batch(B=2, L=40, T=8, seed=1), dev)
        face = batch["face_emb"]
        kw = dict(input_ids=batch["article_ids"], attention_mask=(batch["article_ids"] != 1).long(), decoder_input_ids=batch["caption_ids"],
                  image_features=batch["image_features"], face_features=face, face_mask=(face[:, :, -1] != 1).long(),
                  name_ids=batch["names_art_ids"], name_mask=(batch["names_art_ids"] != 1).long())
        m.eval()
        with torch.no_grad():
            outs[(p_act, "eval")] = m(**kw)["logits"].float().clone()
        m.train()
        out = m(ce_targets=batch["caption_ids"], **kw)
        out["loss"].backward()
        torch.cuda.synchronize()
        assert torch.isfinite(m.store.grad).all()
        outs[(p_act, "train")] = out["logits"].float().clone()
    assert torch.equal(outs[(0.0, "eval")], outs[(0.5, "eval")])                     # inference never drops
    assert (outs[(0.5, "train")] - outs[(0.0, "train")]).abs().mean().item() > 1e-3  # training with p_act > 0 does
