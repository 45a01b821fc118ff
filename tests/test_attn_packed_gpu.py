"""Packed (varlen) mode of the fused attention kernels (csrc/attn_sm100.cu): the rows of all sequences are stored back
to back and every sequence attends only to its own key range -- the device never sees the collate's padding
(DNYT:957-972, TRAIN:255-271).  Checked against per-sequence fp32 torch attention (forward) and its autograd (backward)
on ragged lengths: 1 token, lengths that are not multiples of the 64 / 128-row tiles, a tail of query-only rows appended
to the last sequence (the row padding of a packed batch), regular queries over packed keys (decoder cross-attention),
packed queries over regular keys (prefix cross-attention) and the causal case.  Rows / keys the geometry does not name
must not be written."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _geometry(case, dev):
    if case == "self_ragged_tail":        # encoder self-attention: 5 articles + 37 query-only tail rows on the last one
        klen = [100, 257, 1, 128, 300]
        kstart = [0, 100, 357, 358, 486]
        qlen = klen[:-1] + [klen[-1] + 37]
        return kstart, qlen, kstart, klen, sum(klen) + 37, sum(klen) + 37, False
    if case == "causal":                   # causal mask inside every sequence
        klen = [70, 129, 200]
        kstart = [0, 70, 199]
        return kstart, klen, kstart, klen, sum(klen), sum(klen), True
    if case == "regular_q_packed_k":       # decoder cross-attention: T = 24 queries per caption over its packed article
        klen = [300, 64, 513]
        kstart = [0, 300, 364]
        return [0, 24, 48], [24, 24, 24], kstart, klen, 72, sum(klen) + 11, False
    if case == "packed_q_regular_k":       # prefix cross-attention: packed article rows over a regular [B, 40] key block
        qlen = [130, 5, 260]
        return [0, 130, 135], qlen, [0, 40, 80], [40, 40, 40], sum(qlen) + 9, 120, False
    raise KeyError(case)


@pytest.mark.parametrize("case", ["self_ragged_tail", "causal", "regular_q_packed_k", "packed_q_regular_k"])
def test_packed_attention_fwd_bwd(cuda_device, case):
    from vacnic_b200 import kernels as K
    dev = cuda_device
    H, d = 4, 256
    qs, ql, ks, kl, total_q, total_k, causal = _geometry(case, dev)
    torch.manual_seed(len(case))
    qbuf = (torch.randn(total_q, 3 * d, device=dev) * 1.2).bfloat16()     # q lives at columns [2d, 3d) like the fused [k;v;q] GEMM
    kvbuf = (torch.randn(total_k, 2 * d, device=dev) * 1.2).bfloat16()

    def heads(t, rows, col0):
        ld = t.stride(0)
        return t.as_strided((1, H, rows, 64), (rows * ld, 64, ld, 1), t.storage_offset() + col0)

    q4, k4, v4 = heads(qbuf, total_q, 2 * d), heads(kvbuf, total_k, 0), heads(kvbuf, total_k, d)
    i32 = lambda x: torch.tensor(x, dtype=torch.int32, device=dev)   # noqa: E731
    geo = K.Packed(i32(qs), i32(ql), i32(ks), i32(kl), max(ql), max(kl))
    out, stats = K.attn_fwd(q4, k4, v4, causal=causal, packed=geo)
    assert out.shape == (1, total_q, d) and stats.shape == (1, H, total_q, 2)
    dO = torch.randn(1, total_q, d, device=dev).bfloat16()
    dqbuf, dkvbuf = torch.full_like(qbuf, float("nan")), torch.full_like(kvbuf, float("nan"))
    K.attn_bwd(dO, out, stats, q4, k4, v4, heads(dqbuf, total_q, 2 * d), heads(dkvbuf, total_k, 0), heads(dkvbuf, total_k, d),
               causal=causal, packed=geo)
    torch.cuda.synchronize()
    # per-sequence fp32 reference
    qf = qbuf[:, 2 * d:].float().view(total_q, H, 64).clone().requires_grad_(True)
    kf = kvbuf[:, :d].float().view(total_k, H, 64).clone().requires_grad_(True)
    vf = kvbuf[:, d:].float().view(total_k, H, 64).clone().requires_grad_(True)
    ref = torch.zeros(total_q, H, 64, device=dev)
    covered_q = torch.zeros(total_q, dtype=torch.bool, device=dev)
    covered_k = torch.zeros(total_k, dtype=torch.bool, device=dev)
    for b in range(len(qs)):
        qb, kb, vb = qf[qs[b]:qs[b] + ql[b]], kf[ks[b]:ks[b] + kl[b]], vf[ks[b]:ks[b] + kl[b]]
        s = torch.einsum("qhc,khc->hqk", qb, kb) * 0.125
        if causal:
            s = s.masked_fill(torch.ones(ql[b], kl[b], dtype=torch.bool, device=dev).triu(1), torch.finfo(torch.float32).min)
        ref[qs[b]:qs[b] + ql[b]] = torch.einsum("hqk,khc->qhc", torch.softmax(s, -1), vb)
        covered_q[qs[b]:qs[b] + ql[b]] = True
        covered_k[ks[b]:ks[b] + kl[b]] = True
    (ref * dO.float().view(total_q, H, 64)).sum().backward()
    got = out.float().view(total_q, H, 64)
    err = (got[covered_q] - ref.detach()[covered_q]).abs()
    assert err.max().item() <= 3e-2 and err.mean().item() <= 2e-3, (case, err.max().item(), err.mean().item())
    dq = dqbuf[:, 2 * d:].float().view(total_q, H, 64)
    dk = dkvbuf[:, :d].float().view(total_k, H, 64)
    dv = dkvbuf[:, d:].float().view(total_k, H, 64)
    for name, g, want, cov in (("dq", dq, qf.grad, covered_q), ("dk", dk, kf.grad, covered_k), ("dv", dv, vf.grad, covered_k)):
        assert torch.isfinite(g[cov]).all(), (case, name)
        scale = want.abs().max().item() + 1e-6
        e = (g[cov] - want[cov]).abs().max().item()
        cos = torch.nn.functional.cosine_similarity(g[cov].flatten(), want[cov].flatten(), dim=0).item()
        assert e <= 3e-2 * scale + 2e-3 and cos >= 0.999, (case, name, e, scale, cos)
        if (~cov).any():   # rows outside every sequence are left alone (still the NaN fill)
            assert torch.isnan(g[~cov]).all(), (case, name, "rows outside the geometry were written")
    # query-only tail rows: keys excluded, so they receive no dK / dV and their own dQ is a valid gradient (checked above)
