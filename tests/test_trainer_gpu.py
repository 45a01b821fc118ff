"""TrainStep (TRAIN:253-374 as one captured step): graph replay equals eager execution, the fused AdamW follows
torch.optim.AdamW, and the optional clip_grad_norm_ (TRAIN:365-366) folded into the optimizer matches torch."""
import pytest
import torch

from oracle import model as OM
from vacnic_b200 import spec, synthetic

pytestmark = pytest.mark.gpu


def _models(dev, seed=41, max_pos=128):
    from vacnic_b200.modeling import VacnicBart
    cfg = spec.VacnicConfig(d_model=768, heads=12, ffn=1024, enc_layers=2, dec_layers=2, prompt_size=4, max_pos=max_pos)
    gcfg = spec.VacnicConfig(**{**cfg.as_dict(), "stock": True})
    sd, gsd = spec.test_state_dict(cfg, seed), spec.test_state_dict(gcfg, seed + 100)
    m = VacnicBart(cfg, device=dev, p_drop=0.0)
    m.load_reference_state_dict(sd)
    g = VacnicBart(gcfg, device=dev, p_drop=0.0, frozen=True)
    g.load_reference_state_dict(gsd)
    return cfg, gcfg, sd, gsd, m, g


def test_clip_grad_scale_kernel(cuda_device):
    from vacnic_b200 import kernels as K
    torch.manual_seed(0)
    g = torch.randn(1_000_003, device=cuda_device) * 0.01
    scratch = torch.zeros(2368, device=cuda_device)
    out = torch.zeros(2, device=cuda_device)
    for max_norm, base in ((0.1, 1.0), (100.0, 1.0), (0.1, 0.25)):
        K.clip_grad_scale(g[:1_000_000], max_norm, base, scratch, out[0:1], out[1:2])
        norm = (g[:1_000_000].double().norm() * base).item()
        want = base * min(1.0, max_norm / (norm + 1e-6))
        assert abs(out[1].item() - norm) <= 1e-4 * norm
        assert abs(out[0].item() - want) <= 1e-4 * abs(want)


@pytest.mark.parametrize("max_grad_norm", [None, 0.1])
def test_step_matches_torch_optimizer_and_graph_equals_eager(cuda_device, max_grad_norm):
    from vacnic_b200.trainer import TrainStep
    batch = synthetic.make_batch(B=2, L=40, T=12, seed=9)
    results = []
    for use_graph in (False, True):
        cfg, gcfg, sd, gsd, m, g = _models(cuda_device)
        ts = TrainStep(m, g, lr=1e-3, weight_decay=0.01, use_graph=use_graph, max_grad_norm=max_grad_norm)
        losses = [{k: float(v.detach()) for k, v in ts.step(batch).items()} for _ in range(3)]
        torch.cuda.synchronize()
        results.append((losses, m.store.master.clone()))
    (le, pe), (lg, pg) = results
    # same kernels in the same order; bias / LayerNorm / embedding gradients are accumulated with fp32 atomics, so two runs
    # differ in the last bits and Adam (sign-like for tiny gradients) may flip isolated updates by 2*lr: losses must
    # agree to 2e-3 relative, parameters on average to 1e-5
    for a, b in zip(le, lg):
        for k in a:
            assert abs(a[k] - b[k]) <= 2e-3 * max(1.0, abs(a[k])), (k, le, lg)
    assert (pe - pg).abs().mean().item() <= 1e-5
    # fp32 oracle + torch.optim.AdamW (+ clip_grad_norm_) taking the same three steps
    cfg, gcfg, sd, gsd, m, g = _models(cuda_device)
    osd = {k: v.to(cuda_device).clone().requires_grad_(v.is_floating_point() and k != "final_logits_bias") for k, v in sd.items()
           if k not in spec.TIED_TO_SHARED}
    for k in spec.TIED_TO_SHARED:
        osd[k] = osd["model.shared.weight"]
    ogsd = {k: v.to(cuda_device) for k, v in gsd.items()}
    uniq = list({id(v): v for v in osd.values() if v.requires_grad}.values())
    opt = torch.optim.AdamW(uniq, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01)
    db = synthetic.to_device(batch, cuda_device)
    ol = []
    for _ in range(3):
        o = OM.training_losses(osd, cfg.as_dict(), ogsd, gcfg.as_dict(), db)
        o["loss"].backward()
        if max_grad_norm is not None:
            torch.nn.utils.clip_grad_norm_(uniq, max_norm=max_grad_norm)
        opt.step()
        opt.zero_grad()
        ol.append({k: float(o[k]) for k in ("txt", "margin", "secla")})
    for a, b in zip(le, ol):
        for k in b:
            assert abs(a[k] - b[k]) <= 3e-2 * max(1.0, abs(b[k])), (k, le, ol)
    assert le[2]["txt"] < le[0]["txt"]  # the optimizer is actually applied


def test_steps_enqueued_without_sync_equal_steps_synced_every_time(cuda_device):
    """The host may run any number of graph replays ahead: step counter, warm-up learning rate and Adam bias corrections
    live on the device (vacnic_optim_schedule), so N unsynchronised steps give bit-for-bit the schedule of N synchronised
    ones (round-1 advisor finding: a re-used pinned host buffer was overwritten before the async copy ran)."""
    from vacnic_b200.trainer import TrainStep, linear_schedule
    batch = synthetic.make_batch(B=2, L=40, T=12, seed=9)
    outs = []
    for sync in (True, False):
        cfg, gcfg, sd, gsd, m, g = _models(cuda_device)
        ts = TrainStep(m, g, lr=1e-3, weight_decay=0.01, warmup_steps=6, total_steps=20, use_graph=True)
        devb = {k: v.to(cuda_device) for k, v in TrainStep.prepare(batch, cfg).items()}
        hyper_seen = []
        for i in range(8):
            ts.step(devb, prepared=True)
            if sync:
                torch.cuda.synchronize()
                hyper_seen.append(ts.hyper.cpu().clone())
        torch.cuda.synchronize()
        outs.append((ts.hyper.cpu().clone(), int(ts.step_dev.item()), m.store.master.clone(), hyper_seen))
    (h_sync, t_sync, p_sync, seen), (h_async, t_async, p_async, _) = outs
    assert t_sync == t_async == 8
    assert torch.equal(h_sync, h_async)
    for t, h in enumerate(seen, start=1):   # the device schedule is the reference's (TRAIN:102, 371-374)
        assert abs(h[0].item() - 1e-3 * linear_schedule(t - 1, 6, 20)) <= 1e-9
        assert abs(h[5].item() - (1 - 0.9 ** t)) <= 1e-6 and abs(h[6].item() - (1 - 0.999 ** t)) <= 1e-7
    # same kernels, same order, same hyper-parameters: only the fp32 atomics of bias / LayerNorm gradients may differ
    assert (p_sync - p_async).abs().mean().item() <= 1e-5


def test_plain_loop_still_works_after_a_trainstep_used_the_model(cuda_device):
    """`store.external_step` is reset when TrainStep.step returns: a later forward / backward / optimizer.step loop on the
    same model starts a fresh gradient step (bias / LayerNorm / embedding gradients must not accumulate across steps)."""
    from vacnic_b200.trainer import TrainStep
    cfg, gcfg, sd, gsd, m, g = _models(cuda_device)
    batch = synthetic.make_batch(B=2, L=40, T=12, seed=9)
    ts = TrainStep(m, g, lr=0.0, weight_decay=0.0, use_graph=False)
    ts.step(batch)
    assert m.store.external_step is False
    db = {k: v.to(cuda_device) for k, v in TrainStep.prepare(batch, cfg).items()}
    kw = dict(input_ids=db["article_ids"], attention_mask=db["src_mask"], decoder_input_ids=db["decoder_input_ids"],
              image_features=db["image_features"], face_features=db["face_emb"], face_mask=db["face_mask"],
              name_ids=db["names_art_ids"], name_mask=db["name_mask"], ce_targets=db["caption_ids"])
    grads = []
    for _ in range(2):
        m(**kw)["loss"].backward()
        torch.cuda.synchronize()
        grads.append(m.store.grad.clone())
    z = m.store.z_begin
    denom = grads[0][z:].abs().max().item()
    assert denom > 0 and (grads[0][z:] - grads[1][z:]).abs().max().item() <= 1e-3 * denom   # not doubled


@pytest.mark.parametrize("use_graph", [False, True], ids=["eager", "graph"])
@pytest.mark.parametrize("varlen", [False, True], ids=["padded", "packed"])
def test_side_streams_give_the_single_stream_step(cuda_device, use_graph, varlen):
    """The prefix side of the encoder on a second stream (BartEncoder.forward, TrainStep(side_stream=True)) and the frozen
    guide's forward on a third (TrainStep(guide_stream=True)) change the schedule of the kernels, not their arithmetic:
    losses and gradients of every step, and the weights after three steps, equal the single-stream run up to the fp32
    atomics of the bias / LayerNorm gradients."""
    from vacnic_b200.trainer import TrainStep
    batch = synthetic.make_batch(B=3, L=72, T=12, seed=19)
    runs = []
    for side, gstream in ((False, False), (False, False), (True, False), (False, True), (True, True)):
        cfg, gcfg, sd, gsd, m, g = _models(cuda_device, max_pos=1024)
        # small learning rate: the comparison is about the schedule of the kernels, not about how fast three Adam steps at
        # lr 1e-3 amplify the last-bit noise of the fp32 atomics
        ts = TrainStep(m, g, lr=2e-5, weight_decay=0.01, use_graph=use_graph, varlen=varlen, side_stream=side,
                       guide_stream=gstream)
        assert (ts.side_stream is not None) == side and (ts.guide_stream is not None) == gstream
        losses, grads = [], []
        for _ in range(3):
            losses.append({k: float(v.detach()) for k, v in ts.step(batch).items()})
            torch.cuda.synchronize()
            grads.append(m.store.grad.clone())
        assert not m.rt.keepalive and m.rt.side_stream is None
        runs.append((losses, grads, m.store.master.clone()))
    l0, g0, p0 = runs[0]
    # runs[1] repeats the single-stream run: its distance from runs[0] is the run-to-run noise of the fp32 atomics
    noise = max((g0[i] - runs[1][1][i]).abs().max().item() for i in range(3))
    scale = max(g.abs().max().item() for g in g0)
    for l1, g1, p1 in runs[2:]:
        for a, b in zip(l0, l1):
            for k in a:
                assert abs(a[k] - b[k]) <= 1e-3 * max(1.0, abs(a[k])), (k, l0, l1)
        for i in range(3):   # every step: the same gradient, element by element, within a few times the run-to-run noise
            err = (g0[i] - g1[i]).abs().max().item()
            assert err <= max(4 * noise, 1e-4 * scale), (i, err, noise, scale)
            assert torch.nn.functional.cosine_similarity(g0[i], g1[i], dim=0).item() >= 0.99999
        assert (p0 - p1).abs().mean().item() <= 3e-6   # three steps of lr 2e-5 move a weight by up to 6e-5
