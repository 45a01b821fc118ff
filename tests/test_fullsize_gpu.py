"""Parity at BASELINE.json's FULL sizes and on edge shapes (ragged lengths, single tokens, no faces / no names),
against the fp32 oracle restatement run on the same device with the same weights and inputs.

Full size = BART-large VACNIC (12 + 12 layers, d = 1024), L = 1024 article tokens, T = 64, P = 20 — configs[1] at batch 2.
Tolerances: logits max-abs <= 5e-2 and mean-abs <= 8e-3 (24 bf16 layers against fp32; lm_scale 1), token CE relative
<= 5e-3, CoLaM / SECLA relative <= 2e-2, hidden states cosine >= 0.999."""
import pytest
import torch

from oracle import model as OM
from vacnic_b200 import spec, synthetic

pytestmark = pytest.mark.gpu


def _pair(cfg, seed, dev):
    from vacnic_b200.modeling import VacnicBart
    sd = spec.test_state_dict(cfg, seed)
    m = VacnicBart(cfg, device=dev, p_drop=0.0)
    m.load_reference_state_dict(sd)
    m.eval()
    return m, {k: v.to(dev) for k, v in sd.items()}


def _kw(cfg, batch):
    src = batch["article_ids"]
    kw = dict(input_ids=src, attention_mask=OM.src_mask(src), image_features=batch["image_features"])
    if not cfg.only_image:
        face = batch["face_emb"]
        kw.update(face_features=face, face_mask=OM.src_mask(face[:, :, -1]), name_ids=batch["names_art_ids"],
                  name_mask=OM.src_mask(batch["names_art_ids"]))
    return kw


def _cos(a, b):
    return torch.nn.functional.cosine_similarity(a.flatten().float(), b.flatten().float(), dim=0).item()


def test_config2_full_size_forward_and_losses(cuda_device):
    from vacnic_b200 import blocks as Bk, kernels as K
    from vacnic_b200.modeling import VacnicBart
    dev = cuda_device
    cfg = spec.bart_large()
    gcfg = spec.VacnicConfig(stock=True)
    m, sd = _pair(cfg, 51, dev)
    gsd = spec.test_state_dict(gcfg, 52)
    g = VacnicBart(gcfg, device=dev, p_drop=0.0, frozen=True)
    g.load_reference_state_dict(gsd)
    gsd = {k: v.to(dev) for k, v in gsd.items()}
    batch = synthetic.to_device(synthetic.make_batch(B=2, L=1024, T=64, seed=61), dev)
    tgt = batch["caption_ids"]
    dec_in = OM.shift_tokens_right(tgt, 1, 2)
    with torch.no_grad():
        out = m(decoder_input_ids=dec_in, ce_targets=tgt, **_kw(cfg, batch))
        gout = g(input_ids=batch["article_ids"], attention_mask=OM.src_mask(batch["article_ids"]), decoder_input_ids=dec_in)
        margin = Bk.ColamFn.apply(out["decoder_hidden_states"][-1], gout["decoder_hidden_states"][-1], tgt, 1.0, 1)
        enc = m.model.encoder
        names = K.names_embed(batch["names_ids"], m.store.w16(enc.embed_tokens_ner.weight), m.store.w16(enc.embed_positions_ner.weight),
                              enc.ln_emb_ner.g, enc.ln_emb_ner.b)
        secla = Bk.SeclaFn.apply(out["hidden_states_face"], names)
        o = OM.training_losses(sd, cfg.as_dict(), gsd, gcfg.as_dict(), batch)
    lg, olg = out["logits"].float(), o["out"]["logits"]
    err = (lg - olg).abs()
    assert err.max().item() <= 5e-2 and err.mean().item() <= 8e-3, (err.max().item(), err.mean().item())
    # only positions the reference attends to are compared for the encoder memory (pad rows are never read downstream)
    valid = OM.src_mask(batch["article_ids"]).bool()
    assert _cos(out["encoder_last_hidden_state"][valid], o["out"]["encoder_last_hidden_state"][valid]) >= 0.999
    assert _cos(out["encoder_last_hidden_state"], o["out"]["encoder_last_hidden_state"]) >= 0.999   # pad rows match too
    assert _cos(out["decoder_hidden_states"][-1], o["out"]["decoder_hidden_states"][-1]) >= 0.999
    assert _cos(out["hidden_states_face"], o["out"]["hidden_states_face"]) >= 0.999
    assert abs(out["loss"].item() - float(o["txt"])) <= 5e-3 * float(o["txt"])
    assert abs(margin.item() - float(o["margin"])) <= 2e-2 * max(1.0, abs(float(o["margin"])))
    assert abs(secla.item() - float(o["secla"])) <= 2e-2 * max(1.0, abs(float(o["secla"])))


@pytest.mark.parametrize("B,L,T,seed", [(1, 17, 1, 1), (3, 130, 5, 2), (2, 257, 33, 3), (1, 1024, 64, 4)])
def test_edge_shapes_match_oracle(cuda_device, B, L, T, seed):
    """Ragged article lengths (not multiples of the 64/128-wide tiles), a single decoder token (no causal mask), batch 1,
    captions with no faces (all-ones pad vectors, mask all zero) and no names (<NONAME>), maximum article length."""
    dev = cuda_device
    cfg = spec.VacnicConfig(d_model=1024, heads=16, ffn=2048, enc_layers=2, dec_layers=2, prompt_size=20, max_pos=1024)
    m, sd = _pair(cfg, 70 + seed, dev)
    batch = synthetic.make_batch(B=B, L=L, T=max(T, 2), seed=seed)
    batch["caption_ids"] = batch["caption_ids"][:, :T].contiguous()   # T = 1: a single decoder position
    batch["face_emb"][-1] = 1.0                                   # last sample: no faces at all (DNYT:831)
    batch["names_art_ids"][-1] = 1
    batch["names_art_ids"][-1, :3] = torch.tensor([0, 50266, 2])   # ... and no names: [<s>, <NONAME>, </s>, pad...]
    batch = synthetic.to_device(batch, dev)
    dec_in = OM.shift_tokens_right(batch["caption_ids"], 1, 2)
    with torch.no_grad():
        out = m(decoder_input_ids=dec_in, **_kw(cfg, batch))
        o = OM.model_forward(sd, cfg.as_dict(), decoder_input_ids=dec_in, **_kw(cfg, batch))
    err = (out["logits"].float() - o["logits"]).abs()
    assert err.max().item() <= 3e-2, err.max().item()
    for k in ("hidden_states_face", "hidden_states_ner", "hidden_states_img", "encoder_last_hidden_state"):
        assert _cos(out[k], o[k]) >= 0.9995, k


@pytest.mark.parametrize("which", ["config2_full_model", "config5_only_visual"])
def test_full_depth_losses_and_gradients_match_oracle(cuda_device, which):
    """Gradient parity at FULL depth (12 + 12 layers, d = 1024) and BASELINE.json sizes: configs[1]/[3] = the full model
    (MFULL) at L = 1024, T = 64 with CE + 0.5 CoLaM + SECLA, and configs[4] = the only-visual-prompt model (MVIS,
    run_onlyvis_train.sh) at L = 512 with token CE only -- batch 2, against the fp32 oracle on the same device.
    Every parameter: cosine >= 0.99 and norm within 5 % (analytically-zero gradients: only bf16 noise)."""
    from vacnic_b200 import blocks as Bk, kernels as K
    from vacnic_b200.modeling import VacnicBart
    dev = cuda_device
    vis = which == "config5_only_visual"
    cfg = spec.bart_large(only_image=vis)
    L = 512 if vis else 1024
    sd = spec.test_state_dict(cfg, 81)
    m = VacnicBart(cfg, device=dev, p_drop=0.0)
    m.load_reference_state_dict(sd)
    m.train()
    g = gsd = gcfg = None
    if not vis:
        gcfg = spec.VacnicConfig(stock=True)
        gsd = spec.test_state_dict(gcfg, 82)
        g = VacnicBart(gcfg, device=dev, p_drop=0.0, frozen=True)
        g.load_reference_state_dict(gsd)
        g.eval()
        gsd = {k: v.to(dev) for k, v in gsd.items()}
    batch = synthetic.to_device(synthetic.make_batch(B=2, L=L, T=64, seed=91), dev)
    tgt = batch["caption_ids"]
    dec_in = OM.shift_tokens_right(tgt, 1, 2)
    m.store.begin_step()
    out = m(decoder_input_ids=dec_in, ce_targets=tgt, **_kw(cfg, batch))
    heads, grads, losses = [out["loss"]], [torch.ones(1, device=dev)], {"txt": out["loss"]}
    if not vis:
        with torch.no_grad():
            gout = g(input_ids=batch["article_ids"], attention_mask=OM.src_mask(batch["article_ids"]), decoder_input_ids=dec_in)
        margin = Bk.ColamFn.apply(out["decoder_hidden_states"][-1], gout["decoder_hidden_states"][-1], tgt, 1.0, 1)
        enc = m.model.encoder
        names = K.names_embed(batch["names_ids"], m.store.w16(enc.embed_tokens_ner.weight),
                              m.store.w16(enc.embed_positions_ner.weight), enc.ln_emb_ner.g, enc.ln_emb_ner.b)
        secla = Bk.SeclaFn.apply(out["hidden_states_face"], names)
        heads += [margin, secla]
        grads += [torch.full((1,), 0.5, device=dev), torch.ones(1, device=dev)]
        losses.update(margin=margin, secla=secla)
    torch.autograd.backward(heads, grads)
    m.store.finish_backward()
    torch.cuda.synchronize()
    sdd = {k: v.to(dev).clone().requires_grad_(v.is_floating_point() and k != "final_logits_bias") for k, v in sd.items()
           if k not in spec.TIED_TO_SHARED}
    for k in spec.TIED_TO_SHARED:
        if k in sd:
            sdd[k] = sdd["model.shared.weight"]
    o = OM.training_losses(sdd, cfg.as_dict(), gsd, gcfg.as_dict() if gcfg is not None else None, batch)
    o["loss"].backward()
    for k, v in losses.items():
        want = float(o[k])
        tol = 5e-3 if k == "txt" else 2e-2
        assert abs(v.item() - want) <= tol * max(1.0, abs(want)), (k, v.item(), want)
    worst = {}
    for n, p in m.store.params.items():
        ref, got = sdd[n].grad, p.grad
        if ref is None:
            assert got.abs().max().item() == 0, n
            continue
        rn, gn = ref.norm().item(), got.norm().item()
        if rn < 1e-6:
            assert gn < 5e-3, (n, gn)
            continue
        cos = (ref.flatten() @ got.flatten()).item() / (rn * gn + 1e-30)
        worst[n] = (cos, abs(gn - rn) / rn)
    bad = {n: v for n, v in worst.items() if v[0] < 0.99 or v[1] > 5e-2}
    assert len(worst) > 300 and not bad, (len(bad), len(worst), sorted(bad.items(), key=lambda kv: kv[1][0])[:12])
