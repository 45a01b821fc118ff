"""GPU parity of the B200-native model (vacnic_b200.modeling, all arithmetic in hand-written kernels
through the C ABI) against (a) the golden vectors produced by the unmodified reference classes and
(b) the fp32 oracle restatement run on the same device.

Tolerances (bf16 compute, fp32 accumulate, compared with an FP32 reference; SURVEY.md §8d measured the
reference against itself in bf16 at 1.6e-2 .. 3.3e-2 on logits): hidden states (LayerNorm outputs, |h| up
to ~4, one bf16 ulp there = 1.6e-2) with htol = 2^-8 * 4 * sqrt(residual blocks) (6e-2 for the 2+2-layer
cases, 1e-1 for the 6+6-layer config 1): 99.9 % of elements <= htol, every element <= 2 htol, mean-abs
<= 1e-2; logits max-abs <= 3e-2 * lm_scale,
losses relative <= 1e-2 (5e-3 for token CE at lm_scale 1), gradients: cosine >= 0.99 and relative
norm error <= 5e-2 against the fp32 oracle gradient.
"""
import glob
import os

import pytest
import torch

from oracle import model as OM
from vacnic_b200 import spec, synthetic

pytestmark = pytest.mark.gpu
GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.pt")))
MINI = [p for p in GOLDEN if "mini" in p and "eos" not in p]


def build(path, dev):
    from vacnic_b200.modeling import VacnicBart
    fx = torch.load(path, weights_only=False)
    cfg = spec.VacnicConfig(**fx["cfg"])
    sd = spec.test_state_dict(cfg, fx["weight_seed"], lm_scale=fx["lm_scale"])
    sd["final_logits_bias"][0, cfg.eos_token_id] = fx.get("eos_bias", 0.0)
    m = VacnicBart(cfg, device=dev, p_drop=0.0)
    m.load_reference_state_dict(sd)
    gcfg = spec.VacnicConfig(**{**fx["cfg"], "stock": True, "only_image": False})
    gsd = spec.test_state_dict(gcfg, fx["guide_seed"])
    g = VacnicBart(gcfg, device=dev, p_drop=0.0, frozen=True)
    g.load_reference_state_dict(gsd)
    batch = synthetic.to_device(synthetic.make_batch(**fx["batch_kwargs"]), dev)
    return fx, cfg, sd, gcfg, gsd, m, g, batch


def model_inputs(cfg, batch):
    src = batch["article_ids"]
    kw = dict(input_ids=src, attention_mask=OM.src_mask(src), image_features=batch["image_features"])
    if not cfg.only_image:
        face = batch["face_emb"]
        kw.update(face_features=face, face_mask=OM.src_mask(face[:, :, -1]), name_ids=batch["names_art_ids"],
                  name_mask=OM.src_mask(batch["names_art_ids"]))
    return kw


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-3] for p in GOLDEN])
def test_forward_matches_reference_golden(cuda_device, path):
    fx, cfg, sd, gcfg, gsd, m, g, batch = build(path, cuda_device)
    m.eval()
    tgt = batch["caption_ids"]
    dec_in = OM.shift_tokens_right(tgt, 1, 2)
    with torch.no_grad():
        out = m(decoder_input_ids=dec_in, **model_inputs(cfg, batch))
    lg = out["logits"].float().cpu()
    scale = max(1.0, fx["lm_scale"])
    htol = 2 ** -8 * 4 * (3 * (cfg.enc_layers + cfg.dec_layers) + 3) ** 0.5

    def hcheck(got, want, what):
        err = (got.float().cpu() - want).abs()
        q999 = torch.quantile(err.flatten()[:1_000_000], 0.999).item()
        # mean, 99.9 % quantile and the single worst element (heavy tail where LayerNorm amplifies a rounding)
        assert err.mean().item() <= 1e-2 and q999 <= htol and err.max().item() <= 2 * htol, \
            (what, err.max().item(), q999, err.mean().item(), htol)

    hcheck(out["decoder_hidden_states"][-1], fx["dec_h"], "dec_h")
    hcheck(out["encoder_last_hidden_state"][:, :8], fx["enc_h_sample"], "enc_h")
    hcheck(out["hidden_states_img"], fx["img"], "img")
    if not cfg.only_image:
        hcheck(out["hidden_states_face"], fx["face"], "face")
        hcheck(out["hidden_states_ner"][:, :8], fx["ner"], "ner")
    assert (lg[..., fx["logit_cols"]] - fx["logits_at_cols"]).abs().max().item() <= 3e-2 * scale
    assert (torch.logsumexp(lg, -1) - fx["logits_lse"]).abs().max().item() <= 3e-2 * scale
    # arg-max may only differ where the two candidates are closer than twice the logit tolerance
    top = lg.max(-1).values
    at_ref = lg.gather(-1, fx["logits_argmax"][..., None])[..., 0]
    assert (top - at_ref).max().item() <= 6e-2 * scale


@pytest.mark.parametrize("path", MINI, ids=[os.path.basename(p)[:-3] for p in MINI])
def test_losses_and_gradients_match_oracle(cuda_device, path):
    from vacnic_b200 import blocks as Bk, kernels as K
    fx, cfg, sd, gcfg, gsd, m, g, batch = build(path, cuda_device)
    m.train()  # p_drop = 0: training mode only switches the gradient plumbing on
    g.eval()
    tgt = batch["caption_ids"]
    dec_in = OM.shift_tokens_right(tgt, 1, 2)
    m.store.begin_step()
    out = m(decoder_input_ids=dec_in, ce_targets=tgt, **model_inputs(cfg, batch))
    with torch.no_grad():
        gout = g(input_ids=batch["article_ids"], attention_mask=OM.src_mask(batch["article_ids"]), decoder_input_ids=dec_in)
    margin = Bk.ColamFn.apply(out["decoder_hidden_states"][-1], gout["decoder_hidden_states"][-1], tgt, 1.0, 1)
    losses = {"txt": out["loss"], "margin": margin}
    heads, grads = [out["loss"], margin], [torch.ones(1, device=cuda_device), torch.full((1,), 0.5, device=cuda_device)]
    if not cfg.only_image:
        enc = m.model.encoder
        names = K.names_embed(batch["names_ids"], m.store.w16(enc.embed_tokens_ner.weight),
                              m.store.w16(enc.embed_positions_ner.weight), enc.ln_emb_ner.g, enc.ln_emb_ner.b)
        secla = Bk.SeclaFn.apply(out["hidden_states_face"], names)
        losses["secla"] = secla
        heads.append(secla); grads.append(torch.ones(1, device=cuda_device))
    torch.autograd.backward(heads, grads)
    m.store.finish_backward()
    torch.cuda.synchronize()
    for k, v in losses.items():
        want = fx["losses"][k]
        assert abs(v.item() - want) <= 1e-2 * max(1.0, abs(want)), (k, v.item(), want)
    # oracle gradients in fp32 on the same device
    sdd = {k: v.to(cuda_device).clone().requires_grad_(v.is_floating_point() and k != "final_logits_bias") for k, v in sd.items()
           if k not in spec.TIED_TO_SHARED}
    for k in spec.TIED_TO_SHARED:
        sdd[k] = sdd["model.shared.weight"]
    gsdd = {k: v.to(cuda_device) for k, v in gsd.items()}
    o = OM.training_losses(sdd, cfg.as_dict(), gsdd, gcfg.as_dict(), batch)
    o["loss"].backward()
    worst = {}
    for n, p in m.store.params.items():
        ref = sdd[n].grad
        got = p.grad
        if ref is None:
            assert got.abs().max().item() == 0, n
            continue
        rn, gn = ref.norm().item(), got.norm().item()
        if rn < 1e-6:  # analytically zero (e.g. softmax is invariant to the key bias): only bf16 noise allowed
            assert gn < 2e-3, (n, gn)
            continue
        cos = (ref.flatten() @ got.flatten()).item() / (rn * gn + 1e-30)
        worst[n] = (cos, abs(gn - rn) / rn)
    bad = {n: v for n, v in worst.items() if v[0] < 0.99 or v[1] > 5e-2}
    assert not bad, (len(bad), len(worst), sorted(bad.items(), key=lambda kv: kv[1][0])[:12])
    # golden gradient samples from the reference itself
    for k, gs in fx["grad_samples"].items():
        full = m.store.params[k].grad
        got = full.flatten()[:512].float().cpu()
        if gs.norm().item() == 0:
            assert got.norm().item() == 0, k
        else:
            denom = gs.norm().item() * got.norm().item() + 1e-30
            assert (gs @ got).item() / denom >= 0.99, k
        assert abs(full.norm().item() - fx["grad_norms"][k]) <= 5e-2 * fx["grad_norms"][k] + 1e-6, k


def test_accuracy_next_to_reference_bf16_modes(cuda_device):
    """The GPU-side comparator of SURVEY.md §8d: logit error against the fp32 oracle of (a) this implementation,
    (b) the reference arithmetic under torch.autocast(bfloat16) and (c) the reference arithmetic entirely in bf16
    (model.bfloat16()), on the same device, weights and config-1 inputs (BART-base, B=2, L=512, T=40).
    Measured on B200: ours 0.024 max / 0.0034 mean, autocast 0.0145 / 0.0022, full bf16 0.039 / 0.0057 (BART-large, L=1024:
    0.030 / 0.0046, 0.028 / 0.0039, 0.051 / 0.0081).  Asserted: strictly more accurate than the reference's own full-bf16
    mode, and within 2x of its autocast mode (which keeps the residual stream and LayerNorm inputs in fp32)."""
    import importlib.util
    path = os.path.join(os.path.dirname(os.path.dirname(__file__)), "tools", "accuracy_report.py")
    s = importlib.util.spec_from_file_location("accuracy_report", path)
    mod = importlib.util.module_from_spec(s)
    s.loader.exec_module(mod)
    r = mod.report(spec.bart_base(), cuda_device, B=2, L=512, T=40)
    ours, ac, b16 = r["ours"], r["ref_autocast"], r["ref_bf16"]
    assert ours["max_abs"] <= 0.85 * b16["max_abs"] and ours["mean_abs"] <= 0.85 * b16["mean_abs"], r
    assert ours["max_abs"] <= 2.0 * ac["max_abs"] and ours["mean_abs"] <= 2.0 * ac["mean_abs"], r
    assert ours["max_abs"] <= 3e-2, r
