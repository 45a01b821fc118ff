"""Parity of the B200 path with "the reference's own PyTorch implementation" in the two modes SURVEY.md 8(d) names, on the
same device, weights and inputs: fp32, and torch.autocast(bfloat16) (the mode north_star's tolerance is quoted against).
Reports, for logits (non-pad caption positions) and for the three losses (token CE, CoLaM margin, SECLA):
  ours vs autocast, ours vs fp32, autocast vs fp32 (how far the reference is from itself across its precisions).
`report()` is what tests/test_parity_autocast_gpu.py asserts on.
    python tools/parity_report.py [--large] [--B 2]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from oracle import model as OM  # noqa: E402
from vacnic_b200 import spec, synthetic  # noqa: E402


def ours_losses(m, g, cfg, batch):
    from vacnic_b200 import blocks as Bk, kernels as K
    src, tgt = batch["article_ids"], batch["caption_ids"]
    dec_in = OM.shift_tokens_right(tgt, 1, 2)
    face = batch["face_emb"]
    kw = dict(input_ids=src, attention_mask=OM.src_mask(src), image_features=batch["image_features"])
    if not cfg.only_image:
        kw.update(face_features=face, face_mask=OM.src_mask(face[:, :, -1]), name_ids=batch["names_art_ids"],
                  name_mask=OM.src_mask(batch["names_art_ids"]))
    with torch.no_grad():
        out = m(decoder_input_ids=dec_in, ce_targets=tgt, **kw)
        res = {"logits": out["logits"].float(), "txt": float(out["loss"])}
        if g is not None:
            gout = g(input_ids=src, attention_mask=OM.src_mask(src), decoder_input_ids=dec_in)
            res["margin"] = float(Bk.ColamFn.apply(out["decoder_hidden_states"][-1], gout["decoder_hidden_states"][-1], tgt, 1.0, 1))
        if not cfg.only_image:
            enc = m.model.encoder
            names = K.names_embed(batch["names_ids"], m.store.w16(enc.embed_tokens_ner.weight),
                                  m.store.w16(enc.embed_positions_ner.weight), enc.ln_emb_ner.g, enc.ln_emb_ner.b)
            res["secla"] = float(Bk.SeclaFn.apply(out["hidden_states_face"], names))
    return res


def ref_losses(sd, cfg, gsd, gcfg, batch, autocast):
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
        o = OM.training_losses(sd, cfg.as_dict(), gsd, gcfg.as_dict() if gcfg is not None else None, batch)
    res = {"logits": o["out"]["logits"].float(), "txt": float(o["txt"])}
    for k in ("margin", "secla"):
        if k in o and o[k] is not None:
            res[k] = float(o[k])
    return res


def report(cfg, dev, B, L, T, weight_seed=7, batch_seed=3, lm_scale=1.0, with_guide=True):
    from vacnic_b200.modeling import VacnicBart
    sd_cpu = spec.test_state_dict(cfg, weight_seed, lm_scale=lm_scale)
    sd = {k: v.to(dev) for k, v in sd_cpu.items()}
    m = VacnicBart(cfg, device=dev, p_drop=0.0)
    m.load_reference_state_dict(sd_cpu)
    m.eval()
    g = gsd = gcfg = None
    if with_guide and not cfg.only_image:
        gcfg = spec.VacnicConfig(**{**cfg.as_dict(), "stock": True})
        gsd_cpu = spec.test_state_dict(gcfg, weight_seed + 100)
        gsd = {k: v.to(dev) for k, v in gsd_cpu.items()}
        g = VacnicBart(gcfg, device=dev, p_drop=0.0, frozen=True)
        g.load_reference_state_dict(gsd_cpu)
        g.eval()
    batch = synthetic.to_device(synthetic.make_batch(B=B, L=L, T=T, seed=batch_seed), dev)
    valid = batch["caption_ids"] != 1
    r = {"ours": ours_losses(m, g, cfg, batch), "fp32": ref_losses(sd, cfg, gsd, gcfg, batch, False),
         "autocast": ref_losses(sd, cfg, gsd, gcfg, batch, True)}
    out = {}
    for a, b in (("ours", "autocast"), ("ours", "fp32"), ("autocast", "fp32")):
        err = (r[a]["logits"] - r[b]["logits"]).abs()[valid]
        d = {"logits_max_abs": err.max().item(), "logits_mean_abs": err.mean().item()}
        for k in ("txt", "margin", "secla"):
            if k in r[a] and k in r[b]:
                d[k + "_rel"] = abs(r[a][k] - r[b][k]) / max(1e-6, abs(r[b][k]))
        out[f"{a}_vs_{b}"] = d
    out["losses"] = {k: {kk: vv for kk, vv in v.items() if kk != "logits"} for k, v in r.items()}
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--large", action="store_true")
    ap.add_argument("--vis", action="store_true", help="only-visual model (configs[4]), L=512")
    ap.add_argument("--B", type=int, default=2)
    ap.add_argument("--no-res-fp32", action="store_true")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    if a.vis:
        cfg, L, T = spec.bart_large(only_image=True), 512, 64
    else:
        cfg, L, T = (spec.bart_large(), 1024, 64) if a.large else (spec.bart_base(), 512, 40)
    if a.no_res_fp32:
        from vacnic_b200 import blocks
        _init = blocks.Runtime.__init__

        def _patched(self, *args, **kw):
            _init(self, *args, **kw)
            self.res_fp32 = False
        blocks.Runtime.__init__ = _patched
    r = report(cfg, dev, a.B, L, T)
    name = "BART-large only-visual" if a.vis else ("BART-large" if a.large else "BART-base")
    print(f"{name}, B={a.B}, L={L}, T={T}, res_fp32={not a.no_res_fp32}")
    for k, v in r.items():
        print(f"  {k:20s} " + "  ".join(f"{kk}={vv:.3e}" if isinstance(vv, float) else f"{kk}={vv}" for kk, vv in v.items()))
