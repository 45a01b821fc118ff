"""Generate tests/golden/*.pt from the UNMODIFIED reference classes (run in the build container,
where /root/reference exists):   python tests/golden/make_golden.py

For each case it (1) builds the reference module (oracle/ref_shim.py shims), (2) loads the
deterministic test weights of vacnic_b200.spec.test_state_dict — which also proves that our parameter
inventory equals the reference state_dict key for key, (3) runs the reference forward / loss block /
generate() on a synthetic batch, (4) checks the restatement in oracle/ against it, and (5) stores a
compact fixture: seeds + shapes (weights and inputs are regenerated from the seed; a checksum guards
reproduction), logits at 256 fixed vocabulary columns, per-position logsumexp and argmax, losses,
hidden states, greedy and beam-4 token ids.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import generate as OG  # noqa: E402
from oracle import model as OM  # noqa: E402
from oracle import ref_shim  # noqa: E402
from vacnic_b200 import spec, synthetic  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

CASES = {
    # name: (cfg kwargs, batch kwargs, weight seed, lm_scale)
    "base_full_mini": (dict(d_model=768, heads=12, ffn=1024, enc_layers=2, dec_layers=2, prompt_size=4, max_pos=128),
                       dict(B=3, L=48, T=12, seed=7), 11, 8.0),
    "large_full_mini": (dict(d_model=1024, heads=16, ffn=1024, enc_layers=2, dec_layers=2, prompt_size=6, max_pos=128),
                        dict(B=2, L=40, T=10, seed=8), 12, 8.0),
    "large_vis_mini": (dict(d_model=1024, heads=16, ffn=1024, enc_layers=2, dec_layers=2, prompt_size=6, max_pos=128,
                            only_image=True), dict(B=2, L=40, T=10, seed=9), 13, 8.0),
    # same model with final_logits_bias[eos] raised so that beams finish at different lengths
    "base_full_mini_eos": (dict(d_model=768, heads=12, ffn=1024, enc_layers=2, dec_layers=2, prompt_size=4, max_pos=128),
                           dict(B=4, L=48, T=12, seed=17), 15, 8.0, 13.0),
    "base_full_mini_eos16": (dict(d_model=768, heads=12, ffn=1024, enc_layers=2, dec_layers=2, prompt_size=4, max_pos=128),
                             dict(B=4, L=48, T=12, seed=17), 15, 8.0, 16.0),
    # BASELINE.json configs[0]: BART-base full model, batch 2, 512 article tokens, 40-token caption, P=10
    "config1_base": (dict(d_model=768, heads=12, ffn=3072, enc_layers=6, dec_layers=6, prompt_size=10, max_pos=1024),
                     dict(B=2, L=512, T=40, seed=42), 14, 1.0),
}


def hf_config(cfg: spec.VacnicConfig):
    from transformers import BartConfig
    return BartConfig(vocab_size=cfg.vocab, d_model=cfg.d_model, encoder_layers=cfg.enc_layers,
                      decoder_layers=cfg.dec_layers, encoder_attention_heads=cfg.heads,
                      decoder_attention_heads=cfg.heads, encoder_ffn_dim=cfg.ffn, decoder_ffn_dim=cfg.ffn,
                      max_position_embeddings=cfg.max_pos, output_hidden_states=True, dropout=0.0)


def build_reference(cfg: spec.VacnicConfig, sd):
    Full, Vis = ref_shim.oracle_classes()
    cls = Vis if cfg.only_image else Full
    kw = dict(enc_fusion_layer=list(range(cfg.enc_layers)), dim_common=cfg.d_model, img_size=768,
              prompt_mlp_type="clipcap", map_size=[196, 256, 64, 16], prompt_size=cfg.prompt_size, clip_model=None,
              freeze_clip=False, max_ner_type_len=cfg.max_ner_type_len, max_ner_type_len_gt=cfg.max_ner_type_len_gt,
              only_image=cfg.only_image, init_attn_weight=False)
    if cfg.only_image:
        for k in ("max_ner_type_len", "max_ner_type_len_gt", "only_image", "init_attn_weight"):
            kw.pop(k, None)
    m = cls(hf_config(cfg), **kw).eval()
    ref_keys = list(m.state_dict().keys())
    ours = spec.param_shapes(cfg)
    assert set(ref_keys) == set(ours.keys()), (sorted(set(ref_keys) ^ set(ours.keys())))
    for k, v in m.state_dict().items():
        assert tuple(v.shape) == tuple(ours[k]), (k, v.shape, ours[k])
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not missing and not unexpected, (missing, unexpected)
    return m


def build_guide(cfg: spec.VacnicConfig, seed):
    from transformers import BartForConditionalGeneration
    gcfg = spec.VacnicConfig(**{**cfg.as_dict(), "stock": True, "only_image": False})
    gsd = spec.test_state_dict(gcfg, seed)
    g = BartForConditionalGeneration(hf_config(cfg)).eval()
    keys = set(g.state_dict().keys())
    assert keys == set(spec.param_shapes(gcfg).keys()), sorted(keys ^ set(spec.param_shapes(gcfg).keys()))
    g.load_state_dict(gsd, strict=True)
    return g, gcfg, gsd


def ref_losses(m, guide, batch, cfg, margin=1.0, alpha=0.5, w=1.0):
    """The reference loss block (TRAIN:267-363) executed with the reference modules and torch's own
    loss classes exactly as the script wires them (TRAIN:816-820, 631-660)."""
    src, tgt = batch["article_ids"], batch["caption_ids"]
    tgt_input = OM.shift_tokens_right(tgt, 1, 2)
    sm = OM.src_mask(src)
    if cfg.only_image:
        out = m(input_ids=src, attention_mask=sm, decoder_input_ids=tgt_input, image_features=batch["image_features"])
    else:
        face = batch["face_emb"]
        out = m(input_ids=src, attention_mask=sm, decoder_input_ids=tgt_input, image_features=batch["image_features"],
                face_features=face, face_mask=OM.src_mask(face[:, :, -1]), name_ids=batch["names_art_ids"],
                name_mask=OM.src_mask(batch["names_art_ids"]), add_ner_ffn=True)
    logits = out["logits"]
    txt = torch.nn.CrossEntropyLoss(ignore_index=1)(logits.reshape(-1, logits.shape[-1]), tgt.reshape(-1))
    res = dict(out=out, txt=txt)
    gh = guide(input_ids=src, attention_mask=sm, decoder_input_ids=tgt_input)["decoder_hidden_states"][-1]
    tm = OM.src_mask(tgt)
    a, b = OM.pool(out["decoder_hidden_states"][-1], tm), OM.pool(gh, tm)
    a, b = a / a.norm(dim=1, keepdim=True), b / b.norm(dim=1, keepdim=True)
    scores = torch.matmul(a, b.t())
    res["margin"] = torch.nn.HingeEmbeddingLoss(margin)(scores.diag(), -torch.ones(a.shape[0]))
    loss = txt + alpha * res["margin"]
    if not cfg.only_image:
        enc = m.model.encoder
        hs = []
        with torch.no_grad():
            for i in range(batch["names_ids"].shape[1]):
                ids = batch["names_ids"][:, i, :]
                h = enc.embed_tokens_ner(ids) * enc.embed_scale + enc.embed_positions_ner(ids.size())
                hs.append(torch.mean(enc.layernorm_embedding_ner(h), dim=1))
        names = torch.stack(hs, dim=1)
        face_j = out["hidden_states_face"]
        m1 = torch.matmul(names.unsqueeze(1), face_j.permute(0, 2, 1))
        m2 = torch.matmul(face_j.unsqueeze(1), names.permute(0, 2, 1))

        def bs(x):
            n = x.shape[2]
            lg = x.max(-1).values.sum(-1).div(torch.tensor(n).expand(x.shape[0]).unsqueeze(1).expand(x.shape[0], x.shape[0]))
            return torch.nn.functional.cross_entropy(lg, torch.arange(x.shape[0]))
        res["secla"] = bs(m1) + bs(m2)
        loss = loss + w * res["secla"]
    res["loss"] = loss
    return res


def enc_inputs_of(cfg, batch):
    src = batch["article_ids"]
    d = dict(input_ids=src, attention_mask=OM.src_mask(src), image_features=batch["image_features"])
    if not cfg.only_image:
        face = batch["face_emb"]
        d.update(face_features=face, face_mask=OM.src_mask(face[:, :, -1]), name_ids=batch["names_art_ids"],
                 name_mask=OM.src_mask(batch["names_art_ids"]))
    return d


def vet_generation(sd, cfg, batch, max_len, amp, trials=10):
    """True when greedy and beam-4 ids of the (reference-pinned) oracle do not change under uniform logit noise of
    amplitude `amp` (= the bf16 logit tolerance of the GPU tests) in any of `trials` draws."""
    inp = enc_inputs_of(cfg, batch)
    with torch.no_grad():
        enc = OM.encoder_forward(sd, cfg.as_dict(), **inp)
    OG.LOGIT_NOISE = None
    g0 = OG.greedy(sd, cfg.as_dict(), inp, max_length=max_len, enc=enc)
    b0, _ = OG.beam_search(sd, cfg.as_dict(), inp, num_beams=4, max_length=max_len, length_penalty=2.0, enc=enc)
    try:
        for t in range(trials):
            OG.LOGIT_NOISE = (amp, torch.Generator().manual_seed(1000 + t))
            g = OG.greedy(sd, cfg.as_dict(), inp, max_length=max_len, enc=enc)
            if g.shape != g0.shape or not bool((g == g0).all()):
                return False
            b, _ = OG.beam_search(sd, cfg.as_dict(), inp, num_beams=4, max_length=max_len, length_penalty=2.0, enc=enc)
            if b.shape != b0.shape or not bool((b == b0).all()):
                return False
    finally:
        OG.LOGIT_NOISE = None
    return True


def close(a, b, tol, what):
    err = (a - b).abs().max().item()
    assert err <= tol, f"{what}: max abs err {err} > {tol}"
    return err


def main():
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    only = sys.argv[1:]
    for name, case in CASES.items():
        if only and name not in only:
            continue
        ckw, bkw, wseed, lm_scale = case[:4]
        eos_bias = case[4] if len(case) > 4 else 0.0
        cfg = spec.VacnicConfig(**ckw)
        sd = spec.test_state_dict(cfg, wseed, lm_scale=lm_scale)
        sd["final_logits_bias"][0, cfg.eos_token_id] = eos_bias
        # margin vetting: walk the batch seed until the decoded ids are robust to the bf16 logit tolerance
        max_len = 16 if name != "config1_base" else 24
        # bf16 logit errors measured on B200: max-abs ~1e-2 * lm_scale over all 50k logits, sigma ~2e-3 * lm_scale;
        # vet with uniform noise of amplitude 6.25e-3 * lm_scale (sigma 3.6e-3 * lm_scale) in 10 independent draws
        amp = 6.25e-3 * max(1.0, lm_scale)
        bkw = dict(bkw)
        for attempt in range(3000):
            batch = synthetic.make_batch(**bkw)
            if vet_generation(sd, cfg, batch, max_len, amp):
                break
            bkw["seed"] += 1000
        else:
            raise RuntimeError(name + ": no margin-robust batch seed found")
        print(name, "margin-vetted batch seed", bkw["seed"], "after", attempt + 1, "attempt(s), amp", amp, flush=True)
        m = build_reference(cfg, sd)
        guide, gcfg, gsd = build_guide(cfg, wseed + 100)
        # ---- reference forward + loss block (with grad, for gradient goldens)
        res = ref_losses(m, guide, batch, cfg)
        res["loss"].backward()
        out = res["out"]
        # ---- restatement
        o = OM.training_losses(sd, cfg.as_dict(), gsd, gcfg.as_dict(), batch)
        tol = 2e-4 if name == "config1_base" else 5e-5
        errs = {
            "logits": close(o["out"]["logits"], out["logits"].detach(), tol * max(1.0, lm_scale), name + " logits"),
            "dec_h": close(o["out"]["decoder_hidden_states"][-1], out["decoder_hidden_states"][-1].detach(), tol, name + " dec_h"),
            "enc_h": close(o["out"]["encoder_last_hidden_state"], out["encoder_last_hidden_state"].detach(), tol, name + " enc_h"),
            "img": close(o["out"]["hidden_states_img"], out["hidden_states_img"].detach(), tol, name + " img"),
        }
        for k in ("txt", "margin", "secla", "loss"):
            if k in res:
                errs[k] = close(o[k], res[k].detach(), 1e-5, name + " " + k)
        if not cfg.only_image:
            errs["face"] = close(o["out"]["hidden_states_face"], out["hidden_states_face"].detach(), tol, name + " face")
            errs["ner"] = close(o["out"]["hidden_states_ner"], out["hidden_states_ner"].detach(), tol, name + " ner")
        # ---- generation through the real transformers generate()
        gen = {}
        if name != "config1_base" or os.environ.get("GOLDEN_FULL_GENERATE", "1") == "1":
            src = batch["article_ids"]
            kw = dict(input_ids=src, attention_mask=OM.src_mask(src), image_features=batch["image_features"])
            enc_in = dict(input_ids=src, attention_mask=OM.src_mask(src), image_features=batch["image_features"])
            if not cfg.only_image:
                face = batch["face_emb"]
                extra = dict(face_features=face, face_mask=OM.src_mask(face[:, :, -1]), name_ids=batch["names_art_ids"],
                             name_mask=OM.src_mask(batch["names_art_ids"]))
                kw.update(extra, add_ner_ffn=True)
                enc_in.update(extra)
            with torch.no_grad():
                ids_g = m.generate(**kw, num_beams=1, max_length=max_len, do_sample=False)
                ids_b = m.generate(**kw, num_beams=4, max_length=max_len, length_penalty=2.0)
            ids_g = getattr(ids_g, "sequences", ids_g)
            ids_b = getattr(ids_b, "sequences", ids_b)
            og = OG.greedy(sd, cfg.as_dict(), enc_in, max_length=max_len)
            ob, _ = OG.beam_search(sd, cfg.as_dict(), enc_in, num_beams=4, max_length=max_len, length_penalty=2.0)
            assert og.shape == ids_g.shape and bool((og == ids_g).all()), (name, "greedy ids differ", og, ids_g)
            assert ob.shape == ids_b.shape and bool((ob == ids_b).all()), (name, "beam ids differ", ob, ids_b)
            gen = dict(greedy_ids=ids_g, beam4_ids=ids_b, max_length=max_len, vetted_logit_noise=amp)
            print(name, "greedy", ids_g.tolist(), "beam4", ids_b.tolist())
        # ---- fixture
        gcols = torch.Generator().manual_seed(1234)
        cols = torch.randperm(cfg.vocab, generator=gcols)[:256].sort().values
        lg = out["logits"].detach()
        grads = {k: p.grad.detach().clone() for k, p in m.named_parameters() if p.grad is not None}
        pick = [k for k in ("model.encoder.layers.0.self_attn.q_proj.weight", "model.encoder.layers.1.fc1.bias",
                            "model.encoder.layers.0.cross_attn_img_ner.k_proj.weight",
                            "model.encoder.layers.0.ner_map_up.weight", "model.encoder.visual_map.weight",
                            "model.encoder._linear_1.weight", "model.decoder.layers.1.encoder_attn.v_proj.weight",
                            "model.decoder.layernorm_embedding.weight", "model.encoder.layers.1.img_layer_norm.bias",
                            "model.encoder.embed_tokens_ner.weight", "model.shared.weight",
                            "model.encoder.embed_positions.weight", "lm_head.weight")
                if k in grads]
        fx = dict(
            case=name, cfg=cfg.as_dict(), batch_kwargs=bkw, weight_seed=wseed, guide_seed=wseed + 100, lm_scale=lm_scale,
            eos_bias=eos_bias,
            weight_checksum=float(sum(v.double().sum() for k, v in sd.items() if k not in spec.TIED_TO_SHARED)),
            batch_checksum=float(sum(v.double().sum() for v in batch.values())),
            logit_cols=cols, logits_at_cols=lg[..., cols].clone(), logits_lse=torch.logsumexp(lg, -1),
            logits_argmax=lg.argmax(-1), dec_h=out["decoder_hidden_states"][-1].detach().clone(),
            enc_h_sample=out["encoder_last_hidden_state"].detach()[:, :8].clone(),
            img=out["hidden_states_img"].detach().clone(),
            face=None if cfg.only_image else out["hidden_states_face"].detach().clone(),
            ner=None if cfg.only_image else out["hidden_states_ner"].detach()[:, :8].clone(),
            losses={k: float(res[k]) for k in ("txt", "margin", "secla", "loss") if k in res},
            grad_samples={k: grads[k].flatten()[:512].clone() for k in pick},
            grad_norms={k: float(grads[k].norm()) for k in pick},
            oracle_vs_reference_max_abs=errs, torch_version=torch.__version__, **gen,
        )
        torch.save(fx, os.path.join(OUT, name + ".pt"))
        print(name, "ok", {k: f"{v:.2e}" for k, v in errs.items()}, {k: v.shape for k, v in gen.items() if hasattr(v, "shape")},
              fx["losses"], flush=True)


if __name__ == "__main__":
    main()
