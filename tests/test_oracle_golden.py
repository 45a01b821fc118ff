"""CPU: the oracle restatement (oracle/) against the golden vectors generated from the unmodified
reference classes (tests/golden/make_golden.py).  Tolerances: fp32 re-execution on a possibly
different CPU -> 2e-4 absolute on activations, 1e-5 relative on losses; token ids exact."""
import glob
import os

import pytest
import torch

from oracle import generate as OG
from oracle import model as OM
from vacnic_b200 import spec, synthetic

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.pt")))


def load_case(path):
    fx = torch.load(path, weights_only=False)
    cfg = spec.VacnicConfig(**fx["cfg"])
    sd = spec.test_state_dict(cfg, fx["weight_seed"], lm_scale=fx["lm_scale"])
    sd["final_logits_bias"][0, cfg.eos_token_id] = fx.get("eos_bias", 0.0)
    batch = synthetic.make_batch(**fx["batch_kwargs"])
    gcfg = spec.VacnicConfig(**{**fx["cfg"], "stock": True, "only_image": False})
    gsd = spec.test_state_dict(gcfg, fx["guide_seed"])
    return fx, cfg, sd, batch, gcfg, gsd


def enc_inputs(cfg, batch):
    src = batch["article_ids"]
    d = dict(input_ids=src, attention_mask=OM.src_mask(src), image_features=batch["image_features"])
    if not cfg.only_image:
        face = batch["face_emb"]
        d.update(face_features=face, face_mask=OM.src_mask(face[:, :, -1]), name_ids=batch["names_art_ids"],
                 name_mask=OM.src_mask(batch["names_art_ids"]))
    return d


def test_goldens_exist():
    assert len(GOLDEN) >= 6


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-3] for p in GOLDEN])
def test_oracle_matches_reference_golden(path):
    fx, cfg, sd, batch, gcfg, gsd = load_case(path)
    # the regenerated weights / inputs are the ones the reference saw
    ws = float(sum(v.double().sum() for k, v in sd.items() if k not in spec.TIED_TO_SHARED))
    bs = float(sum(v.double().sum() for v in batch.values()))
    assert abs(ws - fx["weight_checksum"]) <= 1e-6 * max(1.0, abs(ws))
    assert abs(bs - fx["batch_checksum"]) <= 1e-6 * max(1.0, abs(bs))
    for k in sd:
        sd[k] = sd[k].clone().requires_grad_(sd[k].is_floating_point() and k != "final_logits_bias")
    for k in spec.TIED_TO_SHARED:
        sd[k] = sd["model.shared.weight"]
    o = OM.training_losses(sd, cfg.as_dict(), gsd, gcfg.as_dict(), batch)
    out = o["out"]
    tol = 2e-4
    lg = out["logits"].detach()
    assert (lg[..., fx["logit_cols"]] - fx["logits_at_cols"]).abs().max() <= tol * max(1.0, fx["lm_scale"])
    assert (torch.logsumexp(lg, -1) - fx["logits_lse"]).abs().max() <= tol * max(1.0, fx["lm_scale"])
    assert (out["decoder_hidden_states"][-1].detach() - fx["dec_h"]).abs().max() <= tol
    assert (out["encoder_last_hidden_state"].detach()[:, :8] - fx["enc_h_sample"]).abs().max() <= tol
    assert (out["hidden_states_img"].detach() - fx["img"]).abs().max() <= tol
    if not cfg.only_image:
        assert (out["hidden_states_face"].detach() - fx["face"]).abs().max() <= tol
        assert (out["hidden_states_ner"].detach()[:, :8] - fx["ner"]).abs().max() <= tol
    for k, v in fx["losses"].items():
        assert abs(float(o[k]) - v) <= 1e-5 * max(1.0, abs(v)), (k, float(o[k]), v)
    # gradients of the total loss (txt + 0.5 margin + secla), TRAIN:358-364
    o["loss"].backward()
    for k, g in fx["grad_samples"].items():
        got = sd[k].grad.flatten()[:512]
        assert (got - g).abs().max() <= 1e-4 * max(1e-3, g.abs().max().item()) + 1e-7, k
        assert abs(float(sd[k].grad.norm()) - fx["grad_norms"][k]) <= 1e-4 * fx["grad_norms"][k] + 1e-7, k


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-3] for p in GOLDEN])
def test_oracle_generation_matches_reference_generate(path):
    fx, cfg, sd, batch, _, _ = load_case(path)
    if "greedy_ids" not in fx:
        pytest.skip("no generation golden for this case")
    inp = enc_inputs(cfg, batch)
    g = OG.greedy(sd, cfg.as_dict(), inp, max_length=fx["max_length"])
    assert g.shape == fx["greedy_ids"].shape and bool((g == fx["greedy_ids"]).all())
    b, _ = OG.beam_search(sd, cfg.as_dict(), inp, num_beams=4, max_length=fx["max_length"], length_penalty=2.0)
    assert b.shape == fx["beam4_ids"].shape and bool((b == fx["beam4_ids"]).all())


def test_oracle_generation_matches_reference_generate_at_full_size():
    """BART-large (12 + 12 layers), ragged L = 1024 batch, max_length 50: the oracle decodes the ids the unmodified
    reference class produced through transformers' real `generate()` (tests/golden/make_golden_fullsize.py)."""
    path = os.path.join(os.path.dirname(__file__), "golden", "fullsize", "large_full_gen.pt")
    fx = torch.load(path, weights_only=False)
    cfg = spec.VacnicConfig(**fx["cfg"])
    sd = spec.test_state_dict(cfg, fx["weight_seed"], lm_scale=fx["lm_scale"])
    sd["final_logits_bias"][0, cfg.eos_token_id] = fx["eos_bias"]
    sd["final_logits_bias"][0, fx["logit_bias_idx"]] = fx["logit_bias_val"]
    batch = synthetic.make_batch(**fx["batch_kwargs"])
    chk = float(sum(v.double().sum() for k, v in sd.items() if k not in spec.TIED_TO_SHARED))
    assert abs(chk - fx["weight_checksum"]) <= 1e-6 * abs(fx["weight_checksum"])   # the seeded weights reproduce
    assert abs(float(sum(v.double().sum() for v in batch.values())) - fx["batch_checksum"]) <= 1e-6 * abs(fx["batch_checksum"])
    inp = enc_inputs(cfg, batch)
    with torch.no_grad():
        enc = OM.encoder_forward(sd, cfg.as_dict(), **inp)
    g = OG.greedy(sd, cfg.as_dict(), inp, max_length=fx["max_length"], enc=enc)
    assert g.shape == fx["greedy_ids"].shape and bool((g == fx["greedy_ids"]).all())
    b, _ = OG.beam_search(sd, cfg.as_dict(), inp, num_beams=fx["num_beams"], max_length=fx["max_length"],
                          length_penalty=fx["length_penalty"], enc=enc)
    assert b.shape == fx["beam4_ids"].shape and bool((b == fx["beam4_ids"]).all())


def test_fullsize_fixture_is_reproduced_by_its_generator():
    """The committed generator script rebuilds the levelled `final_logits_bias` entries stored in the full-size fixture
    (same tokens; values to fp32 re-execution accuracy) from the seeds alone."""
    import importlib.util
    here = os.path.dirname(__file__)
    spec_ = importlib.util.spec_from_file_location("make_golden_fullsize", os.path.join(here, "golden", "make_golden_fullsize.py"))
    mod = importlib.util.module_from_spec(spec_)
    spec_.loader.exec_module(mod)
    fx = torch.load(os.path.join(here, "golden", "fullsize", "large_full_gen.pt"), weights_only=False)
    assert (fx["weight_seed"], fx["lm_scale"], fx["max_length"], fx["num_beams"]) == (mod.WEIGHT_SEED, mod.LM_SCALE, mod.MAX_LEN, mod.NB)
    _, idx, val = mod.weights(spec.bart_large())
    assert idx.tolist() == fx["logit_bias_idx"].tolist()
    assert torch.allclose(val, fx["logit_bias_val"], rtol=0, atol=2e-3)
