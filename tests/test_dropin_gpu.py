"""The drop-in boundary: the reference's module paths, class name, constructor, forward / generate
signatures, state_dict names, whole-module pickling — driven the way the UNCHANGED reference training script drives
the model (forward -> script-level losses -> loss.backward() -> torch.optim.AdamW.step() -> zero_grad, TRAIN:281-374),
and checked against the fp32 oracle taking the same two optimisation steps."""
import importlib
import io
import os

import pytest
import torch

from oracle import model as OM
from vacnic_b200 import spec, synthetic

pytestmark = pytest.mark.gpu
MFULL = "src.models.modeling_mmbart_clip_inside_vis_clipcap_ent_type_final_fix_len_enc_self_face_name_ids_crossattn"
MVIS = "src.models.modeling_mmbart_clip_inside_vis_clipcap_ent_type_final_fix_len_enc_self_crossattn"


def _hf_config(cfg):
    from transformers import BartConfig
    return BartConfig(vocab_size=cfg.vocab, d_model=cfg.d_model, encoder_layers=cfg.enc_layers, decoder_layers=cfg.dec_layers,
                      encoder_attention_heads=cfg.heads, decoder_attention_heads=cfg.heads, encoder_ffn_dim=cfg.ffn,
                      decoder_ffn_dim=cfg.ffn, max_position_embeddings=cfg.max_pos, output_hidden_states=True, dropout=0.0)


def _ctor_kwargs(cfg):
    return dict(enc_fusion_layer=list(range(cfg.enc_layers)), dim_common=cfg.d_model, img_size=768, prompt_mlp_type="clipcap",
                map_size=[196, 256, 64, 16], prompt_size=cfg.prompt_size, clip_model=None, freeze_clip=False,
                max_ner_type_len=80, max_ner_type_len_gt=20, only_image=False, init_attn_weight=False)


def _inputs(cfg, batch):
    src = batch["article_ids"]
    kw = dict(input_ids=src, attention_mask=OM.src_mask(src), image_features=batch["image_features"])
    if not cfg.only_image:
        face = batch["face_emb"]
        kw.update(face_features=face, face_mask=OM.src_mask(face[:, :, -1]), name_ids=batch["names_art_ids"],
                  name_mask=OM.src_mask(batch["names_art_ids"]), add_ner_ffn=True)
    return kw


def test_reference_training_loop_runs_unchanged_and_tracks_the_oracle(cuda_device):
    mod = importlib.import_module(MFULL)
    cfg = spec.VacnicConfig(d_model=768, heads=12, ffn=1024, enc_layers=2, dec_layers=2, prompt_size=4, max_pos=128)
    sd = spec.test_state_dict(cfg, 31)
    model = mod.BartForMultiModalGeneration(_hf_config(cfg), **_ctor_kwargs(cfg))
    assert type(model).__module__ == MFULL and type(model).__name__ == "BartForMultiModalGeneration"
    ref_keys = set(spec.param_shapes(cfg).keys())
    assert set(model.state_dict().keys()) == ref_keys  # checkpoints interchange with the reference
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected
    batch = synthetic.to_device(synthetic.make_batch(B=2, L=40, T=12, seed=5), cuda_device)
    tgt = batch["caption_ids"]
    dec_in = mod.shift_tokens_right(tgt, 1, 2)
    # exactly what TRAIN:91-107 builds: AdamW over model.model + lm_head parameters
    params = list(model.model.parameters()) + list(model.lm_head.parameters())
    assert len({id(p) for p in params}) == len(params)
    opt = torch.optim.AdamW(params, lr=1e-3, weight_decay=0.01)
    ce = torch.nn.CrossEntropyLoss(ignore_index=1)
    # oracle twin (fp32 torch on the same device)
    osd = {k: v.to(cuda_device).clone().requires_grad_(v.is_floating_point() and k != "final_logits_bias") for k, v in sd.items()
           if k not in spec.TIED_TO_SHARED}
    for k in spec.TIED_TO_SHARED:
        osd[k] = osd["model.shared.weight"]
    uniq = list({id(v): v for v in osd.values() if v.requires_grad}.values())  # tied entries alias one tensor
    oopt = torch.optim.AdamW(uniq, lr=1e-3, weight_decay=0.01)
    model.train()
    losses, olosses = [], []
    for it in range(3):
        out = model(decoder_input_ids=dec_in, **_inputs(cfg, batch))
        logits = out["logits"]
        loss = ce(logits.reshape(-1, logits.shape[-1]), tgt.reshape(-1))      # TRAIN:287 (script-level torch op)
        h = out["decoder_hidden_states"][-1]
        loss = loss + 1e-3 * h.float().pow(2).mean() + 1e-3 * out["hidden_states_face"].float().pow(2).mean()
        loss.backward()
        opt.step()
        opt.zero_grad()                                                      # set_to_none=True: the views must come back
        losses.append(float(loss.detach()))
        o = OM.model_forward(osd, cfg.as_dict(), decoder_input_ids=dec_in, **{k: v for k, v in _inputs(cfg, batch).items() if k != "add_ner_ffn"})
        ol = ce(o["logits"].reshape(-1, o["logits"].shape[-1]), tgt.reshape(-1))
        ol = ol + 1e-3 * o["decoder_hidden_states"][-1].pow(2).mean() + 1e-3 * o["hidden_states_face"].pow(2).mean()
        ol.backward()
        oopt.step()
        oopt.zero_grad()
        olosses.append(float(ol.detach()))
    # the optimizer steps taken by torch are seen by the bf16 compute shadow: the loss moves, and moves like the oracle's
    assert losses[2] < losses[0] - 0.05, losses
    for a, b in zip(losses, olosses):
        assert abs(a - b) <= 2e-2 * max(1.0, abs(b)), (losses, olosses)


def test_pickle_roundtrip_generate_and_resize(cuda_device):
    mod = importlib.import_module(MFULL)
    cfg = spec.VacnicConfig(d_model=768, heads=12, ffn=1024, enc_layers=2, dec_layers=2, prompt_size=4, max_pos=128)
    model = mod.BartForMultiModalGeneration(_hf_config(cfg), **_ctor_kwargs(cfg))
    model.load_state_dict(spec.test_state_dict(cfg, 32, lm_scale=8.0), strict=False)
    model.eval()
    batch = synthetic.to_device(synthetic.make_batch(B=2, L=40, T=12, seed=6), cuda_device)
    kw = _inputs(cfg, batch)
    ids = model.generate(num_beams=4, max_length=10, length_penalty=2.0, **kw)      # INFER:798 call shape
    assert ids.dtype == torch.int64 and ids.shape[0] == 2 and ids.shape[1] <= 10 and bool((ids[:, 0] == 2).all())
    buf = io.BytesIO()
    torch.save(model, buf)                                                          # TRAIN:467
    buf.seek(0)
    clone = torch.load(buf, weights_only=False)                                     # INFER:1087
    assert type(clone).__module__ == MFULL
    ids2 = clone.generate(num_beams=4, max_length=10, length_penalty=2.0, **kw)
    assert torch.equal(ids, ids2)
    n0 = model.model.shared.weight.shape[0]
    model.resize_token_embeddings(n0 + 3)                                           # TRAIN:754
    assert model.model.shared.weight.shape[0] == n0 + 3 and model.lm_head.weight.shape[0] == n0 + 3
    out = model(decoder_input_ids=batch["caption_ids"], **kw)
    assert out["logits"].shape[-1] == n0 + 3


def test_only_visual_module_forward_signature(cuda_device):
    mod = importlib.import_module(MVIS)
    cfg = spec.VacnicConfig(d_model=1024, heads=16, ffn=1024, enc_layers=2, dec_layers=2, prompt_size=6, max_pos=128, only_image=True)
    kw = _ctor_kwargs(cfg)
    for k in ("max_ner_type_len", "max_ner_type_len_gt", "only_image", "init_attn_weight"):
        kw.pop(k)
    model = mod.BartForMultiModalGeneration(_hf_config(cfg), **kw)                  # TRAINVIS:538-542
    batch = synthetic.to_device(synthetic.make_batch(B=2, L=40, T=10, seed=7), cuda_device)
    src = batch["article_ids"]
    out = model(input_ids=src, attention_mask=OM.src_mask(src), decoder_input_ids=batch["caption_ids"],
                image_features=batch["image_features"])                            # TRAINVIS:172
    assert out["logits"].shape == (2, 10, cfg.vocab) and out["hidden_states_img"].shape == (2, 6, 1024)
    with pytest.raises(Exception):
        model(input_ids=src.cpu(), attention_mask=OM.src_mask(src).cpu(), decoder_input_ids=batch["caption_ids"].cpu(),
              image_features=batch["image_features"].cpu())                       # no CPU fallback


def test_save_pretrained_roundtrip_with_reference_names(cuda_device, tmp_path):
    """Checkpoint format: config.json + safetensors under the reference state_dict names."""
    from safetensors.torch import load_file
    mod = importlib.import_module(MFULL)
    cfg = spec.VacnicConfig(d_model=768, heads=12, ffn=1024, enc_layers=2, dec_layers=2, prompt_size=4, max_pos=128)
    model = mod.BartForMultiModalGeneration(_hf_config(cfg), **_ctor_kwargs(cfg))
    model.load_state_dict(spec.test_state_dict(cfg, 33), strict=False)
    model.eval()
    out_dir = model.save_pretrained(str(tmp_path / "ckpt"))
    saved = load_file(os.path.join(out_dir, "model.safetensors"))
    want = set(spec.param_shapes(cfg).keys()) - set(spec.TIED_TO_SHARED)
    assert set(saved.keys()) == want
    clone = mod.BartForMultiModalGeneration.from_pretrained(out_dir, **_ctor_kwargs(cfg))
    batch = synthetic.to_device(synthetic.make_batch(B=2, L=40, T=12, seed=8), cuda_device)
    kw = _inputs(cfg, batch)
    with torch.no_grad():
        a = model(decoder_input_ids=batch["caption_ids"], **kw)["logits"]
        b = clone(decoder_input_ids=batch["caption_ids"], **kw)["logits"]
    assert torch.equal(a, b)
    for k, v in clone.state_dict().items():
        assert torch.equal(v.cpu(), model.state_dict()[k].cpu()), k
