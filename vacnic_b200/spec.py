"""Parameter inventory of the VACNIC multimodal BART (names and shapes identical to the reference
`state_dict`, so checkpoints interchange) and deterministic test initialisation.

Reference constructors: BartForMultiModalGeneration MFULL:1881-1898, BartModel MFULL:1703-1715,
BartEncoder MFULL:1099-1164, BartEncoderLayer MFULL:569-616, BartDecoder MFULL:1393-1425,
BartDecoderLayer MFULL:765-791, BartAttention MFULL:421-449 (only-visual variant: MVIS, same
minus the face/name modules).  `stock=True` is an unmodified HF BART (the CoLaM guide, TRAIN:745).
"""
from __future__ import annotations

from collections import OrderedDict
from dataclasses import dataclass, asdict
from typing import Dict, Tuple

import torch

NER_VOCAB = 50267  # embed_tokens_ner is hard-wired to 50267 rows, MFULL:1150
CLIP_DIM = 768     # MLPClipCap is hard-wired to 768-d CLIP ln_post(CLS) features, MFULL:1136
FACE_DIM = 512     # FaceNet embeddings, _linear_1 MFULL:1162
FACE_FFN = 3072    # _face_up / _face_down hidden width, MFULL:607-608


@dataclass
class VacnicConfig:
    d_model: int = 1024
    heads: int = 16
    ffn: int = 4096
    enc_layers: int = 12
    dec_layers: int = 12
    vocab: int = 50267
    max_pos: int = 1024
    prompt_size: int = 20
    max_ner_type_len: int = 80
    max_ner_type_len_gt: int = 20
    only_image: bool = False
    stock: bool = False
    pad_token_id: int = 1
    decoder_start_token_id: int = 2
    eos_token_id: int = 2

    def as_dict(self) -> dict:
        return asdict(self)

    @property
    def head_dim(self) -> int:
        return self.d_model // self.heads


def bart_large(**kw) -> VacnicConfig:
    return VacnicConfig(**kw)


def bart_base(**kw) -> VacnicConfig:
    base = dict(d_model=768, heads=12, ffn=3072, enc_layers=6, dec_layers=6, prompt_size=10)
    base.update(kw)
    return VacnicConfig(**base)


def _attn(shapes, p, d):
    # registration order inside BartAttention: k_proj, v_proj, q_proj, out_proj (MFULL:444-447)
    for n in ("k_proj", "v_proj", "q_proj", "out_proj"):
        shapes[f"{p}.{n}.weight"] = (d, d)
        shapes[f"{p}.{n}.bias"] = (d,)


def _ln(shapes, p, d):
    shapes[p + ".weight"] = (d,)
    shapes[p + ".bias"] = (d,)


def _lin(shapes, p, out_f, in_f, bias=True):
    shapes[p + ".weight"] = (out_f, in_f)
    if bias:
        shapes[p + ".bias"] = (out_f,)


def param_shapes(cfg: VacnicConfig) -> "OrderedDict[str, Tuple[int, ...]]":
    """Ordered name -> shape of every entry of the reference state_dict (parameters + the
    `final_logits_bias` buffer)."""
    d, f = cfg.d_model, cfg.ffn
    s: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    s["final_logits_bias"] = (1, cfg.vocab)
    s["model.shared.weight"] = (cfg.vocab, d)
    e = "model.encoder."
    s[e + "embed_tokens.weight"] = (cfg.vocab, d)
    s[e + "embed_positions.weight"] = (cfg.max_pos + 2, d)
    for i in range(cfg.enc_layers):
        p = f"{e}layers.{i}."
        _attn(s, p + "self_attn", d)
        _ln(s, p + "self_attn_layer_norm", d)
        _lin(s, p + "fc1", f, d)
        _lin(s, p + "fc2", d, f)
        _ln(s, p + "final_layer_norm", d)
        if not cfg.stock:
            _lin(s, p + "_linear_1up", f, d)
            _lin(s, p + "_linear_1down", d, f)
            _ln(s, p + "img_layer_norm", d)
            if not cfg.only_image:
                _lin(s, p + "ner_map_up", 4 * cfg.max_ner_type_len_gt, cfg.max_ner_type_len)
                _lin(s, p + "ner_map_down", cfg.max_ner_type_len_gt, 4 * cfg.max_ner_type_len_gt)
                _ln(s, p + "ner_map_layer_norm", d)
                _attn(s, p + "self_attn_img_name", d)
                _ln(s, p + "img_name_attn_layer_norm", d)
                _lin(s, p + "_face_up", FACE_FFN, d)
                _lin(s, p + "_face_down", d, FACE_FFN)
                _ln(s, p + "face_layer_norm", d)
                _attn(s, p + "cross_attn_img_ner", d)
                _ln(s, p + "img_ner_attn_layer_norm", d)
            else:
                # MVIS:560-589 keeps the prefix cross-attention but not the face/name modules
                _attn(s, p + "cross_attn_img_ner", d)
                _ln(s, p + "img_ner_attn_layer_norm", d)
    _ln(s, e + "layernorm_embedding", d)
    if not cfg.stock:
        _lin(s, e + "prompt_mlp.model.0", CLIP_DIM * cfg.prompt_size // 2, CLIP_DIM)
        _lin(s, e + "prompt_mlp.model.2", CLIP_DIM * cfg.prompt_size, CLIP_DIM * cfg.prompt_size // 2)
        if d == 1024:
            _lin(s, e + "visual_map", 1024, CLIP_DIM)
        if not cfg.only_image:
            s[e + "embed_tokens_ner.weight"] = (NER_VOCAB, d)
            s[e + "embed_positions_ner.weight"] = (cfg.max_pos + 2, d)
            _ln(s, e + "layernorm_embedding_ner", d)
        _lin(s, e + "_linear_1", d, FACE_DIM)  # present (unused) in MVIS too: MVIS:1076
    q = "model.decoder."
    s[q + "embed_tokens.weight"] = (cfg.vocab, d)
    s[q + "embed_positions.weight"] = (cfg.max_pos + 2, d)
    for i in range(cfg.dec_layers):
        p = f"{q}layers.{i}."
        _attn(s, p + "self_attn", d)
        _ln(s, p + "self_attn_layer_norm", d)
        _attn(s, p + "encoder_attn", d)
        _ln(s, p + "encoder_attn_layer_norm", d)
        _lin(s, p + "fc1", f, d)
        _lin(s, p + "fc2", d, f)
        _ln(s, p + "final_layer_norm", d)
    _ln(s, q + "layernorm_embedding", d)
    s["lm_head.weight"] = (cfg.vocab, d)
    return s


# entries that alias model.shared.weight inside the reference module (same Parameter object)
TIED_TO_SHARED = ("model.encoder.embed_tokens.weight", "model.decoder.embed_tokens.weight")


def test_state_dict(cfg: VacnicConfig, seed: int, lm_scale: float = 1.0, device="cpu") -> Dict[str, torch.Tensor]:
    """Deterministic fp32 weights for parity tests.  Matrices and embeddings ~ N(0, 0.02) like
    `_init_weights` (MFULL:899-908) with the pad row zeroed; unlike the reference init, biases and
    LayerNorm parameters are perturbed so that a wrong bias/affine path cannot hide.  `lm_scale`
    multiplies lm_head to widen top-1/top-2 logit margins for token-id parity."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    for name, shape in param_shapes(cfg).items():
        if name in TIED_TO_SHARED:
            continue
        is_ln = "layer_norm" in name or "layernorm" in name
        if name == "final_logits_bias":
            t = torch.zeros(shape)
        elif is_ln and name.endswith(".weight"):
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif is_ln:
            t = 0.1 * torch.randn(shape, generator=g)
        elif name.endswith(".bias"):
            t = 0.02 * torch.randn(shape, generator=g)
        else:
            t = 0.02 * torch.randn(shape, generator=g)
            if "embed_tokens" in name or name == "model.shared.weight":
                t[cfg.pad_token_id].zero_()
        if name == "lm_head.weight":
            t = t * lm_scale
        sd[name] = t.to(device)
    for name in TIED_TO_SHARED:
        sd[name] = sd["model.shared.weight"]
    if cfg.stock:  # HF BartForConditionalGeneration ties lm_head to the shared embedding
        sd["lm_head.weight"] = sd["model.shared.weight"]
    return sd
