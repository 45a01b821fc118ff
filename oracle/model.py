"""CPU/fp32 restatement of the VACNIC multimodal-BART hot path.  TEST INFRASTRUCTURE ONLY.

This file is the *checker*: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs may import it.  The product (vacnic_b200/, src/models/) never does.

Parity status: PINNED.  The reference ships no tests or golden vectors, so the pin is
the reference itself: tests/golden/make_golden.py imports the unmodified reference classes from
/root/reference in the build container, checks this restatement against them (fp32, CPU, max-abs
<= 1e-5 on every output) and commits golden vectors that tests/test_oracle_golden.py re-checks
on any host.

Everything is a pure function of a `state_dict` (reference parameter names) and a small config
dict; plain torch ops, no nn.Module, so that it runs on CPU or (for comparisons on the GPU box) on a
CUDA device in fp32.  Citations: MFULL = src/models/modeling_mmbart_clip_inside_vis_clipcap_ent_
type_final_fix_len_enc_self_face_name_ids_crossattn.py, MVIS = ..._enc_self_crossattn.py,
TRAIN = train_mmbart_enc_self_face_name_ids_retrieve_crossattn_bart_guide_match.py.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


def make_cfg(d_model=1024, heads=16, ffn=4096, enc_layers=12, dec_layers=12, vocab=50267, max_pos=1024,
             prompt_size=20, max_ner_type_len=80, max_ner_type_len_gt=20, only_image=False, stock=False,
             pad_token_id=1, decoder_start_token_id=2, eos_token_id=2):
    """`stock=True` describes an unmodified HF BART (the frozen CoLaM guide, TRAIN:745)."""
    return dict(d_model=d_model, heads=heads, ffn=ffn, enc_layers=enc_layers, dec_layers=dec_layers, vocab=vocab,
                max_pos=max_pos, prompt_size=prompt_size, max_ner_type_len=max_ner_type_len,
                max_ner_type_len_gt=max_ner_type_len_gt, only_image=only_image, stock=stock,
                pad_token_id=pad_token_id, decoder_start_token_id=decoder_start_token_id, eos_token_id=eos_token_id)


# ------------------------------------------------------------------------------------------ pieces
def linear(sd: SD, name: str, x: torch.Tensor) -> torch.Tensor:
    return F.linear(x, sd[name + ".weight"], sd.get(name + ".bias"))


def layer_norm(sd: SD, name: str, x: torch.Tensor) -> torch.Tensor:
    return F.layer_norm(x, (x.shape[-1],), sd[name + ".weight"], sd[name + ".bias"], 1e-5)


def gelu(x):  # ACT2FN["gelu"] = exact erf GELU (MFULL:579)
    return F.gelu(x)


def expand_mask(mask: torch.Tensor, dtype, tgt_len: Optional[int] = None) -> torch.Tensor:
    """_expand_mask, MFULL:387-398: [B,S] {0,1} -> additive [B,1,tgt,S] with finfo.min at masked keys."""
    bsz, src_len = mask.shape
    tgt_len = src_len if tgt_len is None else tgt_len
    inv = 1.0 - mask[:, None, None, :].expand(bsz, 1, tgt_len, src_len).to(dtype)
    return inv.masked_fill(inv.to(torch.bool), torch.finfo(dtype).min)


def causal_mask(tgt_len: int, dtype, device, past: int = 0) -> torch.Tensor:
    """_make_causal_mask, MFULL:373-385."""
    m = torch.full((tgt_len, tgt_len), torch.finfo(dtype).min, dtype=dtype, device=device)
    cond = torch.arange(tgt_len, device=device)
    m.masked_fill_(cond < (cond + 1).view(tgt_len, 1), 0)
    if past > 0:
        m = torch.cat([torch.zeros(tgt_len, past, dtype=dtype, device=device), m], dim=-1)
    return m[None, None]


def attention(sd: SD, name: str, heads: int, x: torch.Tensor, kv: Optional[torch.Tensor] = None,
              mask: Optional[torch.Tensor] = None, past=None, return_kv: bool = False):
    """BartAttention.forward, MFULL:454-565.  `past` = (k, v) as [B,H,S,hd]: for cross-attention it is
    reused untouched (:474-477), for self-attention the new k/v are appended (:482-487)."""
    B, T, d = x.shape
    hd = d // heads

    def shape(t):
        return t.view(B, -1, heads, hd).transpose(1, 2)

    q = shape(linear(sd, name + ".q_proj", x) * hd ** -0.5)  # scale after bias, MFULL:472
    if kv is not None and past is not None:
        k, v = past
    elif kv is not None:
        k, v = shape(linear(sd, name + ".k_proj", kv)), shape(linear(sd, name + ".v_proj", kv))
    else:
        k, v = shape(linear(sd, name + ".k_proj", x)), shape(linear(sd, name + ".v_proj", x))
        if past is not None:
            k, v = torch.cat([past[0], k], dim=2), torch.cat([past[1], v], dim=2)
    s = q @ k.transpose(-1, -2)
    if mask is not None:
        s = s + mask
    p = torch.softmax(s, dim=-1)
    o = (p @ v).transpose(1, 2).reshape(B, T, d)
    o = linear(sd, name + ".out_proj", o)
    return (o, (k, v)) if return_kv else o


def embed(sd: SD, prefix: str, ids: torch.Tensor, tok: str, pos: str, ln: str, past: int = 0) -> torch.Tensor:
    """LN(E[ids] * 1.0 + Pos[arange + 2]); MFULL:1243-1249 / 1254-1260 / 1555-1563, offset 2 MFULL:409-418."""
    S = ids.shape[-1]
    positions = torch.arange(past, past + S, device=ids.device) + 2
    # nn.Embedding(..., padding_idx=1): the pad row receives no gradient (MFULL:1115, 1150, 1400)
    h = F.embedding(ids, sd[prefix + tok + ".weight"], padding_idx=1) + sd[prefix + pos + ".weight"][positions]
    return layer_norm(sd, prefix + ln, h)


# ------------------------------------------------------------------------------------------ encoder
def encoder_layer(sd: SD, cfg, p: str, h, self_mask, img=None, face=None, ner=None, face_name_mask=None):
    """BartEncoderLayer.forward with every layer a fusion layer (MFULL:645-744; MVIS:591-690).
    Returns (h, face, ner, img)."""
    H = cfg["heads"]
    if not cfg["stock"]:
        # 1. image-prefix FFN, MFULL:647-653
        img = layer_norm(sd, p + "img_layer_norm", img + linear(sd, p + "_linear_1down", gelu(linear(sd, p + "_linear_1up", img))))
        if not cfg["only_image"]:
            # 2. face FFN, MFULL:658-664
            face = layer_norm(sd, p + "face_layer_norm", face + linear(sd, p + "_face_down", gelu(linear(sd, p + "_face_up", face))))
            # 3. names attend to [faces ; names], MFULL:669-679
            ner = layer_norm(sd, p + "img_name_attn_layer_norm",
                             ner + attention(sd, p + "self_attn_img_name", H, ner, kv=torch.cat((face, ner), dim=1), mask=face_name_mask))
            # 4. NER prefix map: reshape is a memory reinterpretation, not a transpose, MFULL:682-688
            B, E, d = ner.shape
            G = cfg["max_ner_type_len_gt"]
            z = gelu(linear(sd, p + "ner_map_up", ner.reshape(B, d, E)))
            z = linear(sd, p + "ner_map_down", z).reshape(B, G, d)
            prefix = layer_norm(sd, p + "ner_map_layer_norm", z)
            kv = torch.cat((img, prefix), dim=1)  # MFULL:691
        else:
            kv = img  # MVIS:626
    # 6. self attention, MFULL:697-707
    h = layer_norm(sd, p + "self_attn_layer_norm", h + attention(sd, p + "self_attn", H, h, mask=self_mask))
    if not cfg["stock"]:
        # 7. prefix cross attention with an all-zero additive mask, MFULL:711-723, 1282-1296
        h = layer_norm(sd, p + "img_ner_attn_layer_norm", h + attention(sd, p + "cross_attn_img_ner", H, h, kv=kv))
    # 8. FFN, MFULL:738-744
    h = layer_norm(sd, p + "final_layer_norm", h + linear(sd, p + "fc2", gelu(linear(sd, p + "fc1", h))))
    return h, face, ner, img


def encoder_forward(sd: SD, cfg, input_ids, attention_mask, image_features=None, face_features=None,
                    face_mask=None, name_ids=None, name_mask=None, prefix="model.encoder."):
    """BartEncoder.forward, MFULL:1172-1381 (only-visual: MVIS:1086-1251)."""
    h = embed(sd, prefix, input_ids, "embed_tokens", "embed_positions", "layernorm_embedding")
    dtype = h.dtype
    img = face = ner = face_name_mask = None
    if not cfg["stock"]:
        if not cfg["only_image"]:
            ner = embed(sd, prefix, name_ids, "embed_tokens_ner", "embed_positions_ner", "layernorm_embedding_ner")
            fm = torch.cat((face_mask, name_mask), dim=1)
            face_name_mask = expand_mask(fm, dtype, tgt_len=cfg["max_ner_type_len"])  # MFULL:1262-1264
            face = linear(sd, prefix + "_linear_1", face_features)  # MFULL:1269
        z = linear(sd, prefix + "prompt_mlp.model.2", torch.tanh(linear(sd, prefix + "prompt_mlp.model.0", image_features)))
        img = z.reshape(z.shape[0], cfg["prompt_size"], 768)  # MFULL:1274-1276
        if cfg["d_model"] == 1024:
            img = linear(sd, prefix + "visual_map", img)  # MFULL:1277-1278
    self_mask = expand_mask(attention_mask, dtype)
    states = []
    for i in range(cfg["enc_layers"]):
        states.append(h)
        h, face, ner, img = encoder_layer(sd, cfg, f"{prefix}layers.{i}.", h, self_mask, img, face, ner, face_name_mask)
    states.append(h)
    return dict(last_hidden_state=h, hidden_states=tuple(states), hidden_states_img=img, hidden_states_ner=ner,
                hidden_states_face=face)


# ------------------------------------------------------------------------------------------ decoder
def decoder_forward(sd: SD, cfg, decoder_input_ids, enc_out, enc_mask, past=None, use_cache=False,
                    prefix="model.decoder."):
    """BartDecoder.forward + BartDecoderLayer.forward, MFULL:1453-1675, 793-890.
    `past` = list over layers of (self_k, self_v, cross_k, cross_v)."""
    B, T = decoder_input_ids.shape
    past_len = past[0][0].shape[2] if past is not None else 0
    h = embed(sd, prefix, decoder_input_ids, "embed_tokens", "embed_positions", "layernorm_embedding", past=past_len)
    self_mask = causal_mask(T, h.dtype, h.device, past_len) if T > 1 else None  # MFULL:1438
    cross_mask = expand_mask(enc_mask, h.dtype, tgt_len=T)
    H = cfg["heads"]
    states, cache = [], []
    for i in range(cfg["dec_layers"]):
        states.append(h)
        p = f"{prefix}layers.{i}."
        lp = past[i] if past is not None else None
        a, skv = attention(sd, p + "self_attn", H, h, mask=self_mask, past=lp[:2] if lp else None, return_kv=True)
        h = layer_norm(sd, p + "self_attn_layer_norm", h + a)
        a, ckv = attention(sd, p + "encoder_attn", H, h, kv=enc_out, mask=cross_mask, past=lp[2:] if lp else None,
                           return_kv=True)
        h = layer_norm(sd, p + "encoder_attn_layer_norm", h + a)
        h = layer_norm(sd, p + "final_layer_norm", h + linear(sd, p + "fc2", gelu(linear(sd, p + "fc1", h))))
        if use_cache:
            cache.append(skv + ckv)
    states.append(h)
    return dict(last_hidden_state=h, hidden_states=tuple(states), past_key_values=cache if use_cache else None)


def lm_logits(sd: SD, h):
    """lm_head(dec_out) + final_logits_bias, MFULL:1997."""
    return F.linear(h, sd["lm_head.weight"]) + sd["final_logits_bias"]


def model_forward(sd: SD, cfg, input_ids, attention_mask, decoder_input_ids, image_features=None,
                  face_features=None, face_mask=None, name_ids=None, name_mask=None):
    """BartForMultiModalGeneration.forward, MFULL:1929-2021 (add_ner_ffn=True; False is broken upstream)."""
    enc = encoder_forward(sd, cfg, input_ids, attention_mask, image_features, face_features, face_mask, name_ids, name_mask)
    dec = decoder_forward(sd, cfg, decoder_input_ids, enc["last_hidden_state"], attention_mask)
    return dict(logits=lm_logits(sd, dec["last_hidden_state"]), decoder_hidden_states=dec["hidden_states"],
                encoder_last_hidden_state=enc["last_hidden_state"], encoder_hidden_states=enc["hidden_states"],
                hidden_states_face=enc["hidden_states_face"], hidden_states_ner=enc["hidden_states_ner"],
                hidden_states_img=enc["hidden_states_img"])


# ------------------------------------------------------------------------------------------ losses
def shift_tokens_right(ids, pad_token_id=1, decoder_start_token_id=2):
    """MFULL:340-353 = TRAIN:196-209."""
    out = ids.new_zeros(ids.shape)
    out[:, 1:] = ids[:, :-1]
    out[:, 0] = decoder_start_token_id
    out.masked_fill_(out == -100, pad_token_id)
    return out


def src_mask(ids):
    """create_src_mask_bart, TRAIN:212-217."""
    return (ids != 1).to(torch.int64)


def token_ce(logits, tgt_ids, pad=1):
    """CrossEntropyLoss(ignore_index=pad), TRAIN:816, 287."""
    return F.cross_entropy(logits.reshape(-1, logits.shape[-1]).float(), tgt_ids.reshape(-1), ignore_index=pad)


def pool(h, mask):
    """TRAIN:178-182."""
    s = h.masked_fill(~mask[..., None].bool(), 0.0).sum(dim=1) / mask.sum(dim=1)[..., None]
    return torch.nan_to_num(s, nan=1.0)


def colam_loss(h, h_guide, tgt_ids, margin=1.0):
    """CoLaM margin loss, TRAIN:292-309, HingeEmbeddingLoss(margin) with target -1 (TRAIN:820)."""
    m = src_mask(tgt_ids)
    a, b = pool(h, m), pool(h_guide, m)
    a = a / a.norm(dim=1, keepdim=True)
    b = b / b.norm(dim=1, keepdim=True)
    diag = (a @ b.t()).diag()
    return F.hinge_embedding_loss(diag, -torch.ones_like(diag), margin=margin)


def names_embedding(sd: SD, names_ids_3d, prefix="model.encoder."):
    """get_embedding_ner, TRAIN:112-133: per span mean over ALL positions (pads included) of the NER
    embedding LayerNorm output."""
    out = []
    for i in range(names_ids_3d.shape[1]):
        h = embed(sd, prefix, names_ids_3d[:, i, :], "embed_tokens_ner", "embed_positions_ner", "layernorm_embedding_ner")
        out.append(h.mean(dim=1))
    return torch.stack(out, dim=1)


def _batch_softmax(match):
    """batch_softmax, TRAIN:631-647."""
    B, _, n, _ = match.shape
    logits = match.max(-1).values.sum(-1) / n
    return F.cross_entropy(logits, torch.arange(B, device=logits.device))


def secla_loss(face, names):
    """BatchSoftmax.forward, TRAIN:654-660.  face [B,F,d], names [B,N,d]."""
    a = torch.matmul(names.unsqueeze(1), face.permute(0, 2, 1))
    b = torch.matmul(face.unsqueeze(1), names.permute(0, 2, 1))
    return _batch_softmax(a) + _batch_softmax(b)


def training_losses(sd: SD, cfg, guide_sd: Optional[SD], guide_cfg, batch, margin=1.0, alpha=0.5, w_secla=1.0):
    """The loss block of train_epoch (TRAIN:267-363) for `--use_secla True --no_clip_loss True`."""
    src, tgt = batch["article_ids"], batch["caption_ids"]
    dec_in = shift_tokens_right(tgt, 1, 2)  # start id = eos, TRAIN:267
    sm = src_mask(src)
    if cfg["only_image"]:
        out = model_forward(sd, cfg, src, sm, dec_in, image_features=batch["image_features"])
    else:
        face = batch["face_emb"]
        out = model_forward(sd, cfg, src, sm, dec_in, image_features=batch["image_features"], face_features=face,
                            face_mask=src_mask(face[:, :, -1]), name_ids=batch["names_art_ids"],
                            name_mask=src_mask(batch["names_art_ids"]))
    res = dict(out=out, txt=token_ce(out["logits"], tgt))
    loss = res["txt"]
    if guide_sd is not None:
        g_enc = encoder_forward(guide_sd, guide_cfg, src, sm)
        g_dec = decoder_forward(guide_sd, guide_cfg, dec_in, g_enc["last_hidden_state"], sm)
        res["margin"] = colam_loss(out["decoder_hidden_states"][-1], g_dec["last_hidden_state"], tgt, margin)
        loss = loss + alpha * res["margin"]
    if not cfg["only_image"]:
        with torch.no_grad():
            names = names_embedding(sd, batch["names_ids"])
        res["secla"] = secla_loss(out["hidden_states_face"], names)
        loss = loss + w_secla * res["secla"]
    res["loss"] = loss
    return res
