"""B200-native VACNIC multimodal BART: the module tree, parameter names and forward signature of the
reference (`BartForMultiModalGeneration`, MFULL:1877-2074 / MVIS:1731-1892), executed by hand-written
sm_100a kernels through the C ABI (vacnic_b200.blocks).  The nn.Linear / nn.LayerNorm / nn.Embedding
members exist as *parameter containers* with the reference's names (so `state_dict`s interchange and
the training script's helper code that reaches into `model.model.encoder.embed_tokens_ner` etc.,
TRAIN:117-129, keeps working); their own torch `forward` is never used on the hot path.

There is no CPU fallback: every forward needs a CUDA device and the built libvacnic_b200.so.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
from torch import nn

from . import blocks as Bk
from . import kernels as K
from .blocks import LN, Runtime
from .spec import CLIP_DIM, FACE_DIM, FACE_FFN, NER_VOCAB, VacnicConfig
from .store import ParamStore


class VacnicOutput(dict):
    """Minimal stand-in for transformers' ModelOutput: key, attribute and index access over the
    non-None fields in declaration order (Seq2SeqLMOutput MFULL:255-311, BaseModelOutput MFULL:125-149)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __getitem__(self, k):
        if isinstance(k, (int, slice)):
            return tuple(v for v in self.values() if v is not None)[k]
        return dict.__getitem__(self, k)

    def to_tuple(self):
        return tuple(v for v in self.values() if v is not None)


def _meta_linear(i, o, bias=True):
    return nn.Linear(i, o, bias=bias, device="meta")


class BartLearnedPositionalEmbedding(nn.Embedding):
    """MFULL:401-418: ids offset by 2."""

    def __init__(self, num_embeddings: int, embedding_dim: int):
        self.offset = 2
        super().__init__(num_embeddings + self.offset, embedding_dim, device="meta")

    def forward(self, input_ids_shape, past_key_values_length: int = 0):
        seq_len = input_ids_shape[1]
        positions = torch.arange(past_key_values_length, past_key_values_length + seq_len, dtype=torch.long,
                                 device=self.weight.device)
        return super().forward(positions + self.offset)


class BartAttention(nn.Module):
    def __init__(self, embed_dim: int, num_heads: int, dropout: float = 0.0, is_decoder: bool = False):
        super().__init__()
        self.embed_dim, self.num_heads, self.dropout = embed_dim, num_heads, dropout
        self.head_dim = embed_dim // num_heads
        if self.head_dim * num_heads != embed_dim:
            raise ValueError(f"embed_dim must be divisible by num_heads (got `embed_dim`: {embed_dim} and `num_heads`: {num_heads}).")
        self.scaling = self.head_dim ** -0.5
        self.is_decoder = is_decoder
        self.k_proj = _meta_linear(embed_dim, embed_dim)
        self.v_proj = _meta_linear(embed_dim, embed_dim)
        self.q_proj = _meta_linear(embed_dim, embed_dim)
        self.out_proj = _meta_linear(embed_dim, embed_dim)

    def bind(self, rt: Runtime, ln_mod, fused_self: bool):
        st = rt.store
        self.lin_qkv = st.lin([self.k_proj.weight, self.v_proj.weight, self.q_proj.weight],
                              [self.k_proj.bias, self.v_proj.bias, self.q_proj.bias]) if fused_self else None
        self.lin_q = st.lin([self.q_proj.weight], [self.q_proj.bias])
        self.lin_kv = st.lin([self.k_proj.weight, self.v_proj.weight], [self.k_proj.bias, self.v_proj.bias])
        self.lin_o = st.lin([self.out_proj.weight], [self.out_proj.bias])
        self.ln = LN(st, ln_mod, rt.new_salt())


class BartEncoderLayer(nn.Module):
    def __init__(self, cfg: VacnicConfig):
        super().__init__()
        d, H = cfg.d_model, cfg.heads
        self.cfg = cfg
        self.embed_dim = d
        self.self_attn = BartAttention(d, H)
        self.self_attn_layer_norm = nn.LayerNorm(d, device="meta")
        self.fc1 = _meta_linear(d, cfg.ffn)
        self.fc2 = _meta_linear(cfg.ffn, d)
        self.final_layer_norm = nn.LayerNorm(d, device="meta")
        if not cfg.stock:
            self._linear_1up = _meta_linear(d, cfg.ffn)
            self._linear_1down = _meta_linear(cfg.ffn, d)
            self.img_layer_norm = nn.LayerNorm(d, device="meta")
            self.only_image = cfg.only_image
            if not cfg.only_image:
                self.ner_map_up = _meta_linear(cfg.max_ner_type_len, 4 * cfg.max_ner_type_len_gt)
                self.ner_map_down = _meta_linear(4 * cfg.max_ner_type_len_gt, cfg.max_ner_type_len_gt)
                self.ner_map_layer_norm = nn.LayerNorm(d, device="meta")
                self.max_ner_type_len_gt = cfg.max_ner_type_len_gt
                self.self_attn_img_name = BartAttention(d, H)
                self.img_name_attn_layer_norm = nn.LayerNorm(d, device="meta")
                self._face_up = _meta_linear(d, FACE_FFN)
                self._face_down = _meta_linear(FACE_FFN, d)
                self.face_layer_norm = nn.LayerNorm(d, device="meta")
            self.cross_attn_img_ner = BartAttention(d, H)
            self.img_ner_attn_layer_norm = nn.LayerNorm(d, device="meta")

    def bind(self, rt: Runtime):
        st, cfg = rt.store, self.cfg
        self.rt = rt
        self.self_attn.bind(rt, self.self_attn_layer_norm, True)
        self.lin_fc1 = st.lin([self.fc1.weight], [self.fc1.bias])
        self.lin_fc2 = st.lin([self.fc2.weight], [self.fc2.bias])
        self.ln_final = LN(st, self.final_layer_norm, rt.new_salt())
        if not cfg.stock:
            self.lin_iup = st.lin([self._linear_1up.weight], [self._linear_1up.bias])
            self.lin_idown = st.lin([self._linear_1down.weight], [self._linear_1down.bias])
            self.ln_img = LN(st, self.img_layer_norm, rt.new_salt())
            self.cross_attn_img_ner.bind(rt, self.img_ner_attn_layer_norm, False)
            if not cfg.only_image:
                self.lin_nup = st.lin([self.ner_map_up.weight], [self.ner_map_up.bias])
                self.lin_ndown = st.lin([self.ner_map_down.weight], [self.ner_map_down.bias])
                self.ln_nmap = LN(st, self.ner_map_layer_norm, rt.new_salt())
                self.self_attn_img_name.bind(rt, self.img_name_attn_layer_norm, False)
                self.lin_fup = st.lin([self._face_up.weight], [self._face_up.bias])
                self.lin_fdown = st.lin([self._face_down.weight], [self._face_down.bias])
                self.ln_face = LN(st, self.face_layer_norm, rt.new_salt())

    def forward(self, h, key_mask, img=None, face=None, ner=None, face_name_mask=None, pack=None):
        """BartEncoderLayer.forward with every layer a fusion layer (MFULL:645-744 / MVIS:591-690).  Every state is a
        pair (bf16 tensor the GEMMs read, fp32 copy carried as the residual -- or None).  `pack` (varlen.ArticlePack):
        the article rows `h` are packed [1, rows, d]; the attention kernels get row ranges instead of `key_mask`."""
        kv, img, face, ner = self.side(img, face, ner, face_name_mask)
        return self.main(h, key_mask, kv, pack), face, ner, img

    def side(self, img=None, face=None, ner=None, face_name_mask=None):
        """The prefix side of the layer (MFULL:645-693): image FFN, face FFN, name attention over [faces; names], NER-prefix
        map.  It reads and writes only the img / face / ner states -- never the article rows -- so the side chains of all
        layers form one dependency chain of their own, which the encoder may run on a second stream (BartEncoder.forward).
        Returns (kv, img, face, ner): kv = the prefix rows this layer's article rows cross-attend to."""
        rt, cfg, H = self.rt, self.cfg, self.cfg.heads
        kv = None
        (img, img32), (face, face32), (ner, ner32) = img or (None, None), face or (None, None), ner or (None, None)
        if not cfg.stock:
            img, img32 = Bk.MlpBlockFn.apply(img, img32, rt.fwd_anchor, rt, self.lin_iup, self.lin_idown, K.ACT_GELU, self.ln_img)
            img_kv, img = Bk.fanout(img, 2)
            if not cfg.only_image:
                face, face32 = Bk.MlpBlockFn.apply(face, face32, rt.fwd_anchor, rt, self.lin_fup, self.lin_fdown, K.ACT_GELU,
                                                   self.ln_face)
                face_kv, face = Bk.fanout(face, 2)
                ner_q, ner_kv = Bk.fanout(ner, 2)
                a = self.self_attn_img_name
                ner, ner32 = Bk.AttnBlockFn.apply(ner_q, ner32, Bk.Concat2Fn.apply(face_kv, ner_kv), None, rt, None, a.lin_q,
                                                  a.lin_kv, a.lin_o, a.ln, H, face_name_mask, False, False, 0, None, False)
                ner_map, ner = Bk.fanout(ner, 2)
                prefix = Bk.NerMapFn.apply(ner_map, rt, self.lin_nup, self.lin_ndown, self.ln_nmap)
                kv = Bk.Concat2Fn.apply(img_kv, prefix)
            else:
                kv = img_kv
        return kv, (img, img32), (face, face32), (ner, ner32)

    def main(self, h, key_mask, kv, pack=None):
        """The article side of the layer (MFULL:695-744): self-attention, cross-attention to the prefix rows `kv`, FFN."""
        rt, cfg, H = self.rt, self.cfg, self.cfg.heads
        h, h32 = h
        a = self.self_attn
        h, h32 = Bk.AttnBlockFn.apply(h, h32, None, None, rt, a.lin_qkv, None, None, a.lin_o, a.ln, H,
                                      key_mask if pack is None else pack.self_mask, False, True, 0, None, False)
        if not cfg.stock:
            a = self.cross_attn_img_ner
            h, h32 = Bk.AttnBlockFn.apply(h, h32, kv, None, rt, None, a.lin_q, a.lin_kv, a.lin_o, a.ln, H,
                                          None if pack is None else pack.prefix_mask, False, True, 0, None, False)
        h, h32 = Bk.MlpBlockFn.apply(h, h32, rt.fwd_anchor, rt, self.lin_fc1, self.lin_fc2, K.ACT_GELU, self.ln_final)
        return (h, h32)


class BartDecoderLayer(nn.Module):
    def __init__(self, cfg: VacnicConfig):
        super().__init__()
        d, H = cfg.d_model, cfg.heads
        self.cfg = cfg
        self.embed_dim = d
        self.self_attn = BartAttention(d, H, is_decoder=True)
        self.self_attn_layer_norm = nn.LayerNorm(d, device="meta")
        self.encoder_attn = BartAttention(d, H, is_decoder=True)
        self.encoder_attn_layer_norm = nn.LayerNorm(d, device="meta")
        self.fc1 = _meta_linear(d, cfg.ffn)
        self.fc2 = _meta_linear(cfg.ffn, d)
        self.final_layer_norm = nn.LayerNorm(d, device="meta")

    def bind(self, rt: Runtime):
        st = rt.store
        self.rt = rt
        self.self_attn.bind(rt, self.self_attn_layer_norm, True)
        self.encoder_attn.bind(rt, self.encoder_attn_layer_norm, False)
        self.lin_fc1 = st.lin([self.fc1.weight], [self.fc1.bias])
        self.lin_fc2 = st.lin([self.fc2.weight], [self.fc2.bias])
        self.ln_final = LN(st, self.final_layer_norm, rt.new_salt())


class BartEncoder(nn.Module):
    def __init__(self, cfg: VacnicConfig, embed_tokens: nn.Embedding):
        super().__init__()
        d = cfg.d_model
        self.cfg = cfg
        self.dropout = 0.1
        self.embed_dim = d
        self.padding_idx = cfg.pad_token_id
        self.embed_scale = 1.0  # scale_embedding=False (MFULL:1112)
        self.embed_tokens = embed_tokens
        self.embed_positions = BartLearnedPositionalEmbedding(cfg.max_pos, d)
        self.layers = nn.ModuleList([BartEncoderLayer(cfg) for _ in range(cfg.enc_layers)])
        self.layernorm_embedding = nn.LayerNorm(d, device="meta")
        if not cfg.stock:
            P = cfg.prompt_size
            self.prompt_mlp = nn.Module()
            self.prompt_mlp.model = nn.Sequential(_meta_linear(CLIP_DIM, CLIP_DIM * P // 2), nn.Tanh(),
                                                  _meta_linear(CLIP_DIM * P // 2, CLIP_DIM * P))
            self.prompt_size = P
            self.prompt_mlp_type = "clipcap"
            if d == 1024:
                self.visual_map = _meta_linear(CLIP_DIM, 1024)
            elif d != CLIP_DIM:
                raise ValueError("d_model must be 768 or 1024: the reference feeds 768-d prefix tokens directly "
                                 "(MFULL:1276) and only maps to 1024 (MFULL:1142)")
            self.only_image = cfg.only_image
            if not cfg.only_image:
                self.embed_tokens_ner = nn.Embedding(NER_VOCAB, d, cfg.pad_token_id, device="meta")
                self.embed_positions_ner = BartLearnedPositionalEmbedding(cfg.max_pos, d)
                self.layernorm_embedding_ner = nn.LayerNorm(d, device="meta")
            self.max_ner_type_len = cfg.max_ner_type_len
            self.max_ner_type_len_gt = cfg.max_ner_type_len_gt
            self._linear_1 = _meta_linear(FACE_DIM, d)

    def bind(self, rt: Runtime):
        st, cfg = rt.store, self.cfg
        self.rt = rt
        self.ln_emb = LN(st, self.layernorm_embedding, rt.new_salt())
        if not cfg.stock:
            m = self.prompt_mlp.model
            self.lin_p0 = st.lin([m[0].weight], [m[0].bias])
            self.lin_p2 = st.lin([m[2].weight], [m[2].bias])
            if cfg.d_model == 1024:
                self.lin_vmap = st.lin([self.visual_map.weight], [self.visual_map.bias])
            if not cfg.only_image:
                self.ln_emb_ner = LN(st, self.layernorm_embedding_ner, rt.new_salt())
                self.lin_face = st.lin([self._linear_1.weight], [self._linear_1.bias])
        for l in self.layers:
            l.bind(rt)

    def forward(self, input_ids=None, attention_mask=None, image_features=None, name_ids=None, name_mask=None,
                face_features=None, face_mask=None, add_ner_ffn=True, output_hidden_states=True, pack=None, **unused):
        """BartEncoder.forward, MFULL:1172-1381 (only-visual MVIS:1086-1251).  `pack` (varlen.ArticlePack, extension): the
        article tokens arrive packed (no padding rows); `input_ids` / `attention_mask` are then ignored and
        `last_hidden_state` is the packed [1, rows, d] memory."""
        rt, cfg = self.rt, self.cfg
        if not (pack.ids if pack is not None else input_ids).is_cuda:
            raise K._l.VacnicError("vacnic_b200 runs on CUDA tensors only (no CPU fallback)")
        if not add_ner_ffn:
            raise ValueError("add_ner_ffn=False is broken in the reference (mask size mismatch, MFULL:666 vs :1296) "
                             "and is not provided")
        rt.training = self.training
        st = rt.store
        if pack is not None:
            B, key_mask = pack.B, None
            h = Bk.EmbedFn.apply(rt.fwd_anchor, pack.ids.view(1, -1), rt, self.embed_tokens.weight, self.embed_positions.weight,
                                 self.ln_emb, 2, cfg.pad_token_id, pack.pos)
        else:
            B, L = input_ids.shape
            if attention_mask is None:
                attention_mask = torch.ones_like(input_ids)
            key_mask = Bk.KeyMask(attention_mask)
            h = Bk.EmbedFn.apply(rt.fwd_anchor, input_ids.contiguous(), rt, self.embed_tokens.weight, self.embed_positions.weight,
                                 self.ln_emb, 2, cfg.pad_token_id)
        # The prefix side (ClipCap MLP, visual_map, name / face embeddings and the img / face / ner chain of every layer)
        # never reads the article rows: with `rt.side_stream` set (TrainStep, VACNIC_SIDE_STREAM) it is enqueued as a whole
        # on a second, high-priority stream and the article side of layer i only waits for that layer's prefix rows, so its
        # ~25 small latency-bound kernels per layer run in the wave tails of the article-side GEMMs instead of between them.
        # The autograd engine runs each backward node on the stream of its forward, so the backward pass forks the same way.
        side = rt.side_stream if (not cfg.stock and torch.is_grad_enabled()) else None
        main = torch.cuda.current_stream()
        if side is not None:
            side.wait_stream(main)
        kvs = []
        with torch.cuda.stream(side if side is not None else main):
            img = face = ner = fn_mask = None
            if not cfg.stock:
                if not cfg.only_image:
                    ner = Bk.EmbedFn.apply(rt.fwd_anchor, name_ids.contiguous(), rt, self.embed_tokens_ner.weight,
                                           self.embed_positions_ner.weight, self.ln_emb_ner, 2, cfg.pad_token_id)
                    fn_mask = Bk.KeyMask(torch.cat((face_mask, name_mask), dim=1))  # MFULL:1262
                    face = (Bk.LinearFn.apply(face_features.to(torch.bfloat16), rt.fwd_anchor, rt, self.lin_face, torch.bfloat16,
                                              False, None, None), None)
                z, _ = Bk.MlpBlockFn.apply(image_features.to(torch.bfloat16), None, rt.fwd_anchor, rt, self.lin_p0, self.lin_p2,
                                           K.ACT_TANH, None)
                img = z.view(B, cfg.prompt_size, CLIP_DIM)  # MFULL:1276
                if cfg.d_model == 1024:
                    img = Bk.LinearFn.apply(img, rt.fwd_anchor, rt, self.lin_vmap, torch.bfloat16, True, None, None)
                img = (img, None)
            if side is not None:
                for layer in self.layers:
                    kv, img, face, ner = layer.side(img, face, ner, fn_mask)
                    ev = torch.cuda.Event()
                    ev.record(side)
                    kvs.append((kv, ev))
                # tensors that cross the two streams stay referenced until the step is over (rt.keepalive is cleared by the
                # owner of the step after the streams have joined): the caching allocator recycles a block on the stream
                # that allocated it as soon as the last reference dies, which the OTHER stream's pending kernels cannot see
                rt.keepalive += [kv for kv, _ in kvs] + [t for pair in (img, face, ner) if pair for t in pair if t is not None]
                if fn_mask is not None:
                    rt.keepalive += [fn_mask.mask, fn_mask.len]
        states = []
        for i, layer in enumerate(self.layers):
            if output_hidden_states:
                states.append(h[0])
            h = (Bk.grad_mark(h[0], rt, ("enc", i)), h[1])  # backward: gradients of encoder layers >= i are final
            if side is not None:
                kv, ev = kvs[i]
                main.wait_event(ev)
                h = layer.main(h, key_mask, kv, pack)
            else:
                h, face, ner, img = layer(h, key_mask, img, face, ner, fn_mask, pack)
        if side is not None:
            main.wait_stream(side)  # the final img / face / ner states are read on the main stream from here on
        h, img, face, ner = h[0], img and img[0], face and face[0], ner and ner[0]
        if output_hidden_states:
            states.append(h)
        # field order matters: encoder_outputs[-1/-2/-3] = face / ner / img (MFULL:1852-1854)
        return VacnicOutput(last_hidden_state=h, hidden_states=tuple(states) if output_hidden_states else None,
                            attentions=None, hidden_states_img=img, hidden_states_ner=ner, hidden_states_face=face)


class BartDecoder(nn.Module):
    def __init__(self, cfg: VacnicConfig, embed_tokens: nn.Embedding):
        super().__init__()
        d = cfg.d_model
        self.cfg = cfg
        self.dropout = 0.1
        self.padding_idx = cfg.pad_token_id
        self.embed_scale = 1.0
        self.embed_tokens = embed_tokens
        self.embed_positions = BartLearnedPositionalEmbedding(cfg.max_pos, d)
        self.layers = nn.ModuleList([BartDecoderLayer(cfg) for _ in range(cfg.dec_layers)])
        self.layernorm_embedding = nn.LayerNorm(d, device="meta")
        self.embed_dim = d

    def bind(self, rt: Runtime):
        st = rt.store
        self.rt = rt
        self.ln_emb = LN(st, self.layernorm_embedding, rt.new_salt())
        for l in self.layers:
            l.bind(rt)
        # hoisted cross-attention K/V projection of the encoder memory: all layers in ONE GEMM
        ws, bs = [], []
        for l in self.layers:
            ws += [l.encoder_attn.k_proj.weight, l.encoder_attn.v_proj.weight]
            bs += [l.encoder_attn.k_proj.bias, l.encoder_attn.v_proj.bias]
        self.lin_cross_kv = st.lin(ws, bs)

    def forward(self, input_ids=None, attention_mask=None, encoder_hidden_states=None, encoder_attention_mask=None,
                output_hidden_states=True, pack=None, **unused):
        """BartDecoder.forward (training / teacher-forced path), MFULL:1453-1675.  `pack`: `encoder_hidden_states` is the
        packed [1, rows, d] memory of varlen.ArticlePack (cross-attention gets row ranges instead of the mask)."""
        rt, cfg = self.rt, self.cfg
        rt.training = self.training
        st = rt.store
        B, T = input_ids.shape
        d, H = cfg.d_model, cfg.heads
        x, x32 = Bk.EmbedFn.apply(rt.fwd_anchor, input_ids.contiguous(), rt, self.embed_tokens.weight, self.embed_positions.weight,
                                  self.ln_emb, 2, cfg.pad_token_id)
        if pack is not None:
            if T != pack.cross_mask.geo.max_q:
                raise ValueError("ArticlePack was built for a different decoder length")
            enc_mask = pack.cross_mask
        else:
            enc_mask = None if encoder_attention_mask is None else Bk.KeyMask(encoder_attention_mask)
        dec_mask = None if attention_mask is None else Bk.KeyMask(attention_mask)
        # backward: once this marker fires, the decoder (incl. the hoisted cross K/V projection) and the LM head are done
        encoder_hidden_states = Bk.grad_mark(encoder_hidden_states, rt, ("dec", 0))
        kv_all = Bk.LinearFn.apply(encoder_hidden_states, rt.fwd_anchor, rt, self.lin_cross_kv, torch.bfloat16, True, None, None)
        dkv_all = torch.empty_like(kv_all) if (torch.is_grad_enabled() and kv_all.requires_grad) else None
        if dkv_all is not None and pack is not None:
            dkv_all[:, -pack.cross_mask.k_tail:].zero_()  # bucket-padding rows: no caption attends to them, nobody writes them
        states = []
        for i, l in enumerate(self.layers):
            if output_hidden_states:
                states.append(x)
            a = l.self_attn
            x, x32 = Bk.AttnBlockFn.apply(x, x32, None, None, rt, a.lin_qkv, None, None, a.lin_o, a.ln, H, dec_mask, T > 1, True, 0,
                                          None, False)
            a = l.encoder_attn
            x, x32 = Bk.AttnBlockFn.apply(x, x32, None, kv_all, rt, None, a.lin_q, None, a.lin_o, a.ln, H, enc_mask, False, True,
                                          i * 2 * d, dkv_all, i == 0)
            x, x32 = Bk.MlpBlockFn.apply(x, x32, rt.fwd_anchor, rt, l.lin_fc1, l.lin_fc2, K.ACT_GELU, l.ln_final)
        if output_hidden_states:
            states.append(x)
        return VacnicOutput(last_hidden_state=x, past_key_values=None,
                            hidden_states=tuple(states) if output_hidden_states else None)


class BartModel(nn.Module):
    def __init__(self, cfg: VacnicConfig):
        super().__init__()
        self.cfg = cfg
        self.shared = nn.Embedding(cfg.vocab, cfg.d_model, cfg.pad_token_id, device="meta")
        self.encoder = BartEncoder(cfg, self.shared)
        self.decoder = BartDecoder(cfg, self.shared)
        self.embed_dim = cfg.d_model

    def get_encoder(self):
        return self.encoder

    def get_decoder(self):
        return self.decoder


class VacnicBart(nn.Module):
    """Core of `BartForMultiModalGeneration` (MFULL:1877) without the transformers plumbing; the drop-in
    classes in src/models/ add `from_pretrained` / `generate` on top of it."""

    def __init__(self, cfg: VacnicConfig, device="cuda", p_drop: float = 0.1, seed: int = 0, tie_lm_head: bool = False,
                 frozen: bool = False, symmetric: Optional[bool] = None, p_attn: float = 0.0, p_act: float = 0.0):
        super().__init__()
        self.cfg = cfg
        self.model = BartModel(cfg)
        self.register_buffer("final_logits_bias", torch.zeros((1, cfg.vocab), device=device))
        self.lm_head = _meta_linear(cfg.d_model, cfg.vocab, bias=False)
        if tie_lm_head or cfg.stock:
            self.lm_head.weight = self.model.shared.weight  # tie_word_embeddings (stock HF BART, TRAIN:745)
        first = []
        for i in range(cfg.dec_layers):
            first += [f"model.decoder.layers.{i}.encoder_attn.k_proj.weight", f"model.decoder.layers.{i}.encoder_attn.v_proj.weight"]
        for i in range(cfg.dec_layers):
            first += [f"model.decoder.layers.{i}.encoder_attn.k_proj.bias", f"model.decoder.layers.{i}.encoder_attn.v_proj.bias"]
        if symmetric is None:
            # one process per GPU under torch.distributed: gradient buffer and bf16 shadow live in symmetric memory so the
            # rank-sharded optimizer step can run over NVLink peer loads / stores (VACNIC_DP_P2P=0 keeps plain allocations
            # and the NCCL all-reduce path)
            import os
            dist = torch.distributed
            symmetric = (not frozen and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
                         and dist.get_backend() == "nccl" and os.environ.get("VACNIC_DP_P2P", "1") != "0")
        self.store = ParamStore(self, device, first=first, frozen=frozen, symmetric=symmetric)
        self.rt = Runtime(self.store, p_drop=p_drop, seed=seed, p_attn=p_attn, p_act=p_act)
        self.model.encoder.bind(self.rt)
        self.model.decoder.bind(self.rt)
        self.lin_lm = self.store.lin([self.lm_head.weight], None)
        self.init_weights(seed)
        self.clip_model = None

    # ------------------------------------------------------------------ weights
    @torch.no_grad()
    def init_weights(self, seed: int = 0, std: float = 0.02):
        """BartPretrainedModel._init_weights, MFULL:899-908: N(0, std) matrices and embeddings (pad row zero),
        zero biases, unit LayerNorm."""
        g = torch.Generator(device=self.store.device).manual_seed(seed)
        for n, p in self.store.params.items():
            if "layer_norm" in n or "layernorm" in n:
                p.data.fill_(1.0 if n.endswith("weight") else 0.0)
            elif n.endswith(".bias"):
                p.data.zero_()
            else:
                p.data.normal_(0.0, std, generator=g)
                if "embed_tokens" in n or "shared" in n:
                    p.data[self.cfg.pad_token_id].zero_()
        self.store.refresh_shadow()

    def load_reference_state_dict(self, sd, strict=True):
        """Load a reference-format state_dict (names as in vacnic_b200.spec.param_shapes)."""
        if "final_logits_bias" in sd:
            self.final_logits_bias.copy_(sd["final_logits_bias"].to(self.final_logits_bias))
        self.store.load_state_dict_flat(sd, strict=strict)

    def _apply(self, fn, recurse=True):  # .to()/.cuda() must not re-allocate the flat views
        return self

    def train(self, mode: bool = True):
        super().train(mode)
        self.rt.training = mode
        # `.data` edits (TRAIN:758) do not bump version counters: every train() / eval() call (once per epoch in the scripts,
        # TRAIN train_epoch) re-casts the bf16 shadow on the next forward.  Per-step callers (trainer.TrainStep,
        # generation.Generator) do not go through here unless the mode really changes.
        self.store._shadow_version = -1
        return self

    def get_encoder(self):
        return self.model.encoder

    def get_decoder(self):
        return self.model.decoder

    def state_dict(self, *args, **kwargs):
        self.store.sync_master()  # rank-sharded optimizer: COLLECTIVE (call on every rank, like FSDP's state_dict)
        return super().state_dict(*args, **kwargs)

    # ------------------------------------------------------------------ forward
    def forward(self, input_ids=None, attention_mask=None, decoder_input_ids=None, decoder_attention_mask=None,
                head_mask=None, decoder_head_mask=None, cross_attn_head_mask=None, encoder_outputs=None,
                past_key_values=None, inputs_embeds=None, decoder_inputs_embeds=None, labels=None, use_cache=None,
                output_attentions=None, output_hidden_states=None, return_dict=None, image_features=None,
                face_features=None, face_mask=None, name_ids=None, name_mask=None, add_ner_ffn=True, ce_targets=None,
                article_pack=None, need_logits: bool = True):
        """Signature of BartForMultiModalGeneration.forward (MFULL:1929-1953; MVIS:1783-1802 lacks the four
        face/name arguments).  `ce_targets` (extension): fuse the script's CrossEntropyLoss(ignore_index=pad)
        (TRAIN:287) into the LM head; the result is returned under "loss".  `article_pack` (extension,
        varlen.ArticlePack): the article arrives as packed rows instead of padded `input_ids` + `attention_mask`;
        `encoder_last_hidden_state` is then the packed memory (varlen.unpack_rows restores the padded layout).
        `need_logits=False` (extension): stop after the decoder -- the CoLaM guide (TRAIN:293) is only read for
        `decoder_hidden_states[-1]`, its LM head (a 1024 x 50267 x 1024 GEMM + 206 MB of fp32 logits) is never looked at."""
        cfg = self.cfg
        for unsupported, name in ((head_mask, "head_mask"), (decoder_head_mask, "decoder_head_mask"),
                                  (cross_attn_head_mask, "cross_attn_head_mask"), (inputs_embeds, "inputs_embeds"),
                                  (decoder_inputs_embeds, "decoder_inputs_embeds"), (past_key_values, "past_key_values")):
            if unsupported is not None:
                raise NotImplementedError(f"{name} is never passed by the VACNIC scripts and is not provided; "
                                          "cached decoding lives in vacnic_b200.generation")
        if output_attentions:
            raise NotImplementedError("attention maps are not materialised by the fused path")
        if self.store.dirty_shadow:
            self.store.refresh_shadow()
        if torch.is_grad_enabled() and self.training and not self.store.frozen and not self.store.external_step:
            # the unchanged reference loop (forward, loss.backward(), optimizer.step(), zero_grad; TRAIN:281-374): every
            # training forward starts a fresh gradient step in the flat buffer (p.grad views are re-attached after
            # zero_grad(set_to_none=True)); gradient accumulation over several forwards needs vacnic_b200.trainer
            self.store.begin_step()
            # end-of-backward node of the plain loop; under torch.distributed (the script wraps the model in
            # DistributedDataParallel, TRAIN:86-87 / TRAINVIS:84) it also makes DDP's reducer hooks fire -- see ParamTouchFn
            dist = torch.distributed
            multi = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
            self.rt.fwd_anchor = Bk.ParamTouchFn.apply(self.store.anchor, self.store,
                                                       *(tuple(self.store.params.values()) if multi else ()))
        else:
            self.rt.fwd_anchor = self.store.anchor
        if labels is not None and decoder_input_ids is None:
            decoder_input_ids = shift_tokens_right(labels, cfg.pad_token_id, cfg.decoder_start_token_id)
        if decoder_input_ids is None:
            if input_ids is None:
                raise ValueError("If no `decoder_input_ids` or `decoder_inputs_embeds` are passed, `input_ids` cannot be "
                                 "`None`. Please pass either `input_ids` or `decoder_input_ids` or `decoder_inputs_embeds`.")
            decoder_input_ids = shift_tokens_right(input_ids, cfg.pad_token_id, cfg.decoder_start_token_id)
        if encoder_outputs is None:
            encoder_outputs = self.model.encoder(input_ids=input_ids, attention_mask=attention_mask,
                                                 image_features=image_features, name_ids=name_ids, name_mask=name_mask,
                                                 face_features=face_features, face_mask=face_mask, add_ner_ffn=add_ner_ffn,
                                                 pack=article_pack)
        enc_h = encoder_outputs["last_hidden_state"] if isinstance(encoder_outputs, dict) else encoder_outputs[0]
        dec = self.model.decoder(input_ids=decoder_input_ids, attention_mask=decoder_attention_mask,
                                 encoder_hidden_states=enc_h, encoder_attention_mask=attention_mask, pack=article_pack)
        x = dec["last_hidden_state"]
        if not need_logits:
            eo = encoder_outputs
            return VacnicOutput(loss=None, logits=None, past_key_values=None, decoder_hidden_states=dec["hidden_states"],
                                decoder_attentions=None, cross_attentions=None, encoder_last_hidden_state=enc_h)
        x_lm, x_out = Bk.fanout(x, 2)
        loss = None
        flb_lin = Bk.Lin(self.lin_lm.w16, self.final_logits_bias.view(-1), self.lin_lm.gw, None, self.lin_lm.key)
        tgt = ce_targets if ce_targets is not None else labels
        if tgt is not None:
            loss, logits = Bk.LmHeadCeFn.apply(x_lm, self.rt, flb_lin, tgt, cfg.pad_token_id if ce_targets is not None else -100)
        else:
            B, T, d = x.shape
            ldp = (cfg.vocab + 7) // 8 * 8
            buf = torch.empty(B * T, ldp, dtype=torch.float32, device=x.device)
            logits = Bk.LinearFn.apply(x_lm, self.rt.fwd_anchor, self.rt, flb_lin, torch.float32, True, buf[:, :cfg.vocab], None)
        dec_states = dec["hidden_states"]
        if dec_states is not None:
            dec_states = dec_states[:-1] + (x_out,)
        eo = encoder_outputs
        return VacnicOutput(
            loss=loss, logits=logits, past_key_values=None, decoder_hidden_states=dec_states, decoder_attentions=None,
            cross_attentions=None, encoder_last_hidden_state=enc_h,
            encoder_hidden_states=eo["hidden_states"] if isinstance(eo, dict) else None, encoder_attentions=None,
            hidden_states_face=eo["hidden_states_face"] if isinstance(eo, dict) else None,
            hidden_states_ner=eo["hidden_states_ner"] if isinstance(eo, dict) else None,
            hidden_states_img=eo["hidden_states_img"] if isinstance(eo, dict) else None)


def shift_tokens_right(input_ids: torch.Tensor, pad_token_id: int, decoder_start_token_id: int):
    """MFULL:340-353 (= TRAIN:196-209): host-side index shuffling on int64 ids."""
    shifted = input_ids.new_zeros(input_ids.shape)
    shifted[:, 1:] = input_ids[:, :-1].clone()
    shifted[:, 0] = decoder_start_token_id
    if pad_token_id is None:
        raise ValueError("self.model.config.pad_token_id has to be defined.")
    shifted.masked_fill_(shifted == -100, pad_token_id)
    return shifted
