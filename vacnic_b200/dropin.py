"""Drop-in classes behind the reference's `src.models` module paths.

`BartForMultiModalGeneration` keeps the reference constructor (MFULL:1881 / MVIS:1735), `from_pretrained(plm,
**vacnic_kwargs)` (TRAIN:743), `forward` (MFULL:1929-1953), `generate` (INFER:798, 867), `resize_token_embeddings`
(MFULL:1906-1918), `get_encoder/get_decoder`, the module tree the training script reaches into (TRAIN:117-129,
140-152, 758) and whole-module pickling (TRAIN:467, INFER:1087) — and routes all device arithmetic to the
sm_100a kernels of libvacnic_b200.so.  transformers is used only for `BartConfig` parsing when available.
"""
from __future__ import annotations

import glob
import os
from typing import Optional

import torch

from .modeling import VacnicBart
from .spec import VacnicConfig

_CTOR_KEYS = ("enc_fusion_layer", "dim_common", "img_size", "prompt_mlp_type", "map_size", "prompt_size", "clip_model",
              "freeze_clip", "max_ner_type_len", "max_ner_type_len_gt", "only_image", "init_attn_weight")


def _cfg_get(config, name, default=None):
    if isinstance(config, dict):
        return config.get(name, default)
    return getattr(config, name, default)


def to_vacnic_config(config, prompt_size, max_ner_type_len, max_ner_type_len_gt, only_image) -> VacnicConfig:
    d = _cfg_get(config, "d_model", 1024)
    if _cfg_get(config, "scale_embedding", False):
        raise NotImplementedError("scale_embedding=True is not used by any VACNIC configuration")
    if _cfg_get(config, "activation_function", "gelu") != "gelu":
        raise NotImplementedError("only the exact-erf GELU of BART (ACT2FN['gelu'], MFULL:579) is provided")
    return VacnicConfig(
        d_model=d, heads=_cfg_get(config, "encoder_attention_heads", 16), ffn=_cfg_get(config, "encoder_ffn_dim", 4096),
        enc_layers=_cfg_get(config, "encoder_layers", 12), dec_layers=_cfg_get(config, "decoder_layers", 12),
        vocab=_cfg_get(config, "vocab_size", 50265), max_pos=_cfg_get(config, "max_position_embeddings", 1024),
        prompt_size=prompt_size, max_ner_type_len=max_ner_type_len, max_ner_type_len_gt=max_ner_type_len_gt,
        only_image=only_image, pad_token_id=_cfg_get(config, "pad_token_id", 1),
        decoder_start_token_id=_cfg_get(config, "decoder_start_token_id", 2), eos_token_id=_cfg_get(config, "eos_token_id", 2))


# what the device-side search implements (transformers `_beam_search` / greedy with a default BartConfig, see
# vacnic_b200.generation); any other value -- passed as a kwarg or inherited from the checkpoint's config /
# generation_config the way HF `generate()` (INFER:798) inherits it -- would change the captions, so it RAISES
_GEN_TOKEN_KEYS = ("eos_token_id", "forced_eos_token_id", "pad_token_id", "bos_token_id", "decoder_start_token_id")
GEN_IMPLEMENTED = {
    "do_sample": (False,), "early_stopping": (False,), "num_return_sequences": (1,), "num_beam_groups": (1,),
    "no_repeat_ngram_size": (0, None), "encoder_no_repeat_ngram_size": (0, None), "repetition_penalty": (1.0, None),
    "encoder_repetition_penalty": (1.0, None), "diversity_penalty": (0.0, None), "min_length": (0, None),
    "min_new_tokens": (None,), "max_new_tokens": (None,), "temperature": (1.0, None), "top_k": (50, None), "top_p": (1.0, None),
    "bad_words_ids": (None,), "force_words_ids": (None,), "forced_bos_token_id": (None,), "suppress_tokens": (None,),
    "begin_suppress_tokens": (None,), "exponential_decay_length_penalty": (None,), "renormalize_logits": (False,),
    "remove_invalid_values": (False, None), "constraints": (None,), "prefix_allowed_tokens_fn": (None,),
    "logits_processor": (None,), "stopping_criteria": (None,), "penalty_alpha": (None,), "output_scores": (False,),
    "output_attentions": (False, True), "output_hidden_states": (False, True),  # no effect on the returned ids (TRAIN:743)
    "output_logits": (False, None),
    "return_dict_in_generate": (False,), "use_cache": (True, False), "assistant_model": (None,), "streamer": (None,),
    "synced_gpus": (False, None), "sequence_bias": (None,), "guidance_scale": (None, 1.0), "decoder_input_ids": (None,),
}

def check_generation_request(cfg: VacnicConfig, config, generation_config, kwargs: dict):
    """Raise NotImplementedError for every generate() argument or inherited checkpoint setting the device-side search does
    not implement (host-only logic, CPU-testable)."""
    sources = [("generate() argument", kwargs)]
    for name, obj in (("generation_config", generation_config), ("config", config)):
        if obj is not None:
            d = obj.to_dict() if hasattr(obj, "to_dict") else dict(obj)
            sources.append((name, {k: v for k, v in d.items() if k not in kwargs}))  # explicit arguments win, as in HF
    for where, d in sources:
        for k, v in d.items():
            if k in GEN_IMPLEMENTED:
                ok = GEN_IMPLEMENTED[k]
                if v is not None and not any(o is not None and v == o for o in ok):
                    raise NotImplementedError(
                        f"generate(): {k}={v!r} (from {where}) is not implemented by the vacnic_b200 search "
                        f"(implemented: {ok[0]!r}); pass {k}={ok[0]!r} explicitly to override a checkpoint setting")
            elif where == "generate() argument" and k not in _GEN_TOKEN_KEYS:
                raise NotImplementedError(f"generate(): unknown / unimplemented argument {k!r}")
    eos = kwargs.get("eos_token_id", cfg.eos_token_id)
    forced_eos = kwargs.get("forced_eos_token_id", _cfg_get(config, "forced_eos_token_id", cfg.eos_token_id))
    if eos != cfg.eos_token_id or forced_eos != cfg.eos_token_id:
        raise NotImplementedError("generate(): eos_token_id / forced_eos_token_id other than the model's EOS are not implemented")
    for k, want in (("pad_token_id", cfg.pad_token_id), ("bos_token_id", None), ("decoder_start_token_id", cfg.decoder_start_token_id)):
        if k in kwargs and want is not None and kwargs[k] != want:
            raise NotImplementedError(f"generate(): {k}={kwargs[k]!r} differs from the model configuration ({want})")


class _DropInBase(VacnicBart):
    ONLY_IMAGE_FILE = False  # the MVIS module has no face/name branch at all

    def __init__(self, config, enc_fusion_layer=None, dim_common=256, img_size=2048, prompt_mlp_type="clipcap",
                 map_size=(192, 256, 64, 16), prompt_size=10, clip_model=None, freeze_clip=False, max_ner_type_len=80,
                 max_ner_type_len_gt=20, only_image=False, init_attn_weight=False, device=None, seed=0):
        if prompt_mlp_type != "clipcap":
            raise NotImplementedError("only --prompt_mlp_type clipcap (the shipped scripts, run_full_train.sh:25) is provided")
        n_enc = _cfg_get(config, "encoder_layers", 12)
        if enc_fusion_layer is not None and sorted(enc_fusion_layer) != list(range(n_enc)):
            raise NotImplementedError("every encoder layer is a fusion layer in all shipped configurations "
                                      "(run_full_train.sh:8); partial fusion is not provided")
        if init_attn_weight:
            raise NotImplementedError("--init_attn_weight True is not used by the shipped scripts")
        if only_image and not self.ONLY_IMAGE_FILE:
            raise NotImplementedError("only_image=True is broken in the reference's full-model file (UnboundLocalError at "
                                      "MFULL:1355); use the ..._enc_self_crossattn module")
        d_model = _cfg_get(config, "d_model", 1024)
        if not (self.ONLY_IMAGE_FILE or only_image) and dim_common != d_model:
            raise ValueError(f"dim_common ({dim_common}) must equal d_model ({d_model}): face states are concatenated "
                             "with name states (MFULL:668)")
        cfg = to_vacnic_config(config, prompt_size, max_ner_type_len, max_ner_type_len_gt, self.ONLY_IMAGE_FILE or only_image)
        if device is None:
            if not torch.cuda.is_available():
                raise RuntimeError("vacnic_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
            device = torch.device("cuda", torch.cuda.current_device())
        self._ctor = dict(prompt_size=prompt_size, max_ner_type_len=max_ner_type_len, max_ner_type_len_gt=max_ner_type_len_gt,
                          only_image=only_image, dim_common=dim_common)
        super().__init__(cfg, device=device, p_drop=float(_cfg_get(config, "dropout", 0.1)), seed=seed,
                         p_attn=float(_cfg_get(config, "attention_dropout", 0.0) or 0.0),
                         p_act=float(_cfg_get(config, "activation_dropout", 0.0) or 0.0))
        object.__setattr__(self, "config", config)
        self.clip_model = clip_model
        if freeze_clip and clip_model is not None:
            for p in clip_model.parameters():
                p.requires_grad = False

    # ------------------------------------------------------------------ loading
    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, *model_args, **kwargs):
        """`cls.from_pretrained(plm, output_hidden_states=True, enc_fusion_layer=..., ...)` (TRAIN:743).  Only local
        checkpoints can be read (no network): a directory holding config.json and pytorch_model.bin /
        model.safetensors.  BART weights are loaded by name; the VACNIC-specific modules keep their N(0, 0.02)
        initialisation, and the NER embedding tables are seeded from the shared table like MFULL:1150-1152."""
        ctor = {k: kwargs.pop(k) for k in list(kwargs) if k in _CTOR_KEYS or k in ("device", "seed")}
        try:
            from transformers import BartConfig
            config = BartConfig.from_pretrained(pretrained_model_name_or_path, **kwargs)
        except ImportError:  # pragma: no cover
            import json
            with open(os.path.join(pretrained_model_name_or_path, "config.json")) as f:
                config = json.load(f)
        model = cls(config, **ctor)
        object.__setattr__(model, "generation_config", _read_generation_settings(pretrained_model_name_or_path))
        sd = _read_checkpoint(pretrained_model_name_or_path)
        if sd is not None:
            model.load_reference_state_dict(sd, strict=False)
            enc = model.model.encoder
            if hasattr(enc, "embed_tokens_ner") and "model.encoder.embed_tokens_ner.weight" not in sd:
                # a plain BART checkpoint: the NER tables start as copies of the token / position tables
                n = min(50265, enc.embed_tokens.weight.shape[0])
                with torch.no_grad():
                    enc.embed_tokens_ner.weight[:n] = enc.embed_tokens.weight[:n]
                    enc.embed_positions_ner.weight.copy_(enc.embed_positions.weight)
                model.store.refresh_shadow()
        return model

    def state_dict(self, *a, **k):
        sd = super().state_dict(*a, **k)
        return sd

    def save_pretrained(self, save_directory: str, safe_serialization: bool = True):
        """Write `config.json` + `model.safetensors` (or `pytorch_model.bin`) with the REFERENCE parameter names, so the
        checkpoint loads into the reference classes (`load_state_dict`) and back through `from_pretrained` (SURVEY §8f,
        replaces the whole-module pickles of TRAIN:467 / INFER:1087).  Entries that alias `model.shared.weight` are
        written once."""
        os.makedirs(save_directory, exist_ok=True)
        sd = {}
        seen = {}
        for k, v in self.state_dict().items():
            key = (v.data_ptr(), tuple(v.shape))
            if key in seen:
                continue  # tied alias (model.encoder/decoder.embed_tokens.weight)
            seen[key] = k
            sd[k] = v.detach().to("cpu").contiguous().clone()
        cfg = self.config
        if hasattr(cfg, "save_pretrained"):
            cfg.save_pretrained(save_directory)
        else:  # pragma: no cover
            import json
            with open(os.path.join(save_directory, "config.json"), "w") as f:
                json.dump(dict(cfg), f)
        if safe_serialization:
            from safetensors.torch import save_file
            save_file(sd, os.path.join(save_directory, "model.safetensors"), metadata={"format": "pt"})
        else:
            torch.save(sd, os.path.join(save_directory, "pytorch_model.bin"))
        return save_directory

    # whole-module pickling (torch.save(model) at TRAIN:467): rebuild the flat store on load
    def __reduce__(self):
        sd = {k: v.detach().cpu() for k, v in self.state_dict().items()}
        cfg = self.config.to_dict() if hasattr(self.config, "to_dict") else dict(self.config)
        return (_rebuild, (type(self), cfg, self._ctor, sd, getattr(self, "generation_config", None)))

    # ------------------------------------------------------------------ embeddings
    def resize_token_embeddings(self, new_num_tokens: int):
        """MFULL:1906-1918: grow `shared`, `lm_head` and `final_logits_bias` (new rows ~ N(0, 0.02))."""
        old = self.cfg.vocab
        if new_num_tokens == old:
            return self.model.shared
        sd = {k: v.detach().clone() for k, v in self.state_dict().items()}
        cfgd = self.config
        if hasattr(cfgd, "vocab_size"):
            cfgd.vocab_size = new_num_tokens
        else:
            cfgd["vocab_size"] = new_num_tokens
        fresh = type(self)(cfgd, enc_fusion_layer=None, clip_model=self.clip_model, device=self.store.device, **self._ctor)
        n = min(old, new_num_tokens)
        fsd = fresh.state_dict()
        with torch.no_grad():
            for k, v in sd.items():
                tgt = fsd[k]
                if v.shape == tgt.shape:
                    tgt.copy_(v)
                elif k == "final_logits_bias":
                    tgt[:, :n].copy_(v[:, :n])
                else:  # [vocab, d] tables: shared / lm_head
                    tgt[:n].copy_(v[:n])
        fresh.store.refresh_shadow()
        self.__dict__.pop("_generators", None)  # captured decode graphs point at the old store / vocabulary
        self.__dict__.update(fresh.__dict__)
        return self.model.shared

    def get_output_embeddings(self):
        return self.lm_head

    # ------------------------------------------------------------------ generation
    @torch.no_grad()
    def generate(self, input_ids=None, attention_mask=None, num_beams: int = 1, max_length: int = 20,
                 length_penalty: float = 1.0, image_features=None, face_features=None, face_mask=None, name_ids=None,
                 name_mask=None, add_ner_ffn=True, **kwargs):
        """Greedy (num_beams=1) or beam search, transformers-5.5 semantics (see vacnic_b200.generation).  Arguments and
        checkpoint generation settings that would alter the search (no_repeat_ngram_size, early_stopping,
        forced_bos_token_id, sampling, ...) raise NotImplementedError instead of being ignored."""
        from . import generation
        if not add_ner_ffn:
            raise ValueError("add_ner_ffn=False is broken in the reference (MFULL:666 vs :1296) and is not provided")
        passthrough = {k: kwargs.pop(k) for k in ("eos_token_id", "forced_eos_token_id", "pad_token_id", "bos_token_id",
                                                  "decoder_start_token_id") if k in kwargs}
        check_generation_request(self.cfg, self.config, getattr(self, "generation_config", None), {**kwargs, **passthrough})
        return generation.generate(self, input_ids=input_ids, attention_mask=attention_mask, num_beams=num_beams,
                                   max_length=max_length, length_penalty=length_penalty, image_features=image_features,
                                   face_features=face_features, face_mask=face_mask, name_ids=name_ids, name_mask=name_mask)

    def prepare_inputs_for_generation(self, decoder_input_ids, past=None, attention_mask=None, encoder_outputs=None, **kw):
        """MFULL:2023-2061 (kept for API compatibility; `generate` above does not go through it)."""
        if past is not None:
            decoder_input_ids = decoder_input_ids[:, -1:]
        out = dict(input_ids=None, encoder_outputs=encoder_outputs, past_key_values=past, decoder_input_ids=decoder_input_ids,
                   attention_mask=attention_mask)
        out.update({k: kw.get(k) for k in ("image_features", "name_ids", "name_mask", "add_ner_ffn", "face_features", "face_mask")})
        return out

    @staticmethod
    def _reorder_cache(past, beam_idx):
        """MFULL:2066-2074: only the self-attention K/V follow the beams; cross K/V are shared."""
        return tuple(tuple(t.index_select(0, beam_idx) for t in layer[:2]) + tuple(layer[2:]) for layer in past)


def _read_generation_settings(path: str) -> dict:
    """Generation settings a checkpoint directory carries (`generation_config.json`, and the legacy generation keys of
    `config.json` that transformers 4.18 -- the reference's pin, vacnic.yml:187 -- applies inside generate()), as a raw
    dict: `generate()` refuses the ones the device-side search does not implement instead of silently dropping them."""
    import json
    out = {}
    for fn in ("config.json", "generation_config.json"):
        f = os.path.join(path, fn)
        if os.path.isfile(f):
            with open(f) as fh:
                raw = json.load(fh)
            out.update({k: raw[k] for k in GEN_IMPLEMENTED if k in raw})
    return out


def _rebuild(cls, cfg, ctor, sd, gen_cfg=None):
    try:
        from transformers import BartConfig
        config = BartConfig(**cfg)
    except ImportError:  # pragma: no cover
        config = cfg
    m = cls(config, **ctor)
    m.load_reference_state_dict(sd, strict=False)
    if gen_cfg is not None:
        object.__setattr__(m, "generation_config", gen_cfg)
    return m


def _read_checkpoint(path: str) -> Optional[dict]:
    if not os.path.isdir(path):
        raise OSError(f"{path}: only local checkpoint directories can be loaded (no network access)")
    st = glob.glob(os.path.join(path, "*.safetensors"))
    if st:
        from safetensors.torch import load_file
        sd = {}
        for f in st:
            sd.update(load_file(f))
    else:
        bins = glob.glob(os.path.join(path, "pytorch_model*.bin"))
        if not bins:
            return None
        sd = {}
        for f in bins:
            sd.update(torch.load(f, map_location="cpu", weights_only=True))
    if "model.shared.weight" in sd:  # tied entries that some checkpoints omit
        sd.setdefault("lm_head.weight", sd["model.shared.weight"])
        sd.setdefault("model.encoder.embed_tokens.weight", sd["model.shared.weight"])
        sd.setdefault("model.decoder.embed_tokens.weight", sd["model.shared.weight"])
    return sd


class BartForMultiModalGenerationFull(_DropInBase):
    """src.models.modeling_mmbart_..._face_name_ids_crossattn.BartForMultiModalGeneration"""
    ONLY_IMAGE_FILE = False


class BartForMultiModalGenerationVis(_DropInBase):
    """src.models.modeling_mmbart_..._enc_self_crossattn.BartForMultiModalGeneration (only-visual-prompt)"""
    ONLY_IMAGE_FILE = True

    def __init__(self, config, enc_fusion_layer=None, dim_common=256, img_size=2048, prompt_mlp_type="clipcap",
                 map_size=(192, 256, 64, 16), prompt_size=10, clip_model=None, freeze_clip=False, device=None, seed=0, **ignored):
        super().__init__(config, enc_fusion_layer=enc_fusion_layer, dim_common=dim_common, img_size=img_size,
                         prompt_mlp_type=prompt_mlp_type, map_size=map_size, prompt_size=prompt_size, clip_model=clip_model,
                         freeze_clip=freeze_clip, only_image=True, device=device, seed=seed)
