"""CPU: host-side logic of the drop-in boundary that needs no device -- `generate()` refuses every argument and every
inherited checkpoint generation setting the device-side search does not implement (it must never return different
captions silently; INFER:798 goes through HF generate(), which honours them)."""
import json

import pytest

from vacnic_b200 import dropin, spec


def _cfgs():
    from transformers import BartConfig
    return spec.bart_large(), BartConfig(output_hidden_states=True)  # TRAIN:743 passes output_hidden_states=True


def test_script_call_shapes_pass():
    cfg, hf = _cfgs()
    dropin.check_generation_request(cfg, hf, None, {})                                  # INFER:798 / 867 pass nothing else
    dropin.check_generation_request(cfg, hf, {}, {"use_cache": True, "early_stopping": False, "do_sample": False})
    dropin.check_generation_request(cfg, hf, None, {"eos_token_id": 2, "pad_token_id": 1, "decoder_start_token_id": 2})


@pytest.mark.parametrize("kw", [{"no_repeat_ngram_size": 3}, {"early_stopping": True}, {"do_sample": True}, {"top_k": 10},
                                {"num_return_sequences": 4}, {"min_length": 5}, {"forced_bos_token_id": 0},
                                {"repetition_penalty": 1.2}, {"bad_words_ids": [[5]]}, {"num_beam_groups": 2},
                                {"eos_token_id": 7}, {"forced_eos_token_id": None}, {"some_future_argument": 1}])
def test_unimplemented_arguments_raise(kw):
    cfg, hf = _cfgs()
    with pytest.raises(NotImplementedError):
        dropin.check_generation_request(cfg, hf, None, kw)


def test_checkpoint_generation_settings_raise_until_overridden(tmp_path):
    """facebook/bart-large ships no_repeat_ngram_size=3, early_stopping=true, forced_bos_token_id=0 in config.json."""
    cfg, hf = _cfgs()
    (tmp_path / "config.json").write_text(json.dumps({"model_type": "bart", "no_repeat_ngram_size": 3, "early_stopping": True,
                                                      "num_beams": 4, "forced_bos_token_id": 0, "forced_eos_token_id": 2}))
    (tmp_path / "generation_config.json").write_text(json.dumps({"min_length": 12}))
    gen = dropin._read_generation_settings(str(tmp_path))
    assert gen == {"no_repeat_ngram_size": 3, "early_stopping": True, "forced_bos_token_id": 0, "min_length": 12}
    with pytest.raises(NotImplementedError, match="no_repeat_ngram_size|early_stopping|forced_bos_token_id|min_length"):
        dropin.check_generation_request(cfg, hf, gen, {})
    # explicit arguments win over the checkpoint, as in HF generate()
    dropin.check_generation_request(cfg, hf, gen, {"no_repeat_ngram_size": 0, "early_stopping": False,
                                                   "forced_bos_token_id": None, "min_length": 0})
