"""Drop-in for the reference module of the same path (only-visual-prompt baseline, MVIS; used by
run_train_mmbart_enc_self_onlyvis_retrieve_crossattn.py:538-542).  Implementation: vacnic_b200."""
from vacnic_b200.dropin import BartForMultiModalGenerationVis as _Impl
from vacnic_b200.modeling import (BartAttention, BartDecoder, BartDecoderLayer, BartEncoder, BartEncoderLayer,  # noqa: F401
                                  BartLearnedPositionalEmbedding, BartModel, shift_tokens_right)


class BartForMultiModalGeneration(_Impl):
    """only-visual-prompt model: same class name and module path as the reference, so `torch.save(model)` / `torch.load` (TRAIN:467,
    INFER:1087) and `from ... import BartForMultiModalGeneration` resolve here."""
