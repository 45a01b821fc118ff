"""Drop-in for the reference module of the same path (only-visual-prompt baseline, MVIS; used by
run_train_mmbart_enc_self_onlyvis_retrieve_crossattn.py:538-542).  Implementation: vacnic_b200."""
from vacnic_b200.dropin import BartForMultiModalGenerationVis as BartForMultiModalGeneration  # noqa: F401
from vacnic_b200.modeling import (BartAttention, BartDecoder, BartDecoderLayer, BartEncoder, BartEncoderLayer,  # noqa: F401
                                  BartLearnedPositionalEmbedding, BartModel, shift_tokens_right)

BartForMultiModalGeneration.__module__ = __name__
