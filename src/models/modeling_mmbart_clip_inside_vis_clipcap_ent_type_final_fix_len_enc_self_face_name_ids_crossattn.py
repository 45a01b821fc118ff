"""Drop-in for the reference module of the same path (full VACNIC model: ClipCap prefix, face/name branch,
visually-aware encoder; MFULL).  The training script and the inference file import
`BartForMultiModalGeneration` from here unchanged (TRAIN:601-604, INFER:1038-1045); pickled models resolve
the class by this module path (INFER:1087).  Implementation: vacnic_b200 (sm_100a kernels, no CPU path)."""
from vacnic_b200.dropin import BartForMultiModalGenerationFull as _Impl
from vacnic_b200.modeling import (BartAttention, BartDecoder, BartDecoderLayer, BartEncoder, BartEncoderLayer,  # noqa: F401
                                  BartLearnedPositionalEmbedding, BartModel, shift_tokens_right)


class BartForMultiModalGeneration(_Impl):
    """full VACNIC model: same class name and module path as the reference, so `torch.save(model)` / `torch.load` (TRAIN:467,
    INFER:1087) and `from ... import BartForMultiModalGeneration` resolve here."""
