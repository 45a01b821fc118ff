"""Reference model classes with GenerationMixin restored (transformers >= 5 dropped it from
PreTrainedModel).  TEST INFRASTRUCTURE ONLY; importable only where /root/reference exists."""
from transformers import GenerationMixin

from .ref_shim import ref_full, ref_vis


class OracleFull(ref_full().BartForMultiModalGeneration, GenerationMixin):
    pass


class OracleVis(ref_vis().BartForMultiModalGeneration, GenerationMixin):
    pass
