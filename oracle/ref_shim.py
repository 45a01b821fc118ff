"""Import shim for the UNMODIFIED reference model classes under /root/reference.

TEST INFRASTRUCTURE ONLY.  Used in the build container (never on the GPU box, where /root/reference
does not exist) by tests/golden/make_golden.py to generate the committed golden vectors, and by
`bench.py --impl reference` when the reference tree is present.  Three shims:
  1. a stub `clip` module (OpenAI clip is imported at MFULL:47 but never called on the path);
  2. on CPU-only hosts, neutralise the hard-coded `.cuda()` calls (MFULL:698,727,855,1241,1268,1553);
  3. transformers 5.x no longer mixes GenerationMixin into PreTrainedModel, so `generate()` needs
     the subclasses below (defined in a real file because transformers reads the class source).
"""
import importlib
import os
import sys
import types

import torch

REFERENCE_ROOT = os.environ.get("VACNIC_REFERENCE_ROOT", "/root/reference")
MFULL = "src.models.modeling_mmbart_clip_inside_vis_clipcap_ent_type_final_fix_len_enc_self_face_name_ids_crossattn"
MVIS = "src.models.modeling_mmbart_clip_inside_vis_clipcap_ent_type_final_fix_len_enc_self_crossattn"


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "models"))


def _import(name):
    if not available():
        raise ImportError(f"reference tree not found at {REFERENCE_ROOT}")
    sys.modules.setdefault("clip", types.ModuleType("clip"))
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self  # noqa: E731  (CPU-only host)
    # our own drop-in package is also called `src`: import the reference under a private alias
    saved = {k: v for k, v in sys.modules.items() if k == "src" or k.startswith("src.")}
    for k in saved:
        del sys.modules[k]
    # the reference's `src/` has no __init__.py (namespace package) and would lose against the regular `src`
    # package of this repo wherever it sits on sys.path: hide the repo root while importing
    repo_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    saved_path = list(sys.path)
    sys.path[:] = [REFERENCE_ROOT] + [p for p in sys.path if os.path.abspath(p or os.getcwd()) != repo_root]
    try:
        mod = importlib.import_module(name)
    finally:
        sys.path[:] = saved_path
        for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
            sys.modules["_vacnic_ref_" + k] = sys.modules.pop(k)
        sys.modules.update(saved)
    return mod


_cache = {}


def ref_full():
    if "full" not in _cache:
        _cache["full"] = _import(MFULL)
    return _cache["full"]


def ref_vis():
    if "vis" not in _cache:
        _cache["vis"] = _import(MVIS)
    return _cache["vis"]


def oracle_classes():
    """(OracleFull, OracleVis): reference classes + GenerationMixin."""
    from . import _ref_generate_classes as g
    return g.OracleFull, g.OracleVis
