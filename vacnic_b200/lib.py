"""ctypes binding of libvacnic_b200.so (the C ABI declared in include/vacnic_b200.h).

PyTorch is used only for device memory and streams: tensors are passed as raw device pointers.
There is no fallback: if the shared library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvacnic_b200.so")
CSRC_DIR = os.path.join(_HERE, "csrc")

DT_BF16, DT_F32 = 0, 1
ACT_NONE, ACT_GELU, ACT_TANH, ACT_QUICKGELU = 0, 1, 2, 3
MASK_NONE, MASK_KEYPAD, MASK_CAUSAL = 0, 1, 2


class VacnicError(RuntimeError):
    pass


class GemmDesc(C.Structure):
    _fields_ = [
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
        ("batch0", C.c_int32), ("batch1", C.c_int32),
        ("a", C.c_void_p), ("lda", C.c_int64), ("a_sb0", C.c_int64), ("a_sb1", C.c_int64),
        ("b", C.c_void_p), ("ldb", C.c_int64), ("b_sb0", C.c_int64), ("b_sb1", C.c_int64),
        ("c", C.c_void_p), ("ldc", C.c_int64), ("c_sb0", C.c_int64), ("c_sb1", C.c_int64),
        ("bias", C.c_void_p), ("aux_out", C.c_void_p), ("aux_in", C.c_void_p),
        ("alpha", C.c_float),
        ("a_mn_major", C.c_int32), ("b_mn_major", C.c_int32), ("c_dtype", C.c_int32),
        ("act", C.c_int32), ("dact", C.c_int32), ("accumulate", C.c_int32), ("tile_n", C.c_int32),
        ("c_chunk_stride", C.c_int64), ("workspace", C.c_void_p), ("workspace_bytes", C.c_int64), ("split_k_min_blocks", C.c_int32),
    ]


class AttnDesc(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("H", C.c_int32), ("Sq", C.c_int32), ("Sk", C.c_int32), ("head_dim", C.c_int32), ("causal", C.c_int32),
        ("q", C.c_void_p), ("ldq", C.c_int64), ("q_sh", C.c_int64), ("q_sb", C.c_int64),
        ("k", C.c_void_p), ("ldk", C.c_int64), ("k_sh", C.c_int64), ("k_sb", C.c_int64),
        ("v", C.c_void_p), ("ldv", C.c_int64), ("v_sh", C.c_int64), ("v_sb", C.c_int64),
        ("out", C.c_void_p), ("ldo", C.c_int64), ("o_sb", C.c_int64),
        ("stats", C.c_void_p), ("key_mask", C.c_void_p), ("key_len", C.c_void_p),
        ("dout", C.c_void_p), ("lddo", C.c_int64), ("do_sb", C.c_int64),
        ("dq", C.c_void_p), ("lddq", C.c_int64), ("dq_sh", C.c_int64), ("dq_sb", C.c_int64),
        ("dk", C.c_void_p), ("lddk", C.c_int64), ("dk_sh", C.c_int64), ("dk_sb", C.c_int64),
        ("dv", C.c_void_p), ("lddv", C.c_int64), ("dv_sh", C.c_int64), ("dv_sb", C.c_int64),
        ("delta", C.c_void_p),
        ("q_start", C.c_void_p), ("q_len", C.c_void_p), ("k_start", C.c_void_p), ("k_len", C.c_void_p),
        ("total_q", C.c_int32), ("total_k", C.c_int32),
        ("p_drop", C.c_float), ("rng_state", C.c_void_p), ("salt", C.c_uint32),
    ]


def build(verbose: bool = False) -> str:
    """Compile the CUDA sources in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-C", CSRC_DIR, "-j", str(os.cpu_count() or 4)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise VacnicError("building libvacnic_b200.so failed:\n" + res.stdout[-4000:] + res.stderr[-4000:])
    if verbose:
        print(res.stdout[-2000:])
    return LIB_PATH


_lib = None


def lib() -> C.CDLL:
    """Load the shared library (fails loudly when it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise VacnicError(
                f"{LIB_PATH} not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU or PyTorch fallback for the VACNIC hot path)")
        _lib = C.CDLL(LIB_PATH)
        _lib.vacnic_last_error.restype = C.c_char_p
        _lib.vacnic_launch_count.restype = C.c_int64
        _declare(_lib)
    return _lib


def _declare(L: C.CDLL) -> None:
    from . import _abi  # noqa: WPS433  (table of argtypes, kept next to the header)
    _abi.declare(L)


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().vacnic_last_error().decode("utf-8", "replace")
        raise VacnicError(f"{what} failed (code {rc}): {msg}")


def launch_count() -> int:
    return int(lib().vacnic_launch_count())


def stream_ptr() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream


def ptr(t) -> int:
    return 0 if t is None else t.data_ptr()
