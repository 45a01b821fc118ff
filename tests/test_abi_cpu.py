"""CPU: the C-ABI shared library loads and exports every symbol include/vacnic_b200.h declares, the
ctypes table mirrors the header, argument validation fails with VACNIC_E* codes (no compute calls: no
GPU here), and the product path refuses to run without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest
import torch

from vacnic_b200 import _abi, lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "vacnic_b200.h")


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vacnic_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def L():
    if not os.path.exists(lib.LIB_PATH):
        lib.build()
    return lib.lib()


def test_header_declares_functions():
    fns = header_functions()
    assert len(fns) >= 25 and "vacnic_gemm" in fns


def test_library_exports_every_declared_symbol(L):
    missing = [f for f in header_functions() if not hasattr(L, f)]
    assert not missing, missing


def test_ctypes_table_matches_header(L):
    declared = set(header_functions())
    table = set(_abi.SIGNATURES) | set(_abi.OTHER)
    assert declared == table, sorted(declared ^ table)
    # argument counts: count the commas of each prototype in the header
    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    for name, args in _abi.SIGNATURES.items():
        m = re.search(r"\b" + name + r"\s*\(([^;]*?)\)\s*;", src, flags=re.S)
        assert m, name
        n = 0 if m.group(1).strip() in ("", "void") else m.group(1).count(",") + 1
        assert n == len(args), (name, n, len(args))


def test_version_and_error_string(L):
    assert L.vacnic_version() >= 100
    assert isinstance(L.vacnic_last_error(), bytes)


def test_invalid_arguments_return_error_codes(L):
    rc = L.vacnic_gemm(None, None)
    assert rc == -1 and b"null" in L.vacnic_last_error()
    d = lib.GemmDesc()
    rc = L.vacnic_gemm(C.byref(d), None)
    assert rc == -1 and b"positive" in L.vacnic_last_error()
    rc = L.vacnic_ce_fwd(None, None, None, None, None, 4, 8, 8, 1, None)
    assert rc == -1


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    from vacnic_b200 import kernels as K
    a = torch.zeros(8, 8, dtype=torch.bfloat16)
    with pytest.raises(lib.VacnicError):
        K.gemm(a, a)
    from vacnic_b200 import spec
    from vacnic_b200.dropin import BartForMultiModalGenerationFull
    with pytest.raises(RuntimeError):
        BartForMultiModalGenerationFull(dict(d_model=768, encoder_layers=1, decoder_layers=1), dim_common=768)
    # a GEMM on a box without an sm_100 device reports VACNIC_EDEVICE instead of computing on the host
    L = lib.lib()
    d = lib.GemmDesc()
    d.M = d.N = d.K = 64
    d.batch0 = d.batch1 = 1
    d.a = d.b = d.c = 4096  # never dereferenced: the device check comes first
    d.lda = d.ldb = d.ldc = 64
    assert L.vacnic_gemm(C.byref(d), None) == -2
    assert spec.bart_large().d_model == 1024


def test_every_pdl_launched_kernel_waits_before_touching_memory():
    """Static guard for programmatic dependent launch (csrc/common.h): a kernel launched through launch_pdl may start
    while its predecessor is still running, so its body must execute pdl_sync() (griddepcontrol.wait + trigger); a
    kernel that is launched with the attribute but never waits would race silently."""
    import glob
    import os
    import re
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "vacnic_b200", "csrc")
    src = {p: open(p).read() for p in glob.glob(os.path.join(root, "*.cu"))}
    launched = set()
    for text in src.values():
        for m in re.finditer(r"launch_pdl\(\s*([A-Za-z_0-9]+)", text):
            launched.add(m.group(1))
    launched.discard("kernel")  # the helper's own parameter name
    launched.discard("kern")    # GEMM launchers pass a function pointer variable: checked by name below
    launched |= {"gemm_sm100_kernel", "gemm2_sm100_kernel"}
    assert len(launched) >= 15, launched
    for name in sorted(launched):
        body = None
        for text in src.values():
            m = re.search(r"__global__[^;{]*?\b" + re.escape(name) + r"\s*\([^;{]*?\)\s*\{", text, re.S)
            if m:
                # body = up to the next kernel definition (or end of file)
                nxt = text.find("__global__", m.end())
                body = text[m.end(): nxt if nxt != -1 else len(text)]
                break
        assert body is not None, f"definition of {name} not found"
        assert "pdl_sync()" in body, f"{name} is launched with the PDL attribute but never calls pdl_sync()"
        # nothing that dereferences global memory may precede the wait: the first statement region up to pdl_sync()
        head = body[: body.index("pdl_sync()")]
        assert "__ldg" not in head and "tma_load" not in head and "atomicAdd" not in head, name
