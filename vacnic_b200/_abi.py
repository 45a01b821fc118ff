"""argtypes/restype table for the C ABI (mirrors include/vacnic_b200.h one to one)."""
import ctypes as C

P = C.c_void_p
I32 = C.c_int32
I64 = C.c_int64
F32 = C.c_float

# name -> argtypes (restype is int for all of these)
SIGNATURES = {
    "vacnic_gemm": [P, P],
}


def declare(L):
    L.vacnic_version.restype = C.c_int
    for name, args in SIGNATURES.items():
        fn = getattr(L, name)  # AttributeError here = header/library mismatch: fail loudly
        fn.argtypes = args
        fn.restype = C.c_int
