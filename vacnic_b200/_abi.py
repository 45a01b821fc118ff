"""argtypes/restype table for the C ABI (mirrors include/vacnic_b200.h one to one)."""
import ctypes as C

P = C.c_void_p
I32 = C.c_int32
U32 = C.c_uint32
I64 = C.c_int64
U64 = C.c_uint64
F32 = C.c_float
F64 = C.c_double

# name -> argtypes (restype is int for all of these)
SIGNATURES = {
    "vacnic_gemm": [P, P],
    "vacnic_add_layernorm_fwd": [P, P, P, P, P, P, P, I64, I32, I64, I64, F32, F32, P, U32, P, P, P, P],
    "vacnic_add_layernorm_bwd": [P, P, P, P, P, P, P, P, P, P, P, I64, I32, I64, I64, F32, P, U32, I32, P],
    "vacnic_embed_ln_fwd": [P, P, P, P, P, P, P, P, I64, I32, I32, I32, F32, F32, P, U32, P, P, P],
    "vacnic_embed_ln_bwd": [P, P, P, P, P, P, P, P, P, P, P, I64, I32, I32, I32, I32, F32, P, U32, P, P],
    "vacnic_names_embed": [P, P, P, P, P, P, I64, I32, I32, F32, P],
    "vacnic_softmax_fwd": [P, P, P, I32, I32, I32, I32, I32, I32, I32, P],
    "vacnic_softmax_bwd": [P, P, P, I64, I32, I32, P],
    "vacnic_colsum": [P, P, I64, I32, I64, P],
    "vacnic_cast_f32_bf16": [P, P, I64, P],
    "vacnic_concat_rows": [P, P, P, I32, I64, I64, I32, P],
    "vacnic_add_bf16": [P, P, P, P, I64, P],
    "vacnic_pad_rows": [P, P, I64, I32, I32, P],
    "vacnic_cast_rows_f32_bf16": [P, P, I64, I32, I64, I32, P],
    "vacnic_sum_partials": [P, P, I32, I64, I32, P],
    "vacnic_adamw": [P, P, P, P, P, I64, P, P],
    "vacnic_optim_schedule": [P, P, F64, F64, F64, F32, F32, I64, I64, F32, P],
    "vacnic_dp_adamw_shard": [P, P, U64, U64, I32, I32, P, P, P, P, I64, I64, I32, P],
    "vacnic_vit_patchify": [P, P, I32, I32, I32, I32, I32, P],
    "vacnic_vit_embed_ln": [P, P, P, P, P, P, I64, I32, I32, F32, P],
    "vacnic_unpack_rows": [P, P, P, P, I32, I32, I32, P],
    "vacnic_rng_advance": [P, P],
    "vacnic_dropout_inplace": [P, I64, F32, P, U32, P],
    "vacnic_clip_grad_scale": [P, I64, F32, F32, P, P, P, P],
    "vacnic_ce_fwd": [P, P, P, P, P, I64, I32, I64, I64, P],
    "vacnic_ce_bwd": [P, P, P, P, P, F32, P, I64, I32, I64, I64, P],
    "vacnic_colam_fwd": [P, P, P, P, P, P, P, I32, I32, I32, I64, F32, P],
    "vacnic_colam_bwd": [P, P, P, P, P, F32, P, I32, I32, I32, I64, I32, P],
    "vacnic_secla_fwd": [P, P, P, P, I32, I32, I32, I32, P],
    "vacnic_secla_bwd": [P, P, P, F32, P, I32, I32, I32, I32, I32, P],
    "vacnic_decode_embed_ln": [P, P, P, P, P, P, P, I32, I32, I32, I32, I32, F32, P, P],
    "vacnic_decode_self_attn": [P, P, P, P, P, P, I32, I32, I32, I32, P],
    "vacnic_decode_cross_attn": [P, I64, P, P, I64, I64, I64, P, P, P, I64, I32, I32, I32, I32, I32, P],
    "vacnic_mask_key_len": [P, P, I32, I32, P],
    "vacnic_decode_topk": [P, I64, I32, I32, I32, P, P, P],
    "vacnic_beam_step": [P, P, P, P, P, P, P, P, P, P, P, P, I32, I32, I32, I32, I32, I32, F32, P],
    "vacnic_greedy_step": [P, P, P, P, P, I32, I32, I32, I32, I32, P],
    "vacnic_advance_len": [P, P],
    "vacnic_attn_fwd": [P, P],
    "vacnic_attn_bwd": [P, P],
}
# entry points that do not return a status code
OTHER = {
    "vacnic_last_error": ([], C.c_char_p),
    "vacnic_version": ([], C.c_int),
    "vacnic_launch_count": ([], C.c_int64),
    "vacnic_secla_workspace_bytes": ([I32, I32, I32], C.c_int64),
}


def declare(L):
    for name, args in SIGNATURES.items():
        fn = getattr(L, name)  # AttributeError here = header/library mismatch: fail loudly
        fn.argtypes = args
        fn.restype = C.c_int
    for name, (args, res) in OTHER.items():
        fn = getattr(L, name)
        fn.argtypes = args
        fn.restype = res
