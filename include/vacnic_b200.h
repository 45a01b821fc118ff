/*
 * vacnic_b200.h — C ABI of libvacnic_b200.so (sm_100a only).
 *
 * The reference (tingyu215/VACNIC) is 100 % Python and has no FFI of its own; every device
 * operation it performs is a PyTorch-dispatched library call.  Each entry point below therefore
 * cites the reference *call site* whose arithmetic it replaces (file:line under /root/reference,
 * MFULL = src/models/modeling_mmbart_clip_inside_vis_clipcap_ent_type_final_fix_len_enc_self_
 * face_name_ids_crossattn.py, TRAIN = train_mmbart_enc_self_face_name_ids_retrieve_crossattn_
 * bart_guide_match.py).  INTEGRATION.md shows the ctypes binding a maintainer would add.
 *
 * Conventions (all entry points):
 *   - return 0 on success, a negative VACNIC_E* code otherwise; vacnic_last_error() returns a
 *     thread-local human-readable message for the last failure on the calling thread;
 *   - every pointer is a DEVICE pointer owned by the caller; the library never allocates device
 *     memory and never synchronises: work is enqueued on `stream` (a cudaStream_t passed as void*);
 *   - activations / weight shadows are bf16, row-major; LayerNorm parameters, biases, statistics,
 *     loss scalars and gradient accumulators are fp32; token ids are int64; lengths int32;
 *   - there is no CPU fallback: calling without a usable sm_100 device returns VACNIC_EDEVICE.
 */
#ifndef VACNIC_B200_H_
#define VACNIC_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VACNIC_OK 0
#define VACNIC_EINVAL (-1)  /* bad argument (shape, alignment, null pointer) */
#define VACNIC_EDEVICE (-2) /* no sm_100 device / driver entry point missing */
#define VACNIC_ECUDA (-3)   /* a CUDA runtime / driver call failed         */

#define VACNIC_DT_BF16 0
#define VACNIC_DT_F32 1

#define VACNIC_ACT_NONE 0
#define VACNIC_ACT_GELU 1 /* exact erf GELU: ACT2FN["gelu"], MFULL:579 */
#define VACNIC_ACT_TANH 2 /* MLPClipCap, MFULL:111-123 */

#define VACNIC_MASK_NONE 0   /* all-zero additive mask (prefix cross-attention, MFULL:1282-1296) */
#define VACNIC_MASK_KEYPAD 1 /* per-batch key mask bytes, _expand_mask MFULL:387-398 */
#define VACNIC_MASK_CAUSAL 2 /* _make_causal_mask MFULL:373-385 (+ optional key mask) */

const char* vacnic_last_error(void);
int vacnic_version(void);
/* Number of kernels this library has launched in the calling process (bench.py gpu_launches). */
int64_t vacnic_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * Batched bf16 GEMM on tcgen05 tensor cores with TMEM accumulators and TMA-fed operands.
 * Replaces every nn.Linear / torch.bmm on the path (MFULL:465-563 projections and bmm,
 * MFULL:648-650,659-661,683-685,739-741 FFNs, MFULL:1274-1278 prefix MLP, MFULL:1997 LM head)
 * and their autograd dgrad / wgrad counterparts.
 *
 * For every batch (b0, b1):   acc[m,n] = sum_k A(m,k) * B(n,k)
 *   A(m,k) = a[b0*a_sb0 + b1*a_sb1 + (a_mn_major ? k*lda + m : m*lda + k)]
 *   B(n,k) = b[b0*b_sb0 + b1*b_sb1 + (b_mn_major ? k*ldb + n : n*ldb + k)]
 * epilogue, in this order:
 *   v = acc;  if (bias) v += bias[n];  v *= alpha;  if (aux_out) aux_out[m,n] = bf16(v);
 *   v = act(v);  if (dact) v *= act'(aux_in[m,n])   (GELU': aux_in = pre-activation,
 *                                                    TANH': aux_in = activation output);
 *   if (accumulate) v += C[m,n];   C[m,n] = (c_dtype) v
 * C / aux_out / aux_in share ldc, c_sb0, c_sb1 (element strides).
 * Requirements: a, b 16-byte aligned, lda/ldb/batch strides multiples of 8 elements.
 * ------------------------------------------------------------------------------------------ */
typedef struct vacnic_gemm_desc {
  int32_t M, N, K;
  int32_t batch0, batch1;
  const void* a;
  int64_t lda, a_sb0, a_sb1;
  const void* b;
  int64_t ldb, b_sb0, b_sb1;
  void* c;
  int64_t ldc, c_sb0, c_sb1;
  const float* bias;
  void* aux_out;
  const void* aux_in;
  float alpha;
  int32_t a_mn_major;
  int32_t b_mn_major;
  int32_t c_dtype;
  int32_t act;
  int32_t dact;
  int32_t accumulate;
  int32_t tile_n; /* 0 = choose; else 64 / 128 / 256 */
} vacnic_gemm_desc;

int vacnic_gemm(const vacnic_gemm_desc* d, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VACNIC_B200_H_ */
