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
#define VACNIC_ACT_QUICKGELU 3 /* x * sigmoid(1.702 x): MLP of the frozen CLIP ViT image tower (clip.model.QuickGELU), forward only */

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
 * C / aux_out / aux_in share ldc, c_sb0, c_sb1, c_chunk_stride (element strides).
 * Requirements: a, b 16-byte aligned, lda/ldb/batch strides multiples of 8 elements.  A batch stride of 0
 * broadcasts that operand over the batch dimension (one weight matrix for every batch).
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
  /* != 0: output column n is stored at (n / 64) * c_chunk_stride + (n % 64) instead of n, i.e. every
   * 64-column group (one attention head) becomes its own [rows][64] block with row stride ldc: the
   * head-major layout of the decode-time cross-attention K/V cache. */
  int64_t c_chunk_stride;
  /* Optional split-K scratch (device memory, zero-initialised ONCE by the caller, >= 64 KiB + partial tiles; reused
   * by every call on the same stream): lets problems with few output tiles and a long K spread over idle SMs.
   * The reduction order is fixed, so results are deterministic.  Null = never split. */
  void* workspace;
  int64_t workspace_bytes;
  int32_t split_k_min_blocks; /* split only when K spans at least this many 64-wide blocks (0 = 4) */
} vacnic_gemm_desc;

int vacnic_gemm(const vacnic_gemm_desc* d, void* stream);

/* ------------------------------------------------------------------------------------------
 * LayerNorm family (bf16 activations, fp32 parameters / statistics, eps = 1e-5: nn.LayerNorm
 * defaults at MFULL:577,583,...).  d must be a multiple of 256 (768 / 1024).
 * Dropout (config.dropout, MFULL:705,721,742 ...) is a counter-based Bernoulli mask keyed by
 * (*rng_state, salt, element index); the backward call regenerates it from the same triple.
 * ------------------------------------------------------------------------------------------ */
/* y = LN(res + dropout(x)).  Output row r lands at y + (r / rows_per_group) * y_group_stride +
 * (r % rows_per_group) * d, so a result can be written straight into a slice of the prefix/NER
 * concat buffer (torch.cat at MFULL:691).  rows_per_group = 0 means contiguous output.
 * res32 (optional, fp32 [rows, d]): the residual in fp32 (used instead of `res`); y32 (optional, fp32 [rows, d]): fp32 copy
 * of the output.  Together they carry the residual stream in fp32 from block to block, like the reference under
 * torch.autocast (LayerNorm runs and returns fp32 there) -- the GEMMs still consume the bf16 `y`.
 * sum32 (optional, fp32 [rows, d]): res + dropout(x) BEFORE normalisation -- the residual stream of a pre-LN block (the CLIP
 * ViT: x = x + attn(ln_1(x))); x may be null when res32 is given (plain LayerNorm of the fp32 stream). */
int vacnic_add_layernorm_fwd(const void* x, const void* res, const float* gamma, const float* beta, void* y,
                             float* mean, float* rstd, int64_t rows, int32_t d, int64_t rows_per_group,
                             int64_t y_group_stride, float eps, float p_drop, const uint64_t* rng_state,
                             uint32_t salt, const float* res32, float* y32, float* sum32, void* stream);
/* Gradients of the above: dsum = d(res + dropout(x)) (written or accumulated), dx = dropout-masked
 * dsum (skipped when dx == dsum or null), dgamma/dbeta/dbias are accumulated atomically (fp32);
 * dbias is the bias gradient of the linear layer that produced x (column sums of dx). */
int vacnic_add_layernorm_bwd(const void* dy, const void* x, const void* res, const float* gamma, const float* mean,
                             const float* rstd, void* dsum, void* dx, float* dgamma, float* dbeta, float* dbias,
                             int64_t rows, int32_t d, int64_t rows_per_group, int64_t dy_group_stride, float p_drop,
                             const uint64_t* rng_state, uint32_t salt, int32_t accumulate_dsum, void* stream);
/* y = dropout(LN(tok[ids] + pos[(r % seq_len) + pos_offset])): MFULL:1243-1249 (article),
 * 1254-1260 (names, embed_tokens_ner), 1555-1563 (decoder, pos_offset = 2 + cached length).
 * y32 (optional): fp32 copy of the output (start of the fp32 residual stream).  pos_ids (optional, int32 [rows]): the
 * position of every row given explicitly instead of r % seq_len -- rows of several articles packed back to back
 * (varlen batches: the collate's padding, DNYT:957-972, never reaches the device). */
int vacnic_embed_ln_fwd(const int64_t* ids, const void* tok, const void* pos, const float* gamma, const float* beta,
                        void* y, float* mean, float* rstd, int64_t rows, int32_t seq_len, int32_t pos_offset,
                        int32_t d, float eps, float p_drop, const uint64_t* rng_state, uint32_t salt, float* y32,
                        const int32_t* pos_ids, void* stream);
/* Scatter-add gradients into the fp32 embedding tables; rows with ids == pad_id get no token
 * gradient (nn.Embedding padding_idx, MFULL:1115,1150). */
int vacnic_embed_ln_bwd(const void* dy, const int64_t* ids, const void* tok, const void* pos, const float* gamma,
                        const float* mean, const float* rstd, float* dtok, float* dpos, float* dgamma, float* dbeta,
                        int64_t rows, int32_t seq_len, int32_t pos_offset, int32_t d, int32_t pad_id, float p_drop,
                        const uint64_t* rng_state, uint32_t salt, const int32_t* pos_ids, void* stream);
/* get_embedding_ner (TRAIN:112-133): out[span] = mean_t LN(tok[ids[span,t]] + pos[t + 2]), fp32. */
int vacnic_names_embed(const int64_t* ids, const void* tok, const void* pos, const float* gamma, const float* beta,
                       float* out, int64_t spans, int32_t len, int32_t d, float eps, void* stream);

/* ------------------------------------------------------------------------------------------
 * Masked softmax for the unfused attention path (MFULL:509-548).  scores fp32 / probs bf16,
 * both [B, H, Sq, ld] with ld >= Sk (pad columns are written as zeros).  key_mask: uint8 [B, Sk],
 * 1 = attend (the reference's attention_mask, _expand_mask MFULL:387-398) or null; causal != 0
 * masks key j > query i + past (_make_causal_mask MFULL:373-385).
 * ------------------------------------------------------------------------------------------ */
int vacnic_softmax_fwd(const float* scores, void* probs, const uint8_t* key_mask, int32_t B, int32_t H, int32_t Sq,
                       int32_t Sk, int32_t ld, int32_t causal, int32_t past, void* stream);
/* dscores = probs * (dprobs - rowsum(probs * dprobs)), bf16 out. */
int vacnic_softmax_bwd(const void* probs, const float* dprobs, void* dscores, int64_t rows, int32_t Sk, int32_t ld,
                       void* stream);

/* ------------------------------------------------------------------------------------------
 * Elementwise / reduction helpers.
 * ------------------------------------------------------------------------------------------ */
int vacnic_colsum(const void* x, float* out, int64_t rows, int32_t n, int64_t ld, void* stream); /* out[c] += sum_r x[r,c] : nn.Linear bias gradients */
int vacnic_cast_f32_bf16(const float* src, void* dst, int64_t n, void* stream);
int vacnic_concat_rows(const void* a, const void* b, void* out, int32_t B, int64_t rows_a, int64_t rows_b, int32_t d,
                       void* stream); /* out[B, ra+rb, d] = cat(a[B,ra,d], b[B,rb,d]) : MFULL:668, 691 */
/* dst[r, 0..ld_dst) = bf16({src[r, 0..n), 0 ...}), src rows ld_src floats apart: fp32 gradients coming back from torch
 * ops (script-level losses, TRAIN:287) -> TMA-readable bf16 operands. */
int vacnic_cast_rows_f32_bf16(const float* src, void* dst, int64_t rows, int32_t n, int64_t ld_src, int32_t ld_dst, void* stream);
/* dst[r, 0..ld_dst) = {src[r, 0..n), 0 ...} (bf16): re-pitch rows for TMA (16-byte row pitch). */
int vacnic_pad_rows(const void* src, void* dst, int64_t rows, int32_t n, int32_t ld_dst, void* stream);
/* dst[j] (+)= sum_p src[p*len + j] (fp32): reduces split partial weight gradients. */
int vacnic_sum_partials(const float* src, float* dst, int32_t parts, int64_t len, int32_t accumulate, void* stream);
int vacnic_add_bf16(const void* a, const void* b, const void* c, void* out, int64_t n, void* stream); /* out = a + b (+ c) */
/* Packed (varlen) rows -> the collate's right-padded layout (DNYT:957-972): dst[b, t, :] = src[start[b] + t, :] for
 * t < len[b], zeros otherwise; src bf16 [rows, d], dst bf16 [batch, L, d], start / len int32 [batch]. */
int vacnic_unpack_rows(const void* src, const int32_t* start, const int32_t* len, void* dst, int32_t batch, int32_t L, int32_t d,
                       void* stream);
/* Fused AdamW over a flat parameter buffer (TRAIN:91-107,371-373).  hyper (device, fp32[8]) =
 * {lr, beta1, beta2, eps, weight_decay, 1-beta1^t, 1-beta2^t, grad_scale}; also refreshes the bf16
 * compute shadow p16 (may be null). */
int vacnic_adamw(float* p, const float* g, float* m, float* v, void* p16, int64_t n, const float* hyper, void* stream);
/* Device-side optimizer step counter + schedule (get_linear_schedule_with_warmup TRAIN:102, scheduler.step() after
 * optimizer.step() TRAIN:371-374; Adam bias corrections of torch.optim.AdamW TRAIN:95-101): *step += 1 (t), then hyper =
 * {base_lr * multiplier(t-1), beta1, beta2, eps, weight_decay, 1-beta1^t, 1-beta2^t, grad_scale}.  Runs inside the captured
 * step graph, so no host buffer is read asynchronously. */
int vacnic_optim_schedule(int64_t* step, float* hyper, double base_lr, double beta1, double beta2, float eps, float weight_decay,
                          int64_t warmup_steps, int64_t total_steps, float grad_scale, void* stream);
/* Data-parallel optimizer step over NVLink peer memory: reduce-scatter + AdamW + all-gather as ONE kernel (replaces the DDP
 * gradient all-reduce TRAIN:86-87 followed by optimizer.step() TRAIN:371).  grad_ptrs[r] / shadow_ptrs[r] (host arrays of
 * `world` device addresses) are rank r's flat fp32 gradient buffer and bf16 compute shadow, mapped into this process
 * (symmetric memory).  This rank owns elements [begin, begin+count) (multiples of 8): it sums that shard of every rank's
 * gradient in rank order (fp32), applies vacnic_adamw's arithmetic to its local master / m / v, and stores the new bf16
 * weights into every rank's shadow.  grad_mc / shadow_mc: NVSwitch multicast addresses of the same buffers (0 = plain peer
 * loads / stores; non-zero = multimem.ld_reduce / multimem.st).  Cross-rank ordering (all ranks finished writing the
 * bucket; all ranks finished the step) is the caller's: a symmetric-memory barrier on the same stream.  max_blocks caps
 * the grid (0 = 4 x SM count) so the kernel can run beside the backward pass. */
int vacnic_dp_adamw_shard(const uint64_t* grad_ptrs, const uint64_t* shadow_ptrs, uint64_t grad_mc, uint64_t shadow_mc,
                          int32_t world, int32_t rank, float* master, float* m, float* v, const float* hyper, int64_t begin,
                          int64_t count, int32_t max_blocks, void* stream);
/* clip_grad_norm_ (TRAIN:365-366) folded into the optimizer: *scale_out = base_scale * min(1, max_norm / (|base_scale| *
 * ||g||_2 + 1e-6)); write it to hyper[7] of vacnic_adamw.  scratch: VACNIC_CLIP_SCRATCH_FLOATS device floats (per-block
 * partial sums, reduced in a fixed order: the norm is bit-reproducible). */
#define VACNIC_CLIP_SCRATCH_FLOATS 2368
int vacnic_clip_grad_scale(const float* g, int64_t n, float max_norm, float base_scale, float* scratch, float* scale_out,
                           float* norm_out, void* stream);
int vacnic_rng_advance(uint64_t* state, void* stream);
/* x = dropout(x) in place (bf16, n elements contiguous): config.activation_dropout after the FFN activation (MFULL:649, 660,
 * 684, 740, 874).  Counter-based mask keyed by (*rng_state, salt, element index): calling it again on the gradient with
 * the same key applies the same mask (backward pass). */
int vacnic_dropout_inplace(void* x, int64_t n, float p_drop, const uint64_t* rng_state, uint32_t salt, void* stream);


/* ------------------------------------------------------------------------------------------
 * Frozen CLIP ViT image tower (extract_clip_img_feat, TRAIN:220-240; OpenAI CLIP VisionTransformer): the step
 * immediately upstream of the ClipCap prefix MLP.  Forward only.
 * ------------------------------------------------------------------------------------------ */
/* conv1 (kernel = stride = patch, no bias) as a GEMM: out bf16 [batch * (H/p) * (W/p), C*p*p] = non-overlapping patches of
 * images fp32 [batch, C, H, W], column order (c, ky, kx) = conv1.weight.view(width, -1). */
int vacnic_vit_patchify(const float* images, void* out, int32_t batch, int32_t channels, int32_t height, int32_t width,
                        int32_t patch, void* stream);
/* y32[b, 0] = LN(class_embedding + pos[0]); y32[b, i] = LN(tok[b, i-1] + pos[i]) (ln_pre), fp32 [batch, tokens, d];
 * tok bf16 [batch, tokens-1, d] = the patch projections. */
int vacnic_vit_embed_ln(const void* tok, const float* cls, const float* pos, const float* gamma, const float* beta, float* y32,
                        int64_t batch, int32_t tokens, int32_t d, float eps, void* stream);

/* ------------------------------------------------------------------------------------------
 * Losses (script-level code in the reference, TRAIN:284-363).
 * ------------------------------------------------------------------------------------------ */
/* CrossEntropyLoss(ignore_index) over fp32 logits [rows, ld] (TRAIN:287,816).  out[0] = mean loss,
 * out[1] = number of counted rows; lse / row_loss are [rows] scratch kept for the backward. */
int vacnic_ce_fwd(const float* logits, const int64_t* targets, float* lse, float* row_loss, float* out, int64_t rows,
                  int32_t V, int64_t ld, int64_t ignore_index, void* stream);
/* dlogits (bf16 [rows, ld]) = (softmax - onehot) * coef * gscale[0] / out[1]. */
int vacnic_ce_bwd(const float* logits, const float* lse, const int64_t* targets, const float* stats,
                  const float* gscale, float coef, void* dlogits, int64_t rows, int32_t V, int64_t ld,
                  int64_t ignore_index, void* stream);
/* CoLaM (TRAIN:292-309): masked mean pool over non-pad label positions (nan -> 1.0), L2 normalise,
 * diagonal cosine, HingeEmbeddingLoss(margin) with target -1.  stats is fp32 [B, 8]. */
int vacnic_colam_fwd(const void* h, const void* h_guide, const int64_t* tgt_ids, float* pooled_a, float* pooled_b,
                     float* stats, float* loss, int32_t B, int32_t T, int32_t d, int64_t pad_id, float margin,
                     void* stream);
int vacnic_colam_bwd(const float* pooled_a, const float* pooled_b, const float* stats, const int64_t* tgt_ids,
                     const float* gscale, float coef, void* dh, int32_t B, int32_t T, int32_t d, int64_t pad_id,
                     int32_t accumulate, void* stream);
/* SECLA BatchSoftmax (TRAIN:631-660): names fp32 [B,N,d] (no grad), face bf16 [B,F,d]. */
int64_t vacnic_secla_workspace_bytes(int32_t B, int32_t N, int32_t F);
int vacnic_secla_fwd(const float* names, const void* face, float* workspace, float* loss, int32_t B, int32_t N,
                     int32_t F, int32_t d, void* stream);
int vacnic_secla_bwd(const float* workspace, const float* names, const float* gscale, float coef, void* dface,
                     int32_t B, int32_t N, int32_t F, int32_t d, int32_t accumulate, void* stream);

/* ------------------------------------------------------------------------------------------
 * Cached decoding + device-side greedy / beam search (generate() call sites INFER:798,867,
 * TRAIN:513-520; hooks MFULL:2023-2074; cached attention branches MFULL:474-501).  The search
 * algorithm is transformers' `_beam_search` / greedy `_sample` (third-party, restated in
 * oracle/generate.py).  Rows r = caption * beams + beam.  Every kernel reads the current sequence
 * length from the device integer `cur_len`, so one captured CUDA graph serves every step.
 * Ping-pong state buffers hold two copies; step t reads copy (t & 1) and writes copy ((t+1) & 1).
 * ------------------------------------------------------------------------------------------ */
/* y[r] = LN(tok[seq[r][cur_len-1]] + pos[cur_len-1 + pos_offset])  (MFULL:1552-1563, cached).
 * seq: int32 [R][maxT] (pingpong = 0) or [2][R][maxT] (pingpong = 1).  y32 (optional): fp32 copy of y. */
int vacnic_decode_embed_ln(const int32_t* seq, const int32_t* cur_len, const void* tok, const void* pos,
                           const float* gamma, const float* beta, void* y, int32_t R, int32_t maxT, int32_t d,
                           int32_t pos_offset, int32_t pingpong, float eps, float* y32, void* stream);
/* Self-attention of the newest token over the cache (MFULL:490-495,509-563).  qkv bf16 [R][3d] in
 * [k|v|q] order; the new k/v are appended to kcache/vcache [R][maxT][d] at position cur_len-1; position
 * s < cur_len-1 is read from row anc[(cur_len&1)][r][s] (null anc = own row: greedy).  out bf16 [R][d]. */
int vacnic_decode_self_attn(const void* qkv, void* kcache, void* vcache, const int32_t* anc, const int32_t* cur_len,
                            void* out, int32_t R, int32_t H, int32_t head_dim, int32_t maxT, void* stream);
/* Cross-attention of the nq beams of each caption over its L encoder keys (MFULL:474-479): q bf16
 * row (c*nq+i) at q + row*ldq; K(c,h,s,:) = k[c*kv_cs + h*kv_hs + s*ldkv + 0..63], same strides for v;
 * key_mask uint8 [captions][L] (1 = attend) and key_len int32 [captions] (keys >= key_len are skipped;
 * see vacnic_mask_key_len) may be null.  out row stride ldo. */
int vacnic_decode_cross_attn(const void* q, int64_t ldq, const void* k, const void* v, int64_t ldkv, int64_t kv_hs,
                             int64_t kv_cs, const uint8_t* key_mask, const int32_t* key_len, void* out, int64_t ldo,
                             int32_t captions, int32_t nq, int32_t L, int32_t H, int32_t head_dim, void* stream);
int vacnic_mask_key_len(const uint8_t* mask, int32_t* key_len, int32_t B, int32_t L, void* stream);
/* Per row of fp32 logits [rows][ld]: log-softmax statistics and the K best entries, sorted (ties: lower
 * index; -inf entries are never returned: missing slots hold -inf / 0x7fffffff).  top_lp[r][k] = logit -
 * logsumexp(row).  V <= 65536, K <= 16, rows <= 65535.  One 8-CTA cluster per row, the row lives in registers. */
int vacnic_decode_topk(const float* logits, int64_t ld, int32_t rows, int32_t V, int32_t K, float* top_lp,
                       int32_t* top_idx, void* stream);
/* One `_beam_search` iteration (top-2*beams continuations, running beams, finished set with
 * score / len^length_penalty, early-stop heuristic for early_stopping=False, forced EOS at
 * max_len-1).  flags int32 [maxT][2]: flags[t] = {some caption can still improve, some candidate was
 * not stopped}; the host continues while both are set. */
int vacnic_beam_step(const float* top_lp, const int32_t* top_idx, int32_t* run_seq, int32_t* run_anc, float* run_score,
                     int32_t* fin_seq, float* fin_score, int32_t* fin_len, uint8_t* fin_flag, uint8_t* unsat,
                     int32_t* flags, const int32_t* cur_len, int32_t captions, int32_t beams, int32_t maxT,
                     int32_t max_len, int32_t eos, int32_t V, float length_penalty, void* stream);
int vacnic_greedy_step(const int32_t* top_idx, int32_t* seq, uint8_t* unfinished, int32_t* flags, const int32_t* cur_len,
                       int32_t rows, int32_t maxT, int32_t max_len, int32_t eos, int32_t pad, void* stream);
int vacnic_advance_len(int32_t* cur_len, void* stream);

/* ------------------------------------------------------------------------------------------
 * Fused attention core, head_dim 64 (BartAttention.forward MFULL:503-556: bmm(q,k^T), additive mask,
 * softmax, bmm(p,v); the q scaling of MFULL:472 is applied to the fp32 scores).  Scores and
 * probabilities stay on chip (tcgen05 accumulators in TMEM, P in shared memory).
 *   Q(b,h,i,c) = q[b*q_sb + h*q_sh + i*ldq + c]   (same addressing for k, v, dq, dk, dv)
 *   O(b,i,h*64+c) = out[b*o_sb + i*ldo + h*64 + c] (same for dout)
 * key_mask: uint8 [B][Sk], 1 = attend (null = no padding mask); causal != 0 masks key j > query i;
 * masked scores are set to finfo(float32).min like _expand_mask / _make_causal_mask (MFULL:373-398).
 * key_len (int32 [B], optional): keys >= key_len[b] are known to be masked and are skipped.
 * stats: fp32 [B][H][Sq][2] = {row max (log2 domain), 1 / row sum}, written by the forward pass
 * (may be null for inference) and read by the backward pass.
 * ------------------------------------------------------------------------------------------ */
typedef struct vacnic_attn_desc {
  int32_t B, H, Sq, Sk, head_dim, causal;
  const void* q; int64_t ldq, q_sh, q_sb;
  const void* k; int64_t ldk, k_sh, k_sb;
  const void* v; int64_t ldv, v_sh, v_sb;
  void* out; int64_t ldo, o_sb;
  float* stats;
  const uint8_t* key_mask;
  const int32_t* key_len;
  /* backward only */
  const void* dout; int64_t lddo, do_sb;
  void* dq; int64_t lddq, dq_sh, dq_sb;
  void* dk; int64_t lddk, dk_sh, dk_sb;
  void* dv; int64_t lddv, dv_sh, dv_sb;
  float* delta; /* fp32 [B][H][Sq] scratch: rowsum(dO * O) */
  /* Packed (varlen) mode -- active when q_start != NULL.  The collate's padding (DNYT:957-972, TRAIN:255-271) never
   * reaches the device: the rows of all sequences are stored back to back.  Sequence b (0 <= b < B) owns query rows
   * [q_start[b], q_start[b] + q_len[b]) of q / out / dout / dq (each a [total_q rows] x [H*64] matrix: row stride ld*,
   * head stride *_sh, batch strides ignored) and key rows [k_start[b], k_start[b] + k_len[b]) of k / v / dk / dv
   * (total_k rows).  Sq / Sk are then the MAXIMUM q_len / k_len (they size the grid), key_mask / key_len must be NULL,
   * causal compares positions inside the sequence, stats is fp32 [H][total_q][2] and delta fp32 [H][total_q].
   * All four arrays are int32 device pointers of length B. */
  const int32_t* q_start; const int32_t* q_len; const int32_t* k_start; const int32_t* k_len;
  int32_t total_q, total_k;
  /* Attention dropout (config.attention_dropout: `nn.functional.dropout(attn_weights, p=self.dropout)` after the softmax,
   * MFULL:546).  p_drop = 0 disables it.  Counter-based Bernoulli mask keyed by (*rng_state, salt, (sequence, head, query
   * row), key): vacnic_attn_bwd regenerates it from the same values, nothing is stored. */
  float p_drop; const uint64_t* rng_state; uint32_t salt;
} vacnic_attn_desc;
int vacnic_attn_fwd(const vacnic_attn_desc* d, void* stream);
/* dq, dk, dv of the above (dq/dk carry the head_dim^-0.5 factor).  `out` must hold the forward result. */
int vacnic_attn_bwd(const vacnic_attn_desc* d, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VACNIC_B200_H_ */
