// Cached single-token decoding and device-side greedy / beam search for the VACNIC caption generator
// (generation hooks MFULL:2023-2074, cached BartAttention branches MFULL:474-501, decoder embedding
// MFULL:1552-1563; the search loop itself is transformers' `_beam_search` / greedy `_sample`, the
// algorithm restated in oracle/generate.py and pinned to the real generate()).
//
// Everything here is HBM / latency bound.  All kernels read the current sequence length from a DEVICE
// integer (`cur_len`), so that ONE captured CUDA graph of a decoding step is replayed for every step.
//
// Decode state layout (rows r = caption * beams + beam, R rows in total):
//   self K/V cache  [R][maxT][d] bf16 per layer, written once per (row, position) and NEVER reordered:
//                   the beam search keeps an ancestry table anc[R][maxT] (int32, transformers'
//                   `running_beam_indices`) and attention gathers position s from row anc[r][s];
//   cross K/V       [captions * L][2d] bf16 per layer (k | v), ONE copy per caption shared by its beams
//                   (the reference keeps `beams` identical copies, MFULL:2066-2074);
//   beam state      ping-pong buffers indexed by (cur_len & 1): running sequences, ancestry, finished set.
#include <float.h>

#include "common.h"
#include "ptx.cuh"

namespace vb {

__device__ __forceinline__ void unpack8f(const uint4& u, float (&f)[8]) {
  float2 t;
  t = unpack_bf16x2(u.x); f[0] = t.x; f[1] = t.y;
  t = unpack_bf16x2(u.y); f[2] = t.x; f[3] = t.y;
  t = unpack_bf16x2(u.z); f[4] = t.x; f[5] = t.y;
  t = unpack_bf16x2(u.w); f[6] = t.x; f[7] = t.y;
}

// ------------------------------------------------------------------------------------------------
// x[r] = LN(tok[seq[r][cur_len-1]] + pos[cur_len-1 + pos_offset])      (MFULL:1552-1563 with a cache)
// ------------------------------------------------------------------------------------------------
template <int VPL>
__global__ void __launch_bounds__(256)
decode_embed_ln_kernel(const int32_t* __restrict__ seq, const int32_t* __restrict__ cur_len_p,
                       const __nv_bfloat16* __restrict__ tok, const __nv_bfloat16* __restrict__ pos,
                       const float* __restrict__ gamma, const float* __restrict__ beta, __nv_bfloat16* __restrict__ y,
                       int R, int maxT, int d, int pos_offset, int pingpong, float eps, float* __restrict__ y32) {
  pdl_sync();
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= R) return;
  const int t = *cur_len_p;
  const int32_t* s = seq + (pingpong ? static_cast<long long>(t & 1) * R * maxT : 0);
  const long long id = s[static_cast<long long>(row) * maxT + t - 1];
  const uint4* tp = reinterpret_cast<const uint4*>(tok + id * d);
  const uint4* pp = reinterpret_cast<const uint4*>(pos + static_cast<long long>(t - 1 + pos_offset) * d);
  float v[VPL][8];
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < VPL; ++j) {
    float a[8], b[8];
    unpack8f(__ldg(tp + lane + 32 * j), a);
    unpack8f(__ldg(pp + lane + 32 * j), b);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      // the reference adds two fp32 tensors; here both tables are bf16 shadows, summed in fp32
      v[j][e] = a[e] + b[e];
      sum += v[j][e];
    }
  }
  const float mean = warp_sum(sum) / d;
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < VPL; ++j)
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float c = v[j][e] - mean;
      q += c * c;
    }
  const float rstd = rsqrtf(warp_sum(q) / d + eps);
  uint4* yp = reinterpret_cast<uint4*>(y + static_cast<long long>(row) * d);
#pragma unroll
  for (int j = 0; j < VPL; ++j) {
    const int vi = lane + 32 * j;
    float o[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = (v[j][e] - mean) * rstd * __ldg(gamma + vi * 8 + e) + __ldg(beta + vi * 8 + e);
    uint4 u;
    u.x = pack_bf16x2(o[0], o[1]); u.y = pack_bf16x2(o[2], o[3]);
    u.z = pack_bf16x2(o[4], o[5]); u.w = pack_bf16x2(o[6], o[7]);
    yp[vi] = u;
    if (y32) {  // fp32 residual stream (same as the teacher-forced forward)
      float4* y4 = reinterpret_cast<float4*>(y32 + static_cast<long long>(row) * d) + vi * 2;
      y4[0] = make_float4(o[0], o[1], o[2], o[3]);
      y4[1] = make_float4(o[4], o[5], o[6], o[7]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Cached causal self-attention of the newest token (MFULL:490-495 cache branch + :509-563).
// grid (H, R), 64 threads, head_dim == 64.  qkv row = [k | v | q] (fused projection order).
// ------------------------------------------------------------------------------------------------
constexpr int kMaxT = 256;

__global__ void __launch_bounds__(64)
decode_self_attn_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ kcache,
                        __nv_bfloat16* __restrict__ vcache, const int32_t* __restrict__ anc,
                        const int32_t* __restrict__ cur_len_p, __nv_bfloat16* __restrict__ out, int R, int maxT, int d,
                        float scale) {
  pdl_sync();
  const int h = blockIdx.x, r = blockIdx.y, j = threadIdx.x;
  const int t = *cur_len_p;
  const int pos = t - 1;
  __shared__ float qs[64], kcur[64], vcur[64], sc[kMaxT], red[2], part[8][64];
  __shared__ int srow[kMaxT];  // cache row that holds position s of this hypothesis (ancestry table), read once
  const __nv_bfloat16* row = qkv + static_cast<long long>(r) * 3 * d + h * 64;
  const __nv_bfloat16 kb = row[j], vb_ = row[d + j];
  qs[j] = __bfloat162float(row[2 * d + j]) * scale;
  kcur[j] = __bfloat162float(kb);
  vcur[j] = __bfloat162float(vb_);
  const long long self_off = (static_cast<long long>(r) * maxT + pos) * d + h * 64 + j;
  kcache[self_off] = kb;
  vcache[self_off] = vb_;
  __syncthreads();
  const int32_t* an = anc ? anc + static_cast<long long>(t & 1) * R * maxT + static_cast<long long>(r) * maxT : nullptr;
  float mx = -INFINITY;
  for (int s = j; s <= pos; s += 64) {
    float dot = 0.f;
    if (s == pos) {
#pragma unroll
      for (int e = 0; e < 64; ++e) dot += qs[e] * kcur[e];
    } else {
      const int src = an ? an[s] : r;
      srow[s] = src;
      const uint4* kp = reinterpret_cast<const uint4*>(kcache + (static_cast<long long>(src) * maxT + s) * d + h * 64);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float f[8];
        unpack8f(__ldg(kp + c), f);
#pragma unroll
        for (int e = 0; e < 8; ++e) dot += qs[c * 8 + e] * f[e];
      }
    }
    sc[s] = dot;
    mx = fmaxf(mx, dot);
  }
  mx = warp_max(mx);
  if ((j & 31) == 0) red[j >> 5] = mx;
  __syncthreads();
  mx = fmaxf(red[0], red[1]);
  __syncthreads();
  float sum = 0.f;
  for (int s = j; s <= pos; s += 64) {
    const float e = __expf(sc[s] - mx);
    sc[s] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  if ((j & 31) == 0) red[j >> 5] = sum;
  __syncthreads();
  const float inv = 1.f / (red[0] + red[1]);
  // P.V: thread j owns 8 head dims (one 16-byte load per position) of every 8th cached position -- 8 independent
  // position streams per (row, head) instead of one serial walk over the positions with one 2-byte load each -- and the
  // eight partial sums meet in shared memory.  (The walk was the part of the decode step that grew with the position:
  // +16 us per position and step at 12 layers.)
  const int sg = j >> 3, dc = j & 7;
  float acc8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
  for (int s = sg; s < pos; s += 8) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(vcache + (static_cast<long long>(srow[s]) * maxT + s) * d + h * 64) + dc);
    float f[8];
    unpack8f(u, f);
    const float p = sc[s];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc8[e] += p * f[e];
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) part[sg][dc * 8 + e] = acc8[e];
  __syncthreads();
  float acc = sc[pos] * vcur[j];
#pragma unroll
  for (int g = 0; g < 8; ++g) acc += part[g][j];
  out[static_cast<long long>(r) * d + h * 64 + j] = __float2bfloat16_rn(acc * inv);
}

// ------------------------------------------------------------------------------------------------
// Cross-attention of the beams of one caption over that caption's L cached keys (MFULL:474-479 cache branch):
// grid (H, captions), 128 threads.  HBM-bound streaming of K then V, exact max-subtracted softmax through shared
// memory in between.  Keys at or beyond key_len[c] contribute exactly 0 in the reference (exp(finfo.min - max) == 0)
// and are skipped.  The dot products run on the tensor cores (warp-level mma.sync m16n8k16, bf16 x bf16 -> fp32; a
// CUDA-core version was issue-bound at 2.8 TB/s): K / V tiles of 16 keys are staged through per-warp 6-deep
// cp.async rings in XOR-swizzled shared memory and consumed with ldmatrix:
//   pass 1   S^T[key][q]  = K_tile[key][dim] . Q^T[dim][q]      A = K (row-major), B = Q^T in registers
//   softmax  exact, over shared-memory scores (as above)
//   pass 2   O^T[dim][q] += V_tile^T[dim][key] . P^T[key][q]    A = V via ldmatrix.trans, B = P from shared memory
// Up to 8 queries (beams) per caption ride in the n = 8 dimension of one instruction.
// ------------------------------------------------------------------------------------------------
constexpr int kXTile = 16;                    // keys per tile: one m16n8k16 row block per warp step
constexpr int kXStages = 6;                   // cp.async ring depth PER WARP
constexpr int kXTileBytes = kXTile * 128;     // 2 KB
constexpr int kXWarpRing = kXStages * kXTileBytes;  // 12 KB

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// Each of the 4 warps streams its own contiguous quarter of the caption's keys through a private ring (no block-wide
// barrier inside the streaming loops); the warps meet only for the softmax statistics and the final reduction.
__global__ void __launch_bounds__(128)
decode_cross_attn_mma_kernel(const __nv_bfloat16* __restrict__ q, long long ldq, const __nv_bfloat16* __restrict__ kp,
                             const __nv_bfloat16* __restrict__ vp, long long ldkv, long long kv_hs, long long kv_cs,
                             const uint8_t* __restrict__ key_mask, const int32_t* __restrict__ key_len,
                             __nv_bfloat16* __restrict__ out, long long ldo, int nq, int L, float scale) {
  pdl_sync();
  extern __shared__ __align__(128) uint8_t xsm[];
  float* sc = reinterpret_cast<float*>(xsm + 4 * kXWarpRing);  // [nq][Lp]
  float* red = reinterpret_cast<float*>(xsm);                  // [4][64][8] partial O^T, aliases the rings after pass 2
  __shared__ float stat[2][8][4];
  const int h = blockIdx.x, c = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t4 = lane & 3;
  // key_len == 0 (no valid key at all): every key keeps finfo.min and the row degenerates to uniform, as in the reference
  const int Lc = (key_len && key_len[c] > 0) ? min(L, key_len[c]) : L;
  const int per = (Lc + 63) / 64 * kXTile;  // keys per warp (multiple of 16)
  const int ntw = per / kXTile;             // tiles per warp
  const int Lp = 4 * per;                   // score row pitch
  const int wkey0 = warp * per;
  const uint8_t* mk = key_mask ? key_mask + static_cast<long long>(c) * L : nullptr;
  const __nv_bfloat16* kbase = kp + static_cast<long long>(c) * kv_cs + h * kv_hs;
  const __nv_bfloat16* vbase = vp + static_cast<long long>(c) * kv_cs + h * kv_hs;
  uint8_t* wring = xsm + warp * kXWarpRing;
  const uint32_t ring_u = smem_u32(wring);

  // tile loader (one warp): 16 rows x 8 chunks of 16 B; lane i moves chunks i, i+32, i+64, i+96
  auto load_tile = [&](const __nv_bfloat16* base, int tile, int stage) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int idx = lane + 32 * j;
      const int row = idx >> 3, ch = idx & 7;
      const int key = wkey0 + tile * kXTile + row;
      const int off = stage * kXTileBytes + row * 128 + ((ch ^ (row & 7)) << 4);
      if (key < Lc) cp_async16(ring_u + off, base + static_cast<long long>(key) * ldkv + ch * 8);
      else *reinterpret_cast<uint4*>(wring + off) = make_uint4(0, 0, 0, 0);
    }
  };

  // Q^T fragments: b[ks][0..1] for k-steps of 16 dims; query n = g (zero rows beyond nq)
  uint32_t qb[4][2];
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    qb[ks][0] = qb[ks][1] = 0u;
    if (g < nq) {
      const __nv_bfloat16* qr = q + (static_cast<long long>(c) * nq + g) * ldq + h * 64 + ks * 16 + 2 * t4;
      qb[ks][0] = *reinterpret_cast<const uint32_t*>(qr);
      qb[ks][1] = *reinterpret_cast<const uint32_t*>(qr + 8);
    }
  }

  // ---------------- pass 1: scores of this warp's keys
  for (int s = 0; s < kXStages - 1; ++s) {
    if (s < ntw) load_tile(kbase, s, s);
    cp_async_commit();
  }
  const int arow = (lane & 7) + ((lane >> 3) & 1) * 8;  // ldmatrix row (non-transposed A)
  for (int tile = 0; tile < ntw; ++tile) {
    cp_async_wait<kXStages - 2>();
    __syncwarp();
    {  // refill the stage consumed in the previous iteration
      const int nxt = tile + kXStages - 1;
      if (nxt < ntw) load_tile(kbase, nxt, nxt % kXStages);
      cp_async_commit();
    }
    const uint32_t tb = ring_u + (tile % kXStages) * kXTileBytes;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t a[4];
      const int ch = ks * 2 + (lane >> 4);
      ldmatrix_x4(tb + arow * 128 + ((ch ^ (arow & 7)) << 4), a);
      mma_bf16_16816(acc, a, qb[ks][0], qb[ks][1]);
    }
    // C fragment: keys g, g+8 of the tile; queries 2*t4, 2*t4+1
    const int key0 = wkey0 + tile * kXTile + g;
#pragma unroll
    for (int hrow = 0; hrow < 2; ++hrow) {
      const int key = key0 + hrow * 8;
      const bool oob = key >= Lc;
      const bool masked = !oob && mk != nullptr && mk[key] == 0;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int qi = 2 * t4 + e;
        if (qi < nq) sc[qi * Lp + key] = oob ? -INFINITY : (masked ? -FLT_MAX : acc[hrow * 2 + e] * scale);
      }
    }
    __syncwarp();
  }
  cp_async_wait<0>();
  __syncwarp();
  // prefetch the first V tiles of this warp while the softmax runs
  for (int s = 0; s < kXStages - 1; ++s) {
    if (s < ntw) load_tile(vbase, s, s);
    cp_async_commit();
  }
  __syncthreads();
  // ---------------- exact softmax statistics per query (probabilities stay unnormalised in sc)
  for (int i = 0; i < nq; ++i) {
    float m = -INFINITY;
    for (int k = tid; k < Lp; k += 128) m = fmaxf(m, sc[i * Lp + k]);
    m = warp_max(m);
    if (lane == 0) stat[0][i][warp] = m;
  }
  __syncthreads();
  for (int i = 0; i < nq; ++i) {
    const float m = fmaxf(fmaxf(stat[0][i][0], stat[0][i][1]), fmaxf(stat[0][i][2], stat[0][i][3]));
    float sum = 0.f;
    for (int k = tid; k < Lp; k += 128) {
      const float e = __expf(sc[i * Lp + k] - m);
      sc[i * Lp + k] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    if (lane == 0) stat[1][i][warp] = sum;
  }
  __syncthreads();
  // ---------------- pass 2: O^T[dim][q] += V^T P^T over this warp's keys
  float o[4][4];
#pragma unroll
  for (int mt = 0; mt < 4; ++mt)
#pragma unroll
    for (int e = 0; e < 4; ++e) o[mt][e] = 0.f;
  const int krow = (lane & 7) + (lane >> 4) * 8;  // key row for ldmatrix.trans
  for (int tile = 0; tile < ntw; ++tile) {
    cp_async_wait<kXStages - 2>();
    __syncwarp();
    {
      const int nxt = tile + kXStages - 1;
      if (nxt < ntw) load_tile(vbase, nxt, nxt % kXStages);
      cp_async_commit();
    }
    const uint32_t tb = ring_u + (tile % kXStages) * kXTileBytes;
    // B fragment: P[q = g][keys 2*t4, 2*t4+1] and [+8, +9] of the tile
    uint32_t pb0 = 0u, pb1 = 0u;
    if (g < nq) {
      const float* pr = sc + g * Lp + wkey0 + tile * kXTile + 2 * t4;
      const float2 p01 = *reinterpret_cast<const float2*>(pr);
      const float2 p89 = *reinterpret_cast<const float2*>(pr + 8);
      pb0 = pack_bf16x2(p01.x, p01.y);
      pb1 = pack_bf16x2(p89.x, p89.y);
    }
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {
      uint32_t a[4];
      const int ch = mt * 2 + ((lane >> 3) & 1);
      ldmatrix_x4_trans(tb + krow * 128 + ((ch ^ (krow & 7)) << 4), a);
      mma_bf16_16816(o[mt], a, pb0, pb1);
    }
    __syncwarp();
  }
  cp_async_wait<0>();
  __syncthreads();  // every warp is done with its ring: reuse the space for the cross-warp reduction
#pragma unroll
  for (int mt = 0; mt < 4; ++mt)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int dim = mt * 16 + g + (e >> 1) * 8, qi = 2 * t4 + (e & 1);
      red[(warp * 64 + dim) * 8 + qi] = o[mt][e];
    }
  __syncthreads();
  for (int idx = tid; idx < nq * 64; idx += 128) {
    const int i = idx >> 6, dim = idx & 63;
    const float v = red[(0 * 64 + dim) * 8 + i] + red[(1 * 64 + dim) * 8 + i] + red[(2 * 64 + dim) * 8 + i] + red[(3 * 64 + dim) * 8 + i];
    const float ssum = stat[1][i][0] + stat[1][i][1] + stat[1][i][2] + stat[1][i][3];
    out[(static_cast<long long>(c) * nq + i) * ldo + h * 64 + dim] = __float2bfloat16_rn(v / ssum);
  }
}

// key_len[b] = 1 + index of the last non-zero mask byte (0 when the row is fully masked)
__global__ void mask_key_len_kernel(const uint8_t* __restrict__ mask, int32_t* __restrict__ out, int B, int L) {
  const int b = blockIdx.x;
  int last = 0;
  for (int k = threadIdx.x; k < L; k += blockDim.x)
    if (mask[static_cast<long long>(b) * L + k]) last = max(last, k + 1);
  last = __reduce_max_sync(0xffffffffu, last);
  __shared__ int red[32];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = last;
  __syncthreads();
  if (threadIdx.x == 0) {
    int m = 0;
    for (int w = 0; w < (blockDim.x + 31) / 32; ++w) m = max(m, red[w]);
    out[b] = m;
  }
}

// ------------------------------------------------------------------------------------------------
// Per-row log-softmax statistics + top-K of the LM-head logits (the `log_softmax` and the row-local part
// of `torch.topk(..., 2*num_beams)` of _beam_search; greedy uses K = 1).  Total order everywhere: value descending,
// then vocabulary index ascending; -inf entries (suppressed tokens) are never returned.
//
// One thread-block CLUSTER of 8 CTAs per row.  Every lane keeps E elements of the row in REGISTERS (E coalesced 4-byte
// loads, all in flight at once; no shared-memory staging), so the 200 KB row is read from HBM exactly once and several
// CTAs per SM overlap each other's loads and selection.  Selection is by THRESHOLD, not by K extraction rounds:
//   warp:    T_w = the K-th largest of its 32 lane maxima (found by counting, no sorting); at least K elements of the
//            row are >= T_w, so T_w is a lower bound of the row's K-th best
//   CTA:     T_c = best T_w of its 8 warps; every element >= T_c goes to a shared-memory candidate list (a few dozen
//            for K = 8); the K best of the list are found by counting ranks, one candidate per thread
//   cluster: every CTA stores its K best and its (max, sum of exp) into CTA 0's shared memory (distributed shared
//            memory), arrives on the cluster barrier and exits; CTA 0 alone waits, ranks the 8 K candidates and writes
//            the row's K (log-prob, index) pairs.
// Inputs whose finite values sit in fewer than K lanes of every warp (threshold -inf), or with hundreds of values tied at
// the threshold, overflow the candidate list; that CTA then falls back to K rounds of block-wide arg-max over its registers.
// ------------------------------------------------------------------------------------------------
struct ValIdx {
  float v;
  int i;
};
__device__ __forceinline__ bool better(float v, int i, float w, int k) { return v > w || (v == w && i < k); }
__device__ __forceinline__ ValIdx warp_argmax(ValIdx a) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float w = __shfl_xor_sync(0xffffffffu, a.v, o);
    const int k = __shfl_xor_sync(0xffffffffu, a.i, o);
    if (better(w, k, a.v, a.i)) { a.v = w; a.i = k; }
  }
  return a;
}

constexpr int kMaxBeams = 8;
constexpr int kMaxCand = 2 * kMaxBeams * kMaxBeams;  // nb rows x K = 2 nb candidates each
constexpr int kTopK = 2 * kMaxBeams;                 // largest K
constexpr int kTopCluster = 8;                       // CTAs per row
constexpr int kTopWarps = 8;
constexpr int kTopThreads = 32 * kTopWarps;
constexpr int kTopCap = kTopThreads;                 // candidate list: one candidate per thread in the ranking pass
constexpr int kNoIdx = 0x7fffffff;

struct TopPart {  // one CTA's share of a row
  float m, s;     // max, sum of exp(x - m)
  float v[kTopK];
  int i[kTopK];
};

__device__ __forceinline__ uint32_t dsmem_addr(const void* p, uint32_t cta_rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(cta_rank));
  return r;
}
__device__ __forceinline__ void dsmem_st_f32(uint32_t a, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory");
}
__device__ __forceinline__ void dsmem_st_s32(uint32_t a, int v) {
  asm volatile("st.shared::cluster.s32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
// no memory ordering (no MEMBAR / ERRBAR): for warps that published nothing
__device__ __forceinline__ void cluster_arrive_relaxed() { asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

template <int E>
__global__ void __cluster_dims__(kTopCluster, 1, 1) __launch_bounds__(kTopThreads, E <= 26 ? 4 : 3)
decode_topk_kernel(const float* __restrict__ logits, long long ld, int V, int K, float* __restrict__ top_lp,
                   int32_t* __restrict__ top_idx) {
  cluster_arrive_relaxed();  // phase 0: "this CTA is running" -- waited for right before the first remote store
  pdl_sync();
  __shared__ TopPart parts[kTopCluster];  // gather target; only CTA 0's copy is read
  __shared__ float w_m[kTopWarps], w_s[kTopWarps], thr_v[kTopWarps];
  __shared__ __align__(16) float lane_max[kTopWarps][32];
  __shared__ float cand_v[kTopCap], sel_v[kTopK];
  __shared__ int cand_i[kTopCap], sel_i[kTopK];
  __shared__ int n_cand;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int crank = blockIdx.x;  // gridDim.x == cluster size: the CTA's rank in its cluster
  const long long row = blockIdx.y;
  const float* src = logits + row * ld;
  const int base = (crank * kTopWarps + warp) * (32 * E) + lane;  // the lane's first element; stride 32
  float x[E];
  if (base - lane + 32 * E <= V) {  // warp-uniform: the whole segment is inside the row
#pragma unroll
    for (int j = 0; j < E; ++j) x[j] = __ldg(src + base + j * 32);
  } else {
#pragma unroll
    for (int j = 0; j < E; ++j) {
      const int i = base + j * 32;
      x[j] = i < V ? __ldg(src + i) : -INFINITY;
    }
  }
  if (tid == 0) n_cand = 0;
  if (tid < kTopK) { sel_v[tid] = -INFINITY; sel_i[tid] = kNoIdx; }
  // ---- lane maximum, warp max / sum of exp
  float lmv = -INFINITY;
#pragma unroll
  for (int j = 0; j < E; ++j) lmv = fmaxf(lmv, x[j]);
  const float wm = warp_max(lmv);
  float sum = 0.f;
  if (wm > -INFINITY) {
    constexpr float kLog2e = 1.4426950408889634f;
    const float nb = -wm * kLog2e;
#pragma unroll
    for (int j = 0; j < E; ++j) sum += ex2_approx(fmaf(x[j], kLog2e, nb));  // |d lse| ~ 1e-6, far below score gaps
  }
  sum = warp_sum(sum);
  // ---- T_w: the K-th largest of the 32 lane maxima, by counting (values only: equal values form one group, and the
  // filter below accepts everything >= T, so ties never lose a candidate).  -inf when fewer than K lanes hold a finite value.
  lane_max[warp][lane] = lmv;
  __syncwarp();
  int gt = 0, ge = 0;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float4 f = reinterpret_cast<const float4*>(lane_max[warp])[q];  // broadcast read
    gt += (f.x > lmv) + (f.y > lmv) + (f.z > lmv) + (f.w > lmv);
    ge += (f.x >= lmv) + (f.y >= lmv) + (f.z >= lmv) + (f.w >= lmv);
  }
  const unsigned holder = __ballot_sync(0xffffffffu, gt < K && K <= ge && lmv > -INFINITY);
  float T = -INFINITY;
  if (holder) T = __shfl_sync(0xffffffffu, lmv, __ffs(holder) - 1);
  if (lane == 0) { w_m[warp] = wm; w_s[warp] = sum; thr_v[warp] = T; }
  __syncthreads();
#pragma unroll
  for (int w = 0; w < kTopWarps; ++w) T = fmaxf(T, thr_v[w]);
  // ---- candidates: finite elements >= T_c (warp-aggregated append; hits are rare, the branch is warp-uniform)
#pragma unroll
  for (int j = 0; j < E; ++j) {
    const bool p = x[j] >= T && x[j] > -INFINITY;
    const unsigned b = __ballot_sync(0xffffffffu, p);
    if (b) {
      int pos = 0;
      if (lane == 0) pos = atomicAdd(&n_cand, __popc(b));
      pos = __shfl_sync(0xffffffffu, pos, 0) + __popc(b & ((1u << lane) - 1u));
      if (p && pos < kTopCap) { cand_v[pos] = x[j]; cand_i[pos] = base + j * 32; }
    }
  }
  __syncthreads();
  const int n = n_cand;
  if (n <= kTopCap) {
    if (tid < n) {  // rank by counting; indices are distinct, so ranks are
      const float v = cand_v[tid];
      const int i = cand_i[tid];
      int r = 0;
      for (int u = 0; u < n; ++u) r += better(cand_v[u], cand_i[u], v, i) ? 1 : 0;
      if (r < K) { sel_v[r] = v; sel_i[r] = i; }
    }
  } else {
    // fallback: K rounds of block-wide arg-max over the registers (cand_* reused as the per-warp exchange)
    for (int k = 0; k < K; ++k) {
      ValIdx b = {-INFINITY, kNoIdx};
#pragma unroll
      for (int j = 0; j < E; ++j)
        if (x[j] > b.v) { b.v = x[j]; b.i = base + j * 32; }
      b = warp_argmax(b);
      __syncthreads();
      if (lane == 0) { cand_v[warp] = b.v; cand_i[warp] = b.i; }
      __syncthreads();
#pragma unroll
      for (int w = 0; w < kTopWarps; ++w)
        if (better(cand_v[w], cand_i[w], b.v, b.i)) { b.v = cand_v[w]; b.i = cand_i[w]; }
      if (tid == 0) { sel_v[k] = b.v; sel_i[k] = b.i; }
#pragma unroll
      for (int j = 0; j < E; ++j)
        if (base + j * 32 == b.i) x[j] = -INFINITY;
    }
  }
  __syncthreads();
  // ---- hand the CTA's result to CTA 0 of the cluster
  cluster_wait();  // phase 0 complete: every CTA of the cluster has started, its shared memory may be written
  const uint32_t dst = dsmem_addr(&parts[crank], 0);
  if (tid < K) {
    dsmem_st_f32(dst + offsetof(TopPart, v) + 4 * tid, sel_v[tid]);
    dsmem_st_s32(dst + offsetof(TopPart, i) + 4 * tid, sel_i[tid]);
  } else if (tid == 31) {  // K <= 16: a free lane of warp 0, the only warp that publishes
    float cm = w_m[0];
#pragma unroll
    for (int w = 1; w < kTopWarps; ++w) cm = fmaxf(cm, w_m[w]);
    float cs = 0.f;
#pragma unroll
    for (int w = 0; w < kTopWarps; ++w)
      if (w_m[w] > -INFINITY) cs += w_s[w] * __expf(w_m[w] - cm);
    dsmem_st_f32(dst + offsetof(TopPart, m), cm);
    dsmem_st_f32(dst + offsetof(TopPart, s), cs);
  }
  __syncwarp();
  if (warp == 0) cluster_arrive();  // phase 1 (release): this CTA's stores are done
  else cluster_arrive_relaxed();
  if (crank != 0) return;
  cluster_wait();    // phase 1 (acquire): all 8 parts are in
  // ---- CTA 0: rank the 8 K candidates
  const int total = kTopCluster * K;
  if (tid < kTopK) { sel_v[tid] = -INFINITY; sel_i[tid] = kNoIdx; }
  if (tid < total) { cand_v[tid] = parts[tid / K].v[tid % K]; cand_i[tid] = parts[tid / K].i[tid % K]; }
  __syncthreads();
  if (tid < total && cand_i[tid] != kNoIdx) {
    const float v = cand_v[tid];
    const int i = cand_i[tid];
    int r = 0;
    for (int u = 0; u < total; ++u) r += better(cand_v[u], cand_i[u], v, i) ? 1 : 0;
    if (r < K) { sel_v[r] = v; sel_i[r] = i; }
  }
  __syncthreads();
  if (tid < K) {
    float M = parts[0].m;
#pragma unroll
    for (int c = 1; c < kTopCluster; ++c) M = fmaxf(M, parts[c].m);
    float S = 0.f;
#pragma unroll
    for (int c = 0; c < kTopCluster; ++c)
      if (parts[c].m > -INFINITY) S += parts[c].s * expf(parts[c].m - M);
    const float lse = M + logf(S);
    top_lp[row * K + tid] = sel_v[tid] - lse;
    top_idx[row * K + tid] = sel_i[tid];
  }
}

// ------------------------------------------------------------------------------------------------
// One step of transformers' `_beam_search` (steps c-g, oracle/generate.py:beam_search) for every caption:
// one warp per caption.  State buffers are ping-pong pairs indexed by (cur_len & 1) -> ((cur_len+1) & 1).
// ------------------------------------------------------------------------------------------------
struct BeamArgs {
  const float* top_lp;      // [R][K]   row-local log-probs, sorted
  const int32_t* top_idx;   // [R][K]
  int32_t* run_seq;         // [2][C][nb][maxT]
  int32_t* run_anc;         // [2][C][nb][maxT]   global cache row of the ancestor at each position
  float* run_score;         // [2][C][nb]
  int32_t* fin_seq;         // [2][C][nb][maxT]
  float* fin_score;         // [2][C][nb]
  int32_t* fin_len;         // [2][C][nb]   generated tokens of the finished hypothesis (0 = empty slot)
  uint8_t* fin_flag;        // [2][C][nb]
  uint8_t* unsat;           // [C]
  int32_t* flags;           // [maxT][2]   per step: {any caption still improving, any candidate not stopped}
  const int32_t* cur_len_p;
  int C, nb, K, maxT, max_len, eos, V;
  float length_penalty;
};


__global__ void __launch_bounds__(128) beam_step_kernel(const BeamArgs a) {
  pdl_sync();
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int c = blockIdx.x * 4 + wib;
  __shared__ float s_lp[4][kMaxCand];
  __shared__ int s_tok[4][kMaxCand], s_src[4][kMaxCand];
  __shared__ float k_lp[4][2 * kMaxBeams], k_run[4][2 * kMaxBeams], k_fin[4][2 * kMaxBeams];
  __shared__ int k_tok[4][2 * kMaxBeams], k_src[4][2 * kMaxBeams], k_hit[4][2 * kMaxBeams];
  __shared__ int n_sel[4][kMaxBeams], f_sel[4][kMaxBeams];
  __shared__ float m_sc[4][3 * kMaxBeams];
  if (c >= a.C) return;
  const int t = *a.cur_len_p;
  const int nb = a.nb, K = a.K, maxT = a.maxT;
  const int ib = t & 1, ob = ib ^ 1;
  const long long SB = static_cast<long long>(a.C) * nb;  // rows per ping-pong buffer
  const int ncand = nb * K;
  const bool forced = (t == a.max_len - 1);  // ForcedEOSTokenLogitsProcessor
  // ---- c. candidates: running score + row-local log-prob
  for (int i = lane; i < ncand; i += 32) {
    const int b = i / K, k = i % K;
    const long long r = static_cast<long long>(c) * nb + b;
    float lp = a.top_lp[r * K + k];
    int tok = a.top_idx[r * K + k];
    if (forced) {
      lp = (k == 0) ? 0.f : -INFINITY;
      tok = (k == 0) ? a.eos : tok;
      if (k != 0 && tok == a.eos) tok = (a.eos + 1) % a.V;
    }
    s_lp[wib][i] = a.run_score[ib * SB + r] + lp;
    s_tok[wib][i] = tok;
    s_src[wib][i] = b;
  }
  __syncwarp();
  // rank = number of candidates that beat this one (value desc, flat index b*V+tok asc)
  for (int i = lane; i < ncand; i += 32) {
    const float v = s_lp[wib][i];
    const long long fi = static_cast<long long>(s_src[wib][i]) * a.V + s_tok[wib][i];
    int rank = 0;
    for (int j = 0; j < ncand; ++j) {
      const float w = s_lp[wib][j];
      const long long fj = static_cast<long long>(s_src[wib][j]) * a.V + s_tok[wib][j];
      rank += (w > v || (w == v && fj < fi)) ? 1 : 0;
    }
    if (rank < K) {
      k_lp[wib][rank] = v;
      k_tok[wib][rank] = s_tok[wib][i];
      k_src[wib][rank] = s_src[wib][i];
    }
  }
  __syncwarp();
  // ---- d. stopping criteria, e. running log-probs, f. finished log-probs
  const bool uns = a.unsat[c] != 0;
  if (lane < K) {
    const int hit = (k_tok[wib][lane] == a.eos) || (t + 1 >= a.max_len);
    k_hit[wib][lane] = hit;
    k_run[wib][lane] = k_lp[wib][lane] + (hit ? -1.0e9f : 0.f);
    const float denom = static_cast<float>(pow(static_cast<double>(t + 1 - 1), static_cast<double>(a.length_penalty)));
    float f = k_lp[wib][lane] / denom;
    f = f + (uns ? 0.f : -1.0e9f);
    const bool just = hit && lane < nb;
    f = f + (just ? 0.f : -1.0e9f);
    k_fin[wib][lane] = f;
  }
  __syncwarp();
  // ---- e. next running beams = top nb of k_run
  if (lane < K) {
    const float v = k_run[wib][lane];
    int rank = 0;
    for (int j = 0; j < K; ++j) {
      const float w = k_run[wib][j];
      rank += (w > v || (w == v && j < lane)) ? 1 : 0;
    }
    if (rank < nb) n_sel[wib][rank] = lane;
  }
  // ---- f. merge finished set: [old finished (nb) | candidates (K)] -> top nb
  for (int i = lane; i < nb + K; i += 32)
    m_sc[wib][i] = i < nb ? a.fin_score[ib * SB + static_cast<long long>(c) * nb + i] : k_fin[wib][i - nb];
  __syncwarp();
  for (int i = lane; i < nb + K; i += 32) {
    const float v = m_sc[wib][i];
    int rank = 0;
    for (int j = 0; j < nb + K; ++j) {
      const float w = m_sc[wib][j];
      rank += (w > v || (w == v && j < i)) ? 1 : 0;
    }
    if (rank < nb) f_sel[wib][rank] = i;
  }
  __syncwarp();
  // ---- write the next state
  int any_not_hit = 0;
  for (int j = 0; j < K; ++j) any_not_hit |= !k_hit[wib][j];
  for (int j = 0; j < nb; ++j) {
    const int cand = n_sel[wib][j];
    const int srcb = k_src[wib][cand];
    const long long in_row = ib * SB + static_cast<long long>(c) * nb + srcb;
    const long long out_row = ob * SB + static_cast<long long>(c) * nb + j;
    for (int s = lane; s < maxT; s += 32) {
      int tokv = a.run_seq[in_row * maxT + s];
      int ancv = a.run_anc[in_row * maxT + s];
      if (s == t) tokv = k_tok[wib][cand];
      if (s == t - 1) ancv = static_cast<int>(static_cast<long long>(c) * nb + srcb);
      a.run_seq[out_row * maxT + s] = tokv;
      a.run_anc[out_row * maxT + s] = ancv;
    }
    if (lane == 0) a.run_score[out_row] = k_run[wib][cand];
    // finished slot j
    const int m = f_sel[wib][j];
    if (m < nb) {
      const long long fr = ib * SB + static_cast<long long>(c) * nb + m;
      for (int s = lane; s < maxT; s += 32) a.fin_seq[out_row * maxT + s] = a.fin_seq[fr * maxT + s];
      if (lane == 0) {
        a.fin_score[out_row] = a.fin_score[fr];
        a.fin_len[out_row] = a.fin_len[fr];
        a.fin_flag[out_row] = a.fin_flag[fr];
      }
    } else {
      const int cd = m - nb;
      const long long sr = ib * SB + static_cast<long long>(c) * nb + k_src[wib][cd];
      for (int s = lane; s < maxT; s += 32) {
        int tokv = a.run_seq[sr * maxT + s];
        if (s == t) tokv = k_tok[wib][cd];
        a.fin_seq[out_row * maxT + s] = tokv;
      }
      if (lane == 0) {
        const bool just = k_hit[wib][cd] && cd < nb;
        a.fin_score[out_row] = k_fin[wib][cd];
        a.fin_len[out_row] = t;  // generated tokens = cur_len + 1 - prompt
        a.fin_flag[out_row] = just ? 1 : 0;
      }
    }
  }
  __syncwarp();
  // ---- g. early-stop heuristic (early_stopping=False): can the best running beam still improve?
  if (lane == 0) {
    const long long o0 = ob * SB + static_cast<long long>(c) * nb;
    const float denom = static_cast<float>(pow(static_cast<double>(t + 1 - 1), static_cast<double>(a.length_penalty)));
    const float best_possible = k_run[wib][n_sel[wib][0]] / denom;
    float mn = INFINITY;
    for (int j = 0; j < nb; ++j) {
      const int m = f_sel[wib][j];
      mn = fminf(mn, m_sc[wib][m]);
    }
    bool any = false;
    for (int j = 0; j < nb; ++j) {
      const int m = f_sel[wib][j];
      const bool fin = m < nb ? (a.fin_flag[ib * SB + static_cast<long long>(c) * nb + m] != 0)
                              : (k_hit[wib][m - nb] && (m - nb) < nb);
      const float worst = fin ? mn : -1.0e9f;
      any |= best_possible > worst;
    }
    (void)o0;
    const bool nu = uns && any;
    a.unsat[c] = nu ? 1 : 0;
    if (nu) atomicOr(a.flags + 2 * t, 1);
    if (any_not_hit) atomicOr(a.flags + 2 * t + 1, 1);
  }
}

// greedy `_sample` step: next = argmax (forced eos at the last position), finished rows emit pad
__global__ void greedy_step_kernel(const int32_t* __restrict__ top_idx, int32_t* __restrict__ seq,
                                   uint8_t* __restrict__ unfinished, int32_t* __restrict__ flags,
                                   const int32_t* __restrict__ cur_len_p, int R, int maxT, int max_len, int eos, int pad) {
  pdl_sync();
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const int t = *cur_len_p;
  int tok = (t == max_len - 1) ? eos : top_idx[r];
  const bool un = unfinished[r] != 0;
  if (!un) tok = pad;
  seq[static_cast<long long>(r) * maxT + t] = tok;
  const bool nu = un && tok != eos && (t + 1 < max_len);
  unfinished[r] = nu ? 1 : 0;
  if (nu) atomicOr(flags + 2 * t, 1);
  atomicOr(flags + 2 * t + 1, 1);
}

__global__ void advance_len_kernel(int32_t* cur_len_p) {
  pdl_sync();
  *cur_len_p += 1;
}

}  // namespace vb

using namespace vb;

extern "C" int vacnic_decode_embed_ln(const int32_t* seq, const int32_t* cur_len, const void* tok, const void* pos,
                                      const float* gamma, const float* beta, void* y, int32_t R, int32_t maxT, int32_t d,
                                      int32_t pos_offset, int32_t pingpong, float eps, float* y32, void* stream) {
  VB_REQUIRE(seq && cur_len && tok && pos && gamma && beta && y, "decode_embed_ln: null pointer");
  VB_REQUIRE(R > 0 && maxT > 0 && d > 0 && d % 256 == 0 && d <= 1024, "decode_embed_ln: d must be 256/512/768/1024");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int grid = (R + 7) / 8;
  const __nv_bfloat16* tk = static_cast<const __nv_bfloat16*>(tok);
  const __nv_bfloat16* ps = static_cast<const __nv_bfloat16*>(pos);
  __nv_bfloat16* yy = static_cast<__nv_bfloat16*>(y);
  switch (d / 256) {
    case 1: launch_pdl(decode_embed_ln_kernel<1>, grid, dim3(256), 0, s, seq, cur_len, tk, ps, gamma, beta, yy, R, maxT, d, pos_offset, pingpong, eps, y32); break;
    case 2: launch_pdl(decode_embed_ln_kernel<2>, grid, dim3(256), 0, s, seq, cur_len, tk, ps, gamma, beta, yy, R, maxT, d, pos_offset, pingpong, eps, y32); break;
    case 3: launch_pdl(decode_embed_ln_kernel<3>, grid, dim3(256), 0, s, seq, cur_len, tk, ps, gamma, beta, yy, R, maxT, d, pos_offset, pingpong, eps, y32); break;
    default: launch_pdl(decode_embed_ln_kernel<4>, grid, dim3(256), 0, s, seq, cur_len, tk, ps, gamma, beta, yy, R, maxT, d, pos_offset, pingpong, eps, y32); break;
  }
  count_launch();
  return check_last("decode_embed_ln");
}

extern "C" int vacnic_decode_self_attn(const void* qkv, void* kcache, void* vcache, const int32_t* anc,
                                       const int32_t* cur_len, void* out, int32_t R, int32_t H, int32_t head_dim,
                                       int32_t maxT, void* stream) {
  VB_REQUIRE(qkv && kcache && vcache && cur_len && out, "decode_self_attn: null pointer");
  VB_REQUIRE(head_dim == 64, "decode_self_attn: head_dim must be 64 (BART-base / BART-large)");
  VB_REQUIRE(R > 0 && H > 0 && maxT > 0 && maxT <= kMaxT, "decode_self_attn: maxT must be in 1..%d", kMaxT);
  const int d = H * head_dim;
  launch_pdl(decode_self_attn_kernel, dim3(H, R), dim3(64), 0, static_cast<cudaStream_t>(stream), static_cast<const __nv_bfloat16*>(qkv), static_cast<__nv_bfloat16*>(kcache), static_cast<__nv_bfloat16*>(vcache), anc, cur_len, static_cast<__nv_bfloat16*>(out), R, maxT, d, 1.0f / sqrtf(static_cast<float>(head_dim)));
  count_launch();
  return check_last("decode_self_attn");
}

extern "C" int vacnic_decode_cross_attn(const void* q, int64_t ldq, const void* k, const void* v, int64_t ldkv,
                                        int64_t kv_hs, int64_t kv_cs, const uint8_t* key_mask, const int32_t* key_len,
                                        void* out, int64_t ldo, int32_t captions, int32_t nq, int32_t L, int32_t H,
                                        int32_t head_dim, void* stream) {
  VB_REQUIRE(q && k && v && out, "decode_cross_attn: null pointer");
  VB_REQUIRE(head_dim == 64, "decode_cross_attn: head_dim must be 64");
  VB_REQUIRE(captions > 0 && nq >= 1 && nq <= 8 && L > 0 && L <= 4096 && H > 0, "decode_cross_attn: bad shape (nq <= 8, L <= 4096)");
  VB_REQUIRE(ldq % 8 == 0 && ldkv % 8 == 0 && kv_hs % 8 == 0 && kv_cs % 8 == 0, "decode_cross_attn: strides must be multiples of 8 elements");
  VB_REQUIRE((reinterpret_cast<uintptr_t>(q) & 15) == 0 && (reinterpret_cast<uintptr_t>(k) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(v) & 15) == 0, "decode_cross_attn: misaligned");
  const float scale = 1.0f / sqrtf(static_cast<float>(head_dim));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  {
    // tensor-core streaming kernel (mma.sync over cp.async-staged tiles)
    const int Lp = (L + 63) / 64 * 64;
    const size_t smem = 4 * static_cast<size_t>(kXWarpRing) + static_cast<size_t>(nq) * Lp * sizeof(float);
    static size_t configured = 0;
    if (smem > configured) {  // static (8.4 KB) + dynamic shared memory crosses the 48 KB opt-in line early: always opt in
      cudaError_t e = cudaFuncSetAttribute(decode_cross_attn_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
      if (e != cudaSuccess) return fail(VACNIC_ECUDA, "decode_cross_attn: smem %zu: %s", smem, cudaGetErrorString(e));
      configured = smem;
    }
    VB_REQUIRE(smem <= 200 * 1024, "decode_cross_attn: nq * L too large for the shared-memory score buffer");
    launch_pdl(decode_cross_attn_mma_kernel, dim3(H, captions), dim3(128), smem, s, static_cast<const __nv_bfloat16*>(q), ldq, static_cast<const __nv_bfloat16*>(k), static_cast<const __nv_bfloat16*>(v), ldkv, kv_hs, kv_cs, key_mask, key_len, static_cast<__nv_bfloat16*>(out), ldo, nq, L, scale);
    count_launch();
    return check_last("decode_cross_attn");
  }
}

extern "C" int vacnic_mask_key_len(const uint8_t* mask, int32_t* key_len, int32_t B, int32_t L, void* stream) {
  VB_REQUIRE(mask && key_len && B > 0 && L > 0, "mask_key_len: bad arguments");
  mask_key_len_kernel<<<B, 256, 0, static_cast<cudaStream_t>(stream)>>>(mask, key_len, B, L);
  count_launch();
  return check_last("mask_key_len");
}

extern "C" int vacnic_decode_topk(const float* logits, int64_t ld, int32_t rows, int32_t V, int32_t K, float* top_lp,
                                  int32_t* top_idx, void* stream) {
  VB_REQUIRE(logits && top_lp && top_idx, "decode_topk: null pointer");
  VB_REQUIRE(rows > 0 && rows <= 65535 && V > 0 && K >= 1 && K <= 2 * kMaxBeams && K <= V,
             "decode_topk: bad shape (rows <= 65535, K <= %d)", 2 * kMaxBeams);
  VB_REQUIRE(V <= kTopCluster * kTopThreads * 32, "decode_topk: vocabulary %d exceeds %d", V, kTopCluster * kTopThreads * 32);
  const dim3 grid(kTopCluster, rows), block(kTopThreads);
  const cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int per_lane = (V + kTopCluster * kTopThreads - 1) / (kTopCluster * kTopThreads);  // elements each lane holds
  if (per_lane <= 2) launch_pdl(decode_topk_kernel<2>, grid, block, 0, st, logits, ld, V, K, top_lp, top_idx);
  else if (per_lane <= 8) launch_pdl(decode_topk_kernel<8>, grid, block, 0, st, logits, ld, V, K, top_lp, top_idx);
  else if (per_lane <= 26) launch_pdl(decode_topk_kernel<26>, grid, block, 0, st, logits, ld, V, K, top_lp, top_idx);
  else launch_pdl(decode_topk_kernel<32>, grid, block, 0, st, logits, ld, V, K, top_lp, top_idx);
  count_launch();
  return check_last("decode_topk");
}

extern "C" int vacnic_beam_step(const float* top_lp, const int32_t* top_idx, int32_t* run_seq, int32_t* run_anc,
                                float* run_score, int32_t* fin_seq, float* fin_score, int32_t* fin_len,
                                uint8_t* fin_flag, uint8_t* unsat, int32_t* flags, const int32_t* cur_len,
                                int32_t captions, int32_t beams, int32_t maxT, int32_t max_len, int32_t eos, int32_t V,
                                float length_penalty, void* stream) {
  VB_REQUIRE(top_lp && top_idx && run_seq && run_anc && run_score && fin_seq && fin_score && fin_len && fin_flag && unsat &&
                 flags && cur_len, "beam_step: null pointer");
  VB_REQUIRE(captions > 0 && beams >= 1 && beams <= kMaxBeams, "beam_step: beams must be in 1..%d", kMaxBeams);
  VB_REQUIRE(max_len >= 2 && max_len <= maxT, "beam_step: max_len must be in 2..maxT");
  BeamArgs a;
  a.top_lp = top_lp; a.top_idx = top_idx; a.run_seq = run_seq; a.run_anc = run_anc; a.run_score = run_score;
  a.fin_seq = fin_seq; a.fin_score = fin_score; a.fin_len = fin_len; a.fin_flag = fin_flag; a.unsat = unsat;
  a.flags = flags; a.cur_len_p = cur_len;
  a.C = captions; a.nb = beams; a.K = 2 * beams; a.maxT = maxT; a.max_len = max_len; a.eos = eos; a.V = V;
  a.length_penalty = length_penalty;
  launch_pdl(beam_step_kernel, dim3((captions + 3) / 4), dim3(128), 0, static_cast<cudaStream_t>(stream), a);
  count_launch();
  return check_last("beam_step");
}

extern "C" int vacnic_greedy_step(const int32_t* top_idx, int32_t* seq, uint8_t* unfinished, int32_t* flags,
                                  const int32_t* cur_len, int32_t rows, int32_t maxT, int32_t max_len, int32_t eos,
                                  int32_t pad, void* stream) {
  VB_REQUIRE(top_idx && seq && unfinished && flags && cur_len, "greedy_step: null pointer");
  VB_REQUIRE(rows > 0 && max_len >= 2 && max_len <= maxT, "greedy_step: bad shape");
  launch_pdl(greedy_step_kernel, dim3((rows + 127) / 128), dim3(128), 0, static_cast<cudaStream_t>(stream), top_idx, seq, unfinished, flags, cur_len, rows, maxT, max_len, eos, pad);
  count_launch();
  return check_last("greedy_step");
}

extern "C" int vacnic_advance_len(int32_t* cur_len, void* stream) {
  VB_REQUIRE(cur_len, "advance_len: null pointer");
  launch_pdl(advance_len_kernel, dim3(1), dim3(1), 0, static_cast<cudaStream_t>(stream), cur_len);
  count_launch();
  return check_last("advance_len");
}
