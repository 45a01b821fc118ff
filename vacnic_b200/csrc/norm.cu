// LayerNorm family for the VACNIC hot path (HBM-bound, one warp per row, 16-byte vector access):
//   add_layernorm fwd/bwd : y = LN(res + dropout(x))            (MFULL:652-653, 663-664, 678-679,
//                                                                 688, 705-707, 721-723, 742-744, 839-841, ...)
//   embed_ln fwd/bwd      : y = dropout(LN(E[ids] + Pos[t + off]))   (MFULL:1243-1249, 1254-1260, 1555-1563)
//   names_embed           : mean_t LN(E[ids] + Pos[t + 2])       (get_embedding_ner, TRAIN:112-133)
// d_model must be a multiple of 256 (768 and 1024 are the only widths the reference can build:
// MFULL:1136, 1142).  Statistics and parameter gradients are fp32; activations bf16.
#include "common.h"
#include "ptx.cuh"

namespace vb {

constexpr int kWarpsPerBlock = 8;

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  float2 t;
  t = unpack_bf16x2(u.x); f[0] = t.x; f[1] = t.y;
  t = unpack_bf16x2(u.y); f[2] = t.x; f[3] = t.y;
  t = unpack_bf16x2(u.z); f[4] = t.x; f[5] = t.y;
  t = unpack_bf16x2(u.w); f[6] = t.x; f[7] = t.y;
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 u;
  u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]);
  u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
  return u;
}

struct LnArgs {
  const __nv_bfloat16* x;    // [rows, d]
  const __nv_bfloat16* res;  // [rows, d] or null
  const float* gamma;
  const float* beta;
  __nv_bfloat16* y;          // row r at y + (r / rpg) * y_gs + (r % rpg) * d
  float* mean;
  float* rstd;
  long long rows;
  int d;
  long long rpg, y_gs;
  float eps;
  float p_drop;
  const unsigned long long* rng;  // device step counter (null when p_drop == 0)
  uint32_t salt;
  const float* res32;  // [rows, d] fp32 residual (takes precedence over `res`), or null
  float* y32;          // [rows, d] fp32 copy of the output = the next block's residual (contiguous rows), or null
  float* sum32;        // [rows, d] fp32 res + dropout(x) BEFORE normalisation (pre-LN residual stream of the ViT), or null
};

// v[j][e] for lane: vector index (lane + 32*j), element e.
template <int VPL>
__device__ __forceinline__ void ln_row_stats(float (&v)[VPL][8], int d, float eps, float& mean, float& rstd) {
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < VPL; ++j)
#pragma unroll
    for (int e = 0; e < 8; ++e) s += v[j][e];
  mean = warp_sum(s) / d;
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < VPL; ++j)
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float c = v[j][e] - mean;
      q += c * c;
    }
  rstd = rsqrtf(warp_sum(q) / d + eps);
}

template <int VPL>
__global__ void __launch_bounds__(kWarpsPerBlock * 32) add_layernorm_fwd_kernel(const LnArgs a) {
  pdl_sync();
  const int lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * kWarpsPerBlock + (threadIdx.x >> 5);
  if (row >= a.rows) return;
  const uint4* xp = a.x ? reinterpret_cast<const uint4*>(a.x + row * a.d) : nullptr;
  const uint4* rp = a.res ? reinterpret_cast<const uint4*>(a.res + row * a.d) : nullptr;
  const bool drop = a.p_drop > 0.f;
  const uint32_t seed = drop ? static_cast<uint32_t>(*a.rng) : 0u;
  const uint32_t thr = drop_thresh16(a.p_drop);
  const float inv_keep = drop ? 65536.f / static_cast<float>(65536u - thr) : 1.f;
  // Every load of the row is issued before the first use (the branches are warp-uniform): up to 3 x VPL 16-byte loads in
  // flight per lane instead of 3, which is what a one-row-per-warp kernel needs to cover the HBM latency.
  const float4* r4 = a.res32 ? reinterpret_cast<const float4*>(a.res32 + row * a.d) : nullptr;
  uint4 rx[VPL], rr[VPL];
  float4 r0[VPL], r1[VPL];
#pragma unroll
  for (int j = 0; j < VPL; ++j) {
    const int vi = lane + 32 * j;
    rx[j] = xp ? __ldg(xp + vi) : make_uint4(0u, 0u, 0u, 0u);
    if (r4) {
      r0[j] = __ldg(r4 + vi * 2);
      r1[j] = __ldg(r4 + vi * 2 + 1);
    } else {
      rr[j] = rp ? __ldg(rp + vi) : make_uint4(0u, 0u, 0u, 0u);
    }
  }
  const uint32_t keep = drop ? keep_row_bits<VPL>(seed, a.salt, row, a.d, lane, thr) : 0xffffffffu;
  float v[VPL][8];
#pragma unroll
  for (int j = 0; j < VPL; ++j) {
    const int vi = lane + 32 * j;
    unpack8(rx[j], v[j]);
    if (drop) {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[j][e] = ((keep >> (j * 8 + e)) & 1u) ? v[j][e] * inv_keep : 0.f;
    }
    if (r4) {  // residual stream carried in fp32 (what torch.autocast keeps: LayerNorm outputs stay fp32)
      v[j][0] += r0[j].x; v[j][1] += r0[j].y; v[j][2] += r0[j].z; v[j][3] += r0[j].w;
      v[j][4] += r1[j].x; v[j][5] += r1[j].y; v[j][6] += r1[j].z; v[j][7] += r1[j].w;
    } else if (rp) {
      float r[8];
      unpack8(rr[j], r);
#pragma unroll
      for (int e = 0; e < 8; ++e) v[j][e] += r[e];
    }
    if (a.sum32) {
      float4* s4 = reinterpret_cast<float4*>(a.sum32 + row * a.d) + vi * 2;
      s4[0] = make_float4(v[j][0], v[j][1], v[j][2], v[j][3]);
      s4[1] = make_float4(v[j][4], v[j][5], v[j][6], v[j][7]);
    }
  }
  float mean, rstd;
  ln_row_stats<VPL>(v, a.d, a.eps, mean, rstd);
  if (lane == 0) {
    if (a.mean) a.mean[row] = mean;
    if (a.rstd) a.rstd[row] = rstd;
  }
  uint4* yp = reinterpret_cast<uint4*>(a.y + (row / a.rpg) * a.y_gs + (row % a.rpg) * a.d);
  float4* y4 = a.y32 ? reinterpret_cast<float4*>(a.y32 + row * a.d) : nullptr;
#pragma unroll
  for (int j = 0; j < VPL; ++j) {
    const int vi = lane + 32 * j;
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(a.gamma) + vi * 2);
    const float4 g1 = __ldg(reinterpret_cast<const float4*>(a.gamma) + vi * 2 + 1);
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(a.beta) + vi * 2);
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(a.beta) + vi * 2 + 1);
    const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    float o[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = (v[j][e] - mean) * rstd * g[e] + b[e];
    yp[vi] = pack8(o);
    if (y4) {
      y4[vi * 2] = make_float4(o[0], o[1], o[2], o[3]);
      y4[vi * 2 + 1] = make_float4(o[4], o[5], o[6], o[7]);
    }
  }
}

// ViT token assembly + ln_pre of the CLIP image tower (clip.model.VisionTransformer.forward as used by
// extract_clip_img_feat, TRAIN:220-232): row (b, 0) = class_embedding + pos[0]; row (b, i) = patch_tok[b, i-1] + pos[i];
// y32 = LN(row) in fp32 -- the start of the (pre-LN) residual stream.
template <int VPL>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
vit_embed_ln_kernel(const __nv_bfloat16* __restrict__ tok, const float* __restrict__ cls, const float* __restrict__ pos,
                    const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ y32,
                    long long rows, int tokens, int d, float eps) {
  pdl_sync();
  const int lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * kWarpsPerBlock + (threadIdx.x >> 5);
  if (row >= rows) return;
  const long long b = row / tokens;
  const int i = static_cast<int>(row % tokens);
  float v[VPL][8];
#pragma unroll
  for (int j = 0; j < VPL; ++j) {
    const int vi = lane + 32 * j;
    const float4* p4 = reinterpret_cast<const float4*>(pos + static_cast<long long>(i) * d) + vi * 2;
    const float4 p0 = __ldg(p4), p1 = __ldg(p4 + 1);
    if (i == 0) {
      const float4* c4 = reinterpret_cast<const float4*>(cls) + vi * 2;
      const float4 c0 = __ldg(c4), c1 = __ldg(c4 + 1);
      v[j][0] = c0.x; v[j][1] = c0.y; v[j][2] = c0.z; v[j][3] = c0.w; v[j][4] = c1.x; v[j][5] = c1.y; v[j][6] = c1.z; v[j][7] = c1.w;
    } else {
      unpack8(__ldg(reinterpret_cast<const uint4*>(tok + (b * (tokens - 1) + i - 1) * d) + vi), v[j]);
    }
    v[j][0] += p0.x; v[j][1] += p0.y; v[j][2] += p0.z; v[j][3] += p0.w;
    v[j][4] += p1.x; v[j][5] += p1.y; v[j][6] += p1.z; v[j][7] += p1.w;
  }
  float mean, rstd;
  ln_row_stats<VPL>(v, d, eps, mean, rstd);
  float4* y4 = reinterpret_cast<float4*>(y32 + row * d);
#pragma unroll
  for (int j = 0; j < VPL; ++j) {
    const int vi = lane + 32 * j;
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma) + vi * 2), g1 = __ldg(reinterpret_cast<const float4*>(gamma) + vi * 2 + 1);
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta) + vi * 2), b1 = __ldg(reinterpret_cast<const float4*>(beta) + vi * 2 + 1);
    y4[vi * 2] = make_float4((v[j][0] - mean) * rstd * g0.x + b0.x, (v[j][1] - mean) * rstd * g0.y + b0.y,
                             (v[j][2] - mean) * rstd * g0.z + b0.z, (v[j][3] - mean) * rstd * g0.w + b0.w);
    y4[vi * 2 + 1] = make_float4((v[j][4] - mean) * rstd * g1.x + b1.x, (v[j][5] - mean) * rstd * g1.y + b1.y,
                                 (v[j][6] - mean) * rstd * g1.z + b1.z, (v[j][7] - mean) * rstd * g1.w + b1.w);
  }
}

struct LnBwdArgs {
  const __nv_bfloat16* dy;   // row r at dy + (r / rpg) * dy_gs + (r % rpg) * d
  const __nv_bfloat16* x;
  const __nv_bfloat16* res;  // or null
  const float* gamma;
  const float* mean;
  const float* rstd;
  __nv_bfloat16* dsum;       // [rows,d] gradient of (res + dropout(x)); null allowed when dx given
  __nv_bfloat16* dx;         // [rows,d] gradient of x (= dsum when no dropout); may alias / be null
  float* dgamma;             // += (atomic)
  float* dbeta;              // += (atomic)
  float* dbias;              // += column sums of dx (bias of the linear that produced x), or null
  long long rows;
  int d;
  long long rpg, dy_gs;
  float p_drop;
  const unsigned long long* rng;
  uint32_t salt;
  int accumulate_dsum;       // dsum += instead of = (residual stream fan-in)
};

// The parameter-gradient partial sums (d gamma, d beta, d bias: 3 x d fp32 per warp) live in SHARED memory, one private
// slice per warp (lane l only ever touches its own columns, so no atomics and no barriers inside the row loop).  In
// registers they cost 96 of ~250 registers and held the kernel at one 8-warp CTA per SM; in shared memory the kernel needs
// < 128 registers and two CTAs are resident, i.e. twice the rows -- and bytes -- in flight, which is what bounds it.
template <int VPL>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, 2) add_layernorm_bwd_kernel(const LnBwdArgs a) {
  extern __shared__ __align__(16) float ln_acc[];  // [warp][3][D]
  constexpr int D = VPL * 256;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  float* acc = ln_acc + warp * 3 * D;
  for (int i = lane; i < 3 * D / 4; i += 32) reinterpret_cast<float4*>(acc)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncwarp();
  pdl_sync();
  const bool drop = a.p_drop > 0.f;
  const uint32_t seed = drop ? static_cast<uint32_t>(*a.rng) : 0u;
  const uint32_t thr = drop_thresh16(a.p_drop);
  const float inv_keep = drop ? 65536.f / static_cast<float>(65536u - thr) : 1.f;
  const bool want_dbias = a.dbias != nullptr;
  auto accumulate = [&](int which, int j, const float (&t)[8]) {  // acc[which][(lane + 32 j) * 8 + e] += t[e]
    float4* p = reinterpret_cast<float4*>(acc + which * D) + (lane + 32 * j) * 2;
    float4 u0 = p[0], u1 = p[1];
    u0.x += t[0]; u0.y += t[1]; u0.z += t[2]; u0.w += t[3];
    u1.x += t[4]; u1.y += t[5]; u1.z += t[6]; u1.w += t[7];
    p[0] = u0; p[1] = u1;
  };

  const long long wstride = static_cast<long long>(gridDim.x) * kWarpsPerBlock;
  for (long long row = static_cast<long long>(blockIdx.x) * kWarpsPerBlock + warp; row < a.rows; row += wstride) {
    const uint4* xp = reinterpret_cast<const uint4*>(a.x + row * a.d);
    const uint4* rp = a.res ? reinterpret_cast<const uint4*>(a.res + row * a.d) : nullptr;
    const uint4* dyp = reinterpret_cast<const uint4*>(a.dy + (row / a.rpg) * a.dy_gs + (row % a.rpg) * a.d);
    // Every load of the row is issued before the first use (3 x VPL 16-byte loads in flight per lane): with ~8 resident
    // warps per SM the kernel lives on memory-level parallelism.  The raw bf16 words stay in registers and x_hat / g are
    // recomputed in the second pass instead of being kept as 2 x 32 fp32 values across the row reduction.
    uint4 rx[VPL], rdy[VPL], rr[VPL];
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
      const int vi = lane + 32 * j;
      rx[j] = __ldg(xp + vi);
      rdy[j] = __ldg(dyp + vi);
      rr[j] = rp ? __ldg(rp + vi) : make_uint4(0u, 0u, 0u, 0u);
    }
    const float mean = a.mean[row], rstd = a.rstd[row];
    const uint32_t keep = drop ? keep_row_bits<VPL>(seed, a.salt, row, a.d, lane, thr) : 0xffffffffu;  // bit j*8+e
    auto xhat_g = [&](int j, float (&xh)[8], float (&g)[8], float (&dyv)[8]) {
      const int vi = lane + 32 * j;
      float v[8], r[8];
      unpack8(rx[j], v);
      unpack8(rdy[j], dyv);
      unpack8(rr[j], r);
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(a.gamma) + vi * 2);
      const float4 g1 = __ldg(reinterpret_cast<const float4*>(a.gamma) + vi * 2 + 1);
      const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float xv = ((keep >> (j * 8 + e)) & 1u) ? v[e] * inv_keep : 0.f;
        xh[e] = (xv + r[e] - mean) * rstd;
        g[e] = dyv[e] * gm[e];
      }
    };
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
      float xh[8], g[8], dyv[8];
      xhat_g(j, xh, g, dyv);
      float t[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        s1 += g[e];
        s2 += g[e] * xh[e];
        t[e] = dyv[e] * xh[e];
      }
      accumulate(0, j, t);
      accumulate(1, j, dyv);
    }
    s1 = warp_sum(s1) / a.d;
    s2 = warp_sum(s2) / a.d;
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
      const int vi = lane + 32 * j;
      float xh[8], g[8], dyv[8], ds[8], dxv[8];
      xhat_g(j, xh, g, dyv);
#pragma unroll
      for (int e = 0; e < 8; ++e) ds[e] = rstd * (g[e] - s1 - xh[e] * s2);
      if (a.dsum) {
        uint4* p = reinterpret_cast<uint4*>(a.dsum + row * a.d) + vi;
        if (a.accumulate_dsum) {
          float old[8];
          unpack8(*p, old);
#pragma unroll
          for (int e = 0; e < 8; ++e) ds[e] += old[e];
        }
        *p = pack8(ds);
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) dxv[e] = ((keep >> (j * 8 + e)) & 1u) ? ds[e] * inv_keep : 0.f;
      if (want_dbias) accumulate(2, j, dxv);
      if (a.dx && a.dx != a.dsum) reinterpret_cast<uint4*>(a.dx + row * a.d)[vi] = pack8(dxv);
    }
  }
  // block-level reduction of the parameter gradients, then one atomic per column per block
  __syncthreads();
  for (int which = 0; which < 3; ++which) {
    float* dst = which == 0 ? a.dgamma : (which == 1 ? a.dbeta : a.dbias);
    if (dst == nullptr) continue;
    for (int c = threadIdx.x; c < D; c += kWarpsPerBlock * 32) {
      float sum = 0.f;
#pragma unroll
      for (int w = 0; w < kWarpsPerBlock; ++w) sum += ln_acc[(w * 3 + which) * D + c];
      atomicAdd(dst + c, sum);
    }
  }
}

// ------------------------------------------------------------------ embeddings
struct EmbArgs {
  const long long* ids;        // [rows]
  const __nv_bfloat16* tok;    // [V, d]
  const __nv_bfloat16* pos;    // [max_pos + 2, d]
  const float* gamma;
  const float* beta;
  __nv_bfloat16* y;            // [rows, d]
  float* mean;
  float* rstd;
  long long rows;
  int seq_len, pos_offset, d;
  float eps, p_drop;
  const unsigned long long* rng;
  uint32_t salt;
  float* y32;              // [rows, d] fp32 copy of the output, or null
  const int* pos_ids;      // [rows] explicit position (before pos_offset) of every row, or null = row % seq_len
};

template <int VPL>
__global__ void __launch_bounds__(kWarpsPerBlock * 32) embed_ln_fwd_kernel(const EmbArgs a) {
  const int lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * kWarpsPerBlock + (threadIdx.x >> 5);
  if (row >= a.rows) return;
  const long long id = a.ids[row];
  const int p = (a.pos_ids ? a.pos_ids[row] : static_cast<int>(row % a.seq_len)) + a.pos_offset;
  const uint4* tp = reinterpret_cast<const uint4*>(a.tok + id * a.d);
  const uint4* pp = reinterpret_cast<const uint4*>(a.pos + static_cast<long long>(p) * a.d);
  float v[VPL][8];
#pragma unroll
  for (int j = 0; j < VPL; ++j) {
    float t[8], q[8];
    unpack8(__ldg(tp + lane + 32 * j), t);
    unpack8(__ldg(pp + lane + 32 * j), q);
#pragma unroll
    for (int e = 0; e < 8; ++e) v[j][e] = t[e] + q[e];
  }
  float mean, rstd;
  ln_row_stats<VPL>(v, a.d, a.eps, mean, rstd);
  if (lane == 0) {
    if (a.mean) a.mean[row] = mean;
    if (a.rstd) a.rstd[row] = rstd;
  }
  const bool drop = a.p_drop > 0.f;
  const uint32_t seed = drop ? static_cast<uint32_t>(*a.rng) : 0u;
  const uint32_t thr = drop_thresh(a.p_drop);
  const float inv_keep = drop ? 1.f / (1.f - a.p_drop) : 1.f;
  uint4* yp = reinterpret_cast<uint4*>(a.y + row * a.d);
  float4* y4 = a.y32 ? reinterpret_cast<float4*>(a.y32 + row * a.d) : nullptr;
#pragma unroll
  for (int j = 0; j < VPL; ++j) {
    const int vi = lane + 32 * j;
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(a.gamma) + vi * 2);
    const float4 g1 = __ldg(reinterpret_cast<const float4*>(a.gamma) + vi * 2 + 1);
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(a.beta) + vi * 2);
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(a.beta) + vi * 2 + 1);
    const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    float o[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      o[e] = (v[j][e] - mean) * rstd * g[e] + b[e];
      if (drop) o[e] = keep_elem(seed, a.salt, static_cast<uint64_t>(row) * a.d + vi * 8 + e, thr) ? o[e] * inv_keep : 0.f;
    }
    yp[vi] = pack8(o);
    if (y4) {
      y4[vi * 2] = make_float4(o[0], o[1], o[2], o[3]);
      y4[vi * 2 + 1] = make_float4(o[4], o[5], o[6], o[7]);
    }
  }
}

struct EmbBwdArgs {
  const __nv_bfloat16* dy;
  const long long* ids;
  const __nv_bfloat16* tok;
  const __nv_bfloat16* pos;
  const float* gamma;
  const float* mean;
  const float* rstd;
  float* dtok;   // [V, d]  += (atomic), skipped for ids == pad_id (nn.Embedding padding_idx)
  float* dpos;   // [max_pos + 2, d] += (atomic)
  float* dgamma;
  float* dbeta;
  long long rows;
  int seq_len, pos_offset, d, pad_id;
  float p_drop;
  const unsigned long long* rng;
  uint32_t salt;
  const int* pos_ids;
};

template <int VPL>
__global__ void __launch_bounds__(kWarpsPerBlock * 32) embed_ln_bwd_kernel(const EmbBwdArgs a) {
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const bool drop = a.p_drop > 0.f;
  const uint32_t seed = drop ? static_cast<uint32_t>(*a.rng) : 0u;
  const uint32_t thr = drop_thresh(a.p_drop);
  const float inv_keep = drop ? 1.f / (1.f - a.p_drop) : 1.f;
  float dg[VPL][8], db[VPL][8];
#pragma unroll
  for (int j = 0; j < VPL; ++j)
#pragma unroll
    for (int e = 0; e < 8; ++e) dg[j][e] = db[j][e] = 0.f;
  const long long wstride = static_cast<long long>(gridDim.x) * kWarpsPerBlock;
  for (long long row = static_cast<long long>(blockIdx.x) * kWarpsPerBlock + warp; row < a.rows; row += wstride) {
    const long long id = a.ids[row];
    const int p = (a.pos_ids ? a.pos_ids[row] : static_cast<int>(row % a.seq_len)) + a.pos_offset;
    const uint4* tp = reinterpret_cast<const uint4*>(a.tok + id * a.d);
    const uint4* pp = reinterpret_cast<const uint4*>(a.pos + static_cast<long long>(p) * a.d);
    const uint4* dyp = reinterpret_cast<const uint4*>(a.dy + row * a.d);
    const float mean = a.mean[row], rstd = a.rstd[row];
    float xh[VPL][8], g[VPL][8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
      const int vi = lane + 32 * j;
      float t[8], q[8], dyv[8];
      unpack8(__ldg(tp + vi), t);
      unpack8(__ldg(pp + vi), q);
      unpack8(__ldg(dyp + vi), dyv);
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(a.gamma) + vi * 2);
      const float4 g1 = __ldg(reinterpret_cast<const float4*>(a.gamma) + vi * 2 + 1);
      const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        if (drop) dyv[e] = keep_elem(seed, a.salt, static_cast<uint64_t>(row) * a.d + vi * 8 + e, thr) ? dyv[e] * inv_keep : 0.f;
        xh[j][e] = (t[e] + q[e] - mean) * rstd;
        g[j][e] = dyv[e] * gm[e];
        s1 += g[j][e];
        s2 += g[j][e] * xh[j][e];
        dg[j][e] += dyv[e] * xh[j][e];
        db[j][e] += dyv[e];
      }
    }
    s1 = warp_sum(s1) / a.d;
    s2 = warp_sum(s2) / a.d;
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
      const int vi = lane + 32 * j;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float dh = rstd * (g[j][e] - s1 - xh[j][e] * s2);
        const int col = vi * 8 + e;
        if (a.dtok && id != a.pad_id) atomicAdd(a.dtok + id * a.d + col, dh);
        if (a.dpos) atomicAdd(a.dpos + static_cast<long long>(p) * a.d + col, dh);
      }
    }
  }
  __shared__ float red[kWarpsPerBlock][256];
  for (int which = 0; which < 2; ++which) {
    float* dst = which == 0 ? a.dgamma : a.dbeta;
    if (dst == nullptr) continue;
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
      __syncthreads();
#pragma unroll
      for (int e = 0; e < 8; ++e) red[warp][lane * 8 + e] = which == 0 ? dg[j][e] : db[j][e];
      __syncthreads();
      for (int c = threadIdx.x; c < 256; c += kWarpsPerBlock * 32) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < kWarpsPerBlock; ++w) s += red[w][c];
        atomicAdd(dst + j * 256 + c, s);
      }
    }
  }
}

// out[span, :] = mean_t LN(E[ids[span, t]] + Pos[t + 2]) ; one warp per span, fp32 output.
template <int VPL>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
names_embed_kernel(const long long* __restrict__ ids, const __nv_bfloat16* __restrict__ tok,
                   const __nv_bfloat16* __restrict__ pos, const float* __restrict__ gamma,
                   const float* __restrict__ beta, float* __restrict__ out, long long spans, int len, int d,
                   float eps) {
  const int lane = threadIdx.x & 31;
  const long long span = static_cast<long long>(blockIdx.x) * kWarpsPerBlock + (threadIdx.x >> 5);
  if (span >= spans) return;
  float acc[VPL][8];
#pragma unroll
  for (int j = 0; j < VPL; ++j)
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[j][e] = 0.f;
  for (int t = 0; t < len; ++t) {
    const long long id = ids[span * len + t];
    const uint4* tp = reinterpret_cast<const uint4*>(tok + id * d);
    const uint4* pp = reinterpret_cast<const uint4*>(pos + static_cast<long long>(t + 2) * d);
    float v[VPL][8];
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
      float a[8], b[8];
      unpack8(__ldg(tp + lane + 32 * j), a);
      unpack8(__ldg(pp + lane + 32 * j), b);
#pragma unroll
      for (int e = 0; e < 8; ++e) v[j][e] = a[e] + b[e];
    }
    float mean, rstd;
    ln_row_stats<VPL>(v, d, eps, mean, rstd);
#pragma unroll
    for (int j = 0; j < VPL; ++j)
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[j][e] += (v[j][e] - mean) * rstd;
  }
  const float inv = 1.f / len;
#pragma unroll
  for (int j = 0; j < VPL; ++j) {
    const int vi = lane + 32 * j;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int col = vi * 8 + e;
      // mean_t (xhat * g + b) = g * mean_t xhat + b
      out[span * d + col] = acc[j][e] * inv * __ldg(gamma + col) + __ldg(beta + col);
    }
  }
}

static int bwd_grid(long long rows, int ctas_per_sm) {
  // the backward kernels are persistent over rows: exactly the resident CTAs are launched -- a single wave, and the fewest
  // end-of-CTA atomics on the parameter gradients
  const int sms = sm_count();
  long long want = (rows + kWarpsPerBlock - 1) / kWarpsPerBlock;
  long long cap = static_cast<long long>(sms > 0 ? sms : 148) * ctas_per_sm;
  return static_cast<int>(want < cap ? (want > 0 ? want : 1) : cap);
}

#define VB_DISPATCH_VPL(d, CALL)                                              \
  switch ((d) / 256) {                                                        \
    case 1: { constexpr int VPL = 1; CALL; } break;                           \
    case 2: { constexpr int VPL = 2; CALL; } break;                           \
    case 3: { constexpr int VPL = 3; CALL; } break;                           \
    case 4: { constexpr int VPL = 4; CALL; } break;                           \
    default: return fail(VACNIC_EINVAL, "d_model %d not supported (need 256..1024, multiple of 256)", (d)); \
  }

}  // namespace vb

using namespace vb;

extern "C" int vacnic_add_layernorm_fwd(const void* x, const void* res, const float* gamma, const float* beta,
                                        void* y, float* mean, float* rstd, int64_t rows, int32_t d,
                                        int64_t rows_per_group, int64_t y_group_stride, float eps, float p_drop,
                                        const uint64_t* rng_state, uint32_t salt, const float* res32, float* y32,
                                        float* sum32, void* stream) {
  VB_REQUIRE((x || res32) && gamma && beta && y, "add_layernorm_fwd: null pointer");
  VB_REQUIRE(((reinterpret_cast<uintptr_t>(res32) | reinterpret_cast<uintptr_t>(y32) | reinterpret_cast<uintptr_t>(sum32)) & 15) == 0,
             "add_layernorm_fwd: fp32 residual buffers must be 16-byte aligned");
  VB_REQUIRE(rows >= 0 && d > 0 && d % 256 == 0, "add_layernorm_fwd: bad shape rows=%lld d=%d", (long long)rows, d);
  VB_REQUIRE(p_drop >= 0.f && p_drop < 1.f && (p_drop == 0.f || rng_state), "add_layernorm_fwd: bad dropout args");
  if (rows == 0) return VACNIC_OK;
  LnArgs a;
  a.x = static_cast<const __nv_bfloat16*>(x); a.res = static_cast<const __nv_bfloat16*>(res);
  a.gamma = gamma; a.beta = beta; a.y = static_cast<__nv_bfloat16*>(y); a.mean = mean; a.rstd = rstd;
  a.rows = rows; a.d = d;
  a.rpg = rows_per_group > 0 ? rows_per_group : rows;
  a.y_gs = rows_per_group > 0 ? y_group_stride : 0;
  VB_REQUIRE(a.y_gs % 8 == 0, "add_layernorm_fwd: group stride must be a multiple of 8 elements");
  a.eps = eps; a.p_drop = p_drop; a.rng = reinterpret_cast<const unsigned long long*>(rng_state); a.salt = salt;
  a.res32 = res32; a.y32 = y32; a.sum32 = sum32;
  const int grid = static_cast<int>((rows + kWarpsPerBlock - 1) / kWarpsPerBlock);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  VB_DISPATCH_VPL(d, (launch_pdl(add_layernorm_fwd_kernel<VPL>, dim3(grid), dim3(kWarpsPerBlock * 32), 0, s, a)));
  count_launch();
  return check_last("add_layernorm_fwd");
}

extern "C" int vacnic_add_layernorm_bwd(const void* dy, const void* x, const void* res, const float* gamma,
                                        const float* mean, const float* rstd, void* dsum, void* dx, float* dgamma,
                                        float* dbeta, float* dbias, int64_t rows, int32_t d, int64_t rows_per_group,
                                        int64_t dy_group_stride, float p_drop, const uint64_t* rng_state,
                                        uint32_t salt, int32_t accumulate_dsum, void* stream) {
  VB_REQUIRE(dy && x && gamma && mean && rstd, "add_layernorm_bwd: null pointer");
  VB_REQUIRE(dsum || dx, "add_layernorm_bwd: need dsum or dx");
  VB_REQUIRE(rows >= 0 && d > 0 && d % 256 == 0, "add_layernorm_bwd: bad shape");
  VB_REQUIRE(p_drop >= 0.f && p_drop < 1.f && (p_drop == 0.f || rng_state), "add_layernorm_bwd: bad dropout args");
  if (rows == 0) return VACNIC_OK;
  LnBwdArgs a;
  a.dy = static_cast<const __nv_bfloat16*>(dy); a.x = static_cast<const __nv_bfloat16*>(x);
  a.res = static_cast<const __nv_bfloat16*>(res); a.gamma = gamma; a.mean = mean; a.rstd = rstd;
  a.dsum = static_cast<__nv_bfloat16*>(dsum); a.dx = static_cast<__nv_bfloat16*>(dx);
  a.dgamma = dgamma; a.dbeta = dbeta; a.dbias = dbias; a.rows = rows; a.d = d;
  a.rpg = rows_per_group > 0 ? rows_per_group : rows;
  a.dy_gs = rows_per_group > 0 ? dy_group_stride : 0;
  VB_REQUIRE(a.dy_gs % 8 == 0, "add_layernorm_bwd: group stride must be a multiple of 8 elements");
  a.p_drop = p_drop; a.rng = reinterpret_cast<const unsigned long long*>(rng_state); a.salt = salt;
  a.accumulate_dsum = accumulate_dsum;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int grid = bwd_grid(rows, 2);
  const size_t smem = static_cast<size_t>(kWarpsPerBlock) * 3 * d * sizeof(float);  // 96 KB at d = 1024: two CTAs per SM
  VB_DISPATCH_VPL(d, {
    static bool configured = false;
    if (!configured) {
      cudaError_t e = cudaFuncSetAttribute(add_layernorm_bwd_kernel<VPL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           kWarpsPerBlock * 3 * VPL * 256 * static_cast<int>(sizeof(float)));
      if (e != cudaSuccess) return fail(VACNIC_ECUDA, "add_layernorm_bwd: smem attribute: %s", cudaGetErrorString(e));
      configured = true;
    }
    launch_pdl(add_layernorm_bwd_kernel<VPL>, dim3(grid), dim3(kWarpsPerBlock * 32), smem, s, a);
  });
  count_launch();
  return check_last("add_layernorm_bwd");
}

extern "C" int vacnic_embed_ln_fwd(const int64_t* ids, const void* tok, const void* pos, const float* gamma,
                                   const float* beta, void* y, float* mean, float* rstd, int64_t rows,
                                   int32_t seq_len, int32_t pos_offset, int32_t d, float eps, float p_drop,
                                   const uint64_t* rng_state, uint32_t salt, float* y32, const int32_t* pos_ids,
                                   void* stream) {
  VB_REQUIRE(ids && tok && pos && gamma && beta && y, "embed_ln_fwd: null pointer");
  VB_REQUIRE(rows >= 0 && seq_len > 0 && d > 0 && d % 256 == 0, "embed_ln_fwd: bad shape");
  VB_REQUIRE(p_drop >= 0.f && p_drop < 1.f && (p_drop == 0.f || rng_state), "embed_ln_fwd: bad dropout args");
  if (rows == 0) return VACNIC_OK;
  EmbArgs a;
  a.ids = reinterpret_cast<const long long*>(ids); a.tok = static_cast<const __nv_bfloat16*>(tok);
  a.pos = static_cast<const __nv_bfloat16*>(pos); a.gamma = gamma; a.beta = beta;
  a.y = static_cast<__nv_bfloat16*>(y); a.mean = mean; a.rstd = rstd; a.rows = rows; a.seq_len = seq_len;
  a.pos_offset = pos_offset; a.d = d; a.eps = eps; a.p_drop = p_drop;
  a.rng = reinterpret_cast<const unsigned long long*>(rng_state); a.salt = salt;
  a.y32 = y32; a.pos_ids = pos_ids;
  const int grid = static_cast<int>((rows + kWarpsPerBlock - 1) / kWarpsPerBlock);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  VB_DISPATCH_VPL(d, (embed_ln_fwd_kernel<VPL><<<grid, kWarpsPerBlock * 32, 0, s>>>(a)));
  count_launch();
  return check_last("embed_ln_fwd");
}

extern "C" int vacnic_embed_ln_bwd(const void* dy, const int64_t* ids, const void* tok, const void* pos,
                                   const float* gamma, const float* mean, const float* rstd, float* dtok,
                                   float* dpos, float* dgamma, float* dbeta, int64_t rows, int32_t seq_len,
                                   int32_t pos_offset, int32_t d, int32_t pad_id, float p_drop,
                                   const uint64_t* rng_state, uint32_t salt, const int32_t* pos_ids, void* stream) {
  VB_REQUIRE(dy && ids && tok && pos && gamma && mean && rstd, "embed_ln_bwd: null pointer");
  VB_REQUIRE(rows >= 0 && seq_len > 0 && d > 0 && d % 256 == 0, "embed_ln_bwd: bad shape");
  if (rows == 0) return VACNIC_OK;
  EmbBwdArgs a;
  a.dy = static_cast<const __nv_bfloat16*>(dy); a.ids = reinterpret_cast<const long long*>(ids);
  a.tok = static_cast<const __nv_bfloat16*>(tok); a.pos = static_cast<const __nv_bfloat16*>(pos);
  a.gamma = gamma; a.mean = mean; a.rstd = rstd; a.dtok = dtok; a.dpos = dpos; a.dgamma = dgamma; a.dbeta = dbeta;
  a.rows = rows; a.seq_len = seq_len; a.pos_offset = pos_offset; a.d = d; a.pad_id = pad_id; a.p_drop = p_drop;
  a.rng = reinterpret_cast<const unsigned long long*>(rng_state); a.salt = salt;
  a.pos_ids = pos_ids;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int grid = bwd_grid(rows, 1);
  VB_DISPATCH_VPL(d, (embed_ln_bwd_kernel<VPL><<<grid, kWarpsPerBlock * 32, 0, s>>>(a)));
  count_launch();
  return check_last("embed_ln_bwd");
}

extern "C" int vacnic_names_embed(const int64_t* ids, const void* tok, const void* pos, const float* gamma,
                                  const float* beta, float* out, int64_t spans, int32_t len, int32_t d, float eps,
                                  void* stream) {
  VB_REQUIRE(ids && tok && pos && gamma && beta && out, "names_embed: null pointer");
  VB_REQUIRE(spans >= 0 && len > 0 && d > 0 && d % 256 == 0, "names_embed: bad shape");
  if (spans == 0) return VACNIC_OK;
  const int grid = static_cast<int>((spans + kWarpsPerBlock - 1) / kWarpsPerBlock);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  VB_DISPATCH_VPL(d, (names_embed_kernel<VPL><<<grid, kWarpsPerBlock * 32, 0, s>>>(
                         reinterpret_cast<const long long*>(ids), static_cast<const __nv_bfloat16*>(tok),
                         static_cast<const __nv_bfloat16*>(pos), gamma, beta, out, spans, len, d, eps)));
  count_launch();
  return check_last("names_embed");
}

extern "C" int vacnic_vit_embed_ln(const void* tok, const float* cls, const float* pos, const float* gamma, const float* beta,
                                   float* y32, int64_t batch, int32_t tokens, int32_t d, float eps, void* stream) {
  VB_REQUIRE(tok && cls && pos && gamma && beta && y32, "vit_embed_ln: null pointer");
  VB_REQUIRE(batch >= 0 && tokens >= 2 && d > 0 && d % 256 == 0, "vit_embed_ln: bad shape");
  if (batch == 0) return VACNIC_OK;
  const long long rows = batch * tokens;
  const int grid = static_cast<int>((rows + kWarpsPerBlock - 1) / kWarpsPerBlock);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  VB_DISPATCH_VPL(d, (launch_pdl(vit_embed_ln_kernel<VPL>, dim3(grid), dim3(kWarpsPerBlock * 32), 0, s,
                                 static_cast<const __nv_bfloat16*>(tok), cls, pos, gamma, beta, y32, rows, tokens, d, eps)));
  count_launch();
  return check_last("vit_embed_ln");
}
