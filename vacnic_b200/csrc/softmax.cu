// Masked row softmax and its gradient for the unfused attention path (BartAttention.forward,
// MFULL:509-548): scores arrive as fp32 [B, H, Sq, ld] from the QK^T GEMM (already scaled, the
// reference scales q before the product, MFULL:472), the additive mask of _expand_mask /
// _make_causal_mask (MFULL:373-398) is applied as "+ finfo(float32).min" exactly like the reference
// (so a fully masked row degenerates to a uniform row, as it does there), probabilities leave as bf16.
// One warp per row; HBM-bound.
#include <float.h>

#include "common.h"
#include "ptx.cuh"

namespace vb {

constexpr int kSmWarps = 8;

// Each lane owns float4 chunks: columns 4*(lane + 32*j) .. +3 (ld is a multiple of 8, rows 16-byte aligned).
template <int VPL>
__global__ void __launch_bounds__(kSmWarps * 32)
softmax_fwd_kernel(const float* __restrict__ s, __nv_bfloat16* __restrict__ p, const uint8_t* __restrict__ key_mask,
                   long long rows, int Sq, int Sk, int ld, int rows_per_batch, int causal, int past) {
  const int lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * kSmWarps + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float4* sp = reinterpret_cast<const float4*>(s + row * ld);
  const uint8_t* mp = key_mask ? key_mask + (row / rows_per_batch) * Sk : nullptr;
  const int qi = static_cast<int>(row % Sq);
  const int lim = causal ? qi + past : 0x7fffffff;
  float v[VPL][4];
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < VPL; ++j) {
    const int c0 = 4 * (lane + 32 * j);
    float4 x = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    if (c0 < ld) x = __ldg(sp + lane + 32 * j);
    float* xe = &x.x;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int c = c0 + e;
      float t = -INFINITY;
      if (c < Sk) {
        t = xe[e];
        if (mp && !mp[c]) t += -FLT_MAX;
        if (c > lim) t += -FLT_MAX;
      }
      v[j][e] = t;
      mx = fmaxf(mx, t);
    }
  }
  mx = warp_max(mx);
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < VPL; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      v[j][e] = __expf(v[j][e] - mx);  // exp(-inf) = 0 for the pad columns
      sum += v[j][e];
    }
  const float inv = 1.f / warp_sum(sum);
  uint2* pp = reinterpret_cast<uint2*>(p + row * ld);
#pragma unroll
  for (int j = 0; j < VPL; ++j) {
    const int c0 = 4 * (lane + 32 * j);
    if (c0 < ld) {
      uint2 u;
      u.x = pack_bf16x2(v[j][0] * inv, v[j][1] * inv);
      u.y = pack_bf16x2(v[j][2] * inv, v[j][3] * inv);
      pp[lane + 32 * j] = u;
    }
  }
}

template <int VPL>
__global__ void __launch_bounds__(kSmWarps * 32)
softmax_bwd_kernel(const __nv_bfloat16* __restrict__ p, const float* __restrict__ dp, __nv_bfloat16* __restrict__ ds,
                   long long rows, int Sk, int ld) {
  const int lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * kSmWarps + (threadIdx.x >> 5);
  if (row >= rows) return;
  const uint2* pp = reinterpret_cast<const uint2*>(p + row * ld);
  const float4* gp = reinterpret_cast<const float4*>(dp + row * ld);
  float pv[VPL][4], g[VPL][4];
  float dot = 0.f;
#pragma unroll
  for (int j = 0; j < VPL; ++j) {
    const int c0 = 4 * (lane + 32 * j);
    uint2 u = make_uint2(0, 0);
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c0 < ld) { u = __ldg(pp + lane + 32 * j); x = __ldg(gp + lane + 32 * j); }
    const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
    pv[j][0] = a.x; pv[j][1] = a.y; pv[j][2] = b.x; pv[j][3] = b.y;
    g[j][0] = x.x; g[j][1] = x.y; g[j][2] = x.z; g[j][3] = x.w;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (c0 + e >= Sk) { pv[j][e] = 0.f; g[j][e] = 0.f; }
      dot += pv[j][e] * g[j][e];
    }
  }
  dot = warp_sum(dot);
  uint2* op = reinterpret_cast<uint2*>(ds + row * ld);
#pragma unroll
  for (int j = 0; j < VPL; ++j) {
    const int c0 = 4 * (lane + 32 * j);
    if (c0 < ld) {
      uint2 u;
      u.x = pack_bf16x2(pv[j][0] * (g[j][0] - dot), pv[j][1] * (g[j][1] - dot));
      u.y = pack_bf16x2(pv[j][2] * (g[j][2] - dot), pv[j][3] * (g[j][3] - dot));
      op[lane + 32 * j] = u;
    }
  }
}

#define VB_DISPATCH_EPL(n, CALL)                                  \
  do {                                                            \
    const int vpl_ = ((n) + 127) / 128;                           \
    if (vpl_ <= 1) { constexpr int VPL = 1; CALL; }               \
    else if (vpl_ <= 2) { constexpr int VPL = 2; CALL; }          \
    else if (vpl_ <= 4) { constexpr int VPL = 4; CALL; }          \
    else if (vpl_ <= 8) { constexpr int VPL = 8; CALL; }          \
    else if (vpl_ <= 9) { constexpr int VPL = 9; CALL; }          \
    else return fail(VACNIC_EINVAL, "softmax: row length %d > 1152 not supported", (n)); \
  } while (0)

}  // namespace vb

using namespace vb;

extern "C" int vacnic_softmax_fwd(const float* scores, void* probs, const uint8_t* key_mask, int32_t B, int32_t H,
                                  int32_t Sq, int32_t Sk, int32_t ld, int32_t causal, int32_t past, void* stream) {
  VB_REQUIRE(scores && probs, "softmax_fwd: null pointer");
  VB_REQUIRE(B > 0 && H > 0 && Sq > 0 && Sk > 0 && ld >= Sk && ld % 8 == 0, "softmax_fwd: bad shape (ld must be a multiple of 8)");
  VB_REQUIRE((reinterpret_cast<uintptr_t>(scores) & 15) == 0 && (reinterpret_cast<uintptr_t>(probs) & 7) == 0, "softmax_fwd: misaligned");
  const long long rows = static_cast<long long>(B) * H * Sq;
  const int grid = static_cast<int>((rows + kSmWarps - 1) / kSmWarps);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  VB_DISPATCH_EPL(ld, (softmax_fwd_kernel<VPL><<<grid, kSmWarps * 32, 0, s>>>(
                          scores, static_cast<__nv_bfloat16*>(probs), key_mask, rows, Sq, Sk, ld, H * Sq, causal, past)));
  count_launch();
  return check_last("softmax_fwd");
}

extern "C" int vacnic_softmax_bwd(const void* probs, const float* dprobs, void* dscores, int64_t rows, int32_t Sk,
                                  int32_t ld, void* stream) {
  VB_REQUIRE(probs && dprobs && dscores, "softmax_bwd: null pointer");
  VB_REQUIRE(rows > 0 && Sk > 0 && ld >= Sk && ld % 8 == 0, "softmax_bwd: bad shape (ld must be a multiple of 8)");
  const int grid = static_cast<int>((rows + kSmWarps - 1) / kSmWarps);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  VB_DISPATCH_EPL(ld, (softmax_bwd_kernel<VPL><<<grid, kSmWarps * 32, 0, s>>>(
                          static_cast<const __nv_bfloat16*>(probs), dprobs, static_cast<__nv_bfloat16*>(dscores), rows,
                          Sk, ld)));
  count_launch();
  return check_last("softmax_bwd");
}
