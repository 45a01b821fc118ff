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

template <int EPL>
__global__ void __launch_bounds__(kSmWarps * 32)
softmax_fwd_kernel(const float* __restrict__ s, __nv_bfloat16* __restrict__ p, const uint8_t* __restrict__ key_mask,
                   long long rows, int Sq, int Sk, int ld, int rows_per_batch, int causal, int past) {
  const int lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * kSmWarps + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* sp = s + row * ld;
  const uint8_t* mp = key_mask ? key_mask + (row / rows_per_batch) * Sk : nullptr;
  const int qi = static_cast<int>(row % Sq);
  float v[EPL];
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < EPL; ++j) {
    const int c = lane + 32 * j;
    float x = -INFINITY;
    if (c < Sk) {
      x = sp[c];
      if (mp && !mp[c]) x += -FLT_MAX;
      if (causal && c > qi + past) x += -FLT_MAX;
    }
    v[j] = x;
    mx = fmaxf(mx, x);
  }
  mx = warp_max(mx);
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < EPL; ++j) {
    const int c = lane + 32 * j;
    v[j] = c < Sk ? __expf(v[j] - mx) : 0.f;
    sum += v[j];
  }
  const float inv = 1.f / warp_sum(sum);
  __nv_bfloat16* pp = p + row * ld;
#pragma unroll
  for (int j = 0; j < EPL; ++j) {
    const int c = lane + 32 * j;
    if (c < ld) pp[c] = __float2bfloat16_rn(v[j] * inv);  // pad columns [Sk, ld) get zeros
  }
}

template <int EPL>
__global__ void __launch_bounds__(kSmWarps * 32)
softmax_bwd_kernel(const __nv_bfloat16* __restrict__ p, const float* __restrict__ dp, __nv_bfloat16* __restrict__ ds,
                   long long rows, int Sk, int ld) {
  const int lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * kSmWarps + (threadIdx.x >> 5);
  if (row >= rows) return;
  float pv[EPL], g[EPL];
  float dot = 0.f;
#pragma unroll
  for (int j = 0; j < EPL; ++j) {
    const int c = lane + 32 * j;
    pv[j] = c < Sk ? __bfloat162float(p[row * ld + c]) : 0.f;
    g[j] = c < Sk ? dp[row * ld + c] : 0.f;
    dot += pv[j] * g[j];
  }
  dot = warp_sum(dot);
#pragma unroll
  for (int j = 0; j < EPL; ++j) {
    const int c = lane + 32 * j;
    if (c < ld) ds[row * ld + c] = __float2bfloat16_rn(pv[j] * (g[j] - dot));
  }
}

#define VB_DISPATCH_EPL(n, CALL)                                  \
  do {                                                            \
    const int epl_ = ((n) + 31) / 32;                             \
    if (epl_ <= 1) { constexpr int EPL = 1; CALL; }               \
    else if (epl_ <= 2) { constexpr int EPL = 2; CALL; }          \
    else if (epl_ <= 4) { constexpr int EPL = 4; CALL; }          \
    else if (epl_ <= 8) { constexpr int EPL = 8; CALL; }          \
    else if (epl_ <= 16) { constexpr int EPL = 16; CALL; }        \
    else if (epl_ <= 34) { constexpr int EPL = 34; CALL; }        \
    else return fail(VACNIC_EINVAL, "softmax: row length %d > 1088 not supported", (n)); \
  } while (0)

}  // namespace vb

using namespace vb;

extern "C" int vacnic_softmax_fwd(const float* scores, void* probs, const uint8_t* key_mask, int32_t B, int32_t H,
                                  int32_t Sq, int32_t Sk, int32_t ld, int32_t causal, int32_t past, void* stream) {
  VB_REQUIRE(scores && probs, "softmax_fwd: null pointer");
  VB_REQUIRE(B > 0 && H > 0 && Sq > 0 && Sk > 0 && ld >= Sk, "softmax_fwd: bad shape");
  const long long rows = static_cast<long long>(B) * H * Sq;
  const int grid = static_cast<int>((rows + kSmWarps - 1) / kSmWarps);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  VB_DISPATCH_EPL(ld, (softmax_fwd_kernel<EPL><<<grid, kSmWarps * 32, 0, s>>>(
                          scores, static_cast<__nv_bfloat16*>(probs), key_mask, rows, Sq, Sk, ld, H * Sq, causal, past)));
  count_launch();
  return check_last("softmax_fwd");
}

extern "C" int vacnic_softmax_bwd(const void* probs, const float* dprobs, void* dscores, int64_t rows, int32_t Sk,
                                  int32_t ld, void* stream) {
  VB_REQUIRE(probs && dprobs && dscores, "softmax_bwd: null pointer");
  VB_REQUIRE(rows > 0 && Sk > 0 && ld >= Sk, "softmax_bwd: bad shape");
  const int grid = static_cast<int>((rows + kSmWarps - 1) / kSmWarps);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  VB_DISPATCH_EPL(ld, (softmax_bwd_kernel<EPL><<<grid, kSmWarps * 32, 0, s>>>(
                          static_cast<const __nv_bfloat16*>(probs), dprobs, static_cast<__nv_bfloat16*>(dscores), rows,
                          Sk, ld)));
  count_launch();
  return check_last("softmax_bwd");
}
