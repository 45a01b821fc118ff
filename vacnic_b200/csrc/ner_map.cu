// NER-prefix map of the visually-aware encoder layer (MFULL:682-687):
//   X = ner.reshape(B, d, E)            -- a memory REINTERPRETATION of the contiguous [B, E, d] buffer
//   Y = gelu(X W_up^T + b_up)           W_up [U, E]   (U = 4*G = 80, E = max_ner_type_len = 80)
//   Z = Y W_down^T + b_down             W_down [G, U] (G = max_ner_type_len_gt = 20)
//   prefix = Z.reshape(B, G, d)         -- again a reinterpretation, then LayerNorm (norm.cu)
// With R = B*d rows of E contiguous elements the two contractions have K = 80 and N = 80 / 20: far
// too small (and, for N = 20, too misaligned for TMA: 40-byte rows) for the tensor-core tile path,
// and only 0.2 GFLOP per layer, so this is a CUDA-core kernel: one thread per row, weights staged in
// shared memory as fp32, activations bf16.  The backward kernel produces dX and accumulates the four
// parameter gradients through shared-memory tiles and one atomicAdd per entry per block.
#include "common.h"
#include "ptx.cuh"

namespace vb {

constexpr int kNmThreads = 128;
constexpr int kMaxE = 80, kMaxU = 80, kMaxG = 32;

__global__ void __launch_bounds__(kNmThreads)
ner_map_fwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w_up,
                   const float* __restrict__ b_up, const __nv_bfloat16* __restrict__ w_down,
                   const float* __restrict__ b_down, __nv_bfloat16* __restrict__ z1, __nv_bfloat16* __restrict__ z2,
                   long long rows, int E, int U, int G) {
  extern __shared__ float smem[];
  float* s_up = smem;                 // [U][E]
  float* s_down = s_up + U * E;       // [G][U]
  float* s_bup = s_down + G * U;      // [U]
  float* s_bdown = s_bup + U;         // [G]
  for (int i = threadIdx.x; i < U * E; i += blockDim.x) s_up[i] = __bfloat162float(w_up[i]);
  for (int i = threadIdx.x; i < G * U; i += blockDim.x) s_down[i] = __bfloat162float(w_down[i]);
  for (int i = threadIdx.x; i < U; i += blockDim.x) s_bup[i] = b_up[i];
  for (int i = threadIdx.x; i < G; i += blockDim.x) s_bdown[i] = b_down[i];
  __syncthreads();
  const long long r = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  float xr[kMaxE];
#pragma unroll
  for (int i = 0; i < kMaxE; ++i) xr[i] = i < E ? __bfloat162float(x[r * E + i]) : 0.f;
  float acc[kMaxG];
#pragma unroll
  for (int m = 0; m < kMaxG; ++m) acc[m] = m < G ? s_bdown[m] : 0.f;
  for (int j = 0; j < U; ++j) {
    float z = s_bup[j];
#pragma unroll
    for (int i = 0; i < kMaxE; ++i)
      if (i < E) z += s_up[j * E + i] * xr[i];
    // the GEMM path rounds the pre-activation to bf16 before storing it; keep the same rounding point
    const __nv_bfloat16 zb = __float2bfloat16_rn(z);
    z1[r * U + j] = zb;
    const float y = __bfloat162float(__float2bfloat16_rn(gelu_erf(z)));
#pragma unroll
    for (int m = 0; m < kMaxG; ++m)
      if (m < G) acc[m] += s_down[m * U + j] * y;
  }
#pragma unroll
  for (int m = 0; m < kMaxG; ++m)
    if (m < G) z2[r * G + m] = __float2bfloat16_rn(acc[m]);
}

// Backward, one block = kNmThreads rows.  Phase 1 (thread per row): dy = W_down^T dz2, dz1 = dy*gelu'(z1),
// dx = W_up^T dz1; dz1 / y / x / dz2 of the block are staged in shared memory as bf16.  Phase 2 (all
// threads): dW_up[j,i] += sum_r dz1[r,j] x[r,i], dW_down[m,j] += sum_r dz2[r,m] y[r,j], bias sums.
__global__ void __launch_bounds__(kNmThreads)
ner_map_bwd_kernel(const __nv_bfloat16* __restrict__ dz2, const __nv_bfloat16* __restrict__ z1,
                   const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w_up,
                   const __nv_bfloat16* __restrict__ w_down, __nv_bfloat16* __restrict__ dx,
                   float* __restrict__ dw_up, float* __restrict__ db_up, float* __restrict__ dw_down,
                   float* __restrict__ db_down, long long rows, int E, int U, int G) {
  extern __shared__ float smem[];
  float* s_up = smem;                                    // [U][E] fp32
  float* s_down = s_up + U * E;                          // [G][U] fp32
  __nv_bfloat16* t_dz1 = reinterpret_cast<__nv_bfloat16*>(s_down + G * U);  // [T][U]
  __nv_bfloat16* t_y = t_dz1 + kNmThreads * U;           // [T][U]
  __nv_bfloat16* t_x = t_y + kNmThreads * U;             // [T][E]
  __nv_bfloat16* t_dz2 = t_x + kNmThreads * E;           // [T][G]
  for (int i = threadIdx.x; i < U * E; i += blockDim.x) s_up[i] = __bfloat162float(w_up[i]);
  for (int i = threadIdx.x; i < G * U; i += blockDim.x) s_down[i] = __bfloat162float(w_down[i]);
  __syncthreads();
  const int t = threadIdx.x;
  const long long r = static_cast<long long>(blockIdx.x) * blockDim.x + t;
  const bool live = r < rows;
  float g2[kMaxG];
#pragma unroll
  for (int m = 0; m < kMaxG; ++m) {
    g2[m] = (live && m < G) ? __bfloat162float(dz2[r * G + m]) : 0.f;
    if (m < G) t_dz2[t * G + m] = __float2bfloat16_rn(g2[m]);
  }
  float dxr[kMaxE];
#pragma unroll
  for (int i = 0; i < kMaxE; ++i) {
    dxr[i] = 0.f;
    if (i < E) t_x[t * E + i] = live ? x[r * E + i] : __float2bfloat16_rn(0.f);
  }
  for (int j = 0; j < U; ++j) {
    const float z = live ? __bfloat162float(z1[r * U + j]) : 0.f;
    float dy = 0.f;
#pragma unroll
    for (int m = 0; m < kMaxG; ++m)
      if (m < G) dy += s_down[m * U + j] * g2[m];
    const float dz = __bfloat162float(__float2bfloat16_rn(dy * gelu_erf_grad(z)));
    t_dz1[t * U + j] = __float2bfloat16_rn(live ? dz : 0.f);
    t_y[t * U + j] = __float2bfloat16_rn(live ? gelu_erf(z) : 0.f);
#pragma unroll
    for (int i = 0; i < kMaxE; ++i)
      if (i < E) dxr[i] += s_up[j * E + i] * dz;
  }
  if (live) {
#pragma unroll
    for (int i = 0; i < kMaxE; ++i)
      if (i < E) dx[r * E + i] = __float2bfloat16_rn(dxr[i]);
  }
  __syncthreads();
  for (int e = threadIdx.x; e < U * E; e += blockDim.x) {
    const int j = e / E, i = e % E;
    float s = 0.f;
    for (int q = 0; q < kNmThreads; ++q) s += __bfloat162float(t_dz1[q * U + j]) * __bfloat162float(t_x[q * E + i]);
    atomicAdd(dw_up + e, s);
  }
  for (int e = threadIdx.x; e < G * U; e += blockDim.x) {
    const int m = e / U, j = e % U;
    float s = 0.f;
    for (int q = 0; q < kNmThreads; ++q) s += __bfloat162float(t_dz2[q * G + m]) * __bfloat162float(t_y[q * U + j]);
    atomicAdd(dw_down + e, s);
  }
  for (int j = threadIdx.x; j < U; j += blockDim.x) {
    float s = 0.f;
    for (int q = 0; q < kNmThreads; ++q) s += __bfloat162float(t_dz1[q * U + j]);
    atomicAdd(db_up + j, s);
  }
  for (int m = threadIdx.x; m < G; m += blockDim.x) {
    float s = 0.f;
    for (int q = 0; q < kNmThreads; ++q) s += __bfloat162float(t_dz2[q * G + m]);
    atomicAdd(db_down + m, s);
  }
}

}  // namespace vb

using namespace vb;

extern "C" int vacnic_ner_map_fwd(const void* x, const void* w_up, const float* b_up, const void* w_down,
                                  const float* b_down, void* z1, void* z2, int64_t rows, int32_t E, int32_t U,
                                  int32_t G, void* stream) {
  VB_REQUIRE(x && w_up && b_up && w_down && b_down && z1 && z2, "ner_map_fwd: null pointer");
  VB_REQUIRE(rows > 0 && E > 0 && E <= kMaxE && U > 0 && U <= kMaxU && G > 0 && G <= kMaxG,
             "ner_map_fwd: unsupported sizes E=%d U=%d G=%d (max %d/%d/%d)", E, U, G, kMaxE, kMaxU, kMaxG);
  const size_t smem = (static_cast<size_t>(U) * E + G * U + U + G) * sizeof(float);
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(ner_map_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    configured = true;
  }
  const unsigned grid = static_cast<unsigned>((rows + kNmThreads - 1) / kNmThreads);
  ner_map_fwd_kernel<<<grid, kNmThreads, smem, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(w_up), b_up,
      static_cast<const __nv_bfloat16*>(w_down), b_down, static_cast<__nv_bfloat16*>(z1),
      static_cast<__nv_bfloat16*>(z2), rows, E, U, G);
  count_launch();
  return check_last("ner_map_fwd");
}

extern "C" int vacnic_ner_map_bwd(const void* dz2, const void* z1, const void* x, const void* w_up,
                                  const void* w_down, void* dx, float* dw_up, float* db_up, float* dw_down,
                                  float* db_down, int64_t rows, int32_t E, int32_t U, int32_t G, void* stream) {
  VB_REQUIRE(dz2 && z1 && x && w_up && w_down && dx && dw_up && db_up && dw_down && db_down,
             "ner_map_bwd: null pointer");
  VB_REQUIRE(rows > 0 && E > 0 && E <= kMaxE && U > 0 && U <= kMaxU && G > 0 && G <= kMaxG,
             "ner_map_bwd: unsupported sizes");
  const size_t smem = (static_cast<size_t>(U) * E + G * U) * sizeof(float) +
                      static_cast<size_t>(kNmThreads) * (2 * U + E + G) * sizeof(__nv_bfloat16);
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(ner_map_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
    configured = true;
  }
  const unsigned grid = static_cast<unsigned>((rows + kNmThreads - 1) / kNmThreads);
  ner_map_bwd_kernel<<<grid, kNmThreads, smem, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(dz2), static_cast<const __nv_bfloat16*>(z1),
      static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(w_up),
      static_cast<const __nv_bfloat16*>(w_down), static_cast<__nv_bfloat16*>(dx), dw_up, db_up, dw_down, db_down,
      rows, E, U, G);
  count_launch();
  return check_last("ner_map_bwd");
}
