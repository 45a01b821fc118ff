// Thin inline-PTX wrappers for sm_100a: mbarrier, TMA (cp.async.bulk.tensor),
// tcgen05 (alloc / mma / commit / ld) and small math helpers shared by all kernels.
// Everything here is written for -gencode arch=compute_100a,code=sm_100a only.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace vb {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .b32 %%rx;\n\t"
      ".reg .pred %%px;\n\t"
      "elect.sync %%rx|%%px, %1;\n\t"
      "@%%px mov.s32 %0, 1;\n\t"
      "}\n"
      : "+r"(pred)
      : "r"(0xffffffffu));
  return pred != 0;
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// ---------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// 4-D tiled load, completion signalled on an mbarrier (bytes).
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar,
                                            int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0),
      "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// ---------------------------------------------------------------- tcgen05
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 inputs, fp32 accumulate. One thread issues.
__device__ __forceinline__ void umma_bf16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                             uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// All prior tcgen05.mma of this thread arrive (count 1) on the mbarrier when complete.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread = lane row).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32"
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15,"
      " %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31},"
      "[%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
// 32 registers per thread -> 32 lanes x 32 consecutive fp32 columns (inverse of tmem_ld_32x32)
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0],"
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16,"
      " %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Shared-memory matrix descriptor (SWIZZLE_128B, version 1). Offsets in bytes.
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes,
                                                         uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);            // [0,14)  start address
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;       // [16,30) leading byte offset
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;       // [32,46) stride byte offset
  d |= static_cast<uint64_t>(1) << 46;                               // [46,48) version = 1
  d |= static_cast<uint64_t>(2) << 61;                               // [61,64) SWIZZLE_128B
  return d;
}

// kind::f16 instruction descriptor: bf16 x bf16 -> fp32, dense.
__host__ __device__ constexpr uint32_t make_idesc_bf16(int m, int n, int a_mn_major,
                                                       int b_mn_major) {
  return (1u << 4)                                   // c_format = F32
         | (1u << 7)                                 // a_format = BF16
         | (1u << 10)                                // b_format = BF16
         | (static_cast<uint32_t>(a_mn_major) << 15) // a_major
         | (static_cast<uint32_t>(b_mn_major) << 16) // b_major
         | (static_cast<uint32_t>(n >> 3) << 17)     // n_dim
         | (static_cast<uint32_t>(m >> 4) << 24);    // m_dim
}

// ---------------------------------------------------------------- shared-space vector access (32-bit addresses)
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
  uint4 u;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "r"(addr));
  return u;
}

__device__ __forceinline__ void st_global_v4(void* p, const uint4& u) {
  asm volatile("st.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(u.x), "r"(u.y), "r"(u.z), "r"(u.w) : "memory");
}

// ---------------------------------------------------------------- math
// erf for the GELU epilogues of the FFN GEMMs, which run once per output element on two warps per scheduler and
// must hide behind the tile's main loop: Abramowitz & Stegun 7.1.26 (|error| <= 1.5e-7, far below bf16 resolution)
// written against the raw MUFU operations -- rcp.approx.ftz / ex2.approx.ftz through inline PTX, 2 MUFU + 11 FMA-pipe
// instructions per GELU.  The same formula through __fdividef / __expf / copysignf compiled to 39 SASS instructions
// per element (range fix-ups of the division, denormal handling of the exponential) and made the K = 1024 GELU GEMM
// epilogue-paced.  (An FMA-only odd polynomial of the same cost is limited to ~3e-5 by fp32 cancellation; that was
// enough to move the SECLA loss of the full-size parity test by 2.6 %, so the accurate form is kept.)
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// r = erf(z), e = exp(-z^2) for z >= 0
__device__ __forceinline__ void erf_exp_pos(float z, float& r, float& e) {
  const float t = rcp_approx(fmaf(0.3275911f, z, 1.0f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  e = ex2_approx(-1.44269504088896340736f * z * z);
  r = fmaf(-p * t, e, 1.0f);
}
__device__ __forceinline__ float erf_fast(float x) {
  float r, e;
  erf_exp_pos(fabsf(x), r, e);
  return copysignf(r, x);
}
// tanh = 1 - 2 / (exp(2x) + 1): 1 MUFU.EX2 + 1 MUFU.RCP, |error| ~ 2e-7 (saturates correctly at +-inf)
__device__ __forceinline__ float tanh_fast(float x) {
  const float e = ex2_approx(x * 2.88539008177792681472f);
  return 1.0f - __fdividef(2.0f, e + 1.0f);
}
// QuickGELU of OpenAI CLIP's ViT MLP: x * sigmoid(1.702 x) = x / (1 + exp(-1.702 x)); 1 MUFU.EX2 + 1 MUFU.RCP
__device__ __forceinline__ float quick_gelu(float x) {
  const float e = ex2_approx(x * -2.45546696f);  // -1.702 * log2(e)
  return __fdividef(x, 1.0f + e);
}
__device__ __forceinline__ float gelu_erf(float x) {
  float r, e;
  erf_exp_pos(fabsf(x) * 0.70710678118654752440f, r, e);
  const float h = 0.5f * x;
  return fmaf(h, copysignf(r, x), h);
}
// d/dx [x Phi(x)] = Phi(x) + x phi(x); the exponential of the erf evaluation is exp(-x^2 / 2), i.e. phi up to a factor
__device__ __forceinline__ float gelu_erf_grad(float x) {
  float r, e;
  erf_exp_pos(fabsf(x) * 0.70710678118654752440f, r, e);
  const float cdf = fmaf(0.5f, copysignf(r, x), 0.5f);
  return fmaf(x, 0.39894228040143267794f * e, cdf);
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 t = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(t);
}


// Counter-based dropout (shared by LayerNorm, activation and attention dropout): Bernoulli keep decision for element `idx` of
// call site `salt` in step `seed` -- the backward pass regenerates the mask from the same triple, nothing is stored.
__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x85EBCA6Bu; x ^= x >> 13; x *= 0xC2B2AE35u; x ^= x >> 16;
  return x;
}
__device__ __forceinline__ bool keep_elem(uint32_t seed, uint32_t salt, uint64_t idx, uint32_t thresh) {
  uint32_t h = hash32(static_cast<uint32_t>(idx) * 0x9E3779B1u + seed);
  h = hash32(h ^ (static_cast<uint32_t>(idx >> 32) + salt * 0x7F4A7C15u));
  return h >= thresh;
}
// Two keep decisions per counter hash (16-bit fields): for the LayerNorm kernels, whose instruction count is dominated by
// the generator (two hash rounds per element were 53 % of the forward's and 39 % of the backward's instructions).  p is
// quantised to 1/65536 (0.1 -> 0.10000610); the kernels scale by the matching 65536 / (65536 - t16).
__host__ __device__ inline uint32_t drop_thresh16(float p) {
  const double t = static_cast<double>(p) * 65536.0 + 0.5;
  return t >= 65535.0 ? 65535u : static_cast<uint32_t>(t);
}
__device__ __forceinline__ uint32_t keep_pair(uint32_t seed, uint32_t salt, uint64_t pair, uint32_t t16) {
  uint32_t h = hash32(static_cast<uint32_t>(pair) * 0x9E3779B1u + seed);
  h = hash32(h ^ (static_cast<uint32_t>(pair >> 32) + salt * 0x7F4A7C15u));
  return ((h & 0xffffu) >= t16 ? 1u : 0u) | ((h >> 16) >= t16 ? 2u : 0u);  // bit 0: element 2*pair, bit 1: element 2*pair + 1
}
// keep bits of the d/32 elements one lane holds of a row (vector lane + 32 j, element e -> bit j*8 + e; d <= 1024)
template <int VPL>
__device__ __forceinline__ uint32_t keep_row_bits(uint32_t seed, uint32_t salt, long long row, int d, int lane, uint32_t t16) {
  uint32_t keep = 0u;
#pragma unroll
  for (int j = 0; j < VPL; ++j) {
    const uint64_t pair0 = (static_cast<uint64_t>(row) * d + (lane + 32 * j) * 8) >> 1;
#pragma unroll
    for (int q = 0; q < 4; ++q) keep |= keep_pair(seed, salt, pair0 + q, t16) << (j * 8 + q * 2);
  }
  return keep;
}
__host__ __device__ inline uint32_t drop_thresh(float p) {
  double t = static_cast<double>(p) * 4294967296.0;
  return t >= 4294967295.0 ? 0xFFFFFFFFu : static_cast<uint32_t>(t);
}

// One AdamW element update (torch.optim.AdamW, TRAIN:95-101) with every rounding pinned by intrinsics, so that the flat
// kernel (elementwise.cu) and the rank-sharded peer-memory kernel (dp.cu) produce bit-identical weights from identical
// gradients whatever the compiler's FMA contraction does (a last-bit difference is enough to flip the sign-like Adam
// update of parameters whose true gradient is zero, e.g. every k_proj.bias).
struct AdamwCoef {
  float b1, b2, one_m_b1, one_m_b2, eps, decay, step_size, inv_sqrt_bc2, gs;
};
__device__ __forceinline__ AdamwCoef adamw_coef(const float* __restrict__ hyper) {
  AdamwCoef c;
  const float lr = hyper[0], wd = hyper[4], bc1 = hyper[5], bc2 = hyper[6];
  c.b1 = hyper[1]; c.b2 = hyper[2]; c.eps = hyper[3]; c.gs = hyper[7];
  c.one_m_b1 = __fsub_rn(1.f, c.b1); c.one_m_b2 = __fsub_rn(1.f, c.b2);
  c.decay = __fsub_rn(1.f, __fmul_rn(lr, wd));
  c.step_size = __fdiv_rn(lr, bc1);
  c.inv_sqrt_bc2 = __fdiv_rn(1.f, __fsqrt_rn(bc2));
  return c;
}
__device__ __forceinline__ void adamw_elem(const AdamwCoef& c, float g, float& p, float& m, float& v) {
  const float gr = __fmul_rn(g, c.gs);
  p = __fmul_rn(p, c.decay);  // decoupled weight decay
  m = __fmaf_rn(c.b1, m, __fmul_rn(c.one_m_b1, gr));
  v = __fmaf_rn(c.b2, v, __fmul_rn(__fmul_rn(c.one_m_b2, gr), gr));
  const float denom = __fadd_rn(__fmul_rn(__fsqrt_rn(v), c.inv_sqrt_bc2), c.eps);
  p = __fsub_rn(p, __fdiv_rn(__fmul_rn(c.step_size, m), denom));
}

}  // namespace vb
