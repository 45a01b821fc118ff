// Shared by the 1-CTA (gemm_sm100.cu) and CTA-pair (gemm2_sm100.cu) tcgen05 GEMM kernels: launch
// arguments and the register-level epilogue (bias / alpha / GELU / tanh / activation gradient / accumulate).
#pragma once
#include "common.h"
#include "ptx.cuh"

namespace vb {

struct GemmArgs {
  int M, N, K;
  int batch0, batch1;
  int num_m, num_n;
  long long total_tiles;
  void* c;
  long long ldc, c_sb0, c_sb1;
  const float* bias;
  void* aux_out;
  const void* aux_in;
  float alpha;
  int c_dtype, act, dact, accumulate;
  int vec_ok;  // 16-byte vector access allowed on C / aux rows
  int bias_vec;  // bias pointer 16-byte aligned
  int a_m0, a_m1, b_m0, b_m1;  // batch-coordinate multipliers: 0 = operand is broadcast over that batch dim
  // split-K (single-CTA kernel): work item = (tile, split); partial accumulators go through `ws`, the last CTA to
  // finish a tile (ws_count) sums them in split order and runs the epilogue
  int splits, kb_per_split;
  float* ws;
  int* ws_count;
  long long c_chunk;  // != 0: column n lives at (n / 64) * c_chunk + (n % 64) (head-major outputs)
};

constexpr int kBM = 128;
constexpr int kBK = 64;
constexpr int kEpiWarps = 8;
constexpr int kThreads = 64 + 32 * kEpiWarps;

// BK = K extent of one pipeline stage.  The skinny problems served by BN = 64 are bound by the per-stage round trip
// (TMA -> mbarrier -> MMA -> commit -> refill, ~0.29 us however many bytes ride on it), so they use 128-wide stages:
// half as many round trips for the same bytes in flight.
template <int BN, int BK = kBK>
struct GemmCfg {
  static constexpr int kABytes = kBM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (BN == 256) ? 4 : (BN == 128 ? 6 : (BK == 128 ? 4 : 8));
  static constexpr int kTmemCols = (2 * BN <= 32) ? 32 : 2 * BN;  // 128 / 256 / 512
  static constexpr int kBarBytes = 256;
  static constexpr int kSmemBytes = kStages * kStageBytes + kBarBytes + 1024;
};

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == VACNIC_ACT_GELU) return gelu_erf(v);
  if (act == VACNIC_ACT_TANH) return tanhf(v);
  return v;
}
__device__ __forceinline__ float apply_dact(float aux, int dact) {
  if (dact == VACNIC_ACT_GELU) return gelu_erf_grad(aux);
  if (dact == VACNIC_ACT_TANH) return 1.0f - aux * aux;
  return 1.0f;
}

// Epilogue for 32 consecutive columns of one output row held in registers.
__device__ __forceinline__ void epilogue_row_chunk(const GemmArgs& g, float (&v)[32],
                                                   long long row_off, int n0) {
  const int nvalid = min(32, g.N - n0);
  if (g.bias != nullptr) {
    if (nvalid == 32 && g.bias_vec) {
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(g.bias + n0) + q);
        v[4 * q + 0] += b4.x; v[4 * q + 1] += b4.y; v[4 * q + 2] += b4.z; v[4 * q + 3] += b4.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < nvalid) v[j] += __ldg(g.bias + n0 + j);
    }
  }
  if (g.alpha != 1.0f) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] *= g.alpha;
  }
  const bool full = (nvalid == 32) && g.vec_ok;
  const long long off = row_off + (g.c_chunk != 0 ? static_cast<long long>(n0 >> 6) * g.c_chunk + (n0 & 63) : n0);
  if (g.aux_out != nullptr) {
    __nv_bfloat16* p = reinterpret_cast<__nv_bfloat16*>(g.aux_out) + off;
    if (full) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint4 u;
        u.x = pack_bf16x2(v[8 * q + 0], v[8 * q + 1]);
        u.y = pack_bf16x2(v[8 * q + 2], v[8 * q + 3]);
        u.z = pack_bf16x2(v[8 * q + 4], v[8 * q + 5]);
        u.w = pack_bf16x2(v[8 * q + 6], v[8 * q + 7]);
        reinterpret_cast<uint4*>(p)[q] = u;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < nvalid) p[j] = __float2bfloat16_rn(v[j]);
    }
  }
  // one warp-uniform branch per activation kind (never both evaluated and selected per element)
  if (g.act == VACNIC_ACT_GELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
  } else if (g.act == VACNIC_ACT_TANH) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = tanhf(v[j]);
  }
  if (g.dact != VACNIC_ACT_NONE) {
    const __nv_bfloat16* p = reinterpret_cast<const __nv_bfloat16*>(g.aux_in) + off;
    if (full) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(p) + q);
        float2 f;
        f = unpack_bf16x2(u.x); v[8 * q + 0] *= apply_dact(f.x, g.dact); v[8 * q + 1] *= apply_dact(f.y, g.dact);
        f = unpack_bf16x2(u.y); v[8 * q + 2] *= apply_dact(f.x, g.dact); v[8 * q + 3] *= apply_dact(f.y, g.dact);
        f = unpack_bf16x2(u.z); v[8 * q + 4] *= apply_dact(f.x, g.dact); v[8 * q + 5] *= apply_dact(f.y, g.dact);
        f = unpack_bf16x2(u.w); v[8 * q + 6] *= apply_dact(f.x, g.dact); v[8 * q + 7] *= apply_dact(f.y, g.dact);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < nvalid) v[j] *= apply_dact(__bfloat162float(p[j]), g.dact);
    }
  }
  if (g.c_dtype == VACNIC_DT_F32) {
    float* p = reinterpret_cast<float*>(g.c) + off;
    if (full) {
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float4 o = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        if (g.accumulate) {
          const float4 c = reinterpret_cast<const float4*>(p)[q];
          o.x += c.x; o.y += c.y; o.z += c.z; o.w += c.w;
        }
        reinterpret_cast<float4*>(p)[q] = o;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < nvalid) p[j] = g.accumulate ? p[j] + v[j] : v[j];
    }
  } else {
    __nv_bfloat16* p = reinterpret_cast<__nv_bfloat16*>(g.c) + off;
    if (full) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (g.accumulate) {
          const uint4 c = reinterpret_cast<const uint4*>(p)[q];
          float2 f;
          f = unpack_bf16x2(c.x); v[8 * q + 0] += f.x; v[8 * q + 1] += f.y;
          f = unpack_bf16x2(c.y); v[8 * q + 2] += f.x; v[8 * q + 3] += f.y;
          f = unpack_bf16x2(c.z); v[8 * q + 4] += f.x; v[8 * q + 5] += f.y;
          f = unpack_bf16x2(c.w); v[8 * q + 6] += f.x; v[8 * q + 7] += f.y;
        }
        uint4 u;
        u.x = pack_bf16x2(v[8 * q + 0], v[8 * q + 1]);
        u.y = pack_bf16x2(v[8 * q + 2], v[8 * q + 3]);
        u.z = pack_bf16x2(v[8 * q + 4], v[8 * q + 5]);
        u.w = pack_bf16x2(v[8 * q + 6], v[8 * q + 7]);
        reinterpret_cast<uint4*>(p)[q] = u;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < nvalid)
          p[j] = __float2bfloat16_rn(g.accumulate ? __bfloat162float(p[j]) + v[j] : v[j]);
    }
  }
}

}  // namespace vb
