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
  if (act == VACNIC_ACT_TANH) return tanh_fast(v);
  if (act == VACNIC_ACT_QUICKGELU) return quick_gelu(v);
  return v;
}
__device__ __forceinline__ float apply_dact(float aux, int dact) {
  if (dact == VACNIC_ACT_GELU) return gelu_erf_grad(aux);
  if (dact == VACNIC_ACT_TANH) return 1.0f - aux * aux;
  return 1.0f;
}

// ------------------------------------------------------------------------------------------------------------
// Epilogue for 32 consecutive columns of one output row held in registers, in pieces so that the hot (full chunk,
// 16-byte aligned) code is compact and contiguous and the ragged scalar code lives out of line: the fully inlined
// original was 20 k SASS instructions in the CTA-pair kernel and 27 % of its stall samples were instruction-cache
// misses (profiles/r1_ncu_gemm2_*.md).
// ------------------------------------------------------------------------------------------------------------

// bias (from shared memory when staged there: zero beyond N, 16-byte aligned; else 16-byte vector loads from global
// memory) and alpha for a full chunk
__device__ __forceinline__ void epilogue_bias_alpha(const GemmArgs& g, float (&v)[32], int n0, const float* sbias) {
  if (sbias != nullptr) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 b4 = reinterpret_cast<const float4*>(sbias)[q];
      v[4 * q + 0] += b4.x; v[4 * q + 1] += b4.y; v[4 * q + 2] += b4.z; v[4 * q + 3] += b4.w;
    }
  } else if (g.bias != nullptr) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 b4 = __ldg(reinterpret_cast<const float4*>(g.bias + n0) + q);
      v[4 * q + 0] += b4.x; v[4 * q + 1] += b4.y; v[4 * q + 2] += b4.z; v[4 * q + 3] += b4.w;
    }
  }
  if (g.alpha != 1.0f) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] *= g.alpha;
  }
}
// may this chunk take the vector code?  (warp-uniform)
__device__ __forceinline__ bool epilogue_chunk_is_vec(const GemmArgs& g, int n0, const float* sbias) {
  return g.N - n0 >= 32 && g.vec_ok && (sbias != nullptr || g.bias == nullptr || g.bias_vec);
}

__device__ __forceinline__ long long epilogue_col_off(const GemmArgs& g, int n0) {
  return g.c_chunk != 0 ? static_cast<long long>(n0 >> 6) * g.c_chunk + (n0 & 63) : n0;
}

// one warp-uniform branch per activation kind (never both evaluated and selected per element)
__device__ __forceinline__ void epilogue_act(const GemmArgs& g, float (&v)[32]) {
  if (g.act == VACNIC_ACT_GELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
  } else if (g.act == VACNIC_ACT_TANH) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = tanh_fast(v[j]);
  } else if (g.act == VACNIC_ACT_QUICKGELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = quick_gelu(v[j]);
  }
}

// v *= act'(aux_in) for a full, aligned chunk
__device__ __forceinline__ void epilogue_dact_vec(const GemmArgs& g, float (&v)[32], long long off) {
  const __nv_bfloat16* p = reinterpret_cast<const __nv_bfloat16*>(g.aux_in) + off;
  uint4 u[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) u[q] = __ldg(reinterpret_cast<const uint4*>(p) + q);
  if (g.dact == VACNIC_ACT_GELU) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float2 f;
      f = unpack_bf16x2(u[q].x); v[8 * q + 0] *= gelu_erf_grad(f.x); v[8 * q + 1] *= gelu_erf_grad(f.y);
      f = unpack_bf16x2(u[q].y); v[8 * q + 2] *= gelu_erf_grad(f.x); v[8 * q + 3] *= gelu_erf_grad(f.y);
      f = unpack_bf16x2(u[q].z); v[8 * q + 4] *= gelu_erf_grad(f.x); v[8 * q + 5] *= gelu_erf_grad(f.y);
      f = unpack_bf16x2(u[q].w); v[8 * q + 6] *= gelu_erf_grad(f.x); v[8 * q + 7] *= gelu_erf_grad(f.y);
    }
  } else {  // tanh: aux holds tanh(z)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float2 f;
      f = unpack_bf16x2(u[q].x); v[8 * q + 0] *= 1.0f - f.x * f.x; v[8 * q + 1] *= 1.0f - f.y * f.y;
      f = unpack_bf16x2(u[q].y); v[8 * q + 2] *= 1.0f - f.x * f.x; v[8 * q + 3] *= 1.0f - f.y * f.y;
      f = unpack_bf16x2(u[q].z); v[8 * q + 4] *= 1.0f - f.x * f.x; v[8 * q + 5] *= 1.0f - f.y * f.y;
      f = unpack_bf16x2(u[q].w); v[8 * q + 6] *= 1.0f - f.x * f.x; v[8 * q + 7] *= 1.0f - f.y * f.y;
    }
  }
}

__device__ __forceinline__ void pack_chunk_bf16(const float (&v)[32], uint4 (&u)[4]) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    u[q].x = pack_bf16x2(v[8 * q + 0], v[8 * q + 1]);
    u[q].y = pack_bf16x2(v[8 * q + 2], v[8 * q + 3]);
    u[q].z = pack_bf16x2(v[8 * q + 4], v[8 * q + 5]);
    u[q].w = pack_bf16x2(v[8 * q + 6], v[8 * q + 7]);
  }
}

// thread-owns-row stores of a full, aligned chunk: 32 bf16 (64 B) at p
__device__ __forceinline__ void store_row_bf16_vec(__nv_bfloat16* p, const float (&v)[32]) {
  uint4 u[4];
  pack_chunk_bf16(v, u);
#pragma unroll
  for (int q = 0; q < 4; ++q) reinterpret_cast<uint4*>(p)[q] = u[q];
}

// C store of a full, aligned chunk, thread-owns-row (fp32 or bf16, optional accumulate)
__device__ __forceinline__ void epilogue_store_vec(const GemmArgs& g, float (&v)[32], long long off) {
  if (g.c_dtype == VACNIC_DT_F32) {
    float* p = reinterpret_cast<float*>(g.c) + off;
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
    __nv_bfloat16* p = reinterpret_cast<__nv_bfloat16*>(g.c) + off;
    if (g.accumulate) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint4 c = reinterpret_cast<const uint4*>(p)[q];
        float2 f;
        f = unpack_bf16x2(c.x); v[8 * q + 0] += f.x; v[8 * q + 1] += f.y;
        f = unpack_bf16x2(c.y); v[8 * q + 2] += f.x; v[8 * q + 3] += f.y;
        f = unpack_bf16x2(c.z); v[8 * q + 4] += f.x; v[8 * q + 5] += f.y;
        f = unpack_bf16x2(c.w); v[8 * q + 6] += f.x; v[8 * q + 7] += f.y;
      }
    }
    store_row_bf16_vec(p, v);
  }
}

// The whole epilogue of a ragged (N edge) or unaligned chunk: scalar, rare, out of line (v lives in local memory).
static __device__ __noinline__ void epilogue_chunk_ragged(const GemmArgs& g, float* v, long long off, int n0, int nvalid,
                                                   const float* sbias) {
  if (sbias != nullptr) {
    for (int j = 0; j < nvalid; ++j) v[j] += sbias[j];
  } else if (g.bias != nullptr) {
    for (int j = 0; j < nvalid; ++j) v[j] += __ldg(g.bias + n0 + j);
  }
  if (g.alpha != 1.0f) {
    for (int j = 0; j < nvalid; ++j) v[j] *= g.alpha;
  }
  if (g.aux_out != nullptr) {
    __nv_bfloat16* p = reinterpret_cast<__nv_bfloat16*>(g.aux_out) + off;
    for (int j = 0; j < nvalid; ++j) p[j] = __float2bfloat16_rn(v[j]);
  }
  if (g.act != VACNIC_ACT_NONE) {
    for (int j = 0; j < nvalid; ++j) v[j] = apply_act(v[j], g.act);
  }
  if (g.dact != VACNIC_ACT_NONE) {
    const __nv_bfloat16* p = reinterpret_cast<const __nv_bfloat16*>(g.aux_in) + off;
    for (int j = 0; j < nvalid; ++j) v[j] *= apply_dact(__bfloat162float(p[j]), g.dact);
  }
  if (g.c_dtype == VACNIC_DT_F32) {
    float* p = reinterpret_cast<float*>(g.c) + off;
    for (int j = 0; j < nvalid; ++j) p[j] = g.accumulate ? p[j] + v[j] : v[j];
  } else {
    __nv_bfloat16* p = reinterpret_cast<__nv_bfloat16*>(g.c) + off;
    for (int j = 0; j < nvalid; ++j)
      p[j] = __float2bfloat16_rn(g.accumulate ? __bfloat162float(p[j]) + v[j] : v[j]);
  }
}
__device__ __forceinline__ void epilogue_chunk_ragged_call(const GemmArgs& g, const float (&v)[32], long long off,
                                                           int n0, const float* sbias) {
  float t[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) t[j] = v[j];
  epilogue_chunk_ragged(g, t, off, n0, min(32, g.N - n0), sbias);
}

// The whole epilogue of one row chunk with thread-owns-row stores (single-CTA kernel; fallback of the pair kernel).
__device__ __forceinline__ void epilogue_row_chunk(const GemmArgs& g, float (&v)[32],
                                                   long long row_off, int n0, const float* sbias = nullptr) {
  const long long off = row_off + epilogue_col_off(g, n0);
  if (epilogue_chunk_is_vec(g, n0, sbias)) {
    epilogue_bias_alpha(g, v, n0, sbias);
    if (g.aux_out != nullptr) store_row_bf16_vec(reinterpret_cast<__nv_bfloat16*>(g.aux_out) + off, v);
    epilogue_act(g, v);
    if (g.dact != VACNIC_ACT_NONE) epilogue_dact_vec(g, v, off);
    epilogue_store_vec(g, v, off);
  } else {
    epilogue_chunk_ragged_call(g, v, off, n0, sbias);
  }
}

// Warp-cooperative bf16 store of a 32-row x 32-column chunk (lane = row on entry) through a 2 KB per-warp staging
// buffer: after the XOR-swizzled transpose four adjacent lanes hold the four 16-byte pieces of one row, so one
// STG.128 covers 8 rows x 64 contiguous bytes (8 fully written sector pairs) instead of 32 rows x 16 bytes (32
// half-written sectors) -- 4x fewer LSU / L2 requests for the same bytes.  `p0` = address of (row of lane 0,
// first column of the chunk); rows >= rows_valid are not written.  All 32 lanes must call.
__device__ __forceinline__ void store_chunk_bf16_coalesced(uint32_t stage, int lane, const float (&v)[32],
                                                           __nv_bfloat16* p0, long long ld, int rows_valid) {
  uint4 u[4];
  pack_chunk_bf16(v, u);
  const uint32_t sw = static_cast<uint32_t>(lane >> 1) & 3u;
  const uint32_t wr = stage + static_cast<uint32_t>(lane) * 64u;
#pragma unroll
  for (uint32_t q = 0; q < 4; ++q) st_shared_v4(wr + ((q ^ sw) << 4), u[q].x, u[q].y, u[q].z, u[q].w);
  __syncwarp();
  const uint32_t q = static_cast<uint32_t>(lane) & 3u;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int r = 8 * j + (lane >> 2);
    const uint4 w = ld_shared_v4(stage + static_cast<uint32_t>(r) * 64u + ((q ^ (static_cast<uint32_t>(r >> 1) & 3u)) << 4));
    if (r < rows_valid) st_global_v4(p0 + static_cast<long long>(r) * ld + q * 8, w);
  }
  __syncwarp();
}

}  // namespace vb
