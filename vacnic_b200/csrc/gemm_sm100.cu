// Batched bf16 GEMM for sm_100a: TMA (cp.async.bulk.tensor) -> 128B-swizzled shared memory ->
// tcgen05.mma (cta_group::1, M=128, N=BN, K=16 per instruction) with fp32 accumulators in TMEM
// (double buffered) -> tcgen05.ld epilogue (bias / alpha / GELU / tanh / activation-gradient /
// accumulate) -> global memory.  Persistent: one CTA per SM walks a static tile list.
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2..9 = epilogue: warp w owns TMEM lanes 32*(w%4) .. +31 and one half of the tile's columns
// (two warps per lane quadrant, so the bias / activation math of a tile is spread over all four SM
// sub-partitions twice and hides behind the next tile's main loop).
//
// This one kernel family serves every dense contraction on the VACNIC hot path (see
// include/vacnic_b200.h for the reference call sites): forward linears (A K-major, B K-major),
// dgrad (B MN-major), wgrad (A and B MN-major), QK^T, PV and their gradients (batched, strided).
#include <cuda.h>

#include "common.h"
#include "gemm_epilogue.cuh"
#include "ptx.cuh"

namespace vb {

template <int BN, bool A_MN, bool B_MN, int BK = kBK>
__global__ void __launch_bounds__(kThreads, 1)
gemm_sm100_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const GemmArgs g) {
  using Cfg = GemmCfg<BN, BK>;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B operands need 1024-byte aligned stage buffers.
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + Cfg::kStages;
  uint64_t* tfull_bar = empty_bar + Cfg::kStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  volatile int* last_flag = reinterpret_cast<volatile int*>(tmem_slot + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_k = (g.K + BK - 1) / BK;
  const long long tiles_per_batch = static_cast<long long>(g.num_m) * g.num_n;
  const long long total_work = g.total_tiles * g.splits;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], kEpiWarps);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_sync();  // on-chip setup above; global memory (operands, split-K workspace, bias, aux, C) only below

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (long long w = blockIdx.x; w < total_work; w += gridDim.x) {
        const long long t = w / g.splits;
        const int kb0 = static_cast<int>(w % g.splits) * g.kb_per_split;
        const int kb1 = min(num_k, kb0 + g.kb_per_split);
        const int batch = static_cast<int>(t / tiles_per_batch);
        const int r = static_cast<int>(t % tiles_per_batch);
        const int m0 = (r % g.num_m) * kBM;
        const int n0 = (r / g.num_m) * BN;
        const int b0 = batch % g.batch0;
        const int b1 = batch / g.batch0;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          uint8_t* sa = smem + stage * Cfg::kStageBytes;
          uint8_t* sb = sa + Cfg::kABytes;
          mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          const int k0 = kb * BK;
          // K-major operand: one [rows x 64] swizzle atom per 64 k-elements; MN-major: per 64-wide M/N chunk a
          // [BK k-rows x 128 B] block, filled 64 k-rows per TMA box
#pragma unroll
          for (int kc = 0; kc < BK / 64; ++kc) {
            if constexpr (!A_MN) {
              tma_load_4d(sa + kc * (kBM * 128), &tmA, &full_bar[stage], k0 + kc * 64, m0, b0 * g.a_m0, b1 * g.a_m1);
            } else {
#pragma unroll
              for (int c = 0; c < kBM / 64; ++c)
                tma_load_4d(sa + c * (BK * 128) + kc * (64 * 128), &tmA, &full_bar[stage], m0 + c * 64, k0 + kc * 64,
                            b0 * g.a_m0, b1 * g.a_m1);
            }
            if constexpr (!B_MN) {
              tma_load_4d(sb + kc * (BN * 128), &tmB, &full_bar[stage], k0 + kc * 64, n0, b0 * g.b_m0, b1 * g.b_m1);
            } else {
#pragma unroll
              for (int c = 0; c < BN / 64; ++c)
                tma_load_4d(sb + c * (BK * 128) + kc * (64 * 128), &tmB, &full_bar[stage], n0 + c * 64, k0 + kc * 64,
                            b0 * g.b_m0, b1 * g.b_m1);
            }
          }
          if (++stage == Cfg::kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(kBM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      // K-major: 8-row groups 1024 B apart (SBO), LBO unused.  MN-major: 64-element chunks along
      // M/N are BK*128 B apart (LBO), 8-k-row groups 1024 B apart (SBO).
      constexpr uint32_t kLboA = A_MN ? BK * 128 : 16, kLboB = B_MN ? BK * 128 : 16;
      int stage = 0;
      uint32_t phase = 0;
      uint32_t it = 0;
      for (long long w = blockIdx.x; w < total_work; w += gridDim.x, ++it) {
        const int kb0 = static_cast<int>(w % g.splits) * g.kb_per_split;
        const int kb1 = min(num_k, kb0 + g.kb_per_split);
        const uint32_t as = it & 1u;
        const uint32_t aphase = (it >> 1) & 1u;
        mbar_wait(&tempty_bar[as], aphase ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint32_t sb = sa + Cfg::kABytes;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // K-major: 16 k-elements = 32 B inside a 64-wide atom, atoms (rows x 128 B) follow each other;
            // MN-major: 16 k-rows = 2048 B
            const uint32_t offA = A_MN ? k * 2048 : (k >> 2) * (kBM * 128) + (k & 3) * 32;
            const uint32_t offB = B_MN ? k * 2048 : (k >> 2) * (BN * 128) + (k & 3) * 32;
            const uint64_t da = make_smem_desc_sw128(sa + offA, kLboA, 1024);
            const uint64_t db = make_smem_desc_sw128(sb + offB, kLboB, 1024);
            umma_bf16_ss(tmem_d, da, db, idesc, (kb > kb0 || k != 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // frees this smem stage when the MMAs retire
          if (++stage == Cfg::kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        umma_commit(&tfull_bar[as]);  // accumulator complete -> epilogue
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue warps 2..9
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;  // which half of the tile's columns
    uint32_t it = 0;
    for (long long w = blockIdx.x; w < total_work; w += gridDim.x, ++it) {
      const long long t = w / g.splits;
      const int sp = static_cast<int>(w % g.splits);
      const int batch = static_cast<int>(t / tiles_per_batch);
      const int r = static_cast<int>(t % tiles_per_batch);
      const int m0 = (r % g.num_m) * kBM;
      const int n0 = (r / g.num_m) * BN;
      const int b0 = batch % g.batch0;
      const int b1 = batch / g.batch0;
      const uint32_t as = it & 1u;
      const uint32_t aphase = (it >> 1) & 1u;
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      const int row = m0 + quad * 32 + lane;
      const long long row_off = static_cast<long long>(b0) * g.c_sb0 +
                                static_cast<long long>(b1) * g.c_sb1 +
                                static_cast<long long>(row) * g.ldc;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + as * BN;
      if (g.splits == 1) {
#pragma unroll 1
        for (int c = half * (BN / 64); c < (half + 1) * (BN / 64); ++c) {
          if (n0 + c * 32 >= g.N) break;
          uint32_t r32[32];
          tmem_ld_32x32(taddr + c * 32, r32);
          tmem_ld_wait();
          if (row < g.M) {
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r32[j]);
            epilogue_row_chunk(g, v, row_off, n0 + c * 32);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[as]);
      } else {
        // ---- split-K: publish the raw partial tile, the last split to arrive reduces in split order
        float* wtile = g.ws + (t * g.splits) * (kBM * BN) + static_cast<long long>(quad * 32 + lane) * BN;
        float* wmine = wtile + static_cast<long long>(sp) * (kBM * BN);
#pragma unroll 1
        for (int c = half * (BN / 64); c < (half + 1) * (BN / 64); ++c) {
          uint32_t r32[32];
          tmem_ld_32x32(taddr + c * 32, r32);
          tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 8; ++q)
            __stcg(reinterpret_cast<float4*>(wmine + c * 32) + q,
                   make_float4(__uint_as_float(r32[4 * q]), __uint_as_float(r32[4 * q + 1]), __uint_as_float(r32[4 * q + 2]),
                               __uint_as_float(r32[4 * q + 3])));
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[as]);
        __threadfence();
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (threadIdx.x == 64) {
          const int old = atomicAdd(g.ws_count + t, 1);
          const int last = (old == g.splits - 1) ? 1 : 0;
          if (last) g.ws_count[t] = 0;  // self-resetting: the next launch finds zeros
          *last_flag = last;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (*last_flag) {
          __threadfence();
#pragma unroll 1
          for (int c = half * (BN / 64); c < (half + 1) * (BN / 64); ++c) {
            if (n0 + c * 32 >= g.N) break;
            if (row < g.M) {
              float v[32];
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = 0.f;
              for (int s2 = 0; s2 < g.splits; ++s2) {
                const float4* src = reinterpret_cast<const float4*>(wtile + static_cast<long long>(s2) * (kBM * BN) + c * 32);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                  const float4 f = __ldcg(src + q);
                  v[4 * q] += f.x; v[4 * q + 1] += f.y; v[4 * q + 2] += f.z; v[4 * q + 3] += f.w;
                }
              }
              epilogue_row_chunk(g, v, row_off, n0 + c * 32);
            }
          }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");  // last_flag is rewritten by the next work item
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) ==
            cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// Tensor map for one operand.  K-major: dims (K, rows, b0, b1), box (64, box_rows).
// MN-major: dims (rows, K, b0, b1), box (64, 64).
int make_operand_map(CUtensorMap* tm, const void* base, bool mn_major, int rows, int K,
                     long long ld, int batch0, long long sb0, int batch1, long long sb1,
                     int box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  if (enc == nullptr) return fail(VACNIC_EDEVICE, "cuTensorMapEncodeTiled entry point not found");
  cuuint64_t dims[4];
  cuuint64_t strides[3];
  cuuint32_t box[4];
  cuuint32_t estr[4] = {1, 1, 1, 1};
  if (!mn_major) {
    dims[0] = static_cast<cuuint64_t>(K);
    dims[1] = static_cast<cuuint64_t>(rows);
    box[0] = 64;
    box[1] = static_cast<cuuint32_t>(box_rows);
  } else {
    dims[0] = static_cast<cuuint64_t>(rows);
    dims[1] = static_cast<cuuint64_t>(K);
    box[0] = 64;
    box[1] = 64;
  }
  dims[2] = static_cast<cuuint64_t>(batch0);
  dims[3] = static_cast<cuuint64_t>(batch1);
  box[2] = 1;
  box[3] = 1;
  strides[0] = static_cast<cuuint64_t>(ld) * 2;
  strides[1] = static_cast<cuuint64_t>(batch0 > 1 ? sb0 : ld) * 2;
  strides[2] = static_cast<cuuint64_t>(batch1 > 1 ? sb1 : ld) * 2;
  for (int i = 0; i < 3; ++i)
    if (strides[i] % 16 != 0 || strides[i] == 0)
      return fail(VACNIC_EINVAL, "gemm: operand stride %d (%llu bytes) not a multiple of 16", i,
                  static_cast<unsigned long long>(strides[i]));
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(VACNIC_ECUDA, "cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(r));
  return VACNIC_OK;
}

template <int BN, bool A_MN, bool B_MN, int BK = kBK>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmArgs& g,
                       cudaStream_t stream) {
  using Cfg = GemmCfg<BN, BK>;
  auto kern = gemm_sm100_kernel<BN, A_MN, B_MN, BK>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e =
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess)
      return fail(VACNIC_ECUDA, "gemm: cudaFuncSetAttribute(smem=%d): %s", Cfg::kSmemBytes,
                  cudaGetErrorString(e));
    configured = true;
  }
  const int sms = sm_count();
  if (sms <= 0) return fail(VACNIC_EDEVICE, "gemm: no CUDA device");
  const long long work = g.total_tiles * g.splits;
  const int grid = static_cast<int>(work < sms ? work : sms);
  launch_pdl(kern, dim3(grid), dim3(kThreads), Cfg::kSmemBytes, stream, tmA, tmB, g);
  count_launch();
  return check_last("gemm launch");
}

template <int BN>
static int dispatch_major(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmArgs& g, bool a_mn,
                          bool b_mn, cudaStream_t stream) {
  if (!a_mn && !b_mn) return launch_gemm<BN, false, false>(tmA, tmB, g, stream);
  if (!a_mn && b_mn) return launch_gemm<BN, false, true>(tmA, tmB, g, stream);
  if (a_mn && !b_mn) return launch_gemm<BN, true, false>(tmA, tmB, g, stream);
  return launch_gemm<BN, true, true>(tmA, tmB, g, stream);
}

int launch_gemm_pair(int bn, bool a_mn, bool b_mn, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmArgs& g,
                     cudaStream_t stream);  // gemm2_sm100.cu

// CTA-pair (256 x BN tiles) kernel: worth it once every SM pair gets at least one tile.  Returns 0 for "use the
// single-CTA kernel".
static int choose_pair_tile_n(const vacnic_gemm_desc* d, int sms) {
  // The CTA-pair kernel is used when some tile width fills every SM pair at least once; among the widths the
  // cheaper one by (waves x tile width) wins, with the 128-wide tile charged 15 % for its lower MMA efficiency
  // (measured: 5.9 us vs 2 x 3.7 us per 1024-deep tile).  E.g. the fc1 weight gradient (4096 x 1024 outputs):
  // 64 tiles of 256 in ONE wave beat 128 tiles of 128 in two.
  if (d->M < 256) return 0;
  const long long batch = static_cast<long long>(d->batch0) * d->batch1;
  const long long num_m = (d->M + 255) / 256;
  const long long pairs = sms / 2;
  const int cands[2] = {256, 128};
  int best = 0;
  double best_cost = 0.0;
  bool fills = false;
  for (int i = 0; i < 2; ++i) {
    const int bn = cands[i];
    if (d->N < bn) continue;
    const long long tiles = batch * num_m * ((d->N + bn - 1) / bn);
    if (tiles >= pairs) fills = true;
    const double cost = static_cast<double>((tiles + pairs - 1) / pairs) * bn * (bn == 128 ? 1.15 : 1.0);
    if (best == 0 || cost < best_cost) {
      best = bn;
      best_cost = cost;
    }
  }
  return fills ? best : 0;
}

static int choose_tile_n(const vacnic_gemm_desc* d, int sms) {
  // Prefer the widest tile that still yields at least one full wave of CTAs; tiny N gets 64.
  const long long batch = static_cast<long long>(d->batch0) * d->batch1;
  const long long num_m = (d->M + kBM - 1) / kBM;
  const int cands[3] = {256, 128, 64};
  for (int i = 0; i < 3; ++i) {
    const int bn = cands[i];
    if (d->N <= bn / 2 && bn > 64) continue;
    const long long tiles = batch * num_m * ((d->N + bn - 1) / bn);
    if (tiles >= sms || bn == 64) return bn;
  }
  return 64;
}

}  // namespace vb

extern "C" int vacnic_gemm(const vacnic_gemm_desc* d, void* stream_v) {
  using namespace vb;
  if (d == nullptr) return fail(VACNIC_EINVAL, "gemm: null descriptor");
  VB_REQUIRE(d->M > 0 && d->N > 0 && d->K > 0, "gemm: M,N,K must be positive (%d,%d,%d)", d->M, d->N, d->K);
  VB_REQUIRE(d->batch0 > 0 && d->batch1 > 0, "gemm: batch dims must be positive");
  VB_REQUIRE(d->a && d->b && d->c, "gemm: null operand pointer");
  VB_REQUIRE((reinterpret_cast<uintptr_t>(d->a) & 15) == 0 && (reinterpret_cast<uintptr_t>(d->b) & 15) == 0,
             "gemm: operands must be 16-byte aligned");
  VB_REQUIRE(d->c_dtype == VACNIC_DT_BF16 || d->c_dtype == VACNIC_DT_F32, "gemm: bad c_dtype");
  VB_REQUIRE(d->dact == VACNIC_ACT_NONE || d->aux_in != nullptr, "gemm: dact needs aux_in");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  const int sms = sm_count();
  if (sms <= 0) return fail(VACNIC_EDEVICE, "gemm: no CUDA device visible");

  // tile_n: 0 = choose; 64 / 128 / 256 = single-CTA kernel with that tile width; 1128 / 1256 = CTA-pair kernel
  int bn = d->tile_n;
  int pair_bn = 0;
  int splits = 1;
  const int num_k_blocks = (d->K + kBK - 1) / kBK;
  if (bn == 0) {
    pair_bn = choose_pair_tile_n(d, sms);
    bn = pair_bn != 0 ? pair_bn : choose_tile_n(d, sms);
    if (pair_bn == 0 && d->workspace != nullptr && num_k_blocks >= (d->split_k_min_blocks > 0 ? d->split_k_min_blocks : 4)) {
      // few output tiles and a long reduction: spread K over otherwise idle SMs (wide tiles keep the L2 -> smem
      // traffic per FLOP low, split-K supplies the parallelism)
      const int wide = d->N >= 256 ? 256 : (d->N >= 128 ? 128 : 64);
      const long long tiles = static_cast<long long>(d->batch0) * d->batch1 * ((d->M + kBM - 1) / kBM) * ((d->N + wide - 1) / wide);
      if (tiles * 2 <= sms) {
        int s = static_cast<int>(sms / tiles);
        if (s > 8) s = 8;
        if (s > num_k_blocks / 2) s = num_k_blocks / 2;
        const long long need = 65536 + tiles * s * kBM * wide * 4;
        if (s >= 2 && tiles <= 16384 && need <= d->workspace_bytes) {
          bn = wide;
          splits = s;
        }
      }
    }
  } else if (bn == 1128 || bn == 1256) {
    pair_bn = bn - 1000;
    bn = pair_bn;
  }
  VB_REQUIRE(bn == 64 || bn == 128 || bn == 256, "gemm: tile_n must be 0/64/128/256/1128/1256");

  GemmArgs g;
  g.M = d->M; g.N = d->N; g.K = d->K;
  g.batch0 = d->batch0; g.batch1 = d->batch1;
  g.num_m = pair_bn != 0 ? (d->M + 2 * kBM - 1) / (2 * kBM) : (d->M + kBM - 1) / kBM;
  g.num_n = (d->N + bn - 1) / bn;
  g.total_tiles = static_cast<long long>(d->batch0) * d->batch1 * g.num_m * g.num_n;
  g.c = d->c; g.ldc = d->ldc; g.c_sb0 = d->c_sb0; g.c_sb1 = d->c_sb1;
  g.bias = d->bias; g.aux_out = d->aux_out; g.aux_in = d->aux_in;
  g.alpha = d->alpha; g.c_dtype = d->c_dtype; g.act = d->act; g.dact = d->dact;
  g.accumulate = d->accumulate;
  g.c_chunk = d->c_chunk_stride;
  g.kb_per_split = (num_k_blocks + splits - 1) / splits;
  g.splits = (num_k_blocks + g.kb_per_split - 1) / g.kb_per_split;  // every split owns at least one k-block
  g.ws_count = static_cast<int*>(d->workspace);
  g.ws = d->workspace ? reinterpret_cast<float*>(static_cast<char*>(d->workspace) + 65536) : nullptr;
  // Vector (16 B) row access needs every row start of C / aux to be 16-byte aligned.
  const int c_es = d->c_dtype == VACNIC_DT_F32 ? 4 : 2;
  auto aligned = [&](const void* p, int es) {
    return p == nullptr ||
           ((reinterpret_cast<uintptr_t>(p) & 15) == 0 && (d->ldc * es) % 16 == 0 && (d->c_chunk_stride * es) % 16 == 0 &&
            (d->c_sb0 * es) % 16 == 0 && (d->c_sb1 * es) % 16 == 0);
  };
  g.vec_ok = aligned(d->c, c_es) && aligned(d->aux_out, 2) && aligned(d->aux_in, 2) ? 1 : 0;
  g.bias_vec = (reinterpret_cast<uintptr_t>(d->bias) & 15) == 0 ? 1 : 0;

  // a zero batch stride means "the same matrix for every batch": the map gets extent 1 and the coordinate is pinned to 0
  g.a_m0 = (d->batch0 > 1 && d->a_sb0 == 0) ? 0 : 1;
  g.a_m1 = (d->batch1 > 1 && d->a_sb1 == 0) ? 0 : 1;
  g.b_m0 = (d->batch0 > 1 && d->b_sb0 == 0) ? 0 : 1;
  g.b_m1 = (d->batch1 > 1 && d->b_sb1 == 0) ? 0 : 1;
  CUtensorMap tmA, tmB;
  int rc = make_operand_map(&tmA, d->a, d->a_mn_major != 0, d->M, d->K, d->lda, g.a_m0 ? d->batch0 : 1, d->a_sb0,
                            g.a_m1 ? d->batch1 : 1, d->a_sb1, kBM);
  if (rc != VACNIC_OK) return rc;
  rc = make_operand_map(&tmB, d->b, d->b_mn_major != 0, d->N, d->K, d->ldb, g.b_m0 ? d->batch0 : 1, d->b_sb0,
                        g.b_m1 ? d->batch1 : 1, d->b_sb1, pair_bn != 0 ? bn / 2 : bn);
  if (rc != VACNIC_OK) return rc;
  if (pair_bn != 0) return launch_gemm_pair(bn, d->a_mn_major != 0, d->b_mn_major != 0, tmA, tmB, g, stream);

  const bool a_mn = d->a_mn_major != 0, b_mn = d->b_mn_major != 0;
  if (bn == 256) return dispatch_major<256>(tmA, tmB, g, a_mn, b_mn, stream);
  if (bn == 128) return dispatch_major<128>(tmA, tmB, g, a_mn, b_mn, stream);
  if (g.splits == 1 && d->K > 128) {  // skinny, latency-bound: 128-wide pipeline stages
    if (!a_mn && !b_mn) return launch_gemm<64, false, false, 128>(tmA, tmB, g, stream);
    if (!a_mn && b_mn) return launch_gemm<64, false, true, 128>(tmA, tmB, g, stream);
    if (a_mn && !b_mn) return launch_gemm<64, true, false, 128>(tmA, tmB, g, stream);
    return launch_gemm<64, true, true, 128>(tmA, tmB, g, stream);
  }
  return dispatch_major<64>(tmA, tmB, g, a_mn, b_mn, stream);
}
