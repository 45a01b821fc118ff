// CTA-pair variant of the batched bf16 GEMM (tcgen05.mma.cta_group::2): two CTAs of a cluster — two SMs of
// one TPC — cooperate on a 256 x BN output tile.  Each CTA loads its own 128 rows of A and ONE HALF of the
// B tile; the leader CTA issues one MMA (M = 256) that reads A and the B halves from both CTAs' shared memory
// and writes 128 x BN fp32 accumulators into each CTA's TMEM.  Per SM this halves the B-operand shared-memory
// traffic and footprint (more pipeline stages), which is what lifts the large K >= 1024 GEMMs of the VACNIC
// step over the single-CTA kernel in gemm_sm100.cu (same epilogue arithmetic and descriptor semantics; bf16 results
// leave through a warp-cooperative coalesced store instead of thread-owns-row stores).
//
// Warp roles per CTA (320 threads): warp 0 TMA producer (own A rows + own B half, completing on the LEADER's
// "full" barrier), warp 1 MMA issuer (leader only) + TMEM allocator, warps 2..9 epilogue (own 128 rows).
// Cross-CTA signalling: tcgen05.commit with .multicast::cluster frees the smem stage / publishes the accumulator
// in both CTAs; the epilogue warps of both CTAs arrive on the leader's "accumulator drained" barrier.
#include <cuda.h>

#include "common.h"
#include "gemm_epilogue.cuh"
#include "ptx.cuh"

namespace vb {

template <int BN>
struct Gemm2Cfg {
  static constexpr int kABytes = kBM * kBK * 2;          // this CTA's 128 rows of A
  static constexpr int kBBytes = (BN / 2) * kBK * 2;     // this CTA's half of the B tile
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (BN == 256) ? 6 : 8;
  static constexpr int kTmemCols = 2 * BN;               // double-buffered accumulator: 512 / 256
  static constexpr int kBarBytes = 256;
  static constexpr int kBiasBytes = kEpiWarps * (BN / 2) * 4;  // each epilogue warp's bias slice of the current tile
  static constexpr int kStoreBytes = kEpiWarps * 32 * 64;      // each epilogue warp's 32 x 32 bf16 transpose buffer
  static constexpr int kSmemBytes = kStages * kStageBytes + kBarBytes + kBiasBytes + kStoreBytes + 1024;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;  // clears the CTA-rank bit of a shared::cluster address -> CTA 0 of the pair

// 4-D tiled load into THIS CTA's shared memory; the transaction bytes complete on the leader CTA's mbarrier.
__device__ __forceinline__ void tma_load_4d_2sm(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                                int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0),
      "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2sm() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_ss_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// All prior MMAs of this thread arrive (count 1) on the barrier at this offset in BOTH CTAs of the pair.
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(static_cast<uint16_t>(3))
      : "memory");
}
// Arrive on the barrier at this offset in the LEADER CTA.  Relaxed: the only thing the MMA thread may not overtake
// is this warp's TMEM reads, which tcgen05.wait::ld + tcgen05.fence::before_thread_sync have already retired; a
// release at cluster scope would also drain the warp's global stores (MEMBAR.ALL.GPU + ERRBAR, ~10 % of the
// kernel's stall samples in profiles/r1_ncu_gemm2_fc1_gelu.md) for no reason.
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile(
      "{\n\t"
      ".reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, 0;\n\t"
      "mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [ra];\n\t"
      "}\n" ::"r"(smem_u32(bar))
      : "memory");
}

template <int BN, bool A_MN, bool B_MN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
gemm2_sm100_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmArgs g) {
  using Cfg = Gemm2Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);  // same offset in both CTAs (same kernel image)
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + Cfg::kStages;
  uint64_t* tfull_bar = empty_bar + Cfg::kStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int num_k = (g.K + kBK - 1) / kBK;
  // g.num_m counts 256-row tiles here
  const long long tiles_per_batch = static_cast<long long>(g.num_m) * g.num_n;
  const long long first_tile = blockIdx.x >> 1;
  const long long tile_stride = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full_bar[s], 1);   // leader's producer arms it with the bytes of both CTAs
      mbar_init(&empty_bar[s], 1);  // multicast commit from the leader's MMA thread
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 2 * kEpiWarps);  // epilogue warps of BOTH CTAs (used on the leader only)
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc_2sm(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_sync();  // everything above is on-chip setup; operands, bias, aux and C are touched only below

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (both CTAs)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (long long t = first_tile; t < g.total_tiles; t += tile_stride) {
        const int batch = static_cast<int>(t / tiles_per_batch);
        const int r = static_cast<int>(t % tiles_per_batch);
        const int m0 = (r % g.num_m) * (2 * kBM) + static_cast<int>(rank) * kBM;
        const int n0 = (r / g.num_m) * BN + static_cast<int>(rank) * (BN / 2);
        const int b0 = batch % g.batch0;
        const int b1 = batch / g.batch0;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          uint8_t* sa = smem + stage * Cfg::kStageBytes;
          uint8_t* sb = sa + Cfg::kABytes;
          if (leader) mbar_expect_tx(&full_bar[stage], 2 * Cfg::kStageBytes);
          const int k0 = kb * kBK;
          if constexpr (!A_MN) {
            tma_load_4d_2sm(sa, &tmA, &full_bar[stage], k0, m0, b0 * g.a_m0, b1 * g.a_m1);
          } else {
#pragma unroll
            for (int c = 0; c < kBM / 64; ++c)
              tma_load_4d_2sm(sa + c * (64 * kBK * 2), &tmA, &full_bar[stage], m0 + c * 64, k0, b0 * g.a_m0, b1 * g.a_m1);
          }
          if constexpr (!B_MN) {
            tma_load_4d_2sm(sb, &tmB, &full_bar[stage], k0, n0, b0 * g.b_m0, b1 * g.b_m1);
          } else {
#pragma unroll
            for (int c = 0; c < BN / 128; ++c)
              tma_load_4d_2sm(sb + c * (64 * kBK * 2), &tmB, &full_bar[stage], n0 + c * 64, k0, b0 * g.b_m0, b1 * g.b_m1);
          }
          if (++stage == Cfg::kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (leader CTA only)
    if (leader && lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(2 * kBM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      constexpr uint32_t kLboA = A_MN ? kBK * 128 : 16, kLboB = B_MN ? kBK * 128 : 16;
      constexpr uint32_t kStepA = A_MN ? 16 * 128 : 32, kStepB = B_MN ? 16 * 128 : 32;
      int stage = 0;
      uint32_t phase = 0;
      uint32_t it = 0;
      for (long long t = first_tile; t < g.total_tiles; t += tile_stride, ++it) {
        const uint32_t as = it & 1u;
        const uint32_t aphase = (it >> 1) & 1u;
        mbar_wait(&tempty_bar[as], aphase ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * BN;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint32_t sb = sa + Cfg::kABytes;
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            const uint64_t da = make_smem_desc_sw128(sa + k * kStepA, kLboA, 1024);
            const uint64_t db = make_smem_desc_sw128(sb + k * kStepB, kLboB, 1024);
            umma_bf16_ss_2sm(tmem_d, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit_2sm(&empty_bar[stage]);  // frees this smem stage in both CTAs when the MMAs retire
          if (++stage == Cfg::kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        umma_commit_2sm(&tfull_bar[as]);  // accumulator complete -> epilogue warps of both CTAs
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue warps 2..9 (own 128 rows)
    // Per tile: each warp stages the bias of its BN/2 columns in its own shared-memory slice while the accumulator
    // is still being produced (no CTA-level barrier), then drains its 32 rows x BN/2 columns in 32-column chunks with the
    // tcgen05.ld of chunk i+1 in flight behind the arithmetic and stores of chunk i, and hands the accumulator
    // back to the MMA thread as soon as its last chunk is in registers (before that chunk is computed / stored).
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    float* sbias = reinterpret_cast<float*>(smem + Cfg::kStages * Cfg::kStageBytes + Cfg::kBarBytes) +
                   (warp - 2) * (BN / 2);
    const uint32_t sstore =
        smem_u32(smem + Cfg::kStages * Cfg::kStageBytes + Cfg::kBarBytes + Cfg::kBiasBytes + (warp - 2) * (32 * 64));
    constexpr int kChunks = BN / 64;  // 32-column chunks per warp
    const bool has_bias = g.bias != nullptr;
    uint32_t it = 0;
    for (long long t = first_tile; t < g.total_tiles; t += tile_stride, ++it) {
      const int batch = static_cast<int>(t / tiles_per_batch);
      const int r = static_cast<int>(t % tiles_per_batch);
      const int m0 = (r % g.num_m) * (2 * kBM) + static_cast<int>(rank) * kBM;
      const int n0 = (r / g.num_m) * BN;
      const int b0 = batch % g.batch0;
      const int b1 = batch / g.batch0;
      const uint32_t as = it & 1u;
      const uint32_t aphase = (it >> 1) & 1u;
      if (has_bias) {
        __syncwarp();  // every lane is done reading the previous tile's slice
#pragma unroll
        for (int j = lane; j < BN / 2; j += 32) {
          const int n = n0 + half * (BN / 2) + j;
          sbias[j] = n < g.N ? __ldg(g.bias + n) : 0.0f;
        }
        __syncwarp();
      }
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      const int row = m0 + quad * 32 + lane;
      const long long row_off = static_cast<long long>(b0) * g.c_sb0 + static_cast<long long>(b1) * g.c_sb1 +
                                static_cast<long long>(row) * g.ldc;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + as * BN;
      const int c0 = half * kChunks;
      const long long warp_off = static_cast<long long>(b0) * g.c_sb0 + static_cast<long long>(b1) * g.c_sb1 +
                                 static_cast<long long>(m0 + quad * 32) * g.ldc;
      const int rows_valid = g.M - (m0 + quad * 32);
      const bool row_ok = lane < rows_valid;
      // bf16 results without read-modify-write leave through the warp-cooperative coalesced store
      const bool coalesced = g.c_dtype == VACNIC_DT_BF16 && !g.accumulate;
      auto process = [&](const uint32_t (&buf)[32], int i) {
        const int n = n0 + (c0 + i) * 32;
        if (n >= g.N) return;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(buf[j]);
        const float* sb = has_bias ? sbias + i * 32 : nullptr;
        const long long coff = epilogue_col_off(g, n);
        if (epilogue_chunk_is_vec(g, n, sb)) {
          epilogue_bias_alpha(g, v, n, sb);
          if (g.aux_out != nullptr) {
            __nv_bfloat16* aux = reinterpret_cast<__nv_bfloat16*>(g.aux_out);
            if (coalesced) store_chunk_bf16_coalesced(sstore, lane, v, aux + warp_off + coff, g.ldc, rows_valid);
            else if (row_ok) store_row_bf16_vec(aux + row_off + coff, v);
          }
          epilogue_act(g, v);
          if (g.dact != VACNIC_ACT_NONE && row_ok) epilogue_dact_vec(g, v, row_off + coff);
          if (coalesced)
            store_chunk_bf16_coalesced(sstore, lane, v, reinterpret_cast<__nv_bfloat16*>(g.c) + warp_off + coff, g.ldc,
                                       rows_valid);
          else if (row_ok)
            epilogue_store_vec(g, v, row_off + coff);
        } else if (row_ok) {
          epilogue_chunk_ragged_call(g, v, row_off + coff, n, sb);
        }
      };
      uint32_t buf_a[32], buf_b[32];
      tmem_ld_32x32(taddr + c0 * 32, buf_a);
#pragma unroll 1
      for (int i = 0; i < kChunks; i += 2) {  // kChunks is 2 or 4: the body handles one chunk pair
        tmem_ld_wait();
        tmem_ld_32x32(taddr + (c0 + i + 1) * 32, buf_b);
        process(buf_a, i);
        tmem_ld_wait();
        if (i + 2 < kChunks) {
          tmem_ld_32x32(taddr + (c0 + i + 2) * 32, buf_a);
        } else {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_leader(&tempty_bar[as]);
        }
        process(buf_b, i + 1);
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();  // both CTAs are done with each other's shared memory and TMEM
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, Cfg::kTmemCols);
  }
}

template <int BN, bool A_MN, bool B_MN>
static int launch_gemm2(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmArgs& g, cudaStream_t stream) {
  using Cfg = Gemm2Cfg<BN>;
  auto kern = gemm2_sm100_kernel<BN, A_MN, B_MN>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess)
      return fail(VACNIC_ECUDA, "gemm2: cudaFuncSetAttribute(smem=%d): %s", Cfg::kSmemBytes, cudaGetErrorString(e));
    configured = true;
  }
  const int sms = sm_count();
  if (sms <= 0) return fail(VACNIC_EDEVICE, "gemm2: no CUDA device");
  const long long pairs = sms / 2;
  const int grid = 2 * static_cast<int>(g.total_tiles < pairs ? g.total_tiles : pairs);
  launch_pdl(kern, dim3(grid), dim3(kThreads), Cfg::kSmemBytes, stream, tmA, tmB, g);
  count_launch();
  return check_last("gemm2 launch");
}

// Entry used by vacnic_gemm (gemm_sm100.cu) for large problems.  `g.num_m` / `g.total_tiles` are in 256-row tiles.
int launch_gemm_pair(int bn, bool a_mn, bool b_mn, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmArgs& g,
                     cudaStream_t stream) {
#define VB_G2(BNV)                                                                     \
  do {                                                                                 \
    if (!a_mn && !b_mn) return launch_gemm2<BNV, false, false>(tmA, tmB, g, stream);   \
    if (!a_mn && b_mn) return launch_gemm2<BNV, false, true>(tmA, tmB, g, stream);     \
    if (a_mn && !b_mn) return launch_gemm2<BNV, true, false>(tmA, tmB, g, stream);     \
    return launch_gemm2<BNV, true, true>(tmA, tmB, g, stream);                         \
  } while (0)
  if (bn == 256) VB_G2(256);
  VB_G2(128);
#undef VB_G2
}

}  // namespace vb
