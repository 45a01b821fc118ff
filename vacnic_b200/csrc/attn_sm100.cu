// Fused scaled-dot-product attention for sm_100a (BartAttention.forward core, MFULL:503-556: bmm(q,k^T) +
// mask -> softmax -> bmm(p,v)), head_dim 64.  Scores never leave the SM: S = Q K^T is produced by
// tcgen05.mma into TMEM, read back with tcgen05.ld by the softmax warps (one thread per query row, so no
// cross-thread reductions), probabilities go to 128B-swizzled shared memory as bf16 and feed the second
// tcgen05.mma (O += P V) whose accumulator also lives in TMEM.
//
// One sweep over the key blocks with a LAZILY updated reference maximum: the first block fixes m_ref = its row maximum;
// a later block only raises m_ref (and rescales the running sum and the O accumulator in TMEM, tcgen05.ld/st) when its
// maximum exceeds m_ref by more than 2^8 — rare after the first blocks — otherwise p = 2^(s - m_ref) <= 256 is simply
// carried in bf16 / fp32.  The result is the exact softmax (any reference point cancels in O / l).  Masks follow the
// reference: masked keys get
// finfo(float32).min (a fully masked row degenerates to a uniform row exactly like the reference),
// keys beyond `key_len[b]` are skipped because they contribute exactly 0.
//
// CTA = 128 query rows of one (batch, head); 320 threads: warp 0 TMA producer, warp 1 MMA issuer + TMEM
// allocator, warps 2..9 softmax/epilogue (thread = one query row x one half of the key block's columns; the
// two halves of a row exchange their block maxima through shared memory).  96 KB of shared memory and 256
// TMEM columns per CTA, so two CTAs share an SM and overlap each other's MMA and exp phases.
#include <cuda.h>
#include <float.h>

#include "common.h"
#include "ptx.cuh"

namespace vb {

constexpr int kAQ = 128;         // query rows per CTA
constexpr int kAK = 128;         // keys per block
constexpr int kHD = 64;          // head dim
constexpr int kTile = kAK * kHD * 2;  // 16 KB: one [128 x 64] bf16 tile
constexpr int kRing = 3;
constexpr int kCompWarps = 8;                        // softmax / elementwise warps: 2 per TMEM lane quadrant
constexpr int kCompThreads = 32 * kCompWarps;         // each thread: one row, one half of the block's columns
constexpr int kAttnThreads = 64 + kCompThreads;
constexpr int kMaxKB = 40;       // key blocks per row (Sk <= 5120)

struct AttnArgs {
  int B, H, Sq, Sk;
  int causal;
  float scale_log2;  // head_dim^-0.5 * log2(e)
  const uint8_t* key_mask;  // [B][Sk], 1 = attend, or null
  const int32_t* key_len;   // [B] or null
  __nv_bfloat16* out;       // out[b*o_sb + row*ldo + h*64 + c]
  long long ldo, o_sb;
  float* stats;             // [B][H][Sq][2] = {row max in the log2 domain, 1 / row sum}, or null
  // packed (varlen) mode, active when q_start != null: sequence b owns query rows [q_start[b], +q_len[b]) of a buffer of
  // total_q rows and key rows [k_start[b], +k_len[b]) of a buffer of total_k rows; the batch strides / Sq / Sk above are
  // then only upper bounds (grid sizing) and stats is [H][total_q][2]
  const int32_t* q_start; const int32_t* q_len; const int32_t* k_start; const int32_t* k_len;
  int total_q;
  // attention dropout (config.attention_dropout, MFULL:546): probabilities are dropped AFTER the softmax normalisation;
  // counter-based mask keyed by (*rng, salt, (sequence, head, query row), key) -- the backward kernels regenerate it
  float p_drop; const unsigned long long* rng; uint32_t salt;
};

// keep decision of attention dropout for (query row id = geo.sbase + row, key index inside the sequence)
__device__ __forceinline__ bool keep_attn(uint32_t seed, uint32_t salt, long long rowid, int key, uint32_t thr) {
  return keep_elem(seed, salt, (static_cast<uint64_t>(rowid) << 16) | static_cast<uint32_t>(key), thr);
}

// per-CTA view of "its" sequence: row counts, first rows in the (possibly packed) buffers, batch coordinate for TMA
struct SeqGeo {
  int Sq, Sk, qr0, kr0, bq;
  long long sbase;  // index of query row 0 of this (sequence, head) in stats / delta
};
template <typename Args>
__device__ __forceinline__ SeqGeo seq_geo(const Args& a, int b, int h) {
  SeqGeo g;
  if (a.q_start != nullptr) {
    g.qr0 = a.q_start[b]; g.Sq = a.q_len[b]; g.kr0 = a.k_start[b]; g.Sk = a.k_len[b]; g.bq = 0;
    g.sbase = static_cast<long long>(h) * a.total_q + g.qr0;
  } else {
    g.qr0 = 0; g.kr0 = 0; g.Sq = a.Sq; g.Sk = a.Sk; g.bq = b;
    g.sbase = (static_cast<long long>(b) * a.H + h) * a.Sq;
  }
  return g;
}

struct AttnSmem {
  uint64_t q_full, ring_full[kRing], ring_empty[kRing], s_full, s_empty, p_full, p_empty, o_full;
  uint32_t tmem_slot;
  uint8_t blk_flag[kMaxKB];  // per key block: 0 = no masking needed, 1 = per-key checks needed
  float xch[2][2][128];      // [block parity][column half][row]: block maxima / final row sums exchanged between halves
};

// named barrier 2 + quad over the two softmax warps (column halves) that own the same 32 query rows
__device__ __forceinline__ void pair_bar_sync(int quad) {  // immediate ids: ptxas reserves only the barriers named
  switch (quad) {
    case 0: asm volatile("bar.sync 2, 64;" ::: "memory"); break;
    case 1: asm volatile("bar.sync 3, 64;" ::: "memory"); break;
    case 2: asm volatile("bar.sync 4, 64;" ::: "memory"); break;
    default: asm volatile("bar.sync 5, 64;" ::: "memory"); break;
  }
}
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// GENERAL = false: no padding-mask bytes and no causal mask (packed rows, prefix / CLIP attention): only the sequence end can
// cut a key block.  Keeping the byte-load path out of this instantiation is what keeps it free of local-memory spills.
template <bool GENERAL, bool DROP>
__global__ void __launch_bounds__(kAttnThreads, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const AttnArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sQ = smem;
  uint8_t* sRing = smem + kTile;
  uint8_t* sP = sRing + kRing * kTile;  // [2 atoms][128 rows][128 B]
  AttnSmem* sh = reinterpret_cast<AttnSmem*>(sP + 2 * kTile);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qb = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int q0 = qb * kAQ;
  pdl_sync();  // key_len / key_mask / sequence geometry below are global reads
  const SeqGeo geo = seq_geo(a, b, h);
  if (q0 >= geo.Sq) return;  // packed mode: this sequence has no query rows in this tile (uniform for the whole CTA)

  // number of key blocks that can contribute
  int kmax = geo.Sk;
  if (a.key_len != nullptr && a.q_start == nullptr) {
    const int kl = a.key_len[b];
    if (kl > 0 && kl < kmax) kmax = kl;  // kl == 0: fully masked rows are uniform over ALL keys -> keep every block
  }
  if (a.causal) kmax = min(kmax, q0 + kAQ);
  const int nkb = (kmax + kAK - 1) / kAK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(&sh->q_full, 1);
    for (int s = 0; s < kRing; ++s) {
      mbar_init(&sh->ring_full[s], 1);
      mbar_init(&sh->ring_empty[s], 1);
    }
    mbar_init(&sh->s_full, 1);
    mbar_init(&sh->s_empty, kCompThreads);
    mbar_init(&sh->p_full, kCompThreads);
    mbar_init(&sh->p_empty, 1);
    mbar_init(&sh->o_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(&sh->tmem_slot, 256);
    tmem_relinquish();
  }
  if (warp >= 2) {
    // classify the key blocks once: does any key of the block need a mask decision?  Geometry first, then the mask
    // bytes in 16-key segments spread over all 256 threads (one thread scanning a block's 128 bytes kept the whole
    // CTA at the barrier below for ~10 % of its life, profiles/r1_ncu_attention_encoder_shape.md)
    const int tid = static_cast<int>(threadIdx.x) - 64;
    for (int j = tid; j < nkb; j += kCompThreads) {
      const int k0 = j * kAK;
      sh->blk_flag[j] = ((k0 + kAK > geo.Sk) || (a.causal && k0 + kAK - 1 > q0)) ? 1 : 0;
    }
    if (a.key_mask != nullptr) {  // (packed mode and mask-free call sites skip the scan and its barrier)
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const uint8_t* mk = a.key_mask + static_cast<long long>(b) * a.Sk;
      const int kend = min(nkb * kAK, geo.Sk);
      for (int k = tid * 16; k < kend; k += kCompThreads * 16) {
        bool z = false;
#pragma unroll
        for (int c = 0; c < 16; ++c) z |= (k + c < kend) && (mk[k + c] == 0);
        if (z) sh->blk_flag[k / kAK] = 1;  // racing writers all store 1
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sh->tmem_slot;
  const uint32_t tmem_s = tmem_base;         // 128 columns
  const uint32_t tmem_o = tmem_base + 128;   // 64 columns

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      mbar_expect_tx(&sh->q_full, kTile);
      tma_load_4d(sQ, &tmQ, &sh->q_full, 0, geo.qr0 + q0, h, geo.bq);
      int slot = 0;
      uint32_t phase = 0;
      auto push = [&](const CUtensorMap* tm, int row0) {
        mbar_wait(&sh->ring_empty[slot], phase ^ 1u);
        mbar_expect_tx(&sh->ring_full[slot], kTile);
        tma_load_4d(sRing + slot * kTile, tm, &sh->ring_full[slot], 0, geo.kr0 + row0, h, geo.bq);
        if (++slot == kRing) { slot = 0; phase ^= 1u; }
      };
      // consumption order of the MMA thread, which runs QK^T one block ahead of PV: K0, K1, V0, K2, V1, ...
      push(&tmK, 0);
      for (int j = 0; j < nkb; ++j) {
        if (j + 1 < nkb) push(&tmK, (j + 1) * kAK);
        push(&tmV, j * kAK);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(kAQ, kAK, 0, 0);
      constexpr uint32_t idesc_o = make_idesc_bf16(kAQ, kHD, 0, 1);
      int slot = 0;
      uint32_t phase = 0;
      uint32_t it = 0;
      mbar_wait(&sh->q_full, 0);
      const uint32_t q_addr = smem_u32(sQ);
      const uint32_t p_addr = smem_u32(sP);
      auto issue_s = [&]() {  // S = Q K^T for the next key block (waits for its K tile and for S to be drained)
        mbar_wait(&sh->ring_full[slot], phase);
        mbar_wait(&sh->s_empty, (it & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t k_addr = smem_u32(sRing + slot * kTile);
#pragma unroll
        for (int k = 0; k < kHD / 16; ++k)
          umma_bf16_ss(tmem_s, make_smem_desc_sw128(q_addr + k * 32, 16, 1024),
                       make_smem_desc_sw128(k_addr + k * 32, 16, 1024), idesc_s, k != 0 ? 1u : 0u);
        umma_commit(&sh->ring_empty[slot]);
        umma_commit(&sh->s_full);
        if (++slot == kRing) { slot = 0; phase ^= 1u; }
        ++it;
      };
      issue_s();
      for (int j = 0; j < nkb; ++j) {
        // QK^T of block j+1 is issued as soon as the softmax warps have pulled S_j out of TMEM, i.e. it runs while they
        // compute the exponentials of block j
        if (j + 1 < nkb) issue_s();
        mbar_wait(&sh->ring_full[slot], phase);
        mbar_wait(&sh->p_full, j & 1u);
        tc_fence_after();
        const uint32_t v_addr = smem_u32(sRing + slot * kTile);
#pragma unroll
        for (int kk = 0; kk < kAK / 16; ++kk)
          umma_bf16_ss(tmem_o, make_smem_desc_sw128(p_addr + (kk >> 2) * kTile + (kk & 3) * 32, 16, 1024),
                       make_smem_desc_sw128(v_addr + kk * 2048, kAK * 128, 1024), idesc_o, (j | kk) != 0 ? 1u : 0u);
        umma_commit(&sh->ring_empty[slot]);
        umma_commit(&sh->p_empty);
        if (++slot == kRing) { slot = 0; phase ^= 1u; }
      }
      umma_commit(&sh->o_full);
    }
  } else {
    // ------------------------------------------------------------ softmax + epilogue
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;  // which 64 of the block's 128 key columns
    const int r = quad * 32 + lane;    // row inside the tile == TMEM lane
    const int row = q0 + r;
    const uint32_t t_s = tmem_s + (static_cast<uint32_t>(quad * 32) << 16) + half * 64;
    const uint8_t* mk = a.key_mask ? a.key_mask + static_cast<long long>(b) * a.Sk : nullptr;
    const float c2 = a.scale_log2;
    uint32_t it = 0;
    // Sequence-end masking (both variants) happens in the RAW score domain, in place -- a key past the end becomes -inf
    // (probability exactly 0) and the scale rides on the FFMA in front of the exponential -- so that cut and uncut blocks
    // leave the scores in the same registers (scaling in place made the compiler shuffle all 64 of them on the uncut
    // path).  Padding-mask bytes and the causal mask (GENERAL variant) are applied to the scaled score: a masked key gets
    // -FLT_MAX = finfo.min, so a fully masked row is uniform like the reference.
    auto masked = [&](float s, int key) -> float {  // GENERAL variant only: scaled + masked score (log2 domain)
      if (key >= geo.Sk) return -INFINITY;
      if ((mk != nullptr && mk[key] == 0) || (a.causal && key > row)) return -FLT_MAX;
      return s * c2;
    };
    // ---- single sweep: block maximum -> (rare) rescale -> p = 2^(t - m_ref), row sum, P -> shared memory
    float m = -INFINITY;  // m_ref
    float l = 0.f;
    const uint32_t atom = smem_u32(sP) + half * kTile + r * 128;
    const uint32_t sw = static_cast<uint32_t>(r & 7);
    const uint32_t t_o = tmem_o + (static_cast<uint32_t>(quad * 32) << 16) + half * 32;  // this thread's 32 O columns
    const uint32_t dseed = DROP ? static_cast<uint32_t>(*a.rng) : 0u, dthr = drop_thresh(a.p_drop);
    const float dinv = DROP ? 1.f / (1.f - a.p_drop) : 1.f;
    for (int j = 0; j < nkb; ++j, ++it) {
      mbar_wait(&sh->s_full, it & 1u);
      tc_fence_after();
      const bool need = sh->blk_flag[j] != 0;
      const int kb = j * kAK + half * 64;
      uint32_t x0[32], x1[32];
      tmem_ld_32x32(t_s, x0);
      tmem_ld_32x32(t_s + 32, x1);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&sh->s_empty);  // all of S_j is in registers: the tensor pipe may overwrite it
      float bm = -INFINITY, kk = c2;
      bool have_max = false;
      if (need) {
        if (!GENERAL || (mk == nullptr && !a.causal)) {
          // only the sequence end can cut this block (packed rows / no padding mask): columns >= nv are out of range
          const int nv = geo.Sk - kb;
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            if (c >= nv) x0[c] = 0xff800000u;       // -inf
            if (c + 32 >= nv) x1[c] = 0xff800000u;
          }
        } else if constexpr (GENERAL) {
          // padding-mask bytes / causal mask: scale and mask in place (kk = 1 below)
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            const float t0 = masked(__uint_as_float(x0[c]), kb + c), t1 = masked(__uint_as_float(x1[c]), kb + 32 + c);
            x0[c] = __float_as_uint(t0);
            x1[c] = __float_as_uint(t1);
            bm = fmaxf(bm, fmaxf(t0, t1));
          }
          kk = 1.0f;
          have_max = true;
        }
      }
      if (!have_max) {
        // block maximum of this half in the log2 domain (c2 > 0, so max(raw) * c2 == max(raw * c2))
#pragma unroll
        for (int c = 0; c < 32; ++c) bm = fmaxf(bm, fmaxf(__uint_as_float(x0[c]), __uint_as_float(x1[c])));
        bm *= c2;
      }
      sh->xch[j & 1][half][r] = bm;
      pair_bar_sync(quad);  // only the two warps that share these 32 rows exchange maxima
      bm = fmaxf(bm, sh->xch[j & 1][half ^ 1][r]);
      // reference maximum: fixed by block 0, raised later only where a row's block maximum exceeds it by more than 2^8
      // (both halves of a row see the same maxima, so they agree)
      float f = 1.0f;
      bool mine = false;
      if (j == 0) {
        m = bm;
      } else if (bm > m + 8.0f) {
        mine = true;
        f = ex2f(m - bm);
        m = bm;
        l *= f;
      }
      // exponentials, packed to bf16 pairs right away (the fp32 scores die as they are consumed: 32 live registers of
      // packed P instead of 64 of fp32 P -- the 64-register version spilled to local memory), before waiting for PV_{j-1}
      uint32_t pk[32];
      float l0 = 0.f, l1 = 0.f;
#pragma unroll
      for (int c = 0; c < 32; c += 2) {
        const float a0 = ex2f(fmaf(__uint_as_float(x0[c]), kk, -m)), a1 = ex2f(fmaf(__uint_as_float(x0[c + 1]), kk, -m));
        const float b0 = ex2f(fmaf(__uint_as_float(x1[c]), kk, -m)), b1 = ex2f(fmaf(__uint_as_float(x1[c + 1]), kk, -m));
        l0 += a0 + a1;  // the row sum normalises the UNdropped probabilities (dropout follows the softmax, MFULL:546)
        l1 += b0 + b1;
        if constexpr (DROP) {
          const long long rid = geo.sbase + row;
          pk[c >> 1] = pack_bf16x2(keep_attn(dseed, a.salt, rid, kb + c, dthr) ? a0 * dinv : 0.f,
                                   keep_attn(dseed, a.salt, rid, kb + c + 1, dthr) ? a1 * dinv : 0.f);
          pk[16 + (c >> 1)] = pack_bf16x2(keep_attn(dseed, a.salt, rid, kb + 32 + c, dthr) ? b0 * dinv : 0.f,
                                          keep_attn(dseed, a.salt, rid, kb + 32 + c + 1, dthr) ? b1 * dinv : 0.f);
        } else {
          pk[c >> 1] = pack_bf16x2(a0, a1);
          pk[16 + (c >> 1)] = pack_bf16x2(b0, b1);
        }
      }
      l += l0 + l1;
      if (j > 0) {
        mbar_wait(&sh->p_empty, (j - 1) & 1u);  // PV_{j-1} retired: P buffer free, O quiescent
        tc_fence_after();
        // tcgen05.ld/st are warp-collective: the whole warp rescales its O slices if any lane raised its maximum
        // (the other lanes multiply by exactly 1)
        if (__any_sync(0xffffffffu, mine)) {
          uint32_t o[32];
          tmem_ld_32x32(t_o, o);
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 32; ++c) o[c] = __float_as_uint(__uint_as_float(o[c]) * f);
          tmem_st_32x32(t_o, o);
          tmem_st_wait();
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q)
        st_shared_v4(atom + ((static_cast<uint32_t>(q) ^ sw) << 4), pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
#pragma unroll
      for (int q = 0; q < 4; ++q)
        st_shared_v4(atom + ((static_cast<uint32_t>(4 + q) ^ sw) << 4), pk[16 + 4 * q], pk[16 + 4 * q + 1], pk[16 + 4 * q + 2],
                     pk[16 + 4 * q + 3]);
      tc_fence_before();
      fence_proxy_async_smem();  // generic-proxy writes of P -> visible to the tensor pipe (async proxy)
      mbar_arrive(&sh->p_full);
    }
    pair_bar_sync(quad);
    sh->xch[0][half][r] = l;
    pair_bar_sync(quad);
    l += sh->xch[0][half ^ 1][r];
    // ---- epilogue: O / l -> bf16 (each half stores 32 of the 64 head-dim columns)
    mbar_wait(&sh->o_full, 0);
    tc_fence_after();
    const float inv = 1.f / l;
    uint32_t o0[32];
    tmem_ld_32x32(tmem_o + (static_cast<uint32_t>(quad * 32) << 16) + half * 32, o0);
    tmem_ld_wait();
    if (row < geo.Sq) {
      __nv_bfloat16* op = a.out + static_cast<long long>(geo.bq) * a.o_sb + static_cast<long long>(geo.qr0 + row) * a.ldo + h * kHD + half * 32;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint4 u;
        u.x = pack_bf16x2(__uint_as_float(o0[8 * q + 0]) * inv, __uint_as_float(o0[8 * q + 1]) * inv);
        u.y = pack_bf16x2(__uint_as_float(o0[8 * q + 2]) * inv, __uint_as_float(o0[8 * q + 3]) * inv);
        u.z = pack_bf16x2(__uint_as_float(o0[8 * q + 4]) * inv, __uint_as_float(o0[8 * q + 5]) * inv);
        u.w = pack_bf16x2(__uint_as_float(o0[8 * q + 6]) * inv, __uint_as_float(o0[8 * q + 7]) * inv);
        reinterpret_cast<uint4*>(op)[q] = u;
      }
      if (a.stats != nullptr && half == 0)
        reinterpret_cast<float2*>(a.stats)[geo.sbase + row] = make_float2(m, inv);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}


// =================================================================================================
// Backward.  Two kernels, both "thread = row of a 128-row tile", both with 256 TMEM columns and two CTAs
// per SM, both recomputing P from the saved row statistics {max, 1/sum}:
//   attn_bwd_dq_kernel   CTA = 128 queries; per 64-key block: S = Q K^T, dP = dO V^T (TMEM) ->
//                        dS = P * (dP - delta) * scale (bf16, shared memory) -> dQ += dS K.  Also computes
//                        delta = rowsum(dO * O) and stores it for the second kernel.
//   attn_bwd_dkv_kernel  CTA = 128 keys; per 64-query block: S^T = K Q^T, dP^T = V dO^T -> P^T and dS^T
//                        (shared memory) -> dV += P^T dO, dK += dS^T Q.
// Nothing is accumulated with atomics: results are deterministic.
// =================================================================================================
constexpr int kHalfTile = 64 * kHD * 2;  // 8 KB: one [64 x 64] bf16 tile

struct AttnBwdArgs {
  int B, H, Sq, Sk;
  int causal;
  float scale, scale_log2;
  const uint8_t* key_mask;
  const int32_t* key_len;
  const float* stats;   // [B][H][Sq][2]
  float* delta;         // [B][H][Sq]
  __nv_bfloat16* dq; long long lddq, dq_sh, dq_sb;
  __nv_bfloat16* dk; long long lddk, dk_sh, dk_sb;
  __nv_bfloat16* dv; long long lddv, dv_sh, dv_sb;
  const int32_t* q_start; const int32_t* q_len; const int32_t* k_start; const int32_t* k_len;  // packed mode, see AttnArgs
  int total_q;
  float p_drop; const unsigned long long* rng; uint32_t salt;  // attention dropout, see AttnArgs
};

struct BwdSmem {
  uint64_t in_full, st_full[2], st_empty[2], sdp_full, s_empty, ds_full, ds_empty, acc_full;
  uint32_t tmem_slot;
  float4 stat[2][64];    // dkv kernel: per query of the block {max, 1/sum, delta, scale/sum}
  uint8_t blk_flag[2 * kMaxKB];  // dq kernel: per 64-key block, 1 = per-key mask checks needed
};

__device__ __forceinline__ void store_row64(__nv_bfloat16* op, const uint32_t (&o0)[32], const uint32_t (&o1)[32], float f) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint4 u;
    u.x = pack_bf16x2(__uint_as_float(o0[8 * q + 0]) * f, __uint_as_float(o0[8 * q + 1]) * f);
    u.y = pack_bf16x2(__uint_as_float(o0[8 * q + 2]) * f, __uint_as_float(o0[8 * q + 3]) * f);
    u.z = pack_bf16x2(__uint_as_float(o0[8 * q + 4]) * f, __uint_as_float(o0[8 * q + 5]) * f);
    u.w = pack_bf16x2(__uint_as_float(o0[8 * q + 6]) * f, __uint_as_float(o0[8 * q + 7]) * f);
    reinterpret_cast<uint4*>(op)[q] = u;
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint4 u;
    u.x = pack_bf16x2(__uint_as_float(o1[8 * q + 0]) * f, __uint_as_float(o1[8 * q + 1]) * f);
    u.y = pack_bf16x2(__uint_as_float(o1[8 * q + 2]) * f, __uint_as_float(o1[8 * q + 3]) * f);
    u.z = pack_bf16x2(__uint_as_float(o1[8 * q + 4]) * f, __uint_as_float(o1[8 * q + 5]) * f);
    u.w = pack_bf16x2(__uint_as_float(o1[8 * q + 6]) * f, __uint_as_float(o1[8 * q + 7]) * f);
    reinterpret_cast<uint4*>(op)[4 + q] = u;
  }
}

__device__ __forceinline__ void store_row32(__nv_bfloat16* op, const uint32_t (&o0)[32]) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint4 u;
    u.x = pack_bf16x2(__uint_as_float(o0[8 * q + 0]), __uint_as_float(o0[8 * q + 1]));
    u.y = pack_bf16x2(__uint_as_float(o0[8 * q + 2]), __uint_as_float(o0[8 * q + 3]));
    u.z = pack_bf16x2(__uint_as_float(o0[8 * q + 4]), __uint_as_float(o0[8 * q + 5]));
    u.w = pack_bf16x2(__uint_as_float(o0[8 * q + 6]), __uint_as_float(o0[8 * q + 7]));
    reinterpret_cast<uint4*>(op)[q] = u;
  }
}

__device__ __forceinline__ void bwd_init(BwdSmem* sh, int warp, int lane) {
  if (warp == 0 && lane == 0) {
    mbar_init(&sh->in_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&sh->st_full[s], 1);
      mbar_init(&sh->st_empty[s], 1);
    }
    mbar_init(&sh->sdp_full, 1);
    mbar_init(&sh->s_empty, kCompThreads);
    mbar_init(&sh->ds_full, kCompThreads);
    mbar_init(&sh->ds_empty, 1);
    mbar_init(&sh->acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(&sh->tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
}

template <bool DROP>
__global__ void __launch_bounds__(kAttnThreads, 2)
attn_bwd_dq_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO,
                   const __grid_constant__ CUtensorMap tmO, const AttnBwdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sQ = smem;                    // [128][64]
  uint8_t* sDO = sQ + kTile;             // [128][64]
  uint8_t* sDS = sDO + kTile;            // [128 rows][64 keys]; holds the O tile first (delta)
  uint8_t* sKV = sDS + kTile;            // 2 stages x (K_j 8 KB | V_j 8 KB)
  BwdSmem* sh = reinterpret_cast<BwdSmem*>(sKV + 2 * kTile);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qb = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  pdl_sync();  // key_len / key_mask / stats below are global reads
  const int q0 = qb * kAQ;
  const SeqGeo geo = seq_geo(a, b, h);
  if (q0 >= geo.Sq) return;  // packed mode: no query rows of this sequence in this tile
  int kmax = geo.Sk;
  if (a.key_len != nullptr && a.q_start == nullptr) {
    const int kl = a.key_len[b];
    if (kl > 0 && kl < kmax) kmax = kl;
  }
  if (a.causal) kmax = min(kmax, q0 + kAQ);
  const int nkb = (kmax + 63) / 64;
  bwd_init(sh, warp, lane);
  const uint32_t tmem_base = sh->tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&tmK);
      tma_prefetch_desc(&tmV);
      mbar_expect_tx(&sh->in_full, 3 * kTile);
      tma_load_4d(sQ, &tmQ, &sh->in_full, 0, geo.qr0 + q0, h, geo.bq);
      tma_load_4d(sDO, &tmDO, &sh->in_full, 0, geo.qr0 + q0, h, geo.bq);
      tma_load_4d(sDS, &tmO, &sh->in_full, 0, geo.qr0 + q0, h, geo.bq);
      for (int j = 0; j < nkb; ++j) {
        const int st = j & 1;
        mbar_wait(&sh->st_empty[st], ((j >> 1) & 1u) ^ 1u);
        mbar_expect_tx(&sh->st_full[st], 2 * kHalfTile);
        tma_load_4d(sKV + st * kTile, &tmK, &sh->st_full[st], 0, geo.kr0 + j * 64, h, geo.bq);
        tma_load_4d(sKV + st * kTile + kHalfTile, &tmV, &sh->st_full[st], 0, geo.kr0 + j * 64, h, geo.bq);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(kAQ, 64, 0, 0);
      constexpr uint32_t idesc_q = make_idesc_bf16(kAQ, kHD, 0, 1);
      const uint32_t q_addr = smem_u32(sQ), do_addr = smem_u32(sDO), ds_addr = smem_u32(sDS);
      mbar_wait(&sh->in_full, 0);
      auto issue_sdp = [&](int j) {  // S = Q K_j^T and dP = dO V_j^T (waits for the tiles and for S / dP to be drained)
        const int st = j & 1;
        mbar_wait(&sh->st_full[st], (j >> 1) & 1u);
        mbar_wait(&sh->s_empty, (j & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t k_addr = smem_u32(sKV + st * kTile), v_addr = k_addr + kHalfTile;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_ss(tmem_base, make_smem_desc_sw128(q_addr + k * 32, 16, 1024),
                       make_smem_desc_sw128(k_addr + k * 32, 16, 1024), idesc_s, k != 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_ss(tmem_base + 64, make_smem_desc_sw128(do_addr + k * 32, 16, 1024),
                       make_smem_desc_sw128(v_addr + k * 32, 16, 1024), idesc_s, k != 0 ? 1u : 0u);
        umma_commit(&sh->sdp_full);
      };
      issue_sdp(0);
      for (int j = 0; j < nkb; ++j) {
        const int st = j & 1;
        if (j + 1 < nkb) issue_sdp(j + 1);  // runs while the elementwise warps turn block j into dS
        const uint32_t k_addr = smem_u32(sKV + st * kTile);
        mbar_wait(&sh->ds_full, j & 1u);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)  // dQ += dS K_j : K = 64 keys, B = K_j [key][hd] (MN-major)
          umma_bf16_ss(tmem_base + 128, make_smem_desc_sw128(ds_addr + kk * 32, 16, 1024),
                       make_smem_desc_sw128(k_addr + kk * 2048, 64 * 128, 1024), idesc_q, (j | kk) != 0 ? 1u : 0u);
        umma_commit(&sh->st_empty[st]);
        umma_commit(&sh->ds_empty);
      }
      umma_commit(&sh->acc_full);
    }
  } else {
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;  // which 32 of the block's 64 key columns
    const int r = quad * 32 + lane;
    const int row = q0 + r;
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const uint8_t* mk = a.key_mask ? a.key_mask + static_cast<long long>(b) * a.Sk : nullptr;
    const float c2 = a.scale_log2;
    const uint32_t sw = static_cast<uint32_t>(r & 7);
    // classify the 64-key blocks: does any key need a mask decision?  (geometry, then the mask bytes in 16-key
    // segments over all 256 threads)
    {
      const int tid = static_cast<int>(threadIdx.x) - 64;
      for (int j = tid; j < nkb; j += kCompThreads) {
        const int kb0 = j * 64;
        sh->blk_flag[j] = ((kb0 + 64 > geo.Sk) || (a.causal && kb0 + 63 > q0)) ? 1 : 0;
      }
      if (mk != nullptr) {  // (packed mode and mask-free call sites skip the scan and its barrier)
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const int kend = min(nkb * 64, geo.Sk);
        for (int k = tid * 16; k < kend; k += kCompThreads * 16) {
          bool z = false;
#pragma unroll
          for (int c = 0; c < 16; ++c) z |= (k + c < kend) && (mk[k + c] == 0);
          if (z) sh->blk_flag[k / 64] = 1;
        }
      }
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    // ---- delta = rowsum(dO * O) (both column halves compute it; half 0 stores it)
    mbar_wait(&sh->in_full, 0);
    float delta = 0.f;
    {
      const uint32_t do_row = smem_u32(sDO) + r * 128, o_row = smem_u32(sDS) + r * 128;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const uint32_t off = (static_cast<uint32_t>(q) ^ sw) << 4;
        float x[8], y[8];
        const uint4 u = ld_shared_v4(do_row + off), w = ld_shared_v4(o_row + off);
        float2 t;
        t = unpack_bf16x2(u.x); x[0] = t.x; x[1] = t.y; t = unpack_bf16x2(u.y); x[2] = t.x; x[3] = t.y;
        t = unpack_bf16x2(u.z); x[4] = t.x; x[5] = t.y; t = unpack_bf16x2(u.w); x[6] = t.x; x[7] = t.y;
        t = unpack_bf16x2(w.x); y[0] = t.x; y[1] = t.y; t = unpack_bf16x2(w.y); y[2] = t.x; y[3] = t.y;
        t = unpack_bf16x2(w.z); y[4] = t.x; y[5] = t.y; t = unpack_bf16x2(w.w); y[6] = t.x; y[7] = t.y;
#pragma unroll
        for (int e = 0; e < 8; ++e) delta += x[e] * y[e];
      }
    }
    // every thread has read its O row before any thread overwrites the buffer with dS
    asm volatile("bar.sync 1, 256;" ::: "memory");
    // rows past the sequence end (TMA zero fill, or -- packed mode -- the next sequence's rows): reference maximum +big, so
    // every p = 2^(t - m) underflows to exactly 0 whatever those rows hold
    float m = 1.0e30f, coef = 0.f;
    if (row < geo.Sq) {
      const long long si = geo.sbase + row;
      const float2 st = reinterpret_cast<const float2*>(a.stats)[si];
      m = st.x;
      coef = st.y * a.scale;  // 1/rowsum * head_dim^-0.5
      if (half == 0) a.delta[si] = delta;
    }
    const uint32_t ds_row = smem_u32(sDS) + r * 128;
    const uint32_t dseed = DROP ? static_cast<uint32_t>(*a.rng) : 0u, dthr = drop_thresh(a.p_drop);
    const float dinv = DROP ? 1.f / (1.f - a.p_drop) : 1.f;
    for (int j = 0; j < nkb; ++j) {
      mbar_wait(&sh->sdp_full, j & 1u);
      tc_fence_after();
      const int kb0 = j * 64 + half * 32;
      const bool need = sh->blk_flag[j] != 0;
      uint32_t xs[32], xp[32];
      tmem_ld_32x32(tmem_base + lane_off + half * 32, xs);
      tmem_ld_32x32(tmem_base + lane_off + 64 + half * 32, xp);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&sh->s_empty);
      uint32_t dsp[16];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float ds[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int c = q * 8 + e;
          float t = fmaf(__uint_as_float(xs[c]), c2, -m);  // log2-domain score minus the row maximum
          if (need) {
            const int key = kb0 + c;
            if (key >= geo.Sk) t = -INFINITY;
            else if ((mk != nullptr && mk[key] == 0) || (a.causal && key > row)) t = -FLT_MAX - m;
          }
          float dp = __uint_as_float(xp[c]);
          if constexpr (DROP) dp = keep_attn(dseed, a.salt, geo.sbase + row, kb0 + c, dthr) ? dp * dinv : 0.f;
          ds[e] = ex2f(t) * (dp - delta) * coef;
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) dsp[q * 4 + e] = pack_bf16x2(ds[2 * e], ds[2 * e + 1]);
      }
      if (j > 0) mbar_wait(&sh->ds_empty, (j - 1) & 1u);  // dQ MMA of block j-1 has consumed the dS buffer
#pragma unroll
      for (int q = 0; q < 4; ++q)
        st_shared_v4(ds_row + ((static_cast<uint32_t>(half * 4 + q) ^ sw) << 4), dsp[q * 4], dsp[q * 4 + 1], dsp[q * 4 + 2],
                     dsp[q * 4 + 3]);
      fence_proxy_async_smem();
      mbar_arrive(&sh->ds_full);
    }
    mbar_wait(&sh->acc_full, 0);
    tc_fence_after();
    uint32_t o0[32];
    tmem_ld_32x32(tmem_base + lane_off + 128 + half * 32, o0);
    tmem_ld_wait();
    if (row < geo.Sq)
      store_row32(a.dq + static_cast<long long>(geo.bq) * a.dq_sb + h * a.dq_sh + static_cast<long long>(geo.qr0 + row) * a.lddq + half * 32, o0);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

template <bool DROP>
__global__ void __launch_bounds__(kAttnThreads, 2)
attn_bwd_dkv_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO,
                    const AttnBwdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sK = smem;                    // [128 keys][64]
  uint8_t* sV = sK + kTile;
  uint8_t* sPT = sV + kTile;             // [128 keys][64 queries]
  uint8_t* sDST = sPT + kTile;
  uint8_t* sQD = sDST + kTile;           // 2 stages x (Q_i 8 KB | dO_i 8 KB)
  BwdSmem* sh = reinterpret_cast<BwdSmem*>(sQD + 2 * kTile);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kb = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  pdl_sync();  // key_len / key_mask / stats below are global reads
  const int k0 = kb * kAK;
  const SeqGeo geo = seq_geo(a, b, h);
  if (k0 >= geo.Sk) return;  // packed mode: no key rows of this sequence in this tile (legacy mode: never true)
  const int kl = (a.key_len != nullptr && a.q_start == nullptr) ? a.key_len[b] : 0;
  if (kl > 0 && k0 >= kl) {
    // every key of this block is masked for every query: dK = dV = 0
    if (warp >= 2 && warp < 6) {
      const int key = k0 + (warp - 2) * 32 + lane;
      if (key < a.Sk) {
        uint4* pk = reinterpret_cast<uint4*>(a.dk + static_cast<long long>(b) * a.dk_sb + h * a.dk_sh + static_cast<long long>(key) * a.lddk);
        uint4* pv = reinterpret_cast<uint4*>(a.dv + static_cast<long long>(b) * a.dv_sb + h * a.dv_sh + static_cast<long long>(key) * a.lddv);
#pragma unroll
        for (int q = 0; q < 8; ++q) { pk[q] = make_uint4(0, 0, 0, 0); pv[q] = make_uint4(0, 0, 0, 0); }
      }
    }
    return;
  }
  const int i0 = a.causal ? k0 / 64 : 0;           // query blocks entirely above the diagonal contribute nothing
  const int nqb = (geo.Sq + 63) / 64;
  bwd_init(sh, warp, lane);
  const uint32_t tmem_base = sh->tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmDO);
      mbar_expect_tx(&sh->in_full, 2 * kTile);
      tma_load_4d(sK, &tmK, &sh->in_full, 0, geo.kr0 + k0, h, geo.bq);
      tma_load_4d(sV, &tmV, &sh->in_full, 0, geo.kr0 + k0, h, geo.bq);
      for (int i = i0, n = 0; i < nqb; ++i, ++n) {
        const int st = n & 1;
        mbar_wait(&sh->st_empty[st], ((n >> 1) & 1u) ^ 1u);
        mbar_expect_tx(&sh->st_full[st], 2 * kHalfTile);
        tma_load_4d(sQD + st * kTile, &tmQ, &sh->st_full[st], 0, geo.qr0 + i * 64, h, geo.bq);
        tma_load_4d(sQD + st * kTile + kHalfTile, &tmDO, &sh->st_full[st], 0, geo.qr0 + i * 64, h, geo.bq);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(kAK, 64, 0, 0);
      constexpr uint32_t idesc_g = make_idesc_bf16(kAK, kHD, 0, 1);
      const uint32_t k_addr = smem_u32(sK), v_addr = smem_u32(sV), pt_addr = smem_u32(sPT), dst_addr = smem_u32(sDST);
      mbar_wait(&sh->in_full, 0);
      auto issue_sdp = [&](int n) {  // S^T = K Q_i^T and dP^T = V dO_i^T for step n
        const int st = n & 1;
        mbar_wait(&sh->st_full[st], (n >> 1) & 1u);
        mbar_wait(&sh->s_empty, (n & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t q_addr = smem_u32(sQD + st * kTile), do_addr = q_addr + kHalfTile;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_ss(tmem_base, make_smem_desc_sw128(k_addr + k * 32, 16, 1024),
                       make_smem_desc_sw128(q_addr + k * 32, 16, 1024), idesc_s, k != 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_ss(tmem_base + 64, make_smem_desc_sw128(v_addr + k * 32, 16, 1024),
                       make_smem_desc_sw128(do_addr + k * 32, 16, 1024), idesc_s, k != 0 ? 1u : 0u);
        umma_commit(&sh->sdp_full);
      };
      const int nsteps = nqb - i0;
      if (nsteps > 0) issue_sdp(0);
      for (int n = 0; n < nsteps; ++n) {
        const int st = n & 1;
        if (n + 1 < nsteps) issue_sdp(n + 1);  // runs while the elementwise warps turn step n into P^T / dS^T
        const uint32_t q_addr = smem_u32(sQD + st * kTile), do_addr = q_addr + kHalfTile;
        mbar_wait(&sh->ds_full, n & 1u);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)  // dV += P^T dO_i
          umma_bf16_ss(tmem_base + 128, make_smem_desc_sw128(pt_addr + kk * 32, 16, 1024),
                       make_smem_desc_sw128(do_addr + kk * 2048, 64 * 128, 1024), idesc_g, (n | kk) != 0 ? 1u : 0u);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)  // dK += dS^T Q_i
          umma_bf16_ss(tmem_base + 192, make_smem_desc_sw128(dst_addr + kk * 32, 16, 1024),
                       make_smem_desc_sw128(q_addr + kk * 2048, 64 * 128, 1024), idesc_g, (n | kk) != 0 ? 1u : 0u);
        umma_commit(&sh->st_empty[st]);
        umma_commit(&sh->ds_empty);
      }
      umma_commit(&sh->acc_full);
    }
  } else {
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;  // which 32 of the block's 64 query columns
    const int r = quad * 32 + lane;
    const int key = k0 + r;
    const int tid = threadIdx.x - 64;  // 0..255
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const bool key_oob = key >= geo.Sk;
    const bool key_masked = !key_oob && a.key_mask != nullptr && a.key_mask[static_cast<long long>(b) * a.Sk + key] == 0;
    const float c2 = a.scale_log2;
    const bool warp_plain = !a.causal && __all_sync(0xffffffffu, !(key_oob || key_masked));
    const uint32_t sw = static_cast<uint32_t>(r & 7);
    const uint32_t pt_row = smem_u32(sPT) + r * 128, dst_row = smem_u32(sDST) + r * 128;
    const long long sbase = geo.sbase;
    const uint32_t dseed = DROP ? static_cast<uint32_t>(*a.rng) : 0u, dthr = drop_thresh(a.p_drop);
    const float dinv = DROP ? 1.f / (1.f - a.p_drop) : 1.f;
    // row statistics of a 64-query block (threads 0..63): {reference max, 1/row sum, delta, scale/row sum}
    auto load_stat = [&](int i) -> float4 {
      const int qrow = i * 64 + tid;
      float4 st = make_float4(1.0e30f, 0.f, 0.f, 0.f);  // query rows beyond the sequence: max = +big, 1/sum = 0 -> p = 0 exactly
      if (qrow < geo.Sq) {
        const float2 s2 = reinterpret_cast<const float2*>(a.stats)[sbase + qrow];
        st = make_float4(s2.x, s2.y, a.delta[sbase + qrow], s2.y * a.scale);
      }
      return st;
    };
    float4 st_next = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tid < 64 && i0 < nqb) st_next = load_stat(i0);
    for (int i = i0, n = 0; i < nqb; ++i, ++n) {
      // stage the row statistics of the 64 queries of this block (double buffered: buffer n&1 was last read in
      // step n-2, which every thread finished before passing the named barrier of step n-1).  The global loads
      // were issued one block ahead, so the barrier below no longer waits for a DRAM / L2 round trip.
      float4* stq = sh->stat[n & 1];
      if (tid < 64) {
        stq[tid] = st_next;
        if (i + 1 < nqb) st_next = load_stat(i + 1);
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      mbar_wait(&sh->sdp_full, n & 1u);
      tc_fence_after();
      const int qb0 = i * 64 + half * 32;
      uint32_t xs[32], xp[32];
      tmem_ld_32x32(tmem_base + lane_off + half * 32, xs);
      tmem_ld_32x32(tmem_base + lane_off + 64 + half * 32, xp);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&sh->s_empty);
      uint32_t ppk[16], dsk[16];
      if (warp_plain && !DROP) {
        // no key of this warp is masked and there is no causal mask: no per-element decisions (the common case)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float pp[8], ds[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int c = q * 8 + e;
            const float4 st = stq[half * 32 + c];
            const float p = ex2f(fmaf(__uint_as_float(xs[c]), c2, -st.x)) * st.y;
            pp[e] = p;
            ds[e] = p * (__uint_as_float(xp[c]) - st.z) * a.scale;
          }
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            ppk[q * 4 + e] = pack_bf16x2(pp[2 * e], pp[2 * e + 1]);
            dsk[q * 4 + e] = pack_bf16x2(ds[2 * e], ds[2 * e + 1]);
          }
        }
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float pp[8], ds[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int c = q * 8 + e;
            const float4 st = stq[half * 32 + c];
            float t = __uint_as_float(xs[c]) * c2;
            if (key_oob) t = -INFINITY;
            else if (key_masked || (a.causal && key > qb0 + c)) t = -FLT_MAX;
            const float p = ex2f(t - st.x) * st.y;
            float dp = __uint_as_float(xp[c]);
            if constexpr (DROP) {
              const bool kp = keep_attn(dseed, a.salt, sbase + qb0 + c, key, dthr);
              pp[e] = kp ? p * dinv : 0.f;   // P^T as the forward pass applied it to V
              dp = kp ? dp * dinv : 0.f;
            } else {
              pp[e] = p;
            }
            ds[e] = p * (dp - st.z) * a.scale;
          }
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            ppk[q * 4 + e] = pack_bf16x2(pp[2 * e], pp[2 * e + 1]);
            dsk[q * 4 + e] = pack_bf16x2(ds[2 * e], ds[2 * e + 1]);
          }
        }
      }
      if (n > 0) mbar_wait(&sh->ds_empty, (n - 1) & 1u);  // the dV / dK MMAs of step n-1 have consumed both buffers
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t off = (static_cast<uint32_t>(half * 4 + q) ^ sw) << 4;
        st_shared_v4(pt_row + off, ppk[q * 4], ppk[q * 4 + 1], ppk[q * 4 + 2], ppk[q * 4 + 3]);
        st_shared_v4(dst_row + off, dsk[q * 4], dsk[q * 4 + 1], dsk[q * 4 + 2], dsk[q * 4 + 3]);
      }
      fence_proxy_async_smem();
      mbar_arrive(&sh->ds_full);
    }
    mbar_wait(&sh->acc_full, 0);
    tc_fence_after();
    uint32_t o0[32], o1[32];
    tmem_ld_32x32(tmem_base + lane_off + 128 + half * 32, o0);
    tmem_ld_32x32(tmem_base + lane_off + 192 + half * 32, o1);
    tmem_ld_wait();
    if (!key_oob) {
      store_row32(a.dv + static_cast<long long>(geo.bq) * a.dv_sb + h * a.dv_sh + static_cast<long long>(geo.kr0 + key) * a.lddv + half * 32, o0);
      store_row32(a.dk + static_cast<long long>(geo.bq) * a.dk_sb + h * a.dk_sh + static_cast<long long>(geo.kr0 + key) * a.lddk + half * 32, o1);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

constexpr int kBwdDqSmemBytes = 5 * kTile + static_cast<int>(sizeof(BwdSmem)) + 1024;
constexpr int kBwdDkvSmemBytes = 6 * kTile + static_cast<int>(sizeof(BwdSmem)) + 1024;

constexpr int kAttnSmemBytes = kTile * (1 + kRing + 2) + static_cast<int>(sizeof(AttnSmem)) + 1024;

}  // namespace vb

using namespace vb;

extern "C" int vacnic_attn_fwd(const vacnic_attn_desc* d, void* stream) {
  if (d == nullptr) return fail(VACNIC_EINVAL, "attn_fwd: null descriptor");
  VB_REQUIRE(d->q && d->k && d->v && d->out, "attn_fwd: null pointer");
  VB_REQUIRE(d->head_dim == 64, "attn_fwd: head_dim must be 64 (BART-base / BART-large)");
  VB_REQUIRE(d->B > 0 && d->H > 0 && d->Sq > 0 && d->Sk > 0 && d->Sk <= kMaxKB * kAK, "attn_fwd: bad shape (Sk <= %d)", kMaxKB * kAK);
  VB_REQUIRE(d->ldo % 8 == 0 && d->o_sb % 8 == 0 && (reinterpret_cast<uintptr_t>(d->out) & 15) == 0, "attn_fwd: out must be 16-byte aligned");
  const bool packed = d->q_start != nullptr;
  if (packed) {
    VB_REQUIRE(d->q_len && d->k_start && d->k_len && d->total_q > 0 && d->total_k > 0, "attn_fwd: packed mode needs q_len, k_start, k_len, total_q, total_k");
    VB_REQUIRE(d->key_mask == nullptr && d->key_len == nullptr, "attn_fwd: packed mode takes no key_mask / key_len");
  }
  // packed mode: one [total rows] x [H*64] matrix per operand (batch dimension of the tensor map = 1)
  const int rows_q = packed ? d->total_q : d->Sq, rows_k = packed ? d->total_k : d->Sk, nb = packed ? 1 : d->B;
  CUtensorMap tq, tk, tv;
  int rc = make_operand_map(&tq, d->q, false, rows_q, 64, d->ldq, d->H, d->q_sh, nb, packed ? 0 : d->q_sb, kAQ);
  if (rc != VACNIC_OK) return rc;
  rc = make_operand_map(&tk, d->k, false, rows_k, 64, d->ldk, d->H, d->k_sh, nb, packed ? 0 : d->k_sb, kAK);
  if (rc != VACNIC_OK) return rc;
  rc = make_operand_map(&tv, d->v, false, rows_k, 64, d->ldv, d->H, d->v_sh, nb, packed ? 0 : d->v_sb, kAK);
  if (rc != VACNIC_OK) return rc;
  AttnArgs a;
  a.B = d->B; a.H = d->H; a.Sq = d->Sq; a.Sk = d->Sk; a.causal = d->causal;
  a.scale_log2 = 1.4426950408889634f / sqrtf(static_cast<float>(d->head_dim));
  a.key_mask = d->key_mask; a.key_len = d->key_len;
  a.out = static_cast<__nv_bfloat16*>(d->out); a.ldo = d->ldo; a.o_sb = d->o_sb;
  a.stats = d->stats;
  a.q_start = d->q_start; a.q_len = d->q_len; a.k_start = d->k_start; a.k_len = d->k_len; a.total_q = d->total_q;
  VB_REQUIRE(d->p_drop >= 0.f && d->p_drop < 1.f && (d->p_drop == 0.f || d->rng_state), "attn_fwd: bad dropout args");
  a.p_drop = d->p_drop; a.rng = reinterpret_cast<const unsigned long long*>(d->rng_state); a.salt = d->salt;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmemBytes);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attn_fwd_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmemBytes);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attn_fwd_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmemBytes);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attn_fwd_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmemBytes);
    if (e != cudaSuccess) return fail(VACNIC_ECUDA, "attn_fwd: cudaFuncSetAttribute(smem=%d): %s", kAttnSmemBytes, cudaGetErrorString(e));
    configured = true;
  }
  const dim3 grid((d->Sq + kAQ - 1) / kAQ, d->H, d->B);
  const bool general = d->key_mask != nullptr || d->causal;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (a.p_drop > 0.f) {
    if (general) launch_pdl(attn_fwd_kernel<true, true>, grid, dim3(kAttnThreads), kAttnSmemBytes, st, tq, tk, tv, a);
    else launch_pdl(attn_fwd_kernel<false, true>, grid, dim3(kAttnThreads), kAttnSmemBytes, st, tq, tk, tv, a);
  } else {
    if (general) launch_pdl(attn_fwd_kernel<true, false>, grid, dim3(kAttnThreads), kAttnSmemBytes, st, tq, tk, tv, a);
    else launch_pdl(attn_fwd_kernel<false, false>, grid, dim3(kAttnThreads), kAttnSmemBytes, st, tq, tk, tv, a);
  }
  count_launch();
  return check_last("attn_fwd launch");
}

extern "C" int vacnic_attn_bwd(const vacnic_attn_desc* d, void* stream) {
  if (d == nullptr) return fail(VACNIC_EINVAL, "attn_bwd: null descriptor");
  VB_REQUIRE(d->q && d->k && d->v && d->out && d->dout && d->dq && d->dk && d->dv && d->stats && d->delta, "attn_bwd: null pointer");
  VB_REQUIRE(d->head_dim == 64, "attn_bwd: head_dim must be 64");
  VB_REQUIRE(d->B > 0 && d->H > 0 && d->Sq > 0 && d->Sk > 0, "attn_bwd: bad shape");
  for (const void* p : {static_cast<const void*>(d->dq), static_cast<const void*>(d->dk), static_cast<const void*>(d->dv)})
    VB_REQUIRE((reinterpret_cast<uintptr_t>(p) & 15) == 0, "attn_bwd: gradient buffers must be 16-byte aligned");
  VB_REQUIRE(d->lddq % 8 == 0 && d->lddk % 8 == 0 && d->lddv % 8 == 0 && d->dq_sh % 8 == 0 && d->dk_sh % 8 == 0 && d->dv_sh % 8 == 0 &&
                 d->dq_sb % 8 == 0 && d->dk_sb % 8 == 0 && d->dv_sb % 8 == 0, "attn_bwd: gradient strides must be multiples of 8");
  const bool packed = d->q_start != nullptr;
  if (packed) {
    VB_REQUIRE(d->q_len && d->k_start && d->k_len && d->total_q > 0 && d->total_k > 0, "attn_bwd: packed mode needs q_len, k_start, k_len, total_q, total_k");
    VB_REQUIRE(d->key_mask == nullptr && d->key_len == nullptr, "attn_bwd: packed mode takes no key_mask / key_len");
  }
  const int rows_q = packed ? d->total_q : d->Sq, rows_k = packed ? d->total_k : d->Sk, nb = packed ? 1 : d->B;
  CUtensorMap q128, q64, k128, k64, v128, v64, do128, do64, o128;
  int rc;
#define VB_MAP(tm, base, rows, ld, sh, sb, box)                                                        \
  rc = make_operand_map(&tm, base, false, rows, 64, ld, d->H, sh, nb, packed ? 0 : sb, box);           \
  if (rc != VACNIC_OK) return rc;
  VB_MAP(q128, d->q, rows_q, d->ldq, d->q_sh, d->q_sb, 128)
  VB_MAP(q64, d->q, rows_q, d->ldq, d->q_sh, d->q_sb, 64)
  VB_MAP(k128, d->k, rows_k, d->ldk, d->k_sh, d->k_sb, 128)
  VB_MAP(k64, d->k, rows_k, d->ldk, d->k_sh, d->k_sb, 64)
  VB_MAP(v128, d->v, rows_k, d->ldv, d->v_sh, d->v_sb, 128)
  VB_MAP(v64, d->v, rows_k, d->ldv, d->v_sh, d->v_sb, 64)
  VB_MAP(do128, d->dout, rows_q, d->lddo, 64, d->do_sb, 128)
  VB_MAP(do64, d->dout, rows_q, d->lddo, 64, d->do_sb, 64)
  VB_MAP(o128, d->out, rows_q, d->ldo, 64, d->o_sb, 128)
#undef VB_MAP
  AttnBwdArgs a;
  a.B = d->B; a.H = d->H; a.Sq = d->Sq; a.Sk = d->Sk; a.causal = d->causal;
  a.scale = 1.0f / sqrtf(static_cast<float>(d->head_dim));
  a.scale_log2 = 1.4426950408889634f * a.scale;
  a.key_mask = d->key_mask; a.key_len = d->key_len; a.stats = d->stats; a.delta = d->delta;
  a.dq = static_cast<__nv_bfloat16*>(d->dq); a.lddq = d->lddq; a.dq_sh = d->dq_sh; a.dq_sb = d->dq_sb;
  a.dk = static_cast<__nv_bfloat16*>(d->dk); a.lddk = d->lddk; a.dk_sh = d->dk_sh; a.dk_sb = d->dk_sb;
  a.dv = static_cast<__nv_bfloat16*>(d->dv); a.lddv = d->lddv; a.dv_sh = d->dv_sh; a.dv_sb = d->dv_sb;
  a.q_start = d->q_start; a.q_len = d->q_len; a.k_start = d->k_start; a.k_len = d->k_len; a.total_q = d->total_q;
  VB_REQUIRE(d->p_drop >= 0.f && d->p_drop < 1.f && (d->p_drop == 0.f || d->rng_state), "attn_bwd: bad dropout args");
  a.p_drop = d->p_drop; a.rng = reinterpret_cast<const unsigned long long*>(d->rng_state); a.salt = d->salt;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_dq_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdDqSmemBytes);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attn_bwd_dkv_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdDkvSmemBytes);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attn_bwd_dq_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdDqSmemBytes);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attn_bwd_dkv_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdDkvSmemBytes);
    if (e != cudaSuccess) return fail(VACNIC_ECUDA, "attn_bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    configured = true;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const dim3 gq((d->Sq + kAQ - 1) / kAQ, d->H, d->B), gk((d->Sk + kAK - 1) / kAK, d->H, d->B);
  if (a.p_drop > 0.f)
    launch_pdl(attn_bwd_dq_kernel<true>, gq, dim3(kAttnThreads), kBwdDqSmemBytes, s, q128, k64, v64, do128, o128, a);
  else
    launch_pdl(attn_bwd_dq_kernel<false>, gq, dim3(kAttnThreads), kBwdDqSmemBytes, s, q128, k64, v64, do128, o128, a);
  count_launch();
  rc = check_last("attn_bwd dq launch");
  if (rc != VACNIC_OK) return rc;
  if (a.p_drop > 0.f)
    launch_pdl(attn_bwd_dkv_kernel<true>, gk, dim3(kAttnThreads), kBwdDkvSmemBytes, s, q64, k128, v128, do64, a);
  else
    launch_pdl(attn_bwd_dkv_kernel<false>, gk, dim3(kAttnThreads), kBwdDkvSmemBytes, s, q64, k128, v128, do64, a);
  count_launch();
  return check_last("attn_bwd dkv launch");
}
