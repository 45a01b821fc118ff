// Data-parallel optimizer step over NVLink peer memory (TRAIN:86-87 DDP gradient averaging + TRAIN:91-107, 371-373
// AdamW, as ONE kernel per address range):
//
//   reduce-scatter : this rank owns a contiguous shard of every gradient bucket; it LOADS that shard of the fp32
//                    gradient from every rank's flat gradient buffer (peer loads through NVSwitch; the buffers are
//                    symmetric-memory allocations mapped into every process) and sums them in rank order in fp32
//   AdamW          : updates its shard of the fp32 master weights and Adam moments (local HBM only)
//   all-gather     : packs the new weights to bf16 and STORES them into every rank's bf16 compute shadow
//
// so no gradient is ever staged, cast or copied, the wire carries (N-1)/N * (4 + 2) bytes per parameter instead of
// 2 * (N-1)/N * 4 (NCCL fp32 all-reduce), the optimizer's HBM traffic drops by N, and every element is produced by exactly
// one rank (replicas are bit-identical by construction; the sum order is fixed, so runs are reproducible).
// Ordering between ranks (bucket ready / step finished) is the caller's job: a symmetric-memory barrier on the
// communication stream before the first launch of a bucket and after the last launch of a step (trainer.py).
//
// A second variant uses the NVSwitch multicast object of the same buffers (`multimem.ld_reduce` sums the N copies
// inside the switch, `multimem.st` writes all N shadows with one store): 1 load + 1 store per 16 bytes instead of N.
#include "common.h"
#include "ptx.cuh"

namespace vb {

constexpr int kMaxRanks = 8;

struct DpShardArgs {
  const float* grad[kMaxRanks];      // flat fp32 gradient buffer of every rank (index = rank), peer-mapped
  __nv_bfloat16* shadow[kMaxRanks];  // flat bf16 compute shadow of every rank
  const float* grad_mc;              // multicast address of the gradient buffers (variant 2), or null
  __nv_bfloat16* shadow_mc;          // multicast address of the shadows (variant 2), or null
  float* master;                     // local: fp32 weights, Adam moments
  float* m;
  float* v;
  const float* hyper;                // {lr, beta1, beta2, eps, wd, 1-beta1^t, 1-beta2^t, grad_scale}
  long long begin;                   // first element of this rank's shard (multiple of 8)
  long long count;                   // elements in the shard (multiple of 8)
};

__device__ __forceinline__ float4 ld_f4(const float* p) {
  float4 r;
  asm volatile("ld.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ float4 ld_reduce_mc_f4(const float* p) {
  float4 r;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ void st_mc_u4(void* p, const uint4& u) {
  asm volatile("multimem.st.relaxed.sys.global.v4.bf16x2 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "r"(u.x), "r"(u.y), "r"(u.z), "r"(u.w) : "memory");
}

template <int W, bool MC>
__global__ void __launch_bounds__(256)
dp_adamw_shard_kernel(const DpShardArgs a) {
  const AdamwCoef c = adamw_coef(a.hyper);
  const long long groups = a.count >> 3;
  for (long long gi = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; gi < groups;
       gi += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long i = a.begin + (gi << 3);
    float g[8];
    if (MC) {
      const float4 s0 = ld_reduce_mc_f4(a.grad_mc + i), s1 = ld_reduce_mc_f4(a.grad_mc + i + 4);
      g[0] = s0.x; g[1] = s0.y; g[2] = s0.z; g[3] = s0.w; g[4] = s1.x; g[5] = s1.y; g[6] = s1.z; g[7] = s1.w;
    } else {
      float4 lo[W], hi[W];
#pragma unroll
      for (int r = 0; r < W; ++r) {  // all peer loads in flight before the first add
        lo[r] = ld_f4(a.grad[r] + i);
        hi[r] = ld_f4(a.grad[r] + i + 4);
      }
      g[0] = lo[0].x; g[1] = lo[0].y; g[2] = lo[0].z; g[3] = lo[0].w;
      g[4] = hi[0].x; g[5] = hi[0].y; g[6] = hi[0].z; g[7] = hi[0].w;
#pragma unroll
      for (int r = 1; r < W; ++r) {  // rank order: the sum is the same whoever owns the shard
        g[0] += lo[r].x; g[1] += lo[r].y; g[2] += lo[r].z; g[3] += lo[r].w;
        g[4] += hi[r].x; g[5] += hi[r].y; g[6] += hi[r].z; g[7] += hi[r].w;
      }
    }
    float4 p0 = *reinterpret_cast<float4*>(a.master + i), p1 = *reinterpret_cast<float4*>(a.master + i + 4);
    float4 m0 = *reinterpret_cast<float4*>(a.m + i), m1 = *reinterpret_cast<float4*>(a.m + i + 4);
    float4 v0 = *reinterpret_cast<float4*>(a.v + i), v1 = *reinterpret_cast<float4*>(a.v + i + 4);
    float p[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
    float mm[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
    float vv[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
    for (int k = 0; k < 8; ++k) adamw_elem(c, g[k], p[k], mm[k], vv[k]);  // shared with adamw_kernel: bit-identical
    *reinterpret_cast<float4*>(a.master + i) = make_float4(p[0], p[1], p[2], p[3]);
    *reinterpret_cast<float4*>(a.master + i + 4) = make_float4(p[4], p[5], p[6], p[7]);
    *reinterpret_cast<float4*>(a.m + i) = make_float4(mm[0], mm[1], mm[2], mm[3]);
    *reinterpret_cast<float4*>(a.m + i + 4) = make_float4(mm[4], mm[5], mm[6], mm[7]);
    *reinterpret_cast<float4*>(a.v + i) = make_float4(vv[0], vv[1], vv[2], vv[3]);
    *reinterpret_cast<float4*>(a.v + i + 4) = make_float4(vv[4], vv[5], vv[6], vv[7]);
    uint4 u;
    u.x = pack_bf16x2(p[0], p[1]); u.y = pack_bf16x2(p[2], p[3]);
    u.z = pack_bf16x2(p[4], p[5]); u.w = pack_bf16x2(p[6], p[7]);
    if (MC) {
      st_mc_u4(a.shadow_mc + i, u);
    } else {
#pragma unroll
      for (int r = 0; r < W; ++r) *reinterpret_cast<uint4*>(a.shadow[r] + i) = u;
    }
  }
}

}  // namespace vb

using namespace vb;

extern "C" int vacnic_dp_adamw_shard(const uint64_t* grad_ptrs, const uint64_t* shadow_ptrs, uint64_t grad_mc,
                                     uint64_t shadow_mc, int32_t world, int32_t rank, float* master, float* m, float* v,
                                     const float* hyper, int64_t begin, int64_t count, int32_t max_blocks, void* stream) {
  VB_REQUIRE(grad_ptrs && shadow_ptrs && master && m && v && hyper, "dp_adamw_shard: null pointer");
  VB_REQUIRE(world == 1 || world == 2 || world == 4 || world == 8, "dp_adamw_shard: world must be 1, 2, 4 or 8 (got %d)", world);
  VB_REQUIRE(rank >= 0 && rank < world, "dp_adamw_shard: bad rank");
  VB_REQUIRE(begin >= 0 && count >= 0 && (begin & 7) == 0 && (count & 7) == 0,
             "dp_adamw_shard: shard [begin, begin+count) must be a multiple of 8 elements");
  VB_REQUIRE((grad_mc == 0) == (shadow_mc == 0), "dp_adamw_shard: both multicast addresses or none");
  if (count == 0) return VACNIC_OK;
  DpShardArgs a = {};
  for (int r = 0; r < world; ++r) {
    VB_REQUIRE(grad_ptrs[r] && shadow_ptrs[r] && (grad_ptrs[r] & 15) == 0 && (shadow_ptrs[r] & 15) == 0,
               "dp_adamw_shard: peer buffers must be non-null and 16-byte aligned");
    a.grad[r] = reinterpret_cast<const float*>(grad_ptrs[r]);
    a.shadow[r] = reinterpret_cast<__nv_bfloat16*>(shadow_ptrs[r]);
  }
  VB_REQUIRE(((reinterpret_cast<uintptr_t>(master) | reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v)) & 15) == 0,
             "dp_adamw_shard: local buffers must be 16-byte aligned");
  a.grad_mc = reinterpret_cast<const float*>(grad_mc);
  a.shadow_mc = reinterpret_cast<__nv_bfloat16*>(shadow_mc);
  a.master = master; a.m = m; a.v = v; a.hyper = hyper; a.begin = begin; a.count = count;
  const long long groups = count >> 3;
  long long blocks = (groups + 255) / 256;
  const long long cap = max_blocks > 0 ? max_blocks : 4LL * sm_count();
  if (blocks > cap) blocks = cap;
  const dim3 grid(static_cast<unsigned>(blocks)), block(256);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (grad_mc != 0) {
    dp_adamw_shard_kernel<1, true><<<grid, block, 0, s>>>(a);
  } else {
    switch (world) {
      case 1: dp_adamw_shard_kernel<1, false><<<grid, block, 0, s>>>(a); break;
      case 2: dp_adamw_shard_kernel<2, false><<<grid, block, 0, s>>>(a); break;
      case 4: dp_adamw_shard_kernel<4, false><<<grid, block, 0, s>>>(a); break;
      default: dp_adamw_shard_kernel<8, false><<<grid, block, 0, s>>>(a); break;
    }
  }
  count_launch();
  return check_last("dp_adamw_shard");
}
