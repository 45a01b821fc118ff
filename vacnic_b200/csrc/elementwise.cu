// Small HBM-bound helpers: bias gradients (column sums), fp32 -> bf16 shadow casts, the fused
// AdamW update (TRAIN:91-107, 371-373: AdamW(lr, betas=(0.9,0.999), eps=1e-8, weight_decay) + linear
// warm-up schedule folded into `lr`), and gradient-buffer utilities.
#include "common.h"
#include "ptx.cuh"

namespace vb {

// out[c] += sum_r x[r, c]; x bf16 [rows, ld]; grid = (col tiles of 256, row slabs)
__global__ void __launch_bounds__(256)
colsum_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ out, long long rows, int n, long long ld,
              int rows_per_block) {
  pdl_sync();
  const int c0 = blockIdx.x * 256 + (threadIdx.x & 31) * 8;
  const int wr = threadIdx.x >> 5;  // 8 warps walk rows
  const long long r0 = static_cast<long long>(blockIdx.y) * rows_per_block;
  const long long r1 = min(rows, r0 + rows_per_block);
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const bool vec = (c0 + 8 <= n) && (ld % 8 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
  auto add8 = [&](const uint4& u) {
    float2 f;
    f = unpack_bf16x2(u.x); acc[0] += f.x; acc[1] += f.y;
    f = unpack_bf16x2(u.y); acc[2] += f.x; acc[3] += f.y;
    f = unpack_bf16x2(u.z); acc[4] += f.x; acc[5] += f.y;
    f = unpack_bf16x2(u.w); acc[6] += f.x; acc[7] += f.y;
  };
  long long r = r0 + wr;
  if (vec) {
    // four independent 16-byte loads in flight per thread (the one-load loop left HBM latency exposed)
    for (; r + 24 < r1; r += 32) {
      const __nv_bfloat16* p = x + r * ld + c0;
      const uint4 u0 = __ldg(reinterpret_cast<const uint4*>(p));
      const uint4 u1 = __ldg(reinterpret_cast<const uint4*>(p + 8 * ld));
      const uint4 u2 = __ldg(reinterpret_cast<const uint4*>(p + 16 * ld));
      const uint4 u3 = __ldg(reinterpret_cast<const uint4*>(p + 24 * ld));
      add8(u0); add8(u1); add8(u2); add8(u3);
    }
    for (; r < r1; r += 8) add8(__ldg(reinterpret_cast<const uint4*>(x + r * ld + c0)));
  } else {
    for (; r < r1; r += 8) {
      const __nv_bfloat16* p = x + r * ld + c0;
#pragma unroll
      for (int e = 0; e < 8; ++e)
        if (c0 + e < n) acc[e] += __bfloat162float(p[e]);
    }
  }
  __shared__ float red[8][256];
#pragma unroll
  for (int e = 0; e < 8; ++e) red[wr][(threadIdx.x & 31) * 8 + e] = acc[e];
  __syncthreads();
  const int c = threadIdx.x;
  float s = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) s += red[w][c];
  if (blockIdx.x * 256 + c < n) atomicAdd(out + blockIdx.x * 256 + c, s);
}

__global__ void __launch_bounds__(256)
cast_f32_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long n) {
  const long long i = (static_cast<long long>(blockIdx.x) * 256 + threadIdx.x) * 8;
  if (i + 8 <= n) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(src + i));
    const float4 b = __ldg(reinterpret_cast<const float4*>(src + i) + 1);
    uint4 u;
    u.x = pack_bf16x2(a.x, a.y); u.y = pack_bf16x2(a.z, a.w);
    u.z = pack_bf16x2(b.x, b.y); u.w = pack_bf16x2(b.z, b.w);
    *reinterpret_cast<uint4*>(dst + i) = u;
  } else {
    for (long long k = i; k < n; ++k) dst[k] = __float2bfloat16_rn(src[k]);
  }
}

// hyper = {lr, beta1, beta2, eps, weight_decay, bias_correction1, bias_correction2, grad_scale}
__global__ void __launch_bounds__(256)
adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
             __nv_bfloat16* __restrict__ p16, long long n, const float* __restrict__ hyper) {
  const AdamwCoef c = adamw_coef(hyper);
  const long long i0 = (static_cast<long long>(blockIdx.x) * 256 + threadIdx.x) * 4;
  if (i0 >= n) return;
  if (i0 + 4 <= n) {
    float4 pp = *reinterpret_cast<float4*>(p + i0);
    const float4 gg = __ldg(reinterpret_cast<const float4*>(g + i0));
    float4 mm = *reinterpret_cast<float4*>(m + i0);
    float4 vv = *reinterpret_cast<float4*>(v + i0);
    float* pa = &pp.x; const float* ga = &gg.x; float* ma = &mm.x; float* va = &vv.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) adamw_elem(c, ga[k], pa[k], ma[k], va[k]);
    *reinterpret_cast<float4*>(p + i0) = pp;
    *reinterpret_cast<float4*>(m + i0) = mm;
    *reinterpret_cast<float4*>(v + i0) = vv;
    if (p16) {
      uint2 u;
      u.x = pack_bf16x2(pp.x, pp.y); u.y = pack_bf16x2(pp.z, pp.w);
      *reinterpret_cast<uint2*>(p16 + i0) = u;
    }
  } else {
    for (long long i = i0; i < n; ++i) {
      float pv = p[i], mv = m[i], vv = v[i];
      adamw_elem(c, g[i], pv, mv, vv);
      p[i] = pv; m[i] = mv; v[i] = vv;
      if (p16) p16[i] = __float2bfloat16_rn(pv);
    }
  }
}

// Step counter and schedule ON THE DEVICE (TRAIN:102 get_linear_schedule_with_warmup + torch.optim.AdamW bias corrections):
// *step += 1, then hyper = {lr_t, beta1, beta2, eps, wd, 1-beta1^t, 1-beta2^t, grad_scale}.  One thread; it is part of the
// captured step graph, so a host that runs several replays ahead can never overwrite a value an earlier replay still needs.
__global__ void optim_schedule_kernel(long long* __restrict__ step, float* __restrict__ hyper, double base_lr, double beta1,
                                      double beta2, float eps, float wd, long long warmup, long long total, float grad_scale) {
  pdl_sync();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const long long t = *step + 1;
  *step = t;
  const long long s = t - 1;  // scheduler.step() follows optimizer.step(): update t uses multiplier(t - 1)
  double mult;
  if (s < warmup) mult = static_cast<double>(s) / static_cast<double>(warmup > 1 ? warmup : 1);
  else {
    const long long den = total - warmup > 1 ? total - warmup : 1;
    mult = static_cast<double>(total - s) / static_cast<double>(den);
    if (mult < 0.0) mult = 0.0;
  }
  hyper[0] = static_cast<float>(base_lr * mult);
  hyper[1] = static_cast<float>(beta1); hyper[2] = static_cast<float>(beta2); hyper[3] = eps; hyper[4] = wd;
  hyper[5] = static_cast<float>(1.0 - pow(beta1, static_cast<double>(t)));  // double like torch.optim.AdamW's Python floats
  hyper[6] = static_cast<float>(1.0 - pow(beta2, static_cast<double>(t)));
  hyper[7] = grad_scale;
}


// im2col of non-overlapping patches (conv1 of the CLIP ViT, kernel = stride = patch; TRAIN:225): images fp32 [B, C, H, W]
// -> bf16 [B * (H/p) * (W/p), C * p * p] with column order (c, ky, kx) = conv1.weight.view(width, -1).  One thread per
// 8 consecutive kx (p is a multiple of 8): two 16-byte loads, one 16-byte store.
__global__ void __launch_bounds__(256)
vit_patchify_kernel(const float* __restrict__ img, __nv_bfloat16* __restrict__ out, int B, int C, int H, int W, int p) {
  pdl_sync();
  const int gw = W / p, gh = H / p;
  const int cols = C * p * p;
  const long long groups = static_cast<long long>(B) * gh * gw * (cols / 8);
  const long long gi = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (gi >= groups) return;
  const int cg = static_cast<int>(gi % (cols / 8));
  const long long prow = gi / (cols / 8);
  const int col = cg * 8;
  const int c = col / (p * p), ky = (col / p) % p, kx = col % p;
  const int px = static_cast<int>(prow % gw), py = static_cast<int>((prow / gw) % gh);
  const long long b = prow / (static_cast<long long>(gw) * gh);
  const float* src = img + ((b * C + c) * H + (py * p + ky)) * static_cast<long long>(W) + px * p + kx;
  const float4 a0 = __ldg(reinterpret_cast<const float4*>(src)), a1 = __ldg(reinterpret_cast<const float4*>(src) + 1);
  uint4 u;
  u.x = pack_bf16x2(a0.x, a0.y); u.y = pack_bf16x2(a0.z, a0.w);
  u.z = pack_bf16x2(a1.x, a1.y); u.w = pack_bf16x2(a1.z, a1.w);
  *reinterpret_cast<uint4*>(out + prow * cols + col) = u;
}


// Packed rows -> right-padded [B, L, d] (bf16): dst[b, t] = src[start[b] + t] for t < len[b], else 0.  Restores the
// collate's layout (DNYT:957-972) where a consumer needs one row block per sample (the per-caption cross-attention K/V
// projection of the decode engine).  One warp per destination row, 16-byte accesses.
__global__ void __launch_bounds__(256)
unpack_rows_kernel(const __nv_bfloat16* __restrict__ src, const int32_t* __restrict__ start, const int32_t* __restrict__ len,
                   __nv_bfloat16* __restrict__ dst, long long rows, int L, int d) {
  pdl_sync();
  const long long row = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const int b = static_cast<int>(row / L), t = static_cast<int>(row % L);
  const bool valid = t < len[b];
  const uint4* sp = reinterpret_cast<const uint4*>(src + (static_cast<long long>(start[b]) + t) * d);
  uint4* dp = reinterpret_cast<uint4*>(dst + row * d);
  for (int v = lane; v < d / 8; v += 32) dp[v] = valid ? __ldg(sp + v) : make_uint4(0, 0, 0, 0);
}


// x = dropout(x) in place (bf16): activation dropout after the GELU of an FFN (config.activation_dropout: MFULL:649, 660, 684,
// 740, 874) and, with the same (seed, salt), the matching mask on the gradient in the backward pass.
__global__ void __launch_bounds__(256)
dropout_inplace_kernel(__nv_bfloat16* __restrict__ x, long long n, float p, const unsigned long long* __restrict__ rng, uint32_t salt) {
  pdl_sync();
  const long long i = (static_cast<long long>(blockIdx.x) * 256 + threadIdx.x) * 8;
  if (i >= n) return;
  const uint32_t seed = static_cast<uint32_t>(*rng), thr = drop_thresh(p);
  const float inv_keep = 1.f / (1.f - p);
  if (i + 8 <= n) {
    uint4 u = *reinterpret_cast<uint4*>(x + i);
    uint32_t* w = &u.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float2 f = unpack_bf16x2(w[k]);
      f.x = keep_elem(seed, salt, static_cast<uint64_t>(i + 2 * k), thr) ? f.x * inv_keep : 0.f;
      f.y = keep_elem(seed, salt, static_cast<uint64_t>(i + 2 * k + 1), thr) ? f.y * inv_keep : 0.f;
      w[k] = pack_bf16x2(f.x, f.y);
    }
    *reinterpret_cast<uint4*>(x + i) = u;
  } else {
    for (long long k = i; k < n; ++k)
      x[k] = keep_elem(seed, salt, static_cast<uint64_t>(k), thr) ? __float2bfloat16_rn(__bfloat162float(x[k]) * inv_keep) : __float2bfloat16_rn(0.f);
  }
}

// out = a + b (+ c) on bf16, fp32 math; gradient fan-in of the residual / state streams
__global__ void __launch_bounds__(256)
add_bf16_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b,
                const __nv_bfloat16* __restrict__ c, __nv_bfloat16* __restrict__ out, long long n) {
  pdl_sync();
  const long long i = (static_cast<long long>(blockIdx.x) * 256 + threadIdx.x) * 8;
  if (i + 8 <= n) {
    const uint4 ua = *reinterpret_cast<const uint4*>(a + i);
    const uint4 ub = *reinterpret_cast<const uint4*>(b + i);
    uint4 uc = make_uint4(0, 0, 0, 0);
    if (c) uc = *reinterpret_cast<const uint4*>(c + i);
    const uint32_t* pa = &ua.x; const uint32_t* pb = &ub.x; const uint32_t* pc = &uc.x;
    uint4 o; uint32_t* po = &o.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 fa = unpack_bf16x2(pa[k]), fb = unpack_bf16x2(pb[k]), fc = unpack_bf16x2(pc[k]);
      po[k] = pack_bf16x2(fa.x + fb.x + fc.x, fa.y + fb.y + fc.y);
    }
    *reinterpret_cast<uint4*>(out + i) = o;
  } else {
    for (long long k = i; k < n; ++k) {
      float v = __bfloat162float(a[k]) + __bfloat162float(b[k]);
      if (c) v += __bfloat162float(c[k]);
      out[k] = __float2bfloat16_rn(v);
    }
  }
}

// out[b] = [a[b] ; c[b]] along the row dimension (torch.cat(dim=1) at MFULL:668, 691); 16-byte vectors
__global__ void __launch_bounds__(256)
concat_rows_kernel(const uint4* __restrict__ a, const uint4* __restrict__ c, uint4* __restrict__ out, int B,
                   long long va, long long vc) {
  pdl_sync();
  const long long per = va + vc;
  const long long n = per * B;
  for (long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * 256) {
    const long long b = i / per, r = i % per;
    out[i] = r < va ? a[b * va + r] : c[b * vc + (r - va)];
  }
}

__global__ void rng_advance_kernel(unsigned long long* state) { *state += 0x9E3779B97F4A7C15ull; }


// dst[r, 0..ld_dst) = {src[r, 0..n), 0...}: re-pitches rows whose byte length is not a multiple of 16 so that TMA can
// read them (the NER-map gradient dz2 has 20-element = 40-byte rows)
__global__ void __launch_bounds__(256)
pad_rows_kernel(const __nv_bfloat16* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long rows, int n, int ld_dst) {
  pdl_sync();
  const long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= rows * ld_dst) return;
  const long long r = i / ld_dst;
  const int c = static_cast<int>(i % ld_dst);
  dst[i] = c < n ? src[r * n + c] : __float2bfloat16_rn(0.f);
}



// Global-norm gradient clipping (torch.nn.utils.clip_grad_norm_, TRAIN:365-366; off in the shipped scripts) folded into the
// fused optimizer: sum of squares of the flat gradient buffer -> grad_scale = base * min(1, max_norm / (base*||g|| + 1e-6)).
__global__ void __launch_bounds__(256)
sqnorm_kernel(const float* __restrict__ g, long long n, float* __restrict__ acc) {
  float s = 0.f;
  const long long stride = static_cast<long long>(gridDim.x) * 256 * 4;
  for (long long i = (static_cast<long long>(blockIdx.x) * 256 + threadIdx.x) * 4; i < n; i += stride) {
    if (i + 4 <= n) {
      const float4 x = __ldg(reinterpret_cast<const float4*>(g + i));
      s += x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;
    } else {
      for (long long k = i; k < n; ++k) s += g[k] * g[k];
    }
  }
  s = warp_sum(s);
  __shared__ float red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    acc[blockIdx.x] = t;  // per-block partial: summed in a fixed order below -> bit-reproducible norm
  }
}
__global__ void __launch_bounds__(256)
clip_scale_kernel(const float* __restrict__ partial, int nparts, float max_norm, float base, float* __restrict__ scale_out,
                  float* __restrict__ norm_out) {
  __shared__ double red[256];
  double s = 0.0;
  for (int i = threadIdx.x; i < nparts; i += 256) s += static_cast<double>(partial[i]);
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float norm = static_cast<float>(sqrt(red[0])) * fabsf(base);
    const float coef = fminf(1.f, max_norm / (norm + 1e-6f));
    *scale_out = base * coef;
    if (norm_out) *norm_out = norm;
  }
}

// dst[r, c] = bf16(src[r, c]) for c < n, rows re-pitched from ld_src to ld_dst (pad columns zero): gradients handed back
// by torch ops (e.g. the script-level CrossEntropyLoss on [rows, 50267] logits) become TMA-readable GEMM operands
__global__ void __launch_bounds__(256)
cast_rows_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long rows, int n, long long ld_src,
                 int ld_dst) {
  const long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= rows * ld_dst) return;
  const long long r = i / ld_dst;
  const int c = static_cast<int>(i % ld_dst);
  dst[i] = __float2bfloat16_rn(c < n ? src[r * ld_src + c] : 0.f);
}

// dst[j] (+)= sum_p src[p, j]: reduction of split-K partial weight gradients (fp32)
__global__ void __launch_bounds__(256)
sum_partials_kernel(const float* __restrict__ src, float* __restrict__ dst, int parts, long long len, int accumulate) {
  pdl_sync();
  const long long j = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (j >= len) return;
  float s = accumulate ? dst[j] : 0.f;
  for (int p = 0; p < parts; ++p) s += __ldg(src + p * len + j);
  dst[j] = s;
}

}  // namespace vb

using namespace vb;

extern "C" int vacnic_colsum(const void* x, float* out, int64_t rows, int32_t n, int64_t ld, void* stream) {
  VB_REQUIRE(x && out, "colsum: null pointer");
  VB_REQUIRE(rows >= 0 && n > 0 && ld >= n, "colsum: bad shape");
  if (rows == 0) return VACNIC_OK;
  const int col_tiles = (n + 255) / 256;
  const int sms = sm_count() > 0 ? sm_count() : 148;
  long long slabs = (4LL * sms + col_tiles - 1) / col_tiles;
  if (slabs > (rows + 63) / 64) slabs = (rows + 63) / 64;
  if (slabs < 1) slabs = 1;
  const int rpb = static_cast<int>((rows + slabs - 1) / slabs);
  dim3 grid(col_tiles, static_cast<unsigned>((rows + rpb - 1) / rpb));
  launch_pdl(colsum_kernel, grid, dim3(256), 0, static_cast<cudaStream_t>(stream), static_cast<const __nv_bfloat16*>(x), out, rows, n, ld, rpb);
  count_launch();
  return check_last("colsum");
}

extern "C" int vacnic_cast_f32_bf16(const float* src, void* dst, int64_t n, void* stream) {
  VB_REQUIRE(src && dst, "cast: null pointer");
  VB_REQUIRE((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0,
             "cast: pointers must be 16-byte aligned");
  if (n <= 0) return VACNIC_OK;
  const long long blocks = (n + 2047) / 2048;
  cast_f32_bf16_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      src, static_cast<__nv_bfloat16*>(dst), n);
  count_launch();
  return check_last("cast_f32_bf16");
}

extern "C" int vacnic_adamw(float* p, const float* g, float* m, float* v, void* p16, int64_t n, const float* hyper,
                            void* stream) {
  VB_REQUIRE(p && g && m && v && hyper, "adamw: null pointer");
  VB_REQUIRE(((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
               reinterpret_cast<uintptr_t>(v)) & 15) == 0 && (reinterpret_cast<uintptr_t>(p16) & 7) == 0,
             "adamw: buffers must be 16-byte aligned");
  if (n <= 0) return VACNIC_OK;
  const long long blocks = (n + 1023) / 1024;
  adamw_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      p, g, m, v, static_cast<__nv_bfloat16*>(p16), n, hyper);
  count_launch();
  return check_last("adamw");
}


extern "C" int vacnic_optim_schedule(int64_t* step, float* hyper, double base_lr, double beta1, double beta2, float eps,
                                     float weight_decay, int64_t warmup_steps, int64_t total_steps, float grad_scale,
                                     void* stream) {
  VB_REQUIRE(step && hyper, "optim_schedule: null pointer");
  VB_REQUIRE((reinterpret_cast<uintptr_t>(step) & 7) == 0, "optim_schedule: step must be 8-byte aligned");
  launch_pdl(optim_schedule_kernel, dim3(1), dim3(32), 0, static_cast<cudaStream_t>(stream), reinterpret_cast<long long*>(step),
             hyper, base_lr, beta1, beta2, eps, weight_decay, static_cast<long long>(warmup_steps),
             static_cast<long long>(total_steps), grad_scale);
  count_launch();
  return check_last("optim_schedule");
}


extern "C" int vacnic_vit_patchify(const float* images, void* out, int32_t batch, int32_t channels, int32_t height, int32_t width,
                                   int32_t patch, void* stream) {
  VB_REQUIRE(images && out, "vit_patchify: null pointer");
  VB_REQUIRE(batch >= 0 && channels > 0 && patch > 0 && patch % 8 == 0 && height % patch == 0 && width % patch == 0 && width % 4 == 0,
             "vit_patchify: patch must be a multiple of 8 that divides height and width");
  VB_REQUIRE((reinterpret_cast<uintptr_t>(images) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0, "vit_patchify: misaligned");
  if (batch == 0) return VACNIC_OK;
  const long long groups = static_cast<long long>(batch) * (height / patch) * (width / patch) * (channels * patch * patch / 8);
  launch_pdl(vit_patchify_kernel, dim3(static_cast<unsigned>((groups + 255) / 256)), dim3(256), 0, static_cast<cudaStream_t>(stream),
             images, static_cast<__nv_bfloat16*>(out), batch, channels, height, width, patch);
  count_launch();
  return check_last("vit_patchify");
}


extern "C" int vacnic_unpack_rows(const void* src, const int32_t* start, const int32_t* len, void* dst, int32_t batch, int32_t L,
                                  int32_t d, void* stream) {
  VB_REQUIRE(src && start && len && dst, "unpack_rows: null pointer");
  VB_REQUIRE(batch >= 0 && L > 0 && d > 0 && d % 8 == 0, "unpack_rows: d must be a multiple of 8");
  VB_REQUIRE((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0, "unpack_rows: misaligned");
  if (batch == 0) return VACNIC_OK;
  const long long rows = static_cast<long long>(batch) * L;
  launch_pdl(unpack_rows_kernel, dim3(static_cast<unsigned>((rows + 7) / 8)), dim3(256), 0, static_cast<cudaStream_t>(stream),
             static_cast<const __nv_bfloat16*>(src), start, len, static_cast<__nv_bfloat16*>(dst), rows, L, d);
  count_launch();
  return check_last("unpack_rows");
}


extern "C" int vacnic_dropout_inplace(void* x, int64_t n, float p_drop, const uint64_t* rng_state, uint32_t salt, void* stream) {
  VB_REQUIRE(x && rng_state, "dropout_inplace: null pointer");
  VB_REQUIRE(p_drop > 0.f && p_drop < 1.f, "dropout_inplace: p must be in (0, 1)");
  VB_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, "dropout_inplace: misaligned");
  if (n <= 0) return VACNIC_OK;
  launch_pdl(dropout_inplace_kernel, dim3(static_cast<unsigned>((n + 2047) / 2048)), dim3(256), 0, static_cast<cudaStream_t>(stream),
             static_cast<__nv_bfloat16*>(x), static_cast<long long>(n), p_drop, reinterpret_cast<const unsigned long long*>(rng_state), salt);
  count_launch();
  return check_last("dropout_inplace");
}

extern "C" int vacnic_add_bf16(const void* a, const void* b, const void* c, void* out, int64_t n, void* stream) {
  VB_REQUIRE(a && b && out, "add_bf16: null pointer");
  VB_REQUIRE(((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c) |
               reinterpret_cast<uintptr_t>(out)) & 15) == 0, "add_bf16: buffers must be 16-byte aligned");
  if (n <= 0) return VACNIC_OK;
  const long long blocks = (n + 2047) / 2048;
  launch_pdl(add_bf16_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, static_cast<cudaStream_t>(stream), static_cast<const __nv_bfloat16*>(a), static_cast<const __nv_bfloat16*>(b), static_cast<const __nv_bfloat16*>(c), static_cast<__nv_bfloat16*>(out), n);
  count_launch();
  return check_last("add_bf16");
}

extern "C" int vacnic_concat_rows(const void* a, const void* b, void* out, int32_t B, int64_t rows_a, int64_t rows_b,
                                  int32_t d, void* stream) {
  VB_REQUIRE(a && b && out, "concat_rows: null pointer");
  VB_REQUIRE(B > 0 && rows_a >= 0 && rows_b >= 0 && d > 0 && d % 8 == 0, "concat_rows: bad shape");
  const long long va = rows_a * d / 8, vc = rows_b * d / 8;
  const long long n = (va + vc) * B;
  if (n == 0) return VACNIC_OK;
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  launch_pdl(concat_rows_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, static_cast<cudaStream_t>(stream), static_cast<const uint4*>(a), static_cast<const uint4*>(b), static_cast<uint4*>(out), B, va, vc);
  count_launch();
  return check_last("concat_rows");
}

extern "C" int vacnic_rng_advance(uint64_t* state, void* stream) {
  VB_REQUIRE(state, "rng_advance: null pointer");
  rng_advance_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<unsigned long long*>(state));
  count_launch();
  return check_last("rng_advance");
}

extern "C" int vacnic_pad_rows(const void* src, void* dst, int64_t rows, int32_t n, int32_t ld_dst, void* stream) {
  VB_REQUIRE(src && dst && rows > 0 && n > 0 && ld_dst >= n, "pad_rows: bad arguments");
  const long long total = rows * ld_dst;
  launch_pdl(pad_rows_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), static_cast<const __nv_bfloat16*>(src), static_cast<__nv_bfloat16*>(dst), rows, n, ld_dst);
  count_launch();
  return check_last("pad_rows");
}

extern "C" int vacnic_sum_partials(const float* src, float* dst, int32_t parts, int64_t len, int32_t accumulate, void* stream) {
  VB_REQUIRE(src && dst && parts > 0 && len > 0, "sum_partials: bad arguments");
  launch_pdl(sum_partials_kernel, dim3(static_cast<unsigned>((len + 255) / 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), src, dst, parts, len, accumulate);
  count_launch();
  return check_last("sum_partials");
}

extern "C" int vacnic_cast_rows_f32_bf16(const float* src, void* dst, int64_t rows, int32_t n, int64_t ld_src, int32_t ld_dst,
                                         void* stream) {
  VB_REQUIRE(src && dst && rows > 0 && n > 0 && ld_src >= n && ld_dst >= n, "cast_rows: bad arguments");
  const long long total = rows * ld_dst;
  cast_rows_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      src, static_cast<__nv_bfloat16*>(dst), rows, n, ld_src, ld_dst);
  count_launch();
  return check_last("cast_rows");
}

extern "C" int vacnic_clip_grad_scale(const float* g, int64_t n, float max_norm, float base_scale, float* scratch,
                                      float* scale_out, float* norm_out, void* stream) {
  VB_REQUIRE(g && scratch && scale_out && n > 0 && max_norm > 0.f, "clip_grad_scale: bad arguments");
  VB_REQUIRE((reinterpret_cast<uintptr_t>(g) & 15) == 0, "clip_grad_scale: gradient buffer must be 16-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  long long blocks = (n + 1023) / 1024;
  if (blocks > VACNIC_CLIP_SCRATCH_FLOATS) blocks = VACNIC_CLIP_SCRATCH_FLOATS;
  sqnorm_kernel<<<static_cast<unsigned>(blocks), 256, 0, s>>>(g, n, scratch);
  clip_scale_kernel<<<1, 256, 0, s>>>(scratch, static_cast<int>(blocks), max_norm, base_scale, scale_out, norm_out);
  count_launch(2);
  return check_last("clip_grad_scale");
}
