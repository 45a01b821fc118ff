// Loss kernels of the VACNIC training step (script-level code in the reference):
//   token cross-entropy  CrossEntropyLoss(ignore_index=pad)                 TRAIN:287, 816
//   CoLaM margin loss    pool -> L2 norm -> diag cosine -> hinge(margin)    TRAIN:292-309, 178-182, 820
//   SECLA                [B,B,N,F] similarity -> max -> mean -> row CE x2   TRAIN:326-330, 631-660
// All reductions are warp-shuffle / shared-memory based and deterministic (no atomics).
#include <float.h>

#include "common.h"
#include "ptx.cuh"

namespace vb {

__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float s = 0.f;
  for (int w = 0; w < nw; ++w) s += red[w];
  return s;
}
__device__ __forceinline__ float block_max(float v, float* red) {
  v = warp_max(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float s = -INFINITY;
  for (int w = 0; w < nw; ++w) s = fmaxf(s, red[w]);
  return s;
}

// ------------------------------------------------------------------ cross entropy
__global__ void __launch_bounds__(512)
ce_fwd_kernel(const float* __restrict__ logits, const long long* __restrict__ targets, float* __restrict__ lse,
              float* __restrict__ row_loss, int V, long long ld, long long ignore_index) {
  __shared__ float red[32];
  const long long row = blockIdx.x;
  const float* lp = logits + row * ld;
  float mx = -INFINITY;
  for (int c = threadIdx.x; c < V; c += blockDim.x) mx = fmaxf(mx, lp[c]);
  mx = block_max(mx, red);
  float s = 0.f;
  for (int c = threadIdx.x; c < V; c += blockDim.x) s += __expf(lp[c] - mx);
  s = block_sum(s, red);
  if (threadIdx.x == 0) {
    const float l = mx + logf(s);
    lse[row] = l;
    const long long t = targets[row];
    row_loss[row] = (t == ignore_index) ? 0.f : l - lp[t];
  }
}

// out[0] = mean loss over non-ignored rows, out[1] = number of non-ignored rows
__global__ void __launch_bounds__(1024)
ce_finalize_kernel(const float* __restrict__ row_loss, const long long* __restrict__ targets, float* __restrict__ out,
                   long long rows, long long ignore_index) {
  __shared__ float red[32];
  float s = 0.f, n = 0.f;
  for (long long r = threadIdx.x; r < rows; r += blockDim.x) {
    s += row_loss[r];
    n += targets[r] != ignore_index ? 1.f : 0.f;
  }
  s = block_sum(s, red);
  n = block_sum(n, red);
  if (threadIdx.x == 0) {
    out[0] = s / n;  // 0/0 = nan, like torch when every target is ignored
    out[1] = n;
  }
}

// dlogits[r, c] = (softmax(logits[r])[c] - [c == t_r]) * gscale[0] * coef / count ; 0 for ignored rows
__global__ void __launch_bounds__(512)
ce_bwd_kernel(const float* __restrict__ logits, const float* __restrict__ lse, const long long* __restrict__ targets,
              const float* __restrict__ stats, const float* __restrict__ gscale, float coef,
              __nv_bfloat16* __restrict__ dlogits, int V, long long ld, long long ignore_index) {
  const long long row = blockIdx.x;
  const long long t = targets[row];
  __nv_bfloat16* dp = dlogits + row * ld;
  if (t == ignore_index) {
    for (int c = threadIdx.x; c < ld; c += blockDim.x) dp[c] = __float2bfloat16_rn(0.f);
    return;
  }
  const float scale = coef * (gscale ? gscale[0] : 1.f) / stats[1];
  const float l = lse[row];
  const float* lp = logits + row * ld;
  for (int c = threadIdx.x; c < ld; c += blockDim.x) {
    float g = 0.f;
    if (c < V) g = (__expf(lp[c] - l) - (c == t ? 1.f : 0.f)) * scale;
    dp[c] = __float2bfloat16_rn(g);
  }
}

// ------------------------------------------------------------------ CoLaM
// One block per caption. stats[b] = {cos, |a|, |b|, count, active}; pooled_a/b fp32 [B, d].
__global__ void __launch_bounds__(256)
colam_fwd_kernel(const __nv_bfloat16* __restrict__ h, const __nv_bfloat16* __restrict__ hg,
                 const long long* __restrict__ tgt, float* __restrict__ pooled_a, float* __restrict__ pooled_b,
                 float* __restrict__ stats, int T, int d, long long pad, float margin) {
  __shared__ float red[32];
  const int b = blockIdx.x;
  float cnt = 0.f;
  for (int t = 0; t < T; ++t) cnt += tgt[static_cast<long long>(b) * T + t] != pad ? 1.f : 0.f;
  float dot = 0.f, na = 0.f, nb = 0.f;
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    float sa = 0.f, sb = 0.f;
    for (int t = 0; t < T; ++t) {
      if (tgt[static_cast<long long>(b) * T + t] != pad) {
        sa += __bfloat162float(h[(static_cast<long long>(b) * T + t) * d + c]);
        sb += __bfloat162float(hg[(static_cast<long long>(b) * T + t) * d + c]);
      }
    }
    // pool(): sum / count, nan_to_num(nan=1.0)  (TRAIN:178-182)
    sa = cnt > 0.f ? sa / cnt : 1.f;
    sb = cnt > 0.f ? sb / cnt : 1.f;
    pooled_a[static_cast<long long>(b) * d + c] = sa;
    pooled_b[static_cast<long long>(b) * d + c] = sb;
    dot += sa * sb; na += sa * sa; nb += sb * sb;
  }
  dot = block_sum(dot, red); na = sqrtf(block_sum(na, red)); nb = sqrtf(block_sum(nb, red));
  if (threadIdx.x == 0) {
    const float cs = dot / (na * nb);
    float* s = stats + b * 8;
    s[0] = cs; s[1] = na; s[2] = nb; s[3] = cnt;
    s[4] = (margin - cs > 0.f) ? 1.f : 0.f;
    s[5] = fmaxf(0.f, margin - cs);
  }
}
__global__ void colam_finalize_kernel(const float* __restrict__ stats, float* __restrict__ loss, int B) {
  float s = 0.f;
  for (int b = threadIdx.x; b < B; b += 32) s += stats[b * 8 + 5];
  s = warp_sum(s);
  if (threadIdx.x == 0) loss[0] = s / B;
}
// dh[b,t,:] (+)= m_t / cnt * dL/da ; grid (B, T)
__global__ void __launch_bounds__(256)
colam_bwd_kernel(const float* __restrict__ pooled_a, const float* __restrict__ pooled_b, const float* __restrict__ stats,
                 const long long* __restrict__ tgt, const float* __restrict__ gscale, float coef,
                 __nv_bfloat16* __restrict__ dh, int B, int T, int d, long long pad, int accumulate) {
  const int b = blockIdx.x, t = blockIdx.y;
  const float* s = stats + b * 8;
  const bool on = tgt[static_cast<long long>(b) * T + t] != pad && s[4] > 0.f && s[3] > 0.f;
  const float cs = s[0], na = s[1], nb = s[2];
  // L = mean_b max(0, margin - cos_b):  dL/dcos = -1/B when active
  const float g = on ? -(gscale ? gscale[0] : 1.f) * coef / B / s[3] : 0.f;
  __nv_bfloat16* out = dh + (static_cast<long long>(b) * T + t) * d;
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    const float a = pooled_a[static_cast<long long>(b) * d + c], bb = pooled_b[static_cast<long long>(b) * d + c];
    float v = g * (bb / (na * nb) - cs * a / (na * na));
    if (accumulate) v += __bfloat162float(out[c]);
    out[c] = __float2bfloat16_rn(v);
  }
}

// ------------------------------------------------------------------ SECLA
// names fp32 [B, N, d] (no grad), face bf16 [B, F, d].  One block.  Workspace (fp32):
//   M  [B*N, B*F] similarity, dM [B*N, B*F] its gradient.
// loss[0] = CE_rows(A) + CE_rows(C), A[i,j] = mean_n max_f M[(i,n),(j,f)], C[i,j] = mean_f max_n M[(j,n),(i,f)].
// M[(b,n), (b',f)] = names[b,n,:] . face[b',f,:]   — one warp per entry, spread over the whole GPU
__global__ void __launch_bounds__(256)
secla_sim_kernel(const float* __restrict__ names, const __nv_bfloat16* __restrict__ face, float* __restrict__ Mw,
                 float* __restrict__ dM, int BN, int BF, int d) {
  const int e = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (e >= BN * BF) return;
  const int r = e / BF, c = e % BF;
  const float* np = names + static_cast<long long>(r) * d;
  const __nv_bfloat16* fp = face + static_cast<long long>(c) * d;
  float s = 0.f;
  for (int k = lane; k < d; k += 32) s += np[k] * __bfloat162float(fp[k]);
  s = warp_sum(s);
  if (lane == 0) { Mw[e] = s; dM[e] = 0.f; }
}

__global__ void __launch_bounds__(1024)
secla_fwd_kernel(const float* __restrict__ names, const __nv_bfloat16* __restrict__ face, float* __restrict__ Mw,
                 float* __restrict__ dM, float* __restrict__ loss, int B, int N, int F, int d) {
  extern __shared__ float sm[];
  float* A = sm;               // [B, B]
  float* C = A + B * B;        // [B, B]
  float* red = C + B * B;      // [32]
  const int BF = B * F;
  // Mw / dM were filled by secla_sim_kernel (previous launch on the same stream)
  for (int e = threadIdx.x; e < B * B; e += blockDim.x) {
    const int i = e / B, j = e % B;
    float a = 0.f;
    for (int n = 0; n < N; ++n) {
      float mx = -INFINITY;
      for (int f = 0; f < F; ++f) mx = fmaxf(mx, Mw[(i * N + n) * BF + j * F + f]);
      a += mx;
    }
    A[e] = a / N;
    float c = 0.f;
    for (int f = 0; f < F; ++f) {
      float mx = -INFINITY;
      for (int n = 0; n < N; ++n) mx = fmaxf(mx, Mw[(j * N + n) * BF + i * F + f]);
      c += mx;
    }
    C[e] = c / F;
  }
  __syncthreads();
  // row-wise cross entropy with target = row index; gradient wrt logits = (softmax - onehot) / B
  float l = 0.f;
  for (int i = threadIdx.x; i < 2 * B; i += blockDim.x) {
    float* L = (i < B ? A : C) + (i % B) * B;
    const int tgt = i % B;
    float mx = -INFINITY;
    for (int j = 0; j < B; ++j) mx = fmaxf(mx, L[j]);
    float s = 0.f;
    for (int j = 0; j < B; ++j) s += __expf(L[j] - mx);
    const float lse = mx + logf(s);
    l += (lse - L[tgt]) / B;
    for (int j = 0; j < B; ++j) {
      const float g = (__expf(L[j] - lse) - (j == tgt ? 1.f : 0.f)) / B;
      L[j] = g;  // overwrite logits with their gradient
    }
  }
  l = block_sum(l, red);
  if (threadIdx.x == 0) loss[0] = l;
  __syncthreads();
  // route gradients through the (first) arg-max, like torch.max(dim).values
  for (int e = threadIdx.x; e < B * B; e += blockDim.x) {
    const int i = e / B, j = e % B;
    for (int n = 0; n < N; ++n) {
      int best = 0; float mx = -INFINITY;
      for (int f = 0; f < F; ++f) { const float v = Mw[(i * N + n) * BF + j * F + f]; if (v > mx) { mx = v; best = f; } }
      atomicAdd(&dM[(i * N + n) * BF + j * F + best], A[e] / N);
    }
    for (int f = 0; f < F; ++f) {
      int best = 0; float mx = -INFINITY;
      for (int n = 0; n < N; ++n) { const float v = Mw[(j * N + n) * BF + i * F + f]; if (v > mx) { mx = v; best = n; } }
      atomicAdd(&dM[(j * N + best) * BF + i * F + f], C[e] / F);
    }
  }
}
// dface[c, :] (+)= gscale * sum_r dM[r, c] * names[r, :]; grid = B*F blocks
__global__ void __launch_bounds__(256)
secla_bwd_kernel(const float* __restrict__ dM, const float* __restrict__ names, const float* __restrict__ gscale,
                 float coef, __nv_bfloat16* __restrict__ dface, int BN, int BF, int d, int accumulate) {
  const int c = blockIdx.x;
  const float gs = coef * (gscale ? gscale[0] : 1.f);
  for (int k = threadIdx.x; k < d; k += blockDim.x) {
    float s = 0.f;
    for (int r = 0; r < BN; ++r) s += dM[r * BF + c] * names[static_cast<long long>(r) * d + k];
    s *= gs;
    __nv_bfloat16* o = dface + static_cast<long long>(c) * d + k;
    if (accumulate) s += __bfloat162float(*o);
    *o = __float2bfloat16_rn(s);
  }
}

}  // namespace vb

using namespace vb;

extern "C" int vacnic_ce_fwd(const float* logits, const int64_t* targets, float* lse, float* row_loss, float* out,
                             int64_t rows, int32_t V, int64_t ld, int64_t ignore_index, void* stream) {
  VB_REQUIRE(logits && targets && lse && row_loss && out, "ce_fwd: null pointer");
  VB_REQUIRE(rows > 0 && V > 0 && ld >= V, "ce_fwd: bad shape");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  ce_fwd_kernel<<<static_cast<unsigned>(rows), 512, 0, s>>>(logits, reinterpret_cast<const long long*>(targets), lse,
                                                            row_loss, V, ld, ignore_index);
  ce_finalize_kernel<<<1, 1024, 0, s>>>(row_loss, reinterpret_cast<const long long*>(targets), out, rows, ignore_index);
  count_launch(2);
  return check_last("ce_fwd");
}

extern "C" int vacnic_ce_bwd(const float* logits, const float* lse, const int64_t* targets, const float* stats,
                             const float* gscale, float coef, void* dlogits, int64_t rows, int32_t V, int64_t ld,
                             int64_t ignore_index, void* stream) {
  VB_REQUIRE(logits && lse && targets && stats && dlogits, "ce_bwd: null pointer");
  VB_REQUIRE(rows > 0 && V > 0 && ld >= V, "ce_bwd: bad shape");
  ce_bwd_kernel<<<static_cast<unsigned>(rows), 512, 0, static_cast<cudaStream_t>(stream)>>>(
      logits, lse, reinterpret_cast<const long long*>(targets), stats, gscale, coef,
      static_cast<__nv_bfloat16*>(dlogits), V, ld, ignore_index);
  count_launch();
  return check_last("ce_bwd");
}

extern "C" int vacnic_colam_fwd(const void* h, const void* h_guide, const int64_t* tgt_ids, float* pooled_a,
                                float* pooled_b, float* stats, float* loss, int32_t B, int32_t T, int32_t d,
                                int64_t pad_id, float margin, void* stream) {
  VB_REQUIRE(h && h_guide && tgt_ids && pooled_a && pooled_b && stats && loss, "colam_fwd: null pointer");
  VB_REQUIRE(B > 0 && T > 0 && d > 0, "colam_fwd: bad shape");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  colam_fwd_kernel<<<B, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(h), static_cast<const __nv_bfloat16*>(h_guide),
                                     reinterpret_cast<const long long*>(tgt_ids), pooled_a, pooled_b, stats, T, d,
                                     pad_id, margin);
  colam_finalize_kernel<<<1, 32, 0, s>>>(stats, loss, B);
  count_launch(2);
  return check_last("colam_fwd");
}

extern "C" int vacnic_colam_bwd(const float* pooled_a, const float* pooled_b, const float* stats,
                                const int64_t* tgt_ids, const float* gscale, float coef, void* dh, int32_t B,
                                int32_t T, int32_t d, int64_t pad_id, int32_t accumulate, void* stream) {
  VB_REQUIRE(pooled_a && pooled_b && stats && tgt_ids && dh, "colam_bwd: null pointer");
  VB_REQUIRE(B > 0 && T > 0 && d > 0, "colam_bwd: bad shape");
  dim3 grid(B, T);
  colam_bwd_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      pooled_a, pooled_b, stats, reinterpret_cast<const long long*>(tgt_ids), gscale, coef,
      static_cast<__nv_bfloat16*>(dh), B, T, d, pad_id, accumulate);
  count_launch();
  return check_last("colam_bwd");
}

extern "C" int64_t vacnic_secla_workspace_bytes(int32_t B, int32_t N, int32_t F) {
  return 2LL * B * N * B * F * static_cast<int64_t>(sizeof(float));
}

extern "C" int vacnic_secla_fwd(const float* names, const void* face, float* workspace, float* loss, int32_t B,
                                int32_t N, int32_t F, int32_t d, void* stream) {
  VB_REQUIRE(names && face && workspace && loss, "secla_fwd: null pointer");
  VB_REQUIRE(B > 0 && N > 0 && F > 0 && d > 0 && B <= 64, "secla_fwd: bad shape (B <= 64)");
  const size_t smem = (2 * B * B + 32) * sizeof(float);
  float* Mw = workspace;
  float* dM = workspace + static_cast<long long>(B) * N * B * F;
  secla_sim_kernel<<<(B * N * B * F + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      names, static_cast<const __nv_bfloat16*>(face), Mw, dM, B * N, B * F, d);
  count_launch();
  secla_fwd_kernel<<<1, 1024, smem, static_cast<cudaStream_t>(stream)>>>(
      names, static_cast<const __nv_bfloat16*>(face), Mw, dM, loss, B, N, F, d);
  count_launch();
  return check_last("secla_fwd");
}

extern "C" int vacnic_secla_bwd(const float* workspace, const float* names, const float* gscale, float coef,
                                void* dface, int32_t B, int32_t N, int32_t F, int32_t d, int32_t accumulate,
                                void* stream) {
  VB_REQUIRE(workspace && names && dface, "secla_bwd: null pointer");
  const float* dM = workspace + static_cast<long long>(B) * N * B * F;
  secla_bwd_kernel<<<B * F, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      dM, names, gscale, coef, static_cast<__nv_bfloat16*>(dface), B * N, B * F, d, accumulate);
  count_launch();
  return check_last("secla_bwd");
}
