// HBM-bound stages of the train step: casts / weight-shadow transposes, embedding gather and
// scatter, reparameterised sample + KL, vocab softmax cross-entropy (fwd + bwd in one pass),
// column sums (bias gradients), global norm and the fused TF-form Adam update.
// Each kernel is coalesced and vectorised; reductions use warp shuffles then one atomic per warp/CTA.
#include <curand_kernel.h>
#include "ops.h"

namespace vc {

static inline int grid_for(long long n, int block, int per_thread = 1) {
  long long g = (n + (long long)block * per_thread - 1) / ((long long)block * per_thread);
  const long long cap = (long long)num_sms() * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

// ------------------------------------------------------------------------------------------
__global__ void k_cast_f32_bf16(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long rows,
                                int cols, long long ld_src, long long ld_dst) {
  const long long total = rows * cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols;
    const int c = (int)(i - r * cols);
    dst[r * ld_dst + c] = __float2bfloat16(src[r * ld_src + c]);
  }
}
int cast_f32_bf16(cudaStream_t s, const float* src, void* dst, long long rows, int cols, long long ld_src,
                  long long ld_dst) {
  {
    ProfScope ps(s, "cast_f32_bf16");
    k_cast_f32_bf16<<<grid_for(rows * cols, 256, 4), 256, 0, s>>>(src, (__nv_bfloat16*)dst, rows, cols, ld_src, ld_dst);
  }
  VC_CUDA(cudaGetLastError());
  return VC_OK;
}

// dst[perm(c), r] = bf16(src[r, c]); perm groups the 4 LSTM gates of `upt` hidden units into one
// contiguous block of 4*upt rows (gate_h > 0), otherwise identity. 32x32 smem tile transpose.
__global__ void k_transpose_cast(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int R, int C,
                                 long long ld_src, long long ld_dst, int gate_h, int upt) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < R && c < C) ? src[(long long)r * ld_src + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (c < C && r < R) {
      int cd = c;
      if (gate_h > 0) {
        const int g = c / gate_h, u = c - g * gate_h;
        cd = (u / upt) * (4 * upt) + g * upt + (u % upt);
      }
      dst[(long long)cd * ld_dst + r] = __float2bfloat16(tile[threadIdx.x][i]);
    }
  }
}
int transpose_cast(cudaStream_t s, const float* src, void* dst, int R, int C, long long ld_src, long long ld_dst,
                   int gate_h, int upt) {
  dim3 grid((C + 31) / 32, (R + 31) / 32), block(32, 8);
  {
    ProfScope ps(s, "transpose_cast");
    k_transpose_cast<<<grid, block, 0, s>>>(src, (__nv_bfloat16*)dst, R, C, ld_src, ld_dst, gate_h, upt);
  }
  VC_CUDA(cudaGetLastError());
  return VC_OK;
}

// tf.nn.dropout / DropoutWrapper keep masks (decoder.py:85-87, rnn_model.py:45-46) drawn on the device when the caller
// gives none: mask[i] = 1 if Philox4x32-10(seed ^ salt, subsequence = i / 4, offset = step) uniform < keep_prob else 0.
__global__ void k_keep_mask(float* __restrict__ mask, long long n, float keep_prob, unsigned long long seed,
                            unsigned long long offset) {
  const long long n4 = (n + 3) / 4;
  for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < n4; q += (long long)gridDim.x * blockDim.x) {
    curandStatePhilox4_32_10_t st;
    curand_init(seed, (unsigned long long)q, offset, &st);
    const float4 u = curand_uniform4(&st);
    const float v[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (4 * q + j < n) mask[4 * q + j] = v[j] <= keep_prob ? 1.f : 0.f;
  }
}
int keep_mask(cudaStream_t s, float* mask, long long n, float keep_prob, unsigned long long seed, unsigned long long offset) {
  {
    ProfScope ps(s, "keep_mask");
    k_keep_mask<<<grid_for((n + 3) / 4, 256, 2), 256, 0, s>>>(mask, n, keep_prob, seed, offset);
  }
  VC_CUDA(cudaGetLastError());
  return VC_OK;
}

// Every bf16 operand shadow of one optimiser step in ONE launch (the per-tensor casts / transposes above were 13 launches
// of 2-15 us each per step). Job kinds: 0 = dst[r, c] = bf16(src[r, c]) (row pitches may differ), vectorised by 4 when
// the row length and both pitches allow; 1 = the gate-interleaving transpose of k_transpose_cast. Blocks are dealt to
// jobs by a prefix table.
__global__ void __launch_bounds__(256) k_refresh_multi(const RefreshJobs jobs) {
  __shared__ float tile[32][33];
  int j = 0;
  while (j + 1 < jobs.n && (int)blockIdx.x >= jobs.j[j + 1].blk0) ++j;
  const RefreshJob& J = jobs.j[j];
  const unsigned int b = blockIdx.x - J.blk0;
  const float* __restrict__ src = J.src;
  __nv_bfloat16* __restrict__ dst = J.dst;
  if (J.kind == 0) {
    const unsigned int total = (unsigned int)J.rows * (unsigned int)J.cols;
    if (J.vec4) {
      const unsigned int c4n = (unsigned int)J.cols >> 2;
      const unsigned int i0 = (b * 256u + threadIdx.x) * 2u;  // two float4 per thread
#pragma unroll
      for (unsigned int k = 0; k < 2; ++k) {
        const unsigned int i = i0 + k;
        if (i * 4u >= total) break;
        const unsigned int r = i / c4n, c = (i - r * c4n) * 4u;
        const float4 v = *reinterpret_cast<const float4*>(src + (size_t)r * J.ld_src + c);
        uint2 o;
        o.x = pack_bf16(v.x, v.y);
        o.y = pack_bf16(v.z, v.w);
        *reinterpret_cast<uint2*>(dst + (size_t)r * J.ld_dst + c) = o;
      }
    } else {
      const unsigned int i0 = b * 2048u + threadIdx.x;
#pragma unroll
      for (unsigned int k = 0; k < 8; ++k) {
        const unsigned int i = i0 + k * 256u;
        if (i >= total) break;
        const unsigned int r = i / (unsigned int)J.cols, c = i - r * (unsigned int)J.cols;
        dst[(size_t)r * J.ld_dst + c] = __float2bfloat16(src[(size_t)r * J.ld_src + c]);
      }
    }
    return;
  }
  if (J.kind == 2) {
    const unsigned int total = (unsigned int)J.rows * (unsigned int)J.cols;
    const unsigned int i0 = b * 2048u + threadIdx.x;
    const unsigned int cin = (unsigned int)J.gate_h, cout = (unsigned int)J.cols;
#pragma unroll
    for (unsigned int k = 0; k < 8; ++k) {
      const unsigned int i = i0 + k * 256u;
      if (i >= total) break;
      const unsigned int co = i % cout, r = i / cout, ci = r % cin, tap = r / cin;
      dst[(size_t)ci * 9 * cout + (8 - tap) * cout + co] = __float2bfloat16(src[i]);
    }
    return;
  }
  const int tiles_c = (J.cols + 31) / 32;
  const int c0 = (int)(b % tiles_c) * 32, r0 = (int)(b / tiles_c) * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + tx;
    tile[i][tx] = (r < J.rows && c < J.cols) ? src[(size_t)r * J.ld_src + c] : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + tx;
    if (c < J.cols && r < J.rows) {
      int cd = c;
      if (J.gate_h > 0) {
        const int g = c / J.gate_h, u = c - g * J.gate_h;
        cd = (u / J.upt) * (4 * J.upt) + g * J.upt + (u % J.upt);
      }
      dst[(size_t)cd * J.ld_dst + r] = __float2bfloat16(tile[tx][i]);
    }
  }
}
int refresh_multi(cudaStream_t s, RefreshJobs& jobs) {
  if (jobs.n <= 0) return VC_OK;
  int blocks = 0;
  for (int i = 0; i < jobs.n; ++i) {
    RefreshJob& J = jobs.j[i];
    J.blk0 = blocks;
    if (J.kind == 2) {
      blocks += (int)(((long long)J.rows * J.cols + 2047) / 2048);
    } else if (J.kind == 0) {
      J.vec4 = (J.cols % 4 == 0 && J.ld_src % 4 == 0 && J.ld_dst % 4 == 0 && (reinterpret_cast<uintptr_t>(J.src) & 15) == 0 &&
                (reinterpret_cast<uintptr_t>(J.dst) & 7) == 0) ? 1 : 0;
      const long long total = (long long)J.rows * J.cols;
      if (total >= (1LL << 31)) return set_error(VC_E_SHAPE, "refresh_multi: tensor too large");
      blocks += (int)((total + 2047) / 2048);
    } else {
      blocks += ((J.cols + 31) / 32) * ((J.rows + 31) / 32);
    }
  }
  {
    ProfScope ps(s, "refresh_shadows");
    k_refresh_multi<<<blocks, 256, 0, s>>>(jobs);
  }
  VC_CUDA(cudaGetLastError());
  return VC_OK;
}

// out_bf16[(b*C + c), :] = bf16(src[b, :])  for c in [0, C): feature tiling (main.py:84-89) applied
// after the projection (Q7); written to up to two destinations (encoder and decoder X slot 0).
__global__ void k_tile_cast(const float* __restrict__ src, __nv_bfloat16* __restrict__ d0, __nv_bfloat16* __restrict__ d1,
                            int B, int C, int E) {
  const long long total = (long long)B * C * E;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int e = (int)(i % E);
    const long long n = i / E;
    const __nv_bfloat16 v = __float2bfloat16(src[(n / C) * E + e]);
    if (d0) d0[i] = v;
    if (d1) d1[i] = v;
  }
}
int tile_cast(cudaStream_t s, const float* src, void* d0, void* d1, int B, int C, int E) {
  {
    ProfScope ps(s, "tile_cast");
    k_tile_cast<<<grid_for((long long)B * C * E, 256, 2), 256, 0, s>>>(src, (__nv_bfloat16*)d0, (__nv_bfloat16*)d1, B, C, E);
  }
  VC_CUDA(cudaGetLastError());
  return VC_OK;
}

// dst[b, e] = sum_c (a[(b*C+c), e] + b2[(b*C+c), e])  -> fp32 and bf16 copies
__global__ void k_tile_reduce(const float* __restrict__ a, const float* __restrict__ b2, float* __restrict__ dst_f,
                              __nv_bfloat16* __restrict__ dst_h, int B, int C, int E) {
  const long long total = (long long)B * E;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int e = (int)(i % E);
    const long long b = i / E;
    float acc = 0.f;
    for (int c = 0; c < C; ++c) {
      const long long o = (b * C + c) * E + e;
      acc += a[o];
      if (b2) acc += b2[o];
    }
    if (dst_f) dst_f[i] = acc;
    if (dst_h) dst_h[i] = __float2bfloat16(acc);
  }
}
int tile_reduce(cudaStream_t s, const float* a, const float* b2, float* dst_f, void* dst_h, int B, int C, int E) {
  {
    ProfScope ps(s, "tile_reduce");
    k_tile_reduce<<<grid_for((long long)B * E, 256), 256, 0, s>>>(a, b2, dst_f, (__nv_bfloat16*)dst_h, B, C, E);
  }
  VC_CUDA(cudaGetLastError());
  return VC_OK;
}

// ------------------------------------------------------------------------------------------
// Embedding gather: X[t, n, :] = table_bf16[tok[n, t], :] (* keep_mask[n, t, :] / keep). One warp per row.
__global__ void k_embed_gather(const __nv_bfloat16* __restrict__ table, const int* __restrict__ tok,
                               __nv_bfloat16* __restrict__ X, const float* __restrict__ keep_mask, float inv_keep, int N,
                               int T, int E, int V) {
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  for (long long row = blockIdx.x * (long long)warps_per_block + (threadIdx.x >> 5); row < (long long)N * T;
       row += (long long)gridDim.x * warps_per_block) {
    const int t = (int)(row / N), n = (int)(row - (long long)t * N);
    int id = tok[(long long)n * T + t];
    id = id < 0 ? 0 : (id >= V ? V - 1 : id);
    const __nv_bfloat16* src = table + (long long)id * E;
    __nv_bfloat16* dst = X + row * E;
    if (keep_mask == nullptr) {
      for (int e = lane * 8; e < E; e += 256) *reinterpret_cast<uint4*>(dst + e) = *reinterpret_cast<const uint4*>(src + e);
    } else {
      const float* mk = keep_mask + ((long long)n * T + t) * E;
      for (int e = lane; e < E; e += 32) dst[e] = __float2bfloat16(__bfloat162float(src[e]) * mk[e] * inv_keep);
    }
  }
}
int embed_gather(cudaStream_t s, const void* table, const int* tok, void* X, const float* keep_mask, float inv_keep,
                 int N, int T, int E, int V) {
  if (E % 8 != 0) return set_error(VC_E_SHAPE, "embed_size must be a multiple of 8");
  {
    ProfScope ps(s, "embed_gather");
    k_embed_gather<<<grid_for((long long)N * T, 8), 256, 0, s>>>((const __nv_bfloat16*)table, tok, (__nv_bfloat16*)X,
                                                                keep_mask, inv_keep, N, T, E, V);
  }
  VC_CUDA(cudaGetLastError());
  return VC_OK;
}

// Embedding backward: per token row g = dX[t, n, :] (* keep_mask / keep); normsq += |g|^2 (the
// IndexedSlices.values norm TF's clip_by_global_norm uses, Q4); grad_table[tok] += g.
__global__ void k_embed_scatter(const float* __restrict__ dX, const int* __restrict__ tok, float* __restrict__ gtable,
                                const float* __restrict__ keep_mask, float inv_keep, float* __restrict__ normsq, int N,
                                int T, int E, int V) {
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  float acc = 0.f;
  for (long long row = blockIdx.x * (long long)warps_per_block + (threadIdx.x >> 5); row < (long long)N * T;
       row += (long long)gridDim.x * warps_per_block) {
    const int t = (int)(row / N), n = (int)(row - (long long)t * N);
    int id = tok[(long long)n * T + t];
    id = id < 0 ? 0 : (id >= V ? V - 1 : id);
    const float* src = dX + row * E;
    const float* mk = keep_mask ? keep_mask + ((long long)n * T + t) * E : nullptr;
    for (int e = lane; e < E; e += 32) {
      float g = src[e];
      if (mk) g *= mk[e] * inv_keep;
      acc += g * g;
      if (g != 0.f) atomicAdd(gtable + (long long)id * E + e, g);
    }
  }
  acc = warp_sum(acc);
  if (lane == 0 && acc != 0.f) atomicAdd(normsq, acc);
}
int embed_scatter(cudaStream_t s, const float* dX, const int* tok, float* gtable, const float* keep_mask, float inv_keep,
                  float* normsq, int N, int T, int E, int V) {
  {
    ProfScope ps(s, "embed_scatter");
    k_embed_scatter<<<grid_for((long long)N * T, 8), 256, 0, s>>>(dX, tok, gtable, keep_mask, inv_keep, normsq, N, T, E, V);
  }
  VC_CUDA(cudaGetLastError());
  return VC_OK;
}

// ------------------------------------------------------------------------------------------
// Reparameterised sample + KL (encoder.py:108-109, main.py:118-145).
// heads [N, ld] fp32: mu at [0, Z), log-std at [zp, zp+Z) (Normal prior) -> std = exp(logstd).
// For GMM/AG the caller provides mu/std directly (mix kernel) and passes heads == nullptr.
// z[s, n, k] = mu + std * eps (bf16, consumed by the z_rnn GEMM through the Q1 row-major view).
// eps: explicit fp32 [S,N,Z] (parity) or Philox(seed, offset) when eps == nullptr.
__device__ __forceinline__ float4 philox_normal4(unsigned long long seed, unsigned long long offset, long long quad) {
  curandStatePhilox4_32_10_t st;
  curand_init(seed, (unsigned long long)quad, offset, &st);
  return curand_normal4(&st);
}

// Posterior parameters from the packed head outputs (encoder.py:59-107). heads [N, ld] fp32: head k has its mean at
// columns [2k*zp, 2k*zp+Z) and its log-std at [(2k+1)*zp, ...).
//   Normal: mu = mean_0, sd = exp(logstd_0)
//   GMM   : k = pick[n] (tf.multinomial draw, encoder.py:72-88): mu = mean_k, sd = exp(logstd_k)
//   AG    : mu = sum_k c_v[n,k] mean_k, sd = sum_k c_v[n,k] exp(logstd_k) (encoder.py:105-107); zero weights are
//           skipped (exactly equivalent, Q17). Also cm[n,:] = c_v[n,:] @ c_means for the AG KL (main.py:142-143).
__global__ void k_heads_mix(const float* __restrict__ heads, long long ld, int zp, int prior, const float* __restrict__ c_v,
                            int K, const int* __restrict__ pick, const float* __restrict__ c_means, float* __restrict__ mu,
                            float* __restrict__ sd, float* __restrict__ cm, int N, int Z) {
  const long long total = (long long)N * Z;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / Z;
    const int z = (int)(i - n * Z);
    const float* hr = heads + n * ld;
    if (prior == 2) {
      float m = 0.f, s = 0.f, c = 0.f;
      const float* w = c_v + n * K;
      for (int k = 0; k < K; ++k) {
        const float wk = w[k];
        if (wk != 0.f) {
          m = fmaf(wk, hr[(2 * k) * zp + z], m);
          s = fmaf(wk, __expf(hr[(2 * k + 1) * zp + z]), s);
          c = fmaf(wk, c_means[(long long)k * Z + z], c);
        }
      }
      mu[i] = m; sd[i] = s; cm[i] = c;
    } else {
      int k = 0;
      if (prior == 1) { k = pick[n]; k = k < 0 ? 0 : (k >= K ? K - 1 : k); }
      mu[i] = hr[(2 * k) * zp + z];
      sd[i] = __expf(hr[(2 * k + 1) * zp + z]);
    }
  }
}
int heads_mix(cudaStream_t s, const float* heads, long long ld, int zp, int prior, const float* c_v, int K, const int* pick,
              const float* c_means, float* mu, float* sd, float* cm, int N, int Z) {
  {
    ProfScope ps(s, "heads_mix");
    k_heads_mix<<<grid_for((long long)N * Z, 256), 256, 0, s>>>(heads, ld, zp, prior, c_v, K, pick, c_means, mu, sd, cm, N, Z);
  }
  VC_CUDA(cudaGetLastError());
  return VC_OK;
}

// Backward of heads_mix: dmu/dsd fp32 [N, Z] -> bf16 gradient of the packed head outputs (pre-zeroed for GMM/AG).
__global__ void k_heads_mix_bwd(const float* __restrict__ dmu, const float* __restrict__ dsd, const float* __restrict__ heads,
                                long long ld, int zp, int prior, const float* __restrict__ c_v, int K,
                                const int* __restrict__ pick, const float* __restrict__ sd,
                                __nv_bfloat16* __restrict__ dheads, int N, int Z) {
  const long long total = (long long)N * Z;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / Z;
    const int z = (int)(i - n * Z);
    __nv_bfloat16* dr = dheads + n * ld;
    const float gm = dmu[i], gs = dsd[i];
    if (prior == 2) {
      const float* hr = heads + n * ld;
      const float* w = c_v + n * K;
      for (int k = 0; k < K; ++k) {
        const float wk = w[k];
        if (wk != 0.f) {
          dr[(2 * k) * zp + z] = __float2bfloat16(wk * gm);
          dr[(2 * k + 1) * zp + z] = __float2bfloat16(wk * gs * __expf(hr[(2 * k + 1) * zp + z]));
        }
      }
    } else {
      int k = 0;
      if (prior == 1) { k = pick[n]; k = k < 0 ? 0 : (k >= K ? K - 1 : k); }
      dr[(2 * k) * zp + z] = __float2bfloat16(gm);
      dr[(2 * k + 1) * zp + z] = __float2bfloat16(gs * sd[i]);  // d/dlogstd = d/dsd * sd
    }
  }
}
int heads_mix_bwd(cudaStream_t s, const float* dmu, const float* dsd, const float* heads, long long ld, int zp, int prior,
                  const float* c_v, int K, const int* pick, const float* sd, void* dheads, int N, int Z) {
  {
    ProfScope ps(s, "heads_mix_bwd");
    k_heads_mix_bwd<<<grid_for((long long)N * Z, 256), 256, 0, s>>>(dmu, dsd, heads, ld, zp, prior, c_v, K, pick, sd,
                                                                   (__nv_bfloat16*)dheads, N, Z);
  }
  VC_CUDA(cudaGetLastError());
  return VC_OK;
}

// tf.multinomial(logits = c_v, 1) (encoder.py:72, Q3: the cluster *probabilities* are used as logits): one draw per
// row from softmax(c_v[n, :]) by inverse CDF on a Philox uniform.
__global__ void k_gmm_pick(const float* __restrict__ c_v, int K, unsigned long long seed, unsigned long long offset,
                           int* __restrict__ pick, int N) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  curandStatePhilox4_32_10_t st;
  curand_init(seed ^ 0x9E3779B97F4A7C15ull, (unsigned long long)n, offset, &st);
  const float u = curand_uniform(&st);  // (0, 1]
  const float* w = c_v + (long long)n * K;
  float mx = -INFINITY, tot = 0.f;
  for (int k = 0; k < K; ++k) mx = fmaxf(mx, w[k]);
  for (int k = 0; k < K; ++k) tot += __expf(w[k] - mx);
  float acc = 0.f;
  int sel = K - 1;
  for (int k = 0; k < K; ++k) {
    acc += __expf(w[k] - mx);
    if (u * tot <= acc) { sel = k; break; }
  }
  pick[n] = sel;
}
int gmm_pick_clusters(cudaStream_t s, const float* c_v, int K, unsigned long long seed, unsigned long long offset, int* pick,
                      int N) {
  {
    ProfScope ps(s, "gmm_pick");
    k_gmm_pick<<<(N + 127) / 128, 128, 0, s>>>(c_v, K, seed, offset, pick, N);
  }
  VC_CUDA(cudaGetLastError());
  return VC_OK;
}

// KL rows. prior 0/1 (Normal/GMM): kl_row[n] = -0.5 * sum_k (1 + log(sd^2+1e-5) - mu^2 - sd^2)
// prior 2 (AG): kl_row[n] = -0.5 * sum_k (0.5 + log(sd+1e-5) - log(cs+1e-5) - ((mu-cm)^2 + sd^2)/(2cs^2+1e-7)),
// cm = c_v[n,:] @ c_means. Also writes dKL/dmu, dKL/dsd (per-row, unscaled) for the backward pass.
__global__ void k_kl_rows(const float* __restrict__ mu, const float* __restrict__ sd, const float* __restrict__ cm,
                          int prior, float* __restrict__ kl_row, float* __restrict__ dkl_dmu,
                          float* __restrict__ dkl_dsd, float* __restrict__ kl_sum, int N, int Z) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= N) return;
  const float cs = 0.1f;
  float acc = 0.f;
  for (int k = lane; k < Z; k += 32) {
    const long long i = (long long)warp * Z + k;
    const float m = mu[i], s = sd[i];
    if (prior != 2) {
      acc += 1.f + logf(s * s + 0.00001f) - m * m - s * s;
      dkl_dmu[i] = m;                                          // -0.5 * (-2 mu)
      dkl_dsd[i] = -0.5f * (2.f * s / (s * s + 0.00001f) - 2.f * s);
    } else {
      const float d = m - cm[i];
      const float den = 2.f * cs * cs + 0.0000001f;
      acc += 0.5f + logf(s + 0.00001f) - logf(cs + 0.00001f) - (d * d + s * s) / den;
      dkl_dmu[i] = -0.5f * (-2.f * d / den);
      dkl_dsd[i] = -0.5f * (1.f / (s + 0.00001f) - 2.f * s / den);
    }
  }
  acc = warp_sum(acc);
  if (lane == 0) {
    kl_row[warp] = -0.5f * acc;
    atomicAdd(kl_sum, -0.5f * acc);
  }
}
int kl_rows(cudaStream_t s, const float* mu, const float* sd, const float* cm, int prior, float* kl_row, float* dkl_dmu,
            float* dkl_dsd, float* kl_sum, int N, int Z) {
  {
    ProfScope ps(s, "kl_rows");
    k_kl_rows<<<(N * 32 + 255) / 256, 256, 0, s>>>(mu, sd, cm, prior, kl_row, dkl_dmu, dkl_dsd, kl_sum, N, Z);
  }
  VC_CUDA(cudaGetLastError());
  return VC_OK;
}

__global__ void k_sample_z(const float* __restrict__ mu, const float* __restrict__ sd, const float* __restrict__ eps,
                           unsigned long long seed, unsigned long long offset, __nv_bfloat16* __restrict__ z,
                           float* __restrict__ z_f32, int S, long long NZ) {
  // one thread per 4 consecutive elements of [S, N*Z]
  const long long quads = ((long long)S * NZ + 3) / 4;
  for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < quads; q += (long long)gridDim.x * blockDim.x) {
    float e[4];
    const long long i0 = q * 4;
    if (eps == nullptr) {
      const float4 r = philox_normal4(seed, offset, q);
      e[0] = r.x; e[1] = r.y; e[2] = r.z; e[3] = r.w;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const long long i = i0 + j;
      if (i >= (long long)S * NZ) break;
      const long long r = i % NZ;
      const float ee = eps ? eps[i] : e[j];
      const float v = mu[r] + sd[r] * ee;
      z[i] = __float2bfloat16(v);
      if (z_f32) z_f32[i] = v;
    }
  }
}
// Same for N*Z % 4 == 0 and 16-byte aligned operands: grid.y = s, one thread per quad of [N*Z] (the Philox quad index
// s * NZ/4 + q is the one k_sample_z and k_dz_reduce use), float4 loads of mu / sd / eps, one 8-byte store of 4 bf16.
__global__ void k_sample_z_v4(const float4* __restrict__ mu, const float4* __restrict__ sd, const float4* __restrict__ eps,
                              unsigned long long seed, unsigned long long offset, uint2* __restrict__ z,
                              float4* __restrict__ z_f32, long long NZ4) {
  const long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (q >= NZ4) return;
  const long long quad = (long long)blockIdx.y * NZ4 + q;
  const float4 e = eps != nullptr ? eps[quad] : philox_normal4(seed, offset, quad);
  const float4 m = __ldg(mu + q), d = __ldg(sd + q);
  const float4 v = make_float4(fmaf(d.x, e.x, m.x), fmaf(d.y, e.y, m.y), fmaf(d.z, e.z, m.z), fmaf(d.w, e.w, m.w));
  z[quad] = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
  if (z_f32 != nullptr) z_f32[quad] = v;
}
static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
int sample_z(cudaStream_t s, const float* mu, const float* sd, const float* eps, unsigned long long seed,
             unsigned long long offset, void* z, float* z_f32, int S, long long NZ) {
  {
    ProfScope ps(s, "sample_z");
    if (NZ % 4 == 0 && S <= 65535 && aligned16(mu) && aligned16(sd) && aligned16(eps) && aligned16(z) && aligned16(z_f32)) {
      const long long nz4 = NZ / 4;
      dim3 grid((unsigned)((nz4 + 255) / 256), (unsigned)S);
      k_sample_z_v4<<<grid, 256, 0, s>>>((const float4*)mu, (const float4*)sd, (const float4*)eps, seed, offset, (uint2*)z,
                                         (float4*)z_f32, nz4);
    } else {
      k_sample_z<<<grid_for(((long long)S * NZ + 3) / 4, 256), 256, 0, s>>>(mu, sd, eps, seed, offset, (__nv_bfloat16*)z,
                                                                            z_f32, S, NZ);
    }
  }
  VC_CUDA(cudaGetLastError());
  return VC_OK;
}

// dmu[r] = sum_s dz[s, r]; dsd[r] = sum_s dz[s, r] * eps[s, r]   (r over N*Z; dz fp32 [S, N*Z]),
// then adds kl_scale * dKL and emits the head gradient in bf16:
//   Normal: dheads[n, k] = dmu, dheads[n, zp + k] = dsd * sd (d/dlogstd).  Otherwise writes dmu/dsd fp32.
__global__ void k_dz_reduce(const float* __restrict__ dz, const float* __restrict__ eps, unsigned long long seed,
                            unsigned long long offset, const float* __restrict__ sd, const float* __restrict__ dkl_dmu,
                            const float* __restrict__ dkl_dsd, float kl_scale, __nv_bfloat16* __restrict__ dheads,
                            long long ld, int zp, float* __restrict__ dmu_out, float* __restrict__ dsd_out, int S, int N,
                            int Z) {
  const long long NZ = (long long)N * Z;
  // each thread owns 4 consecutive r (a Philox quad never straddles rows of s because NZ % 4 == 0 is required
  // for the Philox path; the explicit-eps path has no such constraint)
  const long long quads = (NZ + 3) / 4;
  for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < quads; q += (long long)gridDim.x * blockDim.x) {
    float am[4] = {0, 0, 0, 0}, as[4] = {0, 0, 0, 0};
    for (int s = 0; s < S; ++s) {
      const long long base = (long long)s * NZ + q * 4;
      float e[4];
      if (eps == nullptr) {
        const float4 r = philox_normal4(seed, offset, base / 4);
        e[0] = r.x; e[1] = r.y; e[2] = r.z; e[3] = r.w;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (q * 4 + j >= NZ) break;
        const float g = dz[base + j];
        am[j] += g;
        as[j] += g * (eps ? eps[base + j] : e[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const long long r = q * 4 + j;
      if (r >= NZ) break;
      const float gm = am[j] + kl_scale * dkl_dmu[r];
      const float gs = as[j] + kl_scale * dkl_dsd[r];
      if (dheads) {
        const long long n = r / Z;
        const int k = (int)(r - n * Z);
        dheads[n * ld + k] = __float2bfloat16(gm);
        dheads[n * ld + zp + k] = __float2bfloat16(gs * sd[r]);
      }
      if (dmu_out) dmu_out[r] = gm;
      if (dsd_out) dsd_out[r] = gs;
    }
  }
}
// Same for N*Z % 4 == 0 and 16-byte aligned dz / eps: the sum over the S samples is split over the four warps of a
// CTA (warp w takes s = w, w+4, ...; 32 lanes = 32 consecutive quads, 512 contiguous bytes per load instruction) and
// folded through shared memory in a fixed order, so the result does not depend on scheduling.
__global__ void __launch_bounds__(128)
k_dz_reduce_v4(const float4* __restrict__ dz, const float4* __restrict__ eps, unsigned long long seed,
               unsigned long long offset, const float* __restrict__ sd, const float* __restrict__ dkl_dmu,
               const float* __restrict__ dkl_dsd, float kl_scale, __nv_bfloat16* __restrict__ dheads, long long ld, int zp,
               float* __restrict__ dmu_out, float* __restrict__ dsd_out, int S, long long NZ4, int Z) {
  __shared__ float4 part[2][3][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const long long q = blockIdx.x * 32ll + lane;
  float4 am = make_float4(0.f, 0.f, 0.f, 0.f), as = am;
  if (q < NZ4) {
#pragma unroll 2
    for (int s = w; s < S; s += 4) {
      const long long quad = (long long)s * NZ4 + q;
      const float4 g = dz[quad];
      const float4 e = eps != nullptr ? eps[quad] : philox_normal4(seed, offset, quad);
      am.x += g.x; am.y += g.y; am.z += g.z; am.w += g.w;
      as.x = fmaf(g.x, e.x, as.x); as.y = fmaf(g.y, e.y, as.y); as.z = fmaf(g.z, e.z, as.z); as.w = fmaf(g.w, e.w, as.w);
    }
  }
  if (w > 0) {
    part[0][w - 1][lane] = am;
    part[1][w - 1][lane] = as;
  }
  __syncthreads();
  if (w != 0 || q >= NZ4) return;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float4 a = part[0][k][lane], b = part[1][k][lane];
    am.x += a.x; am.y += a.y; am.z += a.z; am.w += a.w;
    as.x += b.x; as.y += b.y; as.z += b.z; as.w += b.w;
  }
  const float m4[4] = {am.x, am.y, am.z, am.w}, s4[4] = {as.x, as.y, as.z, as.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const long long r = q * 4 + j;
    const float gm = m4[j] + kl_scale * dkl_dmu[r];
    const float gs = s4[j] + kl_scale * dkl_dsd[r];
    if (dheads) {
      const long long n = r / Z;
      const int k = (int)(r - n * Z);
      dheads[n * ld + k] = __float2bfloat16(gm);
      dheads[n * ld + zp + k] = __float2bfloat16(gs * sd[r]);
    }
    if (dmu_out) dmu_out[r] = gm;
    if (dsd_out) dsd_out[r] = gs;
  }
}
int dz_reduce(cudaStream_t s, const float* dz, const float* eps, unsigned long long seed, unsigned long long offset,
              const float* sd, const float* dkl_dmu, const float* dkl_dsd, float kl_scale, void* dheads, long long ld,
              int zp, float* dmu_out, float* dsd_out, int S, int N, int Z) {
  if (eps == nullptr && ((long long)N * Z) % 4 != 0)
    return set_error(VC_E_SHAPE, "Philox sampling needs N*Z to be a multiple of 4");
  {
    ProfScope ps(s, "dz_reduce");
    const long long NZ = (long long)N * Z;
    if (NZ % 4 == 0 && aligned16(dz) && aligned16(eps)) {
      const long long nz4 = NZ / 4;
      k_dz_reduce_v4<<<(unsigned)((nz4 + 31) / 32), 128, 0, s>>>((const float4*)dz, (const float4*)eps, seed, offset, sd,
                                                                 dkl_dmu, dkl_dsd, kl_scale, (__nv_bfloat16*)dheads, ld, zp,
                                                                 dmu_out, dsd_out, S, nz4, Z);
    } else {
      k_dz_reduce<<<grid_for((NZ + 3) / 4, 128), 128, 0, s>>>(dz, eps, seed, offset, sd, dkl_dmu, dkl_dsd, kl_scale,
                                                              (__nv_bfloat16*)dheads, ld, zp, dmu_out, dsd_out, S, N, Z);
    }
  }
  VC_CUDA(cudaGetLastError());
  return VC_OK;
}

// ------------------------------------------------------------------------------------------
// Vocab softmax cross-entropy over bf16 logits rows [rows, ld] (row = t*N + n, time-major), labels
// lbl[n*T + t]. One CTA per row; the row lives in registers between the two passes, so logits are
// read once and (optionally) overwritten in place by dlogits = (softmax - onehot) * mask * gscale.
// Rows are 16-byte aligned (ld % 8 == 0): every thread moves its share as 16-byte vectors, all of
// them in flight before the first use. Accumulates sums[0] += ce*mask, sums[1] += mask.
// gscale = loss_scale / count[0] (count = sum mask).
constexpr int kCeThreads = 256;
constexpr int kCeVecs = 6;  // 8-element vectors per thread: V <= 256 * 6 * 8 = 12288

__global__ void __launch_bounds__(kCeThreads)
k_ce(__nv_bfloat16* __restrict__ logits, long long ld, const int* __restrict__ lbl, int N, int T, int V,
     float* __restrict__ sums, float* __restrict__ ce_rows, const float* __restrict__ count, float loss_scale,
     int write_grad, int row0) {
  const int row = blockIdx.x + row0;
  const int t = row / N, n = row - t * N;
  const int label = lbl[(long long)n * T + t];
  uint4* p = reinterpret_cast<uint4*>(logits + (long long)row * ld);
  const int vecs = (V + 7) / 8;
  uint4 raw[kCeVecs];
#pragma unroll
  for (int c = 0; c < kCeVecs; ++c) {
    const int i = threadIdx.x + c * kCeThreads;
    raw[c] = make_uint4(0xff80ff80u, 0xff80ff80u, 0xff80ff80u, 0xff80ff80u);  // bf16 -inf pairs
    if (i < vecs) raw[c] = p[i];
  }
  float v[kCeVecs * 8];
  float mx = -INFINITY;
#pragma unroll
  for (int c = 0; c < kCeVecs; ++c) {
    const int i = threadIdx.x + c * kCeThreads;
    const uint32_t w[4] = {raw[c].x, raw[c].y, raw[c].z, raw[c].w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {  // bf16 -> fp32 is a 16-bit shift
      v[c * 8 + 2 * k] = __uint_as_float(w[k] << 16);
      v[c * 8 + 2 * k + 1] = __uint_as_float(w[k] & 0xffff0000u);
    }
    if (i == vecs - 1) {  // the row's last vector may run into the pitch padding
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (8 * i + k >= V) v[c * 8 + k] = -INFINITY;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) mx = fmaxf(mx, v[c * 8 + k]);
  }
  __shared__ float red[kCeThreads / 32];
  __shared__ float bcast;
  mx = warp_max(mx);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x < 32) {
    float m2 = threadIdx.x < kCeThreads / 32 ? red[threadIdx.x] : -INFINITY;
    m2 = warp_max(m2);
    if (threadIdx.x == 0) bcast = m2;
  }
  __syncthreads();
  mx = bcast;
  constexpr float kLog2e = 1.4426950408889634f;
  const float mxl = -mx * kLog2e;
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < kCeVecs * 8; ++j) {
    v[j] = exp2f(fmaf(v[j], kLog2e, mxl));  // one FFMA + EX2; the -inf padding becomes 0
    sum += v[j];
  }
  __syncthreads();
  sum = warp_sum(sum);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
  __syncthreads();
  if (threadIdx.x < 32) {
    float s2 = threadIdx.x < kCeThreads / 32 ? red[threadIdx.x] : 0.f;
    s2 = warp_sum(s2);
    if (threadIdx.x == 0) bcast = s2;
  }
  __syncthreads();
  sum = bcast;
  const float mask = label != 0 ? 1.f : 0.f;  // tf.sign(tf.to_float(labels)), labels >= 0
  float ll = 0.f;
  if (threadIdx.x == 0) {
    const int lc = label < 0 ? 0 : (label >= V ? V - 1 : label);
    ll = __bfloat162float(logits[(long long)row * ld + lc]);
    const float ce = logf(sum) + mx - ll;
    if (ce_rows) ce_rows[(long long)n * T + t] = ce;
    if (mask != 0.f) {
      atomicAdd(&sums[0], ce);
      atomicAdd(&sums[1], 1.f);
    }
  }
  if (write_grad) {
    __syncthreads();  // the label logit was read above before anyone overwrites it
    const float gm = mask * loss_scale / fmaxf(count[0], 1.f);
    const float g = gm / sum;
    // softmax * g for every column (padding columns hold exp(-inf) = 0); the one-hot term is patched into the label
    // column by one thread afterwards instead of a compare + select per element
#pragma unroll
    for (int c = 0; c < kCeVecs; ++c) {
      const int i = threadIdx.x + c * kCeThreads;
      if (i < vecs) {
        float a[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] = v[c * 8 + k] * g;
        p[i] = make_uint4(pack_bf16(a[0], a[1]), pack_bf16(a[2], a[3]), pack_bf16(a[4], a[5]), pack_bf16(a[6], a[7]));
      }
    }
    if (gm != 0.f && label >= 0 && label < V) {  // block-uniform
      __syncthreads();
      if (threadIdx.x == 0)
        logits[(long long)row * ld + label] = __float2bfloat16(exp2f(fmaf(ll, kLog2e, mxl)) * g - gm);
    }
  }
}
// Vocabularies wider than the register-resident form above holds (V > 12288): same result, the row is read from
// global memory three times (max, sum of exponentials, gradient) instead of living in registers.
__global__ void __launch_bounds__(kCeThreads)
k_ce_wide(__nv_bfloat16* __restrict__ logits, long long ld, const int* __restrict__ lbl, int N, int T, int V,
          float* __restrict__ sums, float* __restrict__ ce_rows, const float* __restrict__ count, float loss_scale,
          int write_grad, int row0) {
  const int row = blockIdx.x + row0;
  const int t = row / N, n = row - t * N;
  const int label = lbl[(long long)n * T + t];
  __nv_bfloat16* p = logits + (long long)row * ld;
  __shared__ float red[kCeThreads / 32];
  __shared__ float bcast;
  float mx = -INFINITY;
  for (int c = threadIdx.x; c < V; c += kCeThreads) mx = fmaxf(mx, __bfloat162float(p[c]));
  mx = warp_max(mx);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x < 32) {
    float m2 = threadIdx.x < kCeThreads / 32 ? red[threadIdx.x] : -INFINITY;
    m2 = warp_max(m2);
    if (threadIdx.x == 0) bcast = m2;
  }
  __syncthreads();
  mx = bcast;
  float sum = 0.f;
  for (int c = threadIdx.x; c < V; c += kCeThreads) sum += __expf(__bfloat162float(p[c]) - mx);
  __syncthreads();
  sum = warp_sum(sum);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
  __syncthreads();
  if (threadIdx.x < 32) {
    float s2 = threadIdx.x < kCeThreads / 32 ? red[threadIdx.x] : 0.f;
    s2 = warp_sum(s2);
    if (threadIdx.x == 0) bcast = s2;
  }
  __syncthreads();
  sum = bcast;
  const float mask = label != 0 ? 1.f : 0.f;
  if (threadIdx.x == 0) {
    const int lc = label < 0 ? 0 : (label >= V ? V - 1 : label);
    const float ce = logf(sum) + mx - __bfloat162float(p[lc]);
    if (ce_rows) ce_rows[(long long)n * T + t] = ce;
    if (mask != 0.f) {
      atomicAdd(&sums[0], ce);
      atomicAdd(&sums[1], 1.f);
    }
  }
  if (write_grad) {
    __syncthreads();
    const float gm = mask * loss_scale / fmaxf(count[0], 1.f);
    const float g = gm / sum;
    for (int c = threadIdx.x; c < (int)ld; c += kCeThreads) {
      float a = 0.f;
      if (c < V) a = __expf(__bfloat162float(p[c]) - mx) * g - (c == label ? gm : 0.f);
      p[c] = __float2bfloat16(a);
    }
  }
}

int ce_rows(cudaStream_t s, void* logits, long long ld, const int* lbl, int N, int T, int V, float* sums, float* ce_out,
            const float* count, float loss_scale, int write_grad, int row0, int n_rows) {
  if (n_rows < 0) n_rows = N * T - row0;
  if (ld % 8 != 0 || (reinterpret_cast<uintptr_t>(logits) & 15) != 0)
    return set_error(VC_E_SHAPE, "logits rows must be 16-byte aligned (pitch multiple of 8)");
  {
    ProfScope ps(s, "ce");
    if (V > kCeThreads * kCeVecs * 8)
      k_ce_wide<<<n_rows, kCeThreads, 0, s>>>((__nv_bfloat16*)logits, ld, lbl, N, T, V, sums, ce_out, count, loss_scale, write_grad, row0);
    else
      k_ce<<<n_rows, kCeThreads, 0, s>>>((__nv_bfloat16*)logits, ld, lbl, N, T, V, sums, ce_out, count, loss_scale, write_grad, row0);
  }
  VC_CUDA(cudaGetLastError());
  return VC_OK;
}

// count[0] = number of labels != 0 (sum of the loss mask)
__global__ void k_count_mask(const int* __restrict__ lbl, long long n, float* __restrict__ count) {
  float acc = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    acc += lbl[i] != 0 ? 1.f : 0.f;
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0 && acc != 0.f) atomicAdd(count, acc);
}
int count_mask(cudaStream_t s, const int* lbl, long long n, float* count) {
  {
    ProfScope ps(s, "count_mask");
    k_count_mask<<<grid_for(n, 256, 4), 256, 0, s>>>(lbl, n, count);
  }
  VC_CUDA(cudaGetLastError());
  return VC_OK;
}

// ------------------------------------------------------------------------------------------
// Column sums of a bf16 matrix [rows, ld] -> out[col] += sum (fp32 atomics, one per column per CTA).
__global__ void k_colsum_bf16(const __nv_bfloat16* __restrict__ x, long long rows, int cols, long long ld,
                              float* __restrict__ out, int rows_per_cta) {
  const int c2 = (blockIdx.x * blockDim.x + threadIdx.x) * 2;
  if (c2 >= cols) return;
  const long long r0 = (long long)blockIdx.y * rows_per_cta;
  const long long r1 = min(rows, r0 + rows_per_cta);
  float a = 0.f, b = 0.f;
  for (long long r = r0; r < r1; ++r) {
    const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(x + r * ld + c2));
    a += f.x;
    b += f.y;
  }
  atomicAdd(out + c2, a);
  if (c2 + 1 < cols) atomicAdd(out + c2 + 1, b);
}
int colsum_bf16(cudaStream_t s, const void* x, long long rows, int cols, long long ld, float* out) {
  if (ld % 2 != 0) return set_error(VC_E_SHAPE, "colsum pitch must be even");
  const int threads = 128;
  const int gx = ((cols + 1) / 2 + threads - 1) / threads;
  int gy = (num_sms() * 8 + gx - 1) / gx;
  if (gy > rows) gy = (int)rows;
  if (gy < 1) gy = 1;
  const int rpc = (int)((rows + gy - 1) / gy);
  dim3 grid(gx, (unsigned)((rows + rpc - 1) / rpc));
  {
    ProfScope ps(s, "colsum_bf16");
    k_colsum_bf16<<<grid, threads, 0, s>>>((const __nv_bfloat16*)x, rows, cols, ld, out, rpc);
  }
  VC_CUDA(cudaGetLastError());
  return VC_OK;
}

// ------------------------------------------------------------------------------------------
// Global norm (sum of squares into *out) and fused clip + TF-form Adam (Q5):
//   g' = g * clip / max(sqrt(normsq), clip);  m = b1 m + (1-b1) g';  v = b2 v + (1-b2) g'^2;
//   p -= lr_t * m / (sqrt(v) + eps),  lr_t = lr sqrt(1-b2^t)/(1-b1^t) (computed by the host).
// Deterministic: block partials land in `partial`, the block that finishes last adds them up in index order. (Partial sums
// meeting in atomicAdd made the global norm differ in its last bits from run to run -- and between data-parallel replicas,
// whose clip scale and therefore whose parameters then drifted apart by ulps.)
__global__ void k_sumsq(const float* __restrict__ g, long long n, float* __restrict__ out, float* __restrict__ partial,
                        unsigned int* __restrict__ counter) {
  float acc = 0.f;
  const long long n4 = n / 4;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = g4[i];
    acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  for (long long i = n4 * 4 + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    acc += g[i] * g[i];
  acc = warp_sum(acc);
  __shared__ float red[32];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x < 32) {
    float a = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    a = warp_sum(a);
    if (threadIdx.x == 0) {
      partial[blockIdx.x] = a;
      __threadfence();
      last = atomicInc(counter, gridDim.x - 1) == gridDim.x - 1;  // wraps to 0: ready for the next launch
    }
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  float a = 0.f;
  for (unsigned int i = threadIdx.x; i < gridDim.x; i += blockDim.x) a += __ldcg(partial + i);
  a = warp_sum(a);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0) *out += t;  // the caller zeroes `out`; one writer
  }
}
int sumsq(cudaStream_t s, const float* g, long long n, float* out) {
  if (n <= 0) return VC_OK;
  static float* partial = nullptr;  // one GPU per process (include/vaecap.h): scratch shared by every handle's launches
  static unsigned int* counter = nullptr;
  constexpr int kMaxBlocks = 8192;
  if (partial == nullptr) {
    VC_CUDA(cudaMalloc((void**)&partial, kMaxBlocks * sizeof(float)));
    VC_CUDA(cudaMalloc((void**)&counter, sizeof(unsigned int)));
    VC_CUDA(cudaMemset(counter, 0, sizeof(unsigned int)));
  }
  int grid = grid_for(n, 256, 8);
  if (grid > kMaxBlocks) grid = kMaxBlocks;
  {
    ProfScope ps(s, "sumsq");
    k_sumsq<<<grid, 256, 0, s>>>(g, n, out, partial, counter);
  }
  VC_CUDA(cudaGetLastError());
  return VC_OK;
}

// normsq_parts: the squared norm is the sum of up to 4 device scalars (dense grads + per-token embedding slices).
__global__ void k_adam(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                       long long n, const float* __restrict__ normsq_parts, int n_parts, float clip, float gscale,
                       float lr_t, float b1, float b2, float eps, float* __restrict__ norm_out, float wd, int kind) {
  float scale = gscale;
  if (clip > 0.f) {
    float ns = 0.f;
    for (int i = 0; i < n_parts; ++i) ns += normsq_parts[i];
    const float norm = sqrtf(ns) * gscale;
    scale = gscale * clip / fmaxf(norm, clip);
    if (norm_out && blockIdx.x == 0 && threadIdx.x == 0) *norm_out = norm;
  }
  const long long n4 = n / 4;
  float4* p4 = reinterpret_cast<float4*>(p);
  const float4* g4 = reinterpret_cast<const float4*>(g);
  float4* m4 = reinterpret_cast<float4*>(m);
  float4* v4 = reinterpret_cast<float4*>(v);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    if (kind != 0) {
      // tf.train.GradientDescentOptimizer / MomentumOptimizer(momentum = b1) (ops/optimizers.py:33-36, 41-46): lr_t is the
      // staircase-decayed rate; the moment buffer m holds the momentum accumulator, v is not touched
      float4 pp = p4[i], gg = g4[i];
      float4 st = make_float4(gg.x * scale + wd * pp.x, gg.y * scale + wd * pp.y, gg.z * scale + wd * pp.z, gg.w * scale + wd * pp.w);
      if (kind == 2) {
        const float4 mm = m4[i];
        st = make_float4(b1 * mm.x + st.x, b1 * mm.y + st.y, b1 * mm.z + st.z, b1 * mm.w + st.w);
        m4[i] = st;
      }
      pp.x -= lr_t * st.x; pp.y -= lr_t * st.y; pp.z -= lr_t * st.z; pp.w -= lr_t * st.w;
      p4[i] = pp;
      continue;
    }
    float4 pp = p4[i], gg = g4[i], mm = m4[i], vv = v4[i];
#define VC_ADAM1(c)                                 \
  {                                                 \
    const float gs = gg.c * scale + wd * pp.c;      \
    mm.c = b1 * mm.c + (1.f - b1) * gs;             \
    vv.c = b2 * vv.c + (1.f - b2) * gs * gs;        \
    pp.c -= lr_t * mm.c / (sqrtf(vv.c) + eps);      \
  }
    VC_ADAM1(x) VC_ADAM1(y) VC_ADAM1(z) VC_ADAM1(w)
#undef VC_ADAM1
    p4[i] = pp;
    m4[i] = mm;
    v4[i] = vv;
  }
}
int adam_step(cudaStream_t s, float* p, const float* g, float* m, float* v, long long n, const float* normsq_parts,
              int n_parts, float clip, float gscale, float lr_t, float b1, float b2, float eps, float* norm_out, float weight_decay,
              int kind) {
  if (n <= 0) return VC_OK;
  if (n % 4 != 0) return set_error(VC_E_ARG, "adam_step: length must be a multiple of 4 (padded flat buffer)");
  {
    ProfScope ps(s, "adam");
    k_adam<<<grid_for(n / 4, 256, 2), 256, 0, s>>>(p, g, m, v, n, normsq_parts, n_parts, clip, gscale, lr_t, b1, b2, eps, norm_out, weight_decay, kind);
  }
  VC_CUDA(cudaGetLastError());
  return VC_OK;
}

// ------------------------------------------------------------------------------------------
// bf16 [rows, ld] (time-major rows t*N+n) -> fp32 [N*T, V] in the reference's row order n*T+t (debug tap).
__global__ void k_logits_to_ref(const __nv_bfloat16* __restrict__ src, long long ld, float* __restrict__ dst, int N, int T,
                                int V) {
  const long long total = (long long)N * T * V;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % V);
    const long long r = i / V;  // n*T + t
    const int n = (int)(r / T), t = (int)(r - (long long)n * T);
    dst[i] = __bfloat162float(src[((long long)t * N + n) * ld + c]);
  }
}
int logits_to_ref(cudaStream_t s, const void* src, long long ld, float* dst, int N, int T, int V) {
  {
    ProfScope ps(s, "logits_to_ref");
    k_logits_to_ref<<<grid_for((long long)N * T * V, 256, 4), 256, 0, s>>>((const __nv_bfloat16*)src, ld, dst, N, T, V);
  }
  VC_CUDA(cudaGetLastError());
  return VC_OK;
}

__global__ void k_bf16_to_f32(const __nv_bfloat16* __restrict__ src, float* __restrict__ dst, long long rows, int cols,
                              long long ld_src, long long ld_dst) {
  const long long total = rows * cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols;
    const int c = (int)(i - r * cols);
    dst[r * ld_dst + c] = __bfloat162float(src[r * ld_src + c]);
  }
}
int bf16_to_f32(cudaStream_t s, const void* src, float* dst, long long rows, int cols, long long ld_src, long long ld_dst) {
  {
    ProfScope ps(s, "bf16_to_f32");
    k_bf16_to_f32<<<grid_for(rows * cols, 256, 4), 256, 0, s>>>((const __nv_bfloat16*)src, dst, rows, cols, ld_src, ld_dst);
  }
  VC_CUDA(cudaGetLastError());
  return VC_OK;
}

}  // namespace vc
