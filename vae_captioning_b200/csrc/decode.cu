// Generation path (reference: Decoder.px_z_fi(gen_mode=True) vae_model/decoder.py:34-143, online_inference :145-201,
// beam_search :203-320, driven by ops/inference.py). The reference runs one sess.run per token per beam at batch 1;
// here a whole batch of images (x beams) advances per step: embedding gather -> fused-gate LSTM step (tcgen05) ->
// vocab projection (tcgen05) -> per-row softmax / top-k -> per-image beam bookkeeping, all enqueued on one stream
// with no host round trip until the captions are read back.
#include <curand_kernel.h>
#include <algorithm>
#include <utility>
#include <vector>
#include "beam_core.h"
#include "model.h"

namespace vc {

namespace {

// z[b, s*Z + k] = z_mean[b, k] + std * eps[b, s, k]  (zs.Normal('z', z_mean, std, n_samples=S) reshaped [S,1,Z] -> [1,S*Z],
// decoder.py:42-74, 109-110). z_mean = 0, or for the AG prior the mean of the c_means rows of the image's active
// clusters (decoder.py:45-71, per image instead of the reference's batch-1 graph; empty vector: all in-range clusters
// outside the hard-coded unused set, Q18).
__global__ void k_gen_z(const float* __restrict__ c_v, const float* __restrict__ c_means, const float* __restrict__ eps,
                        unsigned long long seed, float stdv, int prior, __nv_bfloat16* __restrict__ z, int B, int S, int Z,
                        int K) {
  const long long total = (long long)B * S * Z;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % Z);
    const long long b = i / ((long long)S * Z);
    float mean = 0.f;
    if (prior == VC_PRIOR_AG) {
      const float* w = c_v + b * K;
      int cnt = 0;
      float acc = 0.f;
      for (int c = 0; c < K; ++c)
        if (w[c] > 0.f) { acc += c_means[(long long)c * Z + k]; ++cnt; }
      if (cnt == 0) {
        for (int c = 0; c < K; ++c) {
          const bool unused = c == 0 || c == 66 || c == 68 || c == 69 || c == 71 || c == 12 || c == 45 || c == 83 ||
                              c == 26 || c == 29 || c == 30;
          if (!unused) { acc += c_means[(long long)c * Z + k]; ++cnt; }
        }
      }
      mean = acc / (float)cnt;
    }
    float e;
    if (eps != nullptr) {
      e = eps[i];
    } else {
      curandStatePhilox4_32_10_t st;
      curand_init(seed, (unsigned long long)i, 0, &st);
      e = curand_normal(&st);
    }
    z[i] = __float2bfloat16(mean + stdv * e);
  }
}

// dst row r <- src row map[r] (state hand-over between beam iterations; map == nullptr: r / group, i.e. broadcast of
// image b's initial state to its `group` beam slots).
__global__ void k_gather_state(const __nv_bfloat16* __restrict__ h_src, const float* __restrict__ c_src,
                               __nv_bfloat16* __restrict__ h_dst, float* __restrict__ c_dst, const int* __restrict__ map,
                               int group, int M, int H) {
  const int r = blockIdx.x;
  if (r >= M) return;
  const int s = map ? map[r] : r / group;
  for (int i = threadIdx.x; i < H; i += blockDim.x) {
    h_dst[(long long)r * H + i] = h_src[(long long)s * H + i];
    c_dst[(long long)r * H + i] = c_src[(long long)s * H + i];
  }
}

// Per-row softmax statistics + top-k of fp32 logits [M, ld]. One CTA per row, the row lives in registers.
// out_idx/out_p [M, k]: descending probability, ties -> lower index. probs (nullable) [M, V]: tf.nn.softmax.
// gumbel_seed != 0: adds Gumbel noise to logits / temperature first, so the top-1 is a tf.multinomial draw
// (decoder.py:136-138).
// One WARP per row, one pass over the row: every lane keeps a running (max, sum of exp) pair and its own sorted top-KMAX
// list while it streams its share of the row with 16-byte loads; the lanes' lists are then merged by k rounds of a
// warp arg-max over the list heads. The first version gave a 256-thread CTA to each row with the whole row in registers
// and k block-wide arg-max rounds: at 5120 rows x 11313 words it ran at 0.6 TB/s (0.37 ms per decode step, 39 % of the
// decode time) because two CTAs per SM spent their time in __syncthreads; this one is bound by the read of the logits.
constexpr int kTopRowsPerCta = 8;

template <int KMAX>
__global__ void __launch_bounds__(kTopRowsPerCta * 32)
k_row_topk(const float* __restrict__ logits, long long ld, int V, int M, int k, int* __restrict__ out_idx,
           float* __restrict__ out_p, float* __restrict__ probs, unsigned long long gumbel_seed, unsigned long long step,
           float inv_temp) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * kTopRowsPerCta + (threadIdx.x >> 5);
  if (row >= M) return;
  const float* p = logits + (long long)row * ld;
  float tv[KMAX];
  int ti[KMAX];
#pragma unroll
  for (int j = 0; j < KMAX; ++j) { tv[j] = -INFINITY; ti[j] = 0x7fffffff; }
  float mx = -INFINITY, sum = 0.f;
  // thr: a lower bound of the row's k-th largest value seen so far (the k-th largest of the 32 list heads, refreshed once
  // per 16 elements per lane). Without it nearly every element makes SOME lane insert, and the whole warp walks the
  // insertion code; with it insertions stop after the first few hundred elements. Ties: thr comes from elements with
  // lower indices than anything visited later, which win ties anyway, so the strict comparison loses nothing.
  float thr = -INFINITY;
  auto refresh_thr = [&]() {
    float h = tv[0], t = -INFINITY;
    for (int j = 0; j < k; ++j) {
      t = warp_max(h);
      const unsigned hit = __ballot_sync(0xffffffffu, h == t);
      if (lane == __ffs(hit) - 1) h = -INFINITY;
    }
    thr = t;
  };
  auto visit = [&](float x, int i) {
    if (gumbel_seed != 0) {
      curandStatePhilox4_32_10_t st;
      curand_init(gumbel_seed, (unsigned long long)row * V + i, step, &st);
      const float u = curand_uniform(&st);
      x = x * inv_temp - __logf(-__logf(u));
    }
    if (x > mx) {  // running softmax denominator relative to the running maximum (ex2-based exp: 2 ulp, summed 11k times)
      sum = sum * __expf(mx - x) + 1.f;
      mx = x;
    } else {
      sum += __expf(x - mx);
    }
    if (x > thr && x > tv[KMAX - 1]) {  // strict: among equal values the earlier (lower) index stays ahead
      tv[KMAX - 1] = x;
      ti[KMAX - 1] = i;
#pragma unroll
      for (int j = KMAX - 1; j > 0; --j) {
        if (tv[j] > tv[j - 1]) {
          const float a = tv[j]; tv[j] = tv[j - 1]; tv[j - 1] = a;
          const int b = ti[j]; ti[j] = ti[j - 1]; ti[j - 1] = b;
        }
      }
    }
  };
  const bool vec = (ld & 3) == 0 && (reinterpret_cast<uintptr_t>(logits) & 15) == 0;
  if (vec) {
    // four independent 16-byte loads per lane in flight (2 KB per warp): the visits form one dependent chain, so the
    // memory-level parallelism has to come from the loads being issued ahead of them
    for (int base = lane * 4; base - lane * 4 < V; base += 512) {  // warp-uniform trip count (refresh_thr is collective)
      if (KMAX > 1 && gumbel_seed == 0) refresh_thr();
      float4 q[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i0 = base + u * 128;
        q[u] = i0 < V ? *reinterpret_cast<const float4*>(p + i0) : make_float4(0.f, 0.f, 0.f, 0.f);  // columns V..ld-1: row padding
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i0 = base + u * 128;
        if (i0 < V) visit(q[u].x, i0);
        if (i0 + 1 < V) visit(q[u].y, i0 + 1);
        if (i0 + 2 < V) visit(q[u].z, i0 + 2);
        if (i0 + 3 < V) visit(q[u].w, i0 + 3);
      }
    }
  } else {
    for (int i = lane; i < V; i += 32) visit(p[i], i);
  }
  // row maximum and denominator
  const float m_row = warp_max(mx);
  const float s_row = warp_sum(mx == -INFINITY ? 0.f : sum * __expf(mx - m_row));
  if (probs != nullptr) {  // tf.nn.softmax of the row (the reference-granularity step call): second pass, L2-resident
    for (int i = lane; i < V; i += 32) probs[(long long)row * V + i] = expf(p[i] - m_row) / s_row;
  }
  // merge: k rounds of arg-max over the heads of the 32 sorted lists (value descending, index ascending)
  for (int j = 0; j < k; ++j) {
    float bv = tv[0];
    int bi = ti[0];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) {
      out_idx[(long long)row * k + j] = bi;
      out_p[(long long)row * k + j] = expf(bv - m_row) / s_row;
    }
    if (ti[0] == bi) {  // the winning lane pops its head
#pragma unroll
      for (int q = 0; q + 1 < KMAX; ++q) { tv[q] = tv[q + 1]; ti[q] = ti[q + 1]; }
      tv[KMAX - 1] = -INFINITY;
      ti[KMAX - 1] = 0x7fffffff;
    }
  }
}

// Greedy / sampling bookkeeping (decoder.py:170-193): append the chosen word, stop at <EOS>.
__global__ void k_greedy_update(const int* __restrict__ top_idx, int* __restrict__ tok, int* __restrict__ out_tokens,
                                int* __restrict__ out_len, int* __restrict__ done, int B, int it, int max_len, int eos) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  if (done[b]) { tok[b] = 0; return; }
  const int w = top_idx[b];
  out_tokens[(long long)b * max_len + it] = w;
  out_len[b] = it + 1;
  tok[b] = w;
  if (w == eos) done[b] = 1;
}

__global__ void k_beam_init(BeamImage* imgs, int2* nodes, int max_nodes, int* tok, int B, int beam, int bos) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  beam_image_init(&imgs[b], nodes + (long long)b * max_nodes, bos);
  for (int j = 0; j < beam; ++j) tok[b * beam + j] = bos;
}

__global__ void k_beam_update(BeamImage* imgs, int2* nodes, int max_nodes, const int* __restrict__ cand_idx,
                              const float* __restrict__ cand_p, int* __restrict__ row_map, int* __restrict__ tok, int B,
                              int beam, int eos, float len_norm) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  int src[kMaxBeam], tk[kMaxBeam];
  for (int j = 0; j < beam; ++j) { src[j] = j; tk[j] = 0; }
  beam_image_update(&imgs[b], nodes + (long long)b * max_nodes, max_nodes, cand_idx + (long long)b * beam * beam,
                    cand_p + (long long)b * beam * beam, beam, eos, len_norm, src, tk);
  for (int j = 0; j < beam; ++j) {
    row_map[b * beam + j] = b * beam + src[j];
    tok[b * beam + j] = tk[j];
  }
}

__global__ void k_beam_finish(BeamImage* imgs, const int2* nodes, int max_nodes, int* out_tokens, int* out_len,
                              float* out_score, int* out_n, int B, int beam, int max_len) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  out_n[b] = beam_image_finish(&imgs[b], nodes + (long long)b * max_nodes, beam, max_len,
                               out_tokens + (long long)b * beam * max_len, out_len + (long long)b * beam,
                               out_score + (long long)b * beam);
}

}  // namespace

// ------------------------------------------------------------------------------------------
struct DecodeWs {
  int capM = 0, capB = 0;
  void *x = nullptr, *h[2] = {nullptr, nullptr}, *feats_h = nullptr, *fv_h = nullptr, *cv_h = nullptr, *cve_h = nullptr,
       *z_h = nullptr, *zd_h = nullptr, *h0 = nullptr, *h1 = nullptr;
  float *c[2] = {nullptr, nullptr}, *c0 = nullptr, *c1 = nullptr, *logits = nullptr, *fv_f = nullptr, *cve_f = nullptr,
        *zd_f = nullptr, *st_feats = nullptr, *st_cv = nullptr, *top_p = nullptr, *out_score = nullptr, *probs = nullptr;
  int *tok = nullptr, *top_idx = nullptr, *row_map = nullptr, *out_tokens = nullptr, *out_len = nullptr, *done = nullptr,
      *out_n = nullptr;
  BeamImage* imgs = nullptr;
  int2* nodes = nullptr;
  int cur = 0, M = 0, B = 0;
  std::vector<void*> allocs;
};

static void decode_free(DecodeWs* w) {
  if (!w) return;
  for (void* p : w->allocs) cudaFree(p);
  delete w;
}

template <class T>
static int ws_alloc(DecodeWs* w, T** p, size_t count) {
  void* q = nullptr;
  cudaError_t e = cudaMalloc(&q, (count ? count : 1) * sizeof(T));
  if (e != cudaSuccess) return set_error(VC_E_NOMEM, "decode workspace: cudaMalloc(%zu) failed: %s", count * sizeof(T), cudaGetErrorString(e));
  cudaMemset(q, 0, (count ? count : 1) * sizeof(T));
  w->allocs.push_back(q);
  *p = (T*)q;
  return VC_OK;
}

constexpr int kDecMaxLen = 64;

int Model::decode_reserve(int B, int beam) {
  const int M = B * beam;
  if (dws != nullptr && dws->capM >= M && dws->capB >= B) return VC_OK;
  decode_free(dws);
  dws = new DecodeWs();
  DecodeWs* w = dws;
  const int E = cfg.embed_size, H = cfg.decoder_hidden, F = cfg.cnn_feature_size, S = cfg.gen_z_samples, Z = cfg.latent_size,
            K = cfg.num_clusters;
  w->capM = M;
  w->capB = B;
  VC_TRY(ws_alloc(w, (uint16_t**)&w->x, (size_t)M * E));
  for (int i = 0; i < 2; ++i) {
    VC_TRY(ws_alloc(w, (uint16_t**)&w->h[i], (size_t)M * H));
    VC_TRY(ws_alloc(w, &w->c[i], (size_t)M * H));
  }
  VC_TRY(ws_alloc(w, (uint16_t**)&w->h0, (size_t)B * H));
  VC_TRY(ws_alloc(w, (uint16_t**)&w->h1, (size_t)B * H));
  VC_TRY(ws_alloc(w, &w->c0, (size_t)B * H));
  VC_TRY(ws_alloc(w, &w->c1, (size_t)B * H));
  VC_TRY(ws_alloc(w, (uint16_t**)&w->feats_h, (size_t)B * F));
  VC_TRY(ws_alloc(w, &w->fv_f, (size_t)B * E));
  VC_TRY(ws_alloc(w, (uint16_t**)&w->fv_h, (size_t)B * E));
  VC_TRY(ws_alloc(w, (uint16_t**)&w->cv_h, (size_t)B * KP));
  VC_TRY(ws_alloc(w, &w->cve_f, (size_t)B * E));
  VC_TRY(ws_alloc(w, (uint16_t**)&w->cve_h, (size_t)B * E));
  VC_TRY(ws_alloc(w, (uint16_t**)&w->z_h, (size_t)B * S * Z));
  VC_TRY(ws_alloc(w, &w->zd_f, (size_t)B * E));
  VC_TRY(ws_alloc(w, (uint16_t**)&w->zd_h, (size_t)B * E));
  VC_TRY(ws_alloc(w, &w->st_feats, (size_t)B * F));
  VC_TRY(ws_alloc(w, &w->st_cv, (size_t)B * K));
  VC_TRY(ws_alloc(w, &w->logits, (size_t)M * VP));
  VC_TRY(ws_alloc(w, &w->tok, (size_t)M));
  VC_TRY(ws_alloc(w, &w->top_idx, (size_t)M * kMaxBeam));
  VC_TRY(ws_alloc(w, &w->top_p, (size_t)M * kMaxBeam));
  VC_TRY(ws_alloc(w, &w->row_map, (size_t)M));
  VC_TRY(ws_alloc(w, &w->out_tokens, (size_t)M * kDecMaxLen));
  VC_TRY(ws_alloc(w, &w->out_len, (size_t)M));
  VC_TRY(ws_alloc(w, &w->out_score, (size_t)M));
  VC_TRY(ws_alloc(w, &w->out_n, (size_t)B));
  VC_TRY(ws_alloc(w, &w->done, (size_t)B));
  VC_TRY(ws_alloc(w, &w->imgs, (size_t)B));
  VC_TRY(ws_alloc(w, &w->nodes, (size_t)B * (1 + kMaxBeam * kDecMaxLen)));
  return VC_OK;
}

void Model::decode_release() {
  decode_free(dws);
  dws = nullptr;
}

// Initial decoder state of B images: LSTM(images_fv) -> [LSTM(c_i)] -> [LSTM(z_dec)] from the zero state
// (decoder.py:96-114), left in dws->h0 / c0 as [B, H].
int Model::decode_begin(const float* feats_dev, const float* c_v_dev, int B, const vc_rng* rng, cudaStream_t s) {
  const int E = cfg.embed_size, H = cfg.decoder_hidden, F = cfg.cnn_feature_size, S = cfg.gen_z_samples, Z = cfg.latent_size,
            K = cfg.num_clusters;
  DecodeWs* w = dws;
  const bool has_cv = cfg.use_c_v || cfg.prior != VC_PRIOR_NORMAL;
  if (has_cv && c_v_dev == nullptr) return set_error(VC_E_ARG, "this configuration needs cluster vectors (c_v)");
  VC_TRY(refresh_decode_shadows(s));
  VC_TRY(cast_f32_bf16(s, feats_dev, w->feats_h, B, F, F, F));
  // Generation runs its two long contractions (imf_emb K = 4096, z_rnn K = S*Z = 15000) un-split: split-K meets in fp32
  // atomics whose order changes from run to run, and a 1e-7 wobble in the initial state is enough to flip a bf16
  // rounding and with it a near-tie between tokens or beams (measured: 6-8 of 1024 captions differed between identical
  // calls). One tile per CTA costs ~30 us per decode call and makes the token output reproducible bit for bit.
  {
    Operand A{w->feats_h, B, F, F, false}, Bw{imf_wt, E, F, F, false};
    EpiStore e{};
    e.out = w->fv_f; e.ld = E; e.bias = pp(pidx("imf_emb/bias")); e.alpha = 1.f;
    ProfTag ptag("imf_emb");
    VC_TRY(gemm_store(s, A, nullptr, 0, Bw, B, E, F, e, 64, 1));
  }
  VC_TRY(cast_f32_bf16(s, w->fv_f, w->fv_h, B, E, E, E));
  VC_CUDA(cudaMemsetAsync(w->h0, 0, (size_t)B * H * 2, s));
  VC_CUDA(cudaMemsetAsync(w->c0, 0, (size_t)B * H * sizeof(float), s));
  void* hp = w->h0; float* cp = w->c0; void* hn = w->h1; float* cn = w->c1;
  auto cell = [&](const void* x) -> int {
    LstmFwdArgs a{};
    a.x = x; a.h_prev = hp; a.c_prev = cp; a.h_out = hn; a.c_out = cn;
    a.w_t_perm = dec.w_t_perm; a.bias = pp(dec.p_bias); a.inv_keep = 1.f; a.t = 0; a.N = B; a.E = E; a.H = H; a.precise = 1;
    VC_TRY(lstm_fwd_step(s, a));
    std::swap(hp, hn);
    std::swap(cp, cn);
    return VC_OK;
  };
  VC_TRY(cell(w->fv_h));
  if (cfg.use_c_v) {
    VC_TRY(cast_f32_bf16(s, c_v_dev, w->cv_h, B, K, K, KP));
    Operand A{w->cv_h, B, K, KP, false}, Bw{cv_wt, E, K, KP, false};
    EpiStore e{};
    e.out = w->cve_f; e.ld = E; e.bias = pp(pidx("cv_emb/bias")); e.alpha = 1.f;
    ProfTag ptag("cv_emb");
    VC_TRY(gemm_store(s, A, nullptr, 0, Bw, B, E, K, e, 64, 1));
    VC_TRY(cast_f32_bf16(s, w->cve_f, w->cve_h, B, E, E, E));
    VC_TRY(cell(w->cve_h));
  }
  if (!cfg.no_encoder) {
    const long long n = (long long)B * S * Z;
    {
      ProfScope ps(s, "gen_z");
      k_gen_z<<<(int)std::min<long long>((n + 255) / 256, 148 * 16), 256, 0, s>>>(
          c_v_dev, c_means, rng ? rng->eps_dev : nullptr, rng ? rng->seed : 0ull, cfg.std, cfg.prior, (__nv_bfloat16*)w->z_h, B,
          S, Z, K);
    }
    VC_CUDA(cudaGetLastError());
    const int SZ = S * Z;
    Operand A{w->z_h, B, SZ, SZ, false}, Bw{z_wt, E, SZ, SZ, false};
    EpiStore e{};
    e.out = w->zd_f; e.ld = E; e.bias = pp(pidx("decoder/net/z_rnn/bias")); e.alpha = 1.f;
    {
      ProfTag ptag("z_rnn");
      VC_TRY(gemm_store(s, A, nullptr, 0, Bw, B, E, SZ, e, 64, 1));
    }
    VC_TRY(cast_f32_bf16(s, w->zd_f, w->zd_h, B, E, E, E));
    VC_TRY(cell(w->zd_h));
  }
  if (hp != w->h0) {  // odd number of cell applications: result sits in h1/c1
    VC_CUDA(cudaMemcpyAsync(w->h0, hp, (size_t)B * H * 2, cudaMemcpyDeviceToDevice, s));
    VC_CUDA(cudaMemcpyAsync(w->c0, cp, (size_t)B * H * sizeof(float), cudaMemcpyDeviceToDevice, s));
  }
  return VC_OK;
}

// One sess.run([sample, out_state]) for M rows (decoder.py:77-142 in gen mode): embedding lookup of dws->tok,
// LSTM step from state buffer `cur` into the other, vocab projection into dws->logits.
int Model::decode_advance(int M, cudaStream_t s) {
  DecodeWs* w = dws;
  const int E = cfg.embed_size, H = cfg.decoder_hidden, V = cfg.vocab_size;
  VC_TRY(embed_gather(s, dec_emb_h, w->tok, w->x, nullptr, 1.f, M, 1, E, V));
  LstmFwdArgs a{};
  a.x = w->x; a.h_prev = w->h[w->cur]; a.c_prev = w->c[w->cur]; a.h_out = w->h[w->cur ^ 1]; a.c_out = w->c[w->cur ^ 1];
  a.w_t_perm = dec.w_t_perm; a.bias = pp(dec.p_bias); a.inv_keep = 1.f; a.t = 0; a.N = M; a.E = E; a.H = H; a.precise = 1;
  VC_TRY(lstm_fwd_step(s, a));
  w->cur ^= 1;
  Operand A{w->h[w->cur], M, H, H, false}, Bw{wo_t, V, H, H, false};
  ProfTag ptag("logits_decode");
  // fp32 logits (token ties are decided at the 1e-4 level) through the TMA-store epilogue: whole 128-byte row segments
  return gemm_tma_rows_f32(s, A, Bw, M, V, H, w->logits, VP, pp(pidx("decoder/rnn_logits/bias")), 256);
}

static int launch_topk(cudaStream_t s, const float* logits, long long ld, int V, int M, int k, int* idx, float* p, float* probs,
                       unsigned long long gumbel_seed, unsigned long long step, float inv_temp) {
  if (k < 1 || k > kMaxBeam) return set_error(VC_E_ARG, "top-k of %d out of range (1..%d)", k, kMaxBeam);
  const int grid = (M + kTopRowsPerCta - 1) / kTopRowsPerCta, block = kTopRowsPerCta * 32;
  {
    ProfScope ps(s, "row_topk");
    if (k == 1)
      k_row_topk<1><<<grid, block, 0, s>>>(logits, ld, V, M, k, idx, p, probs, gumbel_seed, step, inv_temp);
    else if (k <= 8)
      k_row_topk<8><<<grid, block, 0, s>>>(logits, ld, V, M, k, idx, p, probs, gumbel_seed, step, inv_temp);
    else
      k_row_topk<16><<<grid, block, 0, s>>>(logits, ld, V, M, k, idx, p, probs, gumbel_seed, step, inv_temp);
  }
  VC_CUDA(cudaGetLastError());
  return VC_OK;
}

static int gather_state(cudaStream_t s, DecodeWs* w, const void* h_src, const float* c_src, void* h_dst, float* c_dst,
                        const int* map, int group, int M, int H) {
  {
    ProfScope ps(s, "gather_state");
    k_gather_state<<<M, 128, 0, s>>>((const __nv_bfloat16*)h_src, c_src, (__nv_bfloat16*)h_dst, c_dst, map, group, M, H);
  }
  VC_CUDA(cudaGetLastError());
  return VC_OK;
}

// Decoder.online_inference for a batch (decoder.py:145-201). mode 0: greedy (argmax; the temperature rescale is a
// no-op, Q10), 1: sample (tf.multinomial(logits / temperature)). out_tokens [B, max_len] zero padded, out_len [B].
int Model::decode_greedy(const float* feats_dev, const float* c_v_dev, int B, int max_len, int mode, const vc_rng* rng,
                         int bos, int eos, int32_t* out_tokens_host, int32_t* out_len_host, cudaStream_t s) {
  if (B < 1) return set_error(VC_E_SHAPE, "decode: empty batch");
  if (max_len < 1 || max_len > kDecMaxLen) return set_error(VC_E_ARG, "decode: max_len %d out of range (1..%d)", max_len, kDecMaxLen);
  VC_TRY(decode_reserve(B, 1));
  DecodeWs* w = dws;
  const int H = cfg.decoder_hidden, V = cfg.vocab_size;
  VC_TRY(decode_begin(feats_dev, c_v_dev, B, rng, s));
  w->cur = 0;
  VC_TRY(gather_state(s, w, w->h0, w->c0, w->h[0], w->c[0], nullptr, 1, B, H));
  std::vector<int> init(B, bos);
  VC_CUDA(cudaMemcpyAsync(w->tok, init.data(), B * sizeof(int), cudaMemcpyHostToDevice, s));
  VC_CUDA(cudaStreamSynchronize(s));  // `init` is pageable host memory
  VC_CUDA(cudaMemsetAsync(w->done, 0, B * sizeof(int), s));
  VC_CUDA(cudaMemsetAsync(w->out_len, 0, B * sizeof(int), s));
  VC_CUDA(cudaMemsetAsync(w->out_tokens, 0, (size_t)B * max_len * sizeof(int), s));
  const unsigned long long gseed = mode == 1 ? ((rng ? rng->seed : 0ull) | 1ull) : 0ull;
  for (int it = 0; it < max_len; ++it) {
    VC_TRY(decode_advance(B, s));
    VC_TRY(launch_topk(s, w->logits, VP, V, B, 1, w->top_idx, w->top_p, nullptr, gseed, (unsigned long long)it,
                       1.f / cfg.temperature));
    {
      ProfScope ps(s, "greedy_update");
      k_greedy_update<<<(B + 127) / 128, 128, 0, s>>>(w->top_idx, w->tok, w->out_tokens, w->out_len, w->done, B, it, max_len, eos);
    }
    VC_CUDA(cudaGetLastError());
  }
  VC_CUDA(cudaMemcpyAsync(out_tokens_host, w->out_tokens, (size_t)B * max_len * sizeof(int), cudaMemcpyDeviceToHost, s));
  VC_CUDA(cudaMemcpyAsync(out_len_host, w->out_len, B * sizeof(int), cudaMemcpyDeviceToHost, s));
  VC_CUDA(cudaStreamSynchronize(s));
  return VC_OK;
}

// Decoder.beam_search for a batch (decoder.py:203-320, Q9). Outputs per image `beam` sentences sorted by score
// (out_tokens [B, beam, max_len] incl. <BOS>/<EOS>, out_len / out_score [B, beam], out_n [B] = beams returned).
int Model::decode_beam(const float* feats_dev, const float* c_v_dev, int B, int beam, int max_len, float len_norm,
                       const vc_rng* rng, int bos, int eos, int32_t* out_tokens_host, int32_t* out_len_host,
                       float* out_score_host, int32_t* out_n_host, cudaStream_t s) {
  if (B < 1) return set_error(VC_E_SHAPE, "decode: empty batch");
  if (beam < 1 || beam > kMaxBeam) return set_error(VC_E_ARG, "beam_size %d out of range (1..%d)", beam, kMaxBeam);
  if (max_len < 2 || max_len > kDecMaxLen) return set_error(VC_E_ARG, "decode: max_len %d out of range (2..%d)", max_len, kDecMaxLen);
  VC_TRY(decode_reserve(B, beam));
  DecodeWs* w = dws;
  const int H = cfg.decoder_hidden, V = cfg.vocab_size, M = B * beam;
  const int max_nodes = 1 + kMaxBeam * kDecMaxLen;
  VC_TRY(decode_begin(feats_dev, c_v_dev, B, rng, s));
  w->cur = 0;
  VC_TRY(gather_state(s, w, w->h0, w->c0, w->h[0], w->c[0], nullptr, beam, M, H));
  {
    ProfScope ps(s, "beam_init");
    k_beam_init<<<(B + 63) / 64, 64, 0, s>>>(w->imgs, w->nodes, max_nodes, w->tok, B, beam, bos);
  }
  VC_CUDA(cudaGetLastError());
  // decoder.py:230-236: the first run consumes <BOS>; only its state is kept (the probabilities are discarded)
  VC_TRY(decode_advance(M, s));
  for (int it = 0; it < max_len - 1; ++it) {
    VC_TRY(decode_advance(M, s));  // feeds sentence[-1] (<BOS> again on the first iteration, Q9)
    VC_TRY(launch_topk(s, w->logits, VP, V, M, beam, w->top_idx, w->top_p, nullptr, 0ull, 0ull, 1.f));
    {
      ProfScope ps(s, "beam_update");
      k_beam_update<<<(B + 63) / 64, 64, 0, s>>>(w->imgs, w->nodes, max_nodes, w->top_idx, w->top_p, w->row_map, w->tok, B, beam,
                                                 eos, len_norm);
    }
    VC_CUDA(cudaGetLastError());
    VC_TRY(gather_state(s, w, w->h[w->cur], w->c[w->cur], w->h[w->cur ^ 1], w->c[w->cur ^ 1], w->row_map, 1, M, H));
    w->cur ^= 1;
  }
  {
    ProfScope ps(s, "beam_finish");
    k_beam_finish<<<(B + 63) / 64, 64, 0, s>>>(w->imgs, w->nodes, max_nodes, w->out_tokens, w->out_len, w->out_score, w->out_n, B,
                                               beam, max_len);
  }
  VC_CUDA(cudaGetLastError());
  VC_CUDA(cudaMemcpyAsync(out_tokens_host, w->out_tokens, (size_t)M * max_len * sizeof(int), cudaMemcpyDeviceToHost, s));
  VC_CUDA(cudaMemcpyAsync(out_len_host, w->out_len, M * sizeof(int), cudaMemcpyDeviceToHost, s));
  VC_CUDA(cudaMemcpyAsync(out_score_host, w->out_score, M * sizeof(float), cudaMemcpyDeviceToHost, s));
  VC_CUDA(cudaMemcpyAsync(out_n_host, w->out_n, B * sizeof(int), cudaMemcpyDeviceToHost, s));
  VC_CUDA(cudaStreamSynchronize(s));
  return VC_OK;
}

// The reference's own granularity, kept for callers that run their loop in Python: begin = the state-less first
// feed (initial_state left at its default), step = one sess.run([sample, out_state]) for M rows.
int Model::decode_open(const float* feats_dev, const float* c_v_dev, int B, const vc_rng* rng, cudaStream_t s) {
  VC_TRY(decode_reserve(B, 1));
  VC_TRY(decode_begin(feats_dev, c_v_dev, B, rng, s));
  dws->cur = 0;
  dws->M = B;
  return gather_state(s, dws, dws->h0, dws->c0, dws->h[0], dws->c[0], nullptr, 1, B, cfg.decoder_hidden);
}

int Model::decode_step(const int32_t* tok_host, int M, float* probs_host, cudaStream_t s) {
  if (dws == nullptr || dws->M != M) return set_error(VC_E_STATE, "vc_decode_step: call vc_decode_begin with %d rows first", M);
  DecodeWs* w = dws;
  const int V = cfg.vocab_size;
  VC_CUDA(cudaMemcpyAsync(w->tok, tok_host, M * sizeof(int), cudaMemcpyHostToDevice, s));
  VC_CUDA(cudaStreamSynchronize(s));
  VC_TRY(decode_advance(M, s));
  if (probs_host != nullptr) {
    if (w->probs == nullptr) VC_TRY(ws_alloc(w, &w->probs, (size_t)w->capM * V));
    VC_TRY(launch_topk(s, w->logits, VP, V, M, 1, w->top_idx, w->top_p, w->probs, 0ull, 0ull, 1.f));
    VC_CUDA(cudaMemcpyAsync(probs_host, w->probs, (size_t)M * V * sizeof(float), cudaMemcpyDeviceToHost, s));
  }
  VC_CUDA(cudaStreamSynchronize(s));
  return VC_OK;
}

int Model::decode_state(float* c_host, float* h_host, const float* c_in, const float* h_in, cudaStream_t s) {
  if (dws == nullptr || dws->M == 0) return set_error(VC_E_STATE, "no decode in progress");
  DecodeWs* w = dws;
  const int H = cfg.decoder_hidden, M = w->M;
  VC_CUDA(cudaStreamSynchronize(s));
  float* tmp = nullptr;
  VC_CUDA(cudaMalloc((void**)&tmp, (size_t)M * H * sizeof(float)));
  int st = VC_OK;
  if (h_in != nullptr && c_in != nullptr) {  // feed `in_state` (rnn_placeholders, utils/rnn_model.py:7-21)
    if (cudaMemcpy(w->c[w->cur], c_in, (size_t)M * H * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(tmp, h_in, (size_t)M * H * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess)
      st = set_error(VC_E_CUDA, "decode state upload failed");
    if (st == VC_OK) st = cast_f32_bf16(s, tmp, w->h[w->cur], M, H, H, H);
    cudaStreamSynchronize(s);
  }
  if (st == VC_OK && c_host != nullptr && h_host != nullptr) {  // fetch `out_state`
    st = bf16_to_f32(s, w->h[w->cur], tmp, M, H, H, H);
    cudaStreamSynchronize(s);
    if (st == VC_OK && (cudaMemcpy(h_host, tmp, (size_t)M * H * sizeof(float), cudaMemcpyDeviceToHost) != cudaSuccess ||
                        cudaMemcpy(c_host, w->c[w->cur], (size_t)M * H * sizeof(float), cudaMemcpyDeviceToHost) != cudaSuccess))
      st = set_error(VC_E_CUDA, "decode state download failed");
  }
  cudaFree(tmp);
  return st;
}

int Model::decode_stage(const float* feats_host, const float* c_v_host, int B, int beam, const float** feats_dev,
                        const float** c_v_dev, cudaStream_t s) {
  if (B < 1) return set_error(VC_E_SHAPE, "decode: empty batch");
  if (beam < 1 || beam > kMaxBeam) return set_error(VC_E_ARG, "beam_size %d out of range (1..%d)", beam, kMaxBeam);
  VC_TRY(decode_reserve(B, beam));
  const int F = cfg.cnn_feature_size, K = cfg.num_clusters;
  VC_CUDA(cudaMemcpyAsync(dws->st_feats, feats_host, (size_t)B * F * sizeof(float), cudaMemcpyHostToDevice, s));
  *feats_dev = dws->st_feats;
  *c_v_dev = nullptr;
  if (c_v_host != nullptr) {
    VC_CUDA(cudaMemcpyAsync(dws->st_cv, c_v_host, (size_t)B * K * sizeof(float), cudaMemcpyHostToDevice, s));
    *c_v_dev = dws->st_cv;
  }
  return VC_OK;
}

}  // namespace vc

// ------------------------------------------------------------------------------------------
// Host-only driver of the same beam bookkeeping against a callback (no GPU): used to pin beam_core.h against the
// golden vectors produced by the reference's Decoder.beam_search (tests/golden/decode_loops.json).
extern "C" int vc_beam_search_host(int (*step)(void* user, int token, int state_in, float* probs_out), void* user, int V,
                                   int beam, int max_len, int bos, int eos, float len_norm, int32_t* out_tokens,
                                   int32_t* out_len, float* out_score, int32_t* out_n) {
  using namespace vc;
  if (!step || V < 1 || beam < 1 || beam > kMaxBeam || max_len < 2 || max_len > kDecMaxLen)
    return set_error(VC_E_ARG, "vc_beam_search_host: bad arguments");
  std::vector<float> probs((size_t)V);
  std::vector<int2> nodes(1 + kMaxBeam * kDecMaxLen);
  BeamImage im;
  beam_image_init(&im, nodes.data(), bos);
  int states[kMaxBeam], tokens[kMaxBeam];
  states[0] = step(user, bos, -1, probs.data());
  tokens[0] = bos;
  std::vector<int> cand_idx((size_t)beam * beam), order((size_t)V);
  std::vector<float> cand_p((size_t)beam * beam);
  for (int it = 0; it < max_len - 1 && !im.done; ++it) {
    const int np = im.partial.n;
    int new_states[kMaxBeam];
    for (int i = 0; i < np; ++i) {
      new_states[i] = step(user, tokens[i], states[i], probs.data());
      // top-`beam` by probability, ties -> lower index (stable sort on -p)
      std::vector<char> taken((size_t)V, 0);
      for (int j = 0; j < beam; ++j) {
        int best = -1;
        for (int v = 0; v < V; ++v)
          if (!taken[v] && (best < 0 || probs[v] > probs[best])) best = v;
        if (best < 0) { cand_idx[i * beam + j] = 0; cand_p[i * beam + j] = 0.f; continue; }
        taken[best] = 1;
        cand_idx[i * beam + j] = best;
        cand_p[i * beam + j] = probs[best];
      }
    }
    int src[kMaxBeam], tk[kMaxBeam];
    const int n = beam_image_update(&im, nodes.data(), (int)nodes.size(), cand_idx.data(), cand_p.data(), beam, eos, len_norm,
                                    src, tk);
    for (int j = 0; j < n; ++j) {
      states[j] = new_states[src[j]];
      tokens[j] = tk[j];
    }
  }
  *out_n = beam_image_finish(&im, nodes.data(), beam, max_len, out_tokens, out_len, out_score);
  return VC_OK;
}
