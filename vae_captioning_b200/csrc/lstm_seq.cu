// Whole-sequence LSTM kernels: ONE persistent launch runs every step of the recurrence (forward) or of BPTT's
// recurrent part (backward) instead of one launch per step (reference: tf.nn.dynamic_rnn's while_loop,
// vae_model/encoder.py:38-58, decoder.py:91-121; SURVEY 7 "hard part (ii)").
//
// Grid = (N/128 row tiles) x (H/64 unit tiles) CTAs, all co-resident (the host falls back to the per-step kernels
// when the grid exceeds the SM count). CTA (m, n) owns rows [128m, 128m+128) x hidden units [64n, 64n+64) for all
// steps. Per step it needs h_{t-1} (forward) or dG_{t+1} (backward) of ITS rows for ALL units, i.e. the output of the
// H/64 CTAs that share its row tile: a per-(row tile, step) counter in global memory is bumped with a release
// reduction once a CTA has stored its slice, and the TMA producer of each CTA acquires it before loading the
// recurrent operand. Everything that does not depend on the previous step is issued early: the forward x_t part of
// the gate pre-activation accumulates into the other TMEM buffer while the previous step's epilogue is still running.
//
// Roles (384 threads) as in gemm_tc_kernel; the two epilogue groups split the tile's 64 hidden units 32 / 32.
#include "ops.h"
#include "lstm_epi.cuh"

namespace vc {

struct SeqCore {
  int m_tiles, n_tiles, iters;
  int kx_blocks;  // k-blocks of the step-independent A operand (forward: x_t; backward: 0)
  int kh_blocks;  // k-blocks of the recurrent A operand
  int bn, stages;
  int N;          // rows per step (tensor-map row of step s = s * N)
  int reverse;    // 0: step = it (forward), 1: step = iters - 1 - it
  int* flags;     // [m_tiles, iters + 1] zero-initialised; flags[m][it] = CTAs of row tile m that finished iteration it-1
};

// StepEpi: struct State (per-thread registers carried across steps);
//          __device__ void init(State&, int m_blk, int n_blk, int row, int grp) const / finish(...) const;
//          __device__ void step(uint32_t taddr, int st, int m_blk, int n_blk, int row, int grp, State&) const
template <class StepEpi>
__global__ void __launch_bounds__(kGemmThreads, 1)
lstm_seq_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmR,
                const __grid_constant__ CUtensorMap tmB, const SeqCore g, const __grid_constant__ StepEpi epi) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int stage_bytes = gemm_stage_bytes(g.bn);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + g.stages * stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kMaxStages;
  uint64_t* tfull = bars + 2 * kMaxStages;
  uint64_t* tempty = bars + 2 * kMaxStages + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_blk = blockIdx.x % g.m_tiles;
  const int n_blk = blockIdx.x / g.m_tiles;
  const int kb_total = g.kx_blocks + g.kh_blocks;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmR);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < g.stages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 8);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < g.iters; ++it) {
        const int st = g.reverse ? g.iters - 1 - it : it;
        const int row0 = st * g.N + m_blk * kBM;
        // The recurrent operand of this iteration is complete once every CTA of the row tile has published
        // (flags[m][it] == n_tiles). The weight tiles do not depend on it: up to `stages` of them are issued ahead
        // of the flag, their A halves follow as soon as it is seen.
        const int* f = g.flags + m_blk * (g.iters + 1) + it;
        const int rrow = g.reverse ? row0 + g.N : row0;
        bool ready = it == 0;
        int pend = 0, pend_stage0 = 0, pend_kb0 = 0;
        for (int kb = 0; kb < kb_total; ++kb) {
          if (kb >= g.kx_blocks && !ready && (pend == g.stages || ld_relaxed_gpu(f) >= g.n_tiles)) {
            while (ld_relaxed_gpu(f) < g.n_tiles) {
            }
            fence_acq_rel_gpu();
            fence_proxy_async_all();
            ready = true;
            for (int i = 0; i < pend; ++i) {
              const int ps = (pend_stage0 + i) % g.stages;
              tma_load_2d(smem + ps * stage_bytes, &tmR, &full[ps], (pend_kb0 + i - g.kx_blocks) * kBK, rrow);
            }
            pend = 0;
          }
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sa = smem + stage * stage_bytes;
          mbar_expect_tx(&full[stage], stage_bytes);
          tma_load_2d(sa + kABytes, &tmB, &full[stage], kb * kBK, n_blk * g.bn);
          if (kb < g.kx_blocks) {
            tma_load_2d(sa, &tmX, &full[stage], kb * kBK, row0);
          } else if (ready) {
            tma_load_2d(sa, &tmR, &full[stage], (kb - g.kx_blocks) * kBK, rrow);
          } else {
            if (pend == 0) {
              pend_stage0 = stage;
              pend_kb0 = kb;
            }
            ++pend;
          }
          if (++stage == g.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (!ready) {  // fewer recurrent k-blocks than stages: all of them are still waiting for their A half
          while (ld_relaxed_gpu(f) < g.n_tiles) {
          }
          fence_acq_rel_gpu();
          fence_proxy_async_all();
          for (int i = 0; i < pend; ++i) {
            const int ps = (pend_stage0 + i) % g.stages;
            tma_load_2d(smem + ps * stage_bytes, &tmR, &full[ps], (pend_kb0 + i - g.kx_blocks) * kBK, rrow);
          }
        }
      }
    }
  } else if (warp == 1) {
    {  // whole warp walks the steps (uniform control flow), one elected lane issues the tcgen05 instructions
      const uint32_t idesc = make_idesc_bf16(kBM, g.bn, false, false);
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < g.iters; ++it) {
        const int acc = it & 1;
        mbar_wait(&tempty[acc], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 256;
        for (int kb = 0; kb < kb_total; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * stage_bytes);
          const uint64_t adesc = make_smem_desc(sa, 16u, 1024);
          const uint64_t bdesc = make_smem_desc(sa + kABytes, 16u, 1024);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k)
              umma_bf16(d_tmem, adesc + uint64_t(k * 2), bdesc + uint64_t(k * 2), idesc, (kb > 0 || k > 0) ? 1u : 0u);
            umma_commit(&empty[stage]);
          }
          __syncwarp();
          if (++stage == g.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (elect_one()) umma_commit(&tfull[acc]);
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    const int grp = (warp - 4) >> 2;
    typename StepEpi::State state;
    epi.init(state, m_blk, n_blk, q * 32 + lane, grp);
    for (int it = 0; it < g.iters; ++it) {
      const int st = g.reverse ? g.iters - 1 - it : it;
      const int acc = it & 1;
      mbar_wait(&tfull[acc], (it >> 1) & 1);
      tc_fence_after();
      epi.step(tmem_base + acc * 256 + (uint32_t(q * 32) << 16), st, m_blk, n_blk, q * 32 + lane, grp, state);
      tc_fence_before();
      // publish: every epilogue thread's stores -> gpu-scope fence -> CTA barrier -> one release increment
      __threadfence();
      named_bar_sync(1, 256);
      if (warp == 4 && lane == 0) {
        fence_proxy_async_all();
        red_release_gpu_add(g.flags + m_blk * (g.iters + 1) + it + 1, 1);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
    }
    epi.finish(state, m_blk, n_blk, q * 32 + lane, grp);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<512>(tmem_base);
}

// ------------------------------------------------------------------------------------------
struct SeqFwdEpi {
  EpiLstmFwd e;            // per-sequence constants; the per-step pointers below are rebased inside step()
  const float* Cs;         // [steps+1, N, H]
  __nv_bfloat16* Hs;       // [steps+1, N, H]
  __nv_bfloat16* G;        // [steps, N, 4H]
  __nv_bfloat16* out;      // [T, N, H] (nullable)
  const int* lengths;      // applied to caption steps only (t >= 0)
  const float* out_keep;   // [N, T, H] (nullable)
  int pre, T;
  struct State {
    float c[32];  // cell state of this thread's row x 32 hidden units, register-resident across steps
  };
  __device__ __forceinline__ void init(State& s, int, int, int, int) const {
#pragma unroll
    for (int j = 0; j < 32; ++j) s.c[j] = 0.f;  // cell_0.zero_state
  }
  __device__ __forceinline__ void finish(State&, int, int, int, int) const {}
  __device__ __forceinline__ void step(uint32_t taddr, int st, int m_blk, int n_blk, int row, int grp, State& state) const {
    EpiLstmFwd s = e;
    const int t = st - pre;
    const size_t nh = (size_t)e.N * e.H;
    s.c_prev = Cs + (size_t)st * nh;
    s.c_out = const_cast<float*>(Cs) + (size_t)(st + 1) * nh;
    s.h_prev = Hs + (size_t)st * nh;
    s.h_out = Hs + (size_t)(st + 1) * nh;
    s.gates = G + (size_t)st * nh * 4;
    s.out = (out != nullptr && t >= 0) ? out + (size_t)t * nh : nullptr;
    s.lengths = t >= 0 ? lengths : nullptr;
    s.out_keep = (out_keep != nullptr && t >= 0) ? out_keep + (size_t)t * e.H : nullptr;
    s.t = t;
    const int m = m_blk * kBM + row;
    const bool row_ok = m < e.N;
    const bool live = row_ok && (s.lengths == nullptr || t < s.lengths[m]);
    s.template chunk<true>(taddr, m, row_ok, live, n_blk * kUPT, grp * 32, state.c);
    s.template chunk<true>(taddr, m, row_ok, live, n_blk * kUPT, grp * 32 + 16, state.c + 16);
  }
};

struct SeqBwdEpi {
  LstmBwdCommon c;           // per-sequence constants
  const __nv_bfloat16* G;    // [steps, N, 4H]
  const float* Cs;           // [steps+1, N, H]
  const float* d_out;        // [T, N, H] (nullable)
  const float* out_keep;     // [N, T, H] (nullable)
  __nv_bfloat16* dG;         // [steps, N, 4H]
  const int* lengths;
  int pre;
  struct State {
    float dh[32], dc[32];  // pass-through dh / dc of this thread's row x 32 hidden units
  };
  __device__ __forceinline__ void init(State& s, int m_blk, int n_blk, int row, int grp) const {
    const int m = m_blk * kBM + row;
#pragma unroll
    for (int j = 0; j < 32; ++j) s.dh[j] = s.dc[j] = 0.f;
    if (m < c.N) {
      const long long o = (long long)m * c.H + n_blk * 64 + grp * 32;
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        *reinterpret_cast<float4*>(s.dh + j) = *reinterpret_cast<const float4*>(c.dh_carry + o + j);
        *reinterpret_cast<float4*>(s.dc + j) = *reinterpret_cast<const float4*>(c.dc_carry + o + j);
      }
    }
  }
  __device__ __forceinline__ void finish(State& s, int m_blk, int n_blk, int row, int grp) const {
    const int m = m_blk * kBM + row;
    if (m < c.N) {
      const long long o = (long long)m * c.H + n_blk * 64 + grp * 32;
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        *reinterpret_cast<float4*>(c.dh_carry + o + j) = *reinterpret_cast<float4*>(s.dh + j);
        *reinterpret_cast<float4*>(c.dc_carry + o + j) = *reinterpret_cast<float4*>(s.dc + j);
      }
    }
  }
  __device__ __forceinline__ void step(uint32_t taddr, int st, int m_blk, int n_blk, int row, int grp, State& state) const {
    LstmBwdCommon s = c;
    const int t = st - pre;
    const size_t nh = (size_t)c.N * c.H;
    s.gates = G + (size_t)st * nh * 4;
    s.c_prev = Cs + (size_t)st * nh;
    s.c_cur = Cs + (size_t)(st + 1) * nh;
    s.d_out = (d_out != nullptr && t >= 0) ? d_out + (size_t)t * nh : nullptr;
    s.out_keep = (out_keep != nullptr && t >= 0) ? out_keep + (size_t)t * c.H : nullptr;
    s.d_gates = dG + (size_t)st * nh * 4;
    s.lengths = t >= 0 ? lengths : nullptr;
    s.t = t;
    const int m = m_blk * kBM + row;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int col = grp * 32 + h * 16;
      float acc[16];
      __syncwarp();
      tmem_ld16(taddr + col, acc);
      tmem_ld_wait();
      if (m < c.N) s.core(acc, m, n_blk * 64 + col, state.dh + h * 16, state.dc + h * 16);
    }
  }
};

template <class Epi>
static int launch_seq(const CUtensorMap& tmX, const CUtensorMap& tmR, const CUtensorMap& tmB, SeqCore core, const Epi& epi,
                      cudaStream_t stream, const char* tag) {
  static bool configured = false;
  if (!configured) {
    VC_CUDA(cudaFuncSetAttribute(lstm_seq_kernel<Epi>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  core.stages = gemm_pick_stages(core.bn, 0);
  const int smem = gemm_smem_bytes(core.bn, core.stages, 0);
  VC_CUDA(cudaMemsetAsync(core.flags, 0, (size_t)core.m_tiles * (core.iters + 1) * sizeof(int), stream));
  {
    ProfScope ps(stream, tag);
    lstm_seq_kernel<Epi><<<core.m_tiles * core.n_tiles, kGemmThreads, smem, stream>>>(tmX, tmR, tmB, core, epi);
  }
  VC_CUDA(cudaGetLastError());
  return VC_OK;
}

bool lstm_seq_applicable(int N, int H, int steps) {
  const long long ctas = (long long)((N + kBM - 1) / kBM) * (H / 64);
  return H % 64 == 0 && ctas <= num_sms() && steps >= 1;
}

int lstm_seq_flag_count(int N, int steps) { return ((N + kBM - 1) / kBM) * (steps + 2); }

int lstm_fwd_seq(cudaStream_t stream, const LstmSeqFwdArgs& a) {
  if (a.H % 64 != 0 || a.E % kBK != 0) return set_error(VC_E_SHAPE, "LSTM sizes must be multiples of 64 (E=%d H=%d)", a.E, a.H);
  const long long rows = (long long)a.steps * a.N;
  CUtensorMap tmX, tmH, tmW;
  VC_TRY(make_tmap_2d(&tmX, a.X, a.E, rows, a.E, 64, kBM));
  VC_TRY(make_tmap_2d(&tmH, a.Hs, a.H, rows + a.N, a.H, 64, kBM));
  VC_TRY(make_tmap_2d(&tmW, a.w_t_perm, (uint64_t)a.E + a.H, 4ull * a.H, (uint64_t)a.E + a.H, 64, 256));
  SeqCore core{};
  core.m_tiles = (a.N + kBM - 1) / kBM;
  core.n_tiles = a.H / 64;
  core.iters = a.steps;
  core.kx_blocks = a.E / kBK;
  core.kh_blocks = a.H / kBK;
  core.bn = 256;
  core.N = a.N;
  core.reverse = 0;
  core.flags = a.flags;
  SeqFwdEpi epi{};
  epi.e.bias = a.bias;
  epi.e.out_keep_ld = (long long)a.T * a.H;
  epi.e.inv_keep = a.inv_keep;
  epi.e.N = a.N;
  epi.e.H = a.H;
  epi.Cs = a.Cs;
  epi.Hs = (__nv_bfloat16*)a.Hs;
  epi.G = (__nv_bfloat16*)a.G;
  epi.out = (__nv_bfloat16*)a.out;
  epi.lengths = a.lengths;
  epi.out_keep = a.out_keep;
  epi.pre = a.pre;
  epi.T = a.T;
  return launch_seq(tmX, tmH, tmW, core, epi, stream, "lstm_fwd_seq");
}

// Steps steps-2 .. 0 of BPTT's recurrent part (the last step has no recurrent input and runs in k_lstm_bwd_last).
int lstm_bwd_seq(cudaStream_t stream, const LstmSeqBwdArgs& a) {
  if (a.H % 64 != 0) return set_error(VC_E_SHAPE, "LSTM hidden size must be a multiple of 64");
  if (a.steps < 2) return VC_OK;
  const long long rows = (long long)a.steps * a.N;
  CUtensorMap tmG, tmW;
  VC_TRY(make_tmap_2d(&tmG, a.dG, 4ull * a.H, rows, 4ull * a.H, 64, kBM));
  VC_TRY(make_tmap_2d(&tmW, (const __nv_bfloat16*)a.w_nat + (long long)a.E * 4 * a.H, 4ull * a.H, a.H, 4ull * a.H, 64, 64));
  SeqCore core{};
  core.m_tiles = (a.N + kBM - 1) / kBM;
  core.n_tiles = a.H / 64;
  core.iters = a.steps - 1;
  core.kx_blocks = 0;
  core.kh_blocks = 4 * a.H / kBK;
  core.bn = 64;
  core.N = a.N;
  core.reverse = 1;
  core.flags = a.flags;
  SeqBwdEpi epi{};
  epi.c.out_keep_ld = (long long)a.T * a.H;
  epi.c.inv_keep = a.inv_keep;
  epi.c.dh_carry = a.dh_carry;
  epi.c.dc_carry = a.dc_carry;
  epi.c.N = a.N;
  epi.c.H = a.H;
  epi.G = (const __nv_bfloat16*)a.G;
  epi.Cs = a.Cs;
  epi.d_out = a.d_out;
  epi.out_keep = a.out_keep;
  epi.dG = (__nv_bfloat16*)a.dG;
  epi.lengths = a.lengths;
  epi.pre = a.pre;
  return launch_seq(tmG, tmG, tmW, core, epi, stream, "lstm_bwd_seq");
}

}  // namespace vc
