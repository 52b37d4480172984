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
#include <cstdio>
#include <cstdlib>
#include <vector>
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
  long long* trace;  // debug (VC_LSTM_TRACE=1): [iters][4] clock64() stamps of CTA 0: flag seen, accumulator complete,
                     // epilogue done, published
};

// StepEpi: struct State (per-thread registers carried across steps); static constexpr int kSmemBytes (staging);
//          __device__ void init(State&, int m_blk, int n_blk, int row, int grp) const / finish(...) const;
//          __device__ void prefetch(int st, int m_blk, int n_blk, int row, int grp, uint8_t* smem) const  (before the
//          accumulator of the step is awaited: stage whatever the step needs that earlier kernels produced)
//          __device__ void step(uint32_t taddr, int st, int m_blk, int n_blk, int row, int grp, State&, uint8_t* smem) const
template <class StepEpi>
__global__ void __launch_bounds__(kGemmThreads, 1)
lstm_seq_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmR,
                const __grid_constant__ CUtensorMap tmB, const SeqCore g, const __grid_constant__ StepEpi epi) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int stage_bytes = gemm_stage_bytes(g.bn);
  uint8_t* epi_smem = smem + g.stages * stage_bytes;  // StepEpi::kSmemBytes of epilogue staging, then the barriers
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_smem + StepEpi::kSmemBytes);
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
            if (g.trace && blockIdx.x == 0) g.trace[it * 4 + 0] = clock64();
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
      epi.prefetch(st, m_blk, n_blk, q * 32 + lane, grp, epi_smem);  // operands that do not depend on the recurrence
      mbar_wait(&tfull[acc], (it >> 1) & 1);
      tc_fence_after();
      const bool tr = g.trace && blockIdx.x == 0 && warp == 4 && lane == 0;
      if (tr) g.trace[it * 4 + 1] = clock64();
      epi.step(tmem_base + acc * 256 + (uint32_t(q * 32) << 16), st, m_blk, n_blk, q * 32 + lane, grp, state, epi_smem);
      if (tr) g.trace[it * 4 + 2] = clock64();
      tc_fence_before();
      // publish: the CTA barrier orders every epilogue thread's stores before the elected thread, whose gpu-scope fence
      // is cumulative over them (the pattern of a cooperative-groups grid barrier); then one release increment. A
      // __threadfence() in each of the 256 threads did the same job 256 times.
      named_bar_sync(1, 256);
      if (warp == 4 && lane == 0) {
        fence_acq_rel_gpu();
        fence_proxy_async_all();
        red_release_gpu_add(g.flags + m_blk * (g.iters + 1) + it + 1, 1);
        if (tr) g.trace[it * 4 + 3] = clock64();
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
  // A TMEM lane is a tile row, so an epilogue thread owns one row x 32 hidden units and every plane it produces (h, the
  // emitted output, the four activated gates, the cell state) leaves it in 16 / 32-byte pieces 1 KB apart: a warp-wide
  // 16-byte store touches 32 cache lines, and 32 such stores per thread per step kept the load/store unit busy for
  // ~16 k cycles per step (a timeline of CTA 0 showed the epilogue at 19 k of the 32 k cycles of a step). So the pieces
  // go through shared memory first: the two warps that share 32 rows (same TMEM lane quarter, unit halves 16 | 16) stage
  // 32 units of every plane per pass, meet at a 64-thread barrier, and write the planes out as whole 64-byte (bf16) /
  // 128-byte (fp32) row segments -- 8 resp. 4 lines per store instruction. Two passes cover the CTA's 64 units.
  static constexpr int kPairBytes = 6 * 2048 + 4096;  // per 32-row pair: six bf16 planes [32 x 64 B] + cell state [32 x 128 B]
  static constexpr int kSmemBytes = 4 * kPairBytes;   // 64 KB
  __device__ __forceinline__ void prefetch(int, int, int, int, int, uint8_t*) const {}
  __device__ __forceinline__ void step(uint32_t taddr, int st, int m_blk, int n_blk, int row, int grp, State& state,
                                       uint8_t* stage_smem) const {
    const int t = st - pre;
    const size_t nh = (size_t)e.N * e.H;
    float* c_out = const_cast<float*>(Cs) + (size_t)(st + 1) * nh;
    const __nv_bfloat16* h_prev = Hs + (size_t)st * nh;
    __nv_bfloat16* h_out = Hs + (size_t)(st + 1) * nh;
    __nv_bfloat16* gates = G + (size_t)st * nh * 4;
    __nv_bfloat16* outp = (out != nullptr && t >= 0) ? out + (size_t)t * nh : nullptr;
    const float* keep = (out_keep != nullptr && t >= 0) ? out_keep + (size_t)t * e.H : nullptr;
    const int lane = row & 31, q = row >> 5;
    const int m = m_blk * kBM + row;
    const bool row_ok = m < e.N;
    const bool live = row_ok && (t < 0 || lengths == nullptr || t < lengths[m]);
    uint8_t* sb = stage_smem + q * kPairBytes;
    const int pt = grp * 32 + lane;  // thread index inside the pair
#pragma unroll 1
    for (int ps = 0; ps < 2; ++ps) {
      const int ul = ps * 32 + grp * 16;  // this thread's 16 units inside the CTA's 64
      float* cp = state.c + ps * 16;
      float gi[16], gj[16], gf[16], go[16];
      __syncwarp();
      tmem_ld16(taddr + 0 * kUPT + ul, gi);
      tmem_ld16(taddr + 1 * kUPT + ul, gj);
      tmem_ld16(taddr + 2 * kUPT + ul, gf);
      tmem_ld16(taddr + 3 * kUPT + ul, go);
      const int u0 = n_blk * kUPT + ul;
      float bi[16], bj[16], bf[16], bo[16];
#pragma unroll
      for (int j = 0; j < 16; j += 4) {
        *reinterpret_cast<float4*>(bi + j) = __ldg(reinterpret_cast<const float4*>(e.bias + u0 + j));
        *reinterpret_cast<float4*>(bj + j) = __ldg(reinterpret_cast<const float4*>(e.bias + e.H + u0 + j));
        *reinterpret_cast<float4*>(bf + j) = __ldg(reinterpret_cast<const float4*>(e.bias + 2 * e.H + u0 + j));
        *reinterpret_cast<float4*>(bo + j) = __ldg(reinterpret_cast<const float4*>(e.bias + 3 * e.H + u0 + j));
      }
      tmem_ld_wait();
      uint32_t hp[8], op[8], gp[4][8];
      if (live) {
        float hn[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float i_ = sigmoid_fast(gi[j] + bi[j]);
          const float j_ = tanh_fast(gj[j] + bj[j]);
          const float f_ = sigmoid_fast(gf[j] + bf[j] + 1.0f);
          const float o_ = sigmoid_fast(go[j] + bo[j]);
          const float cn = f_ * cp[j] + i_ * j_;
          hn[j] = o_ * tanh_fast(cn);
          cp[j] = cn;
          gi[j] = i_; gj[j] = j_; gf[j] = f_; go[j] = o_;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          hp[j] = pack_bf16(hn[2 * j], hn[2 * j + 1]);
          gp[0][j] = pack_bf16(gi[2 * j], gi[2 * j + 1]);
          gp[1][j] = pack_bf16(gj[2 * j], gj[2 * j + 1]);
          gp[2][j] = pack_bf16(gf[2 * j], gf[2 * j + 1]);
          gp[3][j] = pack_bf16(go[2 * j], go[2 * j + 1]);
        }
        if (keep != nullptr) {  // the emitted value is the bf16 state scaled by the keep mask (the state itself is untouched)
          const float* kp = keep + (long long)m * e.out_keep_ld + u0;
#pragma unroll
          for (int j = 0; j < 16; ++j) hn[j] = __bfloat162float(__float2bfloat16(hn[j])) * kp[j] * e.inv_keep;
#pragma unroll
          for (int j = 0; j < 8; ++j) op[j] = pack_bf16(hn[2 * j], hn[2 * j + 1]);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) op[j] = hp[j];
        }
      } else {
        // past the end of the sequence (or a padding row of the tile): state copied through, emitted output zero
        // (SURVEY 5.2); the gate planes of such rows are never read by BPTT and are written as zeros
        if (row_ok) {
          const uint4* hsrc = reinterpret_cast<const uint4*>(h_prev + (long long)m * e.H + u0);
          *reinterpret_cast<uint4*>(hp) = hsrc[0];
          *reinterpret_cast<uint4*>(hp + 4) = hsrc[1];
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (!row_ok) hp[j] = 0u;
          op[j] = 0u;
          gp[0][j] = gp[1][j] = gp[2][j] = gp[3][j] = 0u;
        }
      }
      // ---- stage: bf16 planes with 64-byte rows (16-byte chunk c of row r at c ^ ((r >> 1) & 3)), cell state with
      // 128-byte rows (chunk c at c ^ (r & 7)): every warp-wide 16-byte access is four conflict-free wavefronts
      {
        const int sw = (lane >> 1) & 3;
        uint8_t* rb = sb + lane * 64;
        const int c0 = ((grp * 2) ^ sw) << 4, c1 = ((grp * 2 + 1) ^ sw) << 4;
        *reinterpret_cast<uint4*>(rb + 0 * 2048 + c0) = *reinterpret_cast<uint4*>(hp);
        *reinterpret_cast<uint4*>(rb + 0 * 2048 + c1) = *reinterpret_cast<uint4*>(hp + 4);
        *reinterpret_cast<uint4*>(rb + 1 * 2048 + c0) = *reinterpret_cast<uint4*>(op);
        *reinterpret_cast<uint4*>(rb + 1 * 2048 + c1) = *reinterpret_cast<uint4*>(op + 4);
#pragma unroll
        for (int gg = 0; gg < 4; ++gg) {
          *reinterpret_cast<uint4*>(rb + (2 + gg) * 2048 + c0) = *reinterpret_cast<uint4*>(gp[gg]);
          *reinterpret_cast<uint4*>(rb + (2 + gg) * 2048 + c1) = *reinterpret_cast<uint4*>(gp[gg] + 4);
        }
        uint8_t* cb = sb + 6 * 2048 + lane * 128;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          *reinterpret_cast<uint4*>(cb + (((grp * 4 + j) ^ (lane & 7)) << 4)) = *reinterpret_cast<uint4*>(cp + 4 * j);
      }
      named_bar_sync(2 + q, 64);
      // ---- write out whole row segments: 32 contiguous units of 32 rows per plane
      {
        const long long m0 = (long long)m_blk * kBM + q * 32;
        const int ucol = n_blk * kUPT + ps * 32;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int e16 = pt + 64 * i;
          const int r = e16 >> 2, c = e16 & 3;
          if (m0 + r < e.N) {
            const uint8_t* src = sb + r * 64 + ((c ^ ((r >> 1) & 3)) << 4);
            const long long o = (m0 + r) * e.H + ucol + c * 8;
            *reinterpret_cast<uint4*>(h_out + o) = *reinterpret_cast<const uint4*>(src);
            if (outp != nullptr) *reinterpret_cast<uint4*>(outp + o) = *reinterpret_cast<const uint4*>(src + 2048);
            __nv_bfloat16* gdst = gates + (m0 + r) * 4 * e.H + ucol + c * 8;
#pragma unroll
            for (int gg = 0; gg < 4; ++gg)
              *reinterpret_cast<uint4*>(gdst + (long long)gg * e.H) = *reinterpret_cast<const uint4*>(src + (2 + gg) * 2048);
          }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int e16 = pt + 64 * i;
          const int r = e16 >> 3, c = e16 & 7;
          if (m0 + r < e.N)
            *reinterpret_cast<uint4*>(c_out + (m0 + r) * e.H + ucol + c * 4) =
                *reinterpret_cast<const uint4*>(sb + 6 * 2048 + r * 128 + ((c ^ (r & 7)) << 4));
        }
      }
      named_bar_sync(2 + q, 64);  // the staging buffer is rewritten by the next pass / step
    }
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
  // Same thread-per-row problem as the forward epilogue, on the load side too: per step a thread read the four gate
  // planes, c_{t-1} and c_t of its row in 16-byte pieces 1-4 KB apart (40 uncoalesced loads) and wrote dG the same way.
  // None of those inputs depends on the recurrence, so the two warps that share 32 rows fetch them as whole row segments
  // into shared memory BEFORE the step's accumulator is awaited (the loads hide under the MMA), each thread then reads
  // its own row from shared memory, and dG leaves through the same staging area as whole row segments.
  static constexpr int kPassBytes = 4 * 2048 + 2 * 4096;  // four bf16 gate planes [32 x 64 B] + c_prev, c_cur [32 x 128 B]
  static constexpr int kPairBytes = 2 * kPassBytes;       // both 32-unit passes
  static constexpr int kSmemBytes = 4 * kPairBytes;       // 128 KB
  struct State {
    float dh[32], dc[32];  // pass-through dh / dc of this thread's row: index ps * 16 + j <-> unit ps * 32 + grp * 16 + j
  };
  __device__ __forceinline__ void init(State& s, int m_blk, int n_blk, int row, int grp) const {
    const int m = m_blk * kBM + row;
#pragma unroll
    for (int j = 0; j < 32; ++j) s.dh[j] = s.dc[j] = 0.f;
    if (m < c.N) {
#pragma unroll
      for (int ps = 0; ps < 2; ++ps) {
        const long long o = (long long)m * c.H + n_blk * 64 + ps * 32 + grp * 16;
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          *reinterpret_cast<float4*>(s.dh + ps * 16 + j) = *reinterpret_cast<const float4*>(c.dh_carry + o + j);
          *reinterpret_cast<float4*>(s.dc + ps * 16 + j) = *reinterpret_cast<const float4*>(c.dc_carry + o + j);
        }
      }
    }
  }
  __device__ __forceinline__ void finish(State& s, int m_blk, int n_blk, int row, int grp) const {
    const int m = m_blk * kBM + row;
    if (m < c.N) {
#pragma unroll
      for (int ps = 0; ps < 2; ++ps) {
        const long long o = (long long)m * c.H + n_blk * 64 + ps * 32 + grp * 16;
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          *reinterpret_cast<float4*>(c.dh_carry + o + j) = *reinterpret_cast<float4*>(s.dh + ps * 16 + j);
          *reinterpret_cast<float4*>(c.dc_carry + o + j) = *reinterpret_cast<float4*>(s.dc + ps * 16 + j);
        }
      }
    }
  }
  __device__ __forceinline__ void prefetch(int st, int m_blk, int n_blk, int row, int grp, uint8_t* stage_smem) const {
    const size_t nh = (size_t)c.N * c.H;
    const __nv_bfloat16* gates = G + (size_t)st * nh * 4;
    const float* c_prev = Cs + (size_t)st * nh;
    const float* c_cur = Cs + (size_t)(st + 1) * nh;
    const int lane = row & 31, q = row >> 5;
    uint8_t* sb = stage_smem + q * kPairBytes;
    const int pt = grp * 32 + lane;
    const long long m0 = (long long)m_blk * kBM + q * 32;
#pragma unroll
    for (int ps = 0; ps < 2; ++ps) {
      uint8_t* pb = sb + ps * kPassBytes;
      const int ucol = n_blk * 64 + ps * 32;
      uint4 v[8];
#pragma unroll
      for (int i = 0; i < 2; ++i) {  // gate planes: 32 rows x 4 chunks of 16 bytes each
        const int e16 = pt + 64 * i;
        const int r = e16 >> 2, ch = e16 & 3;
        const bool ok = m0 + r < c.N;
        const __nv_bfloat16* src = gates + (m0 + r) * 4 * c.H + ucol + ch * 8;
#pragma unroll
        for (int gg = 0; gg < 4; ++gg) v[i * 4 + gg] = ok ? *reinterpret_cast<const uint4*>(src + (long long)gg * c.H) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int e16 = pt + 64 * i;
        const int r = e16 >> 2, ch = e16 & 3;
        uint8_t* dst = pb + r * 64 + ((ch ^ ((r >> 1) & 3)) << 4);
#pragma unroll
        for (int gg = 0; gg < 4; ++gg) *reinterpret_cast<uint4*>(dst + gg * 2048) = v[i * 4 + gg];
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {  // cell states: 32 rows x 8 chunks
        const int e16 = pt + 64 * i;
        const int r = e16 >> 3, ch = e16 & 7;
        const bool ok = m0 + r < c.N;
        const long long o = (m0 + r) * c.H + ucol + ch * 4;
        v[i] = ok ? *reinterpret_cast<const uint4*>(c_prev + o) : make_uint4(0, 0, 0, 0);
        v[4 + i] = ok ? *reinterpret_cast<const uint4*>(c_cur + o) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int e16 = pt + 64 * i;
        const int r = e16 >> 3, ch = e16 & 7;
        uint8_t* dst = pb + 8192 + r * 128 + ((ch ^ (r & 7)) << 4);
        *reinterpret_cast<uint4*>(dst) = v[i];
        *reinterpret_cast<uint4*>(dst + 4096) = v[4 + i];
      }
    }
    named_bar_sync(2 + q, 64);
  }
  __device__ __forceinline__ void step(uint32_t taddr, int st, int m_blk, int n_blk, int row, int grp, State& state,
                                       uint8_t* stage_smem) const {
    const int t = st - pre;
    const size_t nh = (size_t)c.N * c.H;
    const float* dout = (d_out != nullptr && t >= 0) ? d_out + (size_t)t * nh : nullptr;
    const float* keep = (out_keep != nullptr && t >= 0) ? out_keep + (size_t)t * c.H : nullptr;
    __nv_bfloat16* dgates = dG + (size_t)st * nh * 4;
    const int lane = row & 31, q = row >> 5;
    const int m = m_blk * kBM + row;
    const bool row_ok = m < c.N;
    const bool live = row_ok && (t < 0 || lengths == nullptr || t < lengths[m]);
    uint8_t* sb = stage_smem + q * kPairBytes;
    const int pt = grp * 32 + lane;
    const long long m0 = (long long)m_blk * kBM + q * 32;
#pragma unroll 1
    for (int ps = 0; ps < 2; ++ps) {
      uint8_t* pb = sb + ps * kPassBytes;
      const int ul = ps * 32 + grp * 16;
      const int u0 = n_blk * 64 + ul;
      float* dhc = state.dh + ps * 16;
      float* dcc = state.dc + ps * 16;
      float acc[16];
      __syncwarp();
      tmem_ld16(taddr + ul, acc);
      uint32_t dgp[4][8];
      // this thread's row of the staged planes (same swizzles as the forward epilogue's staging)
      const int sw = (lane >> 1) & 3;
      const uint8_t* rb = pb + lane * 64;
      const int c0 = ((grp * 2) ^ sw) << 4, c1 = ((grp * 2 + 1) ^ sw) << 4;
      tmem_ld_wait();
      if (live) {
        float dov[16];
        if (dout != nullptr) {
          const long long o = (long long)m * c.H + u0;
#pragma unroll
          for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(dov + j) = *reinterpret_cast<const float4*>(dout + o + j);
          if (keep != nullptr) {
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              const float4 k = *reinterpret_cast<const float4*>(keep + (long long)m * c.out_keep_ld + u0 + j);
              dov[j] *= k.x * c.inv_keep; dov[j + 1] *= k.y * c.inv_keep; dov[j + 2] *= k.z * c.inv_keep; dov[j + 3] *= k.w * c.inv_keep;
            }
          }
        }
        float gv[4][16], cp[16], cc[16];
#pragma unroll
        for (int gg = 0; gg < 4; ++gg) {
          const uint4 a = *reinterpret_cast<const uint4*>(rb + gg * 2048 + c0), b = *reinterpret_cast<const uint4*>(rb + gg * 2048 + c1);
          const __nv_bfloat162* ha = reinterpret_cast<const __nv_bfloat162*>(&a);
          const __nv_bfloat162* hb = reinterpret_cast<const __nv_bfloat162*>(&b);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 fa = __bfloat1622float2(ha[j]), fb = __bfloat1622float2(hb[j]);
            gv[gg][2 * j] = fa.x; gv[gg][2 * j + 1] = fa.y;
            gv[gg][8 + 2 * j] = fb.x; gv[gg][8 + 2 * j + 1] = fb.y;
          }
        }
        const uint8_t* cb = pb + 8192 + lane * 128;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int off = ((grp * 4 + j) ^ (lane & 7)) << 4;
          *reinterpret_cast<uint4*>(cp + 4 * j) = *reinterpret_cast<const uint4*>(cb + off);
          *reinterpret_cast<uint4*>(cc + 4 * j) = *reinterpret_cast<const uint4*>(cb + 4096 + off);
        }
        float dg[4][16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float gi = gv[0][j], gj = gv[1][j], gf = gv[2][j], go = gv[3][j];
          float dh = acc[j] + dhc[j];
          if (dout != nullptr) dh += dov[j];
          const float tc = tanh_fast(cc[j]);
          const float dc = dcc[j] + dh * go * (1.f - tc * tc);
          dg[3][j] = dh * tc * go * (1.f - go);
          dg[0][j] = dc * gj * gi * (1.f - gi);
          dg[1][j] = dc * gi * (1.f - gj * gj);
          dg[2][j] = dc * cp[j] * gf * (1.f - gf);
          dcc[j] = dc * gf;
          dhc[j] = 0.f;
        }
#pragma unroll
        for (int gg = 0; gg < 4; ++gg)
#pragma unroll
          for (int j = 0; j < 8; ++j) dgp[gg][j] = pack_bf16(dg[gg][2 * j], dg[gg][2 * j + 1]);
      } else {
        if (row_ok) {
#pragma unroll
          for (int j = 0; j < 16; ++j) dhc[j] += acc[j];
        }
#pragma unroll
        for (int gg = 0; gg < 4; ++gg)
#pragma unroll
          for (int j = 0; j < 8; ++j) dgp[gg][j] = 0u;
      }
      named_bar_sync(2 + q, 64);  // both warps have read this pass's inputs: the gate planes become the dG staging
      {
        uint8_t* wb = pb + lane * 64;
#pragma unroll
        for (int gg = 0; gg < 4; ++gg) {
          *reinterpret_cast<uint4*>(wb + gg * 2048 + c0) = *reinterpret_cast<uint4*>(dgp[gg]);
          *reinterpret_cast<uint4*>(wb + gg * 2048 + c1) = *reinterpret_cast<uint4*>(dgp[gg] + 4);
        }
      }
      named_bar_sync(2 + q, 64);
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int e16 = pt + 64 * i;
        const int r = e16 >> 2, ch = e16 & 3;
        if (m0 + r < c.N) {
          const uint8_t* src = pb + r * 64 + ((ch ^ ((r >> 1) & 3)) << 4);
          __nv_bfloat16* dst = dgates + (m0 + r) * 4 * c.H + n_blk * 64 + ps * 32 + ch * 8;
#pragma unroll
          for (int gg = 0; gg < 4; ++gg) *reinterpret_cast<uint4*>(dst + (long long)gg * c.H) = *reinterpret_cast<const uint4*>(src + gg * 2048);
        }
      }
    }
    named_bar_sync(2 + q, 64);  // the staging area is refilled by the next step's prefetch
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
  core.stages = gemm_pick_stages(core.bn, Epi::kSmemBytes);
  const int smem = gemm_smem_bytes(core.bn, core.stages, Epi::kSmemBytes);
  VC_CUDA(cudaMemsetAsync(core.flags, 0, (size_t)core.m_tiles * (core.iters + 1) * sizeof(int), stream));
  static const bool trace_on = [] { const char* e = getenv("VC_LSTM_TRACE"); return e && e[0] == '1'; }();
  static int traced = 0;
  core.trace = nullptr;
  if (trace_on && traced < 8) {
    VC_CUDA(cudaMalloc((void**)&core.trace, (size_t)core.iters * 4 * sizeof(long long)));
    VC_CUDA(cudaMemsetAsync(core.trace, 0, (size_t)core.iters * 4 * sizeof(long long), stream));
  }
  {
    ProfScope ps(stream, tag);
    lstm_seq_kernel<Epi><<<core.m_tiles * core.n_tiles, kGemmThreads, smem, stream>>>(tmX, tmR, tmB, core, epi);
  }
  VC_CUDA(cudaGetLastError());
  if (core.trace) {  // debug only: synchronous dump of CTA 0's per-step timeline (cycles)
    std::vector<long long> h((size_t)core.iters * 4);
    VC_CUDA(cudaStreamSynchronize(stream));
    VC_CUDA(cudaMemcpy(h.data(), core.trace, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    cudaFree(core.trace);
    fprintf(stderr, "[lstm trace %s #%d] grid %dx%d iters %d kx %d kh %d bn %d: it: flag->acc  acc->epi  epi->pub  pub->nextflag (cycles)\n", tag,
            traced, core.m_tiles, core.n_tiles, core.iters, core.kx_blocks, core.kh_blocks, core.bn);
    for (int it = 1; it + 1 < core.iters; ++it)
      fprintf(stderr, "  %2d: %6lld %6lld %6lld %6lld\n", it, h[it * 4 + 1] - h[it * 4 + 0], h[it * 4 + 2] - h[it * 4 + 1],
              h[it * 4 + 3] - h[it * 4 + 2], h[(it + 1) * 4 + 0] - h[it * 4 + 3]);
    ++traced;
  }
  return VC_OK;
}

bool lstm_seq_applicable(int N, int H, int steps) {
  const long long ctas = (long long)((N + kBM - 1) / kBM) * (H / 64);
  return H % 64 == 0 && ctas <= num_sms() && steps >= 1;
}

int lstm_seq_flag_count(int N, int steps) { return ((N + kBM - 1) / kBM) * (steps + 2); }

// VC_LSTM_SEQ=1 selects the first form below (kept for A/B timing); the default is lstm_seq2.cu
static bool use_seq2() {
  static const bool v2 = [] { const char* e = getenv("VC_LSTM_SEQ"); return !(e && e[0] == '1'); }();
  return v2;
}

int lstm_fwd_seq(cudaStream_t stream, const LstmSeqFwdArgs& a) {
  if (use_seq2()) return lstm_fwd_seq2(stream, a);
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
  if (use_seq2()) return lstm_bwd_seq2(stream, a);
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
