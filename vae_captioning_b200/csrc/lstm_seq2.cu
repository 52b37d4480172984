// Whole-sequence LSTM kernels, second form: same grid, operand ring, tcgen05 mainloop and per-(row tile, step)
// release / acquire counters as lstm_seq.cu, but the step's critical path is cut down to what the NEXT step needs.
//
// Timeline of the first form, CTA 0, cycles per step (VC_LSTM_TRACE=1, N = 1280): forward 5.5 k operands + MMA,
// 11.5 k epilogue, 2.2 k publish, 2.5 k waiting for the peers = 21.7 k; backward 14.3 k / 11.4 k / 4.5 k / 3.7 k = 34 k.
// The epilogue wrote every plane of the step (h, emitted output, four gate planes, cell state; or four gate-gradient
// planes) with per-thread stores through a staging area, met at barriers, then ONE thread fenced for all 256 -- and only
// then did the peers learn that the recurrent operand was there. Here:
//   * the recurrent operand (forward: the 128 x 64 h tile; backward: four 128 x 64 gate-gradient tiles) is built in
//     shared memory in the 128-byte-swizzled layout of a TMA box, and a fourth warp (the I/O warp) writes it with bulk
//     tensor stores, waits for their completion and releases the step counter -- no CTA-wide barrier, no per-thread
//     global stores, no membar over 256 threads' traffic in front of the flag;
//   * forward: everything BPTT needs later (gates, cell state, emitted output) leaves after the h tile is handed over;
//   * backward: the step's inputs (activated gates, c_{t-1}, c_t) are fetched by the I/O warp with TMA into the same
//     tiles while the recurrent GEMM runs (they do not depend on it); c_t of one step is c_{t-1} of the one processed
//     before, so only one cell-state tile is loaded per step; gate gradients overwrite the gate tiles in place;
//   * 3-D tensor maps {column, row, step} clip the rows of a ragged last row tile instead of per-thread guards.
// Reference semantics unchanged: tf.nn.dynamic_rnn over LSTMCell (vae_model/encoder.py:38-58, decoder.py:91-121).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "ops.h"
#include "lstm_epi.cuh"

namespace vc {

constexpr int kBarPub = 6;    // 256 epilogue threads arrive, the I/O warp waits: the step's tiles are complete
constexpr int kBarFree = 7;   // the I/O warp arrives, the epilogue threads wait: the tiles may be rewritten
constexpr int kIoThreads = 288;

struct Seq2Core {
  int m_tiles, n_tiles, iters;
  int kx_blocks, kh_blocks, bn, stages;
  int N, reverse;
  int* flags;        // [m_tiles, iters + 1], zeroed by the launcher
  long long* trace;  // VC_LSTM_TRACE=1: [iters][4] clock64() of CTA 0: flag seen, accumulator complete, tiles complete, published
};

// Epi: struct State; kTileBytes (TMA tiles, 1024-byte aligned), kAuxBytes (other staging), kUsesFree;
//   init / finish (State&, m_blk, n_blk, row, grp); prefetch(State&, st, m_blk, n_blk, row, grp)
//   step(taddr, it, st, m_blk, n_blk, row, grp, State&, tiles, aux, in_full)  -- ends with the kBarPub arrival
//   io_begin(tiles, in_full, maps, st0, m_blk, n_blk), io_publish(tiles, maps, st, m_blk, n_blk),
//   io_next(tiles, in_full, maps, it_next, st_next, m_blk, n_blk)            -- lane 0 of the I/O warp
struct IoMaps {
  const CUtensorMap *m0, *m1, *m2, *m3;
};

template <class Epi>
__global__ void __launch_bounds__(kGemmThreads, 1)
lstm_seq2_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmR,
                 const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmIO0,
                 const __grid_constant__ CUtensorMap tmIO1, const __grid_constant__ CUtensorMap tmIO2,
                 const __grid_constant__ CUtensorMap tmIO3, const Seq2Core g,
                 const __grid_constant__ Epi epi) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int stage_bytes = gemm_stage_bytes(g.bn);
  uint8_t* tiles = smem + g.stages * stage_bytes;  // stage sizes are multiples of 1024: still 1024-aligned
  uint8_t* aux = tiles + Epi::kTileBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(aux + Epi::kAuxBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kMaxStages;
  uint64_t* tfull = bars + 2 * kMaxStages;
  uint64_t* tempty = bars + 2 * kMaxStages + 2;
  uint64_t* in_full = bars + 2 * kMaxStages + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 6);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_blk = blockIdx.x % g.m_tiles;
  const int n_blk = blockIdx.x / g.m_tiles;
  const int kb_total = g.kx_blocks + g.kh_blocks;
  const IoMaps maps{&tmIO0, &tmIO1, &tmIO2, &tmIO3};

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmR);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmIO0);
    tma_prefetch_desc(&tmIO1);
    tma_prefetch_desc(&tmIO2);
    tma_prefetch_desc(&tmIO3);
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
    mbar_init(in_full, 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {  // TMA producer of the GEMM operands: identical to lstm_seq.cu
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < g.iters; ++it) {
        const int st = g.reverse ? g.iters - 1 - it : it;
        const int row0 = st * g.N + m_blk * kBM;
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
            if (g.trace && blockIdx.x == 0) g.trace[it * 16 + 0] = clock64();
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
        if (!ready) {
          while (ld_relaxed_gpu(f) < g.n_tiles) {
          }
          fence_acq_rel_gpu();
          fence_proxy_async_all();
          if (g.trace && blockIdx.x == 0) g.trace[it * 16 + 0] = clock64();
          for (int i = 0; i < pend; ++i) {
            const int ps = (pend_stage0 + i) % g.stages;
            tma_load_2d(smem + ps * stage_bytes, &tmR, &full[ps], (pend_kb0 + i - g.kx_blocks) * kBK, rrow);
          }
        }
      }
    }
  } else if (warp == 1) {
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
  } else if (warp == 3) {
    // I/O warp: hands the step's tiles to the peers (bulk tensor stores, completion, release of the step counter) and,
    // backward, fetches the next step's epilogue inputs
    if (lane == 0) epi.io_begin(tiles, in_full, maps, g.reverse ? g.iters - 1 : 0, m_blk, n_blk);
    __syncwarp();
    for (int it = 0; it < g.iters; ++it) {
      const int st = g.reverse ? g.iters - 1 - it : it;
      named_bar_sync(kBarPub, kIoThreads);
      if (lane == 0) {
        const bool tr = g.trace && blockIdx.x == 0;
        if (tr) g.trace[it * 16 + 4] = clock64();
        epi.io_publish(tiles, maps, st, m_blk, n_blk);       // the recurrent operand: its own bulk group
        bulk_commit();
        bulk_wait_all();                                      // its writes are complete
        if (tr) g.trace[it * 16 + 5] = clock64();
        if (tr) g.trace[it * 16 + 6] = clock64();
        // The group's completion makes its writes visible; the release reduction orders them before the counter and the
        // consumer pairs it with fence.acq_rel + fence.proxy.async in front of its TMA loads. (A fence.proxy.async here
        // waited for the SECOND bulk group as well: 4 k cycles; a fence.acq_rel in front of the reduction another 1.5 k.)
        red_release_gpu_add(g.flags + m_blk * (g.iters + 1) + it + 1, 1);
        if (tr) g.trace[it * 16 + 3] = clock64();
        // everything else the step produced drains under the next step (issued AFTER the release: the release waits for
        // every write this thread has in flight, bulk stores included -- 4 k cycles when they were issued in front of it)
        epi.io_publish_rest(tiles, maps, st, m_blk, n_blk);
        bulk_commit();
        bulk_wait_read<0>();                                  // shared memory of every tile has been read
        if (it + 1 < g.iters) epi.io_next(tiles, in_full, maps, it + 1, g.reverse ? st - 1 : st + 1, m_blk, n_blk);
      }
      __syncwarp();
      if (Epi::kUsesFree && it + 1 < g.iters) named_bar_arrive(kBarFree, kIoThreads);
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    const int grp = (warp - 4) >> 2;
    typename Epi::State state;
    epi.setup(aux, n_blk, threadIdx.x - 128);
    named_bar_sync(1, 256);
    epi.init(state, m_blk, n_blk, q * 32 + lane, grp);
    for (int it = 0; it < g.iters; ++it) {
      const int st = g.reverse ? g.iters - 1 - it : it;
      const int acc = it & 1;
      epi.prefetch(state, st, m_blk, n_blk, q * 32 + lane, grp);
      mbar_wait(&tfull[acc], (it >> 1) & 1);
      tc_fence_after();
      const bool tr = g.trace && blockIdx.x == 0 && warp == 4 && lane == 0;
      if (tr) g.trace[it * 16 + 1] = clock64();
      epi.step(tmem_base + acc * 256 + (uint32_t(q * 32) << 16), it, st, m_blk, n_blk, q * 32 + lane, grp, state, tiles, aux,
               in_full, tr ? g.trace + it * 16 : nullptr);
      tc_fence_before();
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
// Byte offset inside a 1024-byte-aligned TMA tile with 128-byte swizzle (Swizzle<3,4,3> on the address: the 16-byte chunk
// index, bits 4-6, is XORed with bits 7-9), for boxes whose rows are 128 bytes. (A 64-byte-row box under the 128-byte
// mode is NOT this function of the dense offset -- measured: wrong data -- so such boxes use the 64-byte mode below.)
__device__ __forceinline__ uint32_t swz128(uint32_t off) { return off ^ (((off >> 7) & 7u) << 4); }
// 64-byte swizzle (boxes with 64-byte rows): Swizzle<2,4,3>, chunk bits 4-5 XOR address bits 7-8
__device__ __forceinline__ uint32_t swz64(uint32_t off) { return off ^ (((off >> 7) & 3u) << 4); }
template <int kRowBytes>
__device__ __forceinline__ uint32_t swz_row(uint32_t off) { return kRowBytes == 64 ? swz64(off) : swz128(off); }

// Forward, kU hidden units per CTA (64: 8 CTAs per row tile, 32: 16 CTAs per row tile when the grid still fits the SMs).
// Every plane the step produces is a tile in TMA box layout: h [128 x kU bf16] (the recurrent operand, handed over first),
// the emitted output, the four activated-gate planes and the cell state ([128 x 32 fp32] per 32 units) -- 128 KB at
// kU = 64, 64 KB at kU = 32; the epilogue threads only write shared memory, the I/O warp stores the tiles (h as its own
// bulk group, whose completion releases the step counter; the planes BPTT reads later as a second group that drains
// under the next step).
template <int kU>
struct Seq2FwdEpi {
  const float* bias;       // [4H] natural gate order
  const int* lengths;      // applied to caption steps only (t >= 0)
  const float* out_keep;   // [N, T, H] (nullable)
  long long out_keep_ld;
  float inv_keep;
  int pre, T, N, H;
  int has_out;             // the emitted outputs of the caption steps are wanted (decoder)
  static constexpr int kUnits = kU;
  static constexpr int kPasses = kU / 32;
  static constexpr int kRow = kU * 2;            // bytes per row of a bf16 plane tile
  static constexpr int kPlane = kBM * kRow;
  static constexpr int kOutOff = kPlane, kGateOff = 2 * kPlane, kCellOff = 6 * kPlane;
  static constexpr int kTileBytes = 6 * kPlane + kPasses * kABytes;
  static constexpr int kBiasOff = 0;             // aux: this CTA's bias slice [4 gates][kU units] fp32
  static constexpr int kAuxBytes = 1024;
  static constexpr bool kUsesFree = true;
  struct State {
    float c[kU / 2];        // cell state of this thread's row x 16 hidden units per pass
    uint32_t h[kU / 4];     // the same units of h as packed bf16 (state copied through past the end of a caption)
    int len;
  };
  __device__ __forceinline__ void setup(uint8_t* aux, int n_blk, int tid) const {
    if (tid < 4 * kU) reinterpret_cast<float*>(aux + kBiasOff)[tid] = bias[(tid / kU) * H + n_blk * kU + (tid % kU)];
  }
  __device__ __forceinline__ void init(State& s, int m_blk, int, int row, int) const {
#pragma unroll
    for (int j = 0; j < kU / 2; ++j) s.c[j] = 0.f;  // cell_0.zero_state
#pragma unroll
    for (int j = 0; j < kU / 4; ++j) s.h[j] = 0u;
    const int m = m_blk * kBM + row;
    s.len = (lengths != nullptr && m < N) ? lengths[m] : 0x7fffffff;
  }
  __device__ __forceinline__ void finish(State&, int, int, int, int) const {}
  __device__ __forceinline__ void prefetch(State&, int, int, int, int, int) const {}
  __device__ __forceinline__ void io_begin(uint8_t*, uint64_t*, const IoMaps&, int, int, int) const {}
  __device__ __forceinline__ void io_next(uint8_t*, uint64_t*, const IoMaps&, int, int, int, int) const {}
  // maps: m0 = Hs {H, N, steps+1}, m1 = out {H, N, T}, m2 = G {4H, N, steps} (boxes kU wide), m3 = Cs fp32 {H, N, steps+1}
  __device__ __forceinline__ void io_publish(uint8_t* tiles, const IoMaps& maps, int st, int m_blk, int n_blk) const {
    tma_store_3d(maps.m0, tiles, n_blk * kU, m_blk * kBM, st + 1);  // Hs[st + 1], rows past N dropped
  }
  __device__ __forceinline__ void io_publish_rest(uint8_t* tiles, const IoMaps& maps, int st, int m_blk, int n_blk) const {
    const int t = st - pre;
    if (has_out && t >= 0) tma_store_3d(maps.m1, tiles + kOutOff, n_blk * kU, m_blk * kBM, t);
#pragma unroll
    for (int gg = 0; gg < 4; ++gg)
      tma_store_3d(maps.m2, tiles + kGateOff + gg * kPlane, gg * H + n_blk * kU, m_blk * kBM, st);
#pragma unroll
    for (int hf = 0; hf < kPasses; ++hf)
      tma_store_3d(maps.m3, tiles + kCellOff + hf * kABytes, n_blk * kU + hf * 32, m_blk * kBM, st + 1);
  }
  __device__ __forceinline__ void step(uint32_t taddr, int it, int st, int m_blk, int n_blk, int row, int grp, State& state,
                                       uint8_t* tiles, uint8_t* aux, uint64_t*, long long* trace) const {
    const int t = st - pre;
    const bool want_out = has_out && t >= 0;
    const float* keep = (out_keep != nullptr && t >= 0) ? out_keep + (size_t)t * H : nullptr;
    const int m = m_blk * kBM + row;
    const bool row_ok = m < N;
    const bool live = row_ok && (t < 0 || t < state.len);
    if (it > 0) named_bar_sync(kBarFree, kIoThreads);  // the previous step's tiles have left shared memory
#pragma unroll
    for (int ps = 0; ps < kPasses; ++ps) {
      const int ul = ps * 32 + grp * 16;  // this thread's 16 units inside the CTA's kU
      float* cp = state.c + ps * 16;
      uint32_t* hs = state.h + ps * 8;
      float gi[16], gj[16], gf[16], go[16];
      __syncwarp();
      tmem_ld16(taddr + 0 * kU + ul, gi);
      tmem_ld16(taddr + 1 * kU + ul, gj);
      tmem_ld16(taddr + 2 * kU + ul, gf);
      tmem_ld16(taddr + 3 * kU + ul, go);
      const int u0 = n_blk * kU + ul;
      const float* bs = reinterpret_cast<const float*>(aux + kBiasOff) + ul;  // warp-wide broadcast reads
      // this thread's two 16-byte chunks of a bf16 plane tile
      const uint32_t c0 = swz_row<kRow>(row * kRow + ul * 2), c1 = swz_row<kRow>(row * kRow + ul * 2 + 16);
      tmem_ld_wait();
      if (trace) trace[7 + ps * 4] = clock64();
      if (live) {
        float hn[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float i_ = sigmoid_fast(gi[j] + bs[j]);
          const float j_ = tanh_fast(gj[j] + bs[kU + j]);
          const float f_ = sigmoid_fast(gf[j] + bs[2 * kU + j] + 1.0f);
          const float o_ = sigmoid_fast(go[j] + bs[3 * kU + j]);
          const float cn = f_ * cp[j] + i_ * j_;
          hn[j] = o_ * tanh_fast(cn);
          cp[j] = cn;
          gi[j] = i_; gj[j] = j_; gf[j] = f_; go[j] = o_;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) hs[j] = pack_bf16(hn[2 * j], hn[2 * j + 1]);
      }
      // the recurrent operand first
      *reinterpret_cast<uint4*>(tiles + c0) = *reinterpret_cast<uint4*>(hs);
      *reinterpret_cast<uint4*>(tiles + c1) = *reinterpret_cast<uint4*>(hs + 4);
      if (trace) trace[8 + ps * 4] = clock64();
      // what BPTT (and the vocabulary projection) read later
      uint32_t pk[8];
      if (want_out) {
        if (live && keep != nullptr) {  // the emitted value is the bf16 state scaled by the keep mask (the state itself is untouched)
          const float* kp = keep + (long long)m * out_keep_ld + u0;
          const __nv_bfloat162* hb = reinterpret_cast<const __nv_bfloat162*>(hs);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float2 v = __bfloat1622float2(hb[j]);
            pk[j] = pack_bf16(v.x * kp[2 * j] * inv_keep, v.y * kp[2 * j + 1] * inv_keep);
          }
        } else {
          // past the end of the sequence the emitted output is zero while the state is copied through
#pragma unroll
          for (int j = 0; j < 8; ++j) pk[j] = live ? hs[j] : 0u;
        }
        *reinterpret_cast<uint4*>(tiles + kOutOff + c0) = *reinterpret_cast<uint4*>(pk);
        *reinterpret_cast<uint4*>(tiles + kOutOff + c1) = *reinterpret_cast<uint4*>(pk + 4);
      }
      {
        float* gsrc[4] = {gi, gj, gf, go};
#pragma unroll
        for (int gg = 0; gg < 4; ++gg) {  // gate planes of rows past their end are never read by BPTT: zeros
#pragma unroll
          for (int j = 0; j < 8; ++j) pk[j] = live ? pack_bf16(gsrc[gg][2 * j], gsrc[gg][2 * j + 1]) : 0u;
          *reinterpret_cast<uint4*>(tiles + kGateOff + gg * kPlane + c0) = *reinterpret_cast<uint4*>(pk);
          *reinterpret_cast<uint4*>(tiles + kGateOff + gg * kPlane + c1) = *reinterpret_cast<uint4*>(pk + 4);
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j)  // cell state: [128 x 32 fp32] tile of this pass, chunks grp * 4 + j
        *reinterpret_cast<uint4*>(tiles + kCellOff + ps * kABytes + swz128(row * 128 + (grp * 4 + j) * 16)) =
            *reinterpret_cast<uint4*>(cp + 4 * j);
      if (trace) trace[9 + ps * 4] = clock64();
    }
    fence_proxy_async();  // generic-proxy writes of the tiles -> visible to the bulk tensor stores
    named_bar_arrive(kBarPub, kIoThreads);
    if (trace) trace[2] = clock64();
  }
};

// ------------------------------------------------------------------------------------------
// Backward (steps steps-2 .. 0), kU hidden units per CTA. Tiles: four gate tiles [128 x kU bf16] (activated gates in,
// gate gradients out, in place) + two cell-state buffers of [128 x 32 fp32] tiles, all in TMA box layout.
template <int kU>
struct Seq2BwdEpi {
  const float* d_out;        // [T, N, H] (nullable)
  const float* out_keep;     // [N, T, H] (nullable)
  long long out_keep_ld;
  float inv_keep;
  float* dh_carry;
  float* dc_carry;
  const int* lengths;
  int pre, N, H;
  static constexpr int kUnits = kU;
  static constexpr int kPasses = kU / 32;
  static constexpr int kRow = kU * 2;
  static constexpr int kPlane = kBM * kRow;
  static constexpr uint32_t kGateBytes = 4 * kPlane, kCellBytes = kPasses * kABytes;
  static constexpr int kTileBytes = kGateBytes + 2 * kCellBytes;
  static constexpr int kAuxBytes = 0;
  static constexpr bool kUsesFree = false;
  __device__ __forceinline__ void setup(uint8_t*, int, int) const {}
  struct State {
    float dh[kU / 2], dc[kU / 2];  // pass-through dh / dc of this thread's row: index ps * 16 + j <-> unit ps * 32 + grp * 16 + j
    float dov[kU / 2];             // d(emitted output), fetched before the accumulator is awaited
    int len;
  };
  __device__ __forceinline__ void init(State& s, int m_blk, int n_blk, int row, int grp) const {
    const int m = m_blk * kBM + row;
#pragma unroll
    for (int j = 0; j < kU / 2; ++j) s.dh[j] = s.dc[j] = 0.f;
    s.len = (lengths != nullptr && m < N) ? lengths[m] : 0x7fffffff;
    if (m < N) {
#pragma unroll
      for (int ps = 0; ps < kPasses; ++ps) {
        const long long o = (long long)m * H + n_blk * kU + ps * 32 + grp * 16;
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          *reinterpret_cast<float4*>(s.dh + ps * 16 + j) = *reinterpret_cast<const float4*>(dh_carry + o + j);
          *reinterpret_cast<float4*>(s.dc + ps * 16 + j) = *reinterpret_cast<const float4*>(dc_carry + o + j);
        }
      }
    }
  }
  __device__ __forceinline__ void finish(State& s, int m_blk, int n_blk, int row, int grp) const {
    const int m = m_blk * kBM + row;
    if (m < N) {
#pragma unroll
      for (int ps = 0; ps < kPasses; ++ps) {
        const long long o = (long long)m * H + n_blk * kU + ps * 32 + grp * 16;
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          *reinterpret_cast<float4*>(dh_carry + o + j) = *reinterpret_cast<float4*>(s.dh + ps * 16 + j);
          *reinterpret_cast<float4*>(dc_carry + o + j) = *reinterpret_cast<float4*>(s.dc + ps * 16 + j);
        }
      }
    }
  }
  __device__ __forceinline__ void load_dout(float* dov, int st, int m, int u0, bool live) const {
    const int t = st - pre;
    if (d_out == nullptr || t < 0 || !live) {
#pragma unroll
      for (int j = 0; j < 16; ++j) dov[j] = 0.f;
      return;
    }
    const long long o = (long long)t * N * H + (long long)m * H + u0;
#pragma unroll
    for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(dov + j) = *reinterpret_cast<const float4*>(d_out + o + j);
    if (out_keep != nullptr) {
      const float* kp = out_keep + (long long)t * H + (long long)m * out_keep_ld + u0;
#pragma unroll
      for (int j = 0; j < 16; j += 4) {
        const float4 k = *reinterpret_cast<const float4*>(kp + j);
        dov[j] *= k.x * inv_keep; dov[j + 1] *= k.y * inv_keep; dov[j + 2] *= k.z * inv_keep; dov[j + 3] *= k.w * inv_keep;
      }
    }
  }
  __device__ __forceinline__ bool is_live(const State& s, int st, int m) const {
    const int t = st - pre;
    return m < N && (t < 0 || t < s.len);
  }
  __device__ __forceinline__ void prefetch(State& s, int st, int m_blk, int n_blk, int row, int grp) const {
    const int m = m_blk * kBM + row;
    const bool live = is_live(s, st, m);
#pragma unroll
    for (int ps = 0; ps < kPasses; ++ps) load_dout(s.dov + ps * 16, st, m, n_blk * kU + ps * 32 + grp * 16, live);
  }
  // I/O warp, lane 0. maps.m0 = dG {4H, N, steps} (store), m1 = G {4H, N, steps} (load), m2 = Cs fp32 {H, N, steps+1}
  __device__ __forceinline__ void load_gates(uint8_t* tiles, uint64_t* in_full, const IoMaps& maps, int st, int m_blk,
                                             int n_blk) const {
#pragma unroll
    for (int gg = 0; gg < 4; ++gg)
      tma_load_3d(tiles + gg * kPlane, maps.m1, in_full, gg * H + n_blk * kU, m_blk * kBM, st);
  }
  __device__ __forceinline__ void load_cell(uint8_t* tiles, uint64_t* in_full, const IoMaps& maps, int buf, int slot, int m_blk,
                                            int n_blk) const {
    uint8_t* cb = tiles + kGateBytes + buf * kCellBytes;
#pragma unroll
    for (int hf = 0; hf < kPasses; ++hf)
      tma_load_3d(cb + hf * kABytes, maps.m2, in_full, n_blk * kU + hf * 32, m_blk * kBM, slot);
  }
  // iteration `it` (step st): c_t (cur) sits in cell buffer it & 1, c_{t-1} (prev) in the other
  __device__ __forceinline__ void io_begin(uint8_t* tiles, uint64_t* in_full, const IoMaps& maps, int st0, int m_blk,
                                           int n_blk) const {
    mbar_expect_tx(in_full, kGateBytes + 2 * kCellBytes);
    load_gates(tiles, in_full, maps, st0, m_blk, n_blk);
    load_cell(tiles, in_full, maps, 0, st0 + 1, m_blk, n_blk);
    load_cell(tiles, in_full, maps, 1, st0, m_blk, n_blk);
  }
  __device__ __forceinline__ void io_publish(uint8_t* tiles, const IoMaps& maps, int st, int m_blk, int n_blk) const {
#pragma unroll
    for (int gg = 0; gg < 4; ++gg) tma_store_3d(maps.m0, tiles + gg * kPlane, gg * H + n_blk * kU, m_blk * kBM, st);
  }
  __device__ __forceinline__ void io_publish_rest(uint8_t*, const IoMaps&, int, int, int) const {}
  __device__ __forceinline__ void io_next(uint8_t* tiles, uint64_t* in_full, const IoMaps& maps, int it, int st, int m_blk,
                                          int n_blk) const {
    // the stores of the previous step have read the tiles (bulk_wait_read) and every epilogue thread is past its reads
    mbar_expect_tx(in_full, kGateBytes + kCellBytes);
    load_gates(tiles, in_full, maps, st, m_blk, n_blk);
    load_cell(tiles, in_full, maps, (it + 1) & 1, st, m_blk, n_blk);  // c_{t-1} of step st replaces the old c_t
  }
  __device__ __forceinline__ void step(uint32_t taddr, int it, int st, int m_blk, int n_blk, int row, int grp, State& state,
                                       uint8_t* tiles, uint8_t*, uint64_t* in_full, long long* trace) const {
    const int m = m_blk * kBM + row;
    const bool row_ok = m < N;
    const bool live = is_live(state, st, m);
    const uint8_t* c_cur = tiles + kGateBytes + (it & 1) * kCellBytes;
    const uint8_t* c_prev = tiles + kGateBytes + ((it + 1) & 1) * kCellBytes;
    mbar_wait(in_full, it & 1);  // gates / cell states of this step have landed
    if (trace) trace[7] = clock64();
#pragma unroll
    for (int ps = 0; ps < kPasses; ++ps) {
      const float* dov = state.dov + ps * 16;
      const int ul = ps * 32 + grp * 16;
      float* dhc = state.dh + ps * 16;
      float* dcc = state.dc + ps * 16;
      float acc[16];
      __syncwarp();
      tmem_ld16(taddr + ul, acc);
      const uint32_t g0 = swz_row<kRow>(row * kRow + ul * 2), g1 = swz_row<kRow>(row * kRow + ul * 2 + 16);
      uint32_t dgp[4][8];
      tmem_ld_wait();
      if (trace) trace[8 + ps * 2] = clock64();
      if (live) {
        float gv[4][16], cp[16], cc[16];
#pragma unroll
        for (int gg = 0; gg < 4; ++gg) {
          const uint4 a = *reinterpret_cast<const uint4*>(tiles + gg * kPlane + g0);
          const uint4 b = *reinterpret_cast<const uint4*>(tiles + gg * kPlane + g1);
          const __nv_bfloat162* ha = reinterpret_cast<const __nv_bfloat162*>(&a);
          const __nv_bfloat162* hb = reinterpret_cast<const __nv_bfloat162*>(&b);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 fa = __bfloat1622float2(ha[j]), fb = __bfloat1622float2(hb[j]);
            gv[gg][2 * j] = fa.x; gv[gg][2 * j + 1] = fa.y;
            gv[gg][8 + 2 * j] = fb.x; gv[gg][8 + 2 * j + 1] = fb.y;
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t off = ps * kABytes + swz128(row * 128 + (grp * 4 + j) * 16);
          *reinterpret_cast<uint4*>(cp + 4 * j) = *reinterpret_cast<const uint4*>(c_prev + off);
          *reinterpret_cast<uint4*>(cc + 4 * j) = *reinterpret_cast<const uint4*>(c_cur + off);
        }
        float dg[4][16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float gi = gv[0][j], gj = gv[1][j], gf = gv[2][j], go = gv[3][j];
          const float dh = acc[j] + dhc[j] + dov[j];
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
      // gate gradients replace the gates this thread has just read (same row, same chunks: no other thread touches them)
#pragma unroll
      for (int gg = 0; gg < 4; ++gg) {
        *reinterpret_cast<uint4*>(tiles + gg * kPlane + g0) = *reinterpret_cast<uint4*>(dgp[gg]);
        *reinterpret_cast<uint4*>(tiles + gg * kPlane + g1) = *reinterpret_cast<uint4*>(dgp[gg] + 4);
      }
      if (trace) trace[9 + ps * 2] = clock64();
    }
    fence_proxy_async();
    named_bar_arrive(kBarPub, kIoThreads);
    if (trace) trace[2] = clock64();
  }
};

// ------------------------------------------------------------------------------------------
template <class Epi>
static int launch_seq2(const CUtensorMap& tmX, const CUtensorMap& tmR, const CUtensorMap& tmB, const CUtensorMap& io0,
                       const CUtensorMap& io1, const CUtensorMap& io2, const CUtensorMap& io3, Seq2Core core, const Epi& epi, cudaStream_t stream,
                       const char* tag) {
  static bool configured = false;
  if (!configured) {
    VC_CUDA(cudaFuncSetAttribute(lstm_seq2_kernel<Epi>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  const int epi_bytes = Epi::kTileBytes + Epi::kAuxBytes;
  core.stages = gemm_pick_stages(core.bn, epi_bytes);
  if (core.stages < 2) return set_error(VC_E_ARG, "lstm_seq2: no room for the operand ring");
  const int smem = gemm_smem_bytes(core.bn, core.stages, epi_bytes);
  VC_CUDA(cudaMemsetAsync(core.flags, 0, (size_t)core.m_tiles * (core.iters + 1) * sizeof(int), stream));
  static const bool trace_on = [] { const char* e = getenv("VC_LSTM_TRACE"); return e && e[0] == '1'; }();
  static int traced = 0;
  core.trace = nullptr;
  if (trace_on && traced < 8) {
    VC_CUDA(cudaMalloc((void**)&core.trace, (size_t)core.iters * 16 * sizeof(long long)));
    VC_CUDA(cudaMemsetAsync(core.trace, 0, (size_t)core.iters * 16 * sizeof(long long), stream));
  }
  {
    ProfScope ps(stream, tag);
    lstm_seq2_kernel<Epi><<<core.m_tiles * core.n_tiles, kGemmThreads, smem, stream>>>(tmX, tmR, tmB, io0, io1, io2, io3, core, epi);
  }
  VC_CUDA(cudaGetLastError());
  if (core.trace) {  // debug only: synchronous dump of CTA 0's per-step timeline (cycles)
    std::vector<long long> h((size_t)core.iters * 16);
    VC_CUDA(cudaStreamSynchronize(stream));
    VC_CUDA(cudaMemcpy(h.data(), core.trace, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    cudaFree(core.trace);
    fprintf(stderr, "[lstm trace v2 %s #%d] grid %dx%d iters %d kx %d kh %d bn %d stages %d: it: flag->acc  acc->tiles  tiles->pub  pub->nextflag (cycles)\n",
            tag, traced, core.m_tiles, core.n_tiles, core.iters, core.kx_blocks, core.kh_blocks, core.bn, core.stages);
    for (int it = 1; it + 1 < core.iters && it < 8; ++it) {
      const long long* r = h.data() + it * 16;
      fprintf(stderr, "  %2d: %6lld %6lld %6lld %6lld | io: bar %lld store %lld fence %lld red %lld | epi (since acc):", it, r[1] - r[0],
              r[2] - r[1], r[3] - r[2], h[(it + 1) * 16 + 0] - r[3], r[4] - r[2], r[5] - r[4], r[6] - r[5], r[3] - r[6]);
      for (int k = 7; k < 15; ++k) fprintf(stderr, " %lld", r[k] ? r[k] - r[1] : 0);
      fprintf(stderr, "\n");
    }
    ++traced;
  }
  return VC_OK;
}

// Hidden units per CTA: 32 (16 CTAs per 128-row tile) while that grid is still co-resident, else 64.
int lstm_seq_units(int N, int H) {
  static const int forced = [] { const char* e = getenv("VC_LSTM_UNITS"); return e ? atoi(e) : 0; }();
  if (forced == 64) return 64;
  const long long ctas32 = (long long)((N + kBM - 1) / kBM) * (H / 32);
  return (H % 32 == 0 && ctas32 <= num_sms()) ? 32 : 64;
}

template <int kU>
static int fwd_seq2(cudaStream_t stream, const LstmSeqFwdArgs& a, const void* w_perm) {
  const long long rows = (long long)a.steps * a.N;
  CUtensorMap tmX, tmH, tmW, tmS, tmO, tmG3, tmC3;
  VC_TRY(make_tmap_2d(&tmX, a.X, a.E, rows, a.E, 64, kBM));
  VC_TRY(make_tmap_2d(&tmH, a.Hs, a.H, rows + a.N, a.H, 64, kBM));
  VC_TRY(make_tmap_2d(&tmW, w_perm, (uint64_t)a.E + a.H, 4ull * a.H, (uint64_t)a.E + a.H, 64, 4 * kU));
  VC_TRY(make_tmap_3d(&tmS, a.Hs, a.H, a.N, a.steps + 1, a.H, (uint64_t)a.N * a.H, kU, kBM, false));
  VC_TRY(make_tmap_3d(&tmG3, a.G, 4ull * a.H, a.N, a.steps, 4ull * a.H, (uint64_t)a.N * 4 * a.H, kU, kBM, false));
  VC_TRY(make_tmap_3d(&tmC3, a.Cs, a.H, a.N, a.steps + 1, a.H, (uint64_t)a.N * a.H, 32, kBM, true));
  tmO = tmS;
  if (a.out != nullptr) VC_TRY(make_tmap_3d(&tmO, a.out, a.H, a.N, a.T, a.H, (uint64_t)a.N * a.H, kU, kBM, false));
  Seq2Core core{};
  core.m_tiles = (a.N + kBM - 1) / kBM;
  core.n_tiles = a.H / kU;
  core.iters = a.steps;
  core.kx_blocks = a.E / kBK;
  core.kh_blocks = a.H / kBK;
  core.bn = 4 * kU;
  core.N = a.N;
  core.reverse = 0;
  core.flags = a.flags;
  Seq2FwdEpi<kU> epi{};
  epi.bias = a.bias;
  epi.has_out = a.out != nullptr ? 1 : 0;
  epi.lengths = a.lengths;
  epi.out_keep = a.out_keep;
  epi.out_keep_ld = (long long)a.T * a.H;
  epi.inv_keep = a.inv_keep;
  epi.pre = a.pre;
  epi.T = a.T;
  epi.N = a.N;
  epi.H = a.H;
  return launch_seq2(tmX, tmH, tmW, tmS, tmO, tmG3, tmC3, core, epi, stream, "lstm_fwd_seq");
}

int lstm_fwd_seq2(cudaStream_t stream, const LstmSeqFwdArgs& a) {
  if (a.H % 64 != 0 || a.E % kBK != 0) return set_error(VC_E_SHAPE, "LSTM sizes must be multiples of 64 (E=%d H=%d)", a.E, a.H);
  if (a.w_t_perm32 != nullptr && lstm_seq_units(a.N, a.H) == 32) return fwd_seq2<32>(stream, a, a.w_t_perm32);
  return fwd_seq2<64>(stream, a, a.w_t_perm);
}

template <int kU>
static int bwd_seq2(cudaStream_t stream, const LstmSeqBwdArgs& a) {
  const long long rows = (long long)a.steps * a.N;
  CUtensorMap tmG, tmW, tmdG3, tmG3, tmC3;
  VC_TRY(make_tmap_2d(&tmG, a.dG, 4ull * a.H, rows, 4ull * a.H, 64, kBM));
  VC_TRY(make_tmap_2d(&tmW, (const __nv_bfloat16*)a.w_nat + (long long)a.E * 4 * a.H, 4ull * a.H, a.H, 4ull * a.H, 64, kU));
  VC_TRY(make_tmap_3d(&tmdG3, a.dG, 4ull * a.H, a.N, a.steps, 4ull * a.H, (uint64_t)a.N * 4 * a.H, kU, kBM, false));
  VC_TRY(make_tmap_3d(&tmG3, a.G, 4ull * a.H, a.N, a.steps, 4ull * a.H, (uint64_t)a.N * 4 * a.H, kU, kBM, false));
  VC_TRY(make_tmap_3d(&tmC3, a.Cs, a.H, a.N, a.steps + 1, a.H, (uint64_t)a.N * a.H, 32, kBM, true));
  Seq2Core core{};
  core.m_tiles = (a.N + kBM - 1) / kBM;
  core.n_tiles = a.H / kU;
  core.iters = a.steps - 1;
  core.kx_blocks = 0;
  core.kh_blocks = 4 * a.H / kBK;
  core.bn = kU;
  core.N = a.N;
  core.reverse = 1;
  core.flags = a.flags;
  Seq2BwdEpi<kU> epi{};
  epi.d_out = a.d_out;
  epi.out_keep = a.out_keep;
  epi.out_keep_ld = (long long)a.T * a.H;
  epi.inv_keep = a.inv_keep;
  epi.dh_carry = a.dh_carry;
  epi.dc_carry = a.dc_carry;
  epi.lengths = a.lengths;
  epi.pre = a.pre;
  epi.N = a.N;
  epi.H = a.H;
  return launch_seq2(tmG, tmG, tmW, tmdG3, tmG3, tmC3, tmC3, core, epi, stream, "lstm_bwd_seq");
}

// Steps steps-2 .. 0 of BPTT's recurrent part (the last step has no recurrent input and runs in k_lstm_bwd_last).
int lstm_bwd_seq2(cudaStream_t stream, const LstmSeqBwdArgs& a) {
  if (a.H % 64 != 0) return set_error(VC_E_SHAPE, "LSTM hidden size must be a multiple of 64");
  if (a.steps < 2) return VC_OK;
  return lstm_seq_units(a.N, a.H) == 32 ? bwd_seq2<32>(stream, a) : bwd_seq2<64>(stream, a);
}

}  // namespace vc
