// Warp-specialised persistent tcgen05 GEMM mainloop shared by every dense contraction on the
// path (VGG16 convolutions as implicit GEMM, fc layers, the 4-gate LSTM projections, vocab logits,
// and their dgrad / wgrad forms).
//
//   D[128 x BN] (fp32, TMEM) = sum_k A[128 x 64] (bf16, smem) * B[BN x 64] (bf16, smem)
//
// Roles (384 threads): warp 0 = TMA producer, warp 1 = tcgen05.mma issuer (one elected lane),
// warp 2 = TMEM allocator, warps 4..7 and 8..11 = two epilogue groups (one TMEM lane quarter per
// warp). Group g drains accumulator buffer g, i.e. every second tile of the CTA, so two tile
// epilogues run concurrently (a lone warp per SM sub-partition cannot hide its own latencies).
// Three pipelines: smem full/empty ring (TMA <-> MMA), TMEM full/empty double buffer
// (MMA <-> epilogue), and a static persistent tile schedule (tile += gridDim.x).
//
// Operands are staged with 128-byte swizzle. Either operand may be K-major (contraction index
// contiguous in global memory; one TMA box {64, rows}) or MN-major (row/column index contiguous;
// boxes {64 mn, 64 k}, 8 KiB apart) -- the UMMA descriptors carry the difference, so wgrad GEMMs
// need no transposes. A may also come from a 4-D NHWC tensor map (3x3 SAME convolution taps with
// TMA out-of-bounds zero fill doing the padding) or be split across two tensor maps.
#pragma once
#include <type_traits>
#include "ptx.cuh"

namespace vc {

constexpr int kBM = 128;        // tile rows (UMMA M)
constexpr int kBK = 64;         // k-block: 64 bf16 = one 128-byte swizzle row
constexpr int kABytes = kBM * kBK * 2;
constexpr int kGemmThreads = 384;
constexpr int kMaxStages = 8;
constexpr int kMaxAcc = 8;       // accumulator buffers in TMEM: 512 columns / tile width (2 at bn = 256 ... 8 at bn <= 64)
__host__ __device__ inline int gemm_acc_buffers(int bn) { return bn > 128 ? 2 : (bn > 64 ? 4 : 8); }
__host__ __device__ inline int gemm_acc_stride(int bn) { return bn > 128 ? 256 : (bn > 64 ? 128 : 64); }

enum AMode : int { A_KMAJOR = 0, A_MNMAJOR = 1, A_CONV3x3 = 2, A_WGRAD3x3 = 3 };

struct GemmCore {
  int m_tiles, n_tiles, k_blocks, splits;
  int bn;        // tile columns (UMMA N): multiple of 16, <= 256; multiple of 64 when B is MN-major
  int stages;
  int a_mode;    // AMode
  int b_mn;      // 1: B is MN-major (2-D map); 2: MN-major gathered from an NHWC map by 64-pixel patches (A_WGRAD3x3)
  int a_switch;  // A_KMAJOR: k-block at which A switches to tmA2 (coordinates restart at 0)
                 // A_MNMAJOR: m-tile at which A switches to tmA2. <0: never.
  int n_fast;    // tile order: 0 = m fastest (B tile shared by consecutive CTAs), 1 = n fastest (the n-tiles of one
                 // m-row run side by side, so a long-K A row panel is fetched from HBM once and hit in L2 after)
  // A_CONV3x3 geometry. A tile is 4 patches of 32 pixels; a patch is pw x ph x pn (w, h, image) with
  // pw*ph*pn == 32, so each epilogue warp owns one patch and 2x2 max-pool partners are lanes of the
  // same warp. Patches are arranged tw x th x (4/(tw*th)) inside the tile.
  // A_WGRAD3x3 (filter gradient of the same convolution): the contraction runs over pixels. k-block kb is one
  // pw x ph x pn = 64-pixel patch (tiles_w x tiles_h patches per pn images); tile row m = tap * Cin + cin, so the
  // two 64-row halves of an A tile are two (tap, 64-channel chunk) pairs, each gathered with its tap's pixel shift
  // (OOB zero fill = SAME padding); B is the same un-shifted patch of the output gradient.
  int pw, ph, pn, tw, th, tiles_w, tiles_h, cpk;  // cpk = Cin / 64 chunks per filter tap
  int n_img;                                      // A_WGRAD3x3: image count (an image index >= n_img zero-fills)
  int dual;      // 1 (bn = 256 only): a tile is TWO adjacent n-tiles sharing one A tile -- both 256-column accumulators fill the
                 // 512 TMEM columns (no double buffering), a stage holds A + two B blocks, epilogue group g drains n-tile 2j + g.
                 // The operand bytes an SM ingests per FLOP drop by a quarter (A 16 KB + 2 x 16 KB instead of 2 x (16 + 16) KB
                 // as a pair): the long-K GEMMs run at the ~90 B/clk an SM ingests, not at the tensor pipe's rate.
  int pair;      // 1: run as CTA pairs (gemm_tc_kernel<Epi, 1>: cluster of 2, tcgen05 cta_group::2). m_tiles is then even
                 // (rounded up; the surplus tile reads zero fill and its stores are clipped) and tmB's box holds bn / 2 rows
  int tap_rows;  // A_CONV3x3 over a window map (conv1_1): k-block kb is filter ROW kb; the three taps of the row and their
                 // channels are one contiguous run of the zero-padded input, so the box moves by (0, kb) instead of (s-1, r-1)
};

struct PatchOrigin {
  int w, h, n;
};
// Origin (w, h, image) of patch q (0..3) of tile m_blk.
__device__ __forceinline__ PatchOrigin conv_patch_origin(const GemmCore& g, int m_blk, int q) {
  const int tiw = m_blk % g.tiles_w;
  const int r = m_blk / g.tiles_w;
  const int tih = r % g.tiles_h;
  const int tin = r / g.tiles_h;
  const int tn = 4 / (g.tw * g.th);
  const int qx = q % g.tw, qy = (q / g.tw) % g.th, qn = q / (g.tw * g.th);
  PatchOrigin o;
  o.w = (tiw * g.tw + qx) * g.pw;
  o.h = (tih * g.th + qy) * g.ph;
  o.n = (tin * tn + qn) * g.pn;
  return o;
}

__host__ __device__ inline int gemm_stage_bytes(int bn) { return kABytes + bn * kBK * 2; }
// dynamic smem = [1024-align slack][stages x (A tile + B tile)][epilogue staging][barriers, 512 B]
__host__ inline int gemm_pick_stages(int bn, int epi_bytes) {
  int s = (227 * 1024 - 1024 - 512 - epi_bytes) / gemm_stage_bytes(bn);
  return s > kMaxStages ? kMaxStages : s;
}
__host__ inline int gemm_smem_bytes(int bn, int stages, int epi_bytes) {
  return stages * gemm_stage_bytes(bn) + epi_bytes + 1024 + 512;
}

struct TileCoord {
  int m_blk, n_blk, split, kb_begin, kb_end;
};

__device__ __forceinline__ TileCoord decode_tile(const GemmCore& g, int tile) {
  TileCoord t;
  const int mn = g.m_tiles * g.n_tiles;
  t.split = tile / mn;
  const int rem = tile - t.split * mn;
  if (g.n_fast) {
    t.m_blk = rem / g.n_tiles;
    t.n_blk = rem - t.m_blk * g.n_tiles;
  } else {
    t.n_blk = rem / g.m_tiles;
    t.m_blk = rem - t.n_blk * g.m_tiles;
  }
  const int kps = (g.k_blocks + g.splits - 1) / g.splits;
  t.kb_begin = t.split * kps;
  t.kb_end = min(g.k_blocks, t.kb_begin + kps);
  return t;
}

// An epilogue may declare `static constexpr bool kPrefetch = true`, a `struct Pre` and
//   __device__ void prefetch(const GemmCore& g, const TileCoord& t, int row, Pre& pre) const
// which the epilogue threads call BEFORE they wait for the tile's accumulator (global loads whose latency then hides behind
// the tile's MMAs); operator() receives `const Pre*` as a last argument.
struct EpiNoPre {};
template <class E, class = void>
struct epi_prefetches {
  static constexpr bool value = false;
  using type = EpiNoPre;
};
template <class E>
struct epi_prefetches<E, std::enable_if_t<E::kPrefetch>> {
  static constexpr bool value = true;
  using type = typename E::Pre;
};

// Epi must provide:
//   static constexpr int kSmemBytes;   // epilogue staging (multiple of 2048), split evenly over the 2 groups
//   static constexpr bool kPairOk;     // may run under gemm_tc_kernel<Epi, 1> (CTA pairs)
//   __device__ void operator()(uint32_t tmem_row_addr, const GemmCore& g, const TileCoord& t, int row,
//                              uint8_t* grp_smem, int grp, int& phase) const
// called by each of the 128 threads of epilogue group `grp` (row = 0..127 = TMEM lane = tile row) once the
// accumulator is complete. tmem_row_addr addresses column 0 of this thread's warp lane quarter;
// grp_smem is the group's half of the staging area; phase is per-thread state carried across tiles.
//   __device__ void finish(uint8_t* grp_smem, int phase) const   // called once per epilogue thread after the last tile
//
// kPair = 1 (launched as clusters of 2): the two CTAs of a pair own m-tiles 2p and 2p + 1 of the same n-tile. Each stages
// its own A tile and HALF of the B tile (rows / columns [rank * bn/2, +bn/2) of it); every TMA load signals the LEADER's
// `full` barrier (cluster rank 0), whose MMA warp issues one 256-row tcgen05.mma.cta_group::2 per k-step and releases the
// ring stage / publishes the accumulator in BOTH CTAs with a multicast commit. Each CTA drains its own 128 TMEM lanes with
// the unchanged epilogue and arrives on the leader's `tempty` barrier. Per SM and k-step the tensor core then reads
// 4 KB + bn/2 rows of operands from shared memory instead of 4 KB + bn rows (the 128 x 256 tile of cta_group::1 is below
// the shared-memory ridge), and a stage is 32 KB instead of 48 KB of L2 -> SM traffic.
template <class Epi, int kPair>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
               const __grid_constant__ CUtensorMap tmB, const GemmCore g, const __grid_constant__ Epi epi) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int b_rows = kPair ? g.bn / 2 : g.bn;  // rows (K-major) / columns (MN-major) of B this CTA stages
  const bool dual = g.dual != 0;
  const int b_block = b_rows * kBK * 2;        // bytes of one staged B block
  const int stage_bytes = kABytes + (dual ? 2 : 1) * b_block;
  uint8_t* epi_smem = smem + g.stages * stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_smem + Epi::kSmemBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kMaxStages;
  uint64_t* tfull = bars + 2 * kMaxStages;
  uint64_t* tempty = bars + 2 * kMaxStages + kMaxAcc;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 2 * kMaxAcc);
  const int nacc = dual ? 1 : gemm_acc_buffers(g.bn);
  const int acc_stride = dual ? 512 : gemm_acc_stride(g.bn);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // pair mode: the schedule runs over PAIR tiles (m_tiles / 2 of them per n-tile), one cluster per pair tile
  const uint32_t rank = kPair ? cluster_ctarank() : 0u;
  const int n_sched = dual ? (g.n_tiles + 1) / 2 : g.n_tiles;  // dual: the schedule runs over pairs of n-tiles
  const int total_tiles = (kPair ? g.m_tiles / 2 : g.m_tiles) * n_sched * g.splits;
  const int tile0 = kPair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int tile_step = kPair ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  GemmCore gs = g;  // what decode_tile sees
  if (kPair) gs.m_tiles = g.m_tiles / 2;
  gs.n_tiles = n_sched;
  auto tile_of = [&](int tile) {
    TileCoord t = decode_tile(gs, tile);
    if (kPair) t.m_blk = t.m_blk * 2 + (int)rank;
    if (dual) t.n_blk *= 2;  // first n-tile of the pair; the second one is t.n_blk + 1 (past N: zero fill in, nothing out)
    return t;
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmA2);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < g.stages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < kMaxAcc; ++i) {
      mbar_init(&tfull[i], 1);
      // pair: the four epilogue warps of both CTAs arrive on the leader's; dual: both epilogue groups drain every tile
      mbar_init(&tempty[i], (kPair ? 8 : 4) * (dual ? 2 : 1));
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    if (kPair) tmem_alloc_pair<512>(tmem_slot);
    else tmem_alloc<512>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();                // the TMEM address written by tcgen05.alloc is read by every thread below
  if (kPair) cluster_sync_all();  // the peer's barriers must be initialised before anything arrives on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    // (whole warp walks the schedule with uniform control flow; one elected lane issues the copies)
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = tile0; tile < total_tiles; tile += tile_step) {
        const TileCoord t = tile_of(tile);
        PatchOrigin po[4];
        if (g.a_mode == A_CONV3x3) {
#pragma unroll
          for (int q = 0; q < 4; ++q) po[q] = conv_patch_origin(g, t.m_blk, q);
        }
        for (int kb = t.kb_begin; kb < t.kb_end; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sa = smem + stage * stage_bytes;
          uint8_t* sb = sa + kABytes;
          if (kPair) {
            // both CTAs load; the bytes of both land on the leader's barrier
            if (elect_one()) {
              if (rank == 0) mbar_expect_tx(&full[stage], 2 * stage_bytes);
              const uint32_t fb = mapa_u32(&full[stage], 0);
              if (g.a_mode == A_KMAJOR) {
                if (g.a_switch >= 0 && kb >= g.a_switch)
                  tma_load_2d_pair(sa, &tmA2, fb, (kb - g.a_switch) * kBK, t.m_blk * kBM);
                else
                  tma_load_2d_pair(sa, &tmA, fb, kb * kBK, t.m_blk * kBM);
              } else if (g.a_mode == A_MNMAJOR) {
                const bool second = g.a_switch >= 0 && t.m_blk >= g.a_switch;
                const CUtensorMap* tm = second ? &tmA2 : &tmA;
                const int m0 = (second ? t.m_blk - g.a_switch : t.m_blk) * kBM;
                tma_load_2d_pair(sa, tm, fb, m0, kb * kBK);
                tma_load_2d_pair(sa + 8192, tm, fb, m0 + 64, kb * kBK);
              } else if (g.a_mode == A_WGRAD3x3) {
                const int tiw = kb % g.tiles_w;
                const int r = kb / g.tiles_w;
                const int w0 = tiw * g.pw, h0 = (r % g.tiles_h) * g.ph, n0 = (r / g.tiles_h) * g.pn;
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                  const int chunk = t.m_blk * 2 + half;
                  const int tap = chunk / g.cpk;
                  const int c0 = (chunk - tap * g.cpk) * 64;
                  const int fr = tap / 3, fs = tap - fr * 3;
                  if (tap < 9)
                    tma_load_4d_pair(sa + half * 8192, &tmA, fb, c0, w0 + fs - 1, h0 + fr - 1, n0);
                  else
                    tma_load_4d_pair(sa + half * 8192, &tmA, fb, 0, 0, 0, g.n_img);  // past the last tap: zeros
                }
                for (int h = 0; h < (dual ? 2 : 1); ++h)
                  for (int j = 0; j < b_rows; j += 64)
                    tma_load_4d_pair(sb + h * b_block + j * 128, &tmB, fb, (t.n_blk + h) * g.bn + (int)rank * b_rows + j, w0, h0, n0);
              } else {  // A_CONV3x3
                const int tap = kb / g.cpk;
                const int c0 = (kb - tap * g.cpk) * kBK;
                const int fr = g.tap_rows ? tap + 1 : tap / 3, fs = g.tap_rows ? 1 : tap - (tap / 3) * 3;
#pragma unroll
                for (int q = 0; q < 4; ++q)
                  tma_load_4d_pair(sa + q * 4096, &tmA, fb, c0, po[q].w + fs - 1, po[q].h + fr - 1, po[q].n);
              }
              for (int h = 0; h < (dual ? 2 : 1); ++h) {
                const int nb = (t.n_blk + h) * g.bn + (int)rank * b_rows;
                uint8_t* sbh = sb + h * b_block;
                if (g.b_mn == 1) {
                  for (int j = 0; j < b_rows; j += 64) tma_load_2d_pair(sbh + j * 128, &tmB, fb, nb + j, kb * kBK);
                } else if (g.b_mn == 0) {
                  tma_load_2d_pair(sbh, &tmB, fb, kb * kBK, nb);
                }
              }
            }
          } else if (elect_one()) {
          mbar_expect_tx(&full[stage], stage_bytes);
          if (g.a_mode == A_KMAJOR) {
            if (g.a_switch >= 0 && kb >= g.a_switch)
              tma_load_2d(sa, &tmA2, &full[stage], (kb - g.a_switch) * kBK, t.m_blk * kBM);
            else
              tma_load_2d(sa, &tmA, &full[stage], kb * kBK, t.m_blk * kBM);
          } else if (g.a_mode == A_MNMAJOR) {
            const bool second = g.a_switch >= 0 && t.m_blk >= g.a_switch;
            const CUtensorMap* tm = second ? &tmA2 : &tmA;
            const int m0 = (second ? t.m_blk - g.a_switch : t.m_blk) * kBM;
            tma_load_2d(sa, tm, &full[stage], m0, kb * kBK);
            tma_load_2d(sa + 8192, tm, &full[stage], m0 + 64, kb * kBK);
          } else if (g.a_mode == A_WGRAD3x3) {
            const int tiw = kb % g.tiles_w;
            const int r = kb / g.tiles_w;
            const int w0 = tiw * g.pw, h0 = (r % g.tiles_h) * g.ph, n0 = (r / g.tiles_h) * g.pn;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              const int chunk = t.m_blk * 2 + half;  // 64-row chunk index over [9 taps][cpk chunks]
              const int tap = chunk / g.cpk;
              const int c0 = (chunk - tap * g.cpk) * 64;
              const int fr = tap / 3, fs = tap - fr * 3;
              if (tap < 9)
                tma_load_4d(sa + half * 8192, &tmA, &full[stage], c0, w0 + fs - 1, h0 + fr - 1, n0);
              else
                tma_load_4d(sa + half * 8192, &tmA, &full[stage], 0, 0, 0, g.n_img);  // past the last tap: zeros
            }
            for (int h = 0; h < (dual ? 2 : 1); ++h)
              for (int j = 0; j < g.bn; j += 64)
                tma_load_4d(sb + h * b_block + j * 128, &tmB, &full[stage], (t.n_blk + h) * g.bn + j, w0, h0, n0);
          } else {
            const int tap = kb / g.cpk;
            const int c0 = (kb - tap * g.cpk) * kBK;
            const int fr = g.tap_rows ? tap + 1 : tap / 3, fs = g.tap_rows ? 1 : tap - (tap / 3) * 3;
#pragma unroll
            for (int q = 0; q < 4; ++q)
              tma_load_4d(sa + q * 4096, &tmA, &full[stage], c0, po[q].w + fs - 1, po[q].h + fr - 1, po[q].n);
          }
          for (int h = 0; h < (dual ? 2 : 1); ++h) {
            uint8_t* sbh = sb + h * b_block;
            if (g.b_mn == 1) {
              for (int j = 0; j < g.bn; j += 64)
                tma_load_2d(sbh + j * 128, &tmB, &full[stage], (t.n_blk + h) * g.bn + j, kb * kBK);
            } else if (g.b_mn == 0) {
              tma_load_2d(sbh, &tmB, &full[stage], kb * kBK, (t.n_blk + h) * g.bn);
            }
          }
          }
          __syncwarp();
          if (++stage == g.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    // the whole warp walks the schedule (uniform control flow, descriptors in uniform registers); one elected lane
    // issues the tcgen05 instructions. In pair mode only the leader CTA issues (for both SMs).
    if (rank == 0) {
      const bool a_mn = g.a_mode == A_MNMAJOR || g.a_mode == A_WGRAD3x3;
      const uint32_t idesc = make_idesc_bf16(kPair ? 2 * kBM : kBM, g.bn, a_mn, g.b_mn != 0);
      const uint32_t a_lbo = a_mn ? 8192u : 16u;
      const uint32_t b_lbo = g.b_mn ? 8192u : 16u;
      const uint32_t a_kstep = a_mn ? (2048u >> 4) : (32u >> 4);
      const uint32_t b_kstep = g.b_mn ? (2048u >> 4) : (32u >> 4);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = tile0; tile < total_tiles; tile += tile_step) {
        const TileCoord t = tile_of(tile);
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * acc_stride;
        for (int kb = t.kb_begin; kb < t.kb_end; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * stage_bytes);
          const uint64_t adesc = make_smem_desc(sa, a_lbo, 1024);
          const uint64_t bdesc = make_smem_desc(sa + kABytes, b_lbo, 1024);
          const uint64_t bdesc1 = make_smem_desc(sa + kABytes + b_block, b_lbo, 1024);  // dual: the second n-tile's block
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k) {
              if (kPair)
                umma_bf16_pair(d_tmem, adesc + uint64_t(k * a_kstep), bdesc + uint64_t(k * b_kstep), idesc,
                               (kb > t.kb_begin || k > 0) ? 1u : 0u);
              else
                umma_bf16(d_tmem, adesc + uint64_t(k * a_kstep), bdesc + uint64_t(k * b_kstep), idesc,
                          (kb > t.kb_begin || k > 0) ? 1u : 0u);
              if (dual) {
                if (kPair)
                  umma_bf16_pair(d_tmem + 256, adesc + uint64_t(k * a_kstep), bdesc1 + uint64_t(k * b_kstep), idesc,
                                 (kb > t.kb_begin || k > 0) ? 1u : 0u);
                else
                  umma_bf16(d_tmem + 256, adesc + uint64_t(k * a_kstep), bdesc1 + uint64_t(k * b_kstep), idesc,
                            (kb > t.kb_begin || k > 0) ? 1u : 0u);
              }
            }
            if (kPair) umma_commit_pair(&empty[stage]);
            else umma_commit(&empty[stage]);
          }
          __syncwarp();
          if (++stage == g.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (elect_one()) {
          if (kPair) umma_commit_pair(&tfull[acc]);
          else umma_commit(&tfull[acc]);
        }
        __syncwarp();
        if (++acc == nacc) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue
    // tile ordinal `ord` of this CTA lives in accumulator buffer ord % nacc; group (warp - 4) / 4 drains the
    // ordinals of its own parity (nacc is even, so a buffer always belongs to the same group)
    const int q = warp & 3;
    const int grp = (warp - 4) >> 2;
    int epi_phase = 0;
    uint8_t* grp_smem = epi_smem + grp * (Epi::kSmemBytes / 2);
    int ord = 0;
    const uint32_t tempty_leader = kPair ? mapa_u32(tempty, 0) : 0u;
    for (int tile = tile0; tile < total_tiles; tile += tile_step, ++ord) {
      if (!dual && (ord & 1) != grp) continue;
      const int acc = ord % nacc;
      const uint32_t acc_phase = (ord / nacc) & 1;
      TileCoord t = tile_of(tile);
      if (dual) t.n_blk += grp;  // dual: group g drains the tile's n-tile g (TMEM columns [256 g, 256 g + 256))
      [[maybe_unused]] typename epi_prefetches<Epi>::type pre;
      if constexpr (epi_prefetches<Epi>::value) epi.prefetch(g, t, q * 32 + lane, pre);
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      if (t.kb_end > t.kb_begin) {
        const uint32_t taddr = tmem_base + acc * acc_stride + (dual ? grp * 256 : 0) + (uint32_t(q * 32) << 16);
        if constexpr (epi_prefetches<Epi>::value) epi(taddr, g, t, q * 32 + lane, grp_smem, grp, epi_phase, &pre);
        else epi(taddr, g, t, q * 32 + lane, grp_smem, grp, epi_phase);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (kPair) mbar_arrive_cluster(tempty_leader + acc * 8);
        else mbar_arrive(&tempty[acc]);
      }
    }
    epi.finish(grp_smem, epi_phase);
  }
  tc_fence_before();
  if (kPair) cluster_sync_all();  // neither CTA may leave while the other can still touch its barriers / shared memory
  else __syncthreads();
  if (warp == 2) {
    if (kPair) tmem_dealloc_pair<512>(tmem_base);
    else tmem_dealloc<512>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------
// Generic store epilogue: out[m, n] = act(acc + bias[n]) as fp32 or bf16, or fp32 atomic
// accumulation (split-K / gradient accumulation).
struct EpiStore {
  void* out;
  const float* bias;  // nullable, indexed by n
  long long ld;       // elements
  int M, N, bn;
  int relu;
  int out_bf16;
  int atomic;  // fp32 atomicAdd (bias added by split 0 only)
  float alpha;
  // staging for the atomic path: one 32 x 32 fp32 block per epilogue warp (2 groups x 4 warps x 4 KB)
  static constexpr int kSmemBytes = 32 * 1024;
  static constexpr bool kPairOk = true;  // tile-local epilogue: runs unchanged in either CTA of a pair
  __device__ __forceinline__ void finish(uint8_t*, int) const {}

  __device__ __forceinline__ void operator()(uint32_t taddr, const GemmCore&, const TileCoord& t, int row,
                                             uint8_t* grp_smem, int, int&) const {
    const int m = t.m_blk * kBM + row;
    const int n_base = t.n_blk * bn;
    const bool row_ok = m < M;
    const int lane = row & 31;
    float* stage = reinterpret_cast<float*>(grp_smem) + (row >> 5) * 1024;
    for (int c = 0; c < bn; c += 32) {
      if (n_base + c >= N) break;  // warp-uniform
      float v[32];
      __syncwarp();
      tmem_ld32(taddr + c, v);
      tmem_ld_wait();
      const int n0 = n_base + c;
      const int nv = min(min(32, bn - c), N - n0);
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float x = v[j] * alpha;
        if (bias != nullptr && j < nv && (!atomic || t.split == 0)) x += bias[n0 + j];
        if (relu) x = fmaxf(x, 0.f);
        v[j] = x;
      }
      // vector stores need every row of the block 16-byte aligned (warp-uniform test)
      const int esize = out_bf16 ? 2 : 4;
      const bool vec_ok = nv == 32 && ((ld * esize) & 15) == 0 &&
                          ((reinterpret_cast<uintptr_t>(out) + (uintptr_t)n0 * esize) & 15) == 0;
      if (atomic || !vec_ok) {
        // A TMEM lane is a tile row, so thread-per-row atomics (or scalar stores) would touch 32 cache lines per warp
        // instruction, one 4-byte piece per 32-byte sector. Transpose the warp's 32 x 32 block through shared memory
        // (XOR-swizzled columns: conflict-free both ways) so each instruction covers 32 consecutive elements of ONE
        // output row. Also the path for ragged / unaligned blocks: no dynamically indexed register array anywhere.
#pragma unroll
        for (int j = 0; j < 32; ++j) stage[lane * 32 + (j ^ lane)] = v[j];
        __syncwarp();
        const int m0 = t.m_blk * kBM + (row & ~31);
        const int rows_ok = min(32, M - m0);
        if (lane < nv) {
          const long long o0 = (long long)m0 * ld + n0 + lane;
          if (atomic) {
            float* o = reinterpret_cast<float*>(out) + o0;
#pragma unroll 4
            for (int r = 0; r < rows_ok; ++r) atomicAdd(o + (long long)r * ld, stage[r * 32 + (lane ^ r)]);
          } else if (out_bf16) {
            __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out) + o0;
#pragma unroll 4
            for (int r = 0; r < rows_ok; ++r) o[(long long)r * ld] = __float2bfloat16(stage[r * 32 + (lane ^ r)]);
          } else {
            float* o = reinterpret_cast<float*>(out) + o0;
#pragma unroll 4
            for (int r = 0; r < rows_ok; ++r) o[(long long)r * ld] = stage[r * 32 + (lane ^ r)];
          }
        }
      } else if (row_ok) {
        if (out_bf16) {
          __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out) + (long long)m * ld + n0;
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            uint4 u;
            u.x = pack_bf16(v[j], v[j + 1]);
            u.y = pack_bf16(v[j + 2], v[j + 3]);
            u.z = pack_bf16(v[j + 4], v[j + 5]);
            u.w = pack_bf16(v[j + 6], v[j + 7]);
            *reinterpret_cast<uint4*>(o + j) = u;
          }
        } else {
          float* o = reinterpret_cast<float*>(out) + (long long)m * ld + n0;
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
      }
    }
  }
};

// ------------------------------------------------------------------------------------------
// bf16 tile store through shared memory + TMA: out = act(acc * alpha + bias[n]) written as 64-column
// (128-byte) swizzled rows into the group's staging buffer, then ONE bulk tensor store per 128-row x 64-column
// chunk (the TMA unit's cost is per instruction, so four 4 KB stores per chunk throttled the 64-channel layers).
// TMA clips rows/columns/pixels that fall outside the tensor, so ragged M, N and image borders need no predicates.
// Three output geometries (the 4 patches of a tile stack into one box, see conv_patch_origin):
//   kRows   : plain [M, N] matrix, box {64, 128}
//   kConv   : NHWC conv output, box {64, pw, ph*th, pn*tn}
//   kConvPool: same with a fused 2x2/2 max-pool (partners are lanes of a warp), box {64, pw/2, ph*th/2, pn*tn}
enum TmaOutMode : int { kRows = 0, kConv = 1, kConvPool = 2 };

// kRelu = true (EpiTmaRelu, kConv mode only; the input gradients of the fine-tune pass): the tile is the gradient with
// respect to a ReLU output, `relu_src` holds that forward activation (same NHWC geometry as the output): the epilogue zeroes
// the gradient where the activation is not positive -- what k_relu_bwd did in a separate pass over three feature maps --
// and accumulates the per-channel sums of the masked (bf16-rounded) gradient, i.e. the bias gradient, in a 2 KB
// shared-memory table per epilogue group that is flushed to `dbias` with one atomic per channel at the end of the CTA.
template <bool kRelu>
struct EpiTmaT {
  CUtensorMap tm;
  const float* bias;  // nullable, indexed by n (16-byte aligned)
  int N, bn, relu, mode;
  float alpha;
  const __nv_bfloat16* relu_src;  // kRelu: forward activation [n_img, H, W, N] bf16
  float* dbias;                   // kRelu: [N] fp32, accumulated
  int W, H, n_img;                // kRelu: extent of the output feature map
  // 2 groups x (128 rows x 128 B staging [+ 512 fp32 channel sums + 4 x 64 partial sums of the current chunk])
  static constexpr int kStageBytes = 16 * 1024;
  static constexpr int kSmemBytes = kRelu ? 38 * 1024 : 32 * 1024;
  static constexpr bool kPairOk = true;
  static constexpr bool kPrefetch = kRelu;
  struct Pre {
    const __nv_bfloat16* row;  // this thread's pixel of relu_src (nullptr: outside the feature map)
    uint4 a[8];                // its first 64 channels of the tile
  };
  __device__ __forceinline__ void prefetch(const GemmCore& g, const TileCoord& t, int row, Pre& pre) const {
    const int lane = row & 31, q = row >> 5;
    const PatchOrigin pq = conv_patch_origin(g, t.m_blk, q);
    const int w = pq.w + lane % g.pw, h = pq.h + (lane / g.pw) % g.ph, n = pq.n + lane / (g.pw * g.ph);
    pre.row = nullptr;
    if (w < W && h < H && n < n_img && t.n_blk * bn + 64 <= N) {
      pre.row = relu_src + (((long long)n * H + h) * W + w) * N;
      const uint4* a4 = reinterpret_cast<const uint4*>(pre.row + t.n_blk * bn);
#pragma unroll
      for (int j = 0; j < 8; ++j) pre.a[j] = __ldg(a4 + j);
    }
  }

  __device__ __forceinline__ void finish(uint8_t* buf, int phase) const {
    if ((threadIdx.x & 127) == 0) bulk_wait_all();  // the issuing thread of each group
    if constexpr (kRelu) {
      if (!(phase & 2)) return;  // this group never saw a tile (group-uniform): its table was never cleared
      named_bar_sync(1 + (((threadIdx.x >> 5) - 4) >> 2), 128);  // every warp's shared-memory atomics are in
      const float* sums = reinterpret_cast<const float*>(buf + kStageBytes);
      for (int c = threadIdx.x & 127; c < N && c < 512; c += 128)
        if (sums[c] != 0.f) atomicAdd(dbias + c, sums[c]);
    }
  }

  // The accumulator row is rounded to packed bf16 pairs right after the bias add; ReLU and the 2x2 max-pool then run
  // on the pairs (max commutes with the monotone rounding, so the result equals pooling / clamping in fp32 first)
  // with half the instructions and half the shuffles.
  __device__ __forceinline__ void operator()(uint32_t taddr, const GemmCore& g, const TileCoord& t, int row,
                                             uint8_t* buf, int grp, int& phase, const Pre* pre = nullptr) const {
    const int lane = row & 31, q = row >> 5;
    const int n_base = t.n_blk * bn;
    PatchOrigin po{0, 0, 0};
    if (mode != kRows) po = conv_patch_origin(g, t.m_blk, 0);
    [[maybe_unused]] const __nv_bfloat16* relu_row = nullptr;  // this thread's pixel of the forward activation
    [[maybe_unused]] float* sums = reinterpret_cast<float*>(buf + kStageBytes);
    if constexpr (kRelu) {
      if (!(phase & 2)) {  // first tile of this group: clear the channel sums (bit 1 of the carried state marks it done)
        for (int c = row; c < 512; c += 128) sums[c] = 0.f;
        named_bar_sync(1 + grp, 128);
        phase |= 2;
      }
      relu_row = pre->row;
    }
#pragma unroll 1
    for (int c = 0; c < bn; c += 64) {
      const int n0 = n_base + c;
      if (n0 >= N) break;  // warp-uniform
      float v[64];
      __syncwarp();
      tmem_ld32(taddr + c, v);
      tmem_ld32(taddr + c + 32, v + 32);  // bn is a multiple of 64 for this epilogue
      tmem_ld_wait();
      uint32_t p[32];
      if (bias != nullptr && n0 + 64 <= N) {
        const float4* b4 = reinterpret_cast<const float4*>(bias + n0);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float4 b = __ldg(b4 + j);
          p[2 * j] = pack_bf16(fmaf(v[4 * j], alpha, b.x), fmaf(v[4 * j + 1], alpha, b.y));
          p[2 * j + 1] = pack_bf16(fmaf(v[4 * j + 2], alpha, b.z), fmaf(v[4 * j + 3], alpha, b.w));
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float b0 = (bias != nullptr && n0 + 2 * j < N) ? bias[n0 + 2 * j] : 0.f;
          const float b1 = (bias != nullptr && n0 + 2 * j + 1 < N) ? bias[n0 + 2 * j + 1] : 0.f;
          p[j] = pack_bf16(fmaf(v[2 * j], alpha, b0), fmaf(v[2 * j + 1], alpha, b1));
        }
      }
      if (relu) {
        const __nv_bfloat162 zero = __floats2bfloat162_rn(0.f, 0.f);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const __nv_bfloat162 x = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&p[j]), zero);
          p[j] = *reinterpret_cast<const uint32_t*>(&x);
        }
      }
      if constexpr (kRelu) {
        // gradient of ReLU: keep where the forward activation is positive; pixels outside the map contribute nothing
        if (relu_row != nullptr && n0 + 64 <= N) {
          const uint4* a4 = reinterpret_cast<const uint4*>(relu_row + n0);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint4 a = c == 0 ? pre->a[j] : __ldg(a4 + j);  // the first chunk was fetched ahead of the accumulator
            const uint32_t aw[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              // bf16 > 0  <=>  sign bit clear and not zero (activations are ReLU outputs: never negative, never NaN)
              const uint32_t lo = (aw[k] & 0x7fffu) != 0 && (aw[k] & 0x8000u) == 0 ? 0x0000ffffu : 0u;
              const uint32_t hi = (aw[k] & 0x7fff0000u) != 0 && (aw[k] & 0x80000000u) == 0 ? 0xffff0000u : 0u;
              p[4 * j + k] &= (lo | hi);
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) p[j] = 0u;
        }
      }
      int srow = row;
      bool writer = true;
      if (mode == kConvPool) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          uint32_t o = __shfl_xor_sync(0xffffffffu, p[j], 1);
          __nv_bfloat162 x = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&p[j]), *reinterpret_cast<__nv_bfloat162*>(&o));
          uint32_t xu = *reinterpret_cast<const uint32_t*>(&x);
          o = __shfl_xor_sync(0xffffffffu, xu, g.pw);
          x = __hmax2(x, *reinterpret_cast<__nv_bfloat162*>(&o));
          p[j] = *reinterpret_cast<const uint32_t*>(&x);
        }
        const int w = lane % g.pw, h = (lane / g.pw) % g.ph, n = lane / (g.pw * g.ph);
        writer = ((w | h) & 1) == 0;
        srow = q * 8 + (w >> 1) + (g.pw >> 1) * ((h >> 1) + (g.ph >> 1) * n);
      }
      // the staging buffer was handed to a bulk store one chunk ago: wait until the TMA engine has read it
      if (row == 0) bulk_wait_read<0>();
      named_bar_sync(1 + grp, 128);
      if (writer) {
        uint8_t* rp = buf + srow * 128;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<uint4*>(rp + ((j ^ (srow & 7)) << 4)) = make_uint4(p[4 * j], p[4 * j + 1], p[4 * j + 2], p[4 * j + 3]);
      }
      fence_proxy_async();
      named_bar_sync(1 + grp, 128);
      if constexpr (kRelu) {
        // channel sums of the staged chunk: thread (rg, cp) adds channel pair cp over rows [32 rg, 32 rg + 32)
        const int cp = row & 31, rg = row >> 5;
        float sx = 0.f, sy = 0.f;
#pragma unroll 8
        for (int r = rg * 32; r < rg * 32 + 32; ++r) {
          const uint32_t u = *reinterpret_cast<const uint32_t*>(buf + r * 128 + (((cp >> 2) ^ (r & 7)) << 4) + (cp & 3) * 4);
          sx += __uint_as_float(u << 16);
          sy += __uint_as_float(u & 0xffff0000u);
        }
        // (no shared-memory float atomics: they are compare-and-swap loops, and four warps meeting on 64 addresses doubled
        // the time of the 64-channel layer) the four row groups park their partial sums, row group 0 owns the table
        float* part = sums + 512;
        part[rg * 64 + 2 * cp] = sx;
        part[rg * 64 + 2 * cp + 1] = sy;
        named_bar_sync(1 + grp, 128);
        if (rg == 0 && n0 + 2 * cp + 1 < 512) {
          sums[n0 + 2 * cp] += part[2 * cp] + part[64 + 2 * cp] + part[128 + 2 * cp] + part[192 + 2 * cp];
          sums[n0 + 2 * cp + 1] += part[2 * cp + 1] + part[64 + 2 * cp + 1] + part[128 + 2 * cp + 1] + part[192 + 2 * cp + 1];
        }
      }
      if (row == 0) {
        if (mode == kRows)
          tma_store_2d(&tm, buf, n0, t.m_blk * kBM);
        else if (mode == kConv)
          tma_store_4d(&tm, buf, n0, po.w, po.h, po.n);
        else
          tma_store_4d(&tm, buf, n0, po.w >> 1, po.h >> 1, po.n);
        bulk_commit();
      }
      phase ^= 1;
    }
  }
};

using EpiTma = EpiTmaT<false>;
using EpiTmaRelu = EpiTmaT<true>;

// ------------------------------------------------------------------------------------------
// fp32 row-matrix tile store through shared memory + TMA: out[m, n] = acc * alpha + bias[n], 32 columns (128 bytes) at
// a time. Same reason as EpiTma: a TMEM lane is a tile row, so EpiStore's per-thread 16-byte stores touch 32 cache lines
// per warp instruction; for the [rows, vocab] fp32 logits of the decode step (232 MB per step at 5120 rows) that made
// the vocab projection run at 0.4 PFLOP/s. TMA clips the ragged last column block and the last rows.
struct EpiTmaF32 {
  CUtensorMap tm;     // fp32 [M, N] map, box {32, 128}, 128-byte swizzle
  const float* bias;  // nullable, indexed by n
  int N, bn;
  float alpha;
  static constexpr int kSmemBytes = 32 * 1024;  // 2 groups x (128 rows x 128 B)
  static constexpr bool kPairOk = true;

  __device__ __forceinline__ void finish(uint8_t*, int) const {
    if ((threadIdx.x & 127) == 0) bulk_wait_all();
  }
  __device__ __forceinline__ void operator()(uint32_t taddr, const GemmCore&, const TileCoord& t, int row, uint8_t* buf, int grp,
                                             int& phase) const {
    const int n_base = t.n_blk * bn;
#pragma unroll 1
    for (int c = 0; c < bn; c += 32) {
      const int n0 = n_base + c;
      if (n0 >= N) break;  // warp-uniform
      float v[32];
      __syncwarp();
      tmem_ld32(taddr + c, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float b = (bias != nullptr && n0 + j < N) ? __ldg(bias + n0 + j) : 0.f;
        v[j] = fmaf(v[j], alpha, b);
      }
      if (row == 0) bulk_wait_read<0>();  // the staging buffer was handed to a bulk store one chunk ago
      named_bar_sync(1 + grp, 128);
      uint8_t* rp = buf + row * 128;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<uint4*>(rp + ((j ^ (row & 7)) << 4)) = *reinterpret_cast<uint4*>(v + 4 * j);
      fence_proxy_async();
      named_bar_sync(1 + grp, 128);
      if (row == 0) {
        tma_store_2d(&tm, buf, n0, t.m_blk * kBM);
        bulk_commit();
      }
      phase ^= 1;
    }
  }
};

// ------------------------------------------------------------------------------------------
// 3x3 SAME convolution for Cin == 64 with the input patch staged ONCE per tile ("halo" form).
//
// gemm_tc_kernel's A_CONV3x3 mode re-reads the 128-pixel input patch from L2 for each of the 9 filter taps
// (9 x 16 KB per tile) and the [bn, 576] filter slice for every tile; at Cout <= 128 that operand traffic, not the
// tensor pipe, bounds the layer (conv1_2 / conv2_1: ~7 TB/s of L2->SM traffic). Here a tile is 8 (w) x 16 (h) output
// pixels; one 4-D TMA box {64 ch, 16 w, 18 h} brings the patch plus its halo into shared memory with a 16-row line
// pitch (128-byte swizzled rows, OOB zero fill = SAME padding), and the 9 taps are 9 shifted UMMA descriptors into
// that buffer: start = base + (fr*16 + fs)*128 B, 8-row groups (one image row of 8 pixels) SBO = 2048 B apart. The
// 128-byte swizzle is a function of the absolute shared-memory address bits, so a start address that is not
// 1024-byte aligned needs no descriptor base offset (measured: base offset 0 is the setting that matches the oracle). The 64-output-
// channel filter slice (9 x 8 KB) is loaded once per CTA and stays resident. Epilogue: EpiTma (bias + ReLU + pool).
constexpr int kHaloLineRows = 16;                               // halo rows (pixels) per image line in smem
constexpr int kHaloLines = 18;                                  // 16 output lines + 2
constexpr int kHaloBytes = kHaloLineRows * kHaloLines * 128;    // 36,864
constexpr int kHaloWBytes = 9 * 64 * 128;                       // resident filter slice: 9 taps x [64 x 64] bf16
constexpr int kHaloStages = 3;

__host__ inline int conv_halo_smem_bytes(int epi_bytes) {
  return kHaloWBytes + kHaloStages * kHaloBytes + epi_bytes + 1024 + 512;
}

// kPair = 1: CTA pairs as in gemm_tc_kernel<Epi, 1>. The pair owns m-tiles 2p and 2p + 1 (each CTA stages its own halo)
// and ONE n-tile of bn = 64 or 128 output channels, of whose resident filter slice each CTA holds bn / 2 rows per tap;
// the leader issues 256-row MMAs. At Cout = 128 an SM then reads 4 KB + 2 KB of operands per 128 x 128 x 16 MMA half
// (64 tensor cycles) instead of 4 KB + 2 KB per 128 x 64 x 16 (32 cycles).
template <class Epi, int kPair>
__global__ void __launch_bounds__(kGemmThreads, 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmCore g,
                 const __grid_constant__ Epi epi) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* wres = smem;
  uint8_t* halo = smem + kHaloWBytes;
  uint8_t* epi_smem = halo + kHaloStages * kHaloBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_smem + Epi::kSmemBytes);
  uint64_t* hfull = bars;
  uint64_t* hempty = bars + kHaloStages;
  uint64_t* tfull = bars + 2 * kHaloStages;
  uint64_t* tempty = bars + 2 * kHaloStages + kMaxAcc;
  uint64_t* wfull = bars + 2 * kHaloStages + 2 * kMaxAcc;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kHaloStages + 2 * kMaxAcc + 1);
  const int bn = kPair ? g.bn : 64;
  const int wtap = (kPair ? bn / 2 : 64) * 128;  // bytes per filter tap held by this CTA
  const int acc_stride = bn > 64 ? 128 : 64;
  const int nacc = 512 / acc_stride;
  const uint32_t rank = kPair ? cluster_ctarank() : 0u;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // static schedule: a CTA keeps one n-tile (its resident filter slice) and strides over the m-tiles
  // (pair mode: the same over pair tiles and clusters; m_first / m_step / m_end then count PAIRS, see m_of)
  const int sched_id = kPair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int sched_n = kPair ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int n_blk = sched_id % g.n_tiles;
  const int m_first = sched_id / g.n_tiles;
  const int m_step = sched_n / g.n_tiles;
  const int m_end = kPair ? (g.m_tiles + 1) / 2 : g.m_tiles;
  auto m_of = [&](int i) { return kPair ? 2 * i + (int)rank : i; };  // a surplus tile of the last pair is all zero fill

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kHaloStages; ++i) {
      mbar_init(&hfull[i], 1);
      mbar_init(&hempty[i], 1);
    }
    for (int i = 0; i < kMaxAcc; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], kPair ? 8 : 4);
    }
    mbar_init(wfull, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    if (kPair) tmem_alloc_pair<512>(tmem_slot);
    else tmem_alloc<512>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  if (kPair) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    {  // whole warp, one elected lane issues (see elect_one)
      if (elect_one()) {
        if (kPair) {
          if (rank == 0) mbar_expect_tx(wfull, 2 * 9 * wtap);
          const uint32_t wb = mapa_u32(wfull, 0);
          for (int tap = 0; tap < 9; ++tap)
            tma_load_2d_pair(wres + tap * wtap, &tmB, wb, tap * 64, n_blk * bn + (int)rank * (bn / 2));
        } else {
          mbar_expect_tx(wfull, kHaloWBytes);
          for (int tap = 0; tap < 9; ++tap) tma_load_2d(wres + tap * 8192, &tmB, wfull, tap * 64, n_blk * 64);
        }
      }
      __syncwarp();
      int stage = 0;
      uint32_t phase = 0;
      for (int mi = m_first; mi < m_end; mi += m_step) {
        const PatchOrigin po = conv_patch_origin(g, m_of(mi), 0);
        mbar_wait(&hempty[stage], phase ^ 1);
        if (elect_one()) {
          if (kPair) {
            if (rank == 0) mbar_expect_tx(&hfull[stage], 2 * kHaloBytes);
            tma_load_4d_pair(halo + stage * kHaloBytes, &tmA, mapa_u32(&hfull[stage], 0), 0, po.w - 1, po.h - 1, po.n);
          } else {
            mbar_expect_tx(&hfull[stage], kHaloBytes);
            tma_load_4d(halo + stage * kHaloBytes, &tmA, &hfull[stage], 0, po.w - 1, po.h - 1, po.n);
          }
        }
        __syncwarp();
        if (++stage == kHaloStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // MMA issuer: the whole warp walks the schedule (uniform control flow), one elected lane issues
    if (rank == 0) {
      const uint32_t idesc = make_idesc_bf16(kPair ? 2 * kBM : kBM, bn, false, false);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      mbar_wait(wfull, 0);
      tc_fence_after();
      const uint32_t wbase = smem_u32(wres);
      for (int mi = m_first; mi < m_end; mi += m_step) {
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        mbar_wait(&hfull[stage], phase);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * acc_stride;
        const uint32_t hbase = smem_u32(halo + stage * kHaloBytes);
        if (elect_one()) {
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            const int fr = tap / 3, fs = tap - fr * 3;
            const uint64_t adesc = make_smem_desc(hbase + (fr * kHaloLineRows + fs) * 128, 16u, kHaloLineRows * 128);
            const uint64_t bdesc = make_smem_desc(wbase + tap * wtap, 16u, 1024);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (kPair)
                umma_bf16_pair(d_tmem, adesc + uint64_t(k * 2), bdesc + uint64_t(k * 2), idesc, (tap > 0 || k > 0) ? 1u : 0u);
              else
                umma_bf16(d_tmem, adesc + uint64_t(k * 2), bdesc + uint64_t(k * 2), idesc, (tap > 0 || k > 0) ? 1u : 0u);
            }
          }
          if (kPair) {
            umma_commit_pair(&hempty[stage]);
            umma_commit_pair(&tfull[acc]);
          } else {
            umma_commit(&hempty[stage]);
            umma_commit(&tfull[acc]);
          }
        }
        __syncwarp();
        if (++stage == kHaloStages) {
          stage = 0;
          phase ^= 1;
        }
        if (++acc == nacc) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    const int grp = (warp - 4) >> 2;
    int epi_phase = 0;
    uint8_t* grp_smem = epi_smem + grp * (Epi::kSmemBytes / 2);
    int ord = 0;
    const uint32_t tempty_leader = kPair ? mapa_u32(tempty, 0) : 0u;
    for (int mi = m_first; mi < m_end; mi += m_step, ++ord) {
      if ((ord & 1) != grp) continue;
      const int acc = ord % nacc;
      const uint32_t acc_phase = (ord / nacc) & 1;
      TileCoord t;
      t.m_blk = m_of(mi); t.n_blk = n_blk; t.split = 0; t.kb_begin = 0; t.kb_end = 9;
      [[maybe_unused]] typename epi_prefetches<Epi>::type pre;
      if constexpr (epi_prefetches<Epi>::value) epi.prefetch(g, t, q * 32 + lane, pre);
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * acc_stride + (uint32_t(q * 32) << 16);
      if constexpr (epi_prefetches<Epi>::value) epi(taddr, g, t, q * 32 + lane, grp_smem, grp, epi_phase, &pre);
      else epi(taddr, g, t, q * 32 + lane, grp_smem, grp, epi_phase);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (kPair) mbar_arrive_cluster(tempty_leader + acc * 8);
        else mbar_arrive(&tempty[acc]);
      }
    }
    epi.finish(grp_smem, epi_phase);
  }
  tc_fence_before();
  if (kPair) cluster_sync_all();
  else __syncthreads();
  if (warp == 2) {
    if (kPair) tmem_dealloc_pair<512>(tmem_base);
    else tmem_dealloc<512>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------
// Halo form for Cin = 128 / 256 with a narrow output (Cout <= 128): conv2_2, and the input gradients of conv2_1,
// conv2_2 and conv3_1 (which run through the forward kernel on tap-reversed weights).
//
// With Cout <= 128 the generic A_CONV3x3 path is bound by L2 -> SM operand traffic, not by the tensor pipe: every
// 128-pixel tile re-reads its input patch once per filter tap (9 x Cin/64 x 16 KB) next to the [bn, 9*Cin] filter
// slice (conv2_2: 295 KB + 295 KB per 37.7 MFLOP tile, ~19 TB/s at tensor speed). Here the patch + halo of each
// 64-channel slice is staged ONCE per tile (one 4-D TMA box, same layout and shifted-descriptor trick as
// conv_halo_kernel) and only the filter slice streams through a second ring, one [bn x 64] block per (channel slice,
// tap): 74 KB + 295 KB per tile for conv2_2. Two producer lanes (warp 0: halos, warp 3: filter blocks) keep the two
// rings independent. The contraction order is (channel slice, tap, k) instead of (tap, channel slice, k): the same
// products, summed in a different order in fp32.
// Ring depths: a filter block is consumed in 4 MMAs (256 cycles at bn = 128, 128 at bn = 64), far less than the L2
// latency, so the filter ring must hold many blocks in flight (an ncu capture with 4 stages showed the MMA warp
// waiting on it 28 % of the time); a halo lasts 9 blocks, two stages are enough.
constexpr int kHaloSStages = 2;
constexpr int kHaloBMaxStages = 12;

__host__ __device__ inline int conv_halo_stream_b_stages(int bn, int epi_bytes) {
  const int n = (227 * 1024 - 1024 - 512 - epi_bytes - kHaloSStages * kHaloBytes) / (bn * 128);
  return n > kHaloBMaxStages ? kHaloBMaxStages : n;
}
__host__ inline int conv_halo_stream_smem_bytes(int bn, int epi_bytes) {
  return kHaloSStages * kHaloBytes + conv_halo_stream_b_stages(bn, epi_bytes) * bn * 128 + epi_bytes + 1024 + 512;
}

// kPair = 1: CTA pairs (see gemm_tc_kernel<Epi, 1>): pair tile = m-tiles 2p, 2p + 1; each CTA stages its own halos and
// bn / 2 rows of every filter block, so the filter ring holds twice as many blocks.
template <class Epi, int kPair>
__global__ void __launch_bounds__(kGemmThreads, 1)
conv_halo_stream_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmCore g,
                        const __grid_constant__ Epi epi) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int b_rows = kPair ? g.bn / 2 : g.bn;
  const int b_bytes = b_rows * 128;
  uint8_t* halo = smem;
  const int b_stages = conv_halo_stream_b_stages(b_rows, Epi::kSmemBytes);
  uint8_t* bring = smem + kHaloSStages * kHaloBytes;
  uint8_t* epi_smem = bring + b_stages * b_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_smem + Epi::kSmemBytes);
  uint64_t* hfull = bars;
  uint64_t* hempty = hfull + kHaloSStages;
  uint64_t* bfull = hempty + kHaloSStages;
  uint64_t* bempty = bfull + kHaloBMaxStages;
  uint64_t* tfull = bempty + kHaloBMaxStages;
  uint64_t* tempty = tfull + kMaxAcc;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + kMaxAcc);
  const int nacc = gemm_acc_buffers(g.bn);
  const int acc_stride = gemm_acc_stride(g.bn);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // pair mode: the schedule runs over pair tiles and clusters; a surplus tile of the last pair is all zero fill
  const uint32_t rank = kPair ? cluster_ctarank() : 0u;
  const int total_tiles = (kPair ? (g.m_tiles + 1) / 2 : g.m_tiles) * g.n_tiles;
  const int tile0 = kPair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int tile_step = kPair ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  auto m_of = [&](int tile) { return kPair ? 2 * (tile / g.n_tiles) + (int)rank : tile / g.n_tiles; };
  const int cin = g.cpk * 64;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kHaloSStages; ++i) {
      mbar_init(&hfull[i], 1);
      mbar_init(&hempty[i], 1);
    }
    for (int i = 0; i < kHaloBMaxStages; ++i) {
      mbar_init(&bfull[i], 1);
      mbar_init(&bempty[i], 1);
    }
    for (int i = 0; i < kMaxAcc; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], kPair ? 8 : 4);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    if (kPair) tmem_alloc_pair<512>(tmem_slot);
    else tmem_alloc<512>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  if (kPair) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ halo producer: one box per (tile, channel slice)
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = tile0; tile < total_tiles; tile += tile_step) {
        const PatchOrigin po = conv_patch_origin(g, m_of(tile), 0);
        for (int c = 0; c < g.cpk; ++c) {
          mbar_wait(&hempty[stage], phase ^ 1);
          if (elect_one()) {
            if (kPair) {
              if (rank == 0) mbar_expect_tx(&hfull[stage], 2 * kHaloBytes);
              tma_load_4d_pair(halo + stage * kHaloBytes, &tmA, mapa_u32(&hfull[stage], 0), c * 64, po.w - 1, po.h - 1, po.n);
            } else {
              mbar_expect_tx(&hfull[stage], kHaloBytes);
              tma_load_4d(halo + stage * kHaloBytes, &tmA, &hfull[stage], c * 64, po.w - 1, po.h - 1, po.n);
            }
          }
          __syncwarp();
          if (++stage == kHaloSStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 3) {
    // ------------------------------------------------------------ filter producer: one [bn x 64] block per (slice, tap)
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = tile0; tile < total_tiles; tile += tile_step) {
        const int n0 = (tile % g.n_tiles) * g.bn;
        for (int c = 0; c < g.cpk; ++c) {
          for (int tap = 0; tap < 9; ++tap) {
            mbar_wait(&bempty[stage], phase ^ 1);
            if (elect_one()) {
              if (kPair) {
                if (rank == 0) mbar_expect_tx(&bfull[stage], 2 * b_bytes);
                tma_load_2d_pair(bring + stage * b_bytes, &tmB, mapa_u32(&bfull[stage], 0), tap * cin + c * 64,
                                 n0 + (int)rank * b_rows);
              } else {
                mbar_expect_tx(&bfull[stage], b_bytes);
                tma_load_2d(bring + stage * b_bytes, &tmB, &bfull[stage], tap * cin + c * 64, n0);
              }
            }
            __syncwarp();
            if (++stage == b_stages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (whole warp, one elected lane issues)
    if (rank == 0) {
      const uint32_t idesc = make_idesc_bf16(kPair ? 2 * kBM : kBM, g.bn, false, false);
      int hs = 0, bs = 0, acc = 0;
      uint32_t hphase = 0, bphase = 0, acc_phase = 0;
      for (int tile = tile0; tile < total_tiles; tile += tile_step) {
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * acc_stride;
        for (int c = 0; c < g.cpk; ++c) {
          mbar_wait(&hfull[hs], hphase);
          const uint32_t hbase = smem_u32(halo + hs * kHaloBytes);
#pragma unroll 1
          for (int tap = 0; tap < 9; ++tap) {
            mbar_wait(&bfull[bs], bphase);
            tc_fence_after();
            const int fr = tap / 3, fs = tap - fr * 3;
            const uint64_t adesc = make_smem_desc(hbase + (fr * kHaloLineRows + fs) * 128, 16u, kHaloLineRows * 128);
            const uint64_t bdesc = make_smem_desc(smem_u32(bring + bs * b_bytes), 16u, 1024);
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                if (kPair)
                  umma_bf16_pair(d_tmem, adesc + uint64_t(k * 2), bdesc + uint64_t(k * 2), idesc, (c > 0 || tap > 0 || k > 0) ? 1u : 0u);
                else
                  umma_bf16(d_tmem, adesc + uint64_t(k * 2), bdesc + uint64_t(k * 2), idesc, (c > 0 || tap > 0 || k > 0) ? 1u : 0u);
              }
              if (kPair) umma_commit_pair(&bempty[bs]);
              else umma_commit(&bempty[bs]);
            }
            __syncwarp();
            if (++bs == b_stages) {
              bs = 0;
              bphase ^= 1;
            }
          }
          if (elect_one()) {
            if (kPair) umma_commit_pair(&hempty[hs]);
            else umma_commit(&hempty[hs]);
          }
          __syncwarp();
          if (++hs == kHaloSStages) {
            hs = 0;
            hphase ^= 1;
          }
        }
        if (elect_one()) {
          if (kPair) umma_commit_pair(&tfull[acc]);
          else umma_commit(&tfull[acc]);
        }
        __syncwarp();
        if (++acc == nacc) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    const int grp = (warp - 4) >> 2;
    int epi_phase = 0;
    uint8_t* grp_smem = epi_smem + grp * (Epi::kSmemBytes / 2);
    int ord = 0;
    const uint32_t tempty_leader = kPair ? mapa_u32(tempty, 0) : 0u;
    for (int tile = tile0; tile < total_tiles; tile += tile_step, ++ord) {
      if ((ord & 1) != grp) continue;
      const int acc = ord % nacc;
      const uint32_t acc_phase = (ord / nacc) & 1;
      TileCoord t;
      t.m_blk = m_of(tile); t.n_blk = tile % g.n_tiles; t.split = 0; t.kb_begin = 0; t.kb_end = 9 * g.cpk;
      [[maybe_unused]] typename epi_prefetches<Epi>::type pre;
      if constexpr (epi_prefetches<Epi>::value) epi.prefetch(g, t, q * 32 + lane, pre);
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * acc_stride + (uint32_t(q * 32) << 16);
      if constexpr (epi_prefetches<Epi>::value) epi(taddr, g, t, q * 32 + lane, grp_smem, grp, epi_phase, &pre);
      else epi(taddr, g, t, q * 32 + lane, grp_smem, grp, epi_phase);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (kPair) mbar_arrive_cluster(tempty_leader + acc * 8);
        else mbar_arrive(&tempty[acc]);
      }
    }
    epi.finish(grp_smem, epi_phase);
  }
  tc_fence_before();
  if (kPair) cluster_sync_all();
  else __syncthreads();
  if (warp == 2) {
    if (kPair) tmem_dealloc_pair<512>(tmem_base);
    else tmem_dealloc<512>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------
// Filter gradient of a 3x3 SAME convolution in halo form, for the wide feature maps with few channels (conv1_2,
// conv2_1, conv2_2: Cin, Cout <= 128), where gemm_tc_kernel's A_WGRAD3x3 mode is bound by L2 -> SM operand traffic:
// it gathers a fresh shifted 64-pixel x patch per (tap, 64-channel chunk) pair and re-reads the dy patch for every
// 128-row tile of dW (conv1_2: 24 KB of operands per 128-cycle k-block = 187 B/cycle per SM, three times what an SM can
// pull from L2).
//
// One CTA owns dW[9 taps x 64 cin (chunk c)][64 cout (chunk n)] -- five [128 x 64] accumulators (tap pairs; the ninth
// tap is paired with itself and its copy is dropped), 320 TMEM columns -- over a range of 8 x 8-pixel patches
// (split-K across CTAs, fp32 atomics at the end). Per patch it loads ONE x box with the halo (64 ch x 16 w x 10 h,
// 16-pixel line pitch as in conv_halo_kernel) and ONE dy box (64 ch x 8 x 8): 28 KB for 20 MMAs (640 cycles). Both
// operands are MN-major (rows = pixels = the contraction). A tap's A operand is the halo buffer read from a shifted
// start row (fr * 16 + fs): 8-pixel k-groups are one line pitch (2048 B) apart, and the two taps of a pair are the two
// 64-row halves of the UMMA M extent, their distance carried by the descriptor's leading-dimension byte offset.
constexpr int kWgHaloBytes = kHaloLineRows * 10 * 128;  // 20,480: x box 16 w x 10 h pixels x 64 channels
constexpr int kWgDyBytes = 64 * 128;                    // 8,192: dy box 8 x 8 pixels x 64 channels
constexpr int kWgStageBytes = kWgHaloBytes + kWgDyBytes;
constexpr int kWgStages = 7;
// kWide (Cout % 128 == 0): the CTA owns 128 output channels (two dy boxes, N = 128 MMAs: 8 KB of operands read from shared
// memory per 64 tensor cycles instead of 6 KB per 32) and HALF of the taps, because 5 accumulators of 128 columns exceed
// the 512 TMEM columns: half 0 = taps 0..5 (three pairs), half 1 = taps 6, 7, 8 (pair (6, 7) and 8 paired with itself).
constexpr int kWgWideStageBytes = kWgHaloBytes + 2 * kWgDyBytes;
constexpr int kWgWideStages = 6;

struct WgradHaloArgs {
  float* dw;          // [9 * Cin, Cout] fp32, accumulated with atomics
  int Cin, Cout;
  int tiles_w, tiles_h, n_img;  // 8 x 8 patches per image row / column (ragged edges: TMA zero fill)
  int k_total, splits;
};
__host__ __device__ constexpr int wg_stage_bytes(bool wide) { return wide ? kWgWideStageBytes : kWgStageBytes; }
__host__ __device__ constexpr int wg_stages(bool wide) { return wide ? kWgWideStages : kWgStages; }

template <bool kWide>
static __global__ void __launch_bounds__(256, 1)
conv_wgrad_halo_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDy,
                       const WgradHaloArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int kStages = wg_stages(kWide), kStageBytes = wg_stage_bytes(kWide);
  constexpr int kBN = kWide ? 128 : 64;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kStages;
  uint64_t* tfull = bars + 2 * kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int unit0 = blockIdx.x / a.splits;
  const int split = blockIdx.x - unit0 * a.splits;
  const int tap_half = kWide ? (unit0 & 1) : 0;
  const int unit = kWide ? (unit0 >> 1) : unit0;
  const int n_chunks = a.Cout / kBN;
  const int c_chunk = unit / n_chunks;
  const int n_chunk = unit - c_chunk * n_chunks;
  // tap pairs of this CTA: pair p covers taps (tap0 + 2p, tap0 + 2p + 1); the last pair of the last group is (8, 8)
  const int tap0 = kWide ? tap_half * 6 : 0;
  const int n_pairs = kWide ? (tap_half == 0 ? 3 : 2) : 5;
  const int kps = (a.k_total + a.splits - 1) / a.splits;
  const int kb_begin = split * kps;
  const int kb_end = min(a.k_total, kb_begin + kps);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmDy);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(tfull, 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        const int tiw = kb % a.tiles_w;
        const int r = kb / a.tiles_w;
        const int w0 = tiw * 8, h0 = (r % a.tiles_h) * 8, n0 = r / a.tiles_h;
        mbar_wait(&empty[stage], phase ^ 1);
        uint8_t* sx = smem + stage * kStageBytes;
        if (elect_one()) {
          mbar_expect_tx(&full[stage], kStageBytes);
          tma_load_4d(sx, &tmX, &full[stage], c_chunk * 64, w0 - 1, h0 - 1, n0);
          tma_load_4d(sx + kWgHaloBytes, &tmDy, &full[stage], n_chunk * kBN, w0, h0, n0);
          if (kWide) tma_load_4d(sx + kWgHaloBytes + kWgDyBytes, &tmDy, &full[stage], n_chunk * kBN + 64, w0, h0, n0);
        }
        __syncwarp();
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    {  // whole warp walks the k-range, one elected lane issues
      const uint32_t idesc = make_idesc_bf16(kBM, kBN, true, true);
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t sx = smem_u32(smem + stage * kStageBytes);
        const uint64_t bdesc = make_smem_desc(sx + kWgHaloBytes, 8192u, 1024);
        if (elect_one()) {
#pragma unroll
          for (int p = 0; p < 5; ++p) {
            if (p >= n_pairs) break;
            const int ta = tap0 + 2 * p, tb = ta < 8 ? ta + 1 : 8;
            const int ra = (ta / 3) * kHaloLineRows + ta % 3, rb = (tb / 3) * kHaloLineRows + tb % 3;
            const uint64_t adesc = make_smem_desc(sx + ra * 128, uint32_t(rb - ra) * 128u, kHaloLineRows * 128);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(tmem_base + p * kBN, adesc + uint64_t(k * ((2 * kHaloLineRows * 128) >> 4)), bdesc + uint64_t(k * (2048 >> 4)),
                        idesc, (kb > kb_begin || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty[stage]);
        }
        __syncwarp();
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (elect_one()) umma_commit(tfull);
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue: 128 threads, one accumulator row each
    if (kb_end > kb_begin) {
      const int q = warp & 3;
      mbar_wait(tfull, 0);
      tc_fence_after();
      // tfull fires after the last MMA has read shared memory and every TMA load has landed: the operand ring is idle
      // and serves as the transpose staging (one 32 x 32 fp32 block per warp, see EpiStore), so that each atomic
      // instruction adds 32 consecutive floats of one dW row instead of one float in each of 32 rows.
      float* stage = reinterpret_cast<float*>(smem) + q * 1024;
#pragma unroll 1
      for (int p = 0; p < n_pairs; ++p) {
        const int tap = tap0 + 2 * p + (q >> 1);
        const bool keep = tap <= 8;  // the second half of the pair (8, 8) is a copy of tap 8 (warp-uniform)
        float* o = a.dw + ((long long)tap * a.Cin + c_chunk * 64 + (q & 1) * 32) * a.Cout + n_chunk * kBN + lane;
#pragma unroll 1
        for (int c = 0; c < kBN; c += 32) {
          float v[32];
          __syncwarp();
          tmem_ld32(tmem_base + (uint32_t(q * 32) << 16) + p * kBN + c, v);
          tmem_ld_wait();
          if (keep) {
#pragma unroll
            for (int j = 0; j < 32; ++j) stage[lane * 32 + (j ^ lane)] = v[j];
            __syncwarp();
#pragma unroll 4
            for (int r = 0; r < 32; ++r) atomicAdd(o + (long long)r * a.Cout + c, stage[r * 32 + (lane ^ r)]);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<512>(tmem_base);
}

}  // namespace vc
