// Host-side helpers: error plumbing, TMA tensor-map encoding (driver entry point fetched at run
// time, so the library carries no link-time dependency on libcuda), GEMM launch.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>
#include "../../include/vaecap.h"
#include "gemm.cuh"

namespace vc {

// Status codes VC_OK / VC_E_* come from the C ABI header (include/vaecap.h).

std::string& last_error();  // thread-local message of the last failing call
int set_error(int code, const char* fmt, ...);

#define VC_CUDA(expr)                                                                            \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess)                                                                       \
      return ::vc::set_error(VC_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                             __LINE__);                                                          \
  } while (0)
#define VC_TRY(expr)           \
  do {                         \
    int _s = (expr);           \
    if (_s != 0) return _s;    \
  } while (0)

int num_sms();
// SMs a persistent GEMM launched inside a GridCap scope may occupy (default: all of them). The backward pass runs the
// weight-gradient GEMMs nothing downstream waits for on a side stream next to the whole-sequence BPTT kernels, which hold
// (N/128) x (H/units) SMs for their whole duration: a 148-CTA grid would queue behind them, a capped one fits beside them.
int sm_budget();
struct GridCap {
  int prev;
  explicit GridCap(int sms);
  ~GridCap();
};
bool prof_is_on();

// ---- launch accounting + optional per-kernel-family CUDA-event timing (bench.py's roofline leg).
// Every kernel launch site opens a ProfScope: it counts the launch and, while profiling is enabled,
// brackets it with two events on the launching stream. ProfTag overrides the family name for the
// launches made inside its lifetime (e.g. "lstm_fwd" around the generic GEMM launcher).
unsigned long long launch_count();
void prof_enable(bool on);                       // clears previous records when switched on
int prof_collect(char* names, int names_cap, float* ms, int* counts, int cap);  // -> number of families
struct ProfTag {
  const char* prev;
  explicit ProfTag(const char* tag);
  ~ProfTag();
};
struct ProfScope {
  cudaStream_t s;
  int slot;
  ProfScope(cudaStream_t stream, const char* default_tag);
  ~ProfScope();
};

// 2-D bf16 tensor map, 128B swizzle. inner = contiguous extent (elements), outer = rows,
// ld = row pitch in elements (ld*2 must be a multiple of 16 bytes).
int make_tmap_2d(CUtensorMap* out, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner,
                 uint32_t box_outer);
// fp32 [outer, inner] row-major (row pitch ld floats), box {32, box_outer}, 128B swizzle (store side of EpiTmaF32)
int make_tmap_2d_f32(CUtensorMap* out, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_outer);
// 3-D maps {inner, rows, slots} for tensors stored as [slots][rows][ld]: a box never crosses a slot, rows past `rows` are
// zero on load and dropped on store (the 2-D maps above run straight into the next slot). Box {box_inner, box_rows, 1}
// with rows of 128 bytes (128B swizzle) or 64 bytes (64B swizzle).
int make_tmap_3d(CUtensorMap* out, const void* ptr, uint64_t inner, uint64_t rows, uint64_t slots, uint64_t ld,
                 uint64_t slot_pitch, uint32_t box_inner, uint32_t box_rows, bool f32);
// 4-D bf16 NHWC tensor map {C, W, H, N} with box {64, bw, bh, bi}, 128B swizzle.
int make_tmap_nhwc(CUtensorMap* out, const void* ptr, int C, int W, int H, int N, int bw, int bh, int bi);

// Patch / tile geometry of a 3x3 SAME convolution over an NHWC tensor (see GemmCore).
struct ConvGeom {
  int W, H, Nimg, Cin, Cout;
  int pw, ph, pn, tw, th;
};
// Picks the patch shape that tiles a W x H feature map exactly (224/112: 16x2x1, 56: 8x4x1, 28: 4x4x2, 14: 2x2x8).
int conv_geometry(ConvGeom* g, int W, int H, int Nimg, int Cin, int Cout);

// A GEMM operand. K-major: matrix [rows = M or N][cols = K]. MN-major: matrix [rows = K][cols = M or N].
struct Operand {
  const void* ptr;
  long long rows, cols, ld;
  bool mn_major;
};

struct GemmPlan {
  CUtensorMap tmA, tmA2, tmB;
  GemmCore core;
};

// Builds tensor maps + tile schedule for D[M,N] = A * B (contraction K). A2 optional (ptr == nullptr):
//   A K-major:  A covers k in [0, a2_at), A2 covers the rest (a2_at multiple of 64).
//   A MN-major: A covers m in [0, a2_at), A2 the rest (a2_at multiple of 128).
// pair_ok = false: never plan a CTA-pair launch (epilogues that are not tile-local)
int plan_gemm(GemmPlan* p, const Operand& A, const Operand* A2, long long a2_at, const Operand& B, int M, int N, int K,
              int bn, int splits, bool pair_ok = true);

// D = conv3x3_same(in NHWC bf16 [Nimg,H,W,Cin], Wt bf16 [Cout, 9*Cin] (tap-major, then cin)); Cin % 64 == 0.
int plan_conv(GemmPlan* p, const void* in, const void* wt, const ConvGeom& g, int bn);

// conv1_1 (Cin = 3) without an im2col matrix: `padded` is the mean-subtracted image as bf16 [Nimg, H+2, W+2, 8] (3 channels
// + 5 zeros per pixel, a zero border of one pixel). The 3 taps x 3 channels of one filter row are then 24 consecutive
// elements starting at pixel (x, y + r) of the padded image, so ONE overlapping-stride tensor map {64, W, H+2, Nimg} with a
// pixel stride of 16 bytes delivers a [128 pixels x 64] k-block per filter row (elements 24..63 run into the neighbouring
// pixels and meet zero weights). wt: bf16 [Cout, 192], k = r*64 + s*8 + c.
int plan_conv1_window(GemmPlan* p, const void* padded, const void* wt, int W, int H, int Nimg, int Cout);

// Filter gradient of the same convolution: dW[9*Cin, Cout] (HWIO row-major) += patches(in)^T x dY, contraction over
// the Nimg*H*W pixels in 64-pixel patches, split-K over `splits` CTAs per output tile (fp32 atomic epilogue).
// in: NHWC bf16 [Nimg,H,W,Cin]; dy: NHWC bf16 [Nimg,H,W,Cout]; Cin, Cout multiples of 64; bn divides Cout.
int plan_conv_wgrad(GemmPlan* p, const void* in, const void* dy, int W, int H, int Nimg, int Cin, int Cout, int bn,
                    int splits);

// Halo form for Cin == 64 layers whose feature map tiles into 8 x 16 pixel blocks (see conv_halo_kernel).
bool conv_halo_applicable(int W, int H, int Cin, int Cout);
int conv_halo_geometry(ConvGeom* g, int W, int H, int Nimg, int Cin, int Cout);
int plan_conv_halo(GemmPlan* p, const void* in, const void* wt, const ConvGeom& g);

// Halo-form filter gradient (see conv_wgrad_halo_kernel): x, dy bf16 NHWC; dw fp32 [9*Cin, Cout] accumulated with
// atomics (the caller zeroes it). VC_WGRAD_HALO=0 in the environment turns it off; VC_WGRAD_HALO_MAXCH caps the channel count.
bool conv_wgrad_halo_applicable(int W, int H, int Cin, int Cout);
int launch_conv_wgrad_halo(cudaStream_t stream, const void* x, const void* dy, float* dw, int W, int H, int Nimg, int Cin,
                           int Cout);

// Streamed-filter halo form for Cin in {128, 256} and Cout in {64, 128} (see conv_halo_stream_kernel); the tile width
// is the whole Cout. VC_CONV_HALO2=0 in the environment turns it off (the generic implicit-GEMM path is used instead).
bool conv_halo_stream_applicable(int W, int H, int Cin, int Cout);
int plan_conv_halo_stream(GemmPlan* p, const void* in, const void* wt, const ConvGeom& g);

// CTA-pair policy (gemm_tc_kernel<Epi, 1>): VC_PAIR=0 never, VC_PAIR=1 wherever the tile shape allows it, unset = where it
// pays (wide tiles, enough m-tiles that rounding their count up to even costs little). set_pair_mode overrides the
// environment (tests): -1 automatic, 0 off, 1 forced.
void set_pair_mode(int mode);
bool gemm_pair_wanted(int m_tiles, int bn, int b_mn, int k_blocks);
// Dual-N tiles (GemmCore::dual): VC_DUAL=0 never, VC_DUAL=1 wherever legal (bn = 256, >= 2 n-tiles), unset = long contractions
// only (the accumulators are not double-buffered, so the epilogue of a tile is exposed). set_dual_mode as set_pair_mode.
void set_dual_mode(int mode);
bool gemm_dual_wanted(int n_tiles, int bn, int k_blocks, int tiles_total, int units);
bool halo_pair_wanted(int bn);

// Launch of a kernel written for CTA pairs: clusters of 2, at most `pairs` of them and never more than can be co-resident.
template <class Kernel, class... Args>
int launch_pair_kernel(Kernel kernel, int* max_clusters, int pairs, int sm_cap, int smem, cudaStream_t stream, const char* tag,
                       const Args&... args) {
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  if (*max_clusters == 0) {
    VC_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    cfg.gridDim = dim3(num_sms() / 2 * 2);
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kernel, &cfg) != cudaSuccess || n <= 0) {
      cudaGetLastError();
      n = num_sms() / 2;
    }
    *max_clusters = n;
  }
  if (pairs <= 0) return VC_OK;
  const int cap = std::max(1, std::min(*max_clusters, sm_cap / 2));
  cfg.gridDim = dim3(2 * (pairs < cap ? pairs : cap));
  {
    ProfScope ps(stream, tag);
    VC_CUDA(cudaLaunchKernelEx(&cfg, kernel, args...));
  }
  VC_CUDA(cudaGetLastError());
  return VC_OK;
}

template <class Epi>
int launch_conv_halo_stream(const GemmPlan& p, const Epi& epi, cudaStream_t stream) {
  if (p.core.pair) {
    static int max_clusters = 0;
    return launch_pair_kernel(conv_halo_stream_kernel<Epi, 1>, &max_clusters, (p.core.m_tiles + 1) / 2 * p.core.n_tiles, num_sms(),
                              conv_halo_stream_smem_bytes(p.core.bn / 2, Epi::kSmemBytes), stream, "conv_halo_stream", p.tmA,
                              p.tmB, p.core, epi);
  }
  static bool configured = false;
  if (!configured) {
    VC_CUDA(cudaFuncSetAttribute(conv_halo_stream_kernel<Epi, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  const int total = p.core.m_tiles * p.core.n_tiles;
  if (total <= 0) return VC_OK;
  const int grid = total < num_sms() ? total : num_sms();
  const int smem = conv_halo_stream_smem_bytes(p.core.bn, Epi::kSmemBytes);
  {
    ProfScope ps(stream, "conv_halo_stream");
    conv_halo_stream_kernel<Epi, 0><<<grid, kGemmThreads, smem, stream>>>(p.tmA, p.tmB, p.core, epi);
  }
  VC_CUDA(cudaGetLastError());
  return VC_OK;
}

template <class Epi>
int launch_conv_halo(const GemmPlan& p, const Epi& epi, cudaStream_t stream) {
  const int nt = p.core.n_tiles;
  if (p.core.pair) {
    // a cluster keeps one n-tile: the cluster count is a multiple of the n-tile count (see the kernel's schedule)
    static int max_clusters = 0;
    if (max_clusters == 0) VC_TRY(launch_pair_kernel(conv_halo_kernel<Epi, 1>, &max_clusters, 0, num_sms(), conv_halo_smem_bytes(Epi::kSmemBytes), stream, "conv_halo", p.tmA, p.tmB, p.core, epi));
    const int pairs = (p.core.m_tiles + 1) / 2;
    int clusters = std::min(max_clusters, num_sms() / 2) / nt * nt;
    if (clusters > pairs * nt) clusters = pairs * nt;
    return launch_pair_kernel(conv_halo_kernel<Epi, 1>, &max_clusters, clusters, 2 * clusters, conv_halo_smem_bytes(Epi::kSmemBytes), stream,
                              "conv_halo", p.tmA, p.tmB, p.core, epi);
  }
  static bool configured = false;
  if (!configured) {
    VC_CUDA(cudaFuncSetAttribute(conv_halo_kernel<Epi, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  int grid = num_sms() / nt * nt;
  if (grid > p.core.m_tiles * nt) grid = p.core.m_tiles * nt;
  if (grid <= 0) return VC_OK;
  const int smem = conv_halo_smem_bytes(Epi::kSmemBytes);
  {
    ProfScope ps(stream, "conv_halo");
    conv_halo_kernel<Epi, 0><<<grid, kGemmThreads, smem, stream>>>(p.tmA, p.tmB, p.core, epi);
  }
  VC_CUDA(cudaGetLastError());
  return VC_OK;
}


template <class Epi>
int launch_gemm_pair(const GemmPlan& p, const Epi& epi, cudaStream_t stream) {
  static int max_clusters = 0;
  GemmCore core = p.core;
  const int b_rows = (core.dual ? 2 : 1) * core.bn / 2;  // B rows staged per CTA and ring stage
  const int n_sched = core.dual ? (core.n_tiles + 1) / 2 : core.n_tiles;
  core.stages = gemm_pick_stages(b_rows, Epi::kSmemBytes);
  if (core.stages < 2) return set_error(VC_E_ARG, "launch_gemm: tile too large for shared memory (bn=%d)", core.bn);
  return launch_pair_kernel(gemm_tc_kernel<Epi, 1>, &max_clusters, core.m_tiles / 2 * n_sched * core.splits, sm_budget(),
                            gemm_smem_bytes(b_rows, core.stages, Epi::kSmemBytes), stream, "gemm", p.tmA, p.tmA2, p.tmB, core, epi);
}

template <class Epi>
int launch_gemm(const GemmPlan& p, const Epi& epi, cudaStream_t stream) {
  if (p.core.pair) {
    if constexpr (Epi::kPairOk) return launch_gemm_pair(p, epi, stream);
    else return set_error(VC_E_ARG, "launch_gemm: this epilogue does not run as a CTA pair");
  }
  static bool configured = false;
  if (!configured) {
    VC_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<Epi, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  GemmCore core = p.core;
  const int b_rows = (core.dual ? 2 : 1) * core.bn;
  const int total = core.m_tiles * (core.dual ? (core.n_tiles + 1) / 2 : core.n_tiles) * core.splits;
  if (total <= 0) return VC_OK;
  const int grid = total < sm_budget() ? total : sm_budget();
  core.stages = gemm_pick_stages(b_rows, Epi::kSmemBytes);
  if (core.stages < 2) return set_error(VC_E_ARG, "launch_gemm: tile too large for shared memory (bn=%d)", core.bn);
  const int smem = gemm_smem_bytes(b_rows, core.stages, Epi::kSmemBytes);
  {
    ProfScope ps(stream, "gemm");
    gemm_tc_kernel<Epi, 0><<<grid, kGemmThreads, smem, stream>>>(p.tmA, p.tmA2, p.tmB, core, epi);
  }
  VC_CUDA(cudaGetLastError());
  return VC_OK;
}

}  // namespace vc
