// Host-side helpers: error plumbing, TMA tensor-map encoding (driver entry point fetched at run
// time, so the library carries no link-time dependency on libcuda), GEMM launch.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
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

// 2-D bf16 tensor map, 128B swizzle. inner = contiguous extent (elements), outer = rows,
// ld = row pitch in elements (ld*2 must be a multiple of 16 bytes).
int make_tmap_2d(CUtensorMap* out, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner,
                 uint32_t box_outer);
// 4-D bf16 NHWC tensor map {C, W, H, N} with box {64, bw, bh, bi}, 128B swizzle.
int make_tmap_nhwc(CUtensorMap* out, const void* ptr, int C, int W, int H, int N, int bw, int bh, int bi);

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
int plan_gemm(GemmPlan* p, const Operand& A, const Operand* A2, long long a2_at, const Operand& B, int M, int N, int K,
              int bn, int splits);

template <class Epi>
int launch_gemm(const GemmPlan& p, const Epi& epi, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    VC_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<Epi>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  const int total = p.core.m_tiles * p.core.n_tiles * p.core.splits;
  if (total <= 0) return VC_OK;
  const int grid = total < num_sms() ? total : num_sms();
  const int smem = gemm_smem_bytes(p.core.bn, p.core.stages);
  gemm_tc_kernel<Epi><<<grid, kGemmThreads, smem, stream>>>(p.tmA, p.tmA2, p.tmB, p.core, epi);
  VC_CUDA(cudaGetLastError());
  return VC_OK;
}

}  // namespace vc
