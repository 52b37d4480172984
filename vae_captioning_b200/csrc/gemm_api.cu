// Generic bf16 GEMM entry (store epilogue). Internal C++ helper + the C-ABI test entry vc_gemm_bf16.
#include "host_util.h"
#include "ops.h"

namespace vc {

int gemm_store(cudaStream_t stream, const Operand& A, const Operand* A2, long long a2_at, const Operand& B, int M,
               int N, int K, const EpiStore& epi_in, int bn, int splits) {
  GemmPlan plan;
  VC_TRY(plan_gemm(&plan, A, A2, a2_at, B, M, N, K, bn, splits));
  // few n-tiles under a long contraction (dgrad of the vocab projection: 2 x 177 k-blocks): run them side by side
  plan.core.n_fast = (plan.core.n_tiles <= 4 && plan.core.k_blocks >= 64 && plan.core.m_tiles > plan.core.n_tiles) ? 1 : 0;
  EpiStore epi = epi_in;
  epi.M = M;
  epi.N = N;
  epi.bn = bn;
  if (plan.core.splits > 1 && !epi.atomic)
    return set_error(VC_E_ARG, "gemm_store: split-K requires the atomic epilogue");
  return launch_gemm(plan, epi, stream);
}

// out[M, N] (bf16, row pitch ldo) = act(A x B + bias) through the TMA-store epilogue: full 128-byte row segments
// leave the SM as bulk tensor stores instead of per-thread 64-byte pieces (used for the [T*N, V] logits).
int gemm_tma_rows(cudaStream_t stream, const Operand& A, const Operand& B, int M, int N, int K, void* out, long long ldo,
                  const float* bias, int relu, int bn) {
  if (bn % 64 != 0) return set_error(VC_E_ARG, "gemm_tma_rows: bn=%d must be a multiple of 64", bn);
  GemmPlan plan;
  VC_TRY(plan_gemm(&plan, A, nullptr, 0, B, M, N, K, bn, 1));
  EpiTma epi{};
  epi.bias = bias;
  epi.N = N;
  epi.bn = bn;
  epi.relu = relu;
  epi.mode = kRows;
  epi.alpha = 1.f;
  VC_TRY(make_tmap_2d(&epi.tm, out, (uint64_t)ldo, (uint64_t)M, (uint64_t)ldo, 64, 128));
  return launch_gemm(plan, epi, stream);
}

// out[M, N] (fp32, row pitch ldo) = A x B + bias through the fp32 TMA-store epilogue (decode: vocab logits per step).
int gemm_tma_rows_f32(cudaStream_t stream, const Operand& A, const Operand& B, int M, int N, int K, float* out, long long ldo,
                      const float* bias, int bn) {
  if (bn % 32 != 0) return set_error(VC_E_ARG, "gemm_tma_rows_f32: bn=%d must be a multiple of 32", bn);
  GemmPlan plan;
  VC_TRY(plan_gemm(&plan, A, nullptr, 0, B, M, N, K, bn, 1));
  EpiTmaF32 epi{};
  epi.bias = bias;
  epi.N = N;
  epi.bn = bn;
  epi.alpha = 1.f;
  VC_TRY(make_tmap_2d_f32(&epi.tm, out, (uint64_t)N, (uint64_t)M, (uint64_t)ldo, 128));
  return launch_gemm(plan, epi, stream);
}

}  // namespace vc

extern "C" int vc_gemm_bf16(const void* A, int a_mn, long long lda, const void* B, int b_mn, long long ldb, void* out,
                            long long ldo, const float* bias, int M, int N, int K, int bn, int splits, int relu,
                            int out_bf16, int atomic, void* stream) {
  using namespace vc;
  Operand a{A, a_mn ? K : M, a_mn ? M : K, lda, a_mn != 0};
  Operand b{B, b_mn ? K : N, b_mn ? N : K, ldb, b_mn != 0};
  EpiStore epi{};
  epi.out = out;
  epi.bias = bias;
  epi.ld = ldo;
  epi.relu = relu;
  epi.out_bf16 = out_bf16;
  epi.atomic = atomic;
  epi.alpha = 1.f;
  return gemm_store(static_cast<cudaStream_t>(stream), a, nullptr, 0, b, M, N, K, epi, bn, splits);
}

extern "C" void vc_test_pair_mode(int mode) { vc::set_pair_mode(mode); }
extern "C" void vc_test_dual_mode(int mode) { vc::set_dual_mode(mode); }

// Test entries for the convolution backward kernels (parity against torch conv2d gradients on the same bf16 inputs):
// x, dy bf16 NHWC; w fp32 HWIO; dw fp32 [9*Cin, Cout] (zeroed here); dx bf16 NHWC. All device pointers.
extern "C" int vc_conv3x3_bwd(const void* x, const void* dy, const float* w, float* dw, void* dx, int B, int hw, int cin,
                              int cout, void* stream) {
  using namespace vc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  void* wt_d = nullptr;
  VC_CUDA(cudaMalloc(&wt_d, (size_t)9 * cin * cout * 2));
  int st = dgrad_shadow(s, w, wt_d, cin, cout);
  if (st == VC_OK && cudaMemsetAsync(dw, 0, (size_t)9 * cin * cout * sizeof(float), s) != cudaSuccess)
    st = set_error(VC_E_CUDA, "memset failed");
  if (st == VC_OK) st = conv3x3_wgrad(s, x, dy, dw, B, hw, cin, cout);
  if (st == VC_OK) st = conv3x3_dgrad(s, dy, wt_d, dx, B, hw, cin, cout);
  cudaStreamSynchronize(s);
  cudaFree(wt_d);
  return st;
}

extern "C" int vc_conv3x3_dgrad_relu(const void* dy, const float* w, const void* act, void* dx, float* dbias, int B, int hw,
                                     int cin, int cout, void* stream) {
  using namespace vc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  void* wt_d = nullptr;
  VC_CUDA(cudaMalloc(&wt_d, (size_t)9 * cin * cout * 2));
  int st = dgrad_shadow(s, w, wt_d, cin, cout);
  if (st == VC_OK) st = conv3x3_dgrad(s, dy, wt_d, dx, B, hw, cin, cout, "conv_dgrad", act, dbias);
  cudaStreamSynchronize(s);
  cudaFree(wt_d);
  return st;
}

// db: fp32 [C] device, accumulated into (the bias gradient = per-channel sum of dY)
extern "C" int vc_relu_pool_bwd(const void* dA, const void* out, void* dY, float* db, int B, int hw, int C, int pooled,
                                void* stream) {
  return vc::relu_pool_bwd(static_cast<cudaStream_t>(stream), dA, out, dY, B, hw, C, pooled != 0, db);
}
