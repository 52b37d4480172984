#include "host_util.h"
#include <cstdarg>
#include <mutex>

namespace vc {

std::string& last_error() {
  static thread_local std::string msg;
  return msg;
}

int set_error(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  last_error() = buf;
  return code;
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int make_tmap_2d(CUtensorMap* out, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner,
                 uint32_t box_outer) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return set_error(VC_E_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || ((ld * 2) & 15))
    return set_error(VC_E_ARG, "TMA operand must be 16-byte aligned with a 16-byte row pitch (ptr=%p ld=%llu)", ptr,
                     (unsigned long long)ld);
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(VC_E_CUDA, "cuTensorMapEncodeTiled(2d inner=%llu outer=%llu ld=%llu box=%ux%u) -> %d",
                     (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)ld, box_inner, box_outer,
                     (int)r);
  return VC_OK;
}

int make_tmap_nhwc(CUtensorMap* out, const void* ptr, int C, int W, int H, int N, int bw, int bh, int bi) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return set_error(VC_E_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)C * W * 2, (cuuint64_t)C * W * H * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bi};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(VC_E_CUDA, "cuTensorMapEncodeTiled(nhwc C=%d W=%d H=%d N=%d box=%dx%dx%d) -> %d", C, W, H, N, bw,
                     bh, bi, (int)r);
  return VC_OK;
}

static int operand_tmap(CUtensorMap* out, const Operand& o, int box_rows_kmajor) {
  if (o.mn_major) return make_tmap_2d(out, o.ptr, o.cols, o.rows, o.ld, 64, 64);
  return make_tmap_2d(out, o.ptr, o.cols, o.rows, o.ld, 64, box_rows_kmajor);
}

int plan_gemm(GemmPlan* p, const Operand& A, const Operand* A2, long long a2_at, const Operand& B, int M, int N, int K,
              int bn, int splits) {
  if (bn < 16 || bn > 256 || (bn % 16) != 0 || (B.mn_major && (bn % 64) != 0))
    return set_error(VC_E_ARG, "plan_gemm: unsupported tile width bn=%d", bn);
  if (M <= 0 || N <= 0 || K <= 0) return set_error(VC_E_SHAPE, "plan_gemm: empty problem %dx%dx%d", M, N, K);
  memset(p, 0, sizeof(*p));
  GemmCore& g = p->core;
  g.m_tiles = (M + kBM - 1) / kBM;
  g.n_tiles = (N + bn - 1) / bn;
  g.k_blocks = (K + kBK - 1) / kBK;
  g.splits = splits < 1 ? 1 : (splits > g.k_blocks ? g.k_blocks : splits);
  g.bn = bn;
  g.stages = gemm_pick_stages(bn);
  g.a_mode = A.mn_major ? A_MNMAJOR : A_KMAJOR;
  g.b_mn = B.mn_major ? 1 : 0;
  g.a_switch = -1;
  VC_TRY(operand_tmap(&p->tmA, A, kBM));
  p->tmA2 = p->tmA;
  if (A2 != nullptr && A2->ptr != nullptr) {
    if (A2->mn_major != A.mn_major) return set_error(VC_E_ARG, "plan_gemm: A and A2 must share major-ness");
    if (a2_at % (A.mn_major ? kBM : kBK) != 0)
      return set_error(VC_E_ARG, "plan_gemm: A2 switch point %lld not tile aligned", a2_at);
    g.a_switch = (int)(a2_at / (A.mn_major ? kBM : kBK));
    VC_TRY(operand_tmap(&p->tmA2, *A2, kBM));
  }
  VC_TRY(operand_tmap(&p->tmB, B, bn));
  return VC_OK;
}

}  // namespace vc
