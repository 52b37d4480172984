#include "host_util.h"
#include <cstdarg>
#include <cstdlib>
#include <algorithm>
#include <atomic>
#include <mutex>
#include <vector>

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

// ------------------------------------------------------------------------------------------
namespace {
struct ProfRec {
  const char* tag;
  cudaEvent_t e0, e1;
};
std::atomic<unsigned long long> g_launches{0};
bool g_prof_on = false;
std::vector<ProfRec> g_recs;
size_t g_used = 0;
thread_local const char* t_tag = nullptr;
}  // namespace

static thread_local int t_sm_cap = 0;
int sm_budget() { return t_sm_cap > 0 && t_sm_cap < num_sms() ? t_sm_cap : num_sms(); }
GridCap::GridCap(int sms) : prev(t_sm_cap) { t_sm_cap = sms; }
GridCap::~GridCap() { t_sm_cap = prev; }
bool prof_is_on() { return g_prof_on; }

unsigned long long launch_count() { return g_launches.load(); }

void prof_enable(bool on) {
  g_prof_on = on;
  if (on) g_used = 0;
}

ProfTag::ProfTag(const char* tag) : prev(t_tag) { t_tag = tag; }
ProfTag::~ProfTag() { t_tag = prev; }

ProfScope::ProfScope(cudaStream_t stream, const char* default_tag) : s(stream), slot(-1) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (!g_prof_on) return;
  if (g_used == g_recs.size()) {
    ProfRec r{};
    if (cudaEventCreate(&r.e0) != cudaSuccess || cudaEventCreate(&r.e1) != cudaSuccess) return;
    g_recs.push_back(r);
  }
  slot = (int)g_used++;
  g_recs[slot].tag = t_tag ? t_tag : default_tag;
  cudaEventRecord(g_recs[slot].e0, s);
}
ProfScope::~ProfScope() {
  if (slot >= 0) cudaEventRecord(g_recs[slot].e1, s);
}

int prof_collect(char* names, int names_cap, float* ms, int* counts, int cap) {
  cudaDeviceSynchronize();
  std::vector<std::string> fam;
  std::vector<float> t;
  std::vector<int> c;
  for (size_t i = 0; i < g_used; ++i) {
    float e = 0.f;
    if (cudaEventElapsedTime(&e, g_recs[i].e0, g_recs[i].e1) != cudaSuccess) continue;
    size_t j = 0;
    for (; j < fam.size(); ++j)
      if (fam[j] == g_recs[i].tag) break;
    if (j == fam.size()) {
      fam.push_back(g_recs[i].tag);
      t.push_back(0.f);
      c.push_back(0);
    }
    t[j] += e;
    c[j] += 1;
  }
  std::string joined;
  int n = 0;
  for (size_t j = 0; j < fam.size() && n < cap; ++j, ++n) {
    if (!joined.empty()) joined += ",";
    joined += fam[j];
    ms[n] = t[j];
    counts[n] = c[j];
  }
  if (names_cap > 0) {
    strncpy(names, joined.c_str(), names_cap - 1);
    names[names_cap - 1] = 0;
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

int make_tmap_2d_f32(CUtensorMap* out, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_outer) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return set_error(VC_E_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || ((ld * 4) & 15))
    return set_error(VC_E_ARG, "TMA operand must be 16-byte aligned with a 16-byte row pitch (ptr=%p ld=%llu)", ptr,
                     (unsigned long long)ld);
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld * 4};
  cuuint32_t box[2] = {32, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(VC_E_CUDA, "cuTensorMapEncodeTiled(f32 inner=%llu outer=%llu ld=%llu) -> %d", (unsigned long long)inner,
                     (unsigned long long)outer, (unsigned long long)ld, (int)r);
  return VC_OK;
}

int make_tmap_3d(CUtensorMap* out, const void* ptr, uint64_t inner, uint64_t rows, uint64_t slots, uint64_t ld,
                 uint64_t slot_pitch, uint32_t box_inner, uint32_t box_rows, bool f32) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return set_error(VC_E_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  const uint64_t es = f32 ? 4 : 2;
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || ((ld * es) & 15) || ((slot_pitch * es) & 15))
    return set_error(VC_E_ARG, "TMA operand must be 16-byte aligned with 16-byte pitches (ptr=%p ld=%llu)", ptr,
                     (unsigned long long)ld);
  cuuint64_t dims[3] = {inner, rows, slots};
  cuuint64_t strides[2] = {ld * es, slot_pitch * es};
  if (box_inner * es != 128 && box_inner * es != 64) return set_error(VC_E_ARG, "make_tmap_3d: box rows must be 64 or 128 bytes");
  cuuint32_t box[3] = {box_inner, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr),
                  dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  box_inner * es == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,  // swizzle span = box row
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(VC_E_CUDA, "cuTensorMapEncodeTiled(3d inner=%llu rows=%llu slots=%llu ld=%llu) -> %d",
                     (unsigned long long)inner, (unsigned long long)rows, (unsigned long long)slots, (unsigned long long)ld, (int)r);
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

// ------------------------------------------------------------------------------------------
static int g_pair_mode = -2;  // -2: read VC_PAIR on first use
void set_pair_mode(int mode) { g_pair_mode = mode < -1 || mode > 1 ? -1 : mode; }
bool gemm_pair_wanted(int m_tiles, int bn, int b_mn, int k_blocks) {
  if (g_pair_mode == -2) {
    const char* e = getenv("VC_PAIR");
    g_pair_mode = (e && e[0] == '0') ? 0 : (e && e[0] == '1') ? 1 : -1;
  }
  if (g_pair_mode == 0) return false;
  // M = 256 needs N % 16 == 0; each CTA stages bn / 2 rows (8-row swizzle groups) or bn / 2 columns in 64-wide MN blocks
  const bool legal = m_tiles >= 2 && bn >= 32 && bn % 32 == 0 && (b_mn == 0 || bn % 128 == 0);
  if (!legal) return false;
  if (g_pair_mode == 1) return true;
  // automatic: everywhere except where rounding an odd m-tile count up to even wastes more than a tenth of the work
  (void)k_blocks;
  return m_tiles % 2 == 0 || m_tiles >= 9;
}

static int g_dual_mode = -2;
void set_dual_mode(int mode) { g_dual_mode = mode < -1 || mode > 1 ? -1 : mode; }
bool gemm_dual_wanted(int n_tiles, int bn, int k_blocks, int tiles_total, int units) {
  if (g_dual_mode == -2) {
    const char* e = getenv("VC_DUAL");
    g_dual_mode = (e && e[0] == '0') ? 0 : (e && e[0] == '1') ? 1 : -1;
  }
  if (g_dual_mode == 0 || bn != 256 || n_tiles < 2) return false;
  if (g_dual_mode == 1) return true;
  // automatic: the tile's epilogue (two 128 x 256 halves, one per epilogue group) is not overlapped with the next tile's
  // MMAs, so the contraction must be long; an odd n-tile count wastes half a tile; and the halved tile count must still
  // fill the machine
  // (tiles_total counts schedule units -- pair tiles or tiles -- before halving; `units` is what runs side by side)
  if (!(k_blocks >= 32 && (n_tiles % 2 == 0 || n_tiles >= 9))) return false;
  const double waves = 0.5 * tiles_total / (double)units;
  const double eff = waves / (double)(long long)(waves + 0.999999);
  // the vocabulary input gradient (100 dual tiles on 74 clusters: 1.35 waves) loses more to the second, two-thirds-empty
  // wave than the tile gains: 0.317 -> 0.323 ms; conv4_x (10.6 waves) 0.80 -> 0.64 ms, conv5_x (2.6 waves) 0.23 -> 0.18
  return eff >= 0.85;
}

bool halo_pair_wanted(int bn) {
  gemm_pair_wanted(2, 128, 0, 16);  // reads the environment once
  (void)bn;  // 64-wide pair tiles pay as well (conv1_2: 1.02 -> 0.85 ms)
  return g_pair_mode != 0;
}

int plan_gemm(GemmPlan* p, const Operand& A, const Operand* A2, long long a2_at, const Operand& B, int M, int N, int K,
              int bn, int splits, bool pair_ok) {
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
  g.stages = gemm_pick_stages(bn, 0);
  g.a_mode = A.mn_major ? A_MNMAJOR : A_KMAJOR;
  g.b_mn = B.mn_major ? 1 : 0;
  g.a_switch = -1;
  g.pair = pair_ok && gemm_pair_wanted(g.m_tiles, bn, g.b_mn, (g.k_blocks + g.splits - 1) / g.splits) ? 1 : 0;
  if (g.pair) g.m_tiles = (g.m_tiles + 1) / 2 * 2;
  g.dual = pair_ok && gemm_dual_wanted(g.n_tiles, bn, (g.k_blocks + g.splits - 1) / g.splits,
                                       g.m_tiles * g.n_tiles * g.splits / (g.pair ? 2 : 1),
                                       g.pair ? num_sms() / 2 : num_sms()) ? 1 : 0;
  VC_TRY(operand_tmap(&p->tmA, A, kBM));
  p->tmA2 = p->tmA;
  if (A2 != nullptr && A2->ptr != nullptr) {
    if (A2->mn_major != A.mn_major) return set_error(VC_E_ARG, "plan_gemm: A and A2 must share major-ness");
    if (a2_at % (A.mn_major ? kBM : kBK) != 0)
      return set_error(VC_E_ARG, "plan_gemm: A2 switch point %lld not tile aligned", a2_at);
    g.a_switch = (int)(a2_at / (A.mn_major ? kBM : kBK));
    VC_TRY(operand_tmap(&p->tmA2, *A2, kBM));
  }
  VC_TRY(operand_tmap(&p->tmB, B, g.pair ? bn / 2 : bn));
  return VC_OK;
}

int conv_geometry(ConvGeom* g, int W, int H, int Nimg, int Cin, int Cout) {
  g->W = W; g->H = H; g->Nimg = Nimg; g->Cin = Cin; g->Cout = Cout;
  if (W % 16 == 0 && H % 8 == 0) { g->pw = 16; g->ph = 2; g->pn = 1; g->tw = 1; g->th = 4; }       // 16 x 8 x 1 tiles
  else if (W % 8 == 0 && H % 8 == 0) { g->pw = 8; g->ph = 4; g->pn = 1; g->tw = 1; g->th = 2; }    // 8 x 8 x 2
  else if (W % 4 == 0 && H % 4 == 0) { g->pw = 4; g->ph = 4; g->pn = 2; g->tw = 1; g->th = 1; }    // 4 x 4 x 8
  else if (W % 2 == 0 && H % 2 == 0) { g->pw = 2; g->ph = 2; g->pn = 8; g->tw = 1; g->th = 1; }    // 2 x 2 x 32
  else return set_error(VC_E_SHAPE, "conv_geometry: %dx%d feature map has odd extent", W, H);
  return VC_OK;
}

bool conv_halo_applicable(int W, int H, int Cin, int Cout) {
  return Cin == 64 && Cout % 64 == 0 && W % 8 == 0 && H % 16 == 0;
}

int conv_halo_geometry(ConvGeom* g, int W, int H, int Nimg, int Cin, int Cout) {
  g->W = W; g->H = H; g->Nimg = Nimg; g->Cin = Cin; g->Cout = Cout;
  g->pw = 8; g->ph = 4; g->pn = 1; g->tw = 1; g->th = 4;  // 8 x 16 x 1 tiles, one 8 x 4 patch per epilogue warp
  return VC_OK;
}

int plan_conv_halo(GemmPlan* p, const void* in, const void* wt, const ConvGeom& cg) {
  if (cg.Cin != 64 || cg.Cout % 64 != 0 || cg.W % 8 != 0 || cg.H % 16 != 0)
    return set_error(VC_E_SHAPE, "plan_conv_halo: unsupported layer %dx%d Cin=%d Cout=%d", cg.W, cg.H, cg.Cin, cg.Cout);
  memset(p, 0, sizeof(*p));
  GemmCore& g = p->core;
  g.pw = cg.pw; g.ph = cg.ph; g.pn = cg.pn; g.tw = cg.tw; g.th = cg.th;
  g.tiles_w = cg.W / 8;
  g.tiles_h = cg.H / 16;
  g.m_tiles = g.tiles_w * g.tiles_h * cg.Nimg;
  g.pair = halo_pair_wanted(cg.Cout % 128 == 0 ? 128 : 64) && g.m_tiles >= 2 ? 1 : 0;
  g.bn = g.pair && cg.Cout % 128 == 0 ? 128 : 64;  // a pair covers 128 output channels with the filter bytes of 64 per CTA
  g.n_tiles = cg.Cout / g.bn;
  g.cpk = 1;
  g.k_blocks = 9;
  g.splits = 1;
  g.stages = kHaloStages;
  g.a_mode = A_CONV3x3;
  g.a_switch = -1;
  VC_TRY(make_tmap_nhwc(&p->tmA, in, cg.Cin, cg.W, cg.H, cg.Nimg, kHaloLineRows, kHaloLines, 1));
  p->tmA2 = p->tmA;
  VC_TRY(make_tmap_2d(&p->tmB, wt, 9ull * cg.Cin, cg.Cout, 9ull * cg.Cin, 64, g.pair ? g.bn / 2 : 64));
  return VC_OK;
}

static int wg_max_ch() {  // VC_WGRAD_HALO_MAXCH: widest layer that takes the halo form (experiment knob)
  static const int v = [] {
    const char* e = getenv("VC_WGRAD_HALO_MAXCH");
    return e ? atoi(e) : 256;
  }();
  return v;
}

bool conv_wgrad_halo_applicable(int W, int H, int Cin, int Cout) {
  static const bool enabled = [] {
    const char* e = getenv("VC_WGRAD_HALO");
    return !(e && e[0] == '0');
  }();
  // maps of 56 x 56 and wider (conv1_2 .. conv3_3): there the generic pixel-contraction gather is L2-operand-bound
  // (measured: wgrad1_2 4.39 -> 0.98 ms, wgrad2_2 2.07 -> 0.95, wgrad3_2 1.19 -> 0.96); 28 x 28 and 14 x 14 maps do not
  // tile into 8 x 8 patches without waste and keep the generic path
  return enabled && Cin % 64 == 0 && Cout % 64 == 0 && Cin <= wg_max_ch() && Cout <= wg_max_ch() && W >= 56 && H >= 56;
}

template <bool kWide>
static int launch_conv_wgrad_halo_t(cudaStream_t stream, const void* x, const void* dy, float* dw, int W, int H, int Nimg,
                                    int Cin, int Cout) {
  static bool configured = false;
  const int smem = wg_stages(kWide) * wg_stage_bytes(kWide) + 1024 + 256;
  if (!configured) {
    VC_CUDA(cudaFuncSetAttribute(conv_wgrad_halo_kernel<kWide>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  WgradHaloArgs a{};
  a.dw = dw;
  a.Cin = Cin;
  a.Cout = Cout;
  a.tiles_w = (W + 7) / 8;
  a.tiles_h = (H + 7) / 8;
  a.n_img = Nimg;
  a.k_total = a.tiles_w * a.tiles_h * Nimg;
  // wide: (64 input channels) x (128 output channels) x (tap half); narrow: (64) x (64), all nine taps
  const int units = kWide ? (Cin / 64) * (Cout / 128) * 2 : (Cin / 64) * (Cout / 64);
  a.splits = std::max(1, std::min(num_sms() / units, a.k_total));
  CUtensorMap tmX, tmDy;
  VC_TRY(make_tmap_nhwc(&tmX, x, Cin, W, H, Nimg, kHaloLineRows, 10, 1));
  VC_TRY(make_tmap_nhwc(&tmDy, dy, Cout, W, H, Nimg, 8, 8, 1));
  {
    ProfScope ps(stream, "conv_wgrad_halo");
    conv_wgrad_halo_kernel<kWide><<<units * a.splits, 256, smem, stream>>>(tmX, tmDy, a);
  }
  VC_CUDA(cudaGetLastError());
  return VC_OK;
}

int launch_conv_wgrad_halo(cudaStream_t stream, const void* x, const void* dy, float* dw, int W, int H, int Nimg, int Cin,
                           int Cout) {
  static const bool wide_ok = [] {
    const char* e = getenv("VC_WGRAD_WIDE");
    return !(e && e[0] == '0');
  }();
  if (wide_ok && Cout % 128 == 0) return launch_conv_wgrad_halo_t<true>(stream, x, dy, dw, W, H, Nimg, Cin, Cout);
  return launch_conv_wgrad_halo_t<false>(stream, x, dy, dw, W, H, Nimg, Cin, Cout);
}

bool conv_halo_stream_applicable(int W, int H, int Cin, int Cout) {
  static const bool enabled = [] {
    const char* e = getenv("VC_CONV_HALO2");
    return !(e && e[0] == '0');
  }();
  // Cout = 256 (conv3_x, 56 x 56): the generic path moves 1152 KB of operands per 128 x 256 tile (125 B/clk per SM at
  // tensor speed -- it runs at 73 % tensor-pipe utilisation, i.e. at the ~90 B/clk an SM ingests), this one 720 KB; the
  // price is the ragged last tile row (56 = 3.5 x 16 lines: 12.5 % of the MMAs multiply zero fill)
  static const bool wide = [] {
    const char* e = getenv("VC_CONV_HALO2_WIDE");
    return !(e && e[0] == '0');
  }();
  return enabled && (Cin == 128 || Cin == 256) && (Cout == 64 || Cout == 128 || (wide && Cout == 256)) && W % 8 == 0 &&
         H % 2 == 0 && H >= 16;
}

int plan_conv_halo_stream(GemmPlan* p, const void* in, const void* wt, const ConvGeom& cg) {
  if (!(cg.Cin == 128 || cg.Cin == 256) || !(cg.Cout == 64 || cg.Cout == 128 || cg.Cout == 256) || cg.W % 8 != 0)
    return set_error(VC_E_SHAPE, "plan_conv_halo_stream: unsupported layer %dx%d Cin=%d Cout=%d", cg.W, cg.H, cg.Cin, cg.Cout);
  memset(p, 0, sizeof(*p));
  GemmCore& g = p->core;
  g.pw = cg.pw; g.ph = cg.ph; g.pn = cg.pn; g.tw = cg.tw; g.th = cg.th;
  g.tiles_w = cg.W / 8;
  g.tiles_h = (cg.H + 15) / 16;  // a ragged last tile row: TMA zero-fills the loads and clips the stores
  g.m_tiles = g.tiles_w * g.tiles_h * cg.Nimg;
  g.n_tiles = 1;
  g.cpk = cg.Cin / 64;
  g.k_blocks = 9 * g.cpk;
  g.splits = 1;
  g.bn = cg.Cout;
  g.pair = halo_pair_wanted(g.bn) && g.m_tiles >= 2 ? 1 : 0;
  g.stages = kHaloStages;
  g.a_mode = A_CONV3x3;
  g.a_switch = -1;
  VC_TRY(make_tmap_nhwc(&p->tmA, in, cg.Cin, cg.W, cg.H, cg.Nimg, kHaloLineRows, kHaloLines, 1));
  p->tmA2 = p->tmA;
  VC_TRY(make_tmap_2d(&p->tmB, wt, 9ull * cg.Cin, cg.Cout, 9ull * cg.Cin, 64, g.pair ? cg.Cout / 2 : cg.Cout));
  return VC_OK;
}

int plan_conv(GemmPlan* p, const void* in, const void* wt, const ConvGeom& cg, int bn) {
  if (cg.Cin % 64 != 0) return set_error(VC_E_SHAPE, "plan_conv: Cin=%d must be a multiple of 64", cg.Cin);
  if (bn % 64 != 0 || bn > 256) return set_error(VC_E_ARG, "plan_conv: bn=%d", bn);
  memset(p, 0, sizeof(*p));
  GemmCore& g = p->core;
  const int tn = 4 / (cg.tw * cg.th);
  g.pw = cg.pw; g.ph = cg.ph; g.pn = cg.pn; g.tw = cg.tw; g.th = cg.th;
  g.tiles_w = (cg.W + cg.pw * cg.tw - 1) / (cg.pw * cg.tw);
  g.tiles_h = (cg.H + cg.ph * cg.th - 1) / (cg.ph * cg.th);
  const int tiles_n = (cg.Nimg + cg.pn * tn - 1) / (cg.pn * tn);
  g.m_tiles = g.tiles_w * g.tiles_h * tiles_n;
  g.n_tiles = (cg.Cout + bn - 1) / bn;
  g.cpk = cg.Cin / 64;
  g.k_blocks = 9 * g.cpk;
  g.splits = 1;
  g.bn = bn;
  g.stages = gemm_pick_stages(bn, 0);
  g.a_mode = A_CONV3x3;
  g.b_mn = 0;
  g.a_switch = -1;
  g.pair = gemm_pair_wanted(g.m_tiles, bn, 0, g.k_blocks) ? 1 : 0;
  if (g.pair) g.m_tiles = (g.m_tiles + 1) / 2 * 2;  // a surplus tile lies past the last image: zero fill in, clipped out
  g.dual = gemm_dual_wanted(g.n_tiles, bn, g.k_blocks, g.m_tiles * g.n_tiles / (g.pair ? 2 : 1),
                            g.pair ? num_sms() / 2 : num_sms()) ? 1 : 0;
  VC_TRY(make_tmap_nhwc(&p->tmA, in, cg.Cin, cg.W, cg.H, cg.Nimg, cg.pw, cg.ph, cg.pn));
  p->tmA2 = p->tmA;
  VC_TRY(make_tmap_2d(&p->tmB, wt, 9ull * cg.Cin, cg.Cout, 9ull * cg.Cin, 64, g.pair ? bn / 2 : bn));
  return VC_OK;
}

int plan_conv1_window(GemmPlan* p, const void* padded, const void* wt, int W, int H, int Nimg, int Cout) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return set_error(VC_E_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  ConvGeom cg;
  VC_TRY(conv_geometry(&cg, W, H, Nimg, 64, Cout));
  memset(p, 0, sizeof(*p));
  GemmCore& g = p->core;
  const int tn = 4 / (cg.tw * cg.th);
  g.pw = cg.pw; g.ph = cg.ph; g.pn = cg.pn; g.tw = cg.tw; g.th = cg.th;
  g.tiles_w = (W + cg.pw * cg.tw - 1) / (cg.pw * cg.tw);
  g.tiles_h = (H + cg.ph * cg.th - 1) / (cg.ph * cg.th);
  g.m_tiles = g.tiles_w * g.tiles_h * ((Nimg + cg.pn * tn - 1) / (cg.pn * tn));
  g.n_tiles = 1;
  g.cpk = 1;
  g.k_blocks = 3;
  g.splits = 1;
  g.bn = Cout;
  g.stages = gemm_pick_stages(Cout, 0);
  g.a_mode = A_CONV3x3;
  g.b_mn = 0;
  g.a_switch = -1;
  g.tap_rows = 1;
  cuuint64_t dims[4] = {64, (cuuint64_t)W, (cuuint64_t)H + 2, (cuuint64_t)Nimg};
  cuuint64_t strides[3] = {16, (cuuint64_t)(W + 2) * 16, (cuuint64_t)(H + 2) * (W + 2) * 16};
  cuuint32_t box[4] = {64, (cuuint32_t)cg.pw, (cuuint32_t)cg.ph, (cuuint32_t)cg.pn};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(&p->tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(padded), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(VC_E_CUDA, "cuTensorMapEncodeTiled(conv1 window map W=%d H=%d N=%d) -> %d", W, H, Nimg, (int)r);
  p->tmA2 = p->tmA;
  VC_TRY(make_tmap_2d(&p->tmB, wt, 192, Cout, 192, 64, Cout));
  return VC_OK;
}

int plan_conv_wgrad(GemmPlan* p, const void* in, const void* dy, int W, int H, int Nimg, int Cin, int Cout, int bn,
                    int splits) {
  if (Cin % 64 != 0 || Cout % 64 != 0) return set_error(VC_E_SHAPE, "plan_conv_wgrad: Cin=%d Cout=%d must be multiples of 64", Cin, Cout);
  if (bn % 64 != 0 || bn > 256 || Cout % bn != 0) return set_error(VC_E_ARG, "plan_conv_wgrad: bn=%d", bn);
  memset(p, 0, sizeof(*p));
  GemmCore& g = p->core;
  // 64-pixel contraction patches that tile the feature map exactly
  if (W % 16 == 0 && H % 4 == 0) { g.pw = 16; g.ph = 4; g.pn = 1; }
  else if (W % 8 == 0 && H % 8 == 0) { g.pw = 8; g.ph = 8; g.pn = 1; }
  else if (W % 4 == 0 && H % 4 == 0) { g.pw = 4; g.ph = 4; g.pn = 4; }
  else if (W % 2 == 0 && H % 2 == 0) { g.pw = 2; g.ph = 2; g.pn = 16; }
  else return set_error(VC_E_SHAPE, "plan_conv_wgrad: %dx%d feature map has odd extent", W, H);
  g.tw = g.th = 1;
  g.tiles_w = W / g.pw;
  g.tiles_h = H / g.ph;
  g.k_blocks = g.tiles_w * g.tiles_h * ((Nimg + g.pn - 1) / g.pn);
  g.cpk = Cin / 64;
  g.m_tiles = (9 * Cin + kBM - 1) / kBM;
  g.n_tiles = Cout / bn;
  g.splits = splits < 1 ? 1 : (splits > g.k_blocks ? g.k_blocks : splits);
  g.bn = bn;
  g.stages = gemm_pick_stages(bn, 0);
  g.a_mode = A_WGRAD3x3;
  g.b_mn = 2;
  g.a_switch = -1;
  g.n_img = Nimg;
  // pairs: each CTA gathers its own (tap, channel chunk) rows and half of the dy patch columns
  g.pair = gemm_pair_wanted(g.m_tiles, bn, 1, (g.k_blocks + g.splits - 1) / g.splits) ? 1 : 0;
  if (g.pair) g.m_tiles = (g.m_tiles + 1) / 2 * 2;
  // dual tiles: the gathered x patches (the expensive operand) are fetched once for two 256-column blocks of dy
  g.dual = Cout % (2 * bn) == 0 && gemm_dual_wanted(g.n_tiles, bn, (g.k_blocks + g.splits - 1) / g.splits,
                                                     g.m_tiles * g.n_tiles * g.splits / (g.pair ? 2 : 1),
                                                     g.pair ? num_sms() / 2 : num_sms()) ? 1 : 0;
  VC_TRY(make_tmap_nhwc(&p->tmA, in, Cin, W, H, Nimg, g.pw, g.ph, g.pn));
  p->tmA2 = p->tmA;
  VC_TRY(make_tmap_nhwc(&p->tmB, dy, Cout, W, H, Nimg, g.pw, g.ph, g.pn));
  return VC_OK;
}

}  // namespace vc
