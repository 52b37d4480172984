// VGG16 fine-tune backward (reference: ops/optimizers.py:49-82 differentiates the loss w.r.t. the 30 cnn/ variables
// of utils/image_embeddings.py:26-238 when --fine_tune is set; main.py:65-79 feeds images instead of features).
//
// Every contraction runs on the tcgen05 mainloop:
//   fc dgrad   : dX[B, in]   = dPre[B, out] x W[in, out]^T            (W's natural layout is the K-major B operand)
//   fc wgrad   : dW[in, out] = X[B, in]^T x dPre[B, out]              (both operands MN-major, contraction over B)
//   conv dgrad : dA_{l-1}    = conv3x3_same(dY_l, rot180(W_l)^T)      (the forward implicit-GEMM kernel on a
//                                                                      tap-reversed weight shadow [Cin, 9*Cout])
//   conv wgrad : dW_l[9*Cin, Cout] = patches(A_{l-1})^T x dY_l        (A_WGRAD3x3 loader: contraction over pixels,
//                                                                      tap-shifted TMA gathers, split-K fp32 atomics)
// The ReLU / 2x2 max-pool / dropout derivatives are HBM-bound elementwise kernels between them. The L2 term
// weight_decay * w (Q11) is folded into the CNN Adam update (Model::apply) and into vc_grad_get.
#include <utility>
#include "model.h"

namespace vc {

const char* vgg_layer_name(int l);

static inline int ew_grid(long long n, int block) {
  long long g = (n + block - 1) / block;
  const long long cap = (long long)num_sms() * 16;
  return (int)(g > cap ? cap : (g < 1 ? 1 : g));
}

// Bias gradient fused into the derivative kernels: C/8 divides the grid stride, so a thread sees the same 8 channels
// in every iteration and keeps their sums in registers; one shared-memory and one global atomic per channel per CTA.
__device__ __forceinline__ void bias_grad_flush(const float (&acc)[8], int c8, int C, float* __restrict__ db) {
  extern __shared__ float s_db[];
  for (int c = threadIdx.x; c < C; c += blockDim.x) s_db[c] = 0.f;
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 8; ++j) atomicAdd(&s_db[c8 * 8 + j], acc[j]);
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) atomicAdd(db + c, s_db[c]);
}

// dY = dA * (out > 0), 8 channels per thread (ReLU derivative; un-pooled layers); db[c] += sum of dY over pixels
__global__ void k_relu_bwd(const uint4* __restrict__ dA, const uint4* __restrict__ out, uint4* __restrict__ dY, long long n8,
                           int C, float* __restrict__ db) {
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const int c8 = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) % (C / 8));
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    uint4 g = dA[i];
    const uint4 o = out[i];
    const __nv_bfloat162* ov = reinterpret_cast<const __nv_bfloat162*>(&o);
    __nv_bfloat162* gv = reinterpret_cast<__nv_bfloat162*>(&g);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 of = __bfloat1622float2(ov[j]);
      float2 gf = __bfloat1622float2(gv[j]);
      gf.x = of.x > 0.f ? gf.x : 0.f;
      gf.y = of.y > 0.f ? gf.y : 0.f;
      gv[j] = __floats2bfloat162_rn(gf.x, gf.y);
      acc[2 * j] += gf.x;
      acc[2 * j + 1] += gf.y;
    }
    dY[i] = g;
  }
  bias_grad_flush(acc, c8, C, db);
}

// 2x2/2 max-pool + ReLU derivative: dA is the gradient of the pooled map [B, H/2, W/2, C]; dY (un-pooled,
// [B, H, W, C]) receives it at the window's first maximum (scan order (0,0),(0,1),(1,0),(1,1), as TF's and
// torch's MaxPoolGrad do) if that maximum is positive, zero elsewhere. One thread = one window x 8 channels.
__global__ void k_pool_relu_bwd(const __nv_bfloat16* __restrict__ dA, const __nv_bfloat16* __restrict__ out,
                                __nv_bfloat16* __restrict__ dY, int B, int H, int W, int C, float* __restrict__ db) {
  const int Ho = H / 2, Wo = W / 2, C8 = C / 8;
  const long long total = (long long)B * Ho * Wo * C8;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const int c8_fix = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) % C8);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % C8);
    long long r = i / C8;
    const int wo = (int)(r % Wo);
    r /= Wo;
    const int ho = (int)(r % Ho);
    const long long b = r / Ho;
    const long long p00 = (((b * H + 2 * ho) * W + 2 * wo) * C) + c8 * 8;
    const long long offs[4] = {p00, p00 + C, p00 + (long long)W * C, p00 + (long long)W * C + C};
    uint4 q[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) q[k] = *reinterpret_cast<const uint4*>(out + offs[k]);
    const uint4 g = *reinterpret_cast<const uint4*>(dA + (((b * Ho + ho) * Wo + wo) * C) + c8 * 8);
    const __nv_bfloat16* gv = reinterpret_cast<const __nv_bfloat16*>(&g);
    uint4 o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) o[k] = make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float best = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(&q[0])[j]);
      int arg = 0;
#pragma unroll
      for (int k = 1; k < 4; ++k) {
        const float v = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(&q[k])[j]);
        if (v > best) { best = v; arg = k; }
      }
      if (best > 0.f) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (k == arg) reinterpret_cast<__nv_bfloat16*>(&o[k])[j] = gv[j];
        acc[j] += __bfloat162float(gv[j]);
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) *reinterpret_cast<uint4*>(dY + offs[k]) = o[k];
  }
  bias_grad_flush(acc, c8_fix, C, db);
}

// fc layers: dPre = dOut * scale where the (post-ReLU, post-dropout) activation y is positive. y > 0 <=> the unit
// passed both the ReLU and the dropout mask, so the keep mask itself is not needed again; scale = 1 / keep_prob.
template <class TY>
__global__ void k_fc_mask_bwd(const float* __restrict__ dout, const TY* __restrict__ y, float scale,
                              __nv_bfloat16* __restrict__ dpre, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float yv = static_cast<float>(y[i]);
    dpre[i] = __float2bfloat16(yv > 0.f ? dout[i] * scale : 0.f);
  }
}

// dst[ci, (8 - tap) * Cout + co] = bf16(W[tap, ci, co]): the dgrad filter (rotated 180 degrees, channels swapped)
// laid out as the K-major B operand [Cin, 9 * Cout] of the forward implicit-GEMM kernel.
__global__ void k_dgrad_shadow(const float* __restrict__ w, __nv_bfloat16* __restrict__ dst, int Cin, int Cout) {
  const long long total = 9LL * Cin * Cout;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int co = (int)(i % Cout);
    const long long r = i / Cout;
    const int ci = (int)(r % Cin);
    const int tap = (int)(r / Cin);
    dst[(long long)ci * 9 * Cout + (8 - tap) * Cout + co] = __float2bfloat16(w[i]);
  }
}

// dW[9*Cin, Cout] (fp32, HWIO row-major, accumulated with atomics: the caller zeroes it) += patches(x)^T x dY
int conv3x3_wgrad(cudaStream_t s, const void* x, const void* dy, float* dw, int B, int hw, int cin, int cout, const char* tag) {
  ProfTag pt(tag);
  if (conv_wgrad_halo_applicable(hw, hw, cin, cout)) return launch_conv_wgrad_halo(s, x, dy, dw, hw, hw, B, cin, cout);
  const int bn = cout >= 256 ? 256 : cout;
  GemmPlan plan;
  const int tiles = ((9 * cin + 127) / 128) * (cout / bn);
  int splits = std::max(1, num_sms() / tiles);
  // dual-N tiles halve the number of schedule units: split the pixel contraction twice as often to keep the SMs busy
  // (the planner's wave rule then accepts the dual form: 36 m-tiles x 2 n-tiles x 4 splits = 72 dual pair tiles on 74 clusters)
  if (bn == 256 && cout % 512 == 0) {
    const long long kb = (long long)B * hw * hw / 64;
    if (gemm_dual_wanted(cout / bn, bn, (int)(kb / (2 * splits)), tiles * 2 * splits / 2, num_sms() / 2)) splits *= 2;
  }
  VC_TRY(plan_conv_wgrad(&plan, x, dy, hw, hw, B, cin, cout, bn, splits));
  EpiStore e{};
  e.out = dw; e.ld = cout; e.alpha = 1.f; e.atomic = 1;
  e.M = 9 * cin; e.N = cout; e.bn = bn;
  return launch_gemm(plan, e, s);
}

// dx [B, hw, hw, Cin] (bf16 NHWC) = conv3x3_same(dy, tap-reversed W^T); wt_d is the [Cin, 9*Cout] shadow
template <class Epi>
static int conv3x3_dgrad_t(cudaStream_t s, const void* dy, const void* wt_d, void* dx, int B, int hw, int cin, int cout,
                           Epi& epi) {
  const int bnd = cin >= 256 ? 256 : cin;
  epi.bias = nullptr; epi.N = cin; epi.bn = bnd; epi.relu = 0; epi.alpha = 1.f; epi.mode = kConv;
  GemmPlan plan;
  ConvGeom g;
  const bool halo = conv_halo_applicable(hw, hw, cout, cin);
  const bool halo2 = !halo && conv_halo_stream_applicable(hw, hw, cout, cin);
  if (halo) {
    VC_TRY(conv_halo_geometry(&g, hw, hw, B, cout, cin));
    VC_TRY(plan_conv_halo(&plan, dy, wt_d, g));
    epi.bn = plan.core.bn;
  } else if (halo2) {
    VC_TRY(conv_halo_geometry(&g, hw, hw, B, cout, cin));
    VC_TRY(plan_conv_halo_stream(&plan, dy, wt_d, g));
    epi.bn = cin;
  } else {
    VC_TRY(conv_geometry(&g, hw, hw, B, cout, cin));
    VC_TRY(plan_conv(&plan, dy, wt_d, g, bnd));
  }
  VC_TRY(make_tmap_nhwc(&epi.tm, dx, cin, hw, hw, B, g.pw * g.tw, g.ph * g.th, g.pn * (4 / (g.tw * g.th))));
  if (halo) return launch_conv_halo(plan, epi, s);
  if (halo2) return launch_conv_halo_stream(plan, epi, s);
  return launch_gemm(plan, epi, s);
}

int conv3x3_dgrad(cudaStream_t s, const void* dy, const void* wt_d, void* dx, int B, int hw, int cin, int cout, const char* tag,
                  const void* relu_src, float* dbias) {
  ProfTag pt(tag);
  if (relu_src != nullptr) {
    if (cin % 64 != 0 || cin > 512 || dbias == nullptr)
      return set_error(VC_E_SHAPE, "conv3x3_dgrad: fused ReLU gradient needs 64 | Cin <= 512 and a bias gradient (Cin=%d)", cin);
    EpiTmaRelu epi{};
    epi.relu_src = (const __nv_bfloat16*)relu_src;
    epi.dbias = dbias;
    epi.W = hw; epi.H = hw; epi.n_img = B;
    return conv3x3_dgrad_t(s, dy, wt_d, dx, B, hw, cin, cout, epi);
  }
  EpiTma epi{};
  return conv3x3_dgrad_t(s, dy, wt_d, dx, B, hw, cin, cout, epi);
}

int dgrad_shadow(cudaStream_t s, const float* w_hwio, void* wt_d, int cin, int cout) {
  {
    ProfScope ps(s, "refresh_shadows");
    k_dgrad_shadow<<<ew_grid(9LL * cin * cout, 256), 256, 0, s>>>(w_hwio, (__nv_bfloat16*)wt_d, cin, cout);
  }
  VC_CUDA(cudaGetLastError());
  return VC_OK;
}

// ReLU (+ 2x2 max-pool when pooled) derivative: dA -> dY (un-pooled), see the kernels above
int relu_pool_bwd(cudaStream_t s, const void* dA, const void* out, void* dY, int B, int hw, int C, bool pooled, float* db) {
  const long long pix = (long long)B * hw * hw;
  if (C % 8 != 0 || 256 % (C / 8) != 0) return set_error(VC_E_SHAPE, "relu_pool_bwd: C=%d must be 8 x a divisor of 256", C);
  const size_t smem = (size_t)C * sizeof(float);
  if (pooled) {
    ProfScope ps(s, "pool_relu_bwd");
    k_pool_relu_bwd<<<ew_grid(pix / 4 * (C / 8), 256), 256, smem, s>>>((const __nv_bfloat16*)dA, (const __nv_bfloat16*)out,
                                                                      (__nv_bfloat16*)dY, B, hw, hw, C, db);
  } else {
    ProfScope ps(s, "relu_bwd");
    k_relu_bwd<<<ew_grid(pix * (C / 8), 256), 256, smem, s>>>((const uint4*)dA, (const uint4*)out, (uint4*)dY, pix * (C / 8), C, db);
  }
  VC_CUDA(cudaGetLastError());
  return VC_OK;
}

// dW[k, co] = D[co, k] + D[64 + co, 32 + k] for the two-pixel-row conv1_1 filter gradient (D is 128 x 64 fp32)
__global__ void k_conv1_fold(const float* __restrict__ D, float* __restrict__ dw) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 27 * 64) {
    const int k = i / 64, co = i % 64;
    dw[i] = D[co * 64 + k] + D[(64 + co) * 64 + 32 + k];
  }
}

int Model::vgg_bwd_init() {
  const int B = cfg.max_batch;
  for (int l = 1; l < 13; ++l) VC_TRY(dalloc((uint16_t**)&vgg[l].wt_d, (size_t)9 * vgg[l].cin * vgg[l].cout));
  const size_t big = (size_t)B * 224 * 224 * 64;
  VC_TRY(dalloc((uint16_t**)&vgg_bwd_a, big));
  VC_TRY(dalloc((uint16_t**)&vgg_bwd_b, big));
  VC_TRY(dalloc((uint16_t**)&dfc2_pre, (size_t)B * 4096));
  VC_TRY(dalloc((uint16_t**)&dfc1_pre, (size_t)B * 4096));
  VC_TRY(dalloc(&dfeats_f, (size_t)B * cfg.cnn_feature_size));
  VC_TRY(dalloc(&conv1_wg, 128 * 64));
  return VC_OK;
}

int Model::vgg_refresh_bwd_shadows(cudaStream_t s) {
  for (int l = 1; l < 13; ++l) VC_TRY(dgrad_shadow(s, pp(vgg[l].p_w), vgg[l].wt_d, vgg[l].cin, vgg[l].cout));
  return VC_OK;
}

// dfeats: device fp32 [B, 4096] = d lower_bound / d fc2 (post-dropout). Fills the cnn/ region of the flat gradient
// buffer (weights_regularizer term excluded, see the header comment). Needs the un-pooled activations of the last
// vgg_forward (keep_unpooled = true).
int Model::vgg_backward(const float* dfeats, int B, cudaStream_t s) {
  if (!cfg.fine_tune) return set_error(VC_E_STATE, "vgg_backward needs a fine_tune handle");
  if (!vgg_have_unpooled || vgg_last_B != B) return set_error(VC_E_STATE, "vgg_backward: no matching forward pass");
  const float inv_keep = 1.f / cfg.cnn_dropout;
  const int p_fc1w = pidx("cnn/fc1/weights"), p_fc1b = pidx("cnn/fc1/biases"), p_fc2w = pidx("cnn/fc2/weights"),
            p_fc2b = pidx("cnn/fc2/biases");
  // conv gradients and all biases accumulate with atomics: zero them (the two fc weight gradients are plain stores)
  const int64_t cnn0 = params[vgg[0].p_w].offset;
  VC_CUDA(cudaMemsetAsync(Gf + cnn0, 0, (size_t)(params[p_fc1w].offset - cnn0) * sizeof(float), s));
  VC_CUDA(cudaMemsetAsync(gp(p_fc1b), 0, 4096 * sizeof(float), s));
  VC_CUDA(cudaMemsetAsync(gp(p_fc2b), 0, 4096 * sizeof(float), s));

  // ---- fc2 (image_embeddings.py:228-238)
  {
    ProfScope ps(s, "fc_mask_bwd");
    k_fc_mask_bwd<float><<<ew_grid((long long)B * 4096, 256), 256, 0, s>>>(dfeats, fc2_f, inv_keep, (__nv_bfloat16*)dfc2_pre,
                                                                          (long long)B * 4096);
  }
  {
    ProfTag pt("fc_wgrad");
    Operand A{fc1_h, B, 4096, 4096, true}, Bm{dfc2_pre, B, 4096, 4096, true};
    EpiStore e{};
    e.out = gp(p_fc2w); e.ld = 4096; e.alpha = 1.f;
    VC_TRY(gemm_store(s, A, nullptr, 0, Bm, 4096, 4096, B, e, 256, 1));
  }
  VC_TRY(colsum_bf16(s, dfc2_pre, B, 4096, 4096, gp(p_fc2b)));
  VC_TRY(grad_ready_params({p_fc2w, p_fc2b}, s));  // data parallel: the cnn/ gradients travel layer group by layer group
  {
    ProfTag pt("fc_dgrad");
    Operand A{dfc2_pre, B, 4096, 4096, false}, Bm{fc2_w, 4096, 4096, 4096, false};
    EpiStore e{};
    e.out = fc_acc; e.ld = 4096; e.alpha = 1.f;
    VC_TRY(gemm_store(s, A, nullptr, 0, Bm, B, 4096, 4096, e, 128, 1));
  }
  // ---- fc1 (image_embeddings.py:214-226)
  {
    ProfScope ps(s, "fc_mask_bwd");
    k_fc_mask_bwd<__nv_bfloat16><<<ew_grid((long long)B * 4096, 256), 256, 0, s>>>(
        fc_acc, (const __nv_bfloat16*)fc1_h, inv_keep, (__nv_bfloat16*)dfc1_pre, (long long)B * 4096);
  }
  {
    ProfTag pt("fc_wgrad");
    Operand A{vgg[12].pooled, B, 25088, 25088, true}, Bm{dfc1_pre, B, 4096, 4096, true};
    EpiStore e{};
    e.out = gp(p_fc1w); e.ld = 4096; e.alpha = 1.f;
    VC_TRY(gemm_store(s, A, nullptr, 0, Bm, 25088, 4096, B, e, 256, 1));
  }
  VC_TRY(colsum_bf16(s, dfc1_pre, B, 4096, 4096, gp(p_fc1b)));
  VC_TRY(grad_ready_params({p_fc1w, p_fc1b}, s));  // 411 MB, under the whole convolutional backward pass
  uint16_t* dA = (uint16_t*)vgg_bwd_a;  // gradient w.r.t. the (pooled) output of the current layer
  uint16_t* dY = (uint16_t*)vgg_bwd_b;  // gradient w.r.t. its pre-activation, un-pooled
  {
    ProfTag pt("fc_dgrad");
    Operand A{dfc1_pre, B, 4096, 4096, false}, Bm{fc1_w, 25088, 4096, 4096, false};
    EpiStore e{};
    e.out = dA; e.ld = 25088; e.alpha = 1.f; e.out_bf16 = 1;
    VC_TRY(gemm_store(s, A, nullptr, 0, Bm, B, 25088, 4096, e, 256, 1));
  }
  // ---- conv5_3 ... conv1_1
  static const bool fuse_relu = [] {
    const char* e = getenv("VC_FUSE_RELU_BWD");
    return !(e && e[0] == '0');
  }();
  bool masked = false;
  for (int l = 12; l >= 0; --l) {
    VggLayer& L = vgg[l];
    const long long pix = (long long)B * L.hw * L.hw;
    if (masked)
      std::swap(dA, dY);  // the dgrad epilogue of layer l + 1 has applied this layer's ReLU derivative and summed its bias gradient
    else
      VC_TRY(relu_pool_bwd(s, dA, L.out, dY, B, L.hw, L.cout, L.pool, gp(L.p_b)));  // + bias gradient
    if (l == 0) {
      // conv1_1: dW[27, 64] = im2col[pixels, 27]^T x dY[pixels, 64]. Both operands are read as two-pixel rows
      // ([pixels/2, 128] and [pixels/2, 64]: full 128-byte TMA rows, see k_conv1_shadow), which yields the 128 x 64
      // matrix D[par*64 + co, par'*32 + k]; the filter gradient is the sum of its two parity-diagonal blocks.
      ProfTag pt("wgrad1_1");
      VC_CUDA(cudaMemsetAsync(conv1_wg, 0, 128 * 64 * sizeof(float), s));
      Operand A{dY, pix / 2, 128, 128, true}, Bm{vgg_im2col, pix / 2, 64, 64, true};
      EpiStore e{};
      e.out = conv1_wg; e.ld = 64; e.alpha = 1.f; e.atomic = 1;
      VC_TRY(gemm_store(s, A, nullptr, 0, Bm, 128, 64, (int)(pix / 2), e, 64, num_sms()));
      {
        ProfScope ps(s, "conv1_fold");
        k_conv1_fold<<<7, 256, 0, s>>>(conv1_wg, gp(L.p_w));
      }
      VC_CUDA(cudaGetLastError());
      VC_TRY(grad_ready_params({vgg[0].p_w, vgg[0].p_b, vgg[1].p_w, vgg[1].p_b}, s));
      break;
    }
    const void* x_in = vgg[l - 1].pool ? vgg[l - 1].pooled : vgg[l - 1].out;
    static const char* kWg[13] = {"wgrad1_1", "wgrad1_2", "wgrad2_1", "wgrad2_2", "wgrad3_1", "wgrad3_2", "wgrad3_3", "wgrad4_1",
                                  "wgrad4_2", "wgrad4_3", "wgrad5_1", "wgrad5_2", "wgrad5_3"};
    static const char* kDg[13] = {"", "dgrad1_2", "dgrad2_1", "dgrad2_2", "dgrad3_1", "dgrad3_2", "dgrad3_3", "dgrad4_1",
                                  "dgrad4_2", "dgrad4_3", "dgrad5_1", "dgrad5_2", "dgrad5_3"};
    VC_TRY(conv3x3_wgrad(s, x_in, dY, gp(L.p_w), B, L.hw, L.cin, L.cout, kWg[l]));
    // the gradient this produces is the one of layer l - 1's output: if that layer is not pooled, its ReLU derivative and
    // bias gradient are taken in the epilogue (no separate pass over dA / out / dY)
    // (not for the 64-channel map of conv1_1: a 128 x 64 tile leaves its epilogue no slack -- dgrad1_2 0.79 -> 1.86 ms
    // fused, against 0.85 ms for the separate pass)
    masked = fuse_relu && !vgg[l - 1].pool && L.cin >= 128;
    VC_TRY(conv3x3_dgrad(s, dY, L.wt_d, dA, B, L.hw, L.cin, L.cout, kDg[l], masked ? vgg[l - 1].out : nullptr,
                         masked ? gp(vgg[l - 1].p_b) : nullptr));
    // one bucket per VGG block (conv5_x, conv4_x, conv3_x, conv2_x; conv1_x after the loop's last iteration)
    if (l == 10 || l == 7 || l == 4)
      VC_TRY(grad_ready_params({vgg[l].p_w, vgg[l].p_b, vgg[l + 1].p_w, vgg[l + 1].p_b, vgg[l + 2].p_w, vgg[l + 2].p_b}, s));
    if (l == 2) VC_TRY(grad_ready_params({vgg[2].p_w, vgg[2].p_b, vgg[3].p_w, vgg[3].p_b}, s));
  }
  return VC_OK;
}

}  // namespace vc
