// VGG16 feature extractor forward (reference: utils/image_embeddings.py:26-238 -- mean subtraction,
// 13 x (3x3 SAME conv + bias + ReLU), 5 x 2x2/2 max-pool, fc1 + ReLU, fc2 + ReLU -> fc2 [B, 4096]).
//
// Every convolution is an implicit GEMM on the tcgen05 mainloop: the A operand is gathered by TMA
// straight from the NHWC bf16 activation (one 4-D box per 32-pixel patch and filter tap, out-of-bounds
// zero fill = SAME padding), the B operand is the [Cout, 9*Cin] weight shadow, and bias + ReLU
// (+ the 2x2 max-pool of the following layer) run in the epilogue, which stores bf16 NHWC through TMA.
// conv1_1 (Cin = 3) goes through a tiny im2col (27 -> 32 columns) fused with the mean subtraction.
#include <curand_kernel.h>
#include "model.h"

namespace vc {

static const char* kVggNames[13] = {"conv1_1", "conv1_2", "conv2_1", "conv2_2", "conv3_1", "conv3_2", "conv3_3",
                                    "conv4_1", "conv4_2", "conv4_3", "conv5_1", "conv5_2", "conv5_3"};
static const int kVggCin[13] = {3, 64, 64, 128, 128, 256, 256, 256, 512, 512, 512, 512, 512};
static const int kVggCout[13] = {64, 64, 128, 128, 256, 256, 256, 512, 512, 512, 512, 512, 512};
static const int kVggHW[13] = {224, 224, 112, 112, 56, 56, 56, 28, 28, 28, 14, 14, 14};
static const bool kVggPool[13] = {false, true, false, true, false, false, true, false, false, true, false, false, true};

const char* vgg_layer_name(int l) { return kVggNames[l]; }
int vgg_layer_cin(int l) { return kVggCin[l]; }
int vgg_layer_cout(int l) { return kVggCout[l]; }
int vgg_layer_hw(int l) { return kVggHW[l]; }
bool vgg_layer_pool(int l) { return kVggPool[l]; }

// ------------------------------------------------------------------------------------------
// images fp32 [B,224,224,3] (RGB 0..255) -> A bf16 [B*224*224, 32]: column (r*3+s)*3+c holds
// (pixel(h+r-1, w+s-1, c) - mean[c]) or 0 outside the image; columns 27..31 are zero.
template <class TIn>
__global__ void k_im2col_rgb(const TIn* __restrict__ img, __nv_bfloat16* __restrict__ A, int B, int H, int W) {
  const long long total = (long long)B * H * W;
  const float mean[3] = {123.68f, 116.779f, 103.939f};  // image_embeddings.py:30-34
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int w = (int)(i % W);
    const long long r0 = i / W;
    const int h = (int)(r0 % H);
    const long long b = r0 / H;
    float v[32];
#pragma unroll
    for (int j = 27; j < 32; ++j) v[j] = 0.f;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const int hh = h + r - 1, ww = w + s - 1;
        const bool ok = hh >= 0 && hh < H && ww >= 0 && ww < W;
        const TIn* p = img + ((b * H + hh) * W + ww) * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) v[(r * 3 + s) * 3 + c] = ok ? static_cast<float>(p[c]) - mean[c] : 0.f;
      }
    }
    uint4* dst = reinterpret_cast<uint4*>(A + i * 32);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint4 u;
      u.x = pack_bf16(v[8 * j + 0], v[8 * j + 1]);
      u.y = pack_bf16(v[8 * j + 2], v[8 * j + 3]);
      u.z = pack_bf16(v[8 * j + 4], v[8 * j + 5]);
      u.w = pack_bf16(v[8 * j + 6], v[8 * j + 7]);
      dst[j] = u;
    }
  }
}

// images [B,H,W,3] (RGB 0..255, fp32 or uint8) -> P bf16 [B, H+2, W+2, 8]: pixel (x+1, y+1) = (r - mean_r, g - mean_g,
// b - mean_b, 0, 0, 0, 0, 0); the one-pixel border is never written and stays zero (SAME padding). One 16-byte store per pixel.
template <class TIn>
__global__ void k_pad_rgb8(const TIn* __restrict__ img, uint4* __restrict__ P, int B, int H, int W) {
  const long long total = (long long)B * H * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int w = (int)(i % W);
    const long long r0 = i / W;
    const int h = (int)(r0 % H);
    const long long b = r0 / H;
    const TIn* p = img + i * 3;
    uint4 o;
    o.x = pack_bf16(static_cast<float>(p[0]) - 123.68f, static_cast<float>(p[1]) - 116.779f);  // image_embeddings.py:30-34
    o.y = pack_bf16(static_cast<float>(p[2]) - 103.939f, 0.f);
    o.z = 0u;
    o.w = 0u;
    P[(b * (H + 2) + h + 1) * (W + 2) + w + 1] = o;
  }
}
// wt bf16 [64, 192] for the window form of conv1_1 (plan_conv1_window): k = r*64 + s*8 + c <- w[r][s][c][co], zeros elsewhere
__global__ void k_conv1_window_shadow(const float* __restrict__ w, __nv_bfloat16* __restrict__ wt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 64 * 192) {
    const int co = i / 192, k = i % 192;
    const int r = k / 64, rem = k % 64, s = rem / 8, c = rem % 8;
    wt[i] = __float2bfloat16((s < 3 && c < 3) ? w[((r * 3 + s) * 3 + c) * 64 + co] : 0.f);
  }
}

// 2x2 stride-2 max-pool over NHWC bf16 (standalone form, used when un-pooled activations are kept).
__global__ void k_maxpool2(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, int B, int H, int W,
                           int C) {
  const int Ho = H / 2, Wo = W / 2, C8 = C / 8;
  const long long total = (long long)B * Ho * Wo * C8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % C8);
    long long r = i / C8;
    const int wo = (int)(r % Wo);
    r /= Wo;
    const int ho = (int)(r % Ho);
    const long long b = r / Ho;
    const __nv_bfloat16* p = in + (((b * H + 2 * ho) * W + 2 * wo) * C) + c8 * 8;
    uint4 q[4] = {*reinterpret_cast<const uint4*>(p), *reinterpret_cast<const uint4*>(p + C),
                  *reinterpret_cast<const uint4*>(p + (long long)W * C), *reinterpret_cast<const uint4*>(p + (long long)W * C + C)};
    uint4 o;
    __nv_bfloat162* ov = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      __nv_bfloat162 m = __hmax2(reinterpret_cast<__nv_bfloat162*>(&q[0])[j], reinterpret_cast<__nv_bfloat162*>(&q[1])[j]);
      m = __hmax2(m, reinterpret_cast<__nv_bfloat162*>(&q[2])[j]);
      ov[j] = __hmax2(m, reinterpret_cast<__nv_bfloat162*>(&q[3])[j]);
    }
    *reinterpret_cast<uint4*>(out + (((b * Ho + ho) * Wo + wo) * C) + c8 * 8) = o;
  }
}

// y = relu(x + bias) (* keep / keep_prob) -> fp32 and/or bf16 (fc layers after the split-K accumulation).
// tf.nn.dropout (image_embeddings.py:225-226, 236-237): explicit 0/1 mask `keep`, or, when philox != 0, a
// Philox4x32-10 uniform per element (subsequence = element index, offset = stream) kept if u < keep_prob.
__global__ void k_bias_relu(const float* __restrict__ x, const float* __restrict__ bias, const float* __restrict__ keep,
                            float inv_keep, float* __restrict__ yf, __nv_bfloat16* __restrict__ yh, long long rows, int cols,
                            int philox, unsigned long long seed, unsigned long long stream) {
  const long long total = rows * cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    float v = fmaxf(x[i] + bias[i % cols], 0.f);
    if (keep) v *= keep[i] * inv_keep;
    else if (philox) {
      curandStatePhilox4_32_10_t st;
      curand_init(seed, (unsigned long long)i, stream, &st);
      v = (curand_uniform(&st) * inv_keep <= 1.f) ? v * inv_keep : 0.f;  // u <= keep_prob
    }
    if (yf) yf[i] = v;
    if (yh) yh[i] = __float2bfloat16(v);
  }
}

// conv1_1 runs as a [pixels/2, 64] x [64, 128] GEMM: one A row holds the 32-column im2col patches of TWO adjacent pixels
// (a full 128-byte TMA row, no out-of-bounds half), B is the block-diagonal [[W, 0], [0, W]] so the 128 output columns
// are the 64 channels of pixel 2i followed by those of pixel 2i+1 -- exactly the NHWC layout of the output.
// wt bf16 [128, 64] (K-major rows n = pixel*64 + cout), bias2 fp32 [128].
__global__ void k_conv1_shadow(const float* __restrict__ w, const float* __restrict__ bias, __nv_bfloat16* __restrict__ wt,
                               float* __restrict__ bias2) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 128 * 64) {
    const int n = i / 64, k = i % 64;
    const int pix = n / 64, co = n % 64, kk = k - pix * 32;
    wt[i] = __float2bfloat16((kk >= 0 && kk < 27) ? w[kk * 64 + co] : 0.f);
  }
  if (i < 128) bias2[i] = bias[i % 64];
}

static inline int ew_grid(long long n, int block) {
  long long g = (n + block - 1) / block;
  const long long cap = (long long)num_sms() * 16;
  return (int)(g > cap ? cap : (g < 1 ? 1 : g));
}

// ------------------------------------------------------------------------------------------
int Model::vgg_init() {
  const int B = cfg.max_batch;
  vgg.resize(13);
  for (int l = 0; l < 13; ++l) {
    VggLayer& L = vgg[l];
    L.cin = kVggCin[l]; L.cout = kVggCout[l]; L.hw = kVggHW[l]; L.pool = kVggPool[l];
    const bool c5 = l >= 10;
    L.p_w = pidx(std::string("cnn/") + kVggNames[l] + (c5 ? "/weights_conv" : "/weights"));
    L.p_b = pidx(std::string("cnn/") + kVggNames[l] + (c5 ? "/biases_conv" : "/biases"));
    L.kpad = l == 0 ? 192 : 9 * L.cin;  // conv1_1: [64, 192] window form, or block-diagonal [128, 64] (k_conv1_shadow)
    VC_TRY(dalloc((uint16_t**)&L.wt, (size_t)L.cout * L.kpad));
    VC_TRY(dalloc((uint16_t**)&L.out, (size_t)B * L.hw * L.hw * L.cout));
    if (L.pool) VC_TRY(dalloc((uint16_t**)&L.pooled, (size_t)B * (L.hw / 2) * (L.hw / 2) * L.cout));
  }
  VC_TRY(dalloc((uint16_t**)&vgg_im2col, (size_t)B * 224 * 224 * 32));
  {  // conv1_1 form: im2col + two-pixel-row GEMM (default) | window (overlapping-stride TMA map, VC_CONV1=window)
    const char* e = getenv("VC_CONV1");
    conv1_window = e && strcmp(e, "window") == 0;
  }
  if (conv1_window) VC_TRY(dalloc((uint16_t**)&vgg_padded, (size_t)B * 226 * 226 * 8 + 256));  // + slack: the last window overruns
  VC_TRY(dalloc((uint16_t**)&fc1_w, (size_t)25088 * 4096));
  VC_TRY(dalloc((uint16_t**)&fc2_w, (size_t)4096 * 4096));
  VC_TRY(dalloc(&conv1_bias2, 128));
  VC_TRY(dalloc(&fc_acc, (size_t)B * 4096));
  VC_TRY(dalloc((uint16_t**)&fc1_h, (size_t)B * 4096));
  VC_TRY(dalloc(&fc2_f, (size_t)B * 4096));
  VC_TRY(dalloc(&st_images, (size_t)B * 224 * 224 * 3));
  vgg_shadows_dirty = true;
  if (cfg.fine_tune) VC_TRY(vgg_bwd_init());
  return VC_OK;
}

int Model::vgg_refresh_shadows(cudaStream_t s) {
  ProfTag ptag("refresh_shadows");
  {
    ProfScope ps(s, "refresh_shadows");
    if (conv1_window)
      k_conv1_window_shadow<<<48, 256, 0, s>>>(pp(vgg[0].p_w), (__nv_bfloat16*)vgg[0].wt);
    else
      k_conv1_shadow<<<32, 256, 0, s>>>(pp(vgg[0].p_w), pp(vgg[0].p_b), (__nv_bfloat16*)vgg[0].wt, conv1_bias2);
  }
  RefreshJobs jobs;  // every other cnn/ shadow in one launch (28 launches / 0.65 ms per fine-tune step before, 0.34 ms now)
  for (int l = 1; l < 13; ++l) {
    VggLayer& L = vgg[l];
    // HWIO [3,3,Cin,Cout] == [9*Cin, Cout] row-major -> [Cout, kpad] (K-major B operand)
    jobs.transpose(pp(L.p_w), L.wt, 9 * L.cin, L.cout, L.cout, L.kpad, 0, 0);
    if (cfg.fine_tune) jobs.dgrad_filter(pp(L.p_w), L.wt_d, L.cin, L.cout);
  }
  jobs.cast(pp(pidx("cnn/fc1/weights")), fc1_w, 25088, 4096, 4096, 4096);
  jobs.cast(pp(pidx("cnn/fc2/weights")), fc2_w, 4096, 4096, 4096, 4096);
  VC_TRY(refresh_multi(s, jobs));
  vgg_shadows_dirty = false;
  return VC_OK;
}

int Model::vgg_conv_layer(int l, const void* in, int B, bool fuse_pool, cudaStream_t s) {
  VggLayer& L = vgg[l];
  const int bn = L.cout >= 256 ? 256 : L.cout;
  EpiTma epi{};
  epi.bias = pp(L.p_b);
  epi.N = L.cout;
  epi.bn = bn;
  epi.relu = 1;
  epi.alpha = 1.f;
  GemmPlan plan;
  ProfTag ptag(kVggNames[l]);
  if (l == 0 && conv1_window) {
    // straight from the padded pixels (no im2col matrix): three k-blocks, one per filter row (plan_conv1_window)
    VC_TRY(plan_conv1_window(&plan, vgg_padded, L.wt, L.hw, L.hw, B, L.cout));
    epi.mode = kConv;
    const GemmCore& g = plan.core;
    VC_TRY(make_tmap_nhwc(&epi.tm, L.out, L.cout, L.hw, L.hw, B, g.pw * g.tw, g.ph * g.th, g.pn * (4 / (g.tw * g.th))));
    return launch_gemm(plan, epi, s);
  }
  if (l == 0) {
    const long long M = (long long)B * L.hw * L.hw / 2;  // two pixels per GEMM row
    Operand A{in, M, 64, 64, false}, Bw{L.wt, 128, 64, 64, false};
    VC_TRY(plan_gemm(&plan, A, nullptr, 0, Bw, (int)M, 128, 64, 128, 1));
    epi.mode = kRows;
    epi.bias = conv1_bias2;
    epi.N = 128;
    epi.bn = 128;
    VC_TRY(make_tmap_2d(&epi.tm, L.out, 128, M, 128, 64, 128));
    return launch_gemm(plan, epi, s);
  }
  ConvGeom g;
  const bool halo = conv_halo_applicable(L.hw, L.hw, L.cin, L.cout);
  const bool halo2 = !halo && conv_halo_stream_applicable(L.hw, L.hw, L.cin, L.cout);
  if (halo) {
    VC_TRY(conv_halo_geometry(&g, L.hw, L.hw, B, L.cin, L.cout));
    VC_TRY(plan_conv_halo(&plan, in, L.wt, g));
    epi.bn = plan.core.bn;
  } else if (halo2) {
    VC_TRY(conv_halo_geometry(&g, L.hw, L.hw, B, L.cin, L.cout));
    VC_TRY(plan_conv_halo_stream(&plan, in, L.wt, g));
    epi.bn = L.cout;
  } else {
    VC_TRY(conv_geometry(&g, L.hw, L.hw, B, L.cin, L.cout));
    VC_TRY(plan_conv(&plan, in, L.wt, g, bn));
  }
  if (fuse_pool && L.pool) {
    epi.mode = kConvPool;
    VC_TRY(make_tmap_nhwc(&epi.tm, L.pooled, L.cout, L.hw / 2, L.hw / 2, B, g.pw * g.tw / 2, g.ph * g.th / 2, g.pn * (4 / (g.tw * g.th))));
  } else {
    epi.mode = kConv;
    VC_TRY(make_tmap_nhwc(&epi.tm, L.out, L.cout, L.hw, L.hw, B, g.pw * g.tw, g.ph * g.th, g.pn * (4 / (g.tw * g.th))));
  }
  if (halo) return launch_conv_halo(plan, epi, s);
  if (halo2) return launch_conv_halo_stream(plan, epi, s);
  return launch_gemm(plan, epi, s);
}

// images: device fp32 [B,224,224,3]; fc2_out: device fp32 [B,4096] (nullable -> only fc2_f is filled).
// keep_unpooled: materialise every conv output (debug taps / fine-tune backward) instead of fusing the pools.
int Model::vgg_forward(const float* images, float* fc2_out, int B, bool keep_unpooled, const float* fc_keep,
                       cudaStream_t s, bool images_u8) {
  if (vgg.empty()) return set_error(VC_E_STATE, "this handle was created without the CNN (with_cnn = 0)");
  if (B < 1 || B > cfg.max_batch) return set_error(VC_E_SHAPE, "vgg_forward: batch %d exceeds max_batch %d", B, cfg.max_batch);
  if (vgg_shadows_dirty) VC_TRY(vgg_refresh_shadows(s));
  if (conv1_window) {
    ProfScope ps(s, "pad_rgb8");
    const long long n = (long long)B * 224 * 224;
    if (images_u8)
      k_pad_rgb8<uint8_t><<<ew_grid(n, 256), 256, 0, s>>>(reinterpret_cast<const uint8_t*>(images), (uint4*)vgg_padded, B, 224, 224);
    else
      k_pad_rgb8<float><<<ew_grid(n, 256), 256, 0, s>>>(images, (uint4*)vgg_padded, B, 224, 224);
  }
  if (!conv1_window || cfg.fine_tune) {  // the im2col matrix: conv1_1's A operand in the older form, and the filter-gradient operand
    ProfScope ps(s, "im2col_rgb");
    const long long n = (long long)B * 224 * 224;
    if (images_u8)  // uint8 pixels as the HDF5 image store keeps them (utils/batch_gen.py:278-294): a quarter of the bytes
      k_im2col_rgb<uint8_t><<<ew_grid(n, 256), 256, 0, s>>>(reinterpret_cast<const uint8_t*>(images), (__nv_bfloat16*)vgg_im2col, B, 224, 224);
    else
      k_im2col_rgb<float><<<ew_grid(n, 256), 256, 0, s>>>(images, (__nv_bfloat16*)vgg_im2col, B, 224, 224);
  }
  VC_CUDA(cudaGetLastError());
  const void* x = vgg_im2col;
  for (int l = 0; l < 13; ++l) {
    VggLayer& L = vgg[l];
    VC_TRY(vgg_conv_layer(l, x, B, !keep_unpooled, s));
    if (L.pool) {
      if (keep_unpooled) {
        ProfScope ps(s, "maxpool2");
        const long long n = (long long)B * (L.hw / 2) * (L.hw / 2) * (L.cout / 8);
        k_maxpool2<<<ew_grid(n, 256), 256, 0, s>>>((const __nv_bfloat16*)L.out, (__nv_bfloat16*)L.pooled, B, L.hw, L.hw,
                                                   L.cout);
        VC_CUDA(cudaGetLastError());
      }
      x = L.pooled;
    } else {
      x = L.out;
    }
  }
  vgg_last_B = B;
  vgg_have_unpooled = keep_unpooled;
  // fc1 / fc2 (image_embeddings.py:214-238): NHWC flatten of pool5 is the row-major [B, 25088] view
  // dropout only exists in the fine-tune training graph (main.py:67-72); explicit masks (parity) win over Philox
  const bool drop = cfg.fine_tune && cfg.cnn_dropout < 1.f;
  const float inv_keep = drop ? 1.f / cfg.cnn_dropout : 1.f;
  if (!drop) fc_keep = nullptr;
  const int philox = (drop && fc_keep == nullptr) ? 1 : 0;
  int fc_no = 0;
  auto fc = [&](const void* a, int K, const void* w, const char* bias_name, const float* keep, float* yf, void* yh) -> int {
    VC_CUDA(cudaMemsetAsync(fc_acc, 0, (size_t)B * 4096 * sizeof(float), s));
    Operand A{a, B, K, K, false}, Bw{w, K, 4096, 4096, true};
    EpiStore e{};
    e.out = fc_acc; e.ld = 4096; e.atomic = 1; e.alpha = 1.f;
    const int tiles = ((B + 127) / 128) * (4096 / 128);
    {
      ProfTag pt("fc");
      VC_TRY(gemm_store(s, A, nullptr, 0, Bw, B, 4096, K, e, 128, std::max(1, num_sms() / tiles)));
    }
    ProfScope ps(s, "bias_relu");
    k_bias_relu<<<ew_grid((long long)B * 4096, 256), 256, 0, s>>>(fc_acc, pp(pidx(bias_name)), keep, inv_keep, yf,
                                                                 (__nv_bfloat16*)yh, B, 4096, philox, vgg_drop_seed,
                                                                 vgg_drop_step * 2 + fc_no);
    ++fc_no;
    return VC_OK;
  };
  VC_TRY(fc(vgg[12].pooled, 25088, fc1_w, "cnn/fc1/biases", fc_keep, nullptr, fc1_h));
  VC_TRY(fc(fc1_h, 4096, fc2_w, "cnn/fc2/biases", fc_keep ? fc_keep + (size_t)B * 4096 : nullptr, fc2_f, nullptr));
  VC_CUDA(cudaGetLastError());
  if (fc2_out != nullptr && fc2_out != fc2_f)
    VC_CUDA(cudaMemcpyAsync(fc2_out, fc2_f, (size_t)B * 4096 * sizeof(float), cudaMemcpyDeviceToDevice, s));
  return VC_OK;
}

int Model::vgg_activation(const char* layer, float* dst_host) {
  if (vgg.empty()) return set_error(VC_E_STATE, "this handle was created without the CNN");
  if (vgg_last_B <= 0) return set_error(VC_E_STATE, "no VGG forward pass has run on this handle");
  VC_CUDA(cudaDeviceSynchronize());
  const int B = vgg_last_B;
  for (int l = 0; l < 13; ++l) {
    if (strcmp(layer, kVggNames[l]) != 0) continue;
    VggLayer& L = vgg[l];
    if (L.pool && !vgg_have_unpooled)
      return set_error(VC_E_STATE, "%s was pooled in the epilogue; run the forward with activations kept", layer);
    const size_t n = (size_t)B * L.hw * L.hw * L.cout;
    float* tmp = nullptr;
    VC_CUDA(cudaMalloc((void**)&tmp, n * sizeof(float)));
    int st = bf16_to_f32(0, L.out, tmp, (long long)B * L.hw * L.hw, L.cout, L.cout, L.cout);
    if (st == VC_OK && cudaMemcpy(dst_host, tmp, n * sizeof(float), cudaMemcpyDeviceToHost) != cudaSuccess)
      st = set_error(VC_E_CUDA, "activation copy failed");
    cudaFree(tmp);
    return st;
  }
  if (strncmp(layer, "pool", 4) == 0) {
    static const int pool_layer[5] = {1, 3, 6, 9, 12};
    const int k = layer[4] - '1';
    if (k >= 0 && k < 5 && layer[5] == 0) {
      VggLayer& L = vgg[pool_layer[k]];
      const size_t n = (size_t)B * (L.hw / 2) * (L.hw / 2) * L.cout;
      float* tmp = nullptr;
      VC_CUDA(cudaMalloc((void**)&tmp, n * sizeof(float)));
      int st = bf16_to_f32(0, L.pooled, tmp, (long long)B * (L.hw / 2) * (L.hw / 2), L.cout, L.cout, L.cout);
      if (st == VC_OK && cudaMemcpy(dst_host, tmp, n * sizeof(float), cudaMemcpyDeviceToHost) != cudaSuccess)
        st = set_error(VC_E_CUDA, "activation copy failed");
      cudaFree(tmp);
      return st;
    }
  }
  if (strcmp(layer, "fc1") == 0 || strcmp(layer, "fc2") == 0) {  // post-ReLU (+ dropout when fine-tuning) [B, 4096]
    const size_t n = (size_t)B * 4096;
    if (layer[2] == '2') {
      VC_CUDA(cudaMemcpy(dst_host, fc2_f, n * sizeof(float), cudaMemcpyDeviceToHost));
      return VC_OK;
    }
    float* tmp = nullptr;
    VC_CUDA(cudaMalloc((void**)&tmp, n * sizeof(float)));
    int st = bf16_to_f32(0, fc1_h, tmp, B, 4096, 4096, 4096);
    if (st == VC_OK && cudaMemcpy(dst_host, tmp, n * sizeof(float), cudaMemcpyDeviceToHost) != cudaSuccess)
      st = set_error(VC_E_CUDA, "activation copy failed");
    cudaFree(tmp);
    return st;
  }
  return set_error(VC_E_ARG, "unknown VGG layer '%s'", layer);
}

}  // namespace vc
