// Epilogue math shared by the per-step LSTM kernels (lstm.cu) and the whole-sequence persistent kernels (lstm_seq.cu):
// tf.contrib.rnn.LSTMCell semantics (gate order i, j, f, o; forget_bias = 1) and dynamic_rnn's
// sequence_length handling (state copied through, zero output past the end).
#pragma once
#include "ops.h"

namespace vc {

constexpr int kUPT = 64;  // hidden units per forward tile (4 gates x 64 = 256 accumulator columns)

struct EpiLstmFwd {
  const float* bias;              // [4H] natural gate order (i | j | f | o)
  const float* c_prev;            // [N, H]
  float* c_out;                   // [N, H]
  const __nv_bfloat16* h_prev;    // [N, H]
  __nv_bfloat16* h_out;           // [N, H] carried state
  __nv_bfloat16* gates;           // [N, 4H] activated gates for BPTT (nullable)
  __nv_bfloat16* out;             // [N, H] emitted output (nullable): zero past the sequence end
  const int* lengths;             // nullable: step is live for every row
  const float* out_keep;          // nullable: DropoutWrapper(output_keep_prob) keep mask [N, H] for this step
  long long out_keep_ld;
  float inv_keep;
  int t, N, H;
  int precise;  // 1: expf / tanhf (generation: token ties are decided at the 1e-4 level); 0: MUFU tanh.approx (training)
  static constexpr int kSmemBytes = 0;
  static constexpr bool kPairOk = false;
  __device__ __forceinline__ void finish(uint8_t*, int) const {}

  __device__ __forceinline__ void operator()(uint32_t taddr, const GemmCore&, const TileCoord& tc, int row, uint8_t*,
                                             int, int&) const {
    run(taddr, tc.m_blk, tc.n_blk, row, 0, kUPT);
  }

  // hidden units [c_begin, c_end) of the tile (multiples of 16); cell state read from / written to global memory
  __device__ __forceinline__ void run(uint32_t taddr, int m_blk, int n_blk, int row, int c_begin, int c_end) const {
    const int m = m_blk * kBM + row;
    const bool row_ok = m < N;
    const bool live = row_ok && (lengths == nullptr || t < lengths[m]);
#pragma unroll 1
    for (int c = c_begin; c < c_end; c += 16) {
      float cp[16];
      chunk<false>(taddr, m, row_ok, live, n_blk * kUPT, c, cp);
    }
  }

  // One 16-unit chunk of one row. kCReg: cp[] holds c_prev on entry (register-resident state of the persistent
  // kernel) and the new cell state on return; otherwise c_prev is read from global memory. c_out is always written
  // (BPTT needs every step's cell state).
  template <bool kCReg>
  __device__ __forceinline__ void chunk(uint32_t taddr, int m, bool row_ok, bool live, int u_base, int c, float* cp) const {
    float gi[16], gj[16], gf[16], go[16];
    __syncwarp();
    tmem_ld16(taddr + 0 * kUPT + c, gi);
    tmem_ld16(taddr + 1 * kUPT + c, gj);
    tmem_ld16(taddr + 2 * kUPT + c, gf);
    tmem_ld16(taddr + 3 * kUPT + c, go);
    const int u0 = u_base + c;
    const long long o = (long long)m * H + u0;
    // operands that do not depend on the accumulator are fetched while the TMEM loads are in flight
    float bi[16], bj[16], bf[16], bo[16];
#pragma unroll
    for (int j = 0; j < 16; j += 4) {
      *reinterpret_cast<float4*>(bi + j) = __ldg(reinterpret_cast<const float4*>(bias + u0 + j));
      *reinterpret_cast<float4*>(bj + j) = __ldg(reinterpret_cast<const float4*>(bias + H + u0 + j));
      *reinterpret_cast<float4*>(bf + j) = __ldg(reinterpret_cast<const float4*>(bias + 2 * H + u0 + j));
      *reinterpret_cast<float4*>(bo + j) = __ldg(reinterpret_cast<const float4*>(bias + 3 * H + u0 + j));
    }
    if (!kCReg && row_ok) {
#pragma unroll
      for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(cp + j) = *reinterpret_cast<const float4*>(c_prev + o + j);
    }
    tmem_ld_wait();
    if (!row_ok) return;
    if (live) {
      float hn[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float i_, j_, f_, o_, cn;
        if (precise) {
          i_ = sigmoidf_(gi[j] + bi[j]);
          j_ = tanhf(gj[j] + bj[j]);
          f_ = sigmoidf_(gf[j] + bf[j] + 1.0f);
          o_ = sigmoidf_(go[j] + bo[j]);
          cn = f_ * cp[j] + i_ * j_;
          hn[j] = o_ * tanhf(cn);
        } else {
          i_ = sigmoid_fast(gi[j] + bi[j]);
          j_ = tanh_fast(gj[j] + bj[j]);
          f_ = sigmoid_fast(gf[j] + bf[j] + 1.0f);
          o_ = sigmoid_fast(go[j] + bo[j]);
          cn = f_ * cp[j] + i_ * j_;
          hn[j] = o_ * tanh_fast(cn);
        }
        cp[j] = cn;
        gi[j] = i_; gj[j] = j_; gf[j] = f_; go[j] = o_;
      }
#pragma unroll
      for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(c_out + o + j) = *reinterpret_cast<float4*>(cp + j);
      uint4 hv[2];
      uint32_t* hw = reinterpret_cast<uint32_t*>(hv);
#pragma unroll
      for (int j = 0; j < 8; ++j) hw[j] = pack_bf16(hn[2 * j], hn[2 * j + 1]);
      *reinterpret_cast<uint4*>(h_out + o) = hv[0];
      *reinterpret_cast<uint4*>(h_out + o + 8) = hv[1];
      if (out != nullptr) {
        if (out_keep != nullptr) {
          float kp[16];
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            *reinterpret_cast<float4*>(kp + j) = *reinterpret_cast<const float4*>(out_keep + (long long)m * out_keep_ld + u0 + j);
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            // the emitted value is the bf16 state scaled by the keep mask (the state itself is untouched)
            const float hb = __bfloat162float(__float2bfloat16(hn[j]));
            hn[j] = hb * kp[j] * inv_keep;
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) hw[j] = pack_bf16(hn[2 * j], hn[2 * j + 1]);
        }
        *reinterpret_cast<uint4*>(out + o) = hv[0];
        *reinterpret_cast<uint4*>(out + o + 8) = hv[1];
      }
      if (gates != nullptr) {
        __nv_bfloat16* gp = gates + (long long)m * 4 * H + u0;
        float* gsrc[4] = {gi, gj, gf, go};
#pragma unroll
        for (int g = 0; g < 4; ++g) {
#pragma unroll
          for (int j = 0; j < 8; ++j) hw[j] = pack_bf16(gsrc[g][2 * j], gsrc[g][2 * j + 1]);
          *reinterpret_cast<uint4*>(gp + (long long)g * H) = hv[0];
          *reinterpret_cast<uint4*>(gp + (long long)g * H + 8) = hv[1];
        }
      }
    } else {
      // past the end of the sequence: state copied through, emitted output zero
#pragma unroll
      for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(c_out + o + j) = *reinterpret_cast<float4*>(cp + j);
      *reinterpret_cast<uint4*>(h_out + o) = *reinterpret_cast<const uint4*>(h_prev + o);
      *reinterpret_cast<uint4*>(h_out + o + 8) = *reinterpret_cast<const uint4*>(h_prev + o + 8);
      if (out != nullptr) {
        *reinterpret_cast<uint4*>(out + o) = make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4*>(out + o + 8) = make_uint4(0, 0, 0, 0);
      }
    }
  }
};

// ------------------------------------------------------------------------------------------
// Gate gradients of one step for 16 consecutive hidden units of one row.
struct LstmBwdCommon {
  const __nv_bfloat16* gates;   // [N, 4H] activated gates of this step
  const float* c_prev;          // [N, H] cell state before this step
  const float* c_cur;           // [N, H] cell state after this step
  const float* d_out;           // [N, H] gradient of the emitted output (nullable)
  const float* out_keep;        // keep mask of the emitted output (nullable)
  long long out_keep_ld;
  float inv_keep;
  float* dh_carry;              // [N, H] in: pass-through/final-state dh; out: pass-through dh for step t-1
  float* dc_carry;              // [N, H] in: dc of state after this step; out: dc of state before this step
  __nv_bfloat16* d_gates;       // [N, 4H] pre-activation gate gradients of this step
  const int* lengths;
  int t, N, H;

  __device__ __forceinline__ void run(const float* acc, int m, int u0) const {
    const long long o = (long long)m * H + u0;
    float dhc[16], dcc[16];
#pragma unroll
    for (int j = 0; j < 16; j += 4) {
      *reinterpret_cast<float4*>(dhc + j) = *reinterpret_cast<const float4*>(dh_carry + o + j);
      *reinterpret_cast<float4*>(dcc + j) = *reinterpret_cast<const float4*>(dc_carry + o + j);
    }
    core(acc, m, u0, dhc, dcc);
#pragma unroll
    for (int j = 0; j < 16; j += 4) {
      *reinterpret_cast<float4*>(dh_carry + o + j) = *reinterpret_cast<float4*>(dhc + j);
      *reinterpret_cast<float4*>(dc_carry + o + j) = *reinterpret_cast<float4*>(dcc + j);
    }
  }

  // dhc / dcc: in = pass-through dh and dc of the state after this step; out = the same for the state before it
  __device__ __forceinline__ void core(const float* acc, int m, int u0, float* dhc, float* dcc) const {
    const long long o = (long long)m * H + u0;
    const bool live = (lengths == nullptr || t < lengths[m]);
    __nv_bfloat16* dgp = d_gates + (long long)m * 4 * H + u0;
    if (live) {
      float dov[16];
      if (d_out != nullptr) {
#pragma unroll
        for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(dov + j) = *reinterpret_cast<const float4*>(d_out + o + j);
        if (out_keep != nullptr) {
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            const float4 k = *reinterpret_cast<const float4*>(out_keep + (long long)m * out_keep_ld + u0 + j);
            dov[j] *= k.x * inv_keep; dov[j + 1] *= k.y * inv_keep; dov[j + 2] *= k.z * inv_keep; dov[j + 3] *= k.w * inv_keep;
          }
        }
      }
      const __nv_bfloat16* gp = gates + (long long)m * 4 * H + u0;
      float gi[16], gj[16], gf[16], go[16], cp[16], cc[16];
      auto ld16 = [](const __nv_bfloat16* p, float* dst) {
        uint4 a = *reinterpret_cast<const uint4*>(p), b = *reinterpret_cast<const uint4*>(p + 8);
        const __nv_bfloat162* ha = reinterpret_cast<const __nv_bfloat162*>(&a);
        const __nv_bfloat162* hb = reinterpret_cast<const __nv_bfloat162*>(&b);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 fa = __bfloat1622float2(ha[j]), fb = __bfloat1622float2(hb[j]);
          dst[2 * j] = fa.x; dst[2 * j + 1] = fa.y;
          dst[8 + 2 * j] = fb.x; dst[8 + 2 * j + 1] = fb.y;
        }
      };
      ld16(gp, gi);
      ld16(gp + H, gj);
      ld16(gp + 2 * H, gf);
      ld16(gp + 3 * H, go);
#pragma unroll
      for (int j = 0; j < 16; j += 4) {
        *reinterpret_cast<float4*>(cp + j) = *reinterpret_cast<const float4*>(c_prev + o + j);
        *reinterpret_cast<float4*>(cc + j) = *reinterpret_cast<const float4*>(c_cur + o + j);
      }
      float dgi[16], dgj[16], dgf[16], dgo[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float dh = acc[j] + dhc[j];
        if (d_out != nullptr) dh += dov[j];
        const float tc = tanh_fast(cc[j]);
        const float dc = dcc[j] + dh * go[j] * (1.f - tc * tc);
        dgo[j] = dh * tc * go[j] * (1.f - go[j]);
        dgi[j] = dc * gj[j] * gi[j] * (1.f - gi[j]);
        dgj[j] = dc * gi[j] * (1.f - gj[j] * gj[j]);
        dgf[j] = dc * cp[j] * gf[j] * (1.f - gf[j]);
        dcc[j] = dc * gf[j];
        dhc[j] = 0.f;
      }
      float* src[4] = {dgi, dgj, dgf, dgo};
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint4 hv[2];
        uint32_t* hw = reinterpret_cast<uint32_t*>(hv);
#pragma unroll
        for (int j = 0; j < 8; ++j) hw[j] = pack_bf16(src[g][2 * j], src[g][2 * j + 1]);
        *reinterpret_cast<uint4*>(dgp + (long long)g * H) = hv[0];
        *reinterpret_cast<uint4*>(dgp + (long long)g * H + 8) = hv[1];
      }
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) dhc[j] += acc[j];
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        *reinterpret_cast<uint4*>(dgp + (long long)g * H) = make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4*>(dgp + (long long)g * H + 8) = make_uint4(0, 0, 0, 0);
      }
    }
  }
};

}  // namespace vc
