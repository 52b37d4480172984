// LSTM recurrence on the tcgen05 mainloop (reference: tf.contrib.rnn.LSTMCell driven by
// cell(...) pre-steps and tf.nn.dynamic_rnn -- vae_model/encoder.py:38-58, decoder.py:91-121,
// utils/rnn_model.py:23-51; gate order i, j, f, o and forget_bias = 1).
//
// Forward step: one GEMM [N, E+H] x [E+H, 4H] whose A operand is the concatenation of x_t (tensor
// map 1) and h_{t-1} (tensor map 2) and whose B operand is the gate-interleaved transposed weight
// shadow, so every 128x256 accumulator tile holds all four gates of 64 hidden units and the
// sigmoid/tanh/cell update is fused into the epilogue (no gate tensor round trip through HBM
// other than the bf16 activations kept for BPTT).
// Backward step: dh_rec = dG_{t+1} x W_h^T with the gate-gradient computation of step t fused
// into the epilogue.
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
  static constexpr int kSmemBytes = 0;
  __device__ __forceinline__ void finish() const {}

  __device__ __forceinline__ void operator()(uint32_t taddr, const GemmCore&, const TileCoord& tc, int row, uint8_t*,
                                             int, int&) const {
    const int m = tc.m_blk * kBM + row;
    const bool row_ok = m < N;
    const bool live = row_ok && (lengths == nullptr || t < lengths[m]);
    const int u_base = tc.n_blk * kUPT;
#pragma unroll 1
    for (int c = 0; c < kUPT; c += 16) {
      float gi[16], gj[16], gf[16], go[16];
      __syncwarp();
      tmem_ld16(taddr + 0 * kUPT + c, gi);
      tmem_ld16(taddr + 1 * kUPT + c, gj);
      tmem_ld16(taddr + 2 * kUPT + c, gf);
      tmem_ld16(taddr + 3 * kUPT + c, go);
      tmem_ld_wait();
      if (!row_ok) continue;
      const int u0 = u_base + c;
      const long long o = (long long)m * H + u0;
      if (live) {
        float cp[16], hn[16];
#pragma unroll
        for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(cp + j) = *reinterpret_cast<const float4*>(c_prev + o + j);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float i_ = sigmoidf_(gi[j] + bias[u0 + j]);
          const float j_ = tanhf(gj[j] + bias[H + u0 + j]);
          const float f_ = sigmoidf_(gf[j] + bias[2 * H + u0 + j] + 1.0f);
          const float o_ = sigmoidf_(go[j] + bias[3 * H + u0 + j]);
          const float cn = f_ * cp[j] + i_ * j_;
          cp[j] = cn;
          hn[j] = o_ * tanhf(cn);
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
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              // the emitted value is the bf16 state scaled by the keep mask (the state itself is untouched)
              const float hb = __bfloat162float(__float2bfloat16(hn[j]));
              hn[j] = hb * out_keep[(long long)m * out_keep_ld + u0 + j] * inv_keep;
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
        for (int j = 0; j < 16; j += 4)
          *reinterpret_cast<float4*>(c_out + o + j) = *reinterpret_cast<const float4*>(c_prev + o + j);
        *reinterpret_cast<uint4*>(h_out + o) = *reinterpret_cast<const uint4*>(h_prev + o);
        *reinterpret_cast<uint4*>(h_out + o + 8) = *reinterpret_cast<const uint4*>(h_prev + o + 8);
        if (out != nullptr) {
          *reinterpret_cast<uint4*>(out + o) = make_uint4(0, 0, 0, 0);
          *reinterpret_cast<uint4*>(out + o + 8) = make_uint4(0, 0, 0, 0);
        }
      }
    }
  }
};

int lstm_fwd_step(cudaStream_t stream, const LstmFwdArgs& a) {
  if (a.H % kUPT != 0 || a.E % kBK != 0 || a.H % kBK != 0)
    return set_error(VC_E_SHAPE, "LSTM sizes must be multiples of 64 (E=%d H=%d)", a.E, a.H);
  Operand X{a.x, a.N, a.E, a.E, false};
  Operand Hp{a.h_prev, a.N, a.H, a.H, false};
  Operand W{a.w_t_perm, 4LL * a.H, (long long)a.E + a.H, (long long)a.E + a.H, false};
  GemmPlan plan;
  VC_TRY(plan_gemm(&plan, X, &Hp, a.E, W, a.N, 4 * a.H, a.E + a.H, 4 * kUPT, 1));
  EpiLstmFwd epi;
  epi.bias = a.bias;
  epi.c_prev = a.c_prev;
  epi.c_out = a.c_out;
  epi.h_prev = (const __nv_bfloat16*)a.h_prev;
  epi.h_out = (__nv_bfloat16*)a.h_out;
  epi.gates = (__nv_bfloat16*)a.gates;
  epi.out = (__nv_bfloat16*)a.out;
  epi.lengths = a.lengths;
  epi.out_keep = a.out_keep;
  epi.out_keep_ld = a.out_keep_ld;
  epi.inv_keep = a.inv_keep;
  epi.t = a.t;
  epi.N = a.N;
  epi.H = a.H;
  ProfTag tag("lstm_fwd_step");
  return launch_gemm(plan, epi, stream);
}

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
    const bool live = (lengths == nullptr || t < lengths[m]);
    float dhc[16], dcc[16];
#pragma unroll
    for (int j = 0; j < 16; j += 4) {
      *reinterpret_cast<float4*>(dhc + j) = *reinterpret_cast<const float4*>(dh_carry + o + j);
      *reinterpret_cast<float4*>(dcc + j) = *reinterpret_cast<const float4*>(dc_carry + o + j);
    }
    __nv_bfloat16* dgp = d_gates + (long long)m * 4 * H + u0;
    if (live) {
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
        if (d_out != nullptr) {
          float d = d_out[o + j];
          if (out_keep != nullptr) d *= out_keep[(long long)m * out_keep_ld + u0 + j] * inv_keep;
          dh += d;
        }
        const float tc = tanhf(cc[j]);
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
#pragma unroll
    for (int j = 0; j < 16; j += 4) {
      *reinterpret_cast<float4*>(dh_carry + o + j) = *reinterpret_cast<float4*>(dhc + j);
      *reinterpret_cast<float4*>(dc_carry + o + j) = *reinterpret_cast<float4*>(dcc + j);
    }
  }
};

constexpr int kBwdBN = 64;

struct EpiLstmBwd {
  LstmBwdCommon c;
  static constexpr int kSmemBytes = 0;
  __device__ __forceinline__ void finish() const {}
  __device__ __forceinline__ void operator()(uint32_t taddr, const GemmCore&, const TileCoord& tc, int row, uint8_t*,
                                             int, int&) const {
    const int m = tc.m_blk * kBM + row;
#pragma unroll 1
    for (int col = 0; col < kBwdBN; col += 16) {
      float acc[16];
      __syncwarp();
      tmem_ld16(taddr + col, acc);
      tmem_ld_wait();
      if (m < c.N) c.run(acc, m, tc.n_blk * kBwdBN + col);
    }
  }
};

__global__ void k_lstm_bwd_last(LstmBwdCommon c) {
  const long long chunks = (long long)c.N * (c.H / 16);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < chunks; i += (long long)gridDim.x * blockDim.x) {
    const int m = (int)(i / (c.H / 16));
    const int u0 = (int)(i - (long long)m * (c.H / 16)) * 16;
    float acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = 0.f;
    c.run(acc, m, u0);
  }
}

int lstm_bwd_step(cudaStream_t stream, const LstmBwdArgs& a) {
  if (a.H % kBwdBN != 0) return set_error(VC_E_SHAPE, "LSTM hidden size must be a multiple of 64");
  LstmBwdCommon c;
  c.gates = (const __nv_bfloat16*)a.gates;
  c.c_prev = a.c_prev;
  c.c_cur = a.c_cur;
  c.d_out = a.d_out;
  c.out_keep = a.out_keep;
  c.out_keep_ld = a.out_keep_ld;
  c.inv_keep = a.inv_keep;
  c.dh_carry = a.dh_carry;
  c.dc_carry = a.dc_carry;
  c.d_gates = (__nv_bfloat16*)a.d_gates;
  c.lengths = a.lengths;
  c.t = a.t;
  c.N = a.N;
  c.H = a.H;
  if (a.d_gates_next == nullptr) {
    const long long chunks = (long long)a.N * (a.H / 16);
    int grid = (int)((chunks + 127) / 128);
    {
      ProfScope ps(stream, "lstm_bwd_last");
      k_lstm_bwd_last<<<grid, 128, 0, stream>>>(c);
    }
    VC_CUDA(cudaGetLastError());
    return VC_OK;
  }
  // dh_rec[N, H] = dG_{t+1}[N, 4H] x W_h^T : B = rows E..E+H of the natural-layout weight shadow [E+H, 4H]
  Operand A{a.d_gates_next, a.N, 4LL * a.H, 4LL * a.H, false};
  Operand B{(const __nv_bfloat16*)a.w_nat + (long long)a.E * 4 * a.H, a.H, 4LL * a.H, 4LL * a.H, false};
  GemmPlan plan;
  VC_TRY(plan_gemm(&plan, A, nullptr, 0, B, a.N, a.H, 4 * a.H, kBwdBN, 1));
  EpiLstmBwd epi{c};
  ProfTag tag("lstm_bwd_step");
  return launch_gemm(plan, epi, stream);
}

}  // namespace vc
