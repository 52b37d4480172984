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
#include "lstm_epi.cuh"

namespace vc {

int lstm_fwd_step(cudaStream_t stream, const LstmFwdArgs& a) {
  if (a.H % kUPT != 0 || a.E % kBK != 0 || a.H % kBK != 0)
    return set_error(VC_E_SHAPE, "LSTM sizes must be multiples of 64 (E=%d H=%d)", a.E, a.H);
  Operand X{a.x, a.N, a.E, a.E, false};
  Operand Hp{a.h_prev, a.N, a.H, a.H, false};
  Operand W{a.w_t_perm, 4LL * a.H, (long long)a.E + a.H, (long long)a.E + a.H, false};
  GemmPlan plan;
  VC_TRY(plan_gemm(&plan, X, &Hp, a.E, W, a.N, 4 * a.H, a.E + a.H, 4 * kUPT, 1, false));
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
  epi.precise = a.precise;
  epi.t = a.t;
  epi.N = a.N;
  epi.H = a.H;
  ProfTag tag("lstm_fwd_step");
  return launch_gemm(plan, epi, stream);
}

constexpr int kBwdBN = 64;

struct EpiLstmBwd {
  LstmBwdCommon c;
  static constexpr int kSmemBytes = 0;
  static constexpr bool kPairOk = false;
  __device__ __forceinline__ void finish(uint8_t*, int) const {}
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
  VC_TRY(plan_gemm(&plan, A, nullptr, 0, B, a.N, a.H, 4 * a.H, kBwdBN, 1, false));
  EpiLstmBwd epi{c};
  ProfTag tag("lstm_bwd_step");
  return launch_gemm(plan, epi, stream);
}

}  // namespace vc
