// Internal host-callable operations of libvaecap (one declaration per kernel family).
#pragma once
#include "host_util.h"

namespace vc {

// gemm_api.cu ------------------------------------------------------------------------------
int gemm_store(cudaStream_t stream, const Operand& A, const Operand* A2, long long a2_at, const Operand& B, int M,
               int N, int K, const EpiStore& epi, int bn, int splits);

int gemm_tma_rows(cudaStream_t stream, const Operand& A, const Operand& B, int M, int N, int K, void* out, long long ldo,
                  const float* bias, int relu, int bn);
int gemm_tma_rows_f32(cudaStream_t stream, const Operand& A, const Operand& B, int M, int N, int K, float* out, long long ldo,
                      const float* bias, int bn);

// lstm.cu ----------------------------------------------------------------------------------
struct LstmFwdArgs {
  const void* x;         // bf16 [N, E] input of this step
  const void* h_prev;    // bf16 [N, H]
  const float* c_prev;   // fp32 [N, H]
  void* h_out;           // bf16 [N, H]
  float* c_out;          // fp32 [N, H]
  void* gates;           // bf16 [N, 4H] (nullable)
  void* out;             // bf16 [N, H] emitted output (nullable)
  const void* w_t_perm;  // bf16 [4H (gate-interleaved), E+H]
  const float* bias;     // fp32 [4H]
  const int* lengths;    // nullable
  const float* out_keep; // nullable
  long long out_keep_ld;
  float inv_keep;
  int t, N, E, H;
  int precise;           // see EpiLstmFwd::precise
};
int lstm_fwd_step(cudaStream_t stream, const LstmFwdArgs& a);

struct LstmBwdArgs {
  const void* d_gates_next;  // bf16 [N, 4H] gate gradients of step t+1 (nullptr for the last step)
  const void* w_nat;         // bf16 [E+H, 4H] natural-layout weight shadow
  const void* gates;         // bf16 [N, 4H] activated gates of step t
  const float* c_prev;
  const float* c_cur;
  const float* d_out;        // nullable fp32 [N, H]
  const float* out_keep;     // nullable
  long long out_keep_ld;
  float inv_keep;
  float* dh_carry;
  float* dc_carry;
  void* d_gates;             // bf16 [N, 4H] out
  const int* lengths;        // nullable
  int t, N, E, H;
};
int lstm_bwd_step(cudaStream_t stream, const LstmBwdArgs& a);

// lstm_seq.cu: whole-sequence persistent kernels (one launch for all steps) --------------------
struct LstmSeqFwdArgs {
  const void* X;         // bf16 [steps, N, E]
  void* Hs;              // bf16 [steps+1, N, H]; slot 0 = initial state (zeroed by the caller)
  float* Cs;             // fp32 [steps+1, N, H]; slot 0 zeroed by the caller
  void* G;               // bf16 [steps, N, 4H] activated gates (for BPTT)
  void* out;             // bf16 [T, N, H] emitted outputs of the caption steps (nullable)
  const void* w_t_perm;  // bf16 [4H gate-interleaved, E+H]
  const void* w_t_perm32 = nullptr;  // the same, gate-interleaved per 32 hidden units (16-CTA-per-row-tile form; nullable)
  const float* bias;
  const int* lengths;    // nullable
  const float* out_keep; // nullable [N, T, H]
  float inv_keep;
  int* flags;            // lstm_seq_flag_count(N, steps) ints of scratch
  int pre, T, steps, N, E, H;
};
struct LstmSeqBwdArgs {
  const void* w_nat;     // bf16 [E+H, 4H]
  const void* G;         // bf16 [steps, N, 4H]
  const float* Cs;       // fp32 [steps+1, N, H]
  const float* d_out;    // fp32 [T, N, H] (nullable)
  const float* out_keep; // nullable
  float inv_keep;
  float* dh_carry;
  float* dc_carry;
  void* dG;              // bf16 [steps, N, 4H]; slot steps-1 already filled (lstm_bwd_step with d_gates_next = nullptr)
  const int* lengths;
  int* flags;
  int pre, T, steps, N, E, H;
};
bool lstm_seq_applicable(int N, int H, int steps);
int lstm_seq_flag_count(int N, int steps);
int lstm_seq_units(int N, int H);  // hidden units per CTA the second form uses for N rows: 32 or 64
int lstm_fwd_seq(cudaStream_t stream, const LstmSeqFwdArgs& a);
int lstm_bwd_seq(cudaStream_t stream, const LstmSeqBwdArgs& a);
int lstm_fwd_seq2(cudaStream_t stream, const LstmSeqFwdArgs& a);  // lstm_seq2.cu: TMA-store hand-over of the recurrent operand
int lstm_bwd_seq2(cudaStream_t stream, const LstmSeqBwdArgs& a);

// vgg_bwd.cu -------------------------------------------------------------------------------
int conv3x3_wgrad(cudaStream_t s, const void* x, const void* dy, float* dw, int B, int hw, int cin, int cout,
                  const char* tag = "conv_wgrad");
// relu_src (nullable): the forward activation whose gradient dx is ([B, hw, hw, cin] bf16, a ReLU output): dx is then
// masked where relu_src <= 0 and dbias[cin] += per-channel sums of the masked dx, all inside the GEMM epilogue
int conv3x3_dgrad(cudaStream_t s, const void* dy, const void* wt_d, void* dx, int B, int hw, int cin, int cout,
                  const char* tag = "conv_dgrad", const void* relu_src = nullptr, float* dbias = nullptr);
int dgrad_shadow(cudaStream_t s, const float* w_hwio, void* wt_d, int cin, int cout);
int relu_pool_bwd(cudaStream_t s, const void* dA, const void* out, void* dY, int B, int hw, int C, bool pooled, float* db);

// comm.cu ----------------------------------------------------------------------------------
int comm_unique_id(void* id128);

// elementwise.cu ---------------------------------------------------------------------------
int cast_f32_bf16(cudaStream_t s, const float* src, void* dst, long long rows, int cols, long long ld_src,
                  long long ld_dst);
int bf16_to_f32(cudaStream_t s, const void* src, float* dst, long long rows, int cols, long long ld_src, long long ld_dst);
int transpose_cast(cudaStream_t s, const float* src, void* dst, int R, int C, long long ld_src, long long ld_dst,
                   int gate_h, int upt);
struct RefreshJob {
  const float* src;
  __nv_bfloat16* dst;
  int rows, cols;       // of src
  int ld_src, ld_dst;
  int kind;             // 0: cast, 1: transpose (+ LSTM gate interleave when gate_h > 0), 2: conv dgrad filter (tap-reversed, rows = 9*Cin taps x channels, cols = Cout; gate_h = Cin)
  int gate_h, upt;
  int vec4, blk0;       // filled by refresh_multi
};
struct RefreshJobs {
  RefreshJob j[32];
  int n = 0;
  void cast(const float* src, void* dst, int rows, int cols, int ld_src, int ld_dst) {
    j[n++] = RefreshJob{src, (__nv_bfloat16*)dst, rows, cols, ld_src, ld_dst, 0, 0, 0, 0, 0};
  }
  void transpose(const float* src, void* dst, int rows, int cols, int ld_src, int ld_dst, int gate_h, int upt) {
    j[n++] = RefreshJob{src, (__nv_bfloat16*)dst, rows, cols, ld_src, ld_dst, 1, gate_h, upt, 0, 0};
  }
  // dst[ci, (8 - tap) * Cout + co] = bf16(W[tap, ci, co]): the tap-reversed [Cin, 9*Cout] filter of the input-gradient convolution
  void dgrad_filter(const float* w_hwio, void* dst, int cin, int cout) {
    j[n++] = RefreshJob{w_hwio, (__nv_bfloat16*)dst, 9 * cin, cout, cout, 9 * cout, 2, cin, 0, 0, 0};
  }
  bool full() const { return n >= 32; }
};
int refresh_multi(cudaStream_t s, RefreshJobs& jobs);
int keep_mask(cudaStream_t s, float* mask, long long n, float keep_prob, unsigned long long seed, unsigned long long offset);
int tile_cast(cudaStream_t s, const float* src, void* d0, void* d1, int B, int C, int E);
int tile_reduce(cudaStream_t s, const float* a, const float* b2, float* dst_f, void* dst_h, int B, int C, int E);
int embed_gather(cudaStream_t s, const void* table, const int* tok, void* X, const float* keep_mask, float inv_keep,
                 int N, int T, int E, int V);
int embed_scatter(cudaStream_t s, const float* dX, const int* tok, float* gtable, const float* keep_mask, float inv_keep,
                  float* normsq, int N, int T, int E, int V);
int heads_mix(cudaStream_t s, const float* heads, long long ld, int zp, int prior, const float* c_v, int K, const int* pick,
              const float* c_means, float* mu, float* sd, float* cm, int N, int Z);
int heads_mix_bwd(cudaStream_t s, const float* dmu, const float* dsd, const float* heads, long long ld, int zp, int prior,
                  const float* c_v, int K, const int* pick, const float* sd, void* dheads, int N, int Z);
int gmm_pick_clusters(cudaStream_t s, const float* c_v, int K, unsigned long long seed, unsigned long long offset, int* pick,
                      int N);
int kl_rows(cudaStream_t s, const float* mu, const float* sd, const float* cm, int prior, float* kl_row, float* dkl_dmu,
            float* dkl_dsd, float* kl_sum, int N, int Z);
int sample_z(cudaStream_t s, const float* mu, const float* sd, const float* eps, unsigned long long seed,
             unsigned long long offset, void* z, float* z_f32, int S, long long NZ);
int dz_reduce(cudaStream_t s, const float* dz, const float* eps, unsigned long long seed, unsigned long long offset,
              const float* sd, const float* dkl_dmu, const float* dkl_dsd, float kl_scale, void* dheads, long long ld,
              int zp, float* dmu_out, float* dsd_out, int S, int N, int Z);
// rows [row0, row0 + n_rows) of the time-major logits matrix (n_rows < 0: all rows from row0)
int ce_rows(cudaStream_t s, void* logits, long long ld, const int* lbl, int N, int T, int V, float* sums, float* ce_out,
            const float* count, float loss_scale, int write_grad, int row0 = 0, int n_rows = -1);
int count_mask(cudaStream_t s, const int* lbl, long long n, float* count);
int colsum_bf16(cudaStream_t s, const void* x, long long rows, int cols, long long ld, float* out);
int sumsq(cudaStream_t s, const float* g, long long n, float* out);
int adam_step(cudaStream_t s, float* p, const float* g, float* m, float* v, long long n, const float* normsq_parts,
              int n_parts, float clip, float gscale, float lr_t, float b1, float b2, float eps, float* norm_out, float weight_decay = 0.f,
              int kind = 0 /* VC_OPT_*: 0 TF-Adam, 1 SGD (lr_t = decayed rate), 2 Momentum (b1 = momentum, m = accumulator) */);
int logits_to_ref(cudaStream_t s, const void* src, long long ld, float* dst, int N, int T, int V);

}  // namespace vc
