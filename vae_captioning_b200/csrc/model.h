// The CVAE captioning model state behind vc_handle: parameter store (TF names), bf16 weight
// shadows, time-major activation workspace, and the train / eval step orchestration.
#pragma once
#include <initializer_list>
#include <map>
#include <string>
#include <vector>
#include "../../include/vaecap.h"
#include "ops.h"

namespace vc {

struct ParamInfo {
  std::string name;
  int ndim;
  int64_t shape[4];
  int64_t count;
  int64_t offset;  // element offset into the flat fp32 buffers
  int region;      // 0: clipped Adam (non-CNN), 1: frozen (created but never updated), 2: CNN
  bool trainable;
  bool hidden = false;     // packed block that backs several variables (not listed by vc_param_info)
  int parent = -1;         // >= 0: this variable is a view into params[parent] ...
  int64_t parent_off = 0;  // ... starting at this element offset ...
  int64_t ld = 0;          // ... with this row pitch (0 = contiguous)
};

struct LstmNet {
  int p_kernel = -1, p_bias = -1;
  int E = 0, H = 0, pre = 0, steps = 0;
  void* w_t_perm = nullptr;  // bf16 [4H interleaved, E+H]
  void* w_t_perm32 = nullptr;  // interleaved per 32 units: lstm_seq2's 16-CTA-per-row-tile form (batches of <= 1152 rows)
  void* w_nat = nullptr;     // bf16 [E+H, 4H]
  void* X = nullptr;         // bf16 [steps, N, E]
  void* Hs = nullptr;        // bf16 [steps+1, N, H]
  float* Cs = nullptr;       // fp32 [steps+1, N, H]
  void* G = nullptr;         // bf16 [steps, N, 4H] activated gates
  void* dG = nullptr;        // bf16 [steps, N, 4H]
  float* dX = nullptr;       // fp32 [steps, N, E]
  float* dh_carry = nullptr;
  float* dc_carry = nullptr;
};

struct VggLayer {
  int cin = 0, cout = 0, hw = 0, kpad = 0;
  bool pool = false;
  int p_w = -1, p_b = -1;
  void* wt = nullptr;      // bf16 [cout, kpad] (tap-major, then cin)
  void* wt_d = nullptr;    // bf16 [cin, 9*cout] tap-reversed (dgrad B operand; fine_tune only)
  void* out = nullptr;     // bf16 NHWC [B, hw, hw, cout] (post-ReLU, un-pooled)
  void* pooled = nullptr;  // bf16 NHWC [B, hw/2, hw/2, cout] when pool
};

struct DecodeWs;  // decode.cu
struct Comm;      // comm.cu

struct StepInputs {
  const float* feats;     // device fp32 [B, F] (or images when fine_tune; uint8 pixels when feats_u8)
  bool feats_u8 = false;
  const int32_t* cap_lbl; // device [N, T]
  const int32_t* cap_in;
  const int32_t* len;
  const float* c_v;       // device fp32 [N, K] or null
  int B, T;
  int64_t global_step;
  vc_rng rng;
};

class Model {
 public:
  vc_config cfg;
  int device = 0;
  int maxN = 0, maxT = 0;
  std::vector<ParamInfo> params;
  std::vector<int> visible;  // indices of the variables the Saver surface lists (creation order)
  std::map<std::string, int> index;
  int64_t n_adam = 0, n_dense = 0, n_frozen = 0, n_cnn = 0, n_total = 0;  // element counts (padded)
  float *Pf = nullptr, *Gf = nullptr, *Mf = nullptr, *Vf = nullptr;  // flat fp32: params, grads, Adam m, v
  float* g_tail = nullptr;  // 4 floats after the Adam region of G: [|enc slices|^2, |dec slices|^2, |dense|^2, -]
  int64_t adam_t = 0;       // TF global_step of the optimiser (1-based after the first update)
  bool shadows_dirty = true, decode_shadows_dirty = true;
  bool have_forward = false, logits_intact = false;
  bool vocab_bwd_done = false;  // the row-chunked forward already produced dOut and the vocabulary bias gradient
  int lastN = 0, lastT = 0;
  float last_ann = 1.f;

  // --- shadows (bf16 unless noted)
  void *imf_wt = nullptr, *cv_wt = nullptr, *cv_nat = nullptr, *heads_nat = nullptr, *z_wt = nullptr, *z_nat = nullptr,
       *wo_t = nullptr, *wo_nat = nullptr, *enc_emb_h = nullptr, *dec_emb_h = nullptr;
  // The posterior heads (encoder.py:59-107) are packed: kernel block [He, heads_cols] and bias block [heads_cols],
  // head k's mean at columns [2k*ZP, 2k*ZP+Z), its log-std at [(2k+1)*ZP, ...). The TF variables
  // encoder/{dense,dense_1,gmm_ll_k/...,ag_ll_k/...} are strided views into the two blocks.
  int p_heads_w = -1, p_heads_b = -1, heads_nh = 0;
  float* heads_bias = nullptr;  // = pp(p_heads_b)
  float* c_means = nullptr;     // fp32 [K, Z] cluster means (utils/vae_utils.py:20-30), AG prior
  int32_t* gmm_pick = nullptr;  // [N] cluster drawn per row (encoder.py:72) when not given explicitly
  const int32_t* last_pick = nullptr;
  int ZP = 0, VP = 0, KP = 0, heads_cols = 0;

  LstmNet enc, dec;

  // --- workspace
  void *feats_h = nullptr, *cv_h = nullptr, *Out = nullptr, *logits = nullptr, *z = nullptr, *dheads = nullptr,
       *dzdec_h = nullptr, *dimf_h = nullptr, *dcv_h = nullptr;
  float *imf_f = nullptr, *cv_f = nullptr, *heads_f = nullptr, *mu = nullptr, *sd = nullptr, *kl_row = nullptr,
        *dkl_dmu = nullptr, *dkl_dsd = nullptr, *zdec_f = nullptr, *dOut = nullptr, *dz = nullptr, *ce_row = nullptr,
        *scal = nullptr, *cm = nullptr, *dmu_t = nullptr, *dsd_t = nullptr;
  // staging for host-pointer entry points
  float *st_feats = nullptr, *st_cv = nullptr;
  int32_t *st_lbl = nullptr, *st_in = nullptr, *st_len = nullptr;
  float* host_scal = nullptr;  // pinned
  // deferred results (vc_step_result_queue / vc_step_result): ring of two pinned 16-float slots + what fetch() needs besides
  struct PendingResult {
    cudaEvent_t ev = nullptr;
    int n = 0;
    float ann = 0.f;
  } pending[2];
  int pending_head = 0, pending_count = 0;
  int result_queue(cudaStream_t s);
  int result_pop(vc_step_out* out);
  void fill_out(vc_step_out* out, const float* h, int n, float ann) const;
  float *emb_keep_buf = nullptr, *out_keep_buf = nullptr;  // Philox keep masks of the decoder dropouts (allocated on first use)
  // Double-buffered feed (vc_stage_batch / vc_train_step_staged): the next step's host buffers are copied into one slot
  // on a copy stream while the current step computes from the other. Slots are allocated on first use.
  struct StageSlot {
    void* px = nullptr;  // features fp32 [B, F], or images (fp32 / uint8) [B, 224, 224, 3]
    float* fc2 = nullptr;  // frozen extractor: fc2 features of the staged images, computed on the copy stream
    float* cv = nullptr;
    int32_t *lbl = nullptr, *in = nullptr, *len = nullptr;
    cudaEvent_t ready = nullptr, consumed = nullptr;
    cudaEvent_t img_ready = nullptr, px_free = nullptr;  // frozen extractor: image copy done / VGG16 forward has read px
    int B = 0, T = 0, kind = -1;
    bool has_cv = false, filled = false, in_use = false, px_used = false;
  };
  StageSlot slots[2];
  cudaStream_t h2d_stream = nullptr;  // image copies of the frozen-extractor feed (so that they overlap the previous forward)
  int stage_slot(int slot, const void* px_host, int kind, const int32_t* lbl, const int32_t* inp, const int32_t* len,
                 const float* cv, int B, int T, cudaStream_t copy_stream);
  int step_from_slot(int slot, int64_t gs, const vc_rng* rng, cudaStream_t s, bool apply_update = true);
  int* seq_flags = nullptr;    // inter-CTA step counters of the persistent LSTM kernels

  // --- VGG16 (vgg.cu)
  std::vector<VggLayer> vgg;
  void* vgg_padded = nullptr;  // bf16 [B, 226, 226, 8]: mean-subtracted, zero-bordered input of the window form of conv1_1
  bool conv1_window = false;  // VC_CONV1=window
  void *vgg_im2col = nullptr, *fc1_w = nullptr, *fc2_w = nullptr, *fc1_h = nullptr;
  float *fc_acc = nullptr, *fc2_f = nullptr, *st_images = nullptr, *conv1_bias2 = nullptr;
  bool vgg_shadows_dirty = true, vgg_keep = false, vgg_have_unpooled = false;
  int vgg_last_B = 0;
  // fine-tune backward (vgg_bwd.cu)
  void *vgg_bwd_a = nullptr, *vgg_bwd_b = nullptr, *imf_nat = nullptr, *dfc2_pre = nullptr, *dfc1_pre = nullptr;
  float *dfeats_f = nullptr, *conv1_wg = nullptr;
  unsigned long long vgg_drop_seed = 0, vgg_drop_step = 0;  // Philox stream of the fc dropout masks (explicit masks win)
  int vgg_bwd_init();
  int vgg_refresh_bwd_shadows(cudaStream_t s);
  int vgg_backward(const float* dfeats, int B, cudaStream_t s);
  int vgg_init();
  int vgg_refresh_shadows(cudaStream_t s);
  int vgg_conv_layer(int l, const void* in, int B, bool fuse_pool, cudaStream_t s);
  int vgg_forward(const float* images, float* fc2_out, int B, bool keep_unpooled, const float* fc_keep, cudaStream_t s,
                  bool images_u8 = false);
  int vgg_activation(const char* layer, float* dst_host);

  // --- generation (decode.cu)
  DecodeWs* dws = nullptr;
  int decode_reserve(int B, int beam);
  void decode_release();
  int decode_stage(const float* feats_host, const float* c_v_host, int B, int beam, const float** feats_dev,
                   const float** c_v_dev, cudaStream_t s);
  int decode_begin(const float* feats_dev, const float* c_v_dev, int B, const vc_rng* rng, cudaStream_t s);
  int decode_advance(int M, cudaStream_t s);
  int decode_greedy(const float* feats_dev, const float* c_v_dev, int B, int max_len, int mode, const vc_rng* rng, int bos,
                    int eos, int32_t* out_tokens_host, int32_t* out_len_host, cudaStream_t s);
  int decode_beam(const float* feats_dev, const float* c_v_dev, int B, int beam, int max_len, float len_norm,
                  const vc_rng* rng, int bos, int eos, int32_t* out_tokens_host, int32_t* out_len_host,
                  float* out_score_host, int32_t* out_n_host, cudaStream_t s);
  int decode_open(const float* feats_dev, const float* c_v_dev, int B, const vc_rng* rng, cudaStream_t s);
  int decode_step(const int32_t* tok_host, int M, float* probs_host, cudaStream_t s);
  int decode_state(float* c_host, float* h_host, const float* c_in, const float* h_in, cudaStream_t s);

  // --- side stream of the backward pass: weight / bias gradients nothing downstream waits for run beside the BPTT kernels
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  int side_sm_cap = 0;
  bool side_used = false;
  bool side_begin(cudaStream_t s);
  int side_join(cudaStream_t s);

  // --- data-parallel gradient all-reduce (comm.cu); null = single device
  Comm* comm = nullptr;
  int comm_init(const void* id128, int rank, int world);
  void comm_release();
  int comm_world() const;
  bool comm_bf16() const;
  int grad_ready(const int64_t (*ranges)[2], int n_ranges, cudaStream_t s);
  int grad_ready_params(std::initializer_list<int> ids, cudaStream_t s);
  int comm_reduce(const int64_t (*ranges)[2], int n_ranges, cudaStream_t s);
  int comm_set_mode(int mode);
  int comm_join(cudaStream_t s);
  int comm_allreduce_all(cudaStream_t s);
  int comm_stats(float* ms, long long* bytes, int* calls);
  float step_scale() const { return 1.f / (float)comm_world(); }  // towers are averaged

  std::vector<void*> allocs;

  ~Model();
  int init(const vc_config& c, int dev);
  int param_index(const char* name) const;
  int param_set(const char* name, const float* src);
  int param_get(const char* name, float* dst);
  int grad_get(const char* name, float* dst);
  int copy_var(int i, float* base, float* host_dst, const float* host_src);
  int set_cluster_means(const float* src_host);

  int refresh_shadows(cudaStream_t s);
  int refresh_decode_shadows(cudaStream_t s);
  int forward(StepInputs& in, bool write_grad, cudaStream_t s);  // may fill in.rng's dropout masks (Philox) for backward()
  int backward(const StepInputs& in, cudaStream_t s);
  int apply(float grad_scale, cudaStream_t s);
  int fetch(vc_step_out* out, cudaStream_t s);
  int stage_inputs(const float* feats, const int32_t* lbl, const int32_t* inp, const int32_t* len, const float* cv,
                   int B, int T, StepInputs* out, cudaStream_t s, bool feats_u8 = false);
  int forward_debug(float* logits_host, float* mu_host, float* std_host, float* z_host, float* kl_host, float* ce_host);

 private:
  template <class T>
  int dalloc(T** p, size_t count, bool zero = true) {
    void* q = nullptr;
    if (count == 0) count = 1;
    cudaError_t e = cudaMalloc(&q, count * sizeof(T));
    if (e != cudaSuccess)
      return set_error(VC_E_NOMEM, "cudaMalloc(%zu bytes) failed: %s", count * sizeof(T), cudaGetErrorString(e));
    if (zero) VC_CUDA(cudaMemset(q, 0, count * sizeof(T)));
    allocs.push_back(q);
    *p = reinterpret_cast<T*>(q);
    return VC_OK;
  }
  int add_param(const std::string& name, std::vector<int64_t> shape, int region, bool hidden = false);
  int add_view(const std::string& name, std::vector<int64_t> shape, int parent, int64_t parent_off, int64_t ld);
  int lstm_forward(LstmNet& L, int N, int T, const int32_t* len, void* out, const float* out_keep, cudaStream_t s);
  int lstm_backward(LstmNet& L, int N, int T, const int32_t* len, const float* d_out, const float* out_keep,
                    cudaStream_t s, bool defer_weight_grads = false);
  int lstm_weight_grads(LstmNet& L, int N, int T, cudaStream_t s);
  float* gp(int pi) { return Gf + params[pi].offset; }
  float* pp(int pi) { return Pf + params[pi].offset; }
  int pidx(const std::string& n) const {
    auto it = index.find(n);
    return it == index.end() ? -1 : it->second;
  }
};

}  // namespace vc
