// C-ABI surface of libvaecap.so (declared in include/vaecap.h). No exceptions cross this boundary.
#include <new>
#include "model.h"

using namespace vc;

struct vc_handle {
  Model m;
};

#define VC_GUARD_BEGIN try {
#define VC_GUARD_END                                                   \
  }                                                                    \
  catch (const std::bad_alloc&) {                                      \
    return set_error(VC_E_NOMEM, "host allocation failed");            \
  }                                                                    \
  catch (...) {                                                        \
    return set_error(VC_E_STATE, "unexpected C++ exception");          \
  }

extern "C" {

const char* vc_last_error(void) { return last_error().c_str(); }
int vc_abi_version(void) { return 1; }
unsigned long long vc_launch_count(void) { return launch_count(); }
int vc_profile_enable(int on) {
  prof_enable(on != 0);
  return VC_OK;
}
int vc_profile_collect(char* names, int names_cap, float* ms, int* counts, int cap) {
  return prof_collect(names, names_cap, ms, counts, cap);
}

int vc_create(const vc_config* cfg, int device, vc_handle** out) {
  VC_GUARD_BEGIN
  if (cfg == nullptr || out == nullptr) return set_error(VC_E_ARG, "vc_create: null argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return set_error(VC_E_CUDA, "no CUDA device: the hot path has no CPU fallback");
  if (device < 0 || device >= ndev) return set_error(VC_E_ARG, "device %d out of range (%d devices)", device, ndev);
  cudaDeviceProp prop;
  VC_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return set_error(VC_E_CUDA, "libvaecap is built for sm_100a only (device is sm_%d%d)", prop.major, prop.minor);
  vc_handle* h = new vc_handle();
  int st = h->m.init(*cfg, device);
  if (st != VC_OK) {
    delete h;
    return st;
  }
  *out = h;
  return VC_OK;
  VC_GUARD_END
}

int vc_destroy(vc_handle* h) {
  if (h == nullptr) return VC_OK;
  cudaSetDevice(h->m.device);
  cudaDeviceSynchronize();
  delete h;
  return VC_OK;
}

int vc_num_params(vc_handle* h) { return h ? (int)h->m.visible.size() : set_error(VC_E_ARG, "null handle"); }

int vc_param_info(vc_handle* h, int index, const char** name, int32_t* ndim, int64_t* shape, int32_t* trainable) {
  if (h == nullptr) return set_error(VC_E_ARG, "null handle");
  if (index < 0 || index >= (int)h->m.visible.size()) return set_error(VC_E_ARG, "parameter index %d out of range", index);
  const ParamInfo& p = h->m.params[h->m.visible[index]];
  if (name) *name = p.name.c_str();
  if (ndim) *ndim = p.ndim;
  if (shape) for (int i = 0; i < 4; ++i) shape[i] = p.shape[i];
  if (trainable) *trainable = p.trainable ? 1 : 0;
  return VC_OK;
}

int vc_param_get(vc_handle* h, const char* name, float* dst) {
  if (!h || !name || !dst) return set_error(VC_E_ARG, "vc_param_get: null argument");
  cudaSetDevice(h->m.device);
  return h->m.param_get(name, dst);
}
int vc_param_set(vc_handle* h, const char* name, const float* src) {
  if (!h || !name || !src) return set_error(VC_E_ARG, "vc_param_set: null argument");
  cudaSetDevice(h->m.device);
  return h->m.param_set(name, src);
}
int vc_grad_get(vc_handle* h, const char* name, float* dst) {
  if (!h || !name || !dst) return set_error(VC_E_ARG, "vc_grad_get: null argument");
  cudaSetDevice(h->m.device);
  return h->m.grad_get(name, dst);
}

int vc_set_cluster_means(vc_handle* h, const float* src) {
  if (!h || !src) return set_error(VC_E_ARG, "vc_set_cluster_means: null argument");
  cudaSetDevice(h->m.device);
  return h->m.set_cluster_means(src);
}

static StepInputs make_inputs(const float* feats, const int32_t* lbl, const int32_t* inp, const int32_t* len,
                              const float* cv, int B, int T, int64_t gs, const vc_rng* rng) {
  StepInputs in{};
  in.feats = feats; in.cap_lbl = lbl; in.cap_in = inp; in.len = len; in.c_v = cv;
  in.B = B; in.T = T; in.global_step = gs;
  if (rng) in.rng = *rng;
  return in;
}

int vc_forward_backward_dev(vc_handle* h, const float* feats, const int32_t* lbl, const int32_t* inp, const int32_t* len,
                            const float* cv, int B, int T, int64_t gs, const vc_rng* rng, void* stream) {
  VC_GUARD_BEGIN
  if (!h || !feats || !lbl || !inp || !len) return set_error(VC_E_ARG, "vc_forward_backward_dev: null argument");
  cudaSetDevice(h->m.device);
  cudaStream_t s = (cudaStream_t)stream;
  StepInputs in = make_inputs(feats, lbl, inp, len, cv, B, T, gs, rng);
  VC_TRY(h->m.forward(in, true, s));
  return h->m.backward(in, s);
  VC_GUARD_END
}

int vc_grad_buffer(vc_handle* h, float** dev_ptr, int64_t* count) {
  if (!h || !dev_ptr || !count) return set_error(VC_E_ARG, "vc_grad_buffer: null argument");
  *dev_ptr = h->m.Gf;
  // gradients + the 64-float tail carrying the embedding-slice squared norms (+ frozen gap and cnn/ region when fine-tuning)
  *count = h->m.cfg.fine_tune ? h->m.n_total : h->m.n_adam + 64;
  return VC_OK;
}

int vc_apply_gradients(vc_handle* h, float grad_scale, vc_step_out* out, void* stream) {
  VC_GUARD_BEGIN
  if (!h) return set_error(VC_E_ARG, "null handle");
  cudaSetDevice(h->m.device);
  cudaStream_t s = (cudaStream_t)stream;
  VC_TRY(h->m.apply(grad_scale, s));
  return h->m.fetch(out, s);
  VC_GUARD_END
}

int vc_train_step_dev(vc_handle* h, const float* feats, const int32_t* lbl, const int32_t* inp, const int32_t* len,
                      const float* cv, int B, int T, int64_t gs, const vc_rng* rng, vc_step_out* out, void* stream) {
  VC_TRY(vc_forward_backward_dev(h, feats, lbl, inp, len, cv, B, T, gs, rng, stream));
  return vc_apply_gradients(h, h->m.step_scale(), out, stream);
}

int vc_comm_unique_id(void* id128) {
  VC_GUARD_BEGIN
  if (!id128) return set_error(VC_E_ARG, "vc_comm_unique_id: null argument");
  return comm_unique_id(id128);
  VC_GUARD_END
}
int vc_comm_init(vc_handle* h, const void* id128, int rank, int world) {
  VC_GUARD_BEGIN
  if (!h || !id128) return set_error(VC_E_ARG, "vc_comm_init: null argument");
  return h->m.comm_init(id128, rank, world);
  VC_GUARD_END
}
int vc_allreduce_gradients(vc_handle* h, void* stream) {
  VC_GUARD_BEGIN
  if (!h) return set_error(VC_E_ARG, "null handle");
  cudaSetDevice(h->m.device);
  return h->m.comm_allreduce_all((cudaStream_t)stream);
  VC_GUARD_END
}
int vc_comm_set_mode(vc_handle* h, int mode) {
  VC_GUARD_BEGIN
  if (!h) return set_error(VC_E_ARG, "null handle");
  return h->m.comm_set_mode(mode);
  VC_GUARD_END
}
int vc_comm_stats(vc_handle* h, float* ms, long long* bytes, int* calls) {
  VC_GUARD_BEGIN
  if (!h) return set_error(VC_E_ARG, "null handle");
  cudaSetDevice(h->m.device);
  return h->m.comm_stats(ms, bytes, calls);
  VC_GUARD_END
}

int vc_train_step(vc_handle* h, const float* feats, const int32_t* lbl, const int32_t* inp, const int32_t* len,
                  const float* cv, int B, int T, int64_t gs, const vc_rng* rng, vc_step_out* out, void* stream) {
  VC_GUARD_BEGIN
  if (!h || !feats || !lbl || !inp || !len) return set_error(VC_E_ARG, "vc_train_step: null argument");
  cudaSetDevice(h->m.device);
  cudaStream_t s = (cudaStream_t)stream;
  StepInputs in{};
  VC_TRY(h->m.stage_inputs(feats, lbl, inp, len, cv, B, T, &in, s));
  in.global_step = gs;
  if (rng) in.rng = *rng;
  VC_TRY(h->m.forward(in, true, s));
  VC_TRY(h->m.backward(in, s));
  VC_TRY(h->m.apply(h->m.step_scale(), s));
  return h->m.fetch(out, s);
  VC_GUARD_END
}

static int train_step_images_impl(vc_handle* h, const void* images, bool u8, const int32_t* lbl, const int32_t* inp,
                                  const int32_t* len, const float* cv, int B, int T, int64_t gs, const vc_rng* rng,
                                  vc_step_out* out, void* stream) {
  VC_GUARD_BEGIN
  if (!h || !images || !lbl || !inp || !len) return set_error(VC_E_ARG, "vc_train_step_images: null argument");
  Model& m = h->m;
  cudaSetDevice(m.device);
  cudaStream_t s = (cudaStream_t)stream;
  if (m.vgg.empty()) return set_error(VC_E_STATE, "this handle was created without the CNN (with_cnn = 0)");
  if (m.cfg.fine_tune) {  // the image batch is the feed of the trainable graph (main.py:46-48)
    StepInputs in{};
    VC_TRY(m.stage_inputs(reinterpret_cast<const float*>(images), lbl, inp, len, cv, B, T, &in, s, u8));
    in.global_step = gs;
    if (rng) in.rng = *rng;
    VC_TRY(m.forward(in, true, s));
    VC_TRY(m.backward(in, s));
    VC_TRY(m.apply(m.step_scale(), s));
    return m.fetch(out, s);
  }
  if (B < 1 || B > m.cfg.max_batch) return set_error(VC_E_SHAPE, "batch %d exceeds max_batch %d", B, m.cfg.max_batch);
  VC_CUDA(cudaMemcpyAsync(m.st_images, images, (size_t)B * 224 * 224 * 3 * (u8 ? 1 : sizeof(float)), cudaMemcpyHostToDevice, s));
  VC_TRY(m.vgg_forward(m.st_images, nullptr, B, false, nullptr, s, u8));
  StepInputs in{};
  // captions / lengths / c_v are staged as usual; the features come from the on-device forward
  VC_TRY(m.stage_inputs(nullptr, lbl, inp, len, cv, B, T, &in, s));
  in.feats = m.fc2_f;
  in.global_step = gs;
  if (rng) in.rng = *rng;
  VC_TRY(m.forward(in, true, s));
  VC_TRY(m.backward(in, s));
  VC_TRY(m.apply(m.step_scale(), s));
  return m.fetch(out, s);
  VC_GUARD_END
}

int vc_train_step_images(vc_handle* h, const float* images, const int32_t* lbl, const int32_t* inp, const int32_t* len,
                         const float* cv, int B, int T, int64_t gs, const vc_rng* rng, vc_step_out* out, void* stream) {
  return train_step_images_impl(h, images, false, lbl, inp, len, cv, B, T, gs, rng, out, stream);
}
int vc_train_step_images_u8(vc_handle* h, const uint8_t* images, const int32_t* lbl, const int32_t* inp, const int32_t* len,
                            const float* cv, int B, int T, int64_t gs, const vc_rng* rng, vc_step_out* out, void* stream) {
  return train_step_images_impl(h, images, true, lbl, inp, len, cv, B, T, gs, rng, out, stream);
}

int vc_stage_batch(vc_handle* h, int slot, const void* px, int kind, const int32_t* lbl, const int32_t* inp, const int32_t* len,
                   const float* cv, int B, int T, void* copy_stream) {
  VC_GUARD_BEGIN
  if (!h || !px || !lbl || !inp || !len) return set_error(VC_E_ARG, "vc_stage_batch: null argument");
  cudaSetDevice(h->m.device);
  return h->m.stage_slot(slot, px, kind, lbl, inp, len, cv, B, T, (cudaStream_t)copy_stream);
  VC_GUARD_END
}

int vc_train_step_staged(vc_handle* h, int slot, int64_t gs, const vc_rng* rng, vc_step_out* out, void* stream) {
  VC_GUARD_BEGIN
  if (!h) return set_error(VC_E_ARG, "null handle");
  cudaSetDevice(h->m.device);
  cudaStream_t s = (cudaStream_t)stream;
  VC_TRY(h->m.step_from_slot(slot, gs, rng, s));
  return h->m.fetch(out, s);
  VC_GUARD_END
}

int vc_step_result_queue(vc_handle* h, void* stream) {
  VC_GUARD_BEGIN
  if (!h) return set_error(VC_E_ARG, "null handle");
  cudaSetDevice(h->m.device);
  return h->m.result_queue((cudaStream_t)stream);
  VC_GUARD_END
}

int vc_step_result(vc_handle* h, vc_step_out* out) {
  VC_GUARD_BEGIN
  if (!h) return set_error(VC_E_ARG, "null handle");
  cudaSetDevice(h->m.device);
  return h->m.result_pop(out);
  VC_GUARD_END
}

int vc_forward_backward_staged(vc_handle* h, int slot, int64_t gs, const vc_rng* rng, void* stream) {
  VC_GUARD_BEGIN
  if (!h) return set_error(VC_E_ARG, "null handle");
  cudaSetDevice(h->m.device);
  return h->m.step_from_slot(slot, gs, rng, (cudaStream_t)stream, false);
  VC_GUARD_END
}

int vc_eval_step(vc_handle* h, const float* feats, const int32_t* lbl, const int32_t* inp, const int32_t* len,
                 const float* cv, int B, int T, const vc_rng* rng, vc_step_out* out, void* stream) {
  VC_GUARD_BEGIN
  if (!h || !feats || !lbl || !inp || !len) return set_error(VC_E_ARG, "vc_eval_step: null argument");
  cudaSetDevice(h->m.device);
  cudaStream_t s = (cudaStream_t)stream;
  StepInputs in{};
  VC_TRY(h->m.stage_inputs(feats, lbl, inp, len, cv, B, T, &in, s));
  in.global_step = h->m.adam_t;
  if (rng) in.rng = *rng;
  VC_TRY(h->m.forward(in, false, s));
  return h->m.fetch(out, s);
  VC_GUARD_END
}

int vc_forward_debug(vc_handle* h, float* logits, float* mu, float* sd, float* z, float* kl_rows, float* ce_rows) {
  VC_GUARD_BEGIN
  if (!h) return set_error(VC_E_ARG, "null handle");
  cudaSetDevice(h->m.device);
  return h->m.forward_debug(logits, mu, sd, z, kl_rows, ce_rows);
  VC_GUARD_END
}

int vc_vgg_keep_activations(vc_handle* h, int on) {
  if (!h) return set_error(VC_E_ARG, "null handle");
  h->m.vgg_keep = on != 0;
  return VC_OK;
}

int vc_vgg_forward_dev(vc_handle* h, const float* images, float* fc2, int B, void* stream) {
  VC_GUARD_BEGIN
  if (!h || !images) return set_error(VC_E_ARG, "vc_vgg_forward_dev: null argument");
  cudaSetDevice(h->m.device);
  return h->m.vgg_forward(images, fc2, B, h->m.vgg_keep, nullptr, (cudaStream_t)stream);
  VC_GUARD_END
}

static int vgg_forward_host(vc_handle* h, const void* images_host, bool u8, float* fc2_host, int B, void* stream) {
  VC_GUARD_BEGIN
  if (!h || !images_host || !fc2_host) return set_error(VC_E_ARG, "vc_vgg_forward: null argument");
  Model& m = h->m;
  cudaSetDevice(m.device);
  cudaStream_t s = (cudaStream_t)stream;
  if (m.vgg.empty()) return set_error(VC_E_STATE, "this handle was created without the CNN (with_cnn = 0)");
  if (B < 1 || B > m.cfg.max_batch) return set_error(VC_E_SHAPE, "vc_vgg_forward: batch %d exceeds max_batch %d", B, m.cfg.max_batch);
  VC_CUDA(cudaMemcpyAsync(m.st_images, images_host, (size_t)B * 224 * 224 * 3 * (u8 ? 1 : sizeof(float)), cudaMemcpyHostToDevice, s));
  VC_TRY(m.vgg_forward(m.st_images, nullptr, B, m.vgg_keep, nullptr, s, u8));
  VC_CUDA(cudaMemcpyAsync(fc2_host, m.fc2_f, (size_t)B * 4096 * sizeof(float), cudaMemcpyDeviceToHost, s));
  VC_CUDA(cudaStreamSynchronize(s));
  return VC_OK;
  VC_GUARD_END
}
int vc_vgg_forward(vc_handle* h, const float* images_host, float* fc2_host, int B, void* stream) {
  return vgg_forward_host(h, images_host, false, fc2_host, B, stream);
}
int vc_vgg_forward_u8(vc_handle* h, const uint8_t* images_host, float* fc2_host, int B, void* stream) {
  return vgg_forward_host(h, images_host, true, fc2_host, B, stream);
}

int vc_decode_greedy(vc_handle* h, const float* feats_host, const float* c_v_host, int B, int max_len, int mode,
                     const vc_rng* rng, int bos, int eos, int32_t* out_tokens_host, int32_t* out_len_host, void* stream) {
  VC_GUARD_BEGIN
  if (!h || !feats_host || !out_tokens_host || !out_len_host) return set_error(VC_E_ARG, "vc_decode_greedy: null argument");
  if (mode != 0 && mode != 1) return set_error(VC_E_ARG, "vc_decode_greedy: mode must be 0 (greedy) or 1 (sample)");
  Model& m = h->m;
  cudaSetDevice(m.device);
  cudaStream_t s = (cudaStream_t)stream;
  const float *fd = nullptr, *cd = nullptr;
  VC_TRY(m.decode_stage(feats_host, c_v_host, B, 1, &fd, &cd, s));
  return m.decode_greedy(fd, cd, B, max_len, mode, rng, bos, eos, out_tokens_host, out_len_host, s);
  VC_GUARD_END
}

int vc_decode_beam(vc_handle* h, const float* feats_host, const float* c_v_host, int B, int beam, int max_len, float len_norm,
                   const vc_rng* rng, int bos, int eos, int32_t* out_tokens_host, int32_t* out_len_host,
                   float* out_score_host, int32_t* out_n_host, void* stream) {
  VC_GUARD_BEGIN
  if (!h || !feats_host || !out_tokens_host || !out_len_host || !out_score_host || !out_n_host)
    return set_error(VC_E_ARG, "vc_decode_beam: null argument");
  Model& m = h->m;
  cudaSetDevice(m.device);
  cudaStream_t s = (cudaStream_t)stream;
  const float *fd = nullptr, *cd = nullptr;
  VC_TRY(m.decode_stage(feats_host, c_v_host, B, beam, &fd, &cd, s));
  return m.decode_beam(fd, cd, B, beam, max_len, len_norm, rng, bos, eos, out_tokens_host, out_len_host, out_score_host,
                       out_n_host, s);
  VC_GUARD_END
}

int vc_decode_begin(vc_handle* h, const float* feats_host, const float* c_v_host, int B, const vc_rng* rng, void* stream) {
  VC_GUARD_BEGIN
  if (!h || !feats_host) return set_error(VC_E_ARG, "vc_decode_begin: null argument");
  Model& m = h->m;
  cudaSetDevice(m.device);
  cudaStream_t s = (cudaStream_t)stream;
  const float *fd = nullptr, *cd = nullptr;
  VC_TRY(m.decode_stage(feats_host, c_v_host, B, 1, &fd, &cd, s));
  return m.decode_open(fd, cd, B, rng, s);
  VC_GUARD_END
}

int vc_decode_step(vc_handle* h, const int32_t* tokens_host, int M, float* probs_host, void* stream) {
  VC_GUARD_BEGIN
  if (!h || !tokens_host) return set_error(VC_E_ARG, "vc_decode_step: null argument");
  cudaSetDevice(h->m.device);
  return h->m.decode_step(tokens_host, M, probs_host, (cudaStream_t)stream);
  VC_GUARD_END
}

int vc_decode_state_get(vc_handle* h, float* c_host, float* h_host, void* stream) {
  VC_GUARD_BEGIN
  if (!h || !c_host || !h_host) return set_error(VC_E_ARG, "vc_decode_state_get: null argument");
  cudaSetDevice(h->m.device);
  return h->m.decode_state(c_host, h_host, nullptr, nullptr, (cudaStream_t)stream);
  VC_GUARD_END
}

int vc_decode_state_set(vc_handle* h, const float* c_host, const float* h_host, void* stream) {
  VC_GUARD_BEGIN
  if (!h || !c_host || !h_host) return set_error(VC_E_ARG, "vc_decode_state_set: null argument");
  cudaSetDevice(h->m.device);
  return h->m.decode_state(nullptr, nullptr, c_host, h_host, (cudaStream_t)stream);
  VC_GUARD_END
}

int vc_vgg_activation(vc_handle* h, const char* layer, float* dst_host) {
  VC_GUARD_BEGIN
  if (!h || !layer || !dst_host) return set_error(VC_E_ARG, "vc_vgg_activation: null argument");
  cudaSetDevice(h->m.device);
  return h->m.vgg_activation(layer, dst_host);
  VC_GUARD_END
}

}  // extern "C"
