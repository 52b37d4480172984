// Train / eval step orchestration of the CVAE captioning graph (reference: main.py:43-191 builds
// it, main.py:229-244 runs it). Everything is enqueued on the caller's stream; the only
// synchronisation is the optional fetch of the step scalars.
#include "model.h"
#include <cmath>

namespace vc {

static inline int64_t round_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

Model::~Model() {
  comm_release();
  decode_release();
  for (void* p : allocs) cudaFree(p);
  for (StageSlot& q : slots) {
    if (q.ready) cudaEventDestroy(q.ready);
    if (q.consumed) cudaEventDestroy(q.consumed);
    if (q.img_ready) cudaEventDestroy(q.img_ready);
    if (q.px_free) cudaEventDestroy(q.px_free);
  }
  if (host_scal) cudaFreeHost(host_scal);
  for (auto& p : pending)
    if (p.ev) cudaEventDestroy(p.ev);
  if (side) cudaStreamDestroy(side);
  if (h2d_stream) cudaStreamDestroy(h2d_stream);
  if (ev_fork) cudaEventDestroy(ev_fork);
  if (ev_join) cudaEventDestroy(ev_join);
}

int Model::add_param(const std::string& name, std::vector<int64_t> shape, int region, bool hidden) {
  ParamInfo pi;
  pi.name = name;
  pi.ndim = (int)shape.size();
  pi.count = 1;
  for (int i = 0; i < 4; ++i) pi.shape[i] = i < pi.ndim ? shape[i] : 1;
  for (auto d : shape) pi.count *= d;
  pi.region = region;
  pi.trainable = region != 1;
  pi.offset = -1;
  pi.hidden = hidden;
  index[name] = (int)params.size();
  if (!hidden) visible.push_back((int)params.size());
  params.push_back(pi);
  return VC_OK;
}

int Model::add_view(const std::string& name, std::vector<int64_t> shape, int parent, int64_t parent_off, int64_t ld) {
  VC_TRY(add_param(name, shape, params[parent].region, false));
  ParamInfo& pi = params.back();
  pi.parent = parent;
  pi.parent_off = parent_off;
  pi.ld = ld;
  return VC_OK;
}

int Model::param_index(const char* name) const { return pidx(name); }

static const char* kVggNames[13] = {"conv1_1", "conv1_2", "conv2_1", "conv2_2", "conv3_1", "conv3_2", "conv3_3",
                                    "conv4_1", "conv4_2", "conv4_3", "conv5_1", "conv5_2", "conv5_3"};
static const int kVggCin[13] = {3, 64, 64, 128, 128, 256, 256, 256, 512, 512, 512, 512, 512};
static const int kVggCout[13] = {64, 64, 128, 128, 256, 256, 256, 512, 512, 512, 512, 512, 512};

int Model::init(const vc_config& c, int dev) {
  cfg = c;
  device = dev;
  VC_CUDA(cudaSetDevice(dev));
  const int E = cfg.embed_size, He = cfg.encoder_hidden, Hd = cfg.decoder_hidden, Z = cfg.latent_size,
            S = cfg.gen_z_samples, V = cfg.vocab_size, F = cfg.cnn_feature_size, K = cfg.num_clusters;
  if (E % 64 || He % 64 || Hd % 64) return set_error(VC_E_SHAPE, "embed/hidden sizes must be multiples of 64");
  if (F % 8) return set_error(VC_E_SHAPE, "cnn_feature_size must be a multiple of 8");
  if (!cfg.no_encoder && ((int64_t)S * Z) % 8) return set_error(VC_E_SHAPE, "gen_z_samples*latent must be a multiple of 8");
  if (cfg.prior < 0 || cfg.prior > 2) return set_error(VC_E_ARG, "unknown prior %d", cfg.prior);
  if (cfg.optimizer < 0 || cfg.optimizer > 2 || cfg.cnn_optimizer < 0 || cfg.cnn_optimizer > 2)
    return set_error(VC_E_ARG, "unknown optimizer %d / %d (VC_OPT_ADAM, VC_OPT_SGD, VC_OPT_MOMENTUM)", cfg.optimizer, cfg.cnn_optimizer);
  if (cfg.num_captions < 1 || cfg.max_batch < 1 || cfg.max_len < 1) return set_error(VC_E_ARG, "bad batch geometry");
  const bool has_cv = cfg.use_c_v || cfg.prior != VC_PRIOR_NORMAL;
  maxN = cfg.max_batch * cfg.num_captions;
  maxT = cfg.max_len;
  ZP = (int)round_up(Z, 8);
  VP = (int)round_up(V, 8);
  KP = (int)round_up(K, 8);

  // ---- parameter table, region 0 first, embeddings last inside region 0
  add_param("imf_emb/kernel", {F, E}, 0);
  add_param("imf_emb/bias", {E}, 0);
  if (has_cv) {
    const int r = cfg.use_c_v ? 0 : 1;  // created for GMM/AG even without --c_v but then unused (no gradient)
    add_param("cv_emb/kernel", {K, E}, r);
    add_param("cv_emb/bias", {E}, r);
  }
  if (!cfg.no_encoder) {
    add_param("encoder/multi_rnn_cell/cell_0/lstm_cell/kernel", {E + He, 4 * He}, 0);
    add_param("encoder/multi_rnn_cell/cell_0/lstm_cell/bias", {4 * He}, 0);
    // posterior heads: two packed blocks + one strided view per TF variable
    heads_nh = cfg.prior == VC_PRIOR_NORMAL ? 1 : K;
    heads_cols = heads_nh * 2 * ZP;
    add_param("__heads/kernel", {He, heads_cols}, 0, true);
    p_heads_w = (int)params.size() - 1;
    add_param("__heads/bias", {heads_cols}, 0, true);
    p_heads_b = (int)params.size() - 1;
    for (int k = 0; k < heads_nh; ++k) {
      char scope[64] = "encoder";
      if (cfg.prior != VC_PRIOR_NORMAL)
        snprintf(scope, sizeof(scope), "encoder/%s_%d", cfg.prior == VC_PRIOR_GMM ? "gmm_ll" : "ag_ll", k);
      for (int w = 0; w < 2; ++w) {
        const int64_t col = ((int64_t)k * 2 + w) * ZP;
        const std::string layer = std::string(scope) + (w == 0 ? "/dense" : "/dense_1");
        add_view(layer + "/kernel", {He, Z}, p_heads_w, col, heads_cols);
        add_view(layer + "/bias", {Z}, p_heads_b, col, 0);
      }
    }
  }
  add_param("decoder/net/multi_rnn_cell/cell_0/lstm_cell/kernel", {E + Hd, 4 * Hd}, 0);
  add_param("decoder/net/multi_rnn_cell/cell_0/lstm_cell/bias", {4 * Hd}, 0);
  if (!cfg.no_encoder) {
    add_param("decoder/net/z_rnn/kernel", {(int64_t)Z * S, E}, 0);
    add_param("decoder/net/z_rnn/bias", {E}, 0);
  }
  add_param("decoder/rnn_logits/kernel", {Hd, V}, 0);
  add_param("decoder/rnn_logits/bias", {V}, 0);
  const int first_emb = (int)params.size();
  if (!cfg.no_encoder) add_param("encoder/enc_embeddings", {V, E}, 0);
  add_param("decoder/net/dec_embeddings", {V, E}, 0);
  if (cfg.with_cnn || cfg.fine_tune) {
    const int r = cfg.fine_tune ? 2 : 1;
    for (int l = 0; l < 13; ++l) {
      const bool c5 = l >= 10;  // conv5_x variables are named weights_conv / biases_conv (image_embeddings.py:176-201)
      add_param(std::string("cnn/") + kVggNames[l] + (c5 ? "/weights_conv" : "/weights"), {3, 3, kVggCin[l], kVggCout[l]}, r);
      add_param(std::string("cnn/") + kVggNames[l] + (c5 ? "/biases_conv" : "/biases"), {kVggCout[l]}, r);
    }
    add_param("cnn/fc1/weights", {25088, 4096}, r);
    add_param("cnn/fc1/biases", {4096}, r);
    add_param("cnn/fc2/weights", {4096, 4096}, r);
    add_param("cnn/fc2/biases", {4096}, r);
  }
  // offsets: region 0 (dense, then embeddings), 64-float tail, region 1, region 2; views resolve to their parent
  int64_t off = 0;
  for (int i = 0; i < (int)params.size(); ++i)
    if (params[i].region == 0 && i < first_emb && params[i].parent < 0) { params[i].offset = off; off += round_up(params[i].count, 64); }
  n_dense = off;
  for (int i = first_emb; i < (int)params.size(); ++i)
    if (params[i].region == 0 && params[i].parent < 0) { params[i].offset = off; off += round_up(params[i].count, 64); }
  n_adam = off;
  off += 64;
  for (auto& p : params)
    if (p.region == 1 && p.parent < 0) { p.offset = off; off += round_up(p.count, 64); }
  n_frozen = off - n_adam - 64;
  for (auto& p : params)
    if (p.region == 2 && p.parent < 0) { p.offset = off; off += round_up(p.count, 64); }
  n_cnn = off - n_adam - 64 - n_frozen;
  n_total = off;
  for (auto& p : params)
    if (p.parent >= 0) p.offset = params[p.parent].offset + p.parent_off;
  VC_TRY(dalloc(&Pf, n_total));
  const int64_t n_opt = cfg.fine_tune ? n_total : n_adam + 64;
  VC_TRY(dalloc(&Gf, n_opt));
  VC_TRY(dalloc(&Mf, n_opt));
  VC_TRY(dalloc(&Vf, n_opt));
  g_tail = Gf + n_adam;

  // ---- shadows
  const int N = maxN, T = maxT;
  VC_TRY(dalloc((uint16_t**)&imf_wt, (size_t)E * F));
  VC_TRY(dalloc((uint16_t**)&imf_nat, (size_t)F * E));
  if (has_cv) VC_TRY(dalloc((uint16_t**)&cv_wt, (size_t)E * KP));
  if (has_cv) VC_TRY(dalloc((uint16_t**)&cv_nat, (size_t)KP * E));
  auto setup_lstm = [&](LstmNet& L, const char* kname, const char* bname, int H, int pre) -> int {
    L.p_kernel = pidx(kname);
    L.p_bias = pidx(bname);
    L.E = E;
    L.H = H;
    L.pre = pre;
    L.steps = pre + T;
    VC_TRY(dalloc((uint16_t**)&L.w_t_perm, (size_t)4 * H * (E + H)));
    if (lstm_seq_units(N, H) == 32) VC_TRY(dalloc((uint16_t**)&L.w_t_perm32, (size_t)4 * H * (E + H)));
    VC_TRY(dalloc((uint16_t**)&L.w_nat, (size_t)4 * H * (E + H)));
    VC_TRY(dalloc((uint16_t**)&L.X, (size_t)L.steps * N * E));
    VC_TRY(dalloc((uint16_t**)&L.Hs, (size_t)(L.steps + 1) * N * H));
    VC_TRY(dalloc(&L.Cs, (size_t)(L.steps + 1) * N * H));
    VC_TRY(dalloc((uint16_t**)&L.G, (size_t)L.steps * N * 4 * H));
    VC_TRY(dalloc((uint16_t**)&L.dG, (size_t)L.steps * N * 4 * H));
    VC_TRY(dalloc(&L.dX, (size_t)L.steps * N * E));
    VC_TRY(dalloc(&L.dh_carry, (size_t)N * H));
    VC_TRY(dalloc(&L.dc_carry, (size_t)N * H));
    return VC_OK;
  };
  const int cvstep = cfg.use_c_v ? 1 : 0;
  if (!cfg.no_encoder)
    VC_TRY(setup_lstm(enc, "encoder/multi_rnn_cell/cell_0/lstm_cell/kernel", "encoder/multi_rnn_cell/cell_0/lstm_cell/bias",
                      He, 1 + cvstep));
  VC_TRY(setup_lstm(dec, "decoder/net/multi_rnn_cell/cell_0/lstm_cell/kernel",
                    "decoder/net/multi_rnn_cell/cell_0/lstm_cell/bias", Hd, 1 + cvstep + (cfg.no_encoder ? 0 : 1)));
  if (!cfg.no_encoder) {
    VC_TRY(dalloc((uint16_t**)&heads_nat, (size_t)He * heads_cols));
    heads_bias = pp(p_heads_b);
    VC_TRY(dalloc((uint16_t**)&z_wt, (size_t)E * S * Z));
    VC_TRY(dalloc((uint16_t**)&z_nat, (size_t)E * S * Z));
    VC_TRY(dalloc((uint16_t**)&enc_emb_h, (size_t)V * E));
  }
  VC_TRY(dalloc((uint16_t**)&wo_t, (size_t)V * Hd));
  VC_TRY(dalloc((uint16_t**)&wo_nat, (size_t)Hd * VP));
  VC_TRY(dalloc((uint16_t**)&dec_emb_h, (size_t)V * E));

  // ---- workspace
  const int B = cfg.max_batch;
  VC_TRY(dalloc((uint16_t**)&feats_h, (size_t)B * F));
  VC_TRY(dalloc(&imf_f, (size_t)B * E));
  VC_TRY(dalloc((uint16_t**)&dimf_h, (size_t)B * E));
  if (has_cv) {
    VC_TRY(dalloc((uint16_t**)&cv_h, (size_t)N * KP));
    VC_TRY(dalloc(&cv_f, (size_t)N * E));
    VC_TRY(dalloc((uint16_t**)&dcv_h, (size_t)N * E));
  }
  VC_TRY(dalloc((uint16_t**)&Out, (size_t)T * N * Hd));
  VC_TRY(dalloc((uint16_t**)&logits, (size_t)T * N * VP));
  VC_TRY(dalloc(&dOut, (size_t)T * N * Hd));
  VC_TRY(dalloc(&ce_row, (size_t)T * N));
  if (!cfg.no_encoder) {
    VC_TRY(dalloc(&heads_f, (size_t)N * heads_cols));
    VC_TRY(dalloc(&mu, (size_t)N * Z));
    VC_TRY(dalloc(&sd, (size_t)N * Z));
    VC_TRY(dalloc(&kl_row, (size_t)N));
    VC_TRY(dalloc(&dkl_dmu, (size_t)N * Z));
    VC_TRY(dalloc(&dkl_dsd, (size_t)N * Z));
    VC_TRY(dalloc(&cm, (size_t)N * Z));
    VC_TRY(dalloc(&dmu_t, (size_t)N * Z));
    VC_TRY(dalloc(&dsd_t, (size_t)N * Z));
    VC_TRY(dalloc(&c_means, (size_t)K * Z));
    VC_TRY(dalloc(&gmm_pick, (size_t)N));
    VC_TRY(dalloc((uint16_t**)&z, (size_t)S * N * Z));
    VC_TRY(dalloc(&zdec_f, (size_t)N * E));
    VC_TRY(dalloc((uint16_t**)&dzdec_h, (size_t)N * E));
    VC_TRY(dalloc(&dz, (size_t)S * N * Z));
    VC_TRY(dalloc((uint16_t**)&dheads, (size_t)N * heads_cols));
  }
  VC_TRY(dalloc(&scal, 64));
  VC_TRY(dalloc(&seq_flags, (size_t)lstm_seq_flag_count(maxN, dec.steps + 1)));
  const size_t img_elems = cfg.fine_tune ? (size_t)224 * 224 * 3 : (size_t)F;
  VC_TRY(dalloc(&st_feats, (size_t)B * img_elems));
  VC_TRY(dalloc(&st_cv, (size_t)N * K));
  VC_TRY(dalloc(&st_lbl, (size_t)N * T));
  VC_TRY(dalloc(&st_in, (size_t)N * T));
  VC_TRY(dalloc(&st_len, (size_t)N));
  VC_CUDA(cudaMallocHost((void**)&host_scal, 64 * sizeof(float)));
  VC_CUDA(cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking));
  VC_CUDA(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
  VC_CUDA(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
  {  // SMs the whole-sequence recurrences leave idle at the handle's batch size (minus a few for NCCL's channels)
    const int lstm_ctas = ((maxN + 127) / 128) * (cfg.decoder_hidden / lstm_seq_units(maxN, cfg.decoder_hidden));
    side_sm_cap = std::max(24, num_sms() - lstm_ctas - 12);
    if (const char* e = getenv("VC_SIDE_SMS")) side_sm_cap = std::max(8, atoi(e));
  }
  if (cfg.with_cnn || cfg.fine_tune) VC_TRY(vgg_init());
  shadows_dirty = true;
  return VC_OK;
}

// ------------------------------------------------------------------------------------------
// Copies variable i between the flat device buffer `base` (Pf or Gf) and a dense host array; strided for views.
int Model::copy_var(int i, float* base, float* host_dst, const float* host_src) {
  const ParamInfo& p = params[i];
  float* dev = base + p.offset;
  if (p.ld > 0 && p.ndim == 2) {
    const size_t w = (size_t)p.shape[1] * sizeof(float);
    if (host_src)
      VC_CUDA(cudaMemcpy2D(dev, (size_t)p.ld * sizeof(float), host_src, w, w, (size_t)p.shape[0], cudaMemcpyHostToDevice));
    else
      VC_CUDA(cudaMemcpy2D(host_dst, w, dev, (size_t)p.ld * sizeof(float), w, (size_t)p.shape[0], cudaMemcpyDeviceToHost));
  } else {
    if (host_src)
      VC_CUDA(cudaMemcpy(dev, host_src, p.count * sizeof(float), cudaMemcpyHostToDevice));
    else
      VC_CUDA(cudaMemcpy(host_dst, dev, p.count * sizeof(float), cudaMemcpyDeviceToHost));
  }
  return VC_OK;
}

int Model::param_set(const char* name, const float* src) {
  const int i = pidx(name);
  if (i < 0 || params[i].hidden) return set_error(VC_E_ARG, "unknown variable '%s'", name);
  VC_CUDA(cudaDeviceSynchronize());
  VC_TRY(copy_var(i, Pf, nullptr, src));
  shadows_dirty = true;
  vgg_shadows_dirty = true;
  return VC_OK;
}
int Model::param_get(const char* name, float* dst) {
  const int i = pidx(name);
  if (i < 0 || params[i].hidden) return set_error(VC_E_ARG, "unknown variable '%s'", name);
  VC_CUDA(cudaDeviceSynchronize());
  return copy_var(i, Pf, dst, nullptr);
}
int Model::grad_get(const char* name, float* dst) {
  const int i = pidx(name);
  if (i < 0 || params[i].hidden) return set_error(VC_E_ARG, "unknown variable '%s'", name);
  if (params[i].region == 1 || (params[i].region == 2 && !cfg.fine_tune))
    return set_error(VC_E_STATE, "variable '%s' has no gradient in this configuration", name);
  VC_CUDA(cudaDeviceSynchronize());
  VC_TRY(copy_var(i, Gf, dst, nullptr));
  if (params[i].region == 2 && cfg.weight_decay != 0.f) {  // + d(weight_decay * sum(w^2) / 2)/dw, Q11
    std::vector<float> w((size_t)params[i].count);
    VC_TRY(copy_var(i, Pf, w.data(), nullptr));
    for (int64_t k = 0; k < params[i].count; ++k) dst[k] += cfg.weight_decay * w[(size_t)k];
  }
  return VC_OK;
}
int Model::set_cluster_means(const float* src) {
  if (c_means == nullptr) return set_error(VC_E_STATE, "this configuration has no encoder (no cluster means)");
  VC_CUDA(cudaDeviceSynchronize());
  VC_CUDA(cudaMemcpy(c_means, src, (size_t)cfg.num_clusters * cfg.latent_size * sizeof(float), cudaMemcpyHostToDevice));
  return VC_OK;
}

// ------------------------------------------------------------------------------------------
// bf16 operand shadows of the fp32 master weights, rebuilt after every optimiser step. The train / eval graph reads every
// dense kernel in its NATURAL [in, out] layout (MN-major B operand of the tcgen05 GEMMs, K-major B of the dgrad GEMMs), so
// the per-step refresh is plain casts plus the gate-interleaved LSTM kernels; the transposed copies the generation path
// uses (decode.cu) are rebuilt lazily, the first time a decode call follows a weight change (refresh_decode_shadows).
int Model::refresh_shadows(cudaStream_t s) {
  const int E = cfg.embed_size, Z = cfg.latent_size, S = cfg.gen_z_samples, V = cfg.vocab_size, F = cfg.cnn_feature_size,
            K = cfg.num_clusters, Hd = cfg.decoder_hidden;
  RefreshJobs jobs;  // one launch for all of them (elementwise.cu: k_refresh_multi)
  jobs.cast(pp(pidx("imf_emb/kernel")), imf_nat, F, E, E, E);
  if (cv_nat) jobs.cast(pp(pidx("cv_emb/kernel")), cv_nat, K, E, E, E);
  auto lstm = [&](LstmNet& L) {
    jobs.transpose(pp(L.p_kernel), L.w_t_perm, L.E + L.H, 4 * L.H, 4 * L.H, L.E + L.H, L.H, 64);
    if (L.w_t_perm32) jobs.transpose(pp(L.p_kernel), L.w_t_perm32, L.E + L.H, 4 * L.H, 4 * L.H, L.E + L.H, L.H, 32);
    jobs.cast(pp(L.p_kernel), L.w_nat, L.E + L.H, 4 * L.H, 4 * L.H, 4 * L.H);
  };
  if (!cfg.no_encoder) {
    lstm(enc);
    const int He = cfg.encoder_hidden;
    jobs.cast(pp(p_heads_w), heads_nat, He, heads_cols, heads_cols, heads_cols);
    jobs.cast(pp(pidx("decoder/net/z_rnn/kernel")), z_nat, S * Z, E, E, E);
    jobs.cast(pp(pidx("encoder/enc_embeddings")), enc_emb_h, V, E, E, E);
  }
  lstm(dec);
  const int po = pidx("decoder/rnn_logits/kernel");
  jobs.transpose(pp(po), wo_t, Hd, V, V, Hd, 0, 0);  // K-major for the forward vocabulary projection (MN-major B measured 20 % slower there)
  jobs.cast(pp(po), wo_nat, Hd, V, V, VP);
  jobs.cast(pp(pidx("decoder/net/dec_embeddings")), dec_emb_h, V, E, E, E);
  VC_TRY(refresh_multi(s, jobs));
  shadows_dirty = false;
  decode_shadows_dirty = true;
  return VC_OK;
}

int Model::refresh_decode_shadows(cudaStream_t s) {
  if (shadows_dirty) VC_TRY(refresh_shadows(s));
  if (!decode_shadows_dirty) return VC_OK;
  ProfTag ptag("refresh_shadows");
  const int E = cfg.embed_size, Z = cfg.latent_size, S = cfg.gen_z_samples, F = cfg.cnn_feature_size, K = cfg.num_clusters;
  VC_TRY(transpose_cast(s, pp(pidx("imf_emb/kernel")), imf_wt, F, E, E, F, 0, 0));
  if (cv_wt) VC_TRY(transpose_cast(s, pp(pidx("cv_emb/kernel")), cv_wt, K, E, E, KP, 0, 0));
  if (!cfg.no_encoder)
    VC_TRY(transpose_cast(s, pp(pidx("decoder/net/z_rnn/kernel")), z_wt, S * Z, E, E, (int64_t)S * Z, 0, 0));
  decode_shadows_dirty = false;
  return VC_OK;
}

// ------------------------------------------------------------------------------------------
int Model::stage_inputs(const float* feats, const int32_t* lbl, const int32_t* inp, const int32_t* len, const float* cv,
                        int B, int T, StepInputs* out, cudaStream_t s, bool feats_u8) {
  if (B < 1 || B > cfg.max_batch || T < 1 || T > maxT)
    return set_error(VC_E_SHAPE, "batch %d x len %d exceeds the handle's max_batch %d / max_len %d", B, T, cfg.max_batch, maxT);
  const int N = B * cfg.num_captions;
  const size_t fe = cfg.fine_tune ? (size_t)224 * 224 * 3 : (size_t)cfg.cnn_feature_size;
  if (feats != nullptr) VC_CUDA(cudaMemcpyAsync(st_feats, feats, B * fe * (feats_u8 ? 1 : sizeof(float)), cudaMemcpyHostToDevice, s));
  out->feats_u8 = feats_u8;
  VC_CUDA(cudaMemcpyAsync(st_lbl, lbl, (size_t)N * T * sizeof(int32_t), cudaMemcpyHostToDevice, s));
  VC_CUDA(cudaMemcpyAsync(st_in, inp, (size_t)N * T * sizeof(int32_t), cudaMemcpyHostToDevice, s));
  VC_CUDA(cudaMemcpyAsync(st_len, len, (size_t)N * sizeof(int32_t), cudaMemcpyHostToDevice, s));
  out->feats = st_feats;
  out->cap_lbl = st_lbl;
  out->cap_in = st_in;
  out->len = st_len;
  out->c_v = nullptr;
  if (cv != nullptr) {
    VC_CUDA(cudaMemcpyAsync(st_cv, cv, (size_t)N * cfg.num_clusters * sizeof(float), cudaMemcpyHostToDevice, s));
    out->c_v = st_cv;
  }
  out->B = B;
  out->T = T;
  return VC_OK;
}

// ------------------------------------------------------------------------------------------
// Double-buffered feed. kind: 0 = what vc_train_step takes (fc2 features, or fp32 images when fine_tune),
// 1 = fp32 images for the on-device VGG16 forward, 2 = uint8 images (fine-tune feed or on-device forward).
int Model::stage_slot(int slot, const void* px_host, int kind, const int32_t* lbl, const int32_t* inp, const int32_t* len,
                      const float* cv, int B, int T, cudaStream_t cs) {
  if (slot < 0 || slot > 1) return set_error(VC_E_ARG, "staging slot must be 0 or 1, got %d", slot);
  if (kind < 0 || kind > 2) return set_error(VC_E_ARG, "unknown feed kind %d", kind);
  if (B < 1 || B > cfg.max_batch || T < 1 || T > maxT)
    return set_error(VC_E_SHAPE, "batch %d x len %d exceeds the handle's max_batch %d / max_len %d", B, T, cfg.max_batch, maxT);
  const bool images = kind != 0 || cfg.fine_tune;
  if (kind != 0 && !cfg.fine_tune && vgg.empty())
    return set_error(VC_E_STATE, "this handle was created without the CNN (with_cnn = 0)");
  StageSlot& q = slots[slot];
  const int N = B * cfg.num_captions;
  if (q.lbl == nullptr) {
    const size_t px_elems = (cfg.fine_tune || !vgg.empty()) ? (size_t)224 * 224 * 3 : (size_t)cfg.cnn_feature_size;
    VC_TRY(dalloc((float**)&q.px, (size_t)cfg.max_batch * std::max(px_elems, (size_t)cfg.cnn_feature_size), false));
    if (!vgg.empty() && !cfg.fine_tune) VC_TRY(dalloc(&q.fc2, (size_t)cfg.max_batch * 4096, false));
    VC_TRY(dalloc(&q.cv, (size_t)maxN * cfg.num_clusters, false));
    VC_TRY(dalloc(&q.lbl, (size_t)maxN * maxT, false));
    VC_TRY(dalloc(&q.in, (size_t)maxN * maxT, false));
    VC_TRY(dalloc(&q.len, (size_t)maxN, false));
    VC_CUDA(cudaEventCreateWithFlags(&q.ready, cudaEventDisableTiming));
    VC_CUDA(cudaEventCreateWithFlags(&q.consumed, cudaEventDisableTiming));
    VC_CUDA(cudaEventCreateWithFlags(&q.img_ready, cudaEventDisableTiming));
    VC_CUDA(cudaEventCreateWithFlags(&q.px_free, cudaEventDisableTiming));
  }
  const size_t per = images ? (size_t)224 * 224 * 3 * (kind == 2 ? 1 : sizeof(float)) : (size_t)cfg.cnn_feature_size * sizeof(float);
  const bool lookahead = kind != 0 && !cfg.fine_tune;  // frozen extractor: VGG16 forward behind the copy (below)
  if (lookahead) {
    // The image buffer is free as soon as the forward pass that read it is done -- long before the step that uses the
    // slot's features and labels is -- and the copy runs on its own stream: in `cs` it would queue behind the forward of
    // the PREVIOUS batch, and the forward of this one would then wait 1.5 ms for 38.7 MB to arrive with only the caption
    // model to keep the SMs busy (10.09 against 9.70 ms per step for the device-resident loop).
    if (h2d_stream == nullptr) VC_CUDA(cudaStreamCreateWithFlags(&h2d_stream, cudaStreamNonBlocking));
    if (q.px_used) VC_CUDA(cudaStreamWaitEvent(h2d_stream, q.px_free, 0));
    VC_CUDA(cudaMemcpyAsync(q.px, px_host, (size_t)B * per, cudaMemcpyHostToDevice, h2d_stream));
    VC_CUDA(cudaEventRecord(q.img_ready, h2d_stream));
  }
  if (q.in_use) VC_CUDA(cudaStreamWaitEvent(cs, q.consumed, 0));  // the step that last read this slot must be past it
  if (!lookahead) VC_CUDA(cudaMemcpyAsync(q.px, px_host, (size_t)B * per, cudaMemcpyHostToDevice, cs));
  VC_CUDA(cudaMemcpyAsync(q.lbl, lbl, (size_t)N * T * sizeof(int32_t), cudaMemcpyHostToDevice, cs));
  VC_CUDA(cudaMemcpyAsync(q.in, inp, (size_t)N * T * sizeof(int32_t), cudaMemcpyHostToDevice, cs));
  VC_CUDA(cudaMemcpyAsync(q.len, len, (size_t)N * sizeof(int32_t), cudaMemcpyHostToDevice, cs));
  q.has_cv = cv != nullptr;
  if (cv) VC_CUDA(cudaMemcpyAsync(q.cv, cv, (size_t)N * cfg.num_clusters * sizeof(float), cudaMemcpyHostToDevice, cs));
  // Frozen extractor: the VGG16 forward of the staged images does not depend on the step in flight (its weights never
  // change), so it runs here, behind the copy on the copy stream, and overlaps the caption model of the previous batch
  // (whose recurrent kernels leave SMs idle). The step then starts from the slot's fc2 features.
  if (lookahead) {
    VC_CUDA(cudaStreamWaitEvent(cs, q.img_ready, 0));
    VC_TRY(vgg_forward(reinterpret_cast<const float*>(q.px), q.fc2, B, false, nullptr, cs, kind == 2));
    VC_CUDA(cudaEventRecord(q.px_free, cs));
    q.px_used = true;
  }
  VC_CUDA(cudaEventRecord(q.ready, cs));
  q.B = B;
  q.T = T;
  q.kind = kind;
  q.filled = true;
  return VC_OK;
}

int Model::step_from_slot(int slot, int64_t gs, const vc_rng* rng, cudaStream_t s, bool apply_update) {
  if (slot < 0 || slot > 1 || !slots[slot].filled)
    return set_error(VC_E_STATE, "vc_train_step_staged: slot %d holds no batch (call vc_stage_batch first)", slot);
  StageSlot& q = slots[slot];
  VC_CUDA(cudaStreamWaitEvent(s, q.ready, 0));
  StepInputs in{};
  in.B = q.B;
  in.T = q.T;
  in.cap_lbl = q.lbl;
  in.cap_in = q.in;
  in.len = q.len;
  in.c_v = q.has_cv ? q.cv : nullptr;
  in.global_step = gs;
  if (rng) in.rng = *rng;
  if (q.kind != 0 && !cfg.fine_tune) {  // frozen extractor: fc2 was computed behind the copy (stage_slot)
    in.feats = q.fc2;
  } else {
    in.feats = reinterpret_cast<const float*>(q.px);
    in.feats_u8 = q.kind == 2;
  }
  VC_TRY(forward(in, true, s));
  VC_TRY(backward(in, s));
  VC_CUDA(cudaEventRecord(q.consumed, s));
  q.in_use = true;
  q.filled = false;
  return apply_update ? apply(step_scale(), s) : VC_OK;
}

// ------------------------------------------------------------------------------------------
int Model::lstm_forward(LstmNet& L, int N, int T, const int32_t* len, void* out, const float* out_keep, cudaStream_t s) {
  const int steps = L.pre + T;
  uint16_t* X = (uint16_t*)L.X;
  uint16_t* Hs = (uint16_t*)L.Hs;
  uint16_t* Gt = (uint16_t*)L.G;
  // zero initial state (cell_0.zero_state, encoder.py:42-44 / decoder.py:96-98); slot 0 is re-zeroed every
  // call because the slot pitch depends on the current batch size
  VC_CUDA(cudaMemsetAsync(Hs, 0, (size_t)N * L.H * 2, s));
  VC_CUDA(cudaMemsetAsync(L.Cs, 0, (size_t)N * L.H * sizeof(float), s));
  if (lstm_seq_applicable(N, L.H, steps)) {  // one persistent launch for the whole sequence
    LstmSeqFwdArgs a{};
    a.X = X; a.Hs = Hs; a.Cs = L.Cs; a.G = Gt; a.out = out;
    a.w_t_perm = L.w_t_perm;
    a.w_t_perm32 = L.w_t_perm32;
    a.bias = pp(L.p_bias);
    a.lengths = len;
    a.out_keep = out_keep;
    a.inv_keep = 1.f / cfg.dec_lstm_drop;
    a.flags = seq_flags;
    a.pre = L.pre; a.T = T; a.steps = steps; a.N = N; a.E = L.E; a.H = L.H;
    return lstm_fwd_seq(s, a);
  }
  for (int st = 0; st < steps; ++st) {
    LstmFwdArgs a{};
    const int t = st - L.pre;
    a.x = X + (size_t)st * N * L.E;
    a.h_prev = Hs + (size_t)st * N * L.H;
    a.c_prev = L.Cs + (size_t)st * N * L.H;
    a.h_out = Hs + (size_t)(st + 1) * N * L.H;
    a.c_out = L.Cs + (size_t)(st + 1) * N * L.H;
    a.gates = Gt + (size_t)st * N * 4 * L.H;
    a.out = (out != nullptr && t >= 0) ? (uint16_t*)out + (size_t)t * N * L.H : nullptr;
    a.w_t_perm = L.w_t_perm;
    a.bias = pp(L.p_bias);
    a.lengths = t >= 0 ? len : nullptr;
    a.out_keep = (out_keep != nullptr && t >= 0) ? out_keep + (size_t)t * L.H : nullptr;  // mask is [N, T, H]
    a.out_keep_ld = (long long)T * L.H;
    a.inv_keep = 1.f / cfg.dec_lstm_drop;
    a.t = t;
    a.N = N;
    a.E = L.E;
    a.H = L.H;
    VC_TRY(lstm_fwd_step(s, a));
  }
  return VC_OK;
}

// Side stream of the backward pass (VC_BWD_OVERLAP=0 turns it off; per-kernel profiling turns it off so that every
// kernel is timed running alone). side_begin makes `side` wait for everything enqueued on `s` so far and returns
// whether work should go there; side_join makes `s` wait for everything enqueued on `side`.
bool Model::side_begin(cudaStream_t s) {
  static const bool enabled = [] { const char* e = getenv("VC_BWD_OVERLAP"); return !(e && e[0] == '0'); }();
  if (!enabled || side == nullptr || prof_is_on()) return false;
  if (cudaEventRecord(ev_fork, s) != cudaSuccess || cudaStreamWaitEvent(side, ev_fork, 0) != cudaSuccess) return false;
  side_used = true;
  return true;
}
int Model::side_join(cudaStream_t s) {
  if (!side_used) return VC_OK;
  VC_CUDA(cudaEventRecord(ev_join, side));
  VC_CUDA(cudaStreamWaitEvent(s, ev_join, 0));
  side_used = false;
  return VC_OK;
}

// defer_weight_grads: only the recurrence and the input gradient run here; the caller runs lstm_weight_grads later (data
// parallel: the decoder's weight / bias gradients wait until the big posterior-head bucket is on the wire, see backward()).
int Model::lstm_backward(LstmNet& L, int N, int T, const int32_t* len, const float* d_out, const float* out_keep,
                         cudaStream_t s, bool defer_weight_grads) {
  const int steps = L.pre + T;
  uint16_t* Gt = (uint16_t*)L.G;
  uint16_t* dGt = (uint16_t*)L.dG;
  const bool seq = lstm_seq_applicable(N, L.H, steps);
  for (int st = steps - 1; st >= (seq ? steps - 1 : 0); --st) {
    LstmBwdArgs a{};
    const int t = st - L.pre;
    a.d_gates_next = st == steps - 1 ? nullptr : dGt + (size_t)(st + 1) * N * 4 * L.H;
    a.w_nat = L.w_nat;
    a.gates = Gt + (size_t)st * N * 4 * L.H;
    a.c_prev = L.Cs + (size_t)st * N * L.H;
    a.c_cur = L.Cs + (size_t)(st + 1) * N * L.H;
    a.d_out = (d_out != nullptr && t >= 0) ? d_out + (size_t)t * N * L.H : nullptr;
    a.out_keep = (out_keep != nullptr && t >= 0) ? out_keep + (size_t)t * L.H : nullptr;
    a.out_keep_ld = (long long)T * L.H;
    a.inv_keep = 1.f / cfg.dec_lstm_drop;
    a.dh_carry = L.dh_carry;
    a.dc_carry = L.dc_carry;
    a.d_gates = dGt + (size_t)st * N * 4 * L.H;
    a.lengths = t >= 0 ? len : nullptr;
    a.t = t;
    a.N = N;
    a.E = L.E;
    a.H = L.H;
    VC_TRY(lstm_bwd_step(s, a));
  }
  if (seq) {  // steps steps-2 .. 0: one persistent launch
    LstmSeqBwdArgs a{};
    a.w_nat = L.w_nat; a.G = Gt; a.Cs = L.Cs; a.d_out = d_out; a.out_keep = out_keep;
    a.inv_keep = 1.f / cfg.dec_lstm_drop;
    a.dh_carry = L.dh_carry; a.dc_carry = L.dc_carry; a.dG = dGt; a.lengths = len; a.flags = seq_flags;
    a.pre = L.pre; a.T = T; a.steps = steps; a.N = N; a.E = L.E; a.H = L.H;
    VC_TRY(lstm_bwd_seq(s, a));
  }
  if (!defer_weight_grads) VC_TRY(lstm_weight_grads(L, N, T, s));
  // input gradient: dX[steps*N, E] = dG x W_x^T  (B = rows 0..E of the natural shadow)
  const long long rows = (long long)steps * N;
  Operand AG{L.dG, rows, 4LL * L.H, 4LL * L.H, false};
  Operand BW{L.w_nat, L.E, 4LL * L.H, 4LL * L.H, false};
  EpiStore ex{};
  ex.out = L.dX;
  ex.ld = L.E;
  ex.alpha = 1.f;
  {
    ProfTag ptag("lstm_dx");
    VC_TRY(gemm_store(s, AG, nullptr, 0, BW, (int)rows, L.E, 4 * L.H, ex, L.E % 256 == 0 ? 256 : 64, 1));
  }
  return VC_OK;
}

int Model::lstm_weight_grads(LstmNet& L, int N, int T, cudaStream_t s) {
  const int steps = L.pre + T;
  // weight gradient: dW[E+H, 4H] = [X ; H_prev]^T x dG over all steps (one GEMM, split-K, fp32 atomics). Nothing in the
  // backward pass waits for it (nor for the bias gradient): with the side stream on, both run there, beside whatever
  // the main stream does next -- for the decoder that is the encoder's BPTT, which leaves 68 of 148 SMs idle.
  // Data parallel: by now the communication stream is moving the vocabulary / decoder / head buckets and NCCL needs the
  // idle SMs more than this GEMM does (8 x B200, cfg 3: 2.80 ms per step with it on the side stream, 2.77 without).
  cudaStream_t main_s = s;
  const bool on_side = comm_world() == 1 && side_begin(s);
  if (on_side) s = side;
  GridCap cap(on_side ? side_sm_cap : 0);
  const long long rows = (long long)steps * N;
  Operand AX{L.X, rows, L.E, L.E, true};
  Operand AH{L.Hs, rows, L.H, L.H, true};
  Operand BG{L.dG, rows, 4LL * L.H, 4LL * L.H, true};
  EpiStore e{};
  e.out = gp(L.p_kernel);
  e.ld = 4 * L.H;
  e.atomic = 1;
  e.alpha = 1.f;
  const int tiles = ((L.E + L.H + 127) / 128) * ((4 * L.H + 255) / 256);
  int splits = std::max(1, num_sms() / std::max(1, tiles));
  {
    ProfTag ptag("lstm_wgrad");
    if (L.E % 128 == 0) {
      VC_TRY(gemm_store(s, AX, &AH, L.E, BG, L.E + L.H, 4 * L.H, (int)rows, e, 256, splits));
    } else {  // E not a multiple of the 128-row tile: two GEMMs
      EpiStore e2 = e;
      VC_TRY(gemm_store(s, AX, nullptr, 0, BG, L.E, 4 * L.H, (int)rows, e, 256, splits));
      e2.out = gp(L.p_kernel) + (size_t)L.E * 4 * L.H;
      VC_TRY(gemm_store(s, AH, nullptr, 0, BG, L.H, 4 * L.H, (int)rows, e2, 256, splits));
    }
  }
  VC_TRY(colsum_bf16(s, L.dG, rows, 4 * L.H, 4 * L.H, gp(L.p_bias)));
  VC_TRY(grad_ready_params({L.p_kernel, L.p_bias}, s));  // data parallel: this bucket leaves from the stream that made it
  (void)main_s;
  return VC_OK;
}

static float annealing_coeff(const vc_config& c, int64_t gs) {  // main.py:162-170
  if (c.fine_tune || c.restore) return 1.f;
  if (c.ann_param > 1.f) return (tanhf(((float)gs - 1000.f * c.ann_param) / 1000.f) + 1.f) / 2.f;
  return 1.f;
}

// ------------------------------------------------------------------------------------------
// scal layout: 0 ce_sum, 1 mask_sum, 2 mask count (pre-pass), 3 kl_sum, 7 global norm, 8 sum(w^2) over cnn/ (fine_tune)
int Model::forward(StepInputs& in, bool write_grad, cudaStream_t s) {
  const int B = in.B, T = in.T, C = cfg.num_captions, N = B * C;
  const int E = cfg.embed_size, Hd = cfg.decoder_hidden, Z = cfg.latent_size, S = cfg.gen_z_samples, V = cfg.vocab_size,
            F = cfg.cnn_feature_size, K = cfg.num_clusters;
  if (B < 1 || B > cfg.max_batch || T < 1 || T > maxT) return set_error(VC_E_SHAPE, "batch/len out of range");
  const bool has_cv = cfg.use_c_v || cfg.prior != VC_PRIOR_NORMAL;
  if (has_cv && in.c_v == nullptr) return set_error(VC_E_ARG, "this configuration needs cluster vectors (c_v)");
  // --dec_drop / --dec_lstm_drop without explicit masks: Philox(seed, global_step) keep masks drawn on the device
  // (tf.nn.dropout draws fresh uniforms every sess.run); the backward pass of the same step reads the same buffers
  if (cfg.dec_keep_rate < 1.f && in.rng.emb_keep_dev == nullptr) {
    if (emb_keep_buf == nullptr) VC_TRY(dalloc(&emb_keep_buf, (size_t)maxN * maxT * E, false));
    VC_TRY(keep_mask(s, emb_keep_buf, (long long)N * T * E, cfg.dec_keep_rate, in.rng.seed ^ 0x243F6A8885A308D3ull,
                     (unsigned long long)in.global_step));
    in.rng.emb_keep_dev = emb_keep_buf;
  }
  if (cfg.dec_lstm_drop < 1.f && in.rng.out_keep_dev == nullptr) {
    if (out_keep_buf == nullptr) VC_TRY(dalloc(&out_keep_buf, (size_t)maxN * maxT * Hd, false));
    VC_TRY(keep_mask(s, out_keep_buf, (long long)N * T * Hd, cfg.dec_lstm_drop, in.rng.seed ^ 0x13198A2E03707344ull,
                     (unsigned long long)in.global_step));
    in.rng.out_keep_dev = out_keep_buf;
  }
  if (shadows_dirty) VC_TRY(refresh_shadows(s));
  VC_CUDA(cudaMemsetAsync(scal, 0, 64 * sizeof(float), s));

  // --fine_tune: the feed is the image batch and fc2 of the trainable VGG16 is the feature (main.py:65-82);
  // un-pooled activations are kept for the backward pass
  const float* feats = in.feats;
  if (cfg.fine_tune) {
    vgg_drop_seed = in.rng.seed;
    vgg_drop_step = (unsigned long long)in.global_step;
    VC_TRY(vgg_forward(in.feats, nullptr, B, true, in.rng.cnn_keep_dev, s, in.feats_u8));
    feats = fc2_f;
    // l2_regularizer(weight_decay) on every cnn/ variable joins rec_loss through get_total_loss (main.py:67-74,
    // 159-160; Q11): scal[8] = sum of squares of the cnn/ region (its padding is zero)
    VC_TRY(sumsq(s, Pf + (n_total - n_cnn), n_cnn, scal + 8));
  }
  // image features -> embedding space (main.py:84-94; projection before tiling, Q7)
  VC_TRY(cast_f32_bf16(s, feats, feats_h, B, F, F, F));
  VC_CUDA(cudaMemsetAsync(imf_f, 0, (size_t)B * E * sizeof(float), s));
  {
    Operand A{feats_h, B, F, F, false}, Bw{imf_nat, F, E, E, true};
    EpiStore e{};
    e.out = imf_f; e.ld = E; e.bias = pp(pidx("imf_emb/bias")); e.atomic = 1; e.alpha = 1.f;
    ProfTag ptag("imf_emb");
    VC_TRY(gemm_store(s, A, nullptr, 0, Bw, B, E, F, e, 64, 8));
  }
  VC_TRY(tile_cast(s, imf_f, cfg.no_encoder ? nullptr : enc.X, dec.X, B, C, E));
  if (has_cv) {  // main.py:103-114
    VC_TRY(cast_f32_bf16(s, in.c_v, cv_h, N, K, K, KP));
    if (cfg.use_c_v) {
      Operand A{cv_h, N, K, KP, false}, Bw{cv_nat, K, E, E, true};
      EpiStore e{};
      e.out = cv_f; e.ld = E; e.bias = pp(pidx("cv_emb/bias")); e.alpha = 1.f;
      ProfTag ptag("cv_emb");
      VC_TRY(gemm_store(s, A, nullptr, 0, Bw, N, E, K, e, 64, 1));
      VC_TRY(tile_cast(s, cv_f, cfg.no_encoder ? nullptr : (uint16_t*)enc.X + (size_t)N * E,
                       (uint16_t*)dec.X + (size_t)N * E, N, 1, E));
    }
  }
  if (!cfg.no_encoder) {
    // q(z | x, I): vae_model/encoder.py:24-110 (the encoder consumes the label sequence, Q8)
    const int He = cfg.encoder_hidden;
    VC_TRY(embed_gather(s, enc_emb_h, in.cap_lbl, (uint16_t*)enc.X + (size_t)enc.pre * N * E, nullptr, 1.f, N, T, E, V));
    VC_TRY(lstm_forward(enc, N, T, in.len, nullptr, nullptr, s));
    const uint16_t* hT = (uint16_t*)enc.Hs + (size_t)(enc.pre + T) * N * He;
    {
      Operand A{hT, N, He, He, false}, Bw{heads_nat, He, heads_cols, heads_cols, true};
      EpiStore e{};
      e.out = heads_f; e.ld = heads_cols; e.bias = heads_bias; e.alpha = 1.f;
      ProfTag ptag("heads");
      VC_TRY(gemm_store(s, A, nullptr, 0, Bw, N, heads_cols, He, e, heads_cols >= 1024 ? 256 : 64, 1));
    }
    last_pick = nullptr;
    if (cfg.prior == VC_PRIOR_GMM) {  // encoder.py:72: explicit draw (parity mode) or in-kernel Philox
      last_pick = in.rng.gmm_cluster_dev;
      if (last_pick == nullptr) {
        VC_TRY(gmm_pick_clusters(s, in.c_v, K, in.rng.seed, (unsigned long long)in.global_step, gmm_pick, N));
        last_pick = gmm_pick;
      }
    }
    VC_TRY(heads_mix(s, heads_f, heads_cols, ZP, cfg.prior, in.c_v, K, last_pick, c_means, mu, sd, cm, N, Z));
    VC_TRY(kl_rows(s, mu, sd, cm, cfg.prior, kl_row, dkl_dmu, dkl_dsd, scal + 3, N, Z));
    VC_TRY(sample_z(s, mu, sd, in.rng.eps_dev, in.rng.seed, (unsigned long long)in.global_step, z, nullptr, S, (long long)N * Z));
    // p(x | z, I): z reshaped row-major [S,N,Z] -> [N, S*Z] (Q1), decoder.py:109-113
    VC_CUDA(cudaMemsetAsync(zdec_f, 0, (size_t)N * E * sizeof(float), s));
    {
      const int SZ = S * Z;
      Operand A{z, N, SZ, SZ, false}, Bw{z_nat, SZ, E, E, true};
      EpiStore e{};
      e.out = zdec_f; e.ld = E; e.bias = pp(pidx("decoder/net/z_rnn/bias")); e.atomic = 1; e.alpha = 1.f;
      const int tiles = ((N + 127) / 128) * ((E + 127) / 128);
      ProfTag ptag("z_rnn");
      VC_TRY(gemm_store(s, A, nullptr, 0, Bw, N, E, SZ, e, 128, std::max(1, num_sms() / tiles)));
    }
    VC_TRY(tile_cast(s, zdec_f, nullptr, (uint16_t*)dec.X + (size_t)(dec.pre - 1) * N * E, N, 1, E));
  }
  // decoder (decoder.py:77-129)
  VC_TRY(embed_gather(s, dec_emb_h, in.cap_in, (uint16_t*)dec.X + (size_t)dec.pre * N * E,
                      cfg.dec_keep_rate < 1.f ? in.rng.emb_keep_dev : nullptr, 1.f / cfg.dec_keep_rate, N, T, E, V));
  VC_TRY(lstm_forward(dec, N, T, in.len, Out, cfg.dec_lstm_drop < 1.f ? in.rng.out_keep_dev : nullptr, s));
  // masked cross-entropy (main.py:152-158); AG differentiates the sum of an [N] lower bound (Q2)
  VC_TRY(count_mask(s, in.cap_lbl, (long long)N * T, scal + 2));
  const float loss_scale = cfg.prior == VC_PRIOR_AG ? (float)N : 1.f;
  const long long rows_all = (long long)T * N;
  // VC_VOCAB_CHUNK=<rows> turns the row-chunked form on (off by default: measured slower, see below and DESIGN section 8)
  static const int chunk_env = [] { const char* e = getenv("VC_VOCAB_CHUNK"); return e ? atoi(e) : 0; }();
  vocab_bwd_done = false;
  if (write_grad && chunk_env > 0 && rows_all > 2 * chunk_env) {
    // Training step: the [T*N, V] logits matrix (580 MB at N = 1280) is produced, turned into dlogits, column-summed
    // and contracted with W_o^T one row chunk at a time, so that a chunk (2048 rows = 46 MB) is still in the 126 MB L2
    // when the next kernel reads it: logits_fwd(chunk) -> CE(chunk) -> bias column sum(chunk) -> dgrad(chunk). Of the six
    // passes the matrix used to make over HBM (fwd write, CE read + write, dgrad read, bias read, wgrad read) the
    // write-back of dlogits and the weight-gradient read remain. The gradient buffers this touches are zeroed here
    // instead of at the top of backward(). MEASURED (B200, caption model alone, ms per step, N = 1280 / N = 640): un-chunked
    // 3.03 / 2.41; 4096-row chunks 3.26 / 2.47; 2048 rows 3.44 / 2.56; 1024 rows 3.86 / 2.77 -- the per-chunk GEMMs lose
    // more to wave quantisation, split-K atomics of the input gradient and 4 launches per chunk than the L2 hits return.
    const int rows_c = (chunk_env + kBM - 1) / kBM * kBM;
    VC_CUDA(cudaMemsetAsync(Gf, 0, (size_t)(n_adam + 64) * sizeof(float), s));
    VC_CUDA(cudaMemsetAsync(dOut, 0, (size_t)rows_all * Hd * sizeof(float), s));
    const float* b_o = pp(pidx("decoder/rnn_logits/bias"));
    float* db_o = gp(pidx("decoder/rnn_logits/bias"));
    for (long long r0 = 0; r0 < rows_all; r0 += rows_c) {
      const int rc = (int)std::min<long long>(rows_c, rows_all - r0);
      uint16_t* lg = (uint16_t*)logits + r0 * VP;
      {
        Operand A{(uint16_t*)Out + r0 * Hd, rc, Hd, Hd, false}, Bw{wo_t, V, Hd, Hd, false};
        ProfTag ptag("logits_fwd");
        VC_TRY(gemm_tma_rows(s, A, Bw, rc, V, Hd, lg, VP, b_o, 0, 256));
      }
      VC_TRY(ce_rows(s, logits, VP, in.cap_lbl, N, T, V, scal, ce_row, scal + 2, loss_scale, 1, (int)r0, rc));
      VC_TRY(colsum_bf16(s, lg, rc, V, VP, db_o));
      {
        Operand A{lg, rc, V, VP, false}, Bw{wo_nat, Hd, V, VP, false};
        EpiStore e{};
        e.out = dOut + r0 * Hd; e.ld = Hd; e.alpha = 1.f; e.atomic = 1;
        const int tiles = ((rc + kBM - 1) / kBM) * ((Hd + 255) / 256);
        ProfTag ptag("logits_dgrad");
        VC_TRY(gemm_store(s, A, nullptr, 0, Bw, rc, Hd, V, e, Hd % 256 == 0 ? 256 : 64, std::max(1, num_sms() / tiles)));
      }
    }
    vocab_bwd_done = true;
  } else {
    {
      Operand A{Out, rows_all, Hd, Hd, false}, Bw{wo_t, V, Hd, Hd, false};
      ProfTag ptag("logits_fwd");
      VC_TRY(gemm_tma_rows(s, A, Bw, T * N, V, Hd, logits, VP, pp(pidx("decoder/rnn_logits/bias")), 0, 256));
    }
    VC_TRY(ce_rows(s, logits, VP, in.cap_lbl, N, T, V, scal, ce_row, scal + 2, loss_scale, write_grad ? 1 : 0));
  }
  last_ann = annealing_coeff(cfg, in.global_step);
  have_forward = true;
  logits_intact = !write_grad;
  lastN = N;
  lastT = T;
  return VC_OK;
}

int Model::backward(const StepInputs& in, cudaStream_t s) {
  const int B = in.B, T = in.T, C = cfg.num_captions, N = B * C;
  const int E = cfg.embed_size, Hd = cfg.decoder_hidden, Z = cfg.latent_size, S = cfg.gen_z_samples, V = cfg.vocab_size,
            F = cfg.cnn_feature_size, K = cfg.num_clusters;
  if (!vocab_bwd_done) VC_CUDA(cudaMemsetAsync(Gf, 0, (size_t)(n_adam + 64) * sizeof(float), s));
  const long long rows = (long long)T * N;
  // logits layer: dOut = dlogits x W_o^T ; dW_o = Out^T x dlogits ; db_o = colsum(dlogits)
  {
    Operand A{logits, rows, V, VP, false}, Bw{wo_nat, Hd, V, VP, false};
    EpiStore e{};
    e.out = dOut; e.ld = Hd; e.alpha = 1.f;
    if (!vocab_bwd_done) {  // (the row-chunked forward has done the input gradient and the bias column sum already)
      ProfTag ptag("logits_dgrad");
      VC_TRY(gemm_store(s, A, nullptr, 0, Bw, (int)rows, Hd, V, e, Hd % 256 == 0 ? 256 : 64, 1));
    }
    // dW_o and db_o: on the side stream when it is on (beside the decoder's BPTT)
    cudaStream_t main_s = s;
    const bool on_side = side_begin(s);
    if (on_side) s = side;
    GridCap cap(on_side ? side_sm_cap : 0);
    Operand A2{Out, rows, Hd, Hd, true}, B2{logits, rows, V, VP, true};
    EpiStore e2{};
    e2.out = gp(pidx("decoder/rnn_logits/kernel")); e2.ld = V; e2.alpha = 1.f;
    {
      // [Hd, V] in 128 x 256 tiles is only (Hd/128) * ceil(V/256) = 180 tiles for 148 SMs (2 waves, the second a fifth
      // full); split the token contraction so the tile count lands just under a whole number of waves. The gradient
      // buffer was zeroed above, the partial sums meet in fp32 atomics.
      ProfTag ptag("logits_wgrad");
      const int tiles = ((Hd + 127) / 128) * ((V + 255) / 256);
      int splits = 1;
      if (rows >= 4096 && tiles < 4 * num_sms()) {
        double best = 0.0;
        for (int sp = 1; sp <= 8; ++sp) {
          const int t = tiles * sp;
          const double eff = (double)t / (double)(((t + num_sms() - 1) / num_sms()) * num_sms());
          if (eff > best + 0.02) { best = eff; splits = sp; }
        }
      }
      e2.atomic = splits > 1 ? 1 : 0;
      VC_TRY(gemm_store(s, A2, nullptr, 0, B2, Hd, V, (int)rows, e2, 256, splits));
    }
    if (!vocab_bwd_done) VC_TRY(colsum_bf16(s, logits, rows, V, VP, gp(pidx("decoder/rnn_logits/bias"))));
    // data parallel: each group of gradients is summed over the ranks as soon as its last producer is enqueued
    // (comm.cu); the vocabulary projection's 23 MB travel under the decoder's BPTT
    VC_TRY(grad_ready_params({pidx("decoder/rnn_logits/kernel"), pidx("decoder/rnn_logits/bias")}, s));
    s = main_s;
  }
  // decoder BPTT
  VC_CUDA(cudaMemsetAsync(dec.dh_carry, 0, (size_t)N * Hd * sizeof(float), s));
  VC_CUDA(cudaMemsetAsync(dec.dc_carry, 0, (size_t)N * Hd * sizeof(float), s));
  // Data parallel with an encoder: the posterior heads' gradient (71 MB under GMM / AG) is the biggest bucket and needs
  // only the decoder's recurrence and input gradient. The decoder's weight / bias gradients and its embedding scatter
  // (0.15 ms of work nothing waits for) run AFTER that bucket is on the wire, under its all-reduce, instead of before it.
  const bool dec_late = comm_world() > 1 && !cfg.no_encoder;
  auto decoder_leaf_grads = [&]() -> int {
    if (dec_late) VC_TRY(lstm_weight_grads(dec, N, T, s));
    VC_TRY(embed_scatter(s, dec.dX + (size_t)dec.pre * N * E, in.cap_in, gp(pidx("decoder/net/dec_embeddings")),
                         cfg.dec_keep_rate < 1.f ? in.rng.emb_keep_dev : nullptr, 1.f / cfg.dec_keep_rate, g_tail + 1, N, T,
                         E, V));
    return grad_ready_params({pidx("decoder/net/dec_embeddings")}, s);
  };
  VC_TRY(lstm_backward(dec, N, T, in.len, dOut, cfg.dec_lstm_drop < 1.f ? in.rng.out_keep_dev : nullptr, s, dec_late));
  if (!dec_late) VC_TRY(decoder_leaf_grads());
  if (!cfg.no_encoder) {
    const int He = cfg.encoder_hidden;
    const int SZ = S * Z;
    // z_rnn: dz = dzdec x W_z^T ; dW_z = z^T x dzdec ; db_z = colsum(dzdec)
    const float* dzdec = dec.dX + (size_t)(dec.pre - 1) * N * E;
    VC_TRY(cast_f32_bf16(s, dzdec, dzdec_h, N, E, E, E));
    {
      Operand A{dzdec_h, N, E, E, false}, Bw{z_nat, SZ, E, E, false};
      EpiStore e{};
      e.out = dz; e.ld = SZ; e.alpha = 1.f;
      {
        ProfTag ptag("z_rnn_dgrad");
        VC_TRY(gemm_store(s, A, nullptr, 0, Bw, N, SZ, E, e, 256, 1));
      }
      Operand A2{z, N, SZ, SZ, true}, B2{dzdec_h, N, E, E, true};
      EpiStore e2{};
      e2.out = gp(pidx("decoder/net/z_rnn/kernel")); e2.ld = E; e2.alpha = 1.f;
      {
        ProfTag ptag("z_rnn_wgrad");
        VC_TRY(gemm_store(s, A2, nullptr, 0, B2, SZ, E, N, e2, E % 256 == 0 ? 256 : 64, 1));
      }
      VC_TRY(colsum_bf16(s, dzdec_h, N, E, E, gp(pidx("decoder/net/z_rnn/bias"))));
    }
    // reparameterisation + KL -> head gradients. lower_bound = rec + ann * KL / 10 (main.py:172-174);
    // Normal/GMM KL is a batch mean, AG a per-row vector whose sum is differentiated (Q2).
    const float ann = annealing_coeff(cfg, in.global_step);
    const float kl_scale = cfg.prior == VC_PRIOR_AG ? ann / 10.f : ann / (10.f * N);
    VC_TRY(dz_reduce(s, dz, in.rng.eps_dev, in.rng.seed, (unsigned long long)in.global_step, sd, dkl_dmu, dkl_dsd, kl_scale,
                     nullptr, heads_cols, ZP, dmu_t, dsd_t, S, N, Z));
    if (cfg.prior != VC_PRIOR_NORMAL) VC_CUDA(cudaMemsetAsync(dheads, 0, (size_t)N * heads_cols * 2, s));
    VC_TRY(heads_mix_bwd(s, dmu_t, dsd_t, heads_f, heads_cols, ZP, cfg.prior, in.c_v, K, last_pick, sd, dheads, N, Z));
    const uint16_t* hT = (uint16_t*)enc.Hs + (size_t)(enc.pre + T) * N * He;
    {
      ProfTag ptag("heads_bwd");
      Operand A{dheads, N, heads_cols, heads_cols, false}, Bw{heads_nat, He, heads_cols, heads_cols, false};
      EpiStore e{};
      e.out = enc.dh_carry; e.ld = He; e.alpha = 1.f;
      if (heads_cols > 4096) {  // long contraction, few output tiles: split-K
        VC_CUDA(cudaMemsetAsync(enc.dh_carry, 0, (size_t)N * He * sizeof(float), s));
        e.atomic = 1;
        const int tiles = ((N + 127) / 128) * ((He + 63) / 64);
        VC_TRY(gemm_store(s, A, nullptr, 0, Bw, N, He, heads_cols, e, 64, std::max(1, num_sms() / tiles)));
      } else {
        VC_TRY(gemm_store(s, A, nullptr, 0, Bw, N, He, heads_cols, e, 64, 1));
      }
      // dW[He, heads_cols] = h_T^T x dheads, written straight into the packed gradient block
      Operand A2{hT, N, He, He, true}, B2{dheads, N, heads_cols, heads_cols, true};
      EpiStore e2{};
      e2.out = gp(p_heads_w); e2.ld = heads_cols; e2.alpha = 1.f;
      VC_TRY(gemm_store(s, A2, nullptr, 0, B2, He, heads_cols, N, e2, heads_cols >= 1024 ? 256 : 64, 1));
      VC_TRY(colsum_bf16(s, dheads, N, heads_cols, heads_cols, gp(p_heads_b)));
    }
    VC_TRY(grad_ready_params({p_heads_w, p_heads_b, pidx("decoder/net/z_rnn/kernel"), pidx("decoder/net/z_rnn/bias")}, s));
    if (dec_late) VC_TRY(decoder_leaf_grads());
    VC_CUDA(cudaMemsetAsync(enc.dc_carry, 0, (size_t)N * He * sizeof(float), s));
    VC_TRY(lstm_backward(enc, N, T, in.len, nullptr, nullptr, s));
    VC_TRY(embed_scatter(s, enc.dX + (size_t)enc.pre * N * E, in.cap_lbl, gp(pidx("encoder/enc_embeddings")), nullptr, 1.f,
                         g_tail + 0, N, T, E, V));
    // both embedding-slice norms are final now: the encoder table and the tail leave here, so that only the two small
    // projections (4 MB) are left for the last, fully exposed bucket
    VC_TRY(grad_ready_params({pidx("encoder/enc_embeddings"), -2}, s));
  }
  // imf_emb: gradient of the tiled projection sums over the C captions of an image (Q7)
  VC_TRY(tile_reduce(s, dec.dX, cfg.no_encoder ? nullptr : enc.dX, nullptr, dimf_h, B, C, E));
  {
    Operand A{feats_h, B, F, F, true}, Bw{dimf_h, B, E, E, true};
    EpiStore e{};
    e.out = gp(pidx("imf_emb/kernel")); e.ld = E; e.alpha = 1.f;
    ProfTag ptag("imf_emb_bwd");
    VC_TRY(gemm_store(s, A, nullptr, 0, Bw, F, E, B, e, E % 256 == 0 ? 256 : 64, 1));
    VC_TRY(colsum_bf16(s, dimf_h, B, E, E, gp(pidx("imf_emb/bias"))));
  }
  if (cfg.use_c_v) {
    VC_TRY(tile_reduce(s, dec.dX + (size_t)N * E, cfg.no_encoder ? nullptr : enc.dX + (size_t)N * E, nullptr, dcv_h, N, 1, E));
    Operand A{cv_h, N, K, KP, true}, Bw{dcv_h, N, E, E, true};
    EpiStore e{};
    e.out = gp(pidx("cv_emb/kernel")); e.ld = E; e.alpha = 1.f;
    ProfTag ptag("cv_emb_bwd");
    VC_TRY(gemm_store(s, A, nullptr, 0, Bw, K, E, N, e, E % 256 == 0 ? 256 : 64, 1));
    VC_TRY(colsum_bf16(s, dcv_h, N, E, E, gp(pidx("cv_emb/bias"))));
  }
  if (cfg.fine_tune) {
    // d lower_bound / d fc2 = dimf x W_imf^T, then down the VGG16 (ops/optimizers.py:49-52)
    Operand A{dimf_h, B, E, E, false}, Bw{imf_nat, F, E, E, false};
    EpiStore e{};
    e.out = dfeats_f; e.ld = F; e.alpha = 1.f;
    {
      ProfTag ptag("imf_emb_dgrad");
      VC_TRY(gemm_store(s, A, nullptr, 0, Bw, B, F, E, e, 128, 1));
    }
    VC_TRY(vgg_backward(dfeats_f, B, s));
  }
  if (comm != nullptr) {
    // last bucket: the two small projections (and, without an encoder, the 64-float tail whose first two entries are
    // the per-token embedding-slice squared norms of this rank; Q4: the norm is over every tower's slices)
    // one NCCL group (one launch) for all of it; params are listed in buffer order so neighbours merge into one range
    VC_TRY(grad_ready_params({pidx("imf_emb/kernel"), pidx("imf_emb/bias"), cfg.use_c_v ? pidx("cv_emb/kernel") : -1,
                              cfg.use_c_v ? pidx("cv_emb/bias") : -1, cfg.no_encoder ? -2 : -1}, s));
  }
  VC_TRY(side_join(s));  // the optimiser (or the caller's own all-reduce) follows on `s`
  // squared norm of the dense (non-embedding) gradients -> tail[2] (apply); embedding slices are in tail[0..1] (Q4)
  return VC_OK;
}

int Model::apply(float grad_scale, cudaStream_t s) {
  // ops/optimizers.py:13-40: clip_by_global_norm(5.0) then Adam(lr, beta1=0.8); TF Adam form (Q5)
  VC_TRY(comm_join(s));  // data parallel: every bucket of the gradient all-reduce has landed
  VC_CUDA(cudaMemsetAsync(g_tail + 2, 0, sizeof(float), s));
  VC_TRY(sumsq(s, Gf, n_dense, g_tail + 2));
  adam_t += 1;
  const double b1 = 0.8, b2 = 0.999;
  // Adam: constant rate with TF's bias correction (Q5). SGD / Momentum(0.9): the rate halves every lr_decay_steps steps
  // (tf.train.exponential_decay, staircase, read before global_step is incremented; ops/optimizers.py:24-46)
  auto rate = [&](int kind, float base) -> float {
    if (kind == VC_OPT_ADAM) return (float)(base * std::sqrt(1.0 - std::pow(b2, (double)adam_t)) / (1.0 - std::pow(b1, (double)adam_t)));
    const int64_t period = cfg.lr_decay_steps > 0 ? cfg.lr_decay_steps : 1;
    return (float)(base * std::pow(0.5, (double)((adam_t - 1) / period)));
  };
  const float lr_t = rate(cfg.optimizer, cfg.learning_rate);
  VC_TRY(adam_step(s, Pf, Gf, Mf, Vf, n_adam, g_tail, 3, cfg.clip_norm, grad_scale, lr_t,
                   cfg.optimizer == VC_OPT_MOMENTUM ? 0.9f : (float)b1, (float)b2, 1e-8f, scal + 7, 0.f, cfg.optimizer));
  if (cfg.fine_tune) {
    // ops/optimizers.py:49-82: Adam(cnn_lr, beta1=0.8) on the 30 cnn/ variables, no clipping; the regulariser
    // gradient weight_decay * w (Q11) is added inside the update
    const float lr_c = rate(cfg.cnn_optimizer, cfg.cnn_lr);
    const int64_t c0 = n_total - n_cnn;
    VC_TRY(adam_step(s, Pf + c0, Gf + c0, Mf + c0, Vf + c0, n_cnn, nullptr, 0, 0.f, grad_scale, lr_c,
                     cfg.cnn_optimizer == VC_OPT_MOMENTUM ? 0.9f : (float)b1, (float)b2, 1e-8f, nullptr, cfg.weight_decay,
                     cfg.cnn_optimizer));
    vgg_shadows_dirty = true;
  }
  shadows_dirty = true;
  VC_TRY(refresh_shadows(s));
  return VC_OK;
}

void Model::fill_out(vc_step_out* out, const float* h, int n, float ann) const {
  const float cnt = h[1] > 0.f ? h[1] : 1.f;
  out->rec_loss = h[0] / cnt;
  if (cfg.fine_tune) out->rec_loss += 0.5f * cfg.weight_decay * h[8];
  out->n_tokens = h[1];
  out->kld = cfg.no_encoder ? 0.f : h[3] / (float)n;
  out->global_norm = h[7];
  out->annealing = ann;
  out->lower_bound = cfg.no_encoder ? out->rec_loss : out->rec_loss + ann * out->kld / 10.f;
}

int Model::fetch(vc_step_out* out, cudaStream_t s) {
  if (out == nullptr) return VC_OK;
  VC_CUDA(cudaMemcpyAsync(host_scal, scal, 16 * sizeof(float), cudaMemcpyDeviceToHost, s));
  VC_CUDA(cudaStreamSynchronize(s));
  fill_out(out, host_scal, lastN, last_ann);
  return VC_OK;
}

// Deferred form: the copy is queued behind the step on `s`, the host comes back for it one step later.
int Model::result_queue(cudaStream_t s) {
  if (!have_forward) return set_error(VC_E_STATE, "no step has run on this handle");
  if (pending_count == 2) return set_error(VC_E_STATE, "two step results are already outstanding: call vc_step_result first");
  PendingResult& p = pending[(pending_head + pending_count) & 1];
  if (p.ev == nullptr) VC_CUDA(cudaEventCreateWithFlags(&p.ev, cudaEventDisableTiming));
  float* slot = host_scal + 16 * (1 + ((pending_head + pending_count) & 1));  // host_scal holds 64 floats: [fetch | slot 0 | slot 1]
  VC_CUDA(cudaMemcpyAsync(slot, scal, 16 * sizeof(float), cudaMemcpyDeviceToHost, s));
  VC_CUDA(cudaEventRecord(p.ev, s));
  p.n = lastN;
  p.ann = last_ann;
  ++pending_count;
  return VC_OK;
}

int Model::result_pop(vc_step_out* out) {
  if (pending_count == 0) return set_error(VC_E_STATE, "no step result is queued");
  if (out == nullptr) return set_error(VC_E_ARG, "vc_step_result: null out");
  PendingResult& p = pending[pending_head];
  VC_CUDA(cudaEventSynchronize(p.ev));
  fill_out(out, host_scal + 16 * (1 + pending_head), p.n, p.ann);
  pending_head ^= 1;
  --pending_count;
  return VC_OK;
}

int Model::forward_debug(float* logits_host, float* mu_host, float* std_host, float* z_host, float* kl_host, float* ce_host) {
  if (!have_forward) return set_error(VC_E_STATE, "no forward pass has run on this handle");
  VC_CUDA(cudaDeviceSynchronize());
  const int N = lastN, T = lastT, V = cfg.vocab_size, Z = cfg.latent_size, S = cfg.gen_z_samples;
  if (logits_host != nullptr) {
    if (!logits_intact)
      return set_error(VC_E_STATE, "logits were overwritten by the backward pass; run vc_eval_step first");
    float* tmp = nullptr;
    VC_CUDA(cudaMalloc((void**)&tmp, (size_t)N * T * V * sizeof(float)));
    int st = logits_to_ref(0, logits, VP, tmp, N, T, V);
    if (st == VC_OK && cudaMemcpy(logits_host, tmp, (size_t)N * T * V * sizeof(float), cudaMemcpyDeviceToHost) != cudaSuccess)
      st = set_error(VC_E_CUDA, "logits copy failed");
    cudaFree(tmp);
    VC_TRY(st);
  }
  if (ce_host != nullptr) VC_CUDA(cudaMemcpy(ce_host, ce_row, (size_t)N * T * sizeof(float), cudaMemcpyDeviceToHost));
  if (cfg.no_encoder) return VC_OK;
  if (mu_host != nullptr) VC_CUDA(cudaMemcpy(mu_host, mu, (size_t)N * Z * sizeof(float), cudaMemcpyDeviceToHost));
  if (std_host != nullptr) VC_CUDA(cudaMemcpy(std_host, sd, (size_t)N * Z * sizeof(float), cudaMemcpyDeviceToHost));
  if (kl_host != nullptr) VC_CUDA(cudaMemcpy(kl_host, kl_row, (size_t)N * sizeof(float), cudaMemcpyDeviceToHost));
  if (z_host != nullptr) {
    float* tmp = nullptr;
    const size_t n = (size_t)S * N * Z;
    VC_CUDA(cudaMalloc((void**)&tmp, n * sizeof(float)));
    int st = bf16_to_f32(0, z, tmp, 1, (int)n, n, n);
    if (st == VC_OK && cudaMemcpy(z_host, tmp, n * sizeof(float), cudaMemcpyDeviceToHost) != cudaSuccess)
      st = set_error(VC_E_CUDA, "z copy failed");
    cudaFree(tmp);
    VC_TRY(st);
  }
  return VC_OK;
}

}  // namespace vc
