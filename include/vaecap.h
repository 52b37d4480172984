/* libvaecap -- C ABI of the B200-native CVAE captioning hot path.
 *
 * The reference (yiyang92/vae_captioning) has no FFI: its seam is the per-step TensorFlow
 * `sess.run(fetches, feed_dict)` call plus the checkpoint variable names. Each entry point below
 * replaces one such call site (cited as file:line of the reference); INTEGRATION.md shows the
 * Python binding a maintainer of the reference would add.
 *
 * Conventions
 *  - every function returns 0 on success or a negative VC_E_* code; vc_last_error() gives the message
 *    (thread-local); nothing throws or aborts across this boundary.
 *  - pointers named *_dev are device pointers, *_host are host pointers (pinned memory makes the
 *    copies asynchronous). The caller owns every I/O buffer; the handle owns parameters, optimiser
 *    state and the activation workspace (sized at vc_create from max_batch / max_len).
 *  - `stream` is a cudaStream_t passed as void*. Calls on one handle are not re-entrant.
 *  - one process drives ONE GPU (one rank per GPU under torchrun, as the data-parallel path does): per-kernel launch
 *    attributes and the SM count are cached per process, so handles on different devices of one process are unsupported.
 *  - all floating-point I/O is fp32 in TensorFlow layouts (dense kernels [in,out], conv HWIO,
 *    LSTM kernel [x;h] x [i|j|f|o]); token ids and lengths are int32.
 */
#ifndef VAECAP_H_
#define VAECAP_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { VC_OK = 0, VC_E_ARG = -1, VC_E_SHAPE = -2, VC_E_CUDA = -3, VC_E_NCCL = -4, VC_E_STATE = -5, VC_E_NOMEM = -6 };
enum { VC_PRIOR_NORMAL = 0, VC_PRIOR_GMM = 1, VC_PRIOR_AG = 2 };
enum { VC_OPT_ADAM = 0, VC_OPT_SGD = 1, VC_OPT_MOMENTUM = 2 };

/* Mirrors the fields of the reference's Parameters (utils/parameters.py:3-66) that shape the graph
 * built in main.py:43-191. */
typedef struct vc_config {
  int32_t vocab_size;       /* data.dictionary.vocab_size, main.py:92 */
  int32_t embed_size;       /* --embed_dim   (multiple of 64) */
  int32_t encoder_hidden;   /* --enc_hid     (multiple of 64) */
  int32_t decoder_hidden;   /* --dec_hid     (multiple of 64) */
  int32_t latent_size;      /* --latent */
  int32_t gen_z_samples;    /* --gen_z_samples */
  int32_t num_clusters;     /* 90 */
  int32_t num_captions;     /* captions per image fed per step (1..5) */
  int32_t cnn_feature_size; /* 4096 */
  int32_t prior;            /* VC_PRIOR_* (--prior) */
  int32_t use_c_v;          /* --c_v */
  int32_t no_encoder;       /* --no_encoder */
  int32_t fine_tune;        /* --fine_tune: images in, VGG16 trained */
  int32_t restore;          /* --restore: annealing pinned to 1 (main.py:163-164) */
  int32_t with_cnn;         /* allocate VGG16 parameters (needed for fine_tune and vc_vgg_forward) */
  int32_t max_batch;        /* images per step (B) */
  int32_t max_len;          /* padded caption length (T) */
  int32_t optimizer;        /* --optimizer: VC_OPT_ADAM / VC_OPT_SGD / VC_OPT_MOMENTUM (ops/optimizers.py:33-46) */
  int32_t cnn_optimizer;    /* Parameters.cnn_optimizer, same values (ops/optimizers.py:68-81) */
  int32_t lr_decay_steps;   /* int(num_ex_per_epoch / (batch_size + 0.001) * num_epochs_per_decay): SGD / Momentum halve
                               their rate every that many steps (staircase, ops/optimizers.py:24-31); Adam ignores it */
  float dec_keep_rate;      /* --dec_drop */
  float dec_lstm_drop;      /* --dec_lstm_drop */
  float cnn_dropout;
  float weight_decay;
  float learning_rate;      /* --lr */
  float cnn_lr;
  float clip_norm;          /* lstm_clip_by_norm = 5.0 */
  float ann_param;          /* --ann_param */
  float std;                /* --std (generation prior) */
  float temperature;
} vc_config;

/* Explicit randomness for one step. NULL pointers select the in-kernel Philox generator seeded
 * with (seed, step). Explicit tensors are the parity mode (the oracle feeds the same values). */
typedef struct vc_rng {
  uint64_t seed;
  const float* eps_dev;           /* [S, N, Z]   N(0,1) draws of zs.Normal, encoder.py:108-109 */
  const float* emb_keep_dev;      /* [N, T, E]   0/1 keep mask of tf.nn.dropout, decoder.py:85-87 */
  const float* out_keep_dev;      /* [N, T, H]   0/1 keep mask of DropoutWrapper, rnn_model.py:45-46 */
  const int32_t* gmm_cluster_dev; /* [N]         tf.multinomial pick, encoder.py:72 */
  const float* cnn_keep_dev;      /* [2, B, 4096] 0/1 keep masks of the fc1 / fc2 dropout, image_embeddings.py:225-237 */
} vc_rng;

/* Fetches of main.py:241-244. */
typedef struct vc_step_out {
  float kld;         /* np.mean(kl) */
  float rec_loss;
  float lower_bound; /* np.mean(lb) */
  float annealing;
  float global_norm; /* tf.clip_by_global_norm's norm (ops/optimizers.py:15-16) */
  float n_tokens;    /* sum of the loss mask */
} vc_step_out;

typedef struct vc_handle vc_handle;

const char* vc_last_error(void);
int vc_abi_version(void);

/* Measurement hooks (no reference counterpart; the reference has no profiling):
 * vc_launch_count: kernels this library has launched in this process so far.
 * vc_profile_enable(1) starts bracketing every launch with CUDA events on its stream; vc_profile_collect
 * synchronises and returns per-kernel-family totals: comma-joined names, milliseconds and launch counts. */
unsigned long long vc_launch_count(void);
int vc_profile_enable(int on);
int vc_profile_collect(char* names, int names_cap, float* ms, int* counts, int cap);

/* Graph construction, main.py:43-191 (+ tf.global_variables_initializer, main.py:196). Parameters start at zero;
 * load them with vc_param_set. */
int vc_create(const vc_config* cfg, int device, vc_handle** out);
int vc_destroy(vc_handle* h);

/* tf.train.Saver surface (main.py:186-191, 211, 288; utils/image_embeddings.py:240-246): variables by TF name. */
int vc_num_params(vc_handle* h);
int vc_param_info(vc_handle* h, int index, const char** name, int32_t* ndim, int64_t* shape /*[4]*/, int32_t* trainable);
int vc_param_get(vc_handle* h, const char* name, float* dst_host);
int vc_param_set(vc_handle* h, const char* name, const float* src_host);
int vc_grad_get(vc_handle* h, const char* name, float* dst_host); /* tf.gradients of the last step (debug tap) */
/* init_clusters (utils/vae_utils.py:6-31, main.py:127/137): the [num_clusters, latent] constant cluster means of the
 * AG prior (the reference persists them in ./pickles/cluster_means.pickle). Zero until set. */
int vc_set_cluster_means(vc_handle* h, const float* src_host);

/* One training sess.run (main.py:229-244): feed {image_f_inputs, ann_inputs_enc, ann_inputs_dec, ann_lengths,
 * anneal, c_i} -> fetch [kld, rec_loss, lower_bound, optimize, optimize_cnn, annealing].
 * feats_host: fp32 [B, 4096] (or [B,224,224,3] images when fine_tune). cap_lbl/cap_in: int32 [B*C, T]
 * (row n = b*C + c, utils/caption_utils.py:16-21). c_v: fp32 [B*C, 90] or NULL. global_step feeds `anneal`.
 * All inputs are HOST pointers; the call copies them to the device on `stream`. out may be NULL (no sync). */
int vc_train_step(vc_handle* h, const float* feats_host, const int32_t* cap_lbl_host, const int32_t* cap_in_host,
                  const int32_t* len_host, const float* c_v_host, int B, int T, int64_t global_step, const vc_rng* rng,
                  vc_step_out* out, void* stream);
/* Same step fed with raw images (fp32 [B,224,224,3] RGB 0..255, host): the VGG16 forward that
 * Data.extract_features_from_dir runs offline (utils/data.py:86-130) happens on the device inside the call; the
 * CNN is not trained (use fine_tune for that). Needs with_cnn. */
int vc_train_step_images(vc_handle* h, const float* images_host, const int32_t* cap_lbl_host, const int32_t* cap_in_host,
                         const int32_t* len_host, const float* c_v_host, int B, int T, int64_t global_step,
                         const vc_rng* rng, vc_step_out* out, void* stream);
/* Same with uint8 pixels [B,224,224,3] (RGB 0..255), the form the reference's HDF5 image store keeps them in
 * (utils/batch_gen.py:278-294, preprocess.py:26-38): a quarter of the host-to-device bytes; the conversion and the
 * mean subtraction happen in the first device kernel. Works on fine_tune handles too (the images are then the feed). */
int vc_train_step_images_u8(vc_handle* h, const uint8_t* images_host, const int32_t* cap_lbl_host, const int32_t* cap_in_host,
                            const int32_t* len_host, const float* c_v_host, int B, int T, int64_t global_step,
                            const vc_rng* rng, vc_step_out* out, void* stream);
/* Double-buffered feed for the same step (the reference feeds one numpy batch per sess.run, main.py:229-244; here the
 * H2D copy of batch i+1 overlaps the compute of batch i). vc_stage_batch copies one step's HOST buffers into staging
 * slot 0 or 1 on `copy_stream` (asynchronous when the host memory is pinned) and records an event;
 * vc_train_step_staged runs the train step on `stream` from that slot, waiting for the copy on the device. A slot may
 * be refilled as soon as the step that consumes it has been enqueued (the copy waits for that step's last read).
 * kind: 0 = the vc_train_step feed (fc2 features fp32 [B, F]; fp32 images when fine_tune), 1 = fp32 images
 * [B,224,224,3], 2 = uint8 images (both: on-device VGG16 forward, or the fine-tune feed). Errors as vc_train_step. */
int vc_stage_batch(vc_handle* h, int slot, const void* feats_or_images_host, int kind, const int32_t* cap_lbl_host,
                   const int32_t* cap_in_host, const int32_t* len_host, const float* c_v_host, int B, int T, void* copy_stream);
int vc_train_step_staged(vc_handle* h, int slot, int64_t global_step, const vc_rng* rng, vc_step_out* out, void* stream);
/* Deferred read-back of a step's scalars (the reference's sess.run returns them synchronously, main.py:229-244; waiting
 * for step i before enqueueing step i+1 leaves the GPU idle for the host's launch time). After a train step called with
 * out == NULL, vc_step_result_queue enqueues the 64-byte device-to-host copy of ITS scalars on `stream` and returns at
 * once; vc_step_result waits for the OLDEST queued copy and fills `out`. At most two results may be outstanding
 * (VC_E_STATE beyond that, and when nothing is queued). */
int vc_step_result_queue(vc_handle* h, void* stream);
int vc_step_result(vc_handle* h, vc_step_out* out);
/* Data-parallel form of the staged step: forward + backward from the slot, gradients left in vc_grad_buffer for the
 * caller's all-reduce; finish with vc_apply_gradients(1/world). */
int vc_forward_backward_staged(vc_handle* h, int slot, int64_t global_step, const vc_rng* rng, void* stream);

/* vc_train_step with inputs already resident in device memory. */
int vc_train_step_dev(vc_handle* h, const float* feats_dev, const int32_t* cap_lbl_dev, const int32_t* cap_in_dev,
                      const int32_t* len_dev, const float* c_v_dev, int B, int T, int64_t global_step,
                      const vc_rng* rng, vc_step_out* out, void* stream);
/* Split form for data parallelism: forward+backward fills the flat gradient buffer, the caller all-reduces it
 * (vc_grad_buffer), then vc_apply_gradients clips and runs Adam. grad_scale multiplies every gradient (1/world). */
int vc_forward_backward_dev(vc_handle* h, const float* feats_dev, const int32_t* cap_lbl_dev, const int32_t* cap_in_dev,
                            const int32_t* len_dev, const float* c_v_dev, int B, int T, int64_t global_step,
                            const vc_rng* rng, void* stream);
int vc_grad_buffer(vc_handle* h, float** dev_ptr, int64_t* count);
int vc_apply_gradients(vc_handle* h, float grad_scale, vc_step_out* out, void* stream);

/* ---- data parallelism (no reference counterpart: the reference is single-device, utils/parameters.py:163-164; the
 * step shards by minibatch). One process per GPU. Rank 0 obtains a 128-byte NCCL unique id with
 * vc_comm_unique_id and hands it to the other ranks by any means (the Python side broadcasts it with
 * torch.distributed); every rank then calls vc_comm_init(handle, id, rank, world) -- collective, creates the handle's
 * communicator over NVLink / NVSwitch. From then on EVERY train-step entry point above (vc_train_step*,
 * vc_train_step_staged, vc_train_step_dev) is the data-parallel step: the backward pass hands each group of gradients
 * to an in-place sum all-reduce on a communication stream as soon as its last producer kernel is enqueued (vocabulary
 * projection first, encoder / CNN last), the optimiser waits for the last bucket and applies the MEAN of the towers
 * (grad scale 1/world; the global norm of Q4 is over every tower's embedding slices, whose squared norms travel in the
 * same buffer). With a communicator attached vc_forward_backward_* already reduce: callers of the split form must not
 * all-reduce vc_grad_buffer again and pass 1/world to vc_apply_gradients. libnccl is bound with dlopen at the first
 * call (VC_E_NCCL if it cannot be found); without vc_comm_init nothing here is touched.
 * vc_allreduce_gradients: the un-overlapped form for callers that fill vc_grad_buffer themselves.
 * vc_comm_stats: after a step, wall time on the communication stream from the first to the last bucket of that step
 * (milliseconds), the bytes it summed and the number of buckets (synchronises the communication stream). */
int vc_comm_unique_id(void* id128);
int vc_comm_init(vc_handle* h, const void* id128, int rank, int world);
int vc_allreduce_gradients(vc_handle* h, void* stream);
int vc_comm_stats(vc_handle* h, float* ms, long long* bytes, int* calls);
/* Measurement switch: 1 (default) bucketed + overlapped, 2 one all-reduce of the whole buffer behind the backward pass
 * (the baseline the overlap is compared with), 0 no reduction at all (the ranks diverge: timing only). */
int vc_comm_set_mode(vc_handle* h, int mode);

/* validate(): sess.run([rec_loss]) on the training graph (main.py:262-284, dropout stays on, Q12). */
int vc_eval_step(vc_handle* h, const float* feats_host, const int32_t* cap_lbl_host, const int32_t* cap_in_host,
                 const int32_t* len_host, const float* c_v_host, int B, int T, const vc_rng* rng, vc_step_out* out,
                 void* stream);

/* Debug taps of the last forward (x_logits main.py:150, qz.distribution.mean/std main.py:122-124).
 * Any pointer may be NULL. logits_host: fp32 [N*T, V] in the reference's row order (n*T + t). */
int vc_forward_debug(vc_handle* h, float* logits_host, float* mu_host, float* std_host, float* z_host,
                     float* kl_rows_host, float* ce_rows_host);

/* Data.extract_features_from_dir's sess.run(features, {input_img}) (utils/data.py:120-125), batched:
 * images fp32 [B,224,224,3] RGB 0..255 (host) -> fc2 fp32 [B,4096] (host). */
int vc_vgg_forward(vc_handle* h, const float* images_host, float* fc2_host, int B, void* stream);
int vc_vgg_forward_dev(vc_handle* h, const float* images_dev, float* fc2_dev, int B, void* stream);
int vc_vgg_forward_u8(vc_handle* h, const uint8_t* images_host, float* fc2_host, int B, void* stream); /* uint8 pixels */
/* Debug taps: vc_vgg_keep_activations(h, 1) makes later forwards materialise every conv output instead of fusing
 * the 2x2 max-pools into the conv epilogues; vc_vgg_activation then returns the NHWC activation of a layer
 * ("conv1_1".."conv5_3", post-ReLU; "pool1".."pool5") of the last forward as fp32; "fc1" / "fc2" return the [B, 4096]
 * outputs of the two dense layers (post-ReLU, post-dropout when fine-tuning). */
int vc_vgg_keep_activations(vc_handle* h, int on);
int vc_vgg_activation(vc_handle* h, const char* layer, float* dst_host);

/* ---- generation (vae_model/decoder.py:145-320, ops/inference.py:16-50). The reference issues one
 * sess.run([sample, out_state], feed) per token per beam at batch 1; these calls run the whole loop for a batch of B
 * images on the device. feats_host fp32 [B, 4096] (features are NOT tiled in generation, main.py:84), c_v_host fp32
 * [B, 90] or NULL. rng->eps_dev (optional, device) fp32 [B, S, Z]: the N(0,1) draws of zs.Normal('z', z_mean, std)
 * (decoder.py:72-74); NULL -> Philox(rng->seed). bos / eos: data_dict.word2idx['<BOS>'] / ['<EOS>'].
 *
 * vc_decode_greedy = Decoder.online_inference (decoder.py:145-201): mode 0 'greedy' (argmax; temperature is a no-op,
 * SURVEY Q10), mode 1 'sample' (tf.multinomial(logits / temperature)). out_tokens_host int32 [B, max_len] (zero
 * padded, the generated ids = cap_raw), out_len_host [B].
 * vc_decode_beam = Decoder.beam_search (decoder.py:203-320) with utils/top_n.py container semantics (heap order,
 * tie handling), <BOS> fed twice (Q9), p < 1e-12 skipped, score / len^len_norm only for completed captions, never
 * mixing complete and partial. Per image the beams sorted by score: out_tokens_host int32 [B, beam, max_len]
 * (sentence incl. <BOS>/<EOS>, zero padded), out_len_host / out_score_host [B, beam], out_n_host [B] beams returned
 * (beams[0] is the caption; all of them = ret_beams). */
int vc_decode_greedy(vc_handle* h, const float* feats_host, const float* c_v_host, int B, int max_len, int mode,
                     const vc_rng* rng, int bos, int eos, int32_t* out_tokens_host, int32_t* out_len_host, void* stream);
int vc_decode_beam(vc_handle* h, const float* feats_host, const float* c_v_host, int B, int beam, int max_len, float len_norm,
                   const vc_rng* rng, int bos, int eos, int32_t* out_tokens_host, int32_t* out_len_host,
                   float* out_score_host, int32_t* out_n_host, void* stream);
/* The reference's own granularity for callers that keep their Python loop: vc_decode_begin computes the state the graph
 * uses when `initial_state` is left at its placeholder default (utils/rnn_model.py:7-21, decoder.py:96-114);
 * vc_decode_step is one sess.run([sample, out_state], {captions: tokens[M,1], in_state: current}) for the M rows:
 * probs_host fp32 [M, V] = tf.nn.softmax(x_logits) (nullable), state advanced in place;
 * vc_decode_state_get / _set read or feed (c, h) fp32 [M, H] each. */
int vc_decode_begin(vc_handle* h, const float* feats_host, const float* c_v_host, int B, const vc_rng* rng, void* stream);
int vc_decode_step(vc_handle* h, const int32_t* tokens_host, int M, float* probs_host, void* stream);
int vc_decode_state_get(vc_handle* h, float* c_host, float* h_host, void* stream);
int vc_decode_state_set(vc_handle* h, const float* c_host, const float* h_host, void* stream);
/* Host-only (no GPU): the beam bookkeeping above driven by a callback `step(user, token, state_in (-1 = initial),
 * probs_out[V]) -> state id`; pins the container semantics against the reference's Decoder.beam_search. Outputs as
 * vc_decode_beam for one image. */
int vc_beam_search_host(int (*step)(void* user, int token, int state_in, float* probs_out), void* user, int V, int beam,
                        int max_len, int bos, int eos, float len_norm, int32_t* out_tokens, int32_t* out_len,
                        float* out_score, int32_t* out_n);

/* Host-only (no GPU): CRC-32C continuing from `crc` (0 to start). The checkpoint writer/reader of the Python host side
 * (tf.train.Saver's V2 tensor-bundle files, main.py:186-191, 211, 288) checksums every tensor and index block with it. */
uint32_t vc_crc32c(uint32_t crc, const void* data, size_t n);

#ifdef __cplusplus
}
#endif
#endif /* VAECAP_H_ */
