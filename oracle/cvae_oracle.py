"""CPU ORACLE (test infrastructure, not product code) -- PARITY UNPINNED BY THE REFERENCE.

A torch-CPU restatement of the graph that yiyang92/vae_captioning builds with TensorFlow 1.x +
zhusuan 0.3 for its ELBO train step and decode loop. TensorFlow/zhusuan cannot be installed in
this environment (Python 3.12, no network) and the reference ships no tests, golden vectors or
checkpoints, so this file follows the reference *source* line by line (citations below) plus the
published semantics of the TF/zhusuan ops it calls. The pure-Python
pieces of the reference that CAN be imported here (utils/top_n.py, utils/caption_utils.py,
utils/parameters.py) are pinned by tests/golden/*.json generated from the reference itself
(tests/golden/make_reference_fixtures.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module. The product path (vae_captioning_b200/) never does.

All randomness is an explicit input (eps, dropout keep-masks, GMM cluster picks): the reference
never seeds TF, so it is not self-reproducible either.

dtype: float64 by default ("exact" mode); float32 for the timed CPU baseline. `emulate_bf16`
rounds tensors to bfloat16 (straight-through gradient) at exactly the points where the CUDA
path stores or feeds bf16, so forward parity can be checked far more tightly than the stated
bf16-vs-fp32 tolerance.
"""
import math

import numpy as np
import torch

NUM_CLUSTERS = 90
C_SIGMA = 0.1  # utils/vae_utils.py:10


class Config:
    """The fields of the reference's Parameters (utils/parameters.py:3-66) that shape the graph."""

    def __init__(self, **kw):
        self.vocab_size = 11313
        self.embed_size = 256
        self.encoder_hidden = 512
        self.decoder_hidden = 512
        self.latent_size = 150
        self.gen_z_samples = 100
        self.num_clusters = NUM_CLUSTERS
        self.num_captions = 5
        self.cnn_feature_size = 4096
        self.prior = "Normal"
        self.use_c_v = False
        self.no_encoder = False
        self.fine_tune = False
        self.restore = False
        self.dec_keep_rate = 1.0
        self.dec_lstm_drop = 1.0
        self.cnn_dropout = 0.5
        self.weight_decay = 0.00004
        self.learning_rate = 0.0005
        self.cnn_lr = 0.00001
        self.lstm_clip_by_norm = 5.0
        self.optimizer = "Adam"        # --optimizer {SGD, Adam, Momentum} (utils/parameters.py:34)
        self.cnn_optimizer = "Adam"
        self.batch_size = 32           # with the two below: the staircase decay period of SGD / Momentum
        self.num_ex_per_epoch = 150000
        self.num_epochs_per_decay = 5
        self.ann_param = 0.0
        self.std = 0.1
        self.temperature = 1.0
        self.gen_max_len = 30
        self.beam_size = 10
        self.sample_gen = "beam_search"
        self.mode = "training"
        for k, v in kw.items():
            if not hasattr(self, k):
                raise AttributeError("unknown config field %s" % k)
            setattr(self, k, v)

    @property
    def has_cv_input(self):
        # main.py:52-54, 103-104
        return self.use_c_v or self.prior in ("GMM", "AG")


# ------------------------------------------------------------------------------------------
# bf16 emulation (straight-through)
class _RoundBF16(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return g


def _r(x, emulate):
    return _RoundBF16.apply(x) if emulate else x


# ------------------------------------------------------------------------------------------
# parameters (TF variable names and layouts of SURVEY.md 5.4)
VGG_LAYERS = [("conv1_1", 3, 64), ("conv1_2", 64, 64), ("conv2_1", 64, 128), ("conv2_2", 128, 128),
              ("conv3_1", 128, 256), ("conv3_2", 256, 256), ("conv3_3", 256, 256), ("conv4_1", 256, 512),
              ("conv4_2", 512, 512), ("conv4_3", 512, 512), ("conv5_1", 512, 512), ("conv5_2", 512, 512),
              ("conv5_3", 512, 512)]
VGG_POOL_AFTER = {"conv1_2", "conv2_2", "conv3_3", "conv4_3", "conv5_3"}


def vgg_param_names():
    """Creation order of image_embeddings.py:37-238 (== load_weights order, :240-246)."""
    names = []
    for name, cin, cout in VGG_LAYERS:
        suffix = "_conv" if name.startswith("conv5") else ""  # image_embeddings.py:176-201
        names.append(("cnn/%s/weights%s" % (name, suffix), (3, 3, cin, cout)))
        names.append(("cnn/%s/biases%s" % (name, suffix), (cout,)))
    names.append(("cnn/fc1/weights", (25088, 4096)))
    names.append(("cnn/fc1/biases", (4096,)))
    names.append(("cnn/fc2/weights", (4096, 4096)))
    names.append(("cnn/fc2/biases", (4096,)))
    return names


def param_shapes(cfg, with_cnn=False):
    """Ordered {name: shape} of the trainable variables the reference creates."""
    E, He, Hd, Z, S, V = (cfg.embed_size, cfg.encoder_hidden, cfg.decoder_hidden, cfg.latent_size, cfg.gen_z_samples,
                          cfg.vocab_size)
    out = {}
    if with_cnn:
        for n, s in vgg_param_names():
            out[n] = s
    out["imf_emb/kernel"] = (cfg.cnn_feature_size, E)  # main.py:94
    out["imf_emb/bias"] = (E,)
    if cfg.has_cv_input:  # main.py:103-108
        out["cv_emb/kernel"] = (cfg.num_clusters, E)
        out["cv_emb/bias"] = (E,)
    if not cfg.no_encoder:
        out["encoder/enc_embeddings"] = (V, E)  # encoder.py:32-35
        out["encoder/multi_rnn_cell/cell_0/lstm_cell/kernel"] = (E + He, 4 * He)
        out["encoder/multi_rnn_cell/cell_0/lstm_cell/bias"] = (4 * He,)
        if cfg.prior == "Normal":  # encoder.py:59-66
            out["encoder/dense/kernel"] = (He, Z)
            out["encoder/dense/bias"] = (Z,)
            out["encoder/dense_1/kernel"] = (He, Z)
            out["encoder/dense_1/bias"] = (Z,)
        else:  # encoder.py:76-81, 92-97
            tag = "gmm_ll" if cfg.prior == "GMM" else "ag_ll"
            for k in range(cfg.num_clusters):
                out["encoder/%s_%d/dense/kernel" % (tag, k)] = (He, Z)
                out["encoder/%s_%d/dense/bias" % (tag, k)] = (Z,)
                out["encoder/%s_%d/dense_1/kernel" % (tag, k)] = (He, Z)
                out["encoder/%s_%d/dense_1/bias" % (tag, k)] = (Z,)
    out["decoder/net/dec_embeddings"] = (V, E)  # decoder.py:76-81
    out["decoder/net/multi_rnn_cell/cell_0/lstm_cell/kernel"] = (E + Hd, 4 * Hd)
    out["decoder/net/multi_rnn_cell/cell_0/lstm_cell/bias"] = (4 * Hd,)
    if not cfg.no_encoder:
        out["decoder/net/z_rnn/kernel"] = (Z * S, E)  # decoder.py:111-112
        out["decoder/net/z_rnn/bias"] = (E,)
    out["decoder/rnn_logits/kernel"] = (Hd, V)  # decoder.py:127-129
    out["decoder/rnn_logits/bias"] = (V,)
    return out


def init_params(cfg, seed=1, with_cnn=False, dtype=torch.float64, scale=1.0):
    """Seeded glorot-uniform weights / zero biases (TF defaults for get_variable / layers.dense /
    LSTMCell). `scale` > 1 sharpens the random model (used by decode tests to get varied tokens)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    params = {}
    for name, shape in param_shapes(cfg, with_cnn).items():
        if len(shape) == 1:
            a = np.zeros(shape)
            if name.endswith("lstm_cell/bias") or "cnn/" in name:
                a = a  # zeros (LSTMCell bias initializer; synthetic VGG biases also zero)
        else:
            if len(shape) == 4:
                fan_in, fan_out = shape[0] * shape[1] * shape[2], shape[0] * shape[1] * shape[3]
            else:
                fan_in, fan_out = shape[0], shape[1]
            lim = math.sqrt(6.0 / (fan_in + fan_out))
            if name.startswith("cnn/"):
                lim = math.sqrt(6.0 / fan_in)  # He-uniform keeps 13 random ReLU conv layers from dying
            a = rng.uniform(-lim, lim, size=shape) * scale
        params[name] = torch.tensor(np.asarray(a, dtype=np.float32)).to(dtype)
    return params


def init_clusters(num_clusters, latent_size, seed=2):
    """utils/vae_utils.py:20-30: num_clusters random unit vectors in [-1,1]^Z (constants)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    m = 2 * rng.random((num_clusters, latent_size)) - 1
    m = m / np.sqrt(np.sum(m ** 2, axis=1, keepdims=True))
    return torch.tensor(m.astype(np.float32))


# ------------------------------------------------------------------------------------------
# building blocks
def lstm_cell(x, h, c, kernel, bias, emulate=False):
    """tf.contrib.rnn.LSTMCell step: gate order i, j, f, o; forget_bias 1.0."""
    g = torch.cat([x, h], 1) @ _r(kernel, emulate) + bias
    i, j, f, o = torch.chunk(g, 4, dim=1)
    c_new = torch.sigmoid(f + 1.0) * c + torch.sigmoid(i) * torch.tanh(j)
    h_new = torch.sigmoid(o) * torch.tanh(c_new)
    return _r(h_new, emulate), c_new  # the CUDA path keeps h in bf16 (it is a GEMM operand), c in fp32


def dynamic_rnn(x_seq, lengths, h, c, kernel, bias, out_keep_mask=None, keep_prob=1.0, emulate=False):
    """tf.nn.dynamic_rnn(sequence_length=...): for t >= len the emitted output is zero
    and the state is copied through. DropoutWrapper(output_keep_prob) scales only the emitted h
    (utils/rnn_model.py:45-46). x_seq [N,T,E] -> outputs [N,T,H], final (h, c)."""
    N, T, _ = x_seq.shape
    outs = []
    for t in range(T):
        h_new, c_new = lstm_cell(x_seq[:, t], h, c, kernel, bias, emulate)
        live = (lengths > t).to(x_seq.dtype).unsqueeze(1)
        out = h_new
        if out_keep_mask is not None and keep_prob < 1.0:
            out = _r(out * out_keep_mask[:, t] / keep_prob, emulate)
        outs.append(out * live)
        h = live * h_new + (1 - live) * h
        c = live * c_new + (1 - live) * c
    return torch.stack(outs, 1), h, c


def dense(x, params, name, emulate=False):
    return x @ _r(params[name + "/kernel"], emulate) + params[name + "/bias"]


def vgg16_fc2(params, images, cfg=None, fc_keep_masks=None, keep=1.0, emulate=False, taps=None, branches=None):
    """utils/image_embeddings.py:26-238. images [B,224,224,3] RGB 0..255 -> fc2 [B,4096].
    fc_keep_masks: optional (mask_fc1, mask_fc2) for cnn dropout when fine-tuning (:225-237).
    branches (parity-test aid, not a reference feature): {layer: activation [B,H,W,C] (conv) or [B,4096] (fc1, fc2)} of
    ANOTHER implementation's forward pass. Where given, the ReLU passes exactly the units that are positive there and
    the 2x2 max-pool routes exactly the window element that is largest there, instead of deciding from this pass's own
    values: the piecewise-linear network is then the SAME linear map in both implementations, so gradients can be
    compared tensor by tensor without the branch flips that rounding noise near zero / near ties causes."""
    import torch.nn.functional as F
    dt = images.dtype
    mean = torch.tensor([123.68, 116.779, 103.939], dtype=dt)
    x = _r(images - mean, emulate).permute(0, 3, 1, 2)  # NCHW for torch
    for name, cin, cout in VGG_LAYERS:
        suffix = "_conv" if name.startswith("conv5") else ""
        w = _r(params["cnn/%s/weights%s" % (name, suffix)], emulate).permute(3, 2, 0, 1)  # HWIO -> OIHW
        b = params["cnn/%s/biases%s" % (name, suffix)]
        y = F.conv2d(x, w, b, padding=1)
        other = None if branches is None or name not in branches else branches[name].to(dt).permute(0, 3, 1, 2)
        x = _r(torch.relu(y) if other is None else y * (other > 0).to(dt), emulate)
        if taps is not None:
            taps[name] = x.permute(0, 2, 3, 1)
        if name in VGG_POOL_AFTER:
            if other is None:
                x = F.max_pool2d(x, 2, 2)
            else:
                _, idx = F.max_pool2d(other, 2, 2, return_indices=True)
                x = x.flatten(2).gather(2, idx.flatten(2)).view(idx.shape)
    flat = x.permute(0, 2, 3, 1).reshape(x.shape[0], -1)  # NHWC flatten, image_embeddings.py:222
    fc1 = flat @ _r(params["cnn/fc1/weights"], emulate) + params["cnn/fc1/biases"]
    if branches is not None and "fc1" in branches:  # positive there <=> ReLU open and the unit kept by the dropout
        fc1 = fc1 * ((branches["fc1"] > 0) | (fc_keep_masks is not None and fc_keep_masks[0] == 0)).to(dt)
    else:
        fc1 = torch.relu(fc1)
    if fc_keep_masks is not None:
        fc1 = fc1 * fc_keep_masks[0] / keep
    fc1 = _r(fc1, emulate)
    if taps is not None:
        taps["fc1"] = fc1
    fc2 = fc1 @ _r(params["cnn/fc2/weights"], emulate) + params["cnn/fc2/biases"]
    if branches is not None and "fc2" in branches:
        fc2 = fc2 * ((branches["fc2"] > 0) | (fc_keep_masks is not None and fc_keep_masks[1] == 0)).to(dt)
    else:
        fc2 = torch.relu(fc2)
    if fc_keep_masks is not None:
        fc2 = fc2 * fc_keep_masks[1] / keep
    if taps is not None:
        taps["fc2"] = fc2
    return fc2


def encoder_q(params, cfg, images_fv, c_i_emb, c_v, cap_lbl, lengths, gmm_cluster=None, emulate=False):
    """vae_model/encoder.py:24-110 -> (mu [N,Z], std [N,Z])."""
    N = images_fv.shape[0]
    H = cfg.encoder_hidden
    dt = images_fv.dtype
    K = params["encoder/multi_rnn_cell/cell_0/lstm_cell/kernel"]
    b = params["encoder/multi_rnn_cell/cell_0/lstm_cell/bias"]
    x = _r(params["encoder/enc_embeddings"], emulate)[cap_lbl]  # encoder.py:31-36 (labels are the input, Q8)
    x.retain_grad() if x.requires_grad else None
    h = torch.zeros(N, H, dtype=dt)
    c = torch.zeros(N, H, dtype=dt)
    h, c = lstm_cell(images_fv, h, c, K, b, emulate)  # encoder.py:46
    if c_i_emb is not None and cfg.use_c_v:  # encoder.py:47-48
        h, c = lstm_cell(c_i_emb, h, c, K, b, emulate)
    _, h, c = dynamic_rnn(x, lengths, h, c, K, b, emulate=emulate)  # encoder.py:49-58
    if cfg.prior == "Normal":  # encoder.py:59-66
        mu = dense(h, params, "encoder/dense", emulate)
        std = torch.exp(dense(h, params, "encoder/dense_1", emulate))
    else:
        tag = "gmm_ll" if cfg.prior == "GMM" else "ag_ll"
        tm = torch.stack([dense(h, params, "encoder/%s_%d/dense" % (tag, k), emulate)
                          for k in range(cfg.num_clusters)], 1)  # [N,90,Z]
        tl = torch.stack([dense(h, params, "encoder/%s_%d/dense_1" % (tag, k), emulate)
                          for k in range(cfg.num_clusters)], 1)
        if cfg.prior == "GMM":  # encoder.py:72-88 (cluster pick is an explicit input, Q3)
            idx = torch.arange(N)
            mu = tm[idx, gmm_cluster]
            std = torch.exp(tl)[idx, gmm_cluster]
        else:  # AG, encoder.py:105-107
            mu = torch.einsum("nk,nkz->nz", c_v, tm)
            std = torch.einsum("nk,nkz->nz", c_v, torch.exp(tl))
    return mu, std, x


def kl_term(cfg, mu, std, c_v=None, c_means=None):
    """main.py:118-145. Normal/GMM: scalar batch mean. AG: [N] vector (Q2)."""
    if cfg.prior in ("Normal", "GMM"):
        return -0.5 * torch.mean(torch.sum(1 + torch.log(std ** 2 + 0.00001) - mu ** 2 - std ** 2, 1))
    cs = torch.tensor(C_SIGMA, dtype=torch.float32).to(mu.dtype)  # tf.constant(0.1) is float32
    kc = 0.5 + torch.log(std + 0.00001) - torch.log(cs + 0.00001) - (
        (mu - c_v @ c_means.to(mu.dtype)) ** 2 + std ** 2) / (2 * cs ** 2 + 0.0000001)
    return -0.5 * torch.sum(kc, 1)


def decoder_logits(params, cfg, images_fv, c_i_emb, z, cap_in, lengths, emb_keep_mask=None, out_keep_mask=None,
                   emulate=False):
    """vae_model/decoder.py:34-143 (train mode). z [S,N,Z] or None (no_encoder).
    Returns x_logits [N*T, V] in the reference's row order (row = n*T + t) and the embedded input."""
    N, T = cap_in.shape
    H = cfg.decoder_hidden
    dt = images_fv.dtype
    K = params["decoder/net/multi_rnn_cell/cell_0/lstm_cell/kernel"]
    b = params["decoder/net/multi_rnn_cell/cell_0/lstm_cell/bias"]
    x_lookup = _r(params["decoder/net/dec_embeddings"], emulate)[cap_in]  # decoder.py:77-83
    x_lookup.retain_grad() if x_lookup.requires_grad else None
    x = x_lookup
    if cfg.dec_keep_rate < 1 and emb_keep_mask is not None:  # decoder.py:85-87
        x = _r(x * emb_keep_mask / cfg.dec_keep_rate, emulate)
    h = torch.zeros(N, H, dtype=dt)
    c = torch.zeros(N, H, dtype=dt)
    h, c = lstm_cell(images_fv, h, c, K, b, emulate)  # decoder.py:100
    if c_i_emb is not None and cfg.use_c_v:  # decoder.py:101-102
        h, c = lstm_cell(c_i_emb, h, c, K, b, emulate)
    if not cfg.no_encoder:
        # Q1: row-major reinterpretation [S,N,Z] -> [N, S*Z] (decoder.py:109-110)
        z_flat = _r(z, emulate).reshape(-1, cfg.latent_size * cfg.gen_z_samples)
        z_dec = _r(dense(z_flat, params, "decoder/net/z_rnn", emulate), emulate)
        h, c = lstm_cell(z_dec, h, c, K, b, emulate)  # decoder.py:113
    out, h, c = dynamic_rnn(x, lengths, h, c, K, b, out_keep_mask, cfg.dec_lstm_drop, emulate)
    logits = dense(out.reshape(N * T, H), params, "decoder/rnn_logits", emulate)  # decoder.py:126-129
    return logits, x_lookup


def annealing_coeff(cfg, global_step):
    """main.py:162-170."""
    if cfg.fine_tune or cfg.restore:
        return 1.0
    if cfg.ann_param > 1:
        return float((np.tanh((np.float32(global_step) - 1000 * cfg.ann_param) / 1000) + 1) / 2)
    return 1.0


def forward(params, cfg, batch, emulate=False, c_means=None):
    """One forward pass of the training graph. batch keys:
       feats [B,4096] (or images [B,224,224,3] when cfg.fine_tune), cap_lbl/cap_in int64 [N,T],
       lengths int64 [N], c_v [N,90] (optional), global_step, eps [S,N,Z], emb_keep_mask [N,T,E],
       out_keep_mask [N,T,H], gmm_cluster [N], cnn_keep_masks.
    Returns dict with logits, mu, std, z, kld, rec_loss, lower_bound, annealing, + taps."""
    dt = params["imf_emb/kernel"].dtype
    C = cfg.num_captions
    l2 = 0.0
    if cfg.fine_tune:
        masks = batch.get("cnn_keep_masks") if cfg.mode == "training" else None
        feats = vgg16_fc2(params, batch["images"].to(dt), cfg, masks, cfg.cnn_dropout, emulate,
                          branches=batch.get("vgg_branches"))
        if cfg.mode == "training":  # Q11: l2_regularizer(weight_decay) on every cnn/ variable
            for n, _ in vgg_param_names():
                l2 = l2 + cfg.weight_decay * torch.sum(params[n] ** 2) / 2
    else:
        feats = batch["feats"].to(dt)
    feats = _r(feats, emulate)
    if C > 1 and cfg.mode == "training":  # main.py:84-89 (Q7)
        feats = feats.unsqueeze(1).expand(-1, C, -1).reshape(-1, cfg.cnn_feature_size)
    images_fv = _r(dense(feats, params, "imf_emb", emulate), emulate)  # main.py:94
    c_v = batch.get("c_v")
    c_i_emb = None
    if cfg.has_cv_input:
        c_v = c_v.to(dt)
        c_i_emb = _r(dense(_r(c_v, emulate), params, "cv_emb", emulate), emulate)  # main.py:108
    cap_lbl, cap_in, lengths = batch["cap_lbl"], batch["cap_in"], batch["lengths"]
    res = {}
    if not cfg.no_encoder:
        mu, std, x_enc = encoder_q(params, cfg, images_fv, c_i_emb, c_v, cap_lbl, lengths, batch.get("gmm_cluster"),
                                   emulate)
        z = mu + std * batch["eps"].to(dt)  # zs.Normal reparameterised sample, encoder.py:108-109
        if c_means is None and cfg.prior == "AG":
            c_means = init_clusters(cfg.num_clusters, cfg.latent_size)
        kld = kl_term(cfg, mu, std, c_v, c_means)
        res.update(mu=mu, std=std, z=z, x_enc=x_enc)
    else:
        z = None
        kld = torch.zeros((), dtype=dt)
    logits, x_dec = decoder_logits(params, cfg, images_fv, c_i_emb, z, cap_in, lengths, batch.get("emb_keep_mask"),
                                   batch.get("out_keep_mask"), emulate)
    if emulate:
        logits = _r(logits, True)  # the CUDA path stores logits in bf16
    labels = cap_lbl.reshape(-1)  # main.py:152-158
    ce = torch.logsumexp(logits, 1) - logits[torch.arange(labels.numel()), labels]
    mask = torch.sign(labels.to(dt))
    rec = torch.sum(ce * mask) / torch.sum(mask) + l2  # get_total_loss adds the regulariser (main.py:159-160)
    ann = annealing_coeff(cfg, batch.get("global_step", 0))
    lb = rec + ann * kld / 10 if not cfg.no_encoder else rec  # main.py:172-177
    res.update(logits=logits, kld=kld, rec_loss=rec, lower_bound=lb, annealing=ann, x_dec=x_dec,
               images_fv=images_fv)
    return res


# ------------------------------------------------------------------------------------------
# optimiser (ops/optimizers.py + TF semantics Q4/Q5)
def trainable_names(cfg, params):
    """ops/optimizers.py:4-12: scopes cv_emb, imf_emb, decoder, (encoder)."""
    names = [n for n in params if n.startswith("cv_emb/")]
    names += [n for n in params if n.startswith("imf_emb/")]
    names += [n for n in params if n.startswith("decoder/")]
    if not cfg.no_encoder:
        names += [n for n in params if n.startswith("encoder/")]
    return names


def compute_grads(params, cfg, batch, emulate=False, c_means=None):
    """tf.gradients(lower_bound, vars). A vector lower bound (AG, Q2) is differentiated as its sum.
    Returns (forward result, {name: grad or None}, global_norm as TF computes it (Q4))."""
    names = trainable_names(cfg, params)
    cnn_names = [n for n, _ in vgg_param_names()] if cfg.fine_tune else []
    leaves = {n: params[n].detach().clone().requires_grad_(True) for n in names + cnn_names}
    p2 = dict(params)
    p2.update(leaves)
    res = forward(p2, cfg, batch, emulate, c_means)
    res["lower_bound"].sum().backward()
    grads = {n: leaves[n].grad for n in names + cnn_names}
    sq = 0.0
    for n in names:
        g = grads[n]
        if g is None:
            continue
        if n.endswith("enc_embeddings"):
            sq = sq + float((res["x_enc"].grad ** 2).sum())  # IndexedSlices.values, not aggregated (Q4)
        elif n.endswith("dec_embeddings"):
            sq = sq + float((res["x_dec"].grad ** 2).sum())
        else:
            sq = sq + float((g ** 2).sum())
    return res, grads, math.sqrt(sq)


def adam_update(p, g, m, v, lr, t, beta1=0.8, beta2=0.999, eps=1e-8):
    """TF1 AdamOptimizer (Q5): epsilon outside the bias correction; t = 1-based step."""
    lr_t = lr * math.sqrt(1 - beta2 ** t) / (1 - beta1 ** t)
    m = beta1 * m + (1 - beta1) * g
    v = beta2 * v + (1 - beta2) * g * g
    p = p - lr_t * m / (torch.sqrt(v) + eps)
    return p, m, v


def decay_steps(cfg):
    """ops/optimizers.py:24-26: int(num_ex_per_epoch / (batch_size + 0.001) * num_epochs_per_decay)."""
    return int(cfg.num_ex_per_epoch / (cfg.batch_size + 0.001) * cfg.num_epochs_per_decay)


def optimizer_update(kind, p, g, m, v, lr, t, cfg):
    """One apply_gradients of ops/optimizers.py:33-46 / 68-81. t = 1-based step; the decayed rate reads global_step
    BEFORE its increment (t - 1), staircase halving every decay_steps; Adam ignores the decay (Q5)."""
    if kind == "Adam":
        return adam_update(p, g, m, v, lr, t)
    lr_d = lr * 0.5 ** ((t - 1) // max(decay_steps(cfg), 1))
    if kind == "SGD":          # tf.train.GradientDescentOptimizer
        return p - lr_d * g, m, v
    if kind == "Momentum":     # tf.train.MomentumOptimizer(lr, 0.9): accum = 0.9 * accum + g ; p -= lr * accum
        m = 0.9 * m + g
        return p - lr_d * m, m, v
    raise ValueError("unknown optimizer %r" % (kind,))


def train_step(params, opt, cfg, batch, emulate=False, c_means=None):
    """One sess.run([kld, rec_loss, lower_bound, optimize, optimize_cnn, annealing]) (main.py:241-244).
    params/opt are updated in place (opt = {"t": int, "m": {}, "v": {}}). Returns the fetches."""
    res, grads, gnorm = compute_grads(params, cfg, batch, emulate, c_means)
    clip = cfg.lstm_clip_by_norm
    scale = clip / max(gnorm, clip)  # tf.clip_by_global_norm
    opt["t"] = opt.get("t", 0) + 1
    for n in trainable_names(cfg, params):
        g = grads[n]
        if g is None:
            continue
        m = opt["m"].get(n, torch.zeros_like(params[n]))
        v = opt["v"].get(n, torch.zeros_like(params[n]))
        params[n], opt["m"][n], opt["v"][n] = optimizer_update(cfg.optimizer, params[n], g * scale, m, v, cfg.learning_rate,
                                                               opt["t"], cfg)
    if cfg.fine_tune:  # ops/optimizers.py:49-82: no clipping, cnn_lr
        for n, _ in vgg_param_names():
            g = grads[n]
            m = opt["m"].get(n, torch.zeros_like(params[n]))
            v = opt["v"].get(n, torch.zeros_like(params[n]))
            params[n], opt["m"][n], opt["v"][n] = optimizer_update(cfg.cnn_optimizer, params[n], g, m, v, cfg.cnn_lr, opt["t"], cfg)
    return {"kld": res["kld"].detach(), "rec_loss": float(res["rec_loss"].detach()), "lower_bound": res["lower_bound"].detach(),
            "annealing": res["annealing"], "global_norm": gnorm, "grads": grads, "res": res}


# ------------------------------------------------------------------------------------------
# synthetic batches
def synthetic_batch(cfg, B, T, seed=0, dtype=torch.float64, ragged=False, with_images=False):
    rng = np.random.Generator(np.random.PCG64(seed))
    C = cfg.num_captions
    N = B * C
    V = cfg.vocab_size
    batch = {}
    if with_images:
        batch["images"] = torch.tensor(rng.integers(0, 256, size=(B, 224, 224, 3)).astype(np.float32)).to(dtype)
    batch["feats"] = torch.tensor(np.maximum(0, rng.standard_normal((B, cfg.cnn_feature_size))).astype(np.float32)).to(dtype)
    body = np.clip(rng.zipf(1.1, size=(N, T + 1)), 0, None) % max(V - 3, 1) + 3
    body = np.minimum(body, V - 1)
    toks = body.astype(np.int64)
    toks[:, 0] = 1  # <BOS>
    lengths = np.full((N,), T, dtype=np.int64)
    if ragged:
        lengths = rng.integers(0, T + 1, size=(N,)).astype(np.int64)
        lengths[0] = T
        if N > 1:
            lengths[1] = 0  # images with < C captions yield empty rows (batch_gen.py:313-317)
    cap_in = np.zeros((N, T), dtype=np.int64)
    cap_lbl = np.zeros((N, T), dtype=np.int64)
    for n in range(N):
        L = int(lengths[n])
        if L == 0:
            continue
        seq = list(toks[n, :L]) + [2]  # <BOS> w1 .. w_{L-1} <EOS>  (len(caption) = L+1)
        cap_in[n, :L] = seq[:L]
        cap_lbl[n, :L] = seq[1:L + 1]
    batch["cap_in"] = torch.tensor(cap_in)
    batch["cap_lbl"] = torch.tensor(cap_lbl)
    batch["lengths"] = torch.tensor(lengths)
    if cfg.has_cv_input:
        used = [k for k in range(1, 91) if k not in {12, 26, 29, 30, 45, 66, 68, 69, 71, 83}]
        cv = np.zeros((B, 91), dtype=np.float32)
        for b in range(B):
            k = int(np.clip(rng.geometric(0.34), 1, 18))
            ids = rng.choice(used, size=k, replace=False)
            cv[b, ids] = 1.0 / k
        cv = np.repeat(cv[:, None, :], C, axis=1).reshape(N, 91)[:, 1:]  # caption_utils.py:22-24, main.py:236
        batch["c_v"] = torch.tensor(cv).to(dtype)
        if cfg.prior == "GMM":
            batch["gmm_cluster"] = torch.tensor(rng.integers(0, cfg.num_clusters, size=(N,)).astype(np.int64))
    if not cfg.no_encoder:
        batch["eps"] = torch.tensor(rng.standard_normal((cfg.gen_z_samples, N, cfg.latent_size)).astype(np.float32)).to(dtype)
    if cfg.dec_keep_rate < 1:
        batch["emb_keep_mask"] = torch.tensor((rng.random((N, T, cfg.embed_size)) < cfg.dec_keep_rate).astype(np.float32)).to(dtype)
    if cfg.dec_lstm_drop < 1:
        batch["out_keep_mask"] = torch.tensor((rng.random((N, T, cfg.decoder_hidden)) < cfg.dec_lstm_drop).astype(np.float32)).to(dtype)
    batch["global_step"] = 0
    return batch
