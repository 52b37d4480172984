"""CPU suite: pins the oracle (test infrastructure) against (1) golden vectors produced by RUNNING the reference's
own code (tests/golden/make_reference_fixtures.py) and (2) analytic known answers for the TF/zhusuan semantics the
reference delegates to. No GPU, no /root/reference at run time."""
import json
import math
import os

import numpy as np
import pytest
import torch

from helpers import O, SMALL, TINY, make_case
from oracle import decode_oracle as D

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def gold(name):
    with open(os.path.join(GOLD, name)) as f:
        return json.load(f)


# ------------------------------------------------------------------ golden vectors from the reference itself
def test_topn_matches_reference():
    for c in gold("topn.json"):
        t = D.TopN(c["n"])
        for i, s in enumerate(c["scores"]):
            t.push(D.Beam([i], None, s, s))
        assert t.size() == c["size"]
        assert [b.sentence[0] for b in t.extract()] == c["extract"]
        t.reset()
        for i, s in enumerate(c["scores"]):
            t.push(D.Beam([i], None, s, s))
        assert [b.sentence[0] for b in t.extract(sort=True)] == c["extract_sorted"]


def test_decode_loops_match_reference():
    """Decoder.online_inference / Decoder.beam_search of the reference, executed with a deterministic fake session,
    vs the oracle's restatement of the two loops on the same step function (incl. Q9: <BOS> fed twice, the
    test-split <PAD> bug, length normalisation only on completed captions)."""
    for c in gold("decode_loops.json"):
        V, bos, eos = c["V"], 1, 2
        idx2word = {i: "w%d" % i for i in range(V)}
        idx2word.update({0: "<PAD>", 1: "<BOS>", 2: "<EOS>"})
        for i in range(c["n_img"]):
            m = D.HashModel(V, c["seed"] * 1000 + i)
            if c["mode"] == "beam_search":
                s = D.beam_search(m.step, c["beam"], c["max_len"], bos, eos)
                cap = D.captions_json([100 + i], [s], idx2word, bos, eos)[0]
                assert cap == c["captions"][i], (c["mode"], c["beam"], V, i)
                m2 = D.HashModel(V, c["seed"] * 1000 + i)
                beams = D.beam_search(m2.step, c["beam"], c["max_len"], bos, eos, ret_beams=True)
                got = [" ".join(idx2word[w] for w in b if w not in (bos, eos)) for b in beams]
                assert got == c["ret_beams"][i]["caption"]
            else:
                mode = "greedy" if c["mode"] == "greedy" else "beam_search"
                raw = D.online_inference(m.step, mode, c["max_len"], 0.7, bos, eos)
                assert raw == c["raw"][i], (c["mode"], V, i)
                cap = D.captions_json([100 + i], [[bos] + raw], idx2word, bos, eos)[0]
                assert cap == c["captions"][i]


def test_parameters_match_reference():
    from vae_captioning_b200.parameters import Parameters
    g = gold("parameters.json")
    for k, v in g["defaults"].items():
        assert getattr(Parameters, k) == v, k
    for case in g["parsed"]:
        p = Parameters().parse_args(case["argv"])
        for k, v in case["attrs"].items():
            assert getattr(p, k) == v, (k, case["argv"])
        assert os.environ["CUDA_VISIBLE_DEVICES"] == case["CUDA_VISIBLE_DEVICES"]
    assert g["no_gpu_error"] == "TypeError"
    with pytest.raises(TypeError):
        Parameters().parse_args([])


def test_caption_utils_match_reference():
    from vae_captioning_b200.caption_utils import preprocess_captions
    for c in gold("caption_utils.json"):
        (i2, l2), ln2, cv2 = preprocess_captions((np.array(c["inputs"]), np.array(c["labels"])), np.array(c["lengths"]),
                                                 np.array(c["cv"]))
        assert i2.tolist() == c["out_inputs"] and l2.tolist() == c["out_labels"]
        assert ln2.tolist() == c["out_lengths"]
        assert np.asarray(cv2).tolist() == c["out_cv"]


def test_vocabulary_matches_reference():
    from vae_captioning_b200.captions import Dictionary, tokenize_caption
    g = gold("vocabulary.json")
    toks = [tokenize_caption(t) for t in g["texts"]]
    assert toks == g["tokens"]
    caption_dict = {}
    for i, t in enumerate(toks):
        caption_dict.setdefault("img%d.jpg" % (i // 2), []).append(t)
    for keep, v in g["vocab"].items():
        d = Dictionary(caption_dict, int(keep), save_path=None)
        assert d.word2idx == v["word2idx"]
        assert d.vocab_size == v["vocab_size"]
        assert [d.seq2dx([w if w in d.word2idx else "<UNK>" for w in t]) for t in toks] == v["indexed"]


def test_inference_json_matches_reference(tmp_path):
    from vae_captioning_b200 import inference as inf
    from vae_captioning_b200.parameters import Parameters

    class Gen(object):
        def next_val_batch(self, get_image_ids=False, use_obj_vectors=False):
            for b in range(2):
                yield (np.full((3, 4), b, np.float32), None, None, [10 * b + i for i in range(3)],
                       np.arange(3 * 91, dtype=np.float32).reshape(3, 91))

        def next_test_batch(self, use_obj_vectors=False):
            yield (np.zeros((2, 4), np.float32), ["t0.jpg", "t1.jpg"], np.arange(2 * 91, dtype=np.float32).reshape(2, 91))

    class Dec(object):
        def __init__(self):
            self.calls = []

        def beam_search(self, ids, feats, c_v=None, beam_size=2):
            self.calls.append(["beam_search", list(ids), None if c_v is None else np.asarray(c_v).shape[1], beam_size])
            return [{"image_id": i, "caption": "beam %s" % i} for i in ids]

        def online_inference(self, ids, feats, c_v=None):
            self.calls.append(["online_inference", list(ids), None if c_v is None else np.asarray(c_v).shape[1]])
            return [{"image_id": i, "caption": "greedy %s" % i} for i in ids], None

    for c in gold("inference_json.json"):
        p = Parameters()
        p.sample_gen, p.use_c_v, p.prior, p.gen_name, p.beam_size = c["sample_gen"], c["use_c_v"], c["prior"], "gg", 3
        d = Dec()
        inf.inference(p, d, Gen(), Gen(), out_dir=str(tmp_path), reference_test_split_bug=True, verbose=False)
        assert json.load(open(tmp_path / "val_gg.json")) == c["val"]
        assert json.load(open(tmp_path / "test_gg.json")) == c["test"]
        assert d.calls == [list(x) for x in c["calls"]]


# ------------------------------------------------------------------ analytic known answers
def test_kat_kl_standard_normal():
    cfg = O.Config(**TINY)
    Z = cfg.latent_size
    mu, std = torch.zeros(3, Z, dtype=torch.float64), torch.ones(3, Z, dtype=torch.float64)
    assert abs(float(O.kl_term(cfg, mu, std)) - (-0.5 * Z * math.log(1 + 1e-5))) < 1e-12


def test_kat_kl_ag_at_prior():
    cfg = O.Config(prior="AG", **TINY)
    Z, K = cfg.latent_size, cfg.num_clusters
    cm = O.init_clusters(K, Z).double()
    c_v = torch.zeros(2, K, dtype=torch.float64)
    c_v[0, 3] = 1.0
    c_v[1, 5] = c_v[1, 7] = 0.5
    mu = c_v @ cm
    std = torch.full((2, Z), 0.1, dtype=torch.float64)
    kl = O.kl_term(cfg, mu, std, c_v, cm)
    cs = float(np.float32(0.1))
    want = -0.5 * Z * (0.5 + math.log(0.1 + 1e-5) - math.log(cs + 1e-5) - 0.01 / (2 * cs * cs + 1e-7))
    assert kl.shape == (2,)  # Q2: a vector
    np.testing.assert_allclose(kl.numpy(), [want, want], rtol=1e-9)


def test_kat_ce_uniform_logits_and_mask():
    cfg, params, batch = make_case(TINY, 2, 5, seed=1, ragged=True)
    params = {k: torch.zeros_like(v) for k, v in params.items()}
    res = O.forward(params, cfg, batch)
    assert abs(float(res["rec_loss"]) - math.log(cfg.vocab_size)) < 1e-12
    assert float(res["logits"].abs().max()) == 0.0


def test_kat_lstm_zero_weights():
    H = 8
    x, h, c = torch.randn(3, 4, dtype=torch.float64), torch.randn(3, H, dtype=torch.float64), torch.randn(3, H, dtype=torch.float64)
    h2, c2 = O.lstm_cell(x, h, c, torch.zeros(4 + H, 4 * H, dtype=torch.float64), torch.zeros(4 * H, dtype=torch.float64))
    sig1 = 1 / (1 + math.exp(-1.0))
    np.testing.assert_allclose(c2.numpy(), sig1 * c.numpy(), rtol=1e-12)  # forget_bias = 1, i*tanh(j) = 0.5*0
    np.testing.assert_allclose(h2.numpy(), 0.5 * np.tanh(c2.numpy()), rtol=1e-12)


def test_kat_lstm_gate_order_ijfo():
    """Gate columns are [i | j | f | o]: drive one gate block at a time."""
    H = 2
    kernel = torch.zeros(1 + H, 4 * H, dtype=torch.float64)
    bias = torch.zeros(4 * H, dtype=torch.float64)
    bias[0:H] = 50.0      # i -> 1
    bias[H:2 * H] = 50.0  # j -> tanh = 1
    bias[2 * H:3 * H] = -50.0  # f -> 0
    bias[3 * H:] = 50.0   # o -> 1
    h2, c2 = O.lstm_cell(torch.zeros(1, 1, dtype=torch.float64), torch.zeros(1, H, dtype=torch.float64),
                         torch.full((1, H), 7.0, dtype=torch.float64), kernel, bias)
    np.testing.assert_allclose(c2.numpy(), 1.0, rtol=1e-9)
    np.testing.assert_allclose(h2.numpy(), math.tanh(1.0), rtol=1e-9)


def test_kat_adam_first_step_and_clip():
    g = torch.tensor([0.3, -2.0, 1e-4], dtype=torch.float64)
    p, m, v = O.adam_update(torch.zeros(3, dtype=torch.float64), g, torch.zeros(3, dtype=torch.float64),
                            torch.zeros(3, dtype=torch.float64), lr=0.0005, t=1)
    lr_t = 0.0005 * math.sqrt(1 - 0.999) / (1 - 0.8)
    want = -lr_t * 0.2 * g / (torch.sqrt(0.001 * g * g) + 1e-8)
    np.testing.assert_allclose(p.numpy(), want.numpy(), rtol=1e-12)
    np.testing.assert_allclose(p.numpy(), -0.0005 * np.sign(g.numpy()), rtol=1e-2)  # ~ lr * sign(g)


def test_clip_identity_below_threshold_and_scaling_above():
    cfg, params, batch = make_case(TINY, 2, 5, seed=2)
    p1 = {k: v.clone() for k, v in params.items()}
    out = O.train_step(p1, {"t": 0, "m": {}, "v": {}}, cfg, batch)
    assert out["global_norm"] < cfg.lstm_clip_by_norm
    cfg2 = O.Config(lstm_clip_by_norm=out["global_norm"] / 4, **TINY)
    p2 = {k: v.clone() for k, v in params.items()}
    out2 = O.train_step(p2, {"t": 0, "m": {}, "v": {}}, cfg2, batch)
    assert abs(out2["global_norm"] - out["global_norm"]) < 1e-12
    # Adam's first step is scale-invariant up to epsilon: clipped and unclipped updates nearly coincide
    n = "decoder/rnn_logits/kernel"
    d1, d2 = (p1[n] - params[n]).numpy(), (p2[n] - params[n]).numpy()
    big = np.abs(grads_of(out, n)) > 1e-4  # epsilon = 1e-8 only matters for vanishing gradients
    np.testing.assert_allclose(d1[big], d2[big], rtol=2e-2)


def grads_of(out, name):
    return out["grads"][name].numpy()


def test_annealing_schedule():
    cfg = O.Config(ann_param=3.0, **TINY)
    assert O.annealing_coeff(cfg, 3000) == pytest.approx(0.5)
    assert O.annealing_coeff(cfg, 0) == pytest.approx((math.tanh(-3.0) + 1) / 2, rel=1e-4)  # float32 arithmetic, as TF
    assert O.annealing_coeff(O.Config(ann_param=1.0, **TINY), 0) == 1.0
    assert O.annealing_coeff(O.Config(ann_param=3.0, restore=True, **TINY), 0) == 1.0


# ------------------------------------------------------------------ properties
def test_padding_logits_equal_bias_and_state_passthrough():
    cfg, params, batch = make_case(SMALL, 2, 6, seed=4, ragged=True)
    res = O.forward(params, cfg, batch)
    N, T = batch["cap_in"].shape
    lg = res["logits"].reshape(N, T, -1)
    b_o = params["decoder/rnn_logits/bias"]
    for n in range(N):
        for t in range(int(batch["lengths"][n]), T):
            assert torch.equal(lg[n, t], b_o)


def test_q1_reshape_mixes_rows():
    """Q1: tf.reshape(z [S,N,Z], [-1, S*Z]) is a row-major reinterpretation: row r holds (s, n) pairs with linear
    index s*N+n in [S*r, S*r+S) -- not example r."""
    S, N, Z = 4, 6, 3
    z = torch.arange(S * N * Z).reshape(S, N, Z)
    flat = z.reshape(-1, Z * S)
    assert flat.shape == (N, S * Z)
    pairs = [(int(v) // (N * Z), (int(v) // Z) % N) for v in flat[1, ::Z]]
    assert pairs == [divmod(4 * 1 + j, N) for j in range(S)]


def test_ag_lower_bound_is_vector_and_gradient_is_sum():
    cfg, params, batch = make_case(TINY, 2, 4, seed=6, prior="AG", use_c_v=True)
    res, grads, gnorm = O.compute_grads(params, cfg, batch)
    N = batch["cap_in"].shape[0]
    assert res["lower_bound"].shape == (N,)  # Q2
    # the rec-loss gradient is multiplied by N: compare with the Normal-style scalar objective
    leaves = {n: params[n].detach().clone().requires_grad_(True) for n in ("decoder/rnn_logits/bias",)}
    p2 = dict(params)
    p2.update(leaves)
    r2 = O.forward(p2, cfg, batch)
    r2["rec_loss"].backward()
    np.testing.assert_allclose(grads["decoder/rnn_logits/bias"].numpy(), N * leaves["decoder/rnn_logits/bias"].grad.numpy(),
                               rtol=1e-9, atol=1e-12)


def test_gmm_gathers_one_head_and_unused_heads_get_zero_grad():
    cfg, params, batch = make_case(TINY, 2, 4, seed=7, prior="GMM", use_c_v=True)
    res, grads, _ = O.compute_grads(params, cfg, batch)
    picked = set(int(k) for k in batch["gmm_cluster"])
    for k in range(cfg.num_clusters):
        g = grads["encoder/gmm_ll_%d/dense/kernel" % k]
        if k in picked:
            assert float(g.abs().max()) > 0
        else:
            assert g is None or float(g.abs().max()) == 0.0


def test_data_parallel_equals_single_device_without_encoder():
    """with equal token counts per shard, the mean of the shard gradients equals the full-batch gradient
    (no_encoder removes the Q1 cross-row coupling)."""
    cfg, params, batch = make_case(TINY, 4, 5, seed=8, no_encoder=True)
    _, g_full, _ = O.compute_grads(params, cfg, batch)
    C = cfg.num_captions
    acc = None
    for r in range(2):
        sl = slice(r * 2 * C, (r + 1) * 2 * C)
        shard = dict(batch)
        shard["feats"] = batch["feats"][r * 2:(r + 1) * 2]
        for k in ("cap_in", "cap_lbl", "lengths"):
            shard[k] = batch[k][sl]
        _, g, _ = O.compute_grads(params, cfg, shard)
        acc = g if acc is None else {k: (acc[k] + g[k]) for k in g}
    for k in g_full:
        np.testing.assert_allclose((acc[k] / 2).numpy(), g_full[k].numpy(), rtol=1e-9, atol=1e-12)


def test_gen_model_matches_training_graph_step():
    """The generation-mode cell (decode oracle, numpy) and the training-mode graph (torch oracle) share weights:
    feeding the same token with the same initial state must give the same next-word distribution."""
    cfg, params, batch = make_case(TINY, 1, 3, seed=9, no_encoder=True, num_captions=1)
    res = O.forward(params, cfg, batch)
    gm = D.GenModel({k: v.numpy() for k, v in params.items()}, cfg)
    step = gm.make_step(batch["feats"][0].numpy(), None, None)
    probs, _ = step(int(batch["cap_in"][0, 0]), None)
    want = torch.softmax(res["logits"][0], 0).numpy()
    np.testing.assert_allclose(probs.ravel(), want, rtol=1e-9, atol=1e-14)


# ----------------------------------------------------------------------------------------------------------------
# Second opinions: TensorFlow cannot run here, so the oracle's restatement of the TF ops is also checked against
# INDEPENDENT implementations of the same published definitions (torch.nn modules written by other people), with the
# TF-specific conventions (gate order, forget bias, padding rule, epsilon placement) applied by re-mapping.
def test_lstm_agrees_with_torch_nn_lstm_after_gate_remap():
    """tf LSTMCell: [x;h] @ W[E+H, 4H] + b, gates i|j|f|o, forget_bias 1. torch.nn.LSTM: gates i|f|g|o, W_ih[4H,E],
    W_hh[4H,H], no forget bias. Same recurrence once columns are permuted and 1.0 is folded into b_f."""
    torch.manual_seed(0)
    N, T, E, H = 5, 7, 6, 4
    kernel = torch.randn(E + H, 4 * H, dtype=torch.float64) * 0.4
    bias = torch.randn(4 * H, dtype=torch.float64) * 0.1
    x = torch.randn(N, T, E, dtype=torch.float64)
    lengths = torch.tensor([7, 3, 0, 5, 1])
    out, h, c = O.dynamic_rnn(x, lengths, torch.zeros(N, H, dtype=torch.float64), torch.zeros(N, H, dtype=torch.float64), kernel, bias)
    i, j, f, o = torch.chunk(kernel, 4, dim=1)
    bi, bj, bf, bo = torch.chunk(bias, 4)
    ref = torch.nn.LSTM(E, H, batch_first=True).double()
    with torch.no_grad():
        w = torch.cat([i, f, j, o], dim=1)  # torch order: input, forget, cell (g = tf's j), output
        ref.weight_ih_l0.copy_(w[:E].t())
        ref.weight_hh_l0.copy_(w[E:].t())
        ref.bias_ih_l0.copy_(torch.cat([bi, bf + 1.0, bj, bo]))
        ref.bias_hh_l0.zero_()
        for n in range(N):  # dynamic_rnn(sequence_length): run each row for its own length, zeros after, state frozen
            L = int(lengths[n])
            if L == 0:
                assert out[n].abs().max() == 0 and h[n].abs().max() == 0 and c[n].abs().max() == 0
                continue
            y, (hn, cn) = ref(x[n:n + 1, :L])
            np.testing.assert_allclose(out[n, :L].numpy(), y[0].numpy(), rtol=1e-10, atol=1e-12)
            assert out[n, L:].abs().max() == 0 if L < T else True
            np.testing.assert_allclose(h[n].numpy(), hn[0, 0].numpy(), rtol=1e-10, atol=1e-12)
            np.testing.assert_allclose(c[n].numpy(), cn[0, 0].numpy(), rtol=1e-10, atol=1e-12)


def test_masked_ce_agrees_with_torch_cross_entropy():
    """main.py:152-158: sparse softmax CE per token, mask = sign(label), token mean == F.cross_entropy(ignore_index=0)."""
    import torch.nn.functional as F
    cfg, params, batch = make_case(TINY, 3, 6, seed=4, ragged=True)
    res = O.forward(params, cfg, batch)
    logits = res["logits"]  # row = n * T + t, the order of cap_lbl.reshape(-1) (decoder.py:126-129, main.py:152)
    labels = batch["cap_lbl"].reshape(-1).long()
    ce = F.cross_entropy(logits, labels, ignore_index=0, reduction="mean")
    np.testing.assert_allclose(float(res["rec_loss"]), float(ce), rtol=1e-10)
    assert int((labels != 0).sum()) > 0


def test_adam_agrees_with_torch_adam_up_to_epsilon_placement():
    """TF1 Adam puts epsilon outside the bias correction (eps_hat = eps / sqrt(1 - beta2^t) relative to the paper form
    torch implements). With eps -> 0 the two coincide; with the TF epsilon they differ exactly by that rescaling."""
    torch.manual_seed(1)
    p0 = torch.randn(50, dtype=torch.float64)
    grads = [torch.randn(50, dtype=torch.float64) for _ in range(4)]
    p, m, v = p0.clone(), torch.zeros(50, dtype=torch.float64), torch.zeros(50, dtype=torch.float64)
    q = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([q], lr=5e-4, betas=(0.8, 0.999), eps=1e-300)
    for t, g in enumerate(grads, start=1):
        p, m, v = O.adam_update(p, g, m, v, lr=5e-4, t=t, eps=0.0)
        q.grad = g.clone()
        opt.step()
        np.testing.assert_allclose(p.numpy(), q.detach().numpy(), rtol=1e-9, atol=1e-15)
    # TF epsilon: p -= lr_t * m / (sqrt(v) + eps)  ==  paper form with eps' = eps / sqrt(1 - beta2^t)
    g = grads[0]
    p1, _, _ = O.adam_update(p0, g, torch.zeros(50, dtype=torch.float64), torch.zeros(50, dtype=torch.float64), lr=5e-4, t=1, eps=1e-8)
    m1, v1 = 0.2 * g, 0.001 * g * g
    paper = p0 - 5e-4 * (m1 / 0.2) / (torch.sqrt(v1 / 0.001) + 1e-8 / math.sqrt(0.001))
    np.testing.assert_allclose(p1.numpy(), paper.numpy(), rtol=1e-9)


def test_vgg_first_layers_agree_with_explicit_window_sums():
    """utils/image_embeddings.py:26-58: mean-subtracted RGB, 3x3 SAME convolution with HWIO filters (TF's
    cross-correlation: out[h,w,o] = sum_{r,s,c} x[h+r-1, w+s-1, c] * W[r,s,c,o]) + bias + ReLU, then 2x2/2 max-pool --
    restated with explicit shifted-window sums in numpy (no convolution routine involved)."""
    torch.manual_seed(2)
    images = torch.randint(0, 256, (1, 224, 224, 3)).double()
    cfg = O.Config()
    params = O.init_params(cfg, seed=3, with_cnn=True, dtype=torch.float64)
    for n in list(params):
        if n.startswith("cnn/") and params[n].dim() == 1:
            params[n] = torch.linspace(-0.5, 0.5, params[n].numel(), dtype=torch.float64)
    taps = {}
    with torch.no_grad():
        O.vgg16_fc2(params, images, taps=taps)
    x = images[0].numpy() - np.array([123.68, 116.779, 103.939])

    def conv_relu(x, w, b):
        H, W, _ = x.shape
        xp = np.pad(x, ((1, 1), (1, 1), (0, 0)))
        out = np.zeros((H, W, w.shape[3]))
        for r in range(3):
            for s_ in range(3):
                out += np.einsum("hwc,co->hwo", xp[r:r + H, s_:s_ + W], w[r, s_])
        return np.maximum(out + b, 0)

    a1 = conv_relu(x, params["cnn/conv1_1/weights"].numpy(), params["cnn/conv1_1/biases"].numpy())
    np.testing.assert_allclose(taps["conv1_1"][0].numpy(), a1, rtol=1e-9, atol=1e-9)
    a2 = conv_relu(a1, params["cnn/conv1_2/weights"].numpy(), params["cnn/conv1_2/biases"].numpy())
    np.testing.assert_allclose(taps["conv1_2"][0].numpy(), a2, rtol=1e-9, atol=1e-8)
    pooled = a2.reshape(112, 2, 112, 2, 64).max(axis=(1, 3))
    a3 = conv_relu(pooled, params["cnn/conv2_1/weights"].numpy(), params["cnn/conv2_1/biases"].numpy())
    np.testing.assert_allclose(taps["conv2_1"][0].numpy(), a3, rtol=1e-9, atol=1e-8)
