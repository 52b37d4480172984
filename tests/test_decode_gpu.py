"""Batched on-device decode (greedy, sample, beam) through the C ABI vs the decode oracle (numpy restatement of
vae_model/decoder.py gen mode + the reference-pinned loops) on the same seeded weights, features and latent draws.

Tolerance (stated): bf16 operands / fp32 accumulate vs the oracle computing in float64 on the same bf16-rounded
weights: next-word probabilities within 2e-2 of the row maximum; token sequences must be identical to the oracle's
unless the oracle itself is at a near-tie at the first divergent step (its two candidates within that same 2e-2), which
the tests assert image by image (64 images per case); beam scores must agree within 2e-2 relative."""
import numpy as np
import pytest
import torch

from helpers import O, SMALL, TINY, engine_for
from oracle import decode_oracle as D
from vae_captioning_b200.decode import Decoder

pytestmark = pytest.mark.gpu
BOS, EOS = 1, 2


class FakeDict(object):
    def __init__(self, V):
        self.idx2word = {i: "w%d" % i for i in range(V)}
        self.idx2word.update({0: "<PAD>", 1: "<BOS>", 2: "<EOS>"})
        self.word2idx = {w: i for i, w in self.idx2word.items()}
        self.vocab_size = V


def make_decode_case(sizes, B, seed=0, scale=4.0, **kw):
    cfg_kw = dict(sizes)
    cfg_kw.update(kw)
    cfg = O.Config(**cfg_kw)
    cfg.gen_max_len = 12
    params = O.init_params(cfg, seed=seed + 1, scale=scale)
    g = np.random.Generator(np.random.PCG64(seed + 3))
    for n in params:
        if params[n].dim() == 1:
            params[n] = torch.tensor(g.uniform(-0.2, 0.2, size=tuple(params[n].shape)).astype(np.float32)).double()
    params["decoder/rnn_logits/bias"][EOS] += 1.0  # captions should terminate before gen_max_len now and then
    feats = np.maximum(0, g.standard_normal((B, cfg.cnn_feature_size))).astype(np.float32)
    c_v = None
    if cfg.has_cv_input:
        c_v = np.zeros((B, cfg.num_clusters), np.float32)
        for b in range(B):
            k = int(g.integers(0, 4))  # includes empty vectors (AG fallback, Q18)
            if k:
                c_v[b, g.choice(cfg.num_clusters, size=k, replace=False)] = 1.0 / k
    eps = g.standard_normal((B, cfg.gen_z_samples, cfg.latent_size)).astype(np.float32)
    return cfg, params, feats, c_v, eps


def oracle_steps(cfg, params, feats, c_v, eps):
    cm = O.init_clusters(cfg.num_clusters, cfg.latent_size).numpy() if cfg.prior == "AG" else None
    gm = D.GenModel({k: v.numpy() for k, v in params.items()}, cfg, c_means=cm, bf16=True)
    return [gm.make_step(feats[b], None if c_v is None else c_v[b], eps[b]) for b in range(len(feats))]


def device_decoder(cfg, params, B, beam):
    eng = engine_for(cfg, params, max(1, (B * beam + cfg.num_captions - 1) // cfg.num_captions), 4)
    return eng, Decoder(eng, cfg, FakeDict(cfg.vocab_size))


CASES = [(TINY, {}), (SMALL, {}), (SMALL, dict(use_c_v=True)), (SMALL, dict(no_encoder=True)),
         (SMALL, dict(prior="AG", use_c_v=True)), (TINY, dict(prior="GMM"))]


@pytest.mark.parametrize("sizes,kw", CASES)
def test_step_probabilities_and_state_feed(sizes, kw):
    B = 6
    cfg, params, feats, c_v, eps = make_decode_case(sizes, B, seed=1, **kw)
    eng, dec = device_decoder(cfg, params, B, 1)
    steps = oracle_steps(cfg, params, feats, c_v, eps)
    rng = {"eps": torch.tensor(eps).cuda()}
    dec.begin(feats, c_v, rng)
    toks = [BOS] * B
    states = [None] * B
    for it in range(4):
        got = dec.step(toks)
        for b in range(B):
            want, states[b] = steps[b](toks[b], states[b])
            want = want.ravel()
            assert np.max(np.abs(got[b] - want)) <= 2e-2 * want.max(), (kw, it, b)
            assert abs(got[b].sum() - 1.0) < 1e-4
        toks = [int(np.argmax(got[b])) for b in range(B)]
    # in_state / out_state round trip (rnn_placeholders): re-feeding the fetched state reproduces the next step
    c0, h0 = dec.get_state()
    p1 = dec.step(toks)
    dec.set_state(c0, h0)
    p2 = dec.step(toks)
    np.testing.assert_allclose(p1, p2, rtol=0, atol=1e-6)
    eng.close()


def _replay(step, sentence):
    """Next-word probability rows of the oracle along `sentence` (ids after <BOS>), fed the way online_inference feeds
    them: rows[i] = distribution from which sentence[i] was chosen."""
    rows, state, prev = [], None, BOS
    for w in sentence:
        probs, state = step(prev, state)
        rows.append(np.asarray(probs).ravel())
        prev = w
    return rows


def _beam_score(step, sentence, len_norm=0.7):
    """Score Decoder.beam_search gives `sentence` (ids incl. <BOS>): <BOS> is consumed by the initial call and fed again
    by the first loop iteration (Q9); length normalisation for completed captions only."""
    _, state = step(BOS, None)
    logprob, prev = 0.0, sentence[0]
    for w in sentence[1:]:
        probs, state = step(prev, state)
        logprob += float(np.log(np.asarray(probs).ravel()[w]))
        prev = w
    return logprob / len(sentence) ** len_norm if sentence[-1] == EOS else logprob


@pytest.mark.parametrize("sizes,kw", CASES)
def test_greedy_matches_oracle(sizes, kw):
    """64 images per case. Token sequences must be IDENTICAL to the oracle's, except where the oracle itself is at a
    near-tie at the first divergent step: the probability it gives its own choice and the device's choice differ by
    less than 2e-2 of the row maximum (the stated next-word probability tolerance). Anything else fails."""
    B = 64
    cfg, params, feats, c_v, eps = make_decode_case(sizes, B, seed=2, **kw)
    eng, dec = device_decoder(cfg, params, B, 1)
    steps = oracle_steps(cfg, params, feats, c_v, eps)
    rng = {"eps": torch.tensor(eps).cuda()}
    ids = list(range(100, 100 + B))
    caps, raw = dec.online_inference(ids, feats, c_v, sample_gen="greedy", rng=rng)
    same, explained = 0, 0
    for b in range(B):
        want = D.online_inference(steps[b], "greedy", cfg.gen_max_len, cfg.temperature, BOS, EOS)
        assert len(raw[b]) <= cfg.gen_max_len and (raw[b][-1] == EOS or len(raw[b]) == cfg.gen_max_len)
        assert caps[b]["image_id"] == ids[b]
        if raw[b] == want:
            same += 1
            continue
        i = next(k for k in range(min(len(raw[b]), len(want))) if raw[b][k] != want[k])
        p = _replay(steps[b], want[:i + 1])[i]
        gap = float(p[want[i]] - p[raw[b][i]])
        assert 0 <= gap <= 2e-2 * float(p.max()), (kw, b, i, gap, float(p.max()))
        explained += 1
    assert same + explained == B and same >= 0.9 * B, (kw, same, explained)
    # the reference's test-split behaviour for sample_gen='beam_search' (Q9): gen_max_len x <PAD>
    caps, raw = dec.online_inference(ids, feats, c_v, sample_gen="beam_search")
    assert raw[0] == [0] * cfg.gen_max_len and caps[0]["caption"] == " ".join(["<PAD>"] * cfg.gen_max_len)
    eng.close()


@pytest.mark.parametrize("beam", [1, 2, 5, 10])
@pytest.mark.parametrize("sizes,kw", [(TINY, {}), (SMALL, dict(use_c_v=True)), (SMALL, dict(prior="AG"))])
def test_beam_search_matches_oracle(sizes, kw, beam):
    """64 images per case. For EVERY image the score the device reports for its best beam equals the score the oracle
    gives that same sentence (2e-2 relative: the device's log-probabilities are right); the returned beam lists must
    be identical to the oracle's, except where the oracle scores the device's best sentence within 2e-2 (relative) of
    its own best -- a near-tie a bf16-level perturbation may flip."""
    B = 64
    cfg, params, feats, c_v, eps = make_decode_case(sizes, B, seed=3, **kw)
    eng, dec = device_decoder(cfg, params, B, beam)
    steps = oracle_steps(cfg, params, feats, c_v, eps)
    rng = {"eps": torch.tensor(eps).cuda()}
    toks, lens, scores, nb = dec.beam_tokens(feats, c_v, beam_size=beam, rng=rng)
    same, explained = 0, 0
    for b in range(B):
        want = D.beam_search(steps[b], beam, cfg.gen_max_len, BOS, EOS, ret_beams=True)
        got = [[int(w) for w in toks[b, j, :lens[b, j]]] for j in range(int(nb[b]))]
        assert got[0][0] == BOS and all(scores[b, j] >= scores[b, j + 1] for j in range(int(nb[b]) - 1))
        s_dev = _beam_score(steps[b], got[0])
        assert abs(float(scores[b, 0]) - s_dev) <= 2e-2 * abs(s_dev) + 1e-3, (kw, beam, b, float(scores[b, 0]), s_dev)
        if got == want:
            same += 1
            continue
        if got[0] != want[0]:
            s_ref = _beam_score(steps[b], want[0])
            assert abs(s_ref - s_dev) <= 2e-2 * abs(s_ref), (kw, beam, b, s_ref, s_dev)
        else:  # same caption, the tail of the list differs: every listed beam must still carry its oracle score
            for j in range(int(nb[b])):
                sj = _beam_score(steps[b], got[j])
                assert abs(float(scores[b, j]) - sj) <= 2e-2 * abs(sj) + 1e-3, (kw, beam, b, j)
        explained += 1
    assert same + explained == B and same >= 0.75 * B, (kw, beam, same, explained)
    caps = dec.beam_search(list(range(B)), feats, c_v, beam_size=beam, rng=rng)
    assert caps[0]["caption"] == " ".join("w%d" % w for w in toks[0, 0, :lens[0, 0]] if w not in (BOS, EOS))
    eng.close()


def test_sample_mode_draws_from_softmax():
    """'sample': tf.multinomial(logits / temperature) -- with a sharpened model the draws concentrate on the argmax,
    and different seeds give different sequences."""
    B = 64
    cfg, params, feats, c_v, eps = make_decode_case(SMALL, B, seed=4, scale=1.0)
    eng, dec = device_decoder(cfg, params, B, 1)
    rng = {"eps": torch.tensor(eps).cuda(), "seed": 11}
    t1, l1 = dec.greedy_tokens(feats, c_v, mode="sample", rng=rng)
    t2, l2 = dec.greedy_tokens(feats, c_v, mode="sample", rng=dict(rng, seed=12))
    t3, l3 = dec.greedy_tokens(feats, c_v, mode="sample", rng=rng)
    assert np.array_equal(t1, t3)          # same seed -> same draws
    assert not np.array_equal(t1, t2)      # different seed -> different draws
    assert t1.max() < cfg.vocab_size and t1.min() >= 0
    eng.close()


def test_decode_argument_errors():
    cfg, params, feats, c_v, eps = make_decode_case(TINY, 2, seed=5)
    eng, dec = device_decoder(cfg, params, 2, 1)
    with pytest.raises(ValueError):
        dec.beam_tokens(feats, c_v, beam_size=17)
    with pytest.raises(ValueError):
        dec.greedy_tokens(feats[:, :5], c_v)
    eng.close()
