"""CPU ORACLE (test infrastructure, not product code) -- decode loops of the reference.

Restates, in plain Python/numpy, the generation side of yiyang92/vae_captioning:
  * Decoder.px_z_fi(gen_mode=True)   vae_model/decoder.py:34-143   -> GenModel.begin / GenModel.step
  * Decoder.online_inference         vae_model/decoder.py:145-201  -> online_inference()
  * Decoder.beam_search              vae_model/decoder.py:203-320  -> beam_search()
  * TopN / Beam                      utils/top_n.py:4-72            -> TopN / Beam
  * inference.inference json layout  ops/inference.py:16-56         -> captions_json()

The two loops are written against an abstract `step(token, state) -> (probs, state)` callable (the role
`sess.run([sample, out_state], feed)` plays in the reference, state=None meaning "initial_state left at its
placeholder default"), so they can be PINNED against the reference's own loop code: tests/golden/
make_reference_fixtures.py imports the reference's decoder.py with stub `tensorflow`/`zhusuan` modules, drives
Decoder.online_inference / Decoder.beam_search with a deterministic fake session (`HashModel` below) and records
the token sequences in tests/golden/decode_loops.json; tests/test_oracle_cpu.py replays them through this file.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.
"""
import heapq
import math

import numpy as np


# ------------------------------------------------------------------------------------------
# utils/top_n.py:4-72
class TopN(object):
    """Size-n min-heap keyed by Beam.score (heapq semantics; ties keep heapq's order)."""

    def __init__(self, n):
        self._n = n
        self._data = []

    def size(self):
        return len(self._data)

    def push(self, x):
        if len(self._data) < self._n:
            heapq.heappush(self._data, x)
        else:
            heapq.heappushpop(self._data, x)

    def extract(self, sort=False):
        data = self._data
        self._data = None
        if sort:
            data.sort(reverse=True)
        return data

    def reset(self):
        self._data = []


class Beam(object):
    def __init__(self, sentence, state, logprob, score):
        self.sentence = sentence
        self.state = state
        self.logprob = logprob
        self.score = score

    def __lt__(self, other):
        return self.score < other.score

    def __eq__(self, other):
        return self.score == other.score


# ------------------------------------------------------------------------------------------
def online_inference(step, sample_gen, gen_max_len, temperature, bos, eos):
    """decoder.py:164-196 for ONE image. Returns the raw generated ids (cap_raw[i]).

    sample_gen 'greedy': argmax of p^(1/t)/sum (temperature is a no-op, Q10). Any other value that is not
    'sample' leaves gen_word_idx at its initial 0 (the test-split bug of Q9: <PAD> x gen_max_len).
    'sample' mode needs the step callable to return the drawn token instead of probabilities."""
    state = None
    sentence = [bos]
    raw = []
    cur_it = 0
    gen_word_idx = 0
    while cur_it < gen_max_len:
        probs, state = step(sentence[-1], state)
        if sample_gen == "greedy":
            p = np.asarray(probs).ravel()
            p = p ** (1 / temperature) / np.sum(p ** (1 / temperature))
            gen_word_idx = int(np.argmax(p))
        elif sample_gen == "sample":
            gen_word_idx = int(probs)
        sentence.append(gen_word_idx)
        raw.append(gen_word_idx)
        cur_it += 1
        if gen_word_idx == eos:
            break
    return raw


def beam_search(step, beam_size, gen_max_len, bos, eos, len_norm_f=0.7, ret_beams=False):
    """decoder.py:227-319 for ONE image. Returns the best beam's full sentence (ids incl. <BOS>/<EOS>), or the
    sorted list of sentences when ret_beams.

    Q9: the first call consumes <BOS> and only its state is kept; the loop then feeds sentence[-1] = <BOS> again."""
    _, state = step(bos, None)
    partial = TopN(beam_size)
    partial.push(Beam([bos], state, 0.0, 0.0))
    complete = TopN(beam_size)
    for _ in range(gen_max_len - 1):
        plist = partial.extract()
        partial.reset()
        outs = [step(c.sentence[-1], c.state) for c in plist]
        for c, (probs, new_state) in zip(plist, outs):
            w_probs = list(enumerate(np.asarray(probs).ravel()))
            w_probs.sort(key=lambda x: -x[1])
            for w, p in w_probs[:beam_size]:
                if p < 1e-12:
                    continue
                sentence = c.sentence + [w]
                logprob = c.logprob + np.log(p)
                score = logprob
                if w == eos:
                    if len_norm_f > 0:
                        score /= len(sentence) ** len_norm_f
                    complete.push(Beam(sentence, new_state, logprob, score))
                else:
                    partial.push(Beam(sentence, new_state, logprob, score))
        if partial.size() == 0:
            break
    if not complete.size():
        complete = partial
    beams = complete.extract(sort=True)
    if ret_beams:
        return [[int(w) for w in b.sentence] for b in beams]
    return [int(w) for w in beams[0].sentence]


def captions_json(image_ids, sentences, idx2word, bos, eos):
    """[{'image_id':..., 'caption': 'w w w'}] as decoder.py:194-196 / :301-306 build it."""
    out = []
    for iid, s in zip(image_ids, sentences):
        out.append({"image_id": iid, "caption": " ".join(idx2word[w] for w in s if w not in (bos, eos))})
    return out


# ------------------------------------------------------------------------------------------
class HashModel(object):
    """Deterministic stand-in for the reference's session in the golden decode fixtures: the next-word
    distribution depends only on the tokens consumed so far (the "state" is that tuple). Bit-reproducible on
    any machine (PCG64 + IEEE double), no BLAS."""

    def __init__(self, vocab, seed, eos=2, eos_boost=0.08, sharp=6.0):
        self.V = vocab
        self.seed = seed
        self.eos = eos
        self.eos_boost = eos_boost
        self.sharp = sharp
        self.calls = 0

    def step(self, token, state):
        self.calls += 1
        hist = tuple(state or ()) + (int(token),)
        key = [self.seed] + [int(t) + 1 for t in hist]
        rng = np.random.Generator(np.random.PCG64(key))
        p = rng.random(self.V) ** self.sharp
        p[0] = 0.0  # <PAD> is never predicted with mass
        p[self.eos] += self.eos_boost * len(hist) * p.max()
        p = (p / p.sum()).astype(np.float32)
        return p.reshape(1, -1), hist


# ------------------------------------------------------------------------------------------
def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


class GenModel(object):
    """px_z_fi(observed={}, gen_mode=True) for one image (decoder.py:41-142), float64 numpy.

    params: {tf_name: array}; eps: [S, Z] standard-normal draws of zs.Normal('z', z_mean, std) for this image
    (the reference redraws them on every sess.run, but they only matter on the first, state-less call)."""

    def __init__(self, params, cfg, c_means=None, bf16=False):
        self.cfg = cfg
        self.bf16 = bf16
        self.p = {k: self._r(np.asarray(v, dtype=np.float64)) if np.asarray(v).ndim > 1 else np.asarray(v, np.float64)
                  for k, v in params.items()}
        self.c_means = None if c_means is None else np.asarray(c_means, np.float64)

    def _r(self, x):
        if not self.bf16:
            return x
        import torch
        return torch.tensor(x).to(torch.bfloat16).to(torch.float64).numpy()

    def _cell(self, x, h, c):
        K = self.p["decoder/net/multi_rnn_cell/cell_0/lstm_cell/kernel"]
        b = self.p["decoder/net/multi_rnn_cell/cell_0/lstm_cell/bias"]
        g = np.concatenate([x, h], 1) @ K + b
        i, j, f, o = np.split(g, 4, axis=1)
        c2 = _sigmoid(f + 1.0) * c + _sigmoid(i) * np.tanh(j)
        h2 = _sigmoid(o) * np.tanh(c2)
        return self._r(h2), c2

    def z_mean(self, c_v):
        """decoder.py:42-71: zeros, or (AG) the mean of the c_means rows of the image's active clusters. The empty-
        vector fallback of the reference indexes row 90 of a 90-row table (Q18); the in-range rows are used here."""
        Z = self.cfg.latent_size
        if self.cfg.prior != "AG":
            return np.zeros((1, Z))
        idx = np.nonzero(np.asarray(c_v).ravel() > 0)[0]
        if idx.size == 0:
            un = {0, 66, 68, 69, 71, 12, 45, 83, 26, 29, 30}
            idx = np.array([i for i in range(self.cfg.num_clusters + 1) if i not in un and i < self.c_means.shape[0]])
        return self.c_means[idx].mean(axis=0).reshape(1, Z)

    def begin(self, feat, c_v, eps):
        """Initial (c, h): LSTM(images_fv) -> [LSTM(c_i)] -> [LSTM(z_dec)]  (decoder.py:96-114)."""
        cfg = self.cfg
        H = cfg.decoder_hidden
        x = self._r(np.asarray(feat, np.float64).reshape(1, -1))
        fv = self._r(x @ self.p["imf_emb/kernel"] + self.p["imf_emb/bias"])
        h = np.zeros((1, H))
        c = np.zeros((1, H))
        h, c = self._cell(fv, h, c)
        if cfg.use_c_v:
            cv = self._r(np.asarray(c_v, np.float64).reshape(1, -1))
            ce = self._r(cv @ self.p["cv_emb/kernel"] + self.p["cv_emb/bias"])
            h, c = self._cell(ce, h, c)
        if not cfg.no_encoder:
            z = self.z_mean(c_v) + cfg.std * np.asarray(eps, np.float64)  # [S, Z]; [S,1,Z] -> [1, S*Z] row-major
            zf = self._r(z.reshape(1, -1))
            zd = self._r(zf @ self.p["decoder/net/z_rnn/kernel"] + self.p["decoder/net/z_rnn/bias"])
            h, c = self._cell(zd, h, c)
        return (c, h)

    def step_from(self, token, state):
        c, h = state
        x = self.p["decoder/net/dec_embeddings"][int(token)].reshape(1, -1)
        h, c = self._cell(x, h, c)
        logits = h @ self.p["decoder/rnn_logits/kernel"] + self.p["decoder/rnn_logits/bias"]
        logits = logits - logits.max()
        e = np.exp(logits)
        return (e / e.sum()), (c, h)

    def make_step(self, feat, c_v, eps):
        init = self.begin(feat, c_v, eps)

        def step(token, state):
            return self.step_from(token, init if state is None else state)
        return step


def log_softmax64(logits):
    m = np.max(logits)
    return logits - m - math.log(np.sum(np.exp(logits - m)))
