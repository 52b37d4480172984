#!/usr/bin/env python
"""Generates tests/golden/*.json by RUNNING the reference's own code (read-only at /root/reference).

The reference's arithmetic lives in TensorFlow 1.x / zhusuan, which cannot be installed here, but its pure-Python
pieces can be executed as they are:
  * utils/top_n.py            TopN / Beam ordering                         -> topn.json
  * utils/caption_utils.py    preprocess_captions row order (n = b*C + c)  -> caption_utils.json
  * utils/parameters.py       Parameters defaults and flag parsing        -> parameters.json
  * utils/captions.py         tokeniser + Dictionary id assignment         -> vocabulary.json
  * vae_model/decoder.py      Decoder.online_inference / Decoder.beam_search host loops, imported with stub
                              `tensorflow` / `zhusuan` modules and driven by a deterministic fake session
                              (oracle.decode_oracle.HashModel)              -> decode_loops.json
  * ops/inference.py          val/test json layout, with fake generators   -> inference_json.json
Nothing from the reference is copied: the fixtures hold only inputs and the outputs the reference produced.
Run from the repo root:  python tests/golden/make_reference_fixtures.py
"""
import contextlib
import io
import json
import os
import sys
import tempfile
import types
from unittest import mock

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from oracle.decode_oracle import HashModel  # noqa: E402


def dump(name, obj):
    with open(os.path.join(HERE, name), "w") as f:
        json.dump(obj, f, indent=0, sort_keys=True)
    print("wrote", name)


def topn_fixture():
    from utils.top_n import Beam, TopN
    cases = []
    rng = np.random.Generator(np.random.PCG64(11))
    for n in (1, 2, 3, 5, 10):
        for count in (0, 1, n, 3 * n + 1):
            scores = [float(np.round(rng.normal(), 3)) for _ in range(count)]
            if count > 3:
                scores[2] = scores[0]  # a tie
            t = TopN(n)
            for i, s in enumerate(scores):
                t.push(Beam([i], None, s, s))
            size = t.size()
            raw = [b.sentence[0] for b in t.extract()]
            t.reset()
            for i, s in enumerate(scores):
                t.push(Beam([i], None, s, s))
            srt = [b.sentence[0] for b in t.extract(sort=True)]
            cases.append({"n": n, "scores": scores, "size": size, "extract": raw, "extract_sorted": srt})
    dump("topn.json", cases)


def caption_utils_fixture():
    from utils.caption_utils import preprocess_captions
    rng = np.random.Generator(np.random.PCG64(5))
    cases = []
    for B, C, T, K in ((2, 5, 4, 91), (3, 2, 6, 7), (1, 1, 3, 0)):
        inp = rng.integers(0, 50, size=(B, C, T))
        lbl = rng.integers(0, 50, size=(B, C, T))
        ln = rng.integers(0, T + 1, size=(B, C)).astype(np.float64)
        cv = rng.random((B, K)) if K else np.array([])
        (i2, l2), ln2, cv2 = preprocess_captions((inp, lbl), ln, cv)
        cases.append({"inputs": inp.tolist(), "labels": lbl.tolist(), "lengths": ln.tolist(), "cv": cv.tolist(),
                      "out_inputs": i2.tolist(), "out_labels": l2.tolist(), "out_lengths": ln2.tolist(),
                      "out_cv": np.asarray(cv2).tolist()})
    dump("caption_utils.json", cases)


def parameters_fixture():
    from utils.parameters import Parameters
    defaults = {k: v for k, v in vars(Parameters).items() if not k.startswith("_") and not callable(v)}
    argvs = [
        ["--gpu", "0"],
        ["--gpu", "3", "--lr", "0.001", "--embed_dim", "128", "--enc_hid", "256", "--dec_hid", "384", "--latent", "64",
         "--bs", "64", "--epochs", "7", "--prior", "GMM", "--c_v", "--dec_drop", "0.7", "--dec_lstm_drop", "0.8",
         "--gen_z_samples", "10", "--ann_param", "3", "--optimizer", "SGD", "--std", "0.25", "--temperature", "0.5"],
        ["--gpu", "1", "--mode", "inference", "--sample_gen", "greedy", "--checkpoint", "ck", "--gen_name", "xx",
         "--restore", "--no_encoder", "--fine_tune", "--save_params", "--prior", "AG", "--coco_dir", "/data/coco/"],
    ]
    parsed = []
    for argv in argvs:
        p = Parameters()
        with mock.patch.object(sys, "argv", ["main.py"] + argv), mock.patch.dict(os.environ, {}, clear=False):
            p.parse_args()
            env = os.environ.get("CUDA_VISIBLE_DEVICES")
        attrs = {k: getattr(p, k) for k in defaults}
        attrs["hdf5_file"] = p.hdf5_file
        parsed.append({"argv": argv, "attrs": attrs, "CUDA_VISIBLE_DEVICES": env})
    p = Parameters()
    err = None
    try:
        with mock.patch.object(sys, "argv", ["main.py"]):
            p.parse_args()
    except Exception as e:  # os.environ[...] = None raises TypeError (parameters.py:164)
        err = type(e).__name__
    dump("parameters.json", {"defaults": defaults, "parsed": parsed, "no_gpu_error": err})


def vocabulary_fixture():
    from utils.captions import Captions, Dictionary
    texts = ["A man holding a hot dog in his hand.", "a man  riding a wave on top of a surfboard",
             "Two dogs, one cat & a man's hat!", "a cat sitting on a hot laptop", "A dog. A cat. A man?",
             "the MAN and the dog"]
    tok = Captions._tokenize_caption(None, texts[0])
    toks = [Captions._tokenize_caption(None, t) for t in texts]
    caption_dict = {"img%d.jpg" % (i // 2): [] for i in range(len(texts))}
    for i, t in enumerate(toks):
        caption_dict["img%d.jpg" % (i // 2)].append(t)
    out = {"texts": texts, "tokens": toks, "first": tok, "vocab": {}}
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as td:
        os.makedirs(os.path.join(td, "pickles"))
        os.chdir(td)
        try:
            for keep in (1, 2, 3):
                with contextlib.redirect_stdout(io.StringIO()):
                    d = Dictionary(caption_dict, keep)
                out["vocab"][str(keep)] = {"word2idx": d.word2idx, "vocab_size": d.vocab_size,
                                           "indexed": [d.seq2dx([w if w in d.word2idx else "<UNK>" for w in t]) for t in toks]}
        finally:
            os.chdir(cwd)
    dump("vocabulary.json", out)


class _FakeDict(object):
    def __init__(self, V):
        self.idx2word = {i: "w%d" % i for i in range(V)}
        self.idx2word.update({0: "<PAD>", 1: "<BOS>", 2: "<EOS>"})
        self.word2idx = {w: i for i, w in self.idx2word.items()}
        self.vocab_size = V


class _FakeSession(object):
    """sess.run([sample, out_state], feed) -> HashModel.step(token, state)."""

    def __init__(self, decoder, models, image_key):
        self.decoder = decoder
        self.models = models
        self.image_key = image_key
        self.lengths_seen = []

    def run(self, fetches, feed):
        assert fetches == ["SAMPLE", "OUT_STATE"]
        tok = int(np.asarray(feed[self.decoder.captions]).reshape(-1)[0])
        assert np.asarray(feed[self.decoder.captions]).shape == (1, 1)
        self.lengths_seen.append(int(feed[self.decoder.lengths][0]))
        img = int(np.asarray(feed[self.image_key]).reshape(-1)[0])
        state = feed.get("IN_STATE")
        return self.models[img].step(tok, state)


def _import_reference_decoder():
    """decoder.py imports tensorflow / zhusuan at module level; only `tf.variable_scope` is touched by the
    host loops, so stub modules are enough to execute them unmodified."""
    tf = types.ModuleType("tensorflow")
    tf.variable_scope = lambda *a, **k: contextlib.nullcontext()
    tf.AUTO_REUSE = object()
    tf.layers = types.ModuleType("tensorflow.layers")
    tf.contrib = mock.MagicMock()
    tf.nn = mock.MagicMock()
    zs = types.ModuleType("zhusuan")
    with mock.patch.dict(sys.modules, {"tensorflow": tf, "tensorflow.layers": tf.layers, "zhusuan": zs}):
        import importlib
        for m in ("utils.rnn_model", "vae_model.decoder"):
            sys.modules.pop(m, None)
        dec = importlib.import_module("vae_model.decoder")
    return dec


def decode_fixture():
    dec_mod = _import_reference_decoder()
    from utils.parameters import Parameters
    cases = []
    for V, seed, n_img, max_len in ((23, 1, 4, 30), (57, 2, 3, 30), (11313, 3, 2, 30), (40, 4, 3, 8)):
        dd = _FakeDict(V)
        for mode, beam in (("greedy", 0), ("beam_search", 1), ("beam_search", 2), ("beam_search", 5), ("beam_search", 10),
                           ("pad_bug", 0)):
            p = Parameters()
            p.gen_max_len = max_len
            p.sample_gen = {"greedy": "greedy", "beam_search": "beam_search", "pad_bug": "beam_search"}[mode]
            p.temperature = 0.7
            d = dec_mod.Decoder("IMAGES_FV", "CAPTIONS_PH", "LENGTHS_PH", p, dd)
            d.px_z_fi = lambda observed, gen_mode=False: (None, None, None, ("IN_STATE", "OUT_STATE", "SAMPLE"))
            models = [HashModel(V, seed * 1000 + i) for i in range(n_img)]
            sess = _FakeSession(d, models, "IMAGE_F_INPUTS")
            pics = np.arange(n_img, dtype=np.float32).reshape(n_img, 1)
            ids = [100 + i for i in range(n_img)]
            if mode == "beam_search":
                caps = d.beam_search(sess, ids, pics, "IMAGE_F_INPUTS", beam_size=beam)
                beams = d.beam_search(sess, ids, pics, "IMAGE_F_INPUTS", beam_size=beam, ret_beams=True)
                cases.append({"V": V, "seed": seed, "n_img": n_img, "max_len": max_len, "mode": mode, "beam": beam,
                              "captions": caps, "ret_beams": beams})
            else:
                caps, raw = d.online_inference(sess, ids, pics, "IMAGE_F_INPUTS")
                cases.append({"V": V, "seed": seed, "n_img": n_img, "max_len": max_len, "mode": mode, "beam": beam,
                              "captions": caps, "raw": [[int(w) for w in r] for r in raw],
                              "lengths_fed": sess.lengths_seen[:max_len]})
    dump("decode_loops.json", cases)


def inference_json_fixture():
    from ops import inference as inf
    from utils.parameters import Parameters

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

        def beam_search(self, sess, ids, feats, ph, c_v, beam_size=2):
            self.calls.append(("beam_search", list(ids), None if c_v is None else np.asarray(c_v).shape[1], beam_size))
            return [{"image_id": i, "caption": "beam %s" % i} for i in ids]

        def online_inference(self, sess, ids, feats, ph, c_v=None):
            self.calls.append(("online_inference", list(ids), None if c_v is None else np.asarray(c_v).shape[1]))
            return [{"image_id": i, "caption": "greedy %s" % i} for i in ids], None

    out = []
    cwd = os.getcwd()
    for sample_gen, use_c_v, prior in (("beam_search", False, "Normal"), ("greedy", True, "Normal"), ("beam_search", False, "AG")):
        p = Parameters()
        p.sample_gen, p.use_c_v, p.prior, p.gen_name, p.beam_size = sample_gen, use_c_v, prior, "gg", 3
        d = Dec()
        with tempfile.TemporaryDirectory() as td:
            os.chdir(td)
            try:
                with contextlib.redirect_stdout(io.StringIO()):
                    inf.inference(p, d, Gen(), Gen(), "PH", mock.MagicMock(), "SESS")
                val = json.load(open("val_gg.json"))
                test = json.load(open("test_gg.json"))
            finally:
                os.chdir(cwd)
        out.append({"sample_gen": sample_gen, "use_c_v": use_c_v, "prior": prior, "val": val, "test": test,
                    "calls": d.calls})
    dump("inference_json.json", out)


if __name__ == "__main__":
    topn_fixture()
    caption_utils_fixture()
    parameters_fixture()
    vocabulary_fixture()
    decode_fixture()
    inference_json_fixture()
