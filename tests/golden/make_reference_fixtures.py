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


def _jsonable(x):
    if isinstance(x, np.ndarray):
        return {"shape": list(x.shape), "dtype": str(x.dtype), "data": x.tolist()}
    if isinstance(x, (tuple, list)):
        return [_jsonable(v) for v in x]
    if isinstance(x, (np.integer,)):
        return int(x)
    if isinstance(x, (np.floating,)):
        return float(x)
    return x


def batch_gen_fixture():
    """utils/captions.py Captions + Dictionary and utils/batch_gen.py Batch_Generator, unmodified, over the miniature
    COCO tree of tests/golden/fake_coco.py. tensorflow / h5py are imported by batch_gen.py at module level but only
    h5py.File is used (for the uint8 image store), so stubs are enough; `glob` is wrapped to return sorted lists so
    the file order does not depend on the file system."""
    import glob as glob_mod
    import importlib
    import random
    sys.path.insert(0, HERE)
    import fake_coco
    h5 = types.ModuleType("h5py")

    class _H5File(dict):
        def __init__(self, path, mode="r"):
            dict.__init__(self, images=np.load(path))
    h5.File = _H5File
    tf = mock.MagicMock()
    out = {}
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as td, mock.patch.dict(sys.modules, {"tensorflow": tf, "h5py": h5}):
        names = fake_coco.build(td)
        os.chdir(td)
        try:
            for m in ("utils.batch_gen", "utils.captions"):
                sys.modules.pop(m, None)
            bg = importlib.import_module("utils.batch_gen")
            cp = importlib.import_module("utils.captions")
            bg.glob = lambda pat: sorted(glob_mod.glob(pat))
            coco = os.path.join(td, "coco") + "/"
            sink = io.StringIO()
            with contextlib.redirect_stdout(sink):
                cap_tr = cp.Captions(coco + "annotations/captions_train2014.json", 100)
                cap_val = cp.Captions(coco + "annotations/captions_val2014.json", 100)
                raw_tr = {k: [list(c) for c in v] for k, v in cap_tr.captions.items()}
                d = cp.Dictionary(cap_tr.captions, 2)
                cap_tr.index_captions(d.word2idx)
                cap_val.index_captions(d.word2idx)
            out["captions"] = {"raw_train": raw_tr, "word2idx": d.word2idx, "indexed_train": dict(cap_tr.captions_indexed),
                               "indexed_val": dict(cap_val.captions_indexed), "fn_to_id": cap_tr.filename_to_imid,
                               "num_captions": cap_tr.num_captions}
            fd_tr = fake_coco.feature_dict(names["train"], 5)
            fd_val = fake_coco.feature_dict(names["val"], 6)
            fd_test = fake_coco.feature_dict(names["test"], 7)
            fake_coco.image_store(td, names["train"] + names["val"])
            tr_dir, val_dir, test_dir = (coco + "images/%s2014/" % s for s in ("train", "val", "test"))
            runs = []

            def record(tag, gen_factory, method, kwargs, seed):
                random.seed(seed)
                with contextlib.redirect_stdout(sink):
                    g = gen_factory()
                    batches = [_jsonable(b) for b in getattr(g, method)(**kwargs)]
                runs.append({"tag": tag, "method": method, "kwargs": kwargs, "seed": seed, "batches": batches,
                             "unused_cap_in": [n.split("/")[-1] for n in (g.unused_cap_in or [])]})

            feats = lambda bs: bg.Batch_Generator(tr_dir, coco + "annotations/captions_train2014.json", cap_tr, bs,
                                                  feature_dict=fd_tr)
            record("train_feats_bs3_c1", lambda: feats(3), "next_batch", {"use_obj_vectors": False, "num_captions": 1}, 1)
            record("train_feats_bs3_c5_cv", lambda: feats(3), "next_batch", {"use_obj_vectors": True, "num_captions": 5}, 2)
            record("train_feats_bs4_c2", lambda: feats(4), "next_batch", {"use_obj_vectors": False, "num_captions": 2}, 3)
            record("train_feats_all", lambda: feats(None), "next_batch", {"use_obj_vectors": True, "num_captions": 1}, 4)

            def store_gen():
                return bg.Batch_Generator(tr_dir, coco + "annotations/captions_train2014.json", cap_tr, 3, use_hdf5=True,
                                          hdf5_file=os.path.join(td, "store.npy"), feature_dict=None)
            record("train_store_bs3_c5", store_gen, "next_batch", {"use_obj_vectors": True, "num_captions": 5}, 5)

            def repart():
                g = feats(4)
                g.repartiton(cap_val, fd_val, 2)
                return g
            record("train_repartition_bs4_c3", repart, "next_batch", {"use_obj_vectors": True, "num_captions": 3}, 6)

            def val_gen():
                return bg.Batch_Generator(val_dir, coco + "annotations/captions_val2014.json", cap_val, 2,
                                          feature_dict=fd_val, get_image_ids=True)
            record("val_ids_cv", val_gen, "next_val_batch", {"get_image_ids": True, "use_obj_vectors": True}, 7)
            record("val_noids", val_gen, "next_val_batch", {"get_image_ids": False, "use_obj_vectors": False}, 8)
            unused = [val_dir + n for n in names["val"][-2:]]
            record("val_unused_only", lambda: bg.Batch_Generator(val_dir, coco + "annotations/captions_val2014.json", cap_val,
                                                                 None, feature_dict=fd_val, get_image_ids=True,
                                                                 val_tr_unused=list(unused)),
                   "next_val_batch", {"get_image_ids": True, "use_obj_vectors": False}, 9)

            def test_gen():
                return bg.Batch_Generator(test_dir, train_cap_json=coco + "annotations/image_info_test2014.json", batch_size=2,
                                          feature_dict=fd_test, get_image_ids=True, get_test_ids=True)
            record("test_cv", test_gen, "next_test_batch", {"use_obj_vectors": True}, 10)
            record("test_nocv", test_gen, "next_test_batch", {"use_obj_vectors": False}, 11)
            out["runs"] = runs
            errs = {}
            for tag, fn in (("empty_dir", lambda: bg.Batch_Generator(os.path.join(td, "nowhere") + "/", batch_size=2)),
                            ("hdf5_no_file", lambda: bg.Batch_Generator(tr_dir, use_hdf5=True)),
                            ("repartition_no_caps", lambda: feats(2).repartiton(None, fd_val, 2)),
                            ("repartition_no_feats", lambda: feats(2).repartiton(cap_val, None, 2))):
                try:
                    with contextlib.redirect_stdout(sink):
                        fn()
                    errs[tag] = None
                except Exception as e:
                    errs[tag] = type(e).__name__
            out["errors"] = errs
            # utils/data.py Data over the same tree, features served from ./pickles/{split}2014.pickle (so the TF graph in
            # extract_features_from_dir is never built); repartition on, like main.py:22-36.
            import pickle
            for split, fd in (("train", fd_tr), ("val", fd_val), ("test", fd_test)):
                with open("./pickles/%s2014.pickle" % split, "wb") as wf:
                    pickle.dump(fd, wf)
            for m in ("utils.data", "utils.image_embeddings", "utils.parameters"):
                sys.modules.pop(m, None)
            dm = importlib.import_module("utils.data")
            dm.Batch_Generator = bg.Batch_Generator
            from utils.parameters import Parameters
            p = Parameters()
            p.coco_dir, p.use_hdf5, p.keep_words, p.cap_max_length = coco, False, 2, 100
            droot = {}
            random.seed(12)
            with contextlib.redirect_stdout(sink):
                data = dm.Data(p, True, "weights.npz", repartiton=True, gen_val_cap=2)
                tr = data.load_train_data_generator(3)
                droot["num_examples"] = data.num_examples
                droot["vocab_size"] = data.dictionary.vocab_size
                droot["unused_cap_in"] = [n.split("/")[-1] for n in tr.unused_cap_in]
                droot["train"] = [_jsonable(b) for b in tr.next_batch(use_obj_vectors=True, num_captions=2)]
                va = data.get_valid_data(2, val_tr_unused=tr.unused_cap_in)
                droot["val"] = [_jsonable(b) for b in va.next_val_batch(get_image_ids=True, use_obj_vectors=True)]
                te = data.get_test_data(2)
                droot["test"] = [_jsonable(b) for b in te.next_test_batch(use_obj_vectors=True)]
            err = None
            try:
                with contextlib.redirect_stdout(sink):
                    dm.Data(p, False, None, repartiton=True, gen_val_cap=None)
            except Exception as e:
                err = type(e).__name__
            droot["repartition_without_count"] = err
            out["data"] = droot
        finally:
            os.chdir(cwd)
    dump("batch_gen.json", out)


if __name__ == "__main__":
    batch_gen_fixture()
    topn_fixture()
    caption_utils_fixture()
    parameters_fixture()
    vocabulary_fixture()
    decode_fixture()
    inference_json_fixture()
