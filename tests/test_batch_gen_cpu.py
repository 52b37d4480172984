"""Host batcher parity: this repo's Captions / Dictionary / Batch_Generator replayed against
tests/golden/batch_gen.json, which holds what the REFERENCE's utils/captions.py and utils/batch_gen.py yielded on the
same miniature COCO tree (tests/golden/fake_coco.py) with the same `random` / `numpy.random` seeds."""
import contextlib
import glob as glob_mod
import io
import json
import os
import random
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import fake_coco  # noqa: E402

from vae_captioning_b200 import batch_gen as bg  # noqa: E402
from vae_captioning_b200.captions import Captions, Dictionary  # noqa: E402

GOLD = json.load(open(os.path.join(HERE, "golden", "batch_gen.json")))


def _same(got, want, where=""):
    if isinstance(want, dict) and set(want) == {"shape", "dtype", "data"}:
        got = np.asarray(got)
        assert list(got.shape) == want["shape"], where
        assert str(got.dtype) == want["dtype"], where
        np.testing.assert_array_equal(got, np.asarray(want["data"], dtype=got.dtype).reshape(want["shape"]), err_msg=where)
    elif isinstance(want, list):
        assert isinstance(got, (list, tuple)) and len(got) == len(want), where
        for i, (g, w) in enumerate(zip(got, want)):
            _same(g, w, "%s[%d]" % (where, i))
    else:
        assert got == want, where


@pytest.fixture()
def tree(tmp_path, monkeypatch):
    names = fake_coco.build(str(tmp_path))
    monkeypatch.chdir(tmp_path)
    monkeypatch.setattr(bg.glob, "glob", lambda pat, _g=glob_mod.glob: sorted(_g(pat)))
    coco = str(tmp_path / "coco") + "/"
    with contextlib.redirect_stdout(io.StringIO()):
        cap_tr = Captions(coco + "annotations/captions_train2014.json", 100)
        cap_val = Captions(coco + "annotations/captions_val2014.json", 100)
        raw_tr = {k: [list(c) for c in v] for k, v in cap_tr.captions.items()}
        d = Dictionary(cap_tr.captions, 2)
        cap_tr.index_captions(d.word2idx)
        cap_val.index_captions(d.word2idx)
    return dict(root=str(tmp_path), coco=coco, names=names, cap_tr=cap_tr, cap_val=cap_val, raw_tr=raw_tr, d=d)


def test_captions_and_vocabulary_match_reference(tree):
    g = GOLD["captions"]
    assert tree["raw_tr"] == g["raw_train"]
    assert tree["d"].word2idx == g["word2idx"]
    assert dict(tree["cap_tr"].captions_indexed) == g["indexed_train"]
    assert dict(tree["cap_val"].captions_indexed) == g["indexed_val"]
    assert tree["cap_tr"].filename_to_imid == g["fn_to_id"]
    assert tree["cap_tr"].num_captions == g["num_captions"]


def _factories(t):
    coco, names, root = t["coco"], t["names"], t["root"]
    fd_tr, fd_val, fd_test = (fake_coco.feature_dict(names[s], k) for s, k in (("train", 5), ("val", 6), ("test", 7)))
    fake_coco.image_store(root, names["train"] + names["val"])
    tr_dir, val_dir, test_dir = (coco + "images/%s2014/" % s for s in ("train", "val", "test"))
    tr_json, val_json = coco + "annotations/captions_train2014.json", coco + "annotations/captions_val2014.json"

    def feats(bs):
        return bg.Batch_Generator(tr_dir, tr_json, t["cap_tr"], bs, feature_dict=fd_tr)

    def repart():
        g = feats(4)
        g.repartiton(t["cap_val"], fd_val, 2)
        return g

    def val_gen():
        return bg.Batch_Generator(val_dir, val_json, t["cap_val"], 2, feature_dict=fd_val, get_image_ids=True)

    def test_gen():
        return bg.Batch_Generator(test_dir, train_cap_json=coco + "annotations/image_info_test2014.json", batch_size=2,
                                  feature_dict=fd_test, get_image_ids=True, get_test_ids=True)

    unused = [val_dir + n for n in names["val"][-2:]]
    return {
        "train_feats_bs3_c1": lambda: feats(3), "train_feats_bs3_c5_cv": lambda: feats(3), "train_feats_bs4_c2": lambda: feats(4),
        "train_feats_all": lambda: feats(None),
        "train_store_bs3_c5": lambda: bg.Batch_Generator(tr_dir, tr_json, t["cap_tr"], 3, use_hdf5=True,
                                                         hdf5_file=os.path.join(root, "store.npy"), feature_dict=None),
        "train_repartition_bs4_c3": repart, "val_ids_cv": val_gen, "val_noids": val_gen,
        "val_unused_only": lambda: bg.Batch_Generator(val_dir, val_json, t["cap_val"], None, feature_dict=fd_val,
                                                      get_image_ids=True, val_tr_unused=list(unused)),
        "test_cv": test_gen, "test_nocv": test_gen,
    }, dict(feats=feats, fd_val=fd_val, tr_dir=tr_dir)


@pytest.mark.parametrize("tag", [r["tag"] for r in GOLD["runs"]])
def test_generators_yield_what_the_reference_yields(tree, tag):
    run = [r for r in GOLD["runs"] if r["tag"] == tag][0]
    factories, _ = _factories(tree)
    random.seed(run["seed"])
    with contextlib.redirect_stdout(io.StringIO()):
        g = factories[tag]()
        batches = list(getattr(g, run["method"])(**run["kwargs"]))
    assert len(batches) == len(run["batches"])
    for i, (got, want) in enumerate(zip(batches, run["batches"])):
        _same(got, want, "%s batch %d" % (tag, i))
    assert [n.split("/")[-1] for n in (g.unused_cap_in or [])] == run["unused_cap_in"]


def test_error_behaviour_matches_reference(tree):
    _, h = _factories(tree)
    got = {}
    for tag, fn in (("empty_dir", lambda: bg.Batch_Generator(os.path.join(tree["root"], "nowhere") + "/", batch_size=2)),
                    ("hdf5_no_file", lambda: bg.Batch_Generator(h["tr_dir"], use_hdf5=True)),
                    ("repartition_no_caps", lambda: h["feats"](2).repartiton(None, h["fd_val"], 2)),
                    ("repartition_no_feats", lambda: h["feats"](2).repartiton(tree["cap_val"], None, 2))):
        try:
            with contextlib.redirect_stdout(io.StringIO()):
                fn()
            got[tag] = None
        except Exception as e:
            got[tag] = type(e).__name__
    assert got == GOLD["errors"]


def test_prefetcher_preserves_order_and_errors(tree):
    factories, _ = _factories(tree)
    with contextlib.redirect_stdout(io.StringIO()):
        random.seed(1)
        plain = list(factories["train_feats_bs3_c1"]().next_batch(num_captions=1))
        random.seed(1)
        ahead = list(bg.Prefetcher(factories["train_feats_bs3_c1"]().next_batch(num_captions=1), depth=2))
    assert len(plain) == len(ahead) == 3
    for a, b in zip(plain, ahead):
        np.testing.assert_array_equal(a[0], b[0])
        np.testing.assert_array_equal(a[1][0], b[1][0])

    def boom():
        yield 1
        raise KeyError("missing caption")
    it = iter(bg.Prefetcher(boom()))
    assert next(it) == 1
    with pytest.raises(KeyError):
        next(it)


def test_image_files_path_and_load_image(tmp_path, monkeypatch):
    """Without a feature dict or image store the generator decodes the jpg files: uint8 RGB [B, 224, 224, 3]."""
    cv2 = pytest.importorskip("cv2")
    names = fake_coco.build(str(tmp_path), n_train=3, n_val=1, n_test=1, write_images=True)
    monkeypatch.chdir(tmp_path)
    monkeypatch.setattr(bg.glob, "glob", lambda pat, _g=glob_mod.glob: sorted(_g(pat)))
    coco = str(tmp_path / "coco") + "/"
    with contextlib.redirect_stdout(io.StringIO()):
        cap = Captions(coco + "annotations/captions_train2014.json")
        d = Dictionary(cap.captions, 1)
        cap.index_captions(d.word2idx)
        g = bg.Batch_Generator(coco + "images/train2014/", coco + "annotations/captions_train2014.json", cap, 2)
        random.seed(0)
        batches = list(g.next_batch(num_captions=2))
    assert [b[0].shape for b in batches] == [(2, 224, 224, 3), (1, 224, 224, 3)]
    assert batches[0][0].dtype == np.uint8
    from vae_captioning_b200.image_utils import load_image
    p = coco + "images/train2014/" + names["train"][0]
    want = cv2.cvtColor(cv2.resize(cv2.imread(p), (224, 224)), cv2.COLOR_BGR2RGB)
    np.testing.assert_array_equal(load_image(p), want)
    with pytest.raises(FileNotFoundError):
        load_image(str(tmp_path / "nope.jpg"))


def test_data_front_end_matches_reference(tree, monkeypatch):
    """utils/data.py Data (captions -> vocabulary -> generators for the three splits, repartition) with the feature
    pickles present, against what the reference's Data yielded on the same tree."""
    import pickle
    from vae_captioning_b200 import data as dm
    from vae_captioning_b200.parameters import Parameters
    names, coco = tree["names"], tree["coco"]
    for split, k in (("train", 5), ("val", 6), ("test", 7)):
        with open("./pickles/%s2014.pickle" % split, "wb") as wf:
            pickle.dump(fake_coco.feature_dict(names[split], k), wf)
    p = Parameters()
    p.coco_dir, p.use_hdf5, p.keep_words, p.cap_max_length = coco, False, 2, 100
    g = GOLD["data"]
    random.seed(12)
    with contextlib.redirect_stdout(io.StringIO()):
        data = dm.Data(p, True, "weights.npz", repartiton=True, gen_val_cap=2)
        tr = data.load_train_data_generator(3)
        assert data.num_examples == g["num_examples"] and data.dictionary.vocab_size == g["vocab_size"]
        assert [n.split("/")[-1] for n in tr.unused_cap_in] == g["unused_cap_in"]
        _same(list(tr.next_batch(use_obj_vectors=True, num_captions=2)), g["train"], "train")
        va = data.get_valid_data(2, val_tr_unused=tr.unused_cap_in)
        _same(list(va.next_val_batch(get_image_ids=True, use_obj_vectors=True)), g["val"], "val")
        te = data.get_test_data(2)
        _same(list(te.next_test_batch(use_obj_vectors=True)), g["test"], "test")
    err = None
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            dm.Data(p, False, None, repartiton=True, gen_val_cap=None)
    except Exception as e:
        err = type(e).__name__
    assert err == g["repartition_without_count"]
    with pytest.raises(ValueError):  # data.py:48-50: extracting features needs the ImageNet weights
        with contextlib.redirect_stdout(io.StringIO()):
            dm.Data(p, True, None)
