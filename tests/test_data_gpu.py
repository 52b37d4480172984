"""The callers either side of the device path, end to end on the GPU:
  * Data.extract_features_from_dir / data.extract_features: jpg files -> batched uint8 VGG16 forward -> the
    reference's {file name: float32 [1, 4096]} dictionary and ./pickles/{split}.pickle cache (utils/data.py:86-130);
  * gen_caption.Generator: params pickle + vocabulary pickle + TF-bundle checkpoint -> caption of one image
    (gen_caption.py:19-160), incl. ret_beams and the cluster-vector hook."""
import contextlib
import io
import os
import pickle
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import fake_coco  # noqa: E402

from vae_captioning_b200 import checkpoint, synthetic  # noqa: E402
from vae_captioning_b200.image_utils import load_image  # noqa: E402
from vae_captioning_b200.parameters import Parameters  # noqa: E402

pytestmark = pytest.mark.gpu


def small_params(**kw):
    p = Parameters()
    p.embed_size, p.encoder_hidden, p.decoder_hidden, p.latent_size, p.gen_z_samples = 64, 64, 128, 8, 4
    p.batch_size, p.num_epochs, p.gen_max_len, p.beam_size, p.keep_words = 2, 1, 8, 3, 1
    p.checkpoint, p.gen_name, p.use_hdf5 = "unit", "unit", False
    for k, v in kw.items():
        setattr(p, k, v)
    return p


@pytest.fixture(scope="module")
def vgg_weights():
    from vae_captioning_b200.engine import Engine
    p = small_params()
    e = Engine(p, vocab_size=50, max_batch=4, max_len=4, with_cnn=True)
    w = {k: v for k, v in synthetic.init_weights(e.variables(), seed=1).items() if k.startswith("cnn/")}
    for k in w:  # non-zero biases so the bias path is exercised
        if w[k].ndim == 1:
            w[k] = (0.01 * np.arange(w[k].size, dtype=np.float32) / w[k].size)
    e.close()
    return w


def test_extract_features_batched_and_cached(tmp_path, monkeypatch, vgg_weights):
    from vae_captioning_b200 import data as dm
    from vae_captioning_b200.engine import Engine
    names = fake_coco.build(str(tmp_path), n_train=5, n_val=2, n_test=1, write_images=True)
    monkeypatch.chdir(tmp_path)
    coco = str(tmp_path / "coco") + "/"
    p = small_params(coco_dir=coco)
    eng = Engine(p, vocab_size=50, max_batch=4, max_len=4, with_cnn=True)
    eng.load_state(vgg_weights)
    paths = sorted(coco + "images/train2014/" + n for n in names["train"])
    seen = []
    fd = dm.extract_features(eng, paths, batch_size=2, progress=lambda done, total: seen.append((done, total)))
    assert list(fd) == [q.split("/")[-1] for q in paths] and seen[-1] == (5, 5) and len(seen) == 3
    for q in paths:  # batched (2 + 2 + 1 images per call) == one image per call, up to split-K summation order
        f = fd[q.split("/")[-1]]
        assert f.shape == (1, 4096) and f.dtype == np.float32
        one = eng.vgg_forward(load_image(q)[None])
        scale = max(1e-6, float(np.abs(one).max()))
        assert float(np.abs(f - one).max()) <= 2e-3 * scale
        assert float(np.abs(one).max()) > 0
    with pytest.raises(ValueError):
        dm.extract_features(eng, paths, batch_size=8)
    # Data: vocabulary + generators; features of all three splits extracted once, then served from ./pickles
    with contextlib.redirect_stdout(io.StringIO()) as log:
        data = dm.Data(p, True, "unused.npz", engine=eng)
        gen = data.load_train_data_generator(2)
        val = data.get_valid_data(2)
    assert "Extracting features" in log.getvalue()
    assert sorted(os.listdir("pickles")) == ["capt_vocab.pickle", "train2014.pickle", "val2014.pickle"]
    cached = pickle.load(open("pickles/train2014.pickle", "rb"))
    np.testing.assert_array_equal(cached[names["train"][0]], data.train_feature_dict[names["train"][0]])
    batch = next(iter(gen.next_batch(num_captions=2)))
    assert batch[0].shape == (2, 4096) and batch[1][0].shape[:2] == (2, 2)
    vb = next(iter(val.next_val_batch(get_image_ids=True)))
    assert vb[0].shape == (2, 4096) and len(vb[3]) == 2
    with contextlib.redirect_stdout(io.StringIO()) as log:
        again = dm.Data(p, True, "unused.npz", engine=None)  # no engine needed: the cache is read
    assert "Loading prepared feature vector" in log.getvalue()
    np.testing.assert_array_equal(again.train_feature_dict[names["train"][1]], cached[names["train"][1]])
    eng.close()


@pytest.mark.parametrize("use_c_v", [False, True])
def test_gen_caption_generator(tmp_path, monkeypatch, vgg_weights, use_c_v):
    from vae_captioning_b200 import gen_caption as G
    from vae_captioning_b200.captions import Captions, Dictionary
    from vae_captioning_b200.decode import Decoder
    from vae_captioning_b200.engine import Engine
    names = fake_coco.build(str(tmp_path), n_train=3, n_val=1, n_test=1, write_images=True)
    monkeypatch.chdir(tmp_path)
    coco = str(tmp_path / "coco") + "/"
    with contextlib.redirect_stdout(io.StringIO()):
        cap = Captions(coco + "annotations/captions_train2014.json")
        d = Dictionary(cap.captions, 1)  # writes ./pickles/capt_vocab.pickle like the training run does
    p = small_params(use_c_v=use_c_v)
    p.vocab_size = d.vocab_size
    eng = Engine(p, vocab_size=d.vocab_size, max_batch=1, max_len=8, with_cnn=True)
    state = synthetic.init_weights(eng.variables(), seed=3)
    state.update(vgg_weights)
    eng.load_state(state)
    ck = checkpoint.save(checkpoint.checkpoint_path(p), eng.state())  # TF bundle incl. the cnn/ variables
    with open("./pickles/params.pickle", "wb") as wf:
        pickle.dump(p, wf)
    img = coco + "images/train2014/" + names["train"][0]
    cv = np.zeros(91, np.float32)
    cv[[5, 17]] = 0.5
    kw = dict(c_v_generator=(lambda image: cv)) if use_c_v else {}
    # expected: the same device path driven by hand
    feat = eng.vgg_forward(load_image(img)[None])
    dec = Decoder(eng, p, d)
    c_v = cv[None, 1:] if use_c_v else None
    want_greedy, _ = dec.online_inference([names["train"][0]], feat, c_v=c_v, sample_gen="greedy")
    want_beam = dec.beam_search([names["train"][0]], feat, c_v, beam_size=3, ret_beams=True)
    eng.close()
    g = G.Generator(ck, "./pickles/params.pickle", "./pickles/capt_vocab.pickle", gen_method="greedy", **kw)
    assert g.data_dict.word2idx == d.word2idx
    got = g.generate_caption(img)
    assert got == want_greedy and isinstance(got[0]["caption"], str) and got[0]["image_id"] == names["train"][0]
    g.close()
    g = G.Generator(ck, "./pickles/params.pickle", "./pickles/capt_vocab.pickle", gen_method="beam_search", **kw)
    beams = g.generate_caption(img, beam_size=3, ret_beams=True)
    assert beams == want_beam and isinstance(beams[0]["caption"], list)
    assert g.generate_caption(img, beam_size=3)[0]["caption"] == want_beam[0]["caption"][0]
    with pytest.raises(ValueError):
        g.generate_caption(str(tmp_path / "missing.jpg"))
    g.close()
    if use_c_v:
        g = G.Generator(ck, "./pickles/params.pickle", "./pickles/capt_vocab.pickle")
        with pytest.raises(ValueError):
            g.generate_caption(img)
        g.close()
    with pytest.raises(ValueError):
        G.Generator(ck, "./pickles/params.pickle", "./pickles/nope.pickle")
