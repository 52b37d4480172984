"""Host-side pieces either side of the hot path that need no GPU: checkpoint container (variables by TF name, the
VGG16 npz assignment order of utils/image_embeddings.py:240-246), the synthetic feeder's reference shapes, the
main.py feed construction (main.py:225-236), cluster-mean initialisation (utils/vae_utils.py:6-31)."""
import os

import numpy as np
import pytest

from vae_captioning_b200 import checkpoint, synthetic
from vae_captioning_b200.main import SyntheticFeeder, _feed
from vae_captioning_b200.parameters import Parameters


@pytest.mark.parametrize("suffix", ["", ".npz"])
def test_checkpoint_roundtrip_keeps_tf_names(tmp_path, suffix):
    state = {"imf_emb/kernel": np.arange(12, dtype=np.float32).reshape(3, 4),
             "decoder/net/multi_rnn_cell/cell_0/lstm_cell/bias": np.ones(8, np.float32),
             "cnn/conv5_1/weights_conv": np.zeros((3, 3, 2, 2), np.float32)}
    p = Parameters()
    p.checkpoint = "unit"
    path = checkpoint.checkpoint_path(p, str(tmp_path / "checkpoints"))
    assert path.endswith("checkpoints/unit.ckpt")  # main.py:211, 288
    assert checkpoint.checkpoint_path(p) == "./checkpoints/unit.ckpt"
    path += suffix
    assert checkpoint.save(path, state) == path
    if not suffix:  # Saver.save leaves the V2 bundle pair and the `checkpoint` state file
        assert sorted(os.listdir(str(tmp_path / "checkpoints"))) == ["checkpoint", "unit.ckpt.data-00000-of-00001", "unit.ckpt.index"]
    back = checkpoint.load(path)
    assert sorted(back) == sorted(state)
    for k in state:
        np.testing.assert_array_equal(back[k], state[k])
    with pytest.raises(FileNotFoundError):
        checkpoint.load(str(tmp_path / ("missing" + suffix)))


@pytest.mark.parametrize("suffix", [".ckpt", ".npz"])
def test_restore_checks_names_and_shapes(tmp_path, suffix):
    class FakeEngine(object):
        def __init__(self):
            self.got = {}

        def variables(self):
            return [("a/kernel", (2, 3), True), ("a/bias", (3,), True)]

        def set_variable(self, n, v):
            self.got[n] = v

    path = str(tmp_path / ("c" + suffix))
    checkpoint.save(path, {"a/kernel": np.zeros((2, 3)), "a/bias": np.zeros(3)})
    e = FakeEngine()
    checkpoint.restore(e, path)
    assert sorted(e.got) == ["a/bias", "a/kernel"]
    checkpoint.save(path, {"a/kernel": np.zeros((2, 3))})
    with pytest.raises(KeyError):
        checkpoint.restore(FakeEngine(), path)
    checkpoint.restore(FakeEngine(), path, strict=False)
    checkpoint.save(path, {"a/kernel": np.zeros((3, 2)), "a/bias": np.zeros(3)})
    with pytest.raises(ValueError):
        checkpoint.restore(FakeEngine(), path)


def test_vgg16_npz_assignment_order(tmp_path):
    """load_weights sorts the npz keys and assigns them to the graph parameters in creation order; fc8 is dropped."""
    names = ["conv%d_%d" % (b, i) for b, n in ((1, 2), (2, 2), (3, 3), (4, 3), (5, 3)) for i in range(1, n + 1)]
    arrays = {}
    for j, n in enumerate(names + ["fc6", "fc7", "fc8"]):
        arrays[n + "_W"] = np.full((1,), 2 * j, np.float32)
        arrays[n + "_b"] = np.full((1,), 2 * j + 1, np.float32)
    path = str(tmp_path / "vgg16_weights.npz")
    np.savez(path, **arrays)
    st = checkpoint.vgg16_npz_state(path)
    assert len(st) == 30 and list(st) == checkpoint.VGG_VARIABLES
    assert st["cnn/conv1_1/weights"][0] == 0 and st["cnn/conv1_1/biases"][0] == 1
    assert st["cnn/conv5_3/weights_conv"][0] == 24 and st["cnn/conv5_3/biases_conv"][0] == 25
    assert st["cnn/fc1/weights"][0] == 26 and st["cnn/fc2/biases"][0] == 29  # fc6 -> fc1, fc7 -> fc2


def test_synthetic_feeder_and_feed_shapes():
    p = Parameters()
    p.batch_size, p.num_captions, p.use_c_v, p.prior = 4, 5, True, "AG"
    f = SyntheticFeeder(p, vocab_size=50, batches=2, T=7)
    batches = list(f.next_batch(use_obj_vectors=True, num_captions=5))
    assert len(batches) == 2
    feats, (inp, lbl), lens, c_v = batches[0]
    assert feats.shape == (4, 4096) and inp.shape == (4, 5, 7) and lbl.shape == (4, 5, 7)
    assert lens.shape == (4, 5) and lens.dtype == np.float64 and c_v.shape == (4, 91)
    feed = _feed(p, feats, (inp, lbl), lens, c_v)
    assert feed["ann_inputs_enc"].shape == (20, 7) and feed["ann_lengths"].shape == (20,)
    assert feed["c_i"].shape == (20, 90)  # column 0 dropped (main.py:236), rows repeated per caption
    np.testing.assert_array_equal(feed["c_i"][0], feed["c_i"][4])
    # the encoder consumes the label sequence, the decoder the <BOS>-led input sequence (Q8)
    live = feed["ann_lengths"] > 0  # images with fewer than C captions yield empty rows (batch_gen.py:313-317)
    assert (feed["ann_inputs_dec"][live, 0] == 1).all() and (feed["ann_inputs_dec"][~live] == 0).all()
    ids = [i for *_, i, _ in f.next_val_batch(get_image_ids=True)]
    assert ids[0] == [0, 1, 2, 3] and ids[1] == [4, 5, 6, 7]


def test_init_clusters_unit_norm_and_persistence(tmp_path):
    m = synthetic.init_clusters(90, 150)
    assert m.shape == (90, 150) and m.dtype == np.float32
    np.testing.assert_allclose(np.linalg.norm(m, axis=1), 1.0, rtol=1e-5)
    fn = str(tmp_path / "pickles" / "cluster_means.pickle")
    a = synthetic.init_clusters(5, 8, c_m_file=fn)
    assert os.path.exists(fn)
    b = synthetic.init_clusters(5, 8, seed=99, c_m_file=fn)  # re-read, not regenerated
    np.testing.assert_array_equal(a, b)
