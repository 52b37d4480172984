"""The main.py-compatible driver end to end on the device: train one epoch on the synthetic feeder (reference prints,
validation, checkpoint save by TF variable name), restore, then `--mode inference` writing val_/test_ json files in
the COCO-caption result format (ops/inference.py:32-56)."""
import json
import os

import numpy as np
import pytest

from vae_captioning_b200 import checkpoint
from vae_captioning_b200 import main as M
from vae_captioning_b200.parameters import Parameters

pytestmark = pytest.mark.gpu


def small_params(**kw):
    p = Parameters()
    p.embed_size, p.encoder_hidden, p.decoder_hidden, p.latent_size, p.gen_z_samples = 64, 64, 128, 8, 4
    p.batch_size, p.num_epochs, p.vocab_size, p.gen_max_len, p.beam_size = 2, 1, 120, 8, 3
    p.checkpoint, p.gen_name = "unit", "unit"
    for k, v in kw.items():
        setattr(p, k, v)
    return p


@pytest.mark.parametrize("kw", [dict(), dict(prior="AG", use_c_v=True)])
def test_train_checkpoint_inference_roundtrip(tmp_path, monkeypatch, kw):
    monkeypatch.chdir(tmp_path)
    lines = []
    p = small_params(**kw)
    feeder = M.SyntheticFeeder(p, p.vocab_size, batches=3, T=6)
    eng = M.run(p, feeder=feeder, out=lines.append, max_len=8, report_every=2)
    assert any(l.startswith("Epoch: 0 Iteration: 2 VLB:") for l in lines)
    assert any(l.startswith("Validation reconstruction loss:") for l in lines)
    assert lines[-1] == "Model saved in file: ./checkpoints/unit.ckpt"
    saved = checkpoint.load("./checkpoints/unit.ckpt")
    assert sorted(saved) == sorted(n for n, _, _ in eng.variables())
    w0 = eng.get_variable("decoder/rnn_logits/kernel")
    np.testing.assert_array_equal(saved["decoder/rnn_logits/kernel"], w0)
    eng.close()
    # --restore continues from the saved variables; --mode inference decodes val/test with them
    p2 = small_params(restore=True, **kw)
    eng2 = M.run(p2, feeder=feeder, out=lines.append, max_len=8)
    assert "Restoring from checkpoint" in lines
    eng2.close()
    p3 = small_params(mode="inference", sample_gen="greedy", **kw)
    eng3 = M.run(p3, feeder=feeder, out=lines.append, max_len=8)
    eng3.close()
    for name in ("val_unit.json", "test_unit.json"):
        caps = json.load(open(name))
        assert len(caps) == 3 * p.batch_size and set(caps[0]) == {"image_id", "caption"}
        assert isinstance(caps[0]["caption"], str)


def test_ag_inference_uses_the_cluster_means_file(tmp_path, monkeypatch):
    """with --prior AG the generation prior's mean is built from c_means (decoder.py:45-71), which the
    reference creates in both modes and keeps in ./pickles/cluster_means.pickle. The inference driver must load them:
    captions decoded with two different cluster-means files differ, and run() leaves the file behind when it made it."""
    import pickle
    monkeypatch.chdir(tmp_path)
    p = small_params(prior="AG", use_c_v=True, latent_size=8, std=0.0001)
    feeder = M.SyntheticFeeder(p, p.vocab_size, batches=2, T=6)
    M.run(p, feeder=feeder, out=lambda *_: None, max_len=8).close()
    assert os.path.exists(M.CLUSTER_MEANS_FILE)
    outs = []
    for scale in (1.0, -40.0):
        means = np.asarray(pickle.load(open(M.CLUSTER_MEANS_FILE, "rb")), dtype=np.float32)
        with open(M.CLUSTER_MEANS_FILE, "wb") as f:
            pickle.dump(means * scale, f)
        M.run(small_params(mode="inference", sample_gen="greedy", prior="AG", use_c_v=True, latent_size=8, std=0.0001),
              feeder=feeder, out=lambda *_: None, max_len=8).close()
        outs.append(json.load(open("val_unit.json")))
    assert [c["caption"] for c in outs[0]] != [c["caption"] for c in outs[1]]


def test_images_through_the_decoder_after_fine_tune():
    """generation after --fine_tune feeds raw images; the decoder runs them through the engine's VGG16
    first and gives the same captions as decoding the fc2 features directly."""
    from test_decode_gpu import FakeDict
    from vae_captioning_b200 import synthetic
    from vae_captioning_b200.decode import Decoder
    from vae_captioning_b200.engine import Engine
    p = small_params(fine_tune=True, mode="inference")
    eng = Engine(p, vocab_size=p.vocab_size, max_batch=2, max_len=8)
    eng.load_state(synthetic.init_weights(eng.variables(), seed=4))
    dec = Decoder(eng, p, FakeDict(p.vocab_size))
    g = np.random.Generator(np.random.PCG64(5))
    images = g.integers(0, 256, size=(3, 224, 224, 3), dtype=np.uint8)  # more images than max_batch: chunked VGG forward
    feats = np.concatenate([eng.vgg_forward(images[i:i + 2]) for i in range(0, 3, 2)])
    t_img, l_img = dec.greedy_tokens(images, None, "greedy", {"seed": 3})
    t_ft, l_ft = dec.greedy_tokens(feats, None, "greedy", {"seed": 3})
    assert np.array_equal(t_img, t_ft) and np.array_equal(l_img, l_ft)
    with pytest.raises(ValueError):
        dec.greedy_tokens(images[:, :100], None, "greedy")
    eng.close()
