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
