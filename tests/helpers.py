"""Shared test plumbing: build an oracle model + batch, mirror it into an Engine, compare."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import cvae_oracle as O  # noqa: E402  (the oracle is test infrastructure)

TINY = dict(vocab_size=37, embed_size=64, encoder_hidden=64, decoder_hidden=64, latent_size=6, gen_z_samples=4,
            num_captions=3, cnn_feature_size=64)
SMALL = dict(vocab_size=1003, embed_size=128, encoder_hidden=128, decoder_hidden=192, latent_size=20, gen_z_samples=10,
             num_captions=5, cnn_feature_size=256)


def make_case(sizes, B, T, seed=0, ragged=False, dtype=torch.float64, **cfg_kw):
    kw = dict(sizes)
    kw.update(cfg_kw)
    cfg = O.Config(**kw)
    params = O.init_params(cfg, seed=seed + 1, dtype=dtype)
    # non-zero biases so bias paths are exercised (TF initialises them to zero; values are arbitrary test data)
    g = np.random.Generator(np.random.PCG64(seed + 7))
    for n in params:
        if params[n].dim() == 1:
            params[n] = torch.tensor(g.uniform(-0.1, 0.1, size=tuple(params[n].shape)).astype(np.float32)).to(dtype)
    batch = O.synthetic_batch(cfg, B, T, seed=seed, dtype=dtype, ragged=ragged)
    return cfg, params, batch


def engine_for(cfg, params, B, T, **kw):
    from vae_captioning_b200.engine import Engine
    eng = Engine(cfg, vocab_size=cfg.vocab_size, max_batch=B, max_len=T, **kw)
    eng.load_state({n: v.to(torch.float32).numpy() for n, v in params.items()})
    if cfg.prior == "AG" and not cfg.no_encoder:
        eng.set_cluster_means(O.init_clusters(cfg.num_clusters, cfg.latent_size).numpy())  # the oracle's default (seed 2)
    return eng


def rng_for(batch, device="cuda"):
    r = {"seed": 0}
    for src, dst in (("eps", "eps"), ("emb_keep_mask", "emb_keep"), ("out_keep_mask", "out_keep")):
        if src in batch:
            r[dst] = batch[src].to(torch.float32).contiguous().to(device)
    if "gmm_cluster" in batch:
        r["gmm_cluster"] = batch["gmm_cluster"].to(torch.int32).contiguous().to(device)
    return r


def feed_of(batch):
    cv = batch["c_v"].to(torch.float32).numpy() if "c_v" in batch else None
    return dict(image_f_inputs=batch["feats"].to(torch.float32).numpy(), ann_inputs_enc=batch["cap_lbl"].numpy(),
                ann_inputs_dec=batch["cap_in"].numpy(), ann_lengths=batch["lengths"].numpy().astype(np.float64), c_i=cv)


def rel_err(got, ref):
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    return float(np.max(np.abs(got - ref)) / max(1e-30, float(np.max(np.abs(ref)))))
