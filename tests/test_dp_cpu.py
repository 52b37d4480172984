"""Data-parallel host logic on CPU: world_size-2 gloo processes run DataParallelStep (shard -> sum all-reduce ->
scaled apply) over a CPU stand-in backend whose gradients come from the oracle, and must land on the parity contract
of SURVEY 8e: the result of ONE process that averages the gradients of W reference replicas, each on B/W images,
with the Q4 global norm taken over the concatenated embedding slices."""
import math
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import O, TINY, make_case
from vae_captioning_b200 import dp


def test_shard_range_partitions():
    for B in (1, 2, 7, 32, 256):
        for W in (1, 2, 3, 8):
            if W > B:
                with pytest.raises(ValueError):
                    dp.shard_feed({"image_f_inputs": np.zeros((B, 4))}, W - 1, W, 5)
                continue
            spans = [dp.shard_range(B, r, W) for r in range(W)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        dp.shard_range(8, 2, 2)


def test_shard_feed_keeps_captions_with_their_image():
    B, C, T = 6, 3, 4
    feed = {"image_f_inputs": np.arange(B)[:, None] * np.ones((1, 5)), "ann_inputs_enc": np.repeat(np.arange(B), C)[:, None] * np.ones((1, T), int),
            "ann_inputs_dec": np.zeros((B * C, T), int), "ann_lengths": np.repeat(np.arange(B), C), "c_i": None, "anneal": 7}
    for r in range(3):
        s = dp.shard_feed(feed, r, 3, C)
        imgs = s["image_f_inputs"][:, 0]
        assert list(s["ann_lengths"]) == list(np.repeat(imgs, C))
        assert s["ann_inputs_enc"].shape == (2 * C, T) and s["c_i"] is None and s["anneal"] == 7
    bad = dict(feed)
    bad["ann_lengths"] = np.zeros(5)
    with pytest.raises(ValueError):
        dp.shard_feed(bad, 0, 3, C)


class OracleBackend(object):
    """CPU stand-in for the Engine: flat gradient buffer = [dense grads | embedding grads | tail(2 slice norms)]."""

    def __init__(self, cfg, params):
        self.cfg, self.params = cfg, {k: v.clone() for k, v in params.items()}
        self.names = O.trainable_names(cfg, params)
        self.sizes = [params[n].numel() for n in self.names]
        self.flat = torch.zeros(sum(self.sizes) + 2, dtype=torch.float64)
        self.opt = {"t": 0, "m": {}, "v": {}}

    def forward_backward(self, batch):
        res, grads, _ = O.compute_grads(self.params, self.cfg, batch)
        off = 0
        for n, k in zip(self.names, self.sizes):
            self.flat[off:off + k] = grads[n].reshape(-1)
            off += k
        self.flat[off] = float((res["x_enc"].grad ** 2).sum())
        self.flat[off + 1] = float((res["x_dec"].grad ** 2).sum())

    def grad_tensor(self):
        return self.flat

    def apply(self, scale, fetch=True):
        off, sq, g = 0, 0.0, {}
        for n, k in zip(self.names, self.sizes):
            g[n] = self.flat[off:off + k].reshape(self.params[n].shape) * scale
            if not n.endswith("embeddings"):
                sq += float((g[n] ** 2).sum())
            off += k
        sq += float(self.flat[off] + self.flat[off + 1]) * scale * scale  # Q4: concatenated slices of all towers
        norm = math.sqrt(sq)
        clip = self.cfg.lstm_clip_by_norm
        self.opt["t"] += 1
        for n in self.names:
            m = self.opt["m"].get(n, torch.zeros_like(self.params[n]))
            v = self.opt["v"].get(n, torch.zeros_like(self.params[n]))
            self.params[n], self.opt["m"][n], self.opt["v"][n] = O.adam_update(
                self.params[n], g[n] * (clip / max(norm, clip)), m, v, self.cfg.learning_rate, self.opt["t"])
        return {"global_norm": norm}


def _case():
    cfg, params, batch = make_case(TINY, 4, 5, seed=13, ragged=True)
    return cfg, params, batch


def _split_batch(cfg, batch, r, W):
    B = batch["feats"].shape[0]
    lo, hi = dp.shard_range(B, r, W)
    C, S = cfg.num_captions, cfg.gen_z_samples
    out = {}
    for k, v in batch.items():
        if k == "feats":
            out[k] = v[lo:hi]
        elif k == "eps":
            out[k] = v[:, lo * C:hi * C]
        elif torch.is_tensor(v) and v.dim() >= 1 and v.shape[0] == B * C:
            out[k] = v[lo * C:hi * C]
        else:
            out[k] = v
    return out


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    cfg, params, batch = _case()
    backend = OracleBackend(cfg, params)
    step = dp.DataParallelStep(backend)
    out = step(_split_batch(cfg, batch, rank, world))
    ret[rank] = ({k: v.clone() for k, v in backend.params.items()}, out["global_norm"])
    dist.barrier()
    dist.destroy_process_group()


def test_dp_world2_gloo_matches_mean_of_replicas():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    # single-process statement of the contract
    cfg, params, batch = _case()
    ref = OracleBackend(cfg, params)
    acc = torch.zeros_like(ref.flat)
    for r in range(world):
        ref.forward_backward(_split_batch(cfg, batch, r, world))
        acc += ref.flat
    ref.flat.copy_(acc)
    out = ref.apply(1.0 / world)
    for r in range(world):
        got, norm = ret[r]
        assert abs(norm - out["global_norm"]) <= 1e-12 * max(1.0, out["global_norm"])
        for n in ref.params:
            assert torch.allclose(got[n], ref.params[n], rtol=0, atol=1e-12), n
    # and both ranks hold identical parameters (replicated state stays in sync)
    for n in ref.params:
        assert torch.equal(ret[0][0][n], ret[1][0][n])
