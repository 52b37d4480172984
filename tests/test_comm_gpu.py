"""The in-library NCCL path (include/vaecap.h: vc_comm_unique_id / vc_comm_init / vc_comm_set_mode / vc_comm_stats) on ONE
GPU: a one-rank communicator still sends every gradient bucket through ncclAllReduce on the communication stream, so the
event / stream plumbing, the bucket ranges and the bf16 transport are exercised by the single-GPU suite; the multi-rank
numbers are checked on hardware by scripts/dp_two_ranks.py (2 ranks) and bench.py's dp_check (2 / 8 ranks).

Contract: with world = 1 the data-parallel step IS the plain step (sum over one tower, scale 1/1)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist

from helpers import SMALL, engine_for, feed_of, make_case, rng_for

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def one_rank_group():
    if not dist.is_initialized():
        dist.init_process_group("gloo", init_method="tcp://127.0.0.1:29577", rank=0, world_size=1)
    yield
    if dist.is_initialized():
        dist.destroy_process_group()


def _step(cfg, params, batch, B, T, comm, mode=1):
    eng = engine_for(cfg, params, B, T)
    if comm:
        eng.attach_comm(0, 1)
        eng.comm_set_mode(mode)
    out = eng.train_step(anneal=0, rng=rng_for(batch), **feed_of(batch))
    torch.cuda.synchronize()
    stats = eng.comm_stats() if comm else None
    state = {n: eng.get_variable(n) for n in ("decoder/rnn_logits/kernel", "encoder/enc_embeddings", "imf_emb/kernel",
                                              "decoder/net/multi_rnn_cell/cell_0/lstm_cell/kernel")}
    eng.close()
    return out, state, stats


@pytest.mark.parametrize("kw", [{}, dict(prior="GMM", use_c_v=True)])
@pytest.mark.parametrize("mode", [1, 2])
def test_one_rank_communicator_is_the_plain_step(one_rank_group, kw, mode):
    B, T = 4, 6
    cfg, params, batch = make_case(SMALL, B, T, seed=19, ragged=True, **kw)
    plain, s0, _ = _step(cfg, params, batch, B, T, comm=False)
    dp, s1, stats = _step(cfg, params, batch, B, T, comm=True, mode=mode)
    assert (stats["buckets"] == 1 if mode == 2 else stats["buckets"] >= 4) and stats["bytes"] > 0
    assert abs(dp["rec_loss"] - plain["rec_loss"]) <= 1e-5 * abs(plain["rec_loss"])
    assert abs(dp["global_norm"] - plain["global_norm"]) <= 1e-4 * plain["global_norm"]
    for n in s0:  # split-K atomics reorder fp32 sums between two runs; Adam's first step is ~lr * sign(g)
        d = np.abs(s1[n].astype(np.float64) - s0[n])
        assert np.mean(d > 1e-6) <= 2e-3, n


def test_bf16_transport_flag(one_rank_group):
    """VC_GRAD_BF16=1: gradients cross the wire as bf16 (half the bytes); the update stays within bf16 rounding."""
    B, T = 4, 6
    cfg, params, batch = make_case(SMALL, B, T, seed=23, ragged=True)
    _, _, fp32 = _step(cfg, params, batch, B, T, comm=True)
    os.environ["VC_GRAD_BF16"] = "1"
    try:
        out, _, bf16 = _step(cfg, params, batch, B, T, comm=True)
    finally:
        del os.environ["VC_GRAD_BF16"]
    assert bf16["bytes"] < 0.51 * fp32["bytes"] + 1024
    assert np.isfinite(out["global_norm"]) and out["global_norm"] > 0


def test_comm_errors(one_rank_group):
    cfg, params, batch = make_case(SMALL, 2, 5, seed=3)
    eng = engine_for(cfg, params, 2, 5)
    from vae_captioning_b200 import lib as L
    with pytest.raises(L.VaecapError):
        eng.comm_stats()  # no communicator yet
    eng.attach_comm(0, 1)
    with pytest.raises(L.VaecapError):
        eng.attach_comm(0, 1)  # a handle has one communicator
    with pytest.raises(ValueError):
        eng.comm_set_mode(7)
    eng.close()
