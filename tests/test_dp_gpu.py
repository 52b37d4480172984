"""Data-parallel split of the train step on the device: vc_forward_backward_dev fills the flat gradient
buffer, the caller sums the buffers of all ranks (here: two shards run back to back on one GPU and summed through the
vc_grad_buffer alias, exactly what NCCL's all-reduce does in bench.py), vc_apply_gradients(1/W) clips and updates.
Contract: equals ONE reference process whose gradient is the mean of W replicas at B/W images each, with the Q4 norm
over the concatenated embedding slices. Tolerances as in test_train_step_gpu.py (bf16 operands)."""
import math

import numpy as np
import pytest
import torch

from helpers import O, SMALL, engine_for, make_case, rng_for
from vae_captioning_b200 import dp
from vae_captioning_b200 import lib as L

pytestmark = pytest.mark.gpu


def _shard(cfg, batch, r, W):
    B = batch["feats"].shape[0]
    lo, hi = dp.shard_range(B, r, W)
    C = cfg.num_captions
    out = {}
    for k, v in batch.items():
        if k == "feats":
            out[k] = v[lo:hi]
        elif k == "eps":
            out[k] = v[:, lo * C:hi * C].contiguous()
        elif torch.is_tensor(v) and v.dim() >= 1 and v.shape[0] == B * C:
            out[k] = v[lo * C:hi * C]
        else:
            out[k] = v
    return out


@pytest.mark.parametrize("kw", [{}, dict(prior="AG", use_c_v=True)])
def test_two_shards_match_mean_of_replicas(kw):
    W, B, T = 2, 4, 6
    cfg, params, batch = make_case(SMALL, B, T, seed=17, ragged=True, **kw)
    eng = engine_for(cfg, params, B // W, T)
    ptr, count = eng.grad_buffer()
    grad = L.alias_tensor(ptr, count, torch.float32, 0)
    total = torch.zeros_like(grad)
    dev = lambda a, dt: torch.tensor(np.ascontiguousarray(a)).to(dt).cuda()
    ref_sum, slice_sq, names = {}, 0.0, O.trainable_names(cfg, params)
    for r in range(W):
        sb = _shard(cfg, batch, r, W)
        cv = dev(sb["c_v"].float().numpy(), torch.float32) if "c_v" in sb else None
        eng.forward_backward_device(dev(sb["feats"].float().numpy(), torch.float32), dev(sb["cap_lbl"].numpy(), torch.int32),
                                    dev(sb["cap_in"].numpy(), torch.int32), dev(sb["lengths"].numpy(), torch.int32), 0,
                                    c_i=cv, rng=rng_for(sb))
        torch.cuda.synchronize()
        total += grad
        res, grads, _ = O.compute_grads(params, cfg, sb)
        for n in names:
            if grads[n] is not None:
                ref_sum[n] = ref_sum.get(n, 0) + grads[n]
        slice_sq += float((res["x_enc"].grad ** 2).sum()) + float((res["x_dec"].grad ** 2).sum())
    grad.copy_(total)  # the all-reduce(sum)
    out = eng.apply_gradients(1.0 / W)
    # reference: mean gradient, Q4 norm over concatenated slices, clip, TF Adam
    sq = slice_sq / (W * W)
    for n, g in ref_sum.items():
        if not n.endswith("embeddings"):
            sq += float(((g / W) ** 2).sum())
    norm = math.sqrt(sq)
    assert abs(out["global_norm"] - norm) <= 3e-2 * norm
    clip = cfg.lstm_clip_by_norm
    for name in ("decoder/rnn_logits/kernel", "decoder/net/multi_rnn_cell/cell_0/lstm_cell/kernel", "imf_emb/kernel"):
        g = ref_sum[name] / W * (clip / max(norm, clip))
        want, _, _ = O.adam_update(params[name], g, torch.zeros_like(g), torch.zeros_like(g), cfg.learning_rate, 1)
        got = eng.get_variable(name).astype(np.float64) - params[name].numpy()
        want = want.numpy() - params[name].numpy()
        assert np.linalg.norm(got - want) <= 0.15 * np.linalg.norm(want), name
    eng.close()
