#!/usr/bin/env python
"""Workload for scripts/gpu_sanitize.sh: one small ELBO train step per prior (persistent LSTM kernels, tcgen05 GEMMs,
CE, Adam), one fine-tune step would be too slow under the tool and is left out; plus one greedy and one beam decode.
Each result is still checked against the oracle so a tool-induced slowdown cannot hide a wrong answer."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import O, SMALL, TINY, engine_for, feed_of, make_case, rng_for  # noqa: E402


def main():
    for sizes, B, T, kw in ((TINY, 3, 5, {}), (SMALL, 4, 6, dict(prior="GMM", use_c_v=True)), (SMALL, 30, 7, dict(prior="AG", use_c_v=True))):
        cfg, params, batch = make_case(sizes, B, T, seed=3, ragged=True, **kw)
        eng = engine_for(cfg, params, B, T)
        out = eng.train_step(anneal=0, rng=rng_for(batch), **feed_of(batch))
        ref = O.train_step({k: v.clone() for k, v in params.items()}, {"t": 0, "m": {}, "v": {}}, cfg, batch)
        assert abs(out["rec_loss"] - ref["rec_loss"]) <= 5e-3 * abs(ref["rec_loss"]), (kw, out, ref["rec_loss"])
        eng.close()
        print("train step ok", kw, out["rec_loss"])
    from test_decode_gpu import device_decoder, make_decode_case
    cfg, params, feats, c_v, eps = make_decode_case(SMALL, 6, seed=2)
    eng, dec = device_decoder(cfg, params, 6, 3)
    rng = {"eps": torch.tensor(eps).cuda()}
    toks, lens = dec.greedy_tokens(feats, c_v, "greedy", rng)
    bt, bl, bs, nb = dec.beam_tokens(feats, c_v, beam_size=3, rng=rng)
    assert lens.min() >= 1 and int(nb.min()) >= 1
    eng.close()
    print("decode ok", lens.tolist())


if __name__ == "__main__":
    main()
