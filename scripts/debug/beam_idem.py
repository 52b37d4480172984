"""Is batched beam decode run-to-run deterministic at 1024 images? Prints mismatching rows of two identical calls."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import O, engine_for
from test_decode_gpu import FakeDict
from vae_captioning_b200.decode import Decoder
cfg = O.Config(); cfg.gen_max_len = 30
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
params = O.init_params(cfg, seed=2, dtype=torch.float32, scale=scale)
params["decoder/rnn_logits/bias"][2] += 1.0
Bd, beam = 1024, 5
eng = engine_for(cfg, params, (Bd * beam + 4) // 5, 4)
dec = Decoder(eng, cfg, FakeDict(cfg.vocab_size))
g = np.random.Generator(np.random.PCG64(8))
feats = np.maximum(0, g.standard_normal((Bd, 4096))).astype(np.float32)
eps = torch.tensor(g.standard_normal((Bd, cfg.gen_z_samples, cfg.latent_size)).astype(np.float32)).cuda()
runs = [dec.beam_tokens(feats, None, beam_size=beam, rng={"seed": 0, "eps": eps}) for _ in range(3)]
t0, l0, s0, n0 = runs[0]
for r in (1, 2):
    t, l, s, n = runs[r]
    bad = [i for i in range(Bd) if not (np.array_equal(t[i], t0[i]) and np.array_equal(l[i], l0[i]))]
    print("run", r, "rows differing in any beam:", len(bad), "best-beam differing:", sum(not np.array_equal(t[i, 0], t0[i, 0]) for i in range(Bd)),
          "max |score diff|", float(np.abs(s - s0).max()), "nb equal", np.array_equal(n, n0))
    for i in bad[:3]:
        print(" row", i, "lens", l0[i], l[i], "scores", s0[i], s[i])
        print("   t0", t0[i, 0, :8], "t", t[i, 0, :8])
print("len hist", np.bincount(l0[:, 0])[:32])
gt = [dec.greedy_tokens(feats, None, "greedy", {"seed": 0, "eps": eps}) for _ in range(2)]
print("greedy identical:", np.array_equal(gt[0][0], gt[1][0]), "len hist", np.bincount(gt[0][1])[:32])
