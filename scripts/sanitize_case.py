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


def kernel_level():
    """CTA-pair forms (cluster of 2, cta_group::2, remote mbarrier arrivals, multicast commits) of the GEMM mainloop and of
    both halo convolution kernels, the wide filter-gradient kernel and the fused ReLU-derivative epilogue, on small
    shapes -- the train steps above are too small to reach most of them."""
    import ctypes
    from vae_captioning_b200 import lib as L
    lib = L.load()
    lib.vc_test_pair_mode(1)
    g = torch.Generator().manual_seed(11)
    for (M, N, K, a_mn, b_mn, bn) in ((512, 256, 320, 0, 0, 256), (384, 256, 256, 1, 1, 128), (300, 200, 192, 0, 0, 64)):
        A = torch.randn(M, K, generator=g).to(torch.bfloat16)
        B = torch.randn(N, K, generator=g).to(torch.bfloat16)
        Ad = (A.t().contiguous() if a_mn else A).cuda()
        Bd = (B.t().contiguous() if b_mn else B).cuda()
        out = torch.zeros(M, N, device="cuda")
        L.check(lib.vc_gemm_bf16(L.ptr(Ad), a_mn, ctypes.c_longlong(M if a_mn else K), L.ptr(Bd), b_mn,
                                 ctypes.c_longlong(N if b_mn else K), L.ptr(out), ctypes.c_longlong(N), None, M, N, K, bn, 1, 0, 0,
                                 0, L.stream_ptr()))
        torch.cuda.synchronize()
        ref = A.float().cuda() @ B.float().cuda().t()
        assert (out - ref).abs().max().item() <= 2e-3 * ref.abs().max().item(), (M, N, K)
    for (Bn, hw, cin, cout) in ((2, 16, 64, 128), (2, 16, 128, 128), (1, 56, 128, 256), (3, 8, 256, 256)):
        x = torch.randn(Bn, hw, hw, cin, generator=g).to(torch.bfloat16).cuda()
        dy = torch.randn(Bn, hw, hw, cout, generator=g).to(torch.bfloat16).cuda()
        w = (torch.randn(3, 3, cin, cout, generator=g) / np.sqrt(9 * cin)).float().cuda()
        act = torch.relu(torch.randn(Bn, hw, hw, cin, generator=g)).to(torch.bfloat16).cuda()
        dw = torch.zeros(9 * cin, cout, device="cuda")
        dx0 = torch.zeros(Bn, hw, hw, cin, dtype=torch.bfloat16, device="cuda")
        dx1 = torch.zeros_like(dx0)
        db = torch.zeros(cin, device="cuda")
        L.check(lib.vc_conv3x3_bwd(L.ptr(x), L.ptr(dy), L.ptr(w), L.ptr(dw), L.ptr(dx0), Bn, hw, cin, cout, L.stream_ptr()))
        L.check(lib.vc_conv3x3_dgrad_relu(L.ptr(dy), L.ptr(w), L.ptr(act), L.ptr(dx1), L.ptr(db), Bn, hw, cin, cout, L.stream_ptr()))
        torch.cuda.synchronize()
        assert torch.equal(dx1.view(torch.int16), torch.where(act > 0, dx0, torch.zeros_like(dx0)).view(torch.int16)), (hw, cin, cout)
    lib.vc_test_pair_mode(-1)
    print("kernel level ok")


def main():
    for sizes, B, T, kw in ((TINY, 3, 5, {}), (SMALL, 4, 6, dict(prior="GMM", use_c_v=True)), (SMALL, 30, 7, dict(prior="AG", use_c_v=True))):
        cfg, params, batch = make_case(sizes, B, T, seed=3, ragged=True, **kw)
        eng = engine_for(cfg, params, B, T)
        out = eng.train_step(anneal=0, rng=rng_for(batch), **feed_of(batch))
        ref = O.train_step({k: v.clone() for k, v in params.items()}, {"t": 0, "m": {}, "v": {}}, cfg, batch)
        assert abs(out["rec_loss"] - ref["rec_loss"]) <= 5e-3 * abs(ref["rec_loss"]), (kw, out, ref["rec_loss"])
        eng.close()
        print("train step ok", kw, out["rec_loss"])
    kernel_level()
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
