"""On-device VGG16 forward (implicit-GEMM tcgen05 convolutions, fused pools, fc1/fc2) vs the CPU oracle
(torch conv2d restatement of utils/image_embeddings.py:26-238) on seeded random weights and images.

Tolerance (stated): bf16 operands / fp32 accumulate. Against the oracle in bf16-emulation mode each layer's
activation must agree within 2e-2 of that layer's max-abs (one bf16 ulp is 0.8 %, and 13 layers of rounding
decisions compound); fc2 within 3e-2 of its max-abs; against the exact fp64 oracle fc2 within 5e-2.
"""
import numpy as np
import pytest
import torch

from helpers import O, TINY, engine_for, rel_err

pytestmark = pytest.mark.gpu


def make_vgg_case(B, seed=0):
    cfg = O.Config(**TINY)
    params = O.init_params(cfg, seed=seed + 1, with_cnn=True, dtype=torch.float32)
    g = np.random.Generator(np.random.PCG64(seed + 5))
    for n in params:
        if n.startswith("cnn/") and params[n].dim() == 1:
            params[n] = torch.tensor(g.uniform(-0.05, 0.05, size=tuple(params[n].shape)).astype(np.float32))
    images = torch.tensor(g.integers(0, 256, size=(B, 224, 224, 3)).astype(np.float32))
    return cfg, params, images


@pytest.mark.parametrize("B", [2, 3])
def test_vgg_forward_layers_and_fc2(B):
    cfg, params, images = make_vgg_case(B)
    eng = engine_for(cfg, params, B, 5, with_cnn=True)
    taps = {}
    with torch.no_grad():
        ref = O.vgg16_fc2(params, images, emulate=True, taps=taps)
    # 1) un-fused path: every conv output materialised
    eng.vgg_keep_activations(True)
    fc2_a = eng.vgg_forward(images.numpy())
    worst = {}
    for name, t in taps.items():
        worst[name] = rel_err(eng.vgg_activation(name, B), t.numpy())
    bad = {k: v for k, v in worst.items() if v > 2e-2}
    assert not bad, (bad, worst)
    assert rel_err(fc2_a, ref.numpy()) <= 3e-2
    # 2) production path: 2x2 max-pools fused into the conv epilogues. pool5 is bit-identical to the un-fused
    # result (max commutes with the monotone bias+ReLU+rounding); fc1/fc2 accumulate split-K partial sums with
    # fp32 atomics, so fc2 agrees to fp32 summation-order noise.
    p5_unfused = eng.vgg_activation("pool5", B)
    eng.vgg_keep_activations(False)
    fc2_b = eng.vgg_forward(images.numpy())
    np.testing.assert_array_equal(p5_unfused, eng.vgg_activation("pool5", B))
    np.testing.assert_allclose(fc2_a, fc2_b, rtol=1e-4, atol=1e-4 * np.abs(fc2_a).max())
    p5 = eng.vgg_activation("pool5", B)
    ref_p5 = torch.nn.functional.max_pool2d(taps["conv5_3"].permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1)
    assert rel_err(p5, ref_p5.numpy()) <= 2e-2
    # 2b) uint8 pixels (the HDF5 store's dtype): same values, a quarter of the bytes -> identical up to atomics order
    fc2_u8 = eng.vgg_forward(images.numpy().astype(np.uint8))
    np.testing.assert_allclose(fc2_b, fc2_u8, rtol=1e-4, atol=1e-4 * np.abs(fc2_b).max())
    # 3) device-resident entry point
    fc2_c = eng.vgg_forward_device(images.cuda()).cpu().numpy()
    np.testing.assert_allclose(fc2_a, fc2_c, rtol=1e-4, atol=1e-4 * np.abs(fc2_a).max())
    eng.close()


def test_vgg_forward_vs_exact_oracle():
    B = 2
    cfg, params, images = make_vgg_case(B, seed=3)
    eng = engine_for(cfg, params, B, 5, with_cnn=True)
    got = eng.vgg_forward(images.numpy())
    with torch.no_grad():
        ref = O.vgg16_fc2({k: v.double() for k, v in params.items()}, images.double())
    assert rel_err(got, ref.numpy()) <= 5e-2
    assert np.all(got >= 0)
    eng.close()


def test_vgg_requires_cnn_handle():
    cfg, params, images = make_vgg_case(1)
    eng = engine_for(cfg, {k: v for k, v in params.items() if not k.startswith("cnn/")}, 1, 5)
    with pytest.raises(Exception):
        eng.vgg_forward(images.numpy())
    eng.close()


def test_vgg_conv_layers_cta_pairs_match_single_ctas_bitwise():
    """Every convolution kernel has a CTA-pair form (cluster of 2, tcgen05 cta_group::2: the generic implicit GEMM and both
    halo kernels). Pairing changes which SM holds which operand half, not the order of the fp32 accumulation, so each
    layer's activation must be bit-identical to the single-CTA form -- B = 3 gives odd m-tile counts (a surplus tile in
    the last pair)."""
    from vae_captioning_b200 import lib as L
    lib = L.load()
    B = 3
    cfg, params, images = make_vgg_case(B, seed=7)
    names = ["conv1_2", "conv2_1", "conv2_2", "conv3_1", "conv3_3", "conv4_2", "conv5_3", "pool5"]
    acts = {}
    try:
        for mode in (0, 1):
            lib.vc_test_pair_mode(mode)
            eng = engine_for(cfg, params, B, 5, with_cnn=True)
            eng.vgg_keep_activations(True)
            eng.vgg_forward(images.numpy())
            acts[mode] = {n: eng.vgg_activation(n, B).copy() for n in names}
            eng.vgg_keep_activations(False)
            eng.vgg_forward(images.numpy())
            acts[mode]["pool5_fused"] = eng.vgg_activation("pool5", B).copy()
            eng.close()
    finally:
        lib.vc_test_pair_mode(-1)
    for n in acts[0]:
        np.testing.assert_array_equal(acts[0][n], acts[1][n], err_msg=n)
