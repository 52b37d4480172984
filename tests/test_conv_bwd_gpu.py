"""3x3 SAME convolution backward kernels (filter gradient with the pixel-contraction TMA gather, input gradient with
the tap-reversed weight shadow, ReLU / 2x2 max-pool derivative) against torch.nn.functional.conv2d autograd on the
SAME bf16-rounded inputs, one layer at a time -- no ReLU-decision noise, so the comparison is tight.

Tolerance (stated): fp32 accumulation of bf16 products in a different order than torch: filter gradient within 2e-3
of its max-abs; input gradient is stored in bf16: within 1e-2 of its max-abs (one bf16 ulp = 0.8 %).
"""
import ctypes

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from vae_captioning_b200 import lib as L

pytestmark = pytest.mark.gpu

# (B, hw, cin, cout): every VGG16 geometry class (224/112: 16x4 patches + halo dgrad at cout = 64, 56: 8x8, 28: 4x4x4
# with an image-count tail, 14: 2x2x16 with a tail), Cin = 64 (two taps per A tile, odd tap count) and Cin >= 128
CASES = [(1, 32, 64, 64), (2, 16, 64, 128), (2, 16, 128, 128), (3, 56, 128, 256), (3, 28, 256, 512), (5, 14, 512, 512),
         (1, 224, 64, 64), (2, 112, 64, 128), (17, 14, 512, 512), (2, 8, 256, 256),
         # streamed-filter halo kernel (contraction over 128 / 256 channels into 64 / 128): full-size dgrad2_2, ragged rows
         (2, 112, 128, 128), (1, 24, 128, 256), (3, 40, 64, 256)]


@pytest.mark.parametrize("B,hw,cin,cout", CASES)
def test_conv3x3_backward(B, hw, cin, cout):
    lib = L.load()
    g = torch.Generator().manual_seed(hw * 1000 + cin + B)
    x = torch.randn(B, hw, hw, cin, generator=g).to(torch.bfloat16)
    dy = (torch.randn(B, hw, hw, cout, generator=g) * (torch.rand(B, hw, hw, cout, generator=g) < 0.5)).to(torch.bfloat16)
    w = (torch.randn(3, 3, cin, cout, generator=g) / np.sqrt(9 * cin)).float()
    xd, dyd, wd = x.cuda(), dy.cuda(), w.cuda()
    dw = torch.full((9 * cin, cout), 7.0, device="cuda")  # the entry zeroes it
    dx = torch.zeros(B, hw, hw, cin, dtype=torch.bfloat16, device="cuda")
    L.check(lib.vc_conv3x3_bwd(L.ptr(xd), L.ptr(dyd), L.ptr(wd), L.ptr(dw), L.ptr(dx), B, hw, cin, cout, L.stream_ptr()))
    torch.cuda.synchronize()
    # reference: autograd through conv2d on the same rounded values (fp32 on the GPU, highest precision)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    xr = xd.float().permute(0, 3, 1, 2).requires_grad_(True)
    wr = wd.to(torch.bfloat16).float().permute(3, 2, 0, 1).requires_grad_(True)  # HWIO -> OIHW
    y = F.conv2d(xr, wr, padding=1)
    y.backward(dyd.float().permute(0, 3, 1, 2))
    dw_ref = wr.grad.permute(2, 3, 1, 0).reshape(9 * cin, cout)  # OIHW -> HWIO rows (tap, cin)
    dx_ref = xr.grad.permute(0, 2, 3, 1)
    e_w = (dw - dw_ref).abs().max().item() / dw_ref.abs().max().item()
    e_x = (dx.float() - dx_ref).abs().max().item() / dx_ref.abs().max().item()
    assert e_w <= 2e-3, ("dW", e_w)
    assert e_x <= 1e-2, ("dX", e_x)


@pytest.mark.parametrize("B,hw,C,pooled", [(2, 8, 64, 1), (3, 14, 512, 1), (2, 16, 128, 0), (1, 224, 64, 1)])
def test_relu_pool_backward_bit_exact(B, hw, C, pooled):
    """dY = dA routed to the first maximum of each 2x2 window (torch / TF MaxPoolGrad order) where it is positive;
    bit-exact, including ties (bf16 activations tie often) and all-zero windows."""
    lib = L.load()
    g = torch.Generator().manual_seed(hw + C)
    out = torch.relu(torch.randn(B, hw, hw, C, generator=g) * 4).round().clamp(max=6).to(torch.bfloat16)  # many ties
    ho = hw // 2 if pooled else hw
    dA = torch.randn(B, ho, ho, C, generator=g).to(torch.bfloat16)
    od, dAd = out.cuda(), dA.cuda()
    dY = torch.full((B, hw, hw, C), 3.0, dtype=torch.bfloat16, device="cuda")
    db = torch.zeros(C, device="cuda")
    L.check(lib.vc_relu_pool_bwd(L.ptr(dAd), L.ptr(od), L.ptr(dY), L.ptr(db), B, hw, C, pooled, L.stream_ptr()))
    torch.cuda.synchronize()
    o32 = out.float().permute(0, 3, 1, 2).requires_grad_(True)  # CPU autograd: first-max routing
    r = torch.relu(o32)  # out is post-ReLU: relu'(out) = out > 0
    y = F.max_pool2d(r, 2, 2) if pooled else r
    y.backward(dA.float().permute(0, 3, 1, 2))
    ref = o32.grad.permute(0, 2, 3, 1).to(torch.bfloat16)
    assert torch.equal(dY.cpu(), ref)
    # fused bias gradient: per-channel sum of dY (fp32 accumulation of the bf16 values)
    want = ref.float().sum(dim=(0, 1, 2))
    assert (db.cpu() - want).abs().max().item() <= 1e-3 * max(1.0, want.abs().max().item())


@pytest.mark.parametrize("B,hw,cin,cout", [(2, 16, 128, 128), (3, 56, 128, 256), (3, 28, 256, 512), (5, 14, 512, 512),
                                           (2, 112, 128, 128), (1, 24, 128, 256), (3, 40, 64, 256), (1, 32, 64, 64)])
def test_conv3x3_dgrad_with_fused_relu_derivative(B, hw, cin, cout):
    """The fine-tune pass takes the ReLU derivative and the bias gradient of an un-pooled layer in the epilogue of the
    input-gradient GEMM above it (EpiTmaRelu). Against the unfused entry on the same inputs the masked gradient must be
    BIT-IDENTICAL (same accumulator, same bf16 rounding, then a mask) and the bias gradient must equal the column sums of
    that bf16 tensor to fp32 summation-order noise. Covers the generic implicit GEMM, both halo kernels, CTA pairs, odd
    tile counts and ragged tiles."""
    lib = L.load()
    g = torch.Generator().manual_seed(hw * 77 + cin + B)
    x = torch.zeros(B, hw, hw, cin, dtype=torch.bfloat16)
    dy = torch.randn(B, hw, hw, cout, generator=g).to(torch.bfloat16)
    w = (torch.randn(3, 3, cin, cout, generator=g) / np.sqrt(9 * cin)).float()
    act = torch.relu(torch.randn(B, hw, hw, cin, generator=g)).to(torch.bfloat16)  # half of it exactly zero
    xd, dyd, wd, actd = x.cuda(), dy.cuda(), w.cuda(), act.cuda()
    dw = torch.zeros(9 * cin, cout, device="cuda")
    dx_plain = torch.zeros(B, hw, hw, cin, dtype=torch.bfloat16, device="cuda")
    L.check(lib.vc_conv3x3_bwd(L.ptr(xd), L.ptr(dyd), L.ptr(wd), L.ptr(dw), L.ptr(dx_plain), B, hw, cin, cout, L.stream_ptr()))
    dx = torch.full((B, hw, hw, cin), 3.0, dtype=torch.bfloat16, device="cuda")
    db = torch.zeros(cin, device="cuda")
    L.check(lib.vc_conv3x3_dgrad_relu(L.ptr(dyd), L.ptr(wd), L.ptr(actd), L.ptr(dx), L.ptr(db), B, hw, cin, cout, L.stream_ptr()))
    torch.cuda.synchronize()
    want = torch.where(actd > 0, dx_plain, torch.zeros_like(dx_plain))
    assert torch.equal(dx.view(torch.int16), want.view(torch.int16))
    ref_db = want.float().sum(dim=(0, 1, 2))
    assert (db - ref_db).abs().max().item() <= 1e-4 * max(1.0, ref_db.abs().max().item())
