"""tcgen05 GEMM mainloop vs a plain fp32 torch reference (all operand major-ness combinations,
ragged shapes, split-K, bias/ReLU, bf16/fp32 outputs)."""
import ctypes

import pytest
import torch

from vae_captioning_b200 import lib as L

pytestmark = pytest.mark.gpu


def run_gemm(M, N, K, a_mn, b_mn, bn, splits=1, relu=0, out_bf16=0, bias=True, seed=0, ldo=None, force_atomic=False):
    lib = L.load()
    g = torch.Generator(device="cpu").manual_seed(seed)
    # pitches padded to 8 elements so rows are 16-byte aligned
    def pad8(x):
        return (x + 7) // 8 * 8
    A_log = torch.randn(M, K, generator=g)
    B_log = torch.randn(N, K, generator=g)
    if a_mn:
        A_buf = torch.zeros(K, pad8(M)); A_buf[:, :M] = A_log.t(); lda = pad8(M)
    else:
        A_buf = torch.zeros(M, pad8(K)); A_buf[:, :K] = A_log; lda = pad8(K)
    if b_mn:
        B_buf = torch.zeros(K, pad8(N)); B_buf[:, :N] = B_log.t(); ldb = pad8(N)
    else:
        B_buf = torch.zeros(N, pad8(K)); B_buf[:, :K] = B_log; ldb = pad8(K)
    A_d = A_buf.to(torch.bfloat16).cuda()
    B_d = B_buf.to(torch.bfloat16).cuda()
    bias_d = torch.randn(N, generator=g).cuda() if bias else None
    ldo = pad8(N) if ldo is None else ldo
    atomic = 1 if (splits > 1 or force_atomic) else 0
    out = torch.zeros(M, ldo, dtype=torch.bfloat16 if out_bf16 else torch.float32, device="cuda")
    st = lib.vc_gemm_bf16(L.ptr(A_d), a_mn, ctypes.c_longlong(lda), L.ptr(B_d), b_mn, ctypes.c_longlong(ldb),
                          L.ptr(out), ctypes.c_longlong(ldo), L.ptr(bias_d), M, N, K, bn, splits, relu, out_bf16,
                          atomic, L.stream_ptr())
    L.check(st)
    torch.cuda.synchronize()
    ref = A_log.to(torch.bfloat16).float().cuda() @ B_log.to(torch.bfloat16).float().cuda().t()
    if bias:
        ref = ref + bias_d
    if relu:
        ref = ref.clamp_min(0)
    got = out[:, :N].float()
    tol = 2e-2 if out_bf16 else 2e-3
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err <= tol * max(1.0, scale), "M=%d N=%d K=%d a_mn=%d b_mn=%d bn=%d err=%g scale=%g" % (
        M, N, K, a_mn, b_mn, bn, err, scale)
    # padding columns must stay untouched
    if ldo > N:
        assert out[:, N:].abs().max().item() == 0


@pytest.mark.parametrize("a_mn", [0, 1])
@pytest.mark.parametrize("b_mn", [0, 1])
@pytest.mark.parametrize("bn", [64, 128, 256])
def test_gemm_major_modes(a_mn, b_mn, bn):
    run_gemm(384, 512, 256, a_mn, b_mn, bn)


@pytest.mark.parametrize("shape", [(6, 37, 8), (130, 300, 520), (1, 16, 64), (257, 11313 // 8, 72), (300, 200, 15000 // 10)])
@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (1, 1), (0, 1), (1, 0)])
def test_gemm_ragged(shape, a_mn, b_mn):
    M, N, K = shape
    run_gemm(M, N, K, a_mn, b_mn, 64)


def test_gemm_small_bn():
    for bn in (16, 32, 48):
        run_gemm(200, 100, 192, 0, 0, bn)


def test_gemm_splitk():
    run_gemm(256, 256, 4096, 0, 0, 128, splits=7)
    run_gemm(256, 256, 4096, 1, 1, 64, splits=4, bias=False)


def test_gemm_epilogue_variants():
    run_gemm(512, 384, 320, 0, 0, 128, relu=1, out_bf16=1)
    run_gemm(512, 384, 320, 0, 0, 128, relu=1, out_bf16=0)


def test_gemm_many_tiles_persistent():
    # more tiles than SMs: exercises the persistent loop, the smem ring wrap and the TMEM double buffer
    run_gemm(4096, 2048, 192, 0, 0, 64)
    run_gemm(2048, 4096, 1024, 0, 0, 256)


@pytest.mark.parametrize("out_bf16", [0, 1])
def test_gemm_unaligned_output_pitch(out_bf16):
    # an output pitch that is not a multiple of 16 bytes (the vocab-projection weight gradient is [H, V] with V = 11313):
    # rows cannot take vector stores, the epilogue transposes the block through shared memory and writes row segments
    run_gemm(300, 203, 320, 0, 0, 64, out_bf16=out_bf16, ldo=203)
    run_gemm(130, 77, 192, 1, 1, 64, out_bf16=out_bf16, ldo=81)
    run_gemm(257, 333, 128, 0, 0, 256, out_bf16=out_bf16, ldo=333)


def test_gemm_splitk_ragged_atomic():
    # split-K partial sums meet in coalesced fp32 atomics: ragged rows (M % 32 != 0), ragged columns, odd pitch
    run_gemm(203, 301, 2048, 0, 0, 64, splits=5, ldo=301)
    run_gemm(70, 45, 4096, 1, 1, 64, splits=8, bias=False, ldo=45)
    run_gemm(512, 1000, 1536, 1, 1, 256, splits=3, ldo=1001)
    run_gemm(97, 130, 256, 0, 0, 128, force_atomic=True)


@pytest.fixture
def pair_forced():
    """Run the mainloop as CTA pairs (cluster of 2, tcgen05 cta_group::2) wherever the tile shape allows it."""
    lib = L.load()
    lib.vc_test_pair_mode(1)
    yield
    lib.vc_test_pair_mode(-1)


@pytest.mark.parametrize("a_mn", [0, 1])
@pytest.mark.parametrize("b_mn", [0, 1])
@pytest.mark.parametrize("bn", [64, 128, 256])
def test_gemm_pair_major_modes(pair_forced, a_mn, b_mn, bn):
    # bn = 64 with an MN-major B is not pairable (a CTA would stage 32 columns): falls back to single CTAs, same result
    run_gemm(384, 512, 256, a_mn, b_mn, bn)
    run_gemm(512, 768, 320, a_mn, b_mn, bn, seed=3)


def test_gemm_pair_odd_tiles_and_ragged(pair_forced):
    # odd m-tile counts (the surplus tile of the last pair reads zero fill and stores nothing), ragged M / N / K
    run_gemm(130, 300, 520, 0, 0, 64)
    run_gemm(257, 11313 // 8, 72, 0, 0, 128)
    run_gemm(300, 200, 1500, 1, 1, 128)
    run_gemm(641, 333, 192, 0, 1, 256, ldo=333)
    run_gemm(200, 100, 192, 0, 0, 32)


def test_gemm_pair_persistent_and_splitk(pair_forced):
    # more pair tiles than clusters: ring wrap, TMEM double buffer, remote tempty arrivals over many tiles
    run_gemm(8192, 2048, 192, 0, 0, 64)
    run_gemm(4096, 4096, 1024, 0, 0, 256)
    run_gemm(4096, 1024, 512, 1, 1, 256, out_bf16=1)
    run_gemm(512, 512, 4096, 0, 0, 128, splits=7)
    run_gemm(512, 1000, 1536, 1, 1, 256, splits=3, ldo=1001)


def test_gemm_pair_matches_single_bitwise():
    # same k order per tile, same fp32 accumulation: pairing changes which SM computes a tile, not the result
    lib = L.load()
    outs = []
    for mode in (0, 1):
        lib.vc_test_pair_mode(mode)
        try:
            g = torch.Generator(device="cpu").manual_seed(5)
            A = torch.randn(1024, 512, generator=g).to(torch.bfloat16).cuda()
            B = torch.randn(768, 512, generator=g).to(torch.bfloat16).cuda()
            out = torch.zeros(1024, 768, device="cuda")
            L.check(lib.vc_gemm_bf16(L.ptr(A), 0, ctypes.c_longlong(512), L.ptr(B), 0, ctypes.c_longlong(512), L.ptr(out),
                                     ctypes.c_longlong(768), None, 1024, 768, 512, 256, 1, 0, 0, 0, L.stream_ptr()))
            torch.cuda.synchronize()
            outs.append(out.clone())
        finally:
            lib.vc_test_pair_mode(-1)
    assert torch.equal(outs[0], outs[1])


@pytest.mark.parametrize("pair", [0, 1])
def test_gemm_dual_n_tiles_match_single_tiles_bitwise(pair):
    """Dual-N tiles (two adjacent 256-column n-tiles share one A tile; both accumulators fill the 512 TMEM columns, epilogue
    group g drains n-tile g) change the schedule, not the arithmetic: every output element sees the same k order. Covers
    K-major and MN-major operands, an odd n-tile count (the last dual tile's second half is all zero fill), ragged M / N,
    split-K with atomics, bf16 output, with and without CTA pairs."""
    lib = L.load()
    cases = [(1024, 1024, 2048, 0, 0, 1, 0), (768, 1280 + 40, 1024, 0, 1, 1, 0), (640, 512, 4096, 1, 1, 1, 0),
             (512, 768, 2048, 0, 0, 1, 1), (300, 700, 2048, 0, 0, 3, 0)]
    for (M, N, K, a_mn, b_mn, splits, out_bf16) in cases:
        g = torch.Generator(device="cpu").manual_seed(M + N)
        A = torch.randn(M, K, generator=g).to(torch.bfloat16)
        B = torch.randn(N, K, generator=g).to(torch.bfloat16)
        Ad = (A.t().contiguous() if a_mn else A).cuda()
        Bd = (B.t().contiguous() if b_mn else B).cuda()
        bias = torch.randn(N, generator=g).cuda()
        outs = []
        for dual in (0, 1):
            lib.vc_test_pair_mode(pair)
            lib.vc_test_dual_mode(dual)
            try:
                out = torch.zeros(M, N, dtype=torch.bfloat16 if out_bf16 else torch.float32, device="cuda")
                L.check(lib.vc_gemm_bf16(L.ptr(Ad), a_mn, ctypes.c_longlong(M if a_mn else K), L.ptr(Bd), b_mn,
                                         ctypes.c_longlong(N if b_mn else K), L.ptr(out), ctypes.c_longlong(N), L.ptr(bias), M, N, K,
                                         256, splits, 0, out_bf16, 1 if splits > 1 else 0, L.stream_ptr()))
                torch.cuda.synchronize()
                outs.append(out.float().clone())
            finally:
                lib.vc_test_pair_mode(-1)
                lib.vc_test_dual_mode(-1)
        ref = A.float().cuda() @ B.float().cuda().t() + bias
        tol = (2e-2 if out_bf16 else 2e-3) * ref.abs().max().item()
        assert (outs[1] - ref).abs().max().item() <= tol, (M, N, K, a_mn, b_mn, splits)
        if splits == 1:
            assert torch.equal(outs[0], outs[1]), (M, N, K, a_mn, b_mn)
        else:  # fp32 atomics meet in a different order
            assert (outs[0] - outs[1]).abs().max().item() <= 1e-3 * ref.abs().max().item()
