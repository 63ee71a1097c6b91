"""GPU parity of every kernel against plain fp32 PyTorch on the same (bf16-rounded) inputs.

Tolerances: GEMM/conv/attention outputs are bf16 with fp32 accumulation -> 1e-2 of max|ref|
(north_star tolerance); fp32-output and normalisation kernels are held tighter.
"""
import math

import numpy as np

import pytest
import torch
import torch.nn.functional as F

from gpu_util import report, setup_exact_fp32

pytestmark = pytest.mark.gpu

DEV = "cuda"


def _ops():
    from pytorch_stable_diffusion_b200 import ops
    return ops


def rnd(*shape, scale=1.0, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed + sum(shape))
    return (torch.randn(*shape, generator=g) * scale).to(DEV)


# ------------------------------------------------------------------------------------ GEMM
@pytest.mark.parametrize("M,K,N,block_n", [
    (128, 64, 128, 128),      # one tile, one k-block
    (256, 128, 64, 64),
    (8192, 320, 960, 160),    # UNet self-attention in_proj @64x64
    (2048, 1280, 320, 160),
    (512, 2560, 1280, 256),
    (154, 768, 320, 160),     # ragged M (context rows)
    (300, 320, 1280, 0),
])
def test_linear_plain(M, K, N, block_n):
    ops = _ops()
    setup_exact_fp32()
    a = rnd(M, K).bfloat16()
    w = rnd(N, K, scale=K ** -0.5, seed=1).bfloat16()
    out = ops.linear(a, w, block_n=block_n, nsplit=1)
    report(f"linear {M}x{K}x{N}", out, a.float() @ w.float().t(), 1e-2)


def test_linear_epilogue_bias_act_residual():
    ops = _ops()
    setup_exact_fp32()
    M, K, N = 1000, 640, 640
    a = rnd(M, K).bfloat16()
    w = rnd(N, K, scale=K ** -0.5, seed=1).bfloat16()
    b = rnd(N, seed=2)
    r = rnd(M, N, seed=3).bfloat16()
    ref = a.float() @ w.float().t() + b
    out = ops.linear(a, w, bias=b)
    report("linear+bias", out, ref, 1e-2)
    out = ops.linear(a, w, bias=b, residual=r)
    report("linear+bias+res", out, ref + r.float(), 1e-2)
    out = ops.linear(a, w, bias=b, act=ops.ACT_QUICK_GELU, residual=r)
    report("linear+bias+qgelu+res", out, ref * torch.sigmoid(1.702 * ref) + r.float(), 1e-2)
    out = ops.linear(a, w, bias=b, out_fp32=True)
    report("linear fp32 out", out, ref, 2e-3)
    out = ops.linear(a, w, bias=b, out_f16=True)
    assert out.dtype == torch.float16
    report("linear f16 out", out, ref, 2e-3)


def test_linear_small_n_fp32_and_row_bias():
    ops = _ops()
    setup_exact_fp32()
    M, K = 4096, 320
    a = rnd(M, K).bfloat16()
    w = rnd(4, K, scale=K ** -0.5, seed=1).bfloat16()
    b = rnd(4, seed=2)
    out = ops.linear(a, w, bias=b, out_fp32=True)
    report("linear N=4", out, a.float() @ w.float().t() + b, 2e-3)
    # swapped-operand projection (V^T = Wv . X^T + bv): bias per output row
    wv = rnd(512, K, scale=K ** -0.5, seed=4).bfloat16()
    bv = rnd(512, seed=5)
    out = ops.gemm(wv, a, M, kind=ops.GEMM_LINEAR, M=512, c0=K, bias=bv, bias_per_row=True)
    report("linear swapped + row bias", out, wv.float() @ a.float().t() + bv[:, None], 1e-2)


def test_linear_strided_and_dual_source_and_splitk():
    ops = _ops()
    setup_exact_fp32()
    M, C0, C1, N = 512, 640, 320, 640
    big = rnd(M, 2 * C0).bfloat16()
    a0 = big[:, C0:]                      # strided view, lda = 2*C0
    a1 = rnd(M, C1, seed=7).bfloat16()
    w = rnd(N, C0 + C1, scale=(C0 + C1) ** -0.5, seed=1).bfloat16()
    ref = torch.cat([a0.float(), a1.float()], 1) @ w.float().t()
    out = ops.gemm(a0, w, N, kind=ops.GEMM_LINEAR, a1=a1, M=M, c0=C0, c1=C1, lda0=2 * C0, nsplit=1)
    report("linear dual-source strided", out, ref, 1e-2)
    out = ops.gemm(a0, w, N, kind=ops.GEMM_LINEAR, a1=a1, M=M, c0=C0, c1=C1, lda0=2 * C0, nsplit=3)
    report("linear dual-source split-K 3", out, ref, 1e-2)
    # strided output / weight views (K^T-style scores): out into a column slice
    wide = torch.zeros(M, 2 * N, device=DEV, dtype=torch.bfloat16)
    ops.gemm(a1, w[:, :C1], N, kind=ops.GEMM_LINEAR, M=M, c0=C1, ldw=C0 + C1, out=wide[:, N:], ldo=2 * N,
             nsplit=1)
    report("linear ldw/ldo", wide[:, N:], a1.float() @ w[:, :C1].float().t(), 1e-2)
    assert float(wide[:, :N].abs().max()) == 0.0


@pytest.mark.parametrize("M,K,N", [(512, 640, 320), (1000, 320, 640), (256, 1280, 1280), (128, 320, 320)])
def test_linear_wide_tiles(M, K, N):
    """block_n = 320: two 160-column accumulators per tile sharing the A operand (single TMEM buffer)."""
    ops = _ops()
    setup_exact_fp32()
    a = rnd(M, K).bfloat16()
    w = rnd(N, K, scale=K ** -0.5, seed=1).bfloat16()
    b = rnd(N, seed=2)
    r = rnd(M, N, seed=3)
    ref = a.float() @ w.float().t() + b
    report(f"linear wide {M}x{K}x{N}", ops.linear(a, w, bias=b, block_n=320, nsplit=1), ref, 1e-2)
    out, out2 = ops.linear(a, w, bias=b, residual=r, out_fp32=True, out2=True, block_n=320, nsplit=1)
    report(f"linear wide fp32+res {M}x{K}x{N}", out, ref + r, 2e-3)
    assert torch.equal(out2, out.bfloat16())


# ------------------------------------------------------------------------------------ TMA epilogue
@pytest.mark.parametrize("M,K,N,block_n", [
    (4096, 320, 320, 160), (4096, 320, 320, 320), (1000, 320, 640, 160), (333, 640, 96, 96),
    (128, 64, 32, 32), (2048, 1280, 1280, 160), (70, 320, 48, 48),
])
def test_linear_tma_epilogue(M, K, N, block_n):
    """Bulk-tensor-store epilogue (epi_mode=2) against fp32 PyTorch and, bit for bit, against the per-lane
    store epilogue (epi_mode=1): bf16 out, fp32 out + fp32 residual (+ bf16 copy), activation, f16 out."""
    ops = _ops()
    setup_exact_fp32()
    a = rnd(M, K).bfloat16()
    w = rnd(N, K, scale=K ** -0.5, seed=1).bfloat16()
    b = rnd(N, seed=2)
    r = rnd(M, N, seed=3)
    ref = a.float() @ w.float().t() + b
    kw = dict(block_n=block_n, nsplit=1)
    got = ops.linear(a, w, bias=b, epi_mode=2, **kw)
    report(f"tma-epi bf16 {M}x{K}x{N}", got, ref, 1e-2)
    assert torch.equal(got, ops.linear(a, w, bias=b, epi_mode=1, **kw))
    got = ops.linear(a, w, epi_mode=2, **kw)
    report(f"tma-epi no bias {M}x{K}x{N}", got, ref - b, 1e-2)
    got, got2 = ops.linear(a, w, bias=b, residual=r, out_fp32=True, out2=True, epi_mode=2, **kw)
    report(f"tma-epi fp32+res {M}x{K}x{N}", got, ref + r, 2e-3)
    assert torch.equal(got2, got.bfloat16())
    old, old2 = ops.linear(a, w, bias=b, residual=r, out_fp32=True, out2=True, epi_mode=1, **kw)
    assert torch.equal(got, old) and torch.equal(got2, old2)
    got = ops.linear(a, w, bias=b, residual=r, out_fp32=True, epi_mode=2, **kw)
    assert torch.equal(got, old)
    got = ops.linear(a, w, bias=b, out_fp32=True, act=ops.ACT_QUICK_GELU, epi_mode=2, **kw)
    report(f"tma-epi fp32 qgelu {M}x{K}x{N}", got, ref * torch.sigmoid(1.702 * ref), 2e-3)
    got = ops.linear(a, w, bias=b, residual=r, epi_mode=2, **kw)          # bf16 out + fp32 residual
    report(f"tma-epi bf16+res {M}x{K}x{N}", got, ref + r, 1e-2)
    if N % 32 == 0:
        got = ops.linear(a, w, bias=b, out_f16=True, epi_mode=2, **kw)
        report(f"tma-epi f16 {M}x{K}x{N}", got, ref, 2e-3)
    # IEEE-half residual + 16-bit result (the token stream of the attention blocks): TMA epilogue with 2 KB residual
    # slots, bit for bit equal to the per-lane epilogue
    rh = r.half()
    for o16, tol in ((torch.float16, 2e-3), (torch.bfloat16, 1e-2)):
        got = ops.linear(a, w, bias=b, residual=rh, out16=o16, epi_mode=2, **kw)
        assert got.dtype == o16
        report(f"tma-epi {o16} + half residual {M}x{K}x{N}", got, ref + rh.float(), tol)
        assert torch.equal(got, ops.linear(a, w, bias=b, residual=rh, out16=o16, epi_mode=1, **kw))
    got = ops.linear(a, w, bias=b, residual=rh, out16=torch.float16, act=ops.ACT_QUICK_GELU, epi_mode=2, **kw)
    report(f"tma-epi f16 qgelu + half residual {M}x{K}x{N}", got, ref * torch.sigmoid(1.702 * ref) + rh.float(), 2e-3)
    got = ops.linear(a, w, bias=b, residual=rh, out_fp32=True, **kw)     # fp32 out + half residual: per-lane path
    report(f"fp32 out + half residual {M}x{K}x{N}", got, ref + rh.float(), 2e-3)


def test_linear_tma_epilogue_row_bias_and_strided_out():
    ops = _ops()
    setup_exact_fp32()
    M, K, N = 512, 320, 1000
    x = rnd(N, K).bfloat16()
    wv = rnd(M, K, scale=K ** -0.5, seed=4).bfloat16()
    bv = rnd(M, seed=5)
    ref = wv.float() @ x.float().t() + bv[:, None]
    out = ops.gemm(wv, x, N, kind=ops.GEMM_LINEAR, M=M, c0=K, bias=bv, bias_per_row=True, epi_mode=2)
    report("tma-epi swapped + row bias", out, ref, 1e-2)
    wide = torch.zeros(M, 2048, device=DEV, dtype=torch.bfloat16)
    ops.gemm(wv, x, N, kind=ops.GEMM_LINEAR, M=M, c0=K, out=wide[:, 1024:1024 + N], ldo=2048, nsplit=1, epi_mode=2)
    report("tma-epi ldo", wide[:, 1024:1024 + N], ref - bv[:, None], 1e-2)
    assert float(wide[:, :1024].abs().max()) == 0.0 and float(wide[:, 1024 + N:].abs().max()) == 0.0


@pytest.mark.parametrize("NB,H,W,C0,C1,Cout,kind", [
    (2, 64, 64, 64, 0, 320, "s1"), (3, 32, 32, 128, 64, 160, "s1"), (2, 16, 16, 192, 0, 96, "s1"),
    (5, 8, 8, 128, 0, 64, "s1"), (1, 8, 8, 64, 0, 32, "s1"), (2, 96, 96, 64, 0, 64, "s1"),
    (2, 48, 48, 64, 0, 64, "s1"), (2, 24, 24, 64, 0, 64, "s1"), (2, 64, 64, 64, 0, 64, "s2"),
    (3, 4, 4, 64, 0, 64, "s1"),
])
def test_conv3x3_tma_epilogue(NB, H, W, C0, C1, Cout, kind):
    """Every tile-box geometry the UNet / VAE levels produce (and the ineligible ones, which must fall back)."""
    ops = _ops()
    setup_exact_fp32()
    x0 = rnd(NB, H, W, C0).bfloat16()
    x1 = rnd(NB, H, W, C1, seed=5).bfloat16() if C1 else None
    w = rnd(Cout, C0 + C1, 3, 3, scale=(9 * (C0 + C1)) ** -0.5, seed=1).bfloat16()
    b = rnd(Cout, seed=2)
    xin = x0.float() if x1 is None else torch.cat([x0.float(), x1.float()], -1)
    s2 = kind == "s2"
    ref = F.conv2d(xin.permute(0, 3, 1, 2), w.float(), b, padding=1, stride=2 if s2 else 1).permute(0, 2, 3, 1)
    ho, wo = ref.shape[1], ref.shape[2]
    ref = ref.reshape(-1, Cout)
    r = rnd(NB * ho * wo, Cout, seed=3)
    k = ops.GEMM_CONV3X3_S2 if s2 else ops.GEMM_CONV3X3_S1
    kw = dict(kind=k, a1=x1, bias=b, conv_dims=(NB, H, W), c0=C0, c1=C1, nsplit=1)
    got, got2 = ops.gemm(x0, pack3x3(w), Cout, residual=r, out_fp32=True, out2=True, epi_mode=2, **kw)
    report(f"conv tma-epi fp32+res {C0 + C1}->{Cout}@{H}", got, ref + r, 3e-3)
    assert torch.equal(got2, got.bfloat16())
    old, old2 = ops.gemm(x0, pack3x3(w), Cout, residual=r, out_fp32=True, out2=True, epi_mode=1, **kw)
    assert torch.equal(got, old) and torch.equal(got2, old2)
    got = ops.gemm(x0, pack3x3(w), Cout, epi_mode=2, **kw)
    report(f"conv tma-epi bf16 {C0 + C1}->{Cout}@{H}", got, ref, 1e-2)
    assert torch.equal(got, ops.gemm(x0, pack3x3(w), Cout, epi_mode=1, **kw))


# ------------------------------------------------------------------------------------ GN statistics in the epilogue
def _check_partials(part, y, n, name):
    """part [n][K][C][2] must add up (over the slabs) to the per-sample column sums / sums of squares of y."""
    C = y.shape[-1]
    yv = y.view(n, -1, C).double()
    got = part.double().sum(1)
    ref_s, ref_q = yv.sum(1), (yv * yv).sum(1)
    report(f"{name} partial sums", got[..., 0], ref_s, 1e-5)
    report(f"{name} partial sums of squares", got[..., 1], ref_q, 1e-5)


@pytest.mark.parametrize("n,hw,K,N", [(2, 4096, 320, 320), (3, 1024, 640, 640), (2, 256, 2560, 1280), (4, 64, 1280, 1280)])
def test_linear_gn_partials(n, hw, K, N):
    """Short reductions take the TMA epilogue, K = 2560 the per-lane one: both must leave the statistics."""
    ops = _ops()
    setup_exact_fp32()
    M = n * hw
    a = rnd(M, K).bfloat16()
    w = rnd(N, K, scale=K ** -0.5, seed=1).bfloat16()
    b = rnd(N, seed=2)
    r = rnd(M, N, seed=3)
    out, out2, part = ops.linear(a, w, bias=b, residual=r, out_fp32=True, out2=True, nsplit=1, gn_samples=n)
    assert part is not None and part.shape == (n, hw // 32, N, 2)
    report(f"linear gn {M}x{K}x{N}", out, a.float() @ w.float().t() + b + r, 2e-3)
    _check_partials(part, out, n, f"linear {M}x{K}x{N}")
    assert torch.equal(out2, out.bfloat16())
    g = rnd(N, seed=4) + 1.0
    be = rnd(N, seed=5)
    y = ops.groupnorm(out.view(n, hw, N), g, be, silu=True, fused=False, part0=part)
    ref = F.silu(F.group_norm(out.view(n, hw, N).permute(0, 2, 1), 32, g, be, 1e-5).permute(0, 2, 1))
    report(f"groupnorm from epilogue partials {M}x{N}", y, ref, 6e-3)


@pytest.mark.parametrize("NB,H,C0,C1,Cout", [(2, 64, 64, 0, 320), (3, 32, 128, 64, 640), (2, 16, 128, 0, 1280),
                                             (5, 8, 64, 0, 320), (2, 96, 64, 0, 64), (2, 48, 64, 0, 64)])
def test_conv3x3_gn_partials(NB, H, C0, C1, Cout):
    ops = _ops()
    setup_exact_fp32()
    x0 = rnd(NB, H, H, C0).bfloat16()
    x1 = rnd(NB, H, H, C1, seed=5).bfloat16() if C1 else None
    w = rnd(Cout, C0 + C1, 3, 3, scale=(9 * (C0 + C1)) ** -0.5, seed=1).bfloat16()
    b = rnd(Cout, seed=2)
    r = rnd(NB * H * H, Cout, seed=3)
    out, out2, part = ops.gemm(x0, pack3x3(w), Cout, kind=ops.GEMM_CONV3X3_S1, a1=x1, bias=b, residual=r,
                               conv_dims=(NB, H, H), c0=C0, c1=C1, out_fp32=True, out2=True, nsplit=1, gn_samples=NB)
    assert part is not None
    xin = x0.float() if x1 is None else torch.cat([x0.float(), x1.float()], -1)
    ref = F.conv2d(xin.permute(0, 3, 1, 2), w.float(), b, padding=1).permute(0, 2, 3, 1).reshape(-1, Cout) + r
    report(f"conv gn {C0 + C1}->{Cout}@{H}", out, ref, 3e-3)
    _check_partials(part, out, NB, f"conv {C0 + C1}->{Cout}@{H}")
    # channel concat of two producers' partials
    g = rnd(2 * Cout, seed=4) + 1.0
    be = rnd(2 * Cout, seed=5)
    o3 = out.view(NB, H * H, Cout)
    y = ops.groupnorm(o3, g, be, x1=o3, silu=False, fused=False, part0=part, part1=part)
    refn = F.group_norm(torch.cat([o3, o3], -1).permute(0, 2, 1), 32, g, be, 1e-5).permute(0, 2, 1)
    report(f"groupnorm concat from partials {Cout}+{Cout}@{H}", y, refn, 6e-3)


@pytest.mark.parametrize("N,H,C,Cx0,Cx1,Cout,nsplit", [(2, 64, 320, 640, 320, 320, 1), (2, 32, 640, 320, 0, 640, 1),
                                                       (2, 16, 1280, 1280, 640, 1280, 1), (2, 8, 1280, 1280, 1280, 1280, 3),
                                                       (3, 16, 128, 64, 0, 128, 1), (2, 24, 128, 128, 64, 128, 1),
                                                       (2, 12, 128, 64, 64, 128, 1), (2, 48, 64, 128, 0, 64, 1),
                                                       (2, 96, 64, 64, 0, 64, 1), (2, 24, 1280, 1280, 640, 1280, 1),
                                                       (2, 12, 1280, 1280, 1280, 1280, 8), (2, 12, 1280, 1280, 1280, 1280, 6),
                                                       (2, 12, 1280, 1280, 1280, 1280, 0), (2, 24, 1280, 1280, 1280, 1280, 0),
                                                       (2, 24, 1280, 1280, 640, 1280, 4), (2, 8, 1280, 1280, 1280, 1280, 8),
                                                       (2, 48, 640, 1280, 640, 640, 0), (2, 48, 640, 320, 0, 640, 0),
                                                       (2, 96, 320, 640, 320, 320, 0)])
def test_conv3x3_extra_1x1_source(N, H, C, Cx0, Cx1, Cout, nsplit):
    """conv3x3(a) + conv1x1(x0 ++ x1) in ONE GEMM (the resblock's skip convolution as extra k-blocks)."""
    ops = _ops()
    setup_exact_fp32()
    a = rnd(N, H, H, C).bfloat16()
    x0 = rnd(N, H, H, Cx0, seed=6).bfloat16()
    x1 = rnd(N, H, H, Cx1, seed=7).bfloat16() if Cx1 else None
    w3 = rnd(Cout, C, 3, 3, scale=(9 * C) ** -0.5, seed=1).bfloat16()
    w1 = rnd(Cout, Cx0 + Cx1, scale=(Cx0 + Cx1) ** -0.5, seed=8).bfloat16()
    b = rnd(Cout, seed=2)
    xin = x0.float() if x1 is None else torch.cat([x0.float(), x1.float()], -1)
    ref = (F.conv2d(a.float().permute(0, 3, 1, 2), w3.float(), b, padding=1) +
           F.conv2d(xin.permute(0, 3, 1, 2), w1.float()[:, :, None, None])).permute(0, 2, 3, 1).reshape(-1, Cout)
    wcat = torch.cat([pack3x3(w3), w1], dim=1).contiguous()
    out = ops.gemm(a, wcat, Cout, kind=ops.GEMM_CONV3X3_S1, bias=b, conv_dims=(N, H, H), c0=C, out_fp32=True,
                   nsplit=nsplit, ax0=x0, ax1=x1)
    report(f"conv3x3 + 1x1 source {C}+{Cx0 + Cx1}->{Cout}@{H} nsplit={nsplit}", out, ref, 3e-3)
    out, out2, part = ops.gemm(a, wcat, Cout, kind=ops.GEMM_CONV3X3_S1, bias=b, conv_dims=(N, H, H), c0=C,
                               out_fp32=True, out2=True, nsplit=1, ax0=x0, ax1=x1, gn_samples=N)
    report(f"conv3x3 + 1x1 source, gn partials {C}+{Cx0 + Cx1}->{Cout}@{H}", out, ref, 3e-3)
    if part is not None:
        _check_partials(part, out, N, "conv3x3 + 1x1 source")


# ------------------------------------------------------------------------------------ conv
def pack3x3(w):
    return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).contiguous()


@pytest.mark.parametrize("N,H,W,Cin,Cout", [
    (1, 16, 16, 64, 64),
    (2, 64, 64, 320, 320),
    (2, 32, 32, 640, 640),
    (2, 16, 16, 1280, 1280),
    (2, 8, 8, 1280, 1280),
    (3, 8, 8, 128, 128),       # odd batch with 2 images per tile
    (2, 24, 24, 128, 256),     # 768^2 latent level
    (2, 12, 12, 128, 128),
    (1, 96, 96, 64, 128),
    (2, 64, 64, 320, 4),       # UNet output conv (fp32 out handled below)
])
def test_conv3x3_s1(N, H, W, Cin, Cout):
    ops = _ops()
    setup_exact_fp32()
    x = rnd(N, H, W, Cin).bfloat16()
    w = rnd(Cout, Cin, 3, 3, scale=(9 * Cin) ** -0.5, seed=1).bfloat16()
    b = rnd(Cout, seed=2)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), b, padding=1).permute(0, 2, 3, 1)
    out = ops.conv3x3(x, pack3x3(w), Cout, bias=b, nsplit=1, out_fp32=(Cout == 4))
    report(f"conv3x3 s1 {N}x{H}x{W} {Cin}->{Cout}", out, ref, 1e-2 if Cout != 4 else 3e-3)


@pytest.mark.parametrize("N,H,W,C0,C1,Cout", [(2, 64, 64, 320, 0, 320), (2, 32, 32, 640, 320, 640), (2, 16, 16, 1280, 0, 1280)])
def test_conv3x3_wide_tiles(N, H, W, C0, C1, Cout):
    ops = _ops()
    setup_exact_fp32()
    x0 = rnd(N, H, W, C0).bfloat16()
    x1 = rnd(N, H, W, C1, seed=5).bfloat16() if C1 else None
    w = rnd(Cout, C0 + C1, 3, 3, scale=(9 * (C0 + C1)) ** -0.5, seed=1).bfloat16()
    b = rnd(Cout, seed=2)
    r = rnd(N * H * W, Cout, seed=3)
    xin = x0.float() if x1 is None else torch.cat([x0.float(), x1.float()], -1)
    ref = F.conv2d(xin.permute(0, 3, 1, 2), w.float(), b, padding=1).permute(0, 2, 3, 1).reshape(-1, Cout) + r
    out, out2 = ops.gemm(x0, pack3x3(w), Cout, kind=ops.GEMM_CONV3X3_S1, a1=x1, bias=b, residual=r,
                         conv_dims=(N, H, W), c0=C0, c1=C1, out_fp32=True, out2=True, block_n=320, nsplit=1)
    report(f"conv3x3 wide {C0 + C1}->{Cout}@{H}", out, ref, 3e-3)
    assert torch.equal(out2, out.bfloat16())


def test_conv3x3_residual_dual_source_splitk():
    ops = _ops()
    setup_exact_fp32()
    N, H, W, C0, C1, Cout = 2, 16, 16, 640, 320, 640
    x0 = rnd(N, H, W, C0).bfloat16()
    x1 = rnd(N, H, W, C1, seed=5).bfloat16()
    w = rnd(Cout, C0 + C1, 3, 3, scale=(9 * (C0 + C1)) ** -0.5, seed=1).bfloat16()
    b = rnd(Cout, seed=2)
    r = rnd(N, H, W, Cout, seed=3).bfloat16()
    xin = torch.cat([x0, x1], -1).float().permute(0, 3, 1, 2)
    ref = F.conv2d(xin, w.float(), b, padding=1).permute(0, 2, 3, 1) + r.float()
    out = ops.conv3x3(x0, pack3x3(w), Cout, bias=b, a1=x1, c1=C1, residual=r.view(-1, Cout), nsplit=1)
    report("conv3x3 dual-source + residual", out, ref, 1e-2)
    out = ops.conv3x3(x0, pack3x3(w), Cout, bias=b, a1=x1, c1=C1, residual=r.view(-1, Cout), nsplit=4)
    report("conv3x3 dual-source + residual split-K", out, ref, 1e-2)
    out = ops.conv3x3(x0, pack3x3(w), Cout, bias=b, a1=x1, c1=C1, residual=r.view(-1, Cout))
    report("conv3x3 dual-source + residual auto split", out, ref, 1e-2)


@pytest.mark.parametrize("N,H,W,C,Cout", [(2, 64, 64, 320, 320), (2, 16, 16, 1280, 1280), (1, 32, 32, 128, 128)])
def test_conv3x3_s2(N, H, W, C, Cout):
    ops = _ops()
    setup_exact_fp32()
    x = rnd(N, H, W, C).bfloat16()
    w = rnd(Cout, C, 3, 3, scale=(9 * C) ** -0.5, seed=1).bfloat16()
    b = rnd(Cout, seed=2)
    xin = x.float().permute(0, 3, 1, 2)
    ref = F.conv2d(xin, w.float(), b, stride=2, padding=1).permute(0, 2, 3, 1)
    out = ops.conv3x3(x, pack3x3(w), Cout, bias=b, kind=ops.GEMM_CONV3X3_S2, nsplit=1)
    report(f"conv3x3 s2 pad1 {H}x{W}", out, ref, 1e-2)
    ref = F.conv2d(F.pad(xin, (0, 1, 0, 1)), w.float(), b, stride=2).permute(0, 2, 3, 1)
    out = ops.conv3x3(x, pack3x3(w), Cout, bias=b, kind=ops.GEMM_CONV3X3_S2_PAD_RB, nsplit=1)
    report(f"conv3x3 s2 pad-rb {H}x{W}", out, ref, 1e-2)


@pytest.mark.parametrize("Cin,Cout,k", [(4, 320, 3), (4, 4, 1), (3, 128, 3), (8, 8, 1), (4, 512, 3)])
def test_conv_direct(Cin, Cout, k):
    ops = _ops()
    setup_exact_fp32()
    N, H, W = 2, 32, 32
    x = rnd(N, H, W, Cin).bfloat16()
    w = rnd(Cout, Cin, k, k, scale=(k * k * Cin) ** -0.5, seed=1)
    b = rnd(Cout, seed=2)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w, b, padding=(k - 1) // 2).permute(0, 2, 3, 1)
    wp = w.permute(0, 2, 3, 1).reshape(Cout, k * k, Cin).contiguous()
    out, shadow = ops.conv_direct(x, wp, b, Cout, k, out_fp32=True, out2=True)
    report(f"conv_direct {Cin}->{Cout} k{k}", out, ref, 1e-4)
    assert torch.equal(shadow, out.bfloat16())
    out = ops.conv_direct(x, wp, b, Cout, k)
    report(f"conv_direct bf16 {Cin}->{Cout} k{k}", out, ref, 1e-2)
    out = ops.conv_direct(x.float(), wp, b, Cout, k, out_fp32=True)          # fp32 input path
    report(f"conv_direct fp32 in {Cin}->{Cout} k{k}", out, ref, 1e-4)


def test_conv_direct_odd_width_takes_the_one_pixel_kernel():
    ops = _ops()
    setup_exact_fp32()
    N, H, W, Cin, Cout, k = 2, 18, 30, 4, 320, 3
    x = rnd(N, H, W, Cin)
    w = rnd(Cout, Cin, k, k, scale=(k * k * Cin) ** -0.5, seed=1)
    b = rnd(Cout, seed=2)
    ref = F.conv2d(x.permute(0, 3, 1, 2), w, b, padding=1).permute(0, 2, 3, 1)
    wp = w.permute(0, 2, 3, 1).reshape(Cout, k * k, Cin).contiguous()
    out, shadow = ops.conv_direct(x, wp, b, Cout, k, out_fp32=True, out2=True)
    report("conv_direct W=30", out, ref, 1e-4)
    assert torch.equal(shadow, out.bfloat16())


# ------------------------------------------------------------------------------------ norms
@pytest.mark.parametrize("N,HW,C0,C1,eps,silu", [
    (2, 4096, 320, 0, 1e-5, True), (2, 64, 1280, 1280, 1e-5, True), (2, 1024, 640, 320, 1e-5, True),
    (2, 256, 1280, 0, 1e-6, False), (1, 16384, 128, 0, 1e-5, True), (2, 1024, 1280, 640, 1e-5, True),
])
def test_groupnorm(N, HW, C0, C1, eps, silu):
    ops = _ops()
    x0 = (rnd(N, HW, C0) * 2 + 0.5).bfloat16()
    x1 = (rnd(N, HW, C1, seed=3) - 1.0).bfloat16() if C1 else None
    C = C0 + C1
    g = rnd(C, seed=4) + 1.0
    b = rnd(C, seed=5)
    out = ops.groupnorm(x0, g, b, x1=x1, eps=eps, silu=silu)
    xc = torch.cat([x0, x1], -1) if C1 else x0
    ref = F.group_norm(xc.float().permute(0, 2, 1), 32, g, b, eps).permute(0, 2, 1)
    if silu:
        ref = F.silu(ref)
    report(f"groupnorm C={C0}+{C1} HW={HW}", out, ref, 6e-3)


@pytest.mark.parametrize("N,HW,C0,C1,silu", [
    (2, 4096, 320, 0, True), (2, 4096, 320, 320, True), (3, 1024, 640, 0, False), (2, 1024, 1280, 640, True),
    (2, 1024, 640, 320, True), (2, 256, 1280, 0, True), (2, 256, 1280, 1280, True), (5, 64, 1280, 0, True),
    (2, 64, 1280, 1280, False), (1, 4096, 512, 0, True), (2, 9216, 320, 0, True), (2, 2304, 640, 0, True),
    (2, 144, 1280, 0, True), (2, 1000, 320, 0, True),
])
def test_groupnorm_fused(N, HW, C0, C1, silu):
    """One-pass cluster GroupNorm (fp32 in, bf16 out) against F.group_norm and against the two-kernel path."""
    ops = _ops()
    from pytorch_stable_diffusion_b200 import _ext
    assert _ext.lib().sdb_groupnorm_fused_supported(HW, C0, C1, 32) >= 1
    x0 = rnd(N, HW, C0) * 2 + 0.5
    x1 = (rnd(N, HW, C1, seed=3) - 1.0) if C1 else None
    C = C0 + C1
    g = rnd(C, seed=4) + 1.0
    b = rnd(C, seed=5)
    out = ops.groupnorm(x0, g, b, x1=x1, eps=1e-5, silu=silu, fused=True)
    xc = torch.cat([x0, x1], -1) if C1 else x0
    ref = F.group_norm(xc.permute(0, 2, 1), 32, g, b, 1e-5).permute(0, 2, 1)
    if silu:
        ref = F.silu(ref)
    report(f"groupnorm fused C={C0}+{C1} HW={HW}", out, ref, 6e-3)
    two = ops.groupnorm(x0, g, b, x1=x1, eps=1e-5, silu=silu, fused=False)
    report(f"groupnorm fused vs two-pass C={C0}+{C1} HW={HW}", out, two, 8e-3)
    again = ops.groupnorm(x0, g, b, x1=x1, eps=1e-5, silu=silu, fused=True)
    assert torch.equal(out, again)          # fixed summation order: bit-reproducible


def test_groupnorm_fused_unsupported_shapes_fall_back():
    ops = _ops()
    from pytorch_stable_diffusion_b200 import _ext
    lib = _ext.lib()
    assert lib.sdb_groupnorm_fused_supported(4096, 640, 320, 32) == 0       # 30 channels per group at 64x64
    assert lib.sdb_groupnorm_fused_supported(512 * 512, 128, 0, 32) == 0    # VAE full resolution
    x0 = rnd(1, 4096, 640)
    x1 = rnd(1, 4096, 320, seed=3)
    g = rnd(960, seed=4) + 1.0
    b = rnd(960, seed=5)
    out = ops.groupnorm(x0, g, b, x1=x1, silu=True)
    ref = F.silu(F.group_norm(torch.cat([x0, x1], -1).permute(0, 2, 1), 32, g, b, 1e-5).permute(0, 2, 1))
    report("groupnorm fallback 960@64", out, ref, 6e-3)


@pytest.mark.parametrize("rows,C", [(8192, 320), (2048, 640), (512, 1280), (160, 768)])
def test_layernorm(rows, C):
    ops = _ops()
    x = (rnd(rows, C) * 3 + 1).bfloat16()
    g = rnd(C, seed=4) + 1.0
    b = rnd(C, seed=5)
    ref = F.layer_norm(x.float(), (C,), g, b, 1e-5)
    report(f"layernorm {rows}x{C}", ops.layernorm(x, g, b), ref, 6e-3)
    report(f"layernorm fp32 {rows}x{C}", ops.layernorm(x, g, b, out_fp32=True), ref, 1e-5)
    xh = (rnd(rows + 3, C, seed=6) * 3 + 1).half()                        # IEEE-half rows (ragged row count)
    refh = F.layer_norm(xh.float(), (C,), g, b, 1e-5)
    report(f"layernorm half in {rows}x{C}", ops.layernorm(xh, g, b, out_dtype=torch.float16), refh, 2e-3)
    report(f"layernorm half in, bf16 out {rows}x{C}", ops.layernorm(xh, g, b), refh, 6e-3)
    report(f"layernorm half in, fp32 out {rows}x{C}", ops.layernorm(xh, g, b, out_fp32=True), refh, 1e-5)


def test_softmax_rows():
    ops = _ops()
    s = rnd(300, 4096) * 20
    ref = torch.softmax(s / math.sqrt(512), -1)
    report("softmax_rows", ops.softmax_rows(s, 1 / math.sqrt(512)), ref, 5e-3)


# ------------------------------------------------------------------------------------ attention
def ref_attention(q, k, v, causal):
    # q: [N, h, S, d], k/v: [N, h, T, d] fp32
    d = q.shape[-1]
    w = q @ k.transpose(-1, -2)
    if causal:
        mask = torch.ones_like(w, dtype=torch.bool).triu(1)
        w.masked_fill_(mask, -torch.inf)
    w = w / math.sqrt(d)
    return torch.softmax(w, -1) @ v


@pytest.mark.parametrize("N,heads,d,S,Skv,causal,amp", [
    (1, 1, 64, 128, 128, False, 1.0),
    (2, 8, 40, 4096, 4096, False, 1.0),
    (2, 8, 80, 1024, 1024, False, 1.0),
    (2, 8, 160, 256, 256, False, 1.0),
    (2, 8, 160, 64, 64, False, 1.0),
    (2, 8, 40, 4096, 77, False, 1.0),     # cross-attention over the 77 CLIP tokens
    (2, 8, 160, 64, 77, False, 1.0),
    (2, 8, 80, 1024, 77, False, 1.0),     # multi-tile form of the one-tile kernel: 4 tiles per CTA
    (2, 8, 40, 1000, 77, False, 1.0),     # ... ragged last tile
    (3, 8, 40, 256, 77, False, 1.0),      # ... 2 tiles per CTA
    (2, 8, 40, 640, 100, False, 2.0),     # ... 5 tiles, 2 per CTA: the last CTA walks one
    (2, 8, 160, 256, 77, False, 1.0),     # d = 160: two key stages, one tile per CTA
    (2, 12, 64, 77, 77, True, 1.0),       # CLIP causal self-attention
    (2, 8, 40, 1024, 1024, False, 6.0),   # peaky logits: exercises the lazy O rescale
    (2, 8, 80, 576, 576, False, 1.0),     # 768^2 levels (ragged tiles)
    (1, 8, 160, 144, 144, False, 1.0),
])
def test_attention(N, heads, d, S, Skv, causal, amp):
    ops = _ops()
    setup_exact_fp32()
    C = heads * d
    Skv_pad = (Skv + 7) // 8 * 8
    q = (rnd(N * S, C) * amp).bfloat16()
    k = (rnd(N, Skv_pad, C, seed=1) * amp).bfloat16()
    v = rnd(N, Skv_pad, C, seed=2).bfloat16()
    vt = v.permute(2, 0, 1).contiguous()          # [C, N, Skv_pad]
    out = torch.empty(N * S, C, device=DEV, dtype=torch.bfloat16)
    ops.attention(q, k.view(N * Skv_pad, C), vt, out, NB=N, heads=heads, d=d, S=S, Skv=Skv,
                  Skv_pad=Skv_pad, ldq=C, ldk=C, ldo=C, causal=causal)
    qf = q.float().view(N, S, heads, d).transpose(1, 2)
    kf = k.float()[:, :Skv].reshape(N, Skv, heads, d).transpose(1, 2)
    vf = v.float()[:, :Skv].reshape(N, Skv, heads, d).transpose(1, 2)
    ref = ref_attention(qf, kf, vf, causal).transpose(1, 2).reshape(N * S, C)
    report(f"attention d={d} S={S} Skv={Skv} causal={causal}", out, ref, 1e-2)
    # all exponentials on the MUFU (exp_poly=1 switches the polynomial quarter off): same answer to bf16 rounding
    out2 = torch.empty_like(out)
    ops.attention(q, k.view(N * Skv_pad, C), vt, out2, NB=N, heads=heads, d=d, S=S, Skv=Skv,
                  Skv_pad=Skv_pad, ldq=C, ldk=C, ldo=C, causal=causal, exp_poly=1)
    report(f"attention (MUFU only) d={d} S={S} Skv={Skv} causal={causal}", out2, ref, 1e-2)
    report(f"attention poly vs MUFU d={d} S={S} Skv={Skv}", out, out2, 8e-3)


@pytest.mark.parametrize("d,S,Skv,p_f16", [
    (40, 4096, 4096, False), (40, 4096, 4096, True), (80, 1024, 1024, False), (80, 1024, 1024, True),
    (40, 1024, 77, False), (40, 1024, 77, True), (80, 576, 576, True), (40, 64, 64, True), (40, 1024, 1024, True),
    (40, 1024, 1024, False), (40, 9216, 9216, False), (40, 1000, 1000, False), (48, 512, 512, False),
])
def test_attention_sum_row(d, S, Skv, p_f16):
    """Denominator accumulated by the P.V product through a ones row in V^T; optionally f16x2 exps."""
    ops = _ops()
    setup_exact_fp32()
    N, heads = 2, 8
    C = heads * d
    R = (d + 1 + 15) // 16 * 16
    Skv_pad = (Skv + 7) // 8 * 8
    amp = 3.0 if (S == 1024 and Skv == 1024) else 1.0
    q = (rnd(N * S, C) * amp).bfloat16()
    k = (rnd(N, Skv_pad, C, seed=1) * amp).bfloat16()
    v = rnd(N, Skv_pad, C, seed=2).bfloat16()
    vt = torch.zeros(heads, R, N, Skv_pad, device=DEV, dtype=torch.bfloat16)
    vt[:, :d] = v.view(N, Skv_pad, heads, d).permute(2, 3, 0, 1)
    vt[:, d] = 1.0
    vt = vt.view(heads * R, N, Skv_pad).contiguous()
    if p_f16:
        vt = vt.to(torch.float16)          # exact: bf16 values of this magnitude are representable in half
    out = torch.empty(N * S, C, device=DEV, dtype=torch.bfloat16)
    ops.attention(q, k.view(N * Skv_pad, C), vt, out, NB=N, heads=heads, d=d, S=S, Skv=Skv,
                  Skv_pad=Skv_pad, ldq=C, ldk=C, ldo=C, sum_row=True, p_f16=p_f16)
    qf = q.float().view(N, S, heads, d).transpose(1, 2)
    kf = k.float()[:, :Skv].reshape(N, Skv, heads, d).transpose(1, 2)
    vf = v.float()[:, :Skv].reshape(N, Skv, heads, d).transpose(1, 2)
    ref = ref_attention(qf, kf, vf, False).transpose(1, 2).reshape(N * S, C)
    report(f"attention sum_row d={d} S={S} Skv={Skv} p_f16={p_f16}", out, ref, 1e-2)
    if not p_f16:
        # queries pre-multiplied by log2(e)/sqrt(d): the kernel takes the scores as log2 units
        qs = (q.float() * (math.log2(math.e) / math.sqrt(d))).bfloat16()
        qsf = qs.float().view(N, S, heads, d).transpose(1, 2) * (math.sqrt(d) / math.log2(math.e))
        ref2 = ref_attention(qsf, kf, vf, False).transpose(1, 2).reshape(N * S, C)
        out2 = torch.empty_like(out)
        ops.attention(qs, k.view(N * Skv_pad, C), vt, out2, NB=N, heads=heads, d=d, S=S, Skv=Skv,
                      Skv_pad=Skv_pad, ldq=C, ldk=C, ldo=C, sum_row=True, q_prescaled=True)
        report(f"attention sum_row prescaled d={d} S={S} Skv={Skv}", out2, ref2, 1e-2)


def test_attention_growing_maximum():
    """Keys whose scores grow block after block (each block raises the row maximum by far more than the 2^8
    lazy-rescale bound): the O / denominator rescale path must give the reference softmax."""
    ops = _ops()
    setup_exact_fp32()
    N, heads, d, S = 2, 8, 40, 1024
    C = heads * d
    R = 48
    q = rnd(N * S, C).bfloat16()
    k = rnd(N, S, C, seed=1)
    ramp = torch.linspace(0.2, 6.0, S, device=DEV).view(1, S, 1)          # later key blocks: larger |scores|
    k = (k * ramp).bfloat16()
    v = rnd(N, S, C, seed=2).bfloat16()
    vt = torch.zeros(heads, R, N, S, device=DEV, dtype=torch.bfloat16)
    vt[:, :d] = v.view(N, S, heads, d).permute(2, 3, 0, 1)
    vt[:, d] = 1.0
    vt = vt.view(heads * R, N, S).contiguous()
    out = torch.empty(N * S, C, device=DEV, dtype=torch.bfloat16)
    ops.attention(q, k.view(N * S, C), vt, out, NB=N, heads=heads, d=d, S=S, Skv=S, Skv_pad=S, ldq=C, ldk=C, ldo=C,
                  sum_row=True)
    qf = q.float().view(N, S, heads, d).transpose(1, 2)
    kf = k.float().reshape(N, S, heads, d).transpose(1, 2)
    vf = v.float().reshape(N, S, heads, d).transpose(1, 2)
    ref = ref_attention(qf, kf, vf, False).transpose(1, 2).reshape(N * S, C)
    report("attention, growing maximum", out, ref, 1e-2)


@pytest.mark.parametrize("S,amp,ramp", [(4096, 1.0, False), (1024, 3.0, False), (1024, 1.0, True), (1000, 2.0, False),
                                        (256, 1.0, False), (384, 1.0, True), (2304, 2.0, True), (9216, 1.0, False)])
def test_attention_qk_fold(S, amp, ramp):
    """Row offset of the softmax folded into Q.K^T (ones column in K, offset column written into the Q tile by the
    kernel): padded 48-column heads, pre-scaled queries, ones-row denominator."""
    ops = _ops()
    setup_exact_fp32()
    N, heads, d, R = 2, 8, 40, 48
    C = heads * d
    sl2 = math.log2(math.e) / math.sqrt(d)
    qf = rnd(N * S, heads, d) * amp
    kf = rnd(N * S, heads, d, seed=1) * amp
    if ramp:
        kf = kf * torch.linspace(0.2, 6.0, S, device=DEV).repeat(N).view(N * S, 1, 1)
    vf = rnd(N * S, heads, d, seed=2)
    q = torch.zeros(N * S, heads, R, device=DEV, dtype=torch.bfloat16)
    k = torch.zeros(N * S, heads, R, device=DEV, dtype=torch.bfloat16)
    q[:, :, :d] = (qf * sl2).bfloat16()
    k[:, :, :d] = kf.bfloat16()
    k[:, :, d] = 1.0
    vt = torch.zeros(heads, R, N, S, device=DEV, dtype=torch.bfloat16)
    vt[:, :d] = vf.bfloat16().view(N, S, heads, d).permute(2, 3, 0, 1)
    vt[:, d] = 1.0
    vt = vt.view(heads * R, N, S).contiguous()
    out = torch.empty(N * S, C, device=DEV, dtype=torch.bfloat16)
    ops.attention(q.view(N * S, heads * R), k.view(N * S, heads * R), vt, out, NB=N, heads=heads, d=d, S=S, Skv=S,
                  Skv_pad=S, ldq=heads * R, ldk=heads * R, ldo=C, sum_row=True, q_prescaled=True, qk_cols=R, qk_fold=True)
    qr = (q[:, :, :d].float() / sl2).view(N, S, heads, d).transpose(1, 2)
    kr = k[:, :, :d].float().view(N, S, heads, d).transpose(1, 2)
    vr = vf.bfloat16().float().view(N, S, heads, d).transpose(1, 2)
    ref = ref_attention(qr, kr, vr, False).transpose(1, 2).reshape(N * S, C)
    report(f"attention qk_fold S={S} amp={amp} ramp={ramp}", out, ref, 1e-2)
    # same inputs without the fold: the two must agree to bf16 rounding of P
    out2 = torch.empty_like(out)
    k2 = k.clone()
    k2[:, :, d] = 0.0
    ops.attention(q.view(N * S, heads * R), k2.view(N * S, heads * R), vt, out2, NB=N, heads=heads, d=d, S=S, Skv=S,
                  Skv_pad=S, ldq=heads * R, ldk=heads * R, ldo=C, sum_row=True, q_prescaled=True, qk_cols=R)
    report(f"attention qk_fold vs plain S={S}", out, out2, 8e-3)


# ------------------------------------------------------------------------------------ elementwise
def test_layout_and_upsample():
    ops = _ops()
    x = rnd(2, 4, 64, 64)
    out = ops.nchw_to_nhwc_bf16(x, repeat=2, scale=0.5)
    ref = (x * 0.5).repeat(2, 1, 1, 1).permute(0, 2, 3, 1)
    report("nchw->nhwc repeat", out, ref.bfloat16().float(), 1e-6)
    y = rnd(2, 16, 16, 64).bfloat16()
    report("nhwc->nchw", ops.nhwc_to_nchw_f32(y), y.float().permute(0, 3, 1, 2), 0.0)
    report("upsample2x", ops.upsample2x(y),
           F.interpolate(y.float().permute(0, 3, 1, 2), scale_factor=2, mode="nearest").permute(0, 2, 3, 1), 0.0)


def test_small_linear_time_path():
    ops = _ops()
    x = rnd(50, 320)
    w = rnd(1280, 320, scale=320 ** -0.5, seed=1).bfloat16()
    b = rnd(1280, seed=2)
    ref = F.silu(x @ w.float().t() + b)
    report("small_linear silu out", ops.small_linear(x, w, b, act_out=ops.ACT_SILU), ref, 1e-5)
    ref2 = F.silu(x[:, :320]) @ w.float().t() + b
    report("small_linear silu in", ops.small_linear(x, w, b, act_in=ops.ACT_SILU), ref2, 1e-5)


def test_cfg_ddpm_step():
    ops = _ops()
    B, C, H, W = 3, 4, 64, 64
    lat = rnd(B, C, H, W)
    eps = rnd(2 * B, H, W, C, seed=1)
    noise = rnd(B, C, H, W, seed=2)
    coef = torch.tensor([[0.99, 0.125, 0.3, 0.8, 0.45], [0.01, 0.999, 0.9, 0.1, 0.0]], device=DEV)
    e = eps.permute(0, 3, 1, 2)
    for step in (0, 1):
        sb, sa, c0, c1, sg = coef[step].tolist()
        mo = 7.5 * (e[:B] - e[B:]) + e[B:]
        x0 = (lat - sb * mo) / sa
        ref = c0 * x0 + c1 * lat + sg * noise
        got = lat.clone()
        nxt = torch.empty(2 * B, H, W, C, device=DEV, dtype=torch.bfloat16)
        ops.cfg_ddpm_step(got, eps, noise, coef, step, 7.5, True, nxt)
        report(f"cfg_ddpm_step {step}", got, ref, 1e-6)
        got32 = lat.clone()
        nxt32 = torch.empty(2 * B, H, W, C, device=DEV, dtype=torch.float32)
        ops.cfg_ddpm_step(got32, eps, noise, coef, step, 7.5, True, nxt32)
        report("next_in fp32", nxt32, ref.repeat(2, 1, 1, 1).permute(0, 2, 3, 1), 1e-6)
        report("next_in", nxt, ref.repeat(2, 1, 1, 1).permute(0, 2, 3, 1).bfloat16().float(), 1e-6)


def test_vae_scramble_tail_uint8_embed():
    ops = _ops()
    N, HW, C = 2, 1024, 512
    y = rnd(N, HW, C).bfloat16()
    r = rnd(N, HW, C, seed=1)
    # reference semantics (sd/decoder.py:62-71): raw re-view of (n, hw, c) as (n, c, h, w), NCHW add
    ref_nchw = y.float().reshape(N, C, HW) + r.permute(0, 2, 1)
    out, out_b = ops.vae_attn_scramble_add(y, r)
    report("vae scramble add", out, ref_nchw.permute(0, 2, 1), 1e-6)
    report("vae scramble add (bf16 shadow)", out_b, ref_nchw.permute(0, 2, 1), 8e-3)
    mom = rnd(2, 8, 8, 8) * 3
    nz = rnd(2, 4, 8, 8, seed=3)
    m = mom.permute(0, 3, 1, 2)
    ref = (m[:, :4] + torch.clamp(m[:, 4:], -30, 20).exp().sqrt() * nz) * 0.18215
    report("vae encode tail", ops.vae_encode_tail(mom, nz), ref, 1e-5)
    img = rnd(1, 16, 16, 3) * 1.2
    t = img.clone()
    t -= -1
    t *= 127.5
    ref8 = t.clamp(0, 255).to(torch.uint8)
    got8 = ops.image_to_uint8(img)
    assert torch.equal(got8, ref8)                      # byte work: bit-exact (same fp32 operation order, truncating cast)
    big = rnd(2, 64, 64, 3, seed=11) * 1.5
    tb = big.clone()
    tb -= -1
    tb *= 127.5
    assert torch.equal(ops.image_to_uint8(big), tb.clamp(0, 255).to(torch.uint8))
    u8 = torch.randint(0, 256, (1, 16, 16, 3), device=DEV, dtype=torch.uint8)
    f = u8.float()
    f *= 2 / 255
    f += -1
    report("uint8->image", ops.uint8_to_image(u8), f, 4e-3)
    assert torch.equal(ops.uint8_to_image(u8, out_fp32=True), f)
    allv = torch.arange(256, device=DEV, dtype=torch.uint8).view(1, 16, 16, 1).expand(1, 16, 16, 3).contiguous()
    fa = allv.float()
    fa *= 2 / 255
    fa += -1
    assert torch.equal(ops.uint8_to_image(allv, out_fp32=True), fa)
    tok = torch.randint(0, 1000, (2, 77), device=DEV)
    table = rnd(1000, 768)
    pos = rnd(77, 768, seed=9)
    e = ops.clip_embed(tok, table, pos, 80)
    report("clip embed", e[:, :77], table[tok] + pos, 1e-6)
    assert float(e[:, 77:].abs().max()) == 0.0
    a, b = rnd(1000), rnd(1000, seed=4)
    report("axpby", ops.axpby(a, b, 0.3, 0.7), 0.3 * a + 0.7 * b, 1e-6)


def test_matmul_f64_pack_time_composition():
    """ops.matmul_f64 (the GEGLU / conv_output fold at pack time) against torch's fp64 matmul."""
    ops = _ops()
    g = torch.Generator(device=DEV).manual_seed(5)
    for m, k, n, da, db in ((320, 1280, 320, torch.float32, torch.float32), (130, 77, 65, torch.float32, torch.float64),
                            (64, 2560, 1, torch.float32, torch.float32), (100, 33, 7, torch.float64, torch.float32)):
        a = torch.randn(m, k, device=DEV, generator=g, dtype=torch.float32).to(da)
        b = torch.randn(k, n, device=DEV, generator=g, dtype=torch.float32).to(db)
        got = ops.matmul_f64(a, b)
        ref = a.double() @ b.double()
        assert got.dtype == torch.float64 and got.shape == (m, n)
        assert float((got - ref).abs().max() / ref.abs().max()) < 1e-13


@pytest.mark.parametrize("N,H,W,C0,C1,Cout,block_n,pair", [
    (1, 16, 16, 320, 0, 160, 0, 0),       # two row tiles -> one CTA pair
    (1, 16, 16, 320, 0, 160, 0, 1),       # the same through single-CTA tiles
    (1, 8, 16, 512, 0, 128, 0, 0),        # ONE row tile (bw = 16, bh = 8): cta_group::1
    (2, 32, 32, 320, 0, 320, 160, 0),
    (2, 32, 32, 320, 0, 320, 320, 0),     # wide tile: two accumulators share the halo box
    (1, 48, 48, 256, 64, 320, 0, 0),      # ragged: 48 = 3 patches of 16 columns, dual source
    (1, 96, 32, 320, 0, 64, 0, 0),        # H != W
    (3, 16, 16, 640, 640, 640, 0, 0),     # odd batch, dual source
    (1, 64, 64, 384, 0, 96, 0, 0),        # block_n = 96
])
def test_conv3x3_filter_column_staging(N, H, W, C0, C1, Cout, block_n, pair):
    """Filter-column staging (GemmTcParams::a3): one halo box per filter column, the three taps of the column read it
    as row-shifted views. Eligible shapes (patch of one sample, > 32 k-blocks, no split-K) against F.conv2d, with a
    residual, the bf16 copy and the GroupNorm partial sums in the same launch."""
    ops = _ops()
    setup_exact_fp32()
    from pytorch_stable_diffusion_b200 import _ext
    assert _ext.lib().sdb_gemm_conv_a3_bytes(N, H, W) > 0 and 9 * (C0 + C1) // 64 > 32
    x0 = rnd(N, H, W, C0).bfloat16()
    x1 = rnd(N, H, W, C1, seed=5).bfloat16() if C1 else None
    w = rnd(Cout, C0 + C1, 3, 3, scale=(9 * (C0 + C1)) ** -0.5, seed=1).bfloat16()
    b = rnd(Cout, seed=2)
    r = rnd(N * H * W, Cout, seed=3)
    xin = x0.float() if x1 is None else torch.cat([x0.float(), x1.float()], -1)
    ref = F.conv2d(xin.permute(0, 3, 1, 2), w.float(), b, padding=1).permute(0, 2, 3, 1).reshape(-1, Cout) + r
    out, out2, part = ops.gemm(x0, pack3x3(w), Cout, kind=ops.GEMM_CONV3X3_S1, a1=x1, bias=b, residual=r,
                               conv_dims=(N, H, W), c0=C0, c1=C1, out_fp32=True, out2=True, block_n=block_n, nsplit=1,
                               cta_pair=pair, gn_samples=N)
    report(f"conv3x3 filter-column {C0 + C1}->{Cout}@{H}x{W} bn={block_n} pair={pair}", out, ref, 3e-3)
    assert torch.equal(out2, out.bfloat16())
    if part is not None:
        _check_partials(part, out, N, "conv3x3 filter-column")


# ------------------------------------------------------------------------------------ IEEE-half operands
@pytest.mark.parametrize("M,K,N,nsplit", [(4096, 320, 320, 1), (8192, 320, 640, 1), (1024, 2560, 320, 4), (300, 128, 96, 1)])
def test_linear_f16_operands(M, K, N, nsplit):
    """sdb_gemm_args::ab_f16: A and W in IEEE half (the UNet's full-resolution level), fp32 accumulation, fp32 result
    + half shadow, fp32 residual - through the TMA epilogue (short K), the per-lane one and split-K."""
    ops = _ops()
    setup_exact_fp32()
    a = rnd(M, K).half()
    w = rnd(N, K, scale=K ** -0.5, seed=1).half()
    b = rnd(N, seed=2)
    r = rnd(M, N, seed=3)
    ref = a.float() @ w.float().t() + b + r
    out, out2 = ops.linear(a, w, bias=b, residual=r, out_fp32=True, out2=True, out16=torch.float16, nsplit=nsplit)
    report(f"linear f16 operands {M}x{K}x{N} split={nsplit}", out, ref, 1e-3)
    assert out2.dtype == torch.float16 and torch.equal(out2, out.half())
    o16 = ops.linear(a, w, bias=b, out16=torch.float16, nsplit=1)
    assert o16.dtype == torch.float16
    report("linear f16 operands, half out", o16, a.float() @ w.float().t() + b, 1e-3)
    with pytest.raises(ValueError):
        ops.linear(a, w.bfloat16(), bias=b)                 # mixed operand types are refused


@pytest.mark.parametrize("N,H,C0,C1,Cout,cx", [(2, 32, 320, 0, 320, 0), (1, 64, 320, 320, 320, 0), (2, 16, 320, 0, 320, 640),
                                               (2, 8, 128, 0, 64, 0)])
def test_conv3x3_f16_operands(N, H, C0, C1, Cout, cx):
    ops = _ops()
    setup_exact_fp32()
    x0 = rnd(N, H, H, C0).half()
    x1 = rnd(N, H, H, C1, seed=5).half() if C1 else None
    w = rnd(Cout, C0 + C1, 3, 3, scale=(9 * (C0 + C1)) ** -0.5, seed=1).half()
    b = rnd(Cout, seed=2)
    xin = x0.float() if x1 is None else torch.cat([x0.float(), x1.float()], -1)
    ref = F.conv2d(xin.permute(0, 3, 1, 2), w.float(), b, padding=1).permute(0, 2, 3, 1).reshape(-1, Cout)
    wp = pack3x3(w)
    ax0 = None
    if cx:
        ax0 = rnd(N, H, H, cx, seed=6).half()
        w1 = rnd(Cout, cx, scale=cx ** -0.5, seed=8).half()
        ref = ref + F.conv2d(ax0.float().permute(0, 3, 1, 2), w1.float()[:, :, None, None]).permute(0, 2, 3, 1).reshape(-1, Cout)
        wp = torch.cat([wp, w1], dim=1).contiguous()
    out, out2, part = ops.gemm(x0, wp, Cout, kind=ops.GEMM_CONV3X3_S1, a1=x1, bias=b, conv_dims=(N, H, H), c0=C0, c1=C1,
                               out_fp32=True, out2=True, out16=torch.float16, nsplit=1, ax0=ax0, gn_samples=N)
    report(f"conv3x3 f16 operands {C0 + C1}(+{cx})->{Cout}@{H}", out, ref, 1e-3)
    assert out2.dtype == torch.float16 and torch.equal(out2, out.half())
    if part is not None:
        _check_partials(part, out, N, "conv3x3 f16")


def test_half_outputs_of_norms_and_shadows():
    ops = _ops()
    x = rnd(2, 32, 32, 320) * 3 + 0.5
    g, be = rnd(320, seed=1), rnd(320, seed=2)
    ref = F.silu(F.group_norm(x.permute(0, 3, 1, 2), 32, g, be, 1e-5)).permute(0, 2, 3, 1)
    for fused in (False, None):
        y = ops.groupnorm(x, g, be, silu=True, fused=fused, out_dtype=torch.float16)
        assert y.dtype == torch.float16
        report(f"groupnorm half out fused={fused}", y, ref, 1.5e-3)
    xh = x.half()                                                              # IEEE-half INPUT (the resblock's hidden tensor)
    refh = F.silu(F.group_norm(xh.float().permute(0, 3, 1, 2), 32, g, be, 1e-5)).permute(0, 2, 3, 1)
    report("groupnorm half in (stats + apply)", ops.groupnorm(xh, g, be, silu=True, fused=False), refh, 6e-3)
    xl = rnd(4096, 320, seed=4)
    y = ops.layernorm(xl, g, be, out_dtype=torch.float16)
    assert y.dtype == torch.float16
    report("layernorm half out", y, F.layer_norm(xl, (320,), g, be, 1e-5), 1e-3)
    y = ops.layernorm(xl.bfloat16(), g, be, out_dtype=torch.float16)           # generic kernel
    report("layernorm (bf16 in) half out", y, F.layer_norm(xl.bfloat16().float(), (320,), g, be, 1e-5), 1e-3)
    big = rnd(1000, 64, seed=7) * 100
    big[0, 0], big[0, 1] = 1e6, -1e6                                             # saturates, never inf
    h = ops.f32_to_bf16(big, torch.float16)
    assert h.dtype == torch.float16 and torch.isfinite(h).all()
    assert torch.equal(h, big.clamp(-65504, 65504).half())
    assert torch.equal(ops.f32_to_bf16(big), big.bfloat16())
    odd = rnd(777, seed=8)
    assert torch.equal(ops.f32_to_bf16(odd, torch.float16), odd.half())
    xs = rnd(2, 16, 16, 4, seed=9)
    wd = rnd(320, 9, 4, scale=1 / 6, seed=10)
    bd = rnd(320, seed=11)
    o, o2 = ops.conv_direct(xs, wd, bd, 320, 3, out_fp32=True, out2=True, out2_dtype=torch.float16)
    assert o2.dtype == torch.float16 and torch.equal(o2, o.half())
    up = ops.upsample2x(o2)
    assert up.dtype == torch.float16 and torch.equal(up[:, ::2, ::2], o2)


@pytest.mark.parametrize("size", [(768, 768), (256, 256), (384, 640), (500, 300), (512, 512), (1024, 520)])
def test_resize_matches_pillow_byte_for_byte(size):
    """Device resize + rescale against PIL.Image.resize on the reference's own test image
    (images/dog.jpg, stored in the img2img golden): identical bytes, identical fp32 pre-processing."""
    from PIL import Image
    from canon import golden
    from pytorch_stable_diffusion_b200 import imageio
    g = golden("img2img_5.pt")
    if g is None:
        pytest.skip("img2img golden (holds dog.jpg) not generated")
    dog = Image.fromarray(g["input"].numpy())
    w, h = size
    ref = torch.from_numpy(np.array(dog.resize((w, h)))).to(DEV)
    batch = torch.from_numpy(np.stack([np.asarray(dog), np.asarray(dog)[::-1].copy()])).to(DEV)
    u8, img = imageio.resize_u8(batch, w, h, to_image=True)
    assert u8.shape == (2, h, w, 3) and torch.equal(u8[0], ref)
    flipped = torch.from_numpy(np.array(Image.fromarray(np.asarray(dog)[::-1].copy()).resize((w, h)))).to(DEV)
    assert torch.equal(u8[1], flipped)
    f = ref.float()
    f *= 2 / 255
    f += -1
    assert torch.equal(img[0], f)
    u8b, imgb = imageio.load_images([dog, dog.resize((300, 200))], w, h, DEV)
    assert torch.equal(u8b[0], ref) and u8b.shape == (2, h, w, 3) and imgb.shape == (2, h, w, 3)


@pytest.mark.parametrize("N,H,W,Cin,Cout,dt", [(2, 8, 8, 1280, 1280, torch.bfloat16), (2, 16, 16, 1280, 1280, torch.bfloat16),
                                               (1, 32, 32, 640, 640, torch.bfloat16), (1, 24, 40, 128, 64, torch.bfloat16),
                                               (3, 12, 12, 256, 160, torch.float16), (1, 48, 48, 640, 640, torch.float16)])
def test_conv_up2x_phases(N, H, W, Cin, Cout, dt):
    """Upsample (nearest x2 -> conv3x3, sd/diffusion.py:412-435) as four parity-phase 2x2 convolutions of the
    low-resolution input (SDB_GEMM_CONV2X2_UP, ops.conv_up2x) against F.conv2d(F.interpolate(x))."""
    ops = _ops()
    setup_exact_fp32()
    from pytorch_stable_diffusion_b200 import engine
    conv = torch.nn.Conv2d(Cin, Cout, 3, padding=1).to(DEV)
    with torch.no_grad():
        conv.weight.copy_(rnd(Cout, Cin, 3, 3, scale=(9 * Cin) ** -0.5, seed=1))
        conv.bias.copy_(rnd(Cout, seed=2))
        x = rnd(N, H, W, Cin).to(dt)
        ref = conv(F.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=2, mode="nearest")).permute(0, 2, 3, 1)
        w4, b = engine.pack_upsample_phases(conv, DEV, dt)
        out, out2, part = ops.conv_up2x(x, w4, Cout, bias=b, out2=True, gn_samples=N, out16=dt)
    assert out.shape == (N, 2 * H, 2 * W, Cout)
    report(f"conv_up2x {Cin}->{Cout} @{H}x{W} {dt}", out, ref, 1e-2 if dt == torch.bfloat16 else 2e-3)
    assert out2.dtype == dt and torch.equal(out2, out.to(dt))
    if part is not None:
        _check_partials(part, out.reshape(-1, Cout), N, "conv_up2x")
    # the single launch over all four phases (up_phase = 4) equals four per-phase launches bit for bit
    sep = torch.empty_like(out)
    sep2 = torch.empty_like(out2)
    part_sep = torch.zeros_like(part) if part is not None else None
    for phase in range(4):
        ops.gemm(x, w4[phase], Cout, kind=ops.GEMM_CONV2X2_UP, bias=b, conv_dims=(N, H, W), c0=Cin, out=sep,
                 out_fp32=True, out2=sep2, out16=dt, up_phase=phase, gn_part=part_sep)
    assert torch.equal(out, sep) and torch.equal(out2, sep2)
    if part is not None:
        assert torch.equal(part, part_sep)


def test_c_host_program(tmp_path):
    """The boundary from the other side: examples/abi_linear.c (C99, links libsdb200.so + cudart only, no Python in the
    process) runs one nn.Linear through sdb_gemm_tc and checks it against a double-precision host loop."""
    import subprocess
    from test_host_cpu import _build_c_example
    exe = _build_c_example(tmp_path)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    print(r.stdout.strip(), r.stderr.strip(), flush=True)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)

