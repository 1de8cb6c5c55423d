"""Parity of the tcgen05 implicit-GEMM kernels (through the C-ABI) against torch fp32 convolutions on bf16-rounded
inputs, plus size-independent properties at BASELINE sizes. Tolerance: outputs are stored in bf16, so the bound is
2^-8 relative to the largest output (2e-2 used; observed 2-4e-3); fp32 wgrad results 5e-3 (observed < 1e-5)."""
import os
import sys

import pytest
import torch

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "tools"))
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def st():
    import selftest_conv

    return selftest_conv


@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (128, 128, 256), (1000, 256, 512), (77, 24, 40), (4096, 512, 1024),
                                   (40000, 128, 256), (50001, 64, 128)])   # last two: 256-row (MT=2) tiles
def test_linear_gemm(st, M, N, K):
    assert st.case_fwd_k1(M, N, K)


def test_linear_gemm_structured_operands_catch_permutations(st):
    assert st.case_fwd_k1(128, 64, 64, structured=True)


@pytest.mark.parametrize("kind,NB,H,Cin,Cout,act,stats", [
    ("k4s2", 8, 8, 64, 64, 0, False), ("k4s2", 32, 32, 128, 256, 2, False), ("k4s2", 20, 8, 512, 1024, 0, True),
    ("k4s2", 3, 16, 16, 8, 2, False),          # ragged batch, channels < one swizzle row
    ("convt", 8, 4, 64, 64, 0, False), ("convt", 24, 8, 512, 256, 1, False), ("convt", 8, 16, 256, 128, 0, True),
    ("convt", 5, 4, 32, 16, 1, False),
    ("k3s1", 8, 8, 128, 128, 0, False), ("k3s1", 6, 4, 64, 256, 1, False),
    ("k4s2", 256, 32, 64, 128, 2, True), ("convt", 128, 16, 256, 128, 1, True), ("k3s1", 160, 16, 64, 64, 0, False),  # MT=2
])
def test_conv_forward_kinds(st, kind, NB, H, Cin, Cout, act, stats):
    assert st.case_fwd_conv(kind, NB, H, Cin, Cout, act=act, stats=stats)


@pytest.mark.parametrize("kind,NB,Hs,Cd,Cg", [
    ("k1s1", 64, 1, 128, 64), ("k1s1", 512, 2, 256, 256), ("k4s2", 8, 4, 128, 64), ("k4s2", 64, 8, 512, 256),
    ("k4s2", 16, 16, 256, 128), ("k3s1", 16, 8, 128, 128), ("k4s2", 7, 4, 16, 8),
])
def test_conv_wgrad_kinds(st, kind, NB, Hs, Cd, Cg):
    assert st.case_wgrad(kind, NB, Hs, Cd, Cg)


def test_full_size_dcgan64_layers_linearity_and_adjoint():
    """BASELINE-size property checks (no CPU oracle at this size): conv is linear in its input, and
    <dy, conv(x, w)> == <wgrad(dy, x), w> (adjoint identity) for D block 2 (256 -> 512, 16x16 -> 8x8) at batch 1024."""
    from gan_playground_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(0)
    NB, H, Cin, Cout = 1024, 16, 256, 512
    x1 = torch.randn(NB, H, H, Cin, device="cuda", generator=g).bfloat16()
    x2 = torch.randn(NB, H, H, Cin, device="cuda", generator=g).bfloat16()
    w = (torch.randn(Cout, Cin, 4, 4, device="cuda", generator=g) * 0.05)
    wp = ops.pack_conv_weight(w, 0)
    y1 = ops.conv_fwd(x1, wp, None, ops.KIND_CONV_K4S2, H // 2, H // 2).float()
    y2 = ops.conv_fwd(x2, wp, None, ops.KIND_CONV_K4S2, H // 2, H // 2).float()
    xs = (x1.float() + x2.float()).bfloat16()          # exact sum may round; compare against conv of the rounded sum
    ys = ops.conv_fwd(xs, wp, None, ops.KIND_CONV_K4S2, H // 2, H // 2).float()
    lin_err = ((y1 + y2) - ys).abs().max() / ys.abs().max()
    assert lin_err < 3e-2, lin_err
    dy = torch.randn(NB, H // 2, H // 2, Cout, device="cuda", generator=g).bfloat16()
    dw = ops.conv_wgrad(dy, x1, ops.KIND_CONV_K4S2, 16)                     # [Cout][16][Cin] fp32
    w_bf = wp.float().view(Cout, 16, Cin)
    lhs = (dy.float() * y1).sum().double()
    rhs = (dw.double() * w_bf.double()).sum()
    assert abs(lhs - rhs) / abs(rhs) < 5e-3, (lhs.item(), rhs.item())
    # dgrad adjoint: <dy, conv(x)> == <convT(dy), x>
    wpd = ops.pack_conv_weight(w, 1)
    dx = ops.conv_fwd(dy, wpd, None, ops.KIND_CONVT_K4S2, H, H).float()
    rhs2 = (dx.double() * x1.double()).sum()
    assert abs(lhs - rhs2) / abs(rhs2) < 5e-3, (lhs.item(), rhs2.item())


# ------------------------------------------------------------------------------ backward-side epilogue fusion (BwdLink)
_DGRAD_CASES = [
    # (kind, NB, Hin, Cin, Nout): the data-gradient GEMM of the NEXT layer; its output (NB, Hout, Wout, Nout) is dA of the
    # producing layer
    ("convt", 320, 4, 512, 256),     # CTA-pair kernel (>= 74 pair tiles)
    ("convt", 24, 8, 256, 128),      # 128-wide tiles, two 128-row sub-tiles
    ("k4s2", 40, 16, 128, 256),      # dgrad of a ConvTranspose2d block
    ("k1", 12, 32, 64, 64),          # image-side layer: 64-wide tiles, ragged last row tile
]


def _dgrad_case(kind, NB, Hin, Cin, Nout):
    from gan_playground_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(1)
    dy = (torch.randn(NB, Hin, Hin, Cin, device="cuda", generator=g) * 0.5).bfloat16()
    if kind == "convt":
        Hout, k, okind = 2 * Hin, ops.KIND_CONVT_K4S2, 16
    elif kind == "k4s2":
        Hout, k, okind = Hin // 2, ops.KIND_CONV_K4S2, 16
    else:
        Hout, k, okind = Hin, ops.KIND_CONV_K1S1, 1
    wp = (torch.randn(Nout, okind * Cin, device="cuda", generator=g) * 0.05).bfloat16()
    shape = (NB, Hout, Hout, Nout)
    return ops, dy, wp, k, Hout, shape, g


@pytest.mark.parametrize("kind,NB,Hin,Cin,Nout", _DGRAD_CASES)
@pytest.mark.parametrize("act", ["relu", "lrelu"])
def test_dgrad_epilogue_applies_the_producers_activation_derivative(kind, NB, Hin, Cin, Nout, act):
    ops, dy, wp, k, Hout, shape, g = _dgrad_case(kind, NB, Hin, Cin, Nout)
    a = torch.randn(shape, device="cuda", generator=g).bfloat16()
    a[0, 0, 0, :8] = 0                                             # the derivative at exactly 0 is the slope
    code = ops.ACT_RELU if act == "relu" else ops.ACT_LRELU
    plain = ops.conv_fwd(dy, wp, None, k, Hout, Hout)
    want = ops.act_bwd(plain, a, code).float()                     # the standalone pass it replaces
    got = ops.conv_fwd(dy, wp, None, k, Hout, Hout, bwd=("mask", a, code)).float()
    torch.cuda.synchronize()
    # one extra bf16 rounding in the two-pass reference (v -> bf16 -> * slope -> bf16)
    assert torch.allclose(got, want, rtol=2 ** -7, atol=1e-6), (got - want).abs().max().item()
    pos = a.float() > 0
    assert torch.equal(got[pos], plain.float()[pos])               # untouched where the activation was active


@pytest.mark.parametrize("kind,NB,Hin,Cin,Nout", _DGRAD_CASES)
@pytest.mark.parametrize("ydtype", [torch.float32, torch.bfloat16])
def test_dgrad_epilogue_accumulates_the_producers_batchnorm_backward_sums(kind, NB, Hin, Cin, Nout, ydtype):
    ops, dy, wp, k, Hout, shape, g = _dgrad_case(kind, NB, Hin, Cin, Nout)
    y = (torch.randn(shape, device="cuda", generator=g) * 1.5 + 0.3).to(ydtype)
    gamma = torch.rand(Nout, device="cuda", generator=g) + 0.5
    beta = torch.randn(Nout, device="cuda", generator=g) * 0.2
    st = ops.bn_stats_f32(y) if ydtype == torch.float32 else ops.bn_stats(y)
    fin = ops.bn_finalize(st, y.numel() // Nout, gamma, beta, None, None, None)
    plain = ops.conv_fwd(dy, wp, None, k, Hout, Hout)
    reduce = ops.bn_bwd_reduce_f32 if ydtype == torch.float32 else ops.bn_bwd_reduce
    want = reduce(plain, y, fin, ops.ACT_LRELU)                    # the standalone pass it replaces
    red = torch.zeros(2, Nout, device="cuda")
    got = ops.conv_fwd(dy, wp, None, k, Hout, Hout, bwd=("bn", y, fin, ops.ACT_LRELU, red))
    torch.cuda.synchronize()
    assert torch.equal(got, plain)                                 # dA itself is unchanged
    scale = want.abs().max(dim=1, keepdim=True).values
    assert ((red - want).abs() / scale).max().item() < 2e-4, ((red - want).abs() / scale).max().item()
