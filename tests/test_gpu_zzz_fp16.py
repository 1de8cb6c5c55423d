"""Kernels of the "fp16" forward mode against torch: the hi/lo -> fp16 staging, the BatchNorm apply writing (bf16, fp16)
copies, and the tcgen05 GEMM with the fp16 instruction descriptor (operands interpreted as fp16: a bf16 interpretation of
the same bits would be off by orders of magnitude, so the tolerances below — fp16 operand rounding is exact here, only the
fp32 accumulation order differs — also prove the descriptor change took effect)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_pair_to_f16_and_weight_staging():
    from gan_playground_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(0)
    v = torch.randn(300, 136, device="cuda", generator=g)
    hi = v.bfloat16()
    lo = (v - hi.float()).bfloat16()
    buf = torch.zeros(300, 2 * 136, device="cuda", dtype=torch.bfloat16)       # hi | lo per row, like a packed weight row
    buf[:, :136], buf[:, 136:] = hi, lo
    out = ops.pair_to_f16(buf, buf[:, 136:], 300, 136, 2 * 136)
    want = (hi.float() + lo.float()).half()
    assert out.dtype == torch.float16 and torch.equal(out, want)
    assert (out.float() - v).abs().max() <= 1.01 * (v.abs().max() * 2 ** -11)
    w = torch.randn(64, 32, 4, 4, device="cuda", generator=g) * 0.02
    for n_dim in (0, 1):
        wh = ops.conv_weight_f16(w, n_dim)
        wb = ops.pack_conv_weight(w, n_dim)
        assert wh.shape == wb.shape and wh.dtype == torch.float16
        assert (wh.float() - wb.float()).abs().max() <= 2 ** -8 * w.abs().max()       # same layout, finer rounding
        assert (wh.float() - wb.float()).abs().max() > 0


def test_bn_apply_act_pair():
    from gan_playground_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(1)
    y = torch.randn(4, 8, 8, 64, device="cuda", generator=g) * 3 + 1
    fin = torch.zeros(4, 64, device="cuda")
    fin[2] = torch.rand(64, device="cuda", generator=g) + 0.5      # scale
    fin[3] = torch.randn(64, device="cuda", generator=g)           # shift
    for act in (ops.ACT_RELU, ops.ACT_LRELU):
        b, h = ops.bn_apply_act_pair(y, fin, act)
        z = y * fin[2] + fin[3]
        want = torch.relu(z) if act == ops.ACT_RELU else torch.nn.functional.leaky_relu(z, 0.2)
        assert torch.equal(b, want.bfloat16()) or (b.float() - want).abs().max() <= 2 ** -8 * want.abs().max()
        assert (h.float() - want).abs().max() <= 2 ** -10 * want.abs().max()
        hi, lo = ops.bn_apply_act_split(y, fin, act)
        assert torch.equal(hi, b)                                  # the bf16 copy is the split mode's hi half


@pytest.mark.parametrize("M,N,K", [(1000, 256, 512), (77, 24, 40), (40000, 128, 256)])
def test_fp16_linear_gemm(M, N, K):
    from gan_playground_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(2)
    x = torch.randn(M, K, device="cuda", generator=g).half()
    w = (torch.randn(N, K, device="cuda", generator=g) * 0.05).half()
    y = ops.conv_fwd(x.view(M, 1, 1, K), w, None, ops.KIND_CONV_K1S1, 1, 1, fp16_in=True, out_mode="f32").view(M, N)
    want = x.double() @ w.double().t()
    err = ((y.double() - want).abs().max() / want.abs().max()).item()
    print("fp16 GEMM %dx%dx%d: max rel err %.2e" % (M, N, K, err))
    assert err < 1e-4                                              # exact operands, fp32 accumulation
    wrong = x.view(torch.bfloat16).double() @ w.view(torch.bfloat16).double().t()   # the bits read as bf16
    assert not ((wrong - want).abs().max() / want.abs().max()).item() < 1e-1       # (NaN / inf also count as "not close")


def test_fp16_conv_pair_output_and_stats():
    """k4s2 conv on fp16 operands with LeakyReLU and the (bf16, fp16) output pair; ConvT with fused statistics + fp32 out."""
    from gan_playground_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(3)
    NB, H, Cin, Cout = 8, 16, 64, 128
    x = torch.randn(NB, H, H, Cin, device="cuda", generator=g).half()
    w = torch.randn(Cout, Cin, 4, 4, device="cuda", generator=g) * 0.05
    bias = torch.randn(Cout, device="cuda", generator=g)
    wh = ops.conv_weight_f16(w, 0)
    a, ah = ops.conv_fwd(x, wh, bias, ops.KIND_CONV_K4S2, H // 2, H // 2, ops.ACT_LRELU, fp16_in=True, out_mode="pair")
    w16 = wh.float().view(Cout, 4, 4, Cin).permute(0, 3, 1, 2)
    want = torch.nn.functional.leaky_relu(
        torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), w16, bias, stride=2, padding=1), 0.2).permute(0, 2, 3, 1)
    s = want.abs().max()
    assert a.dtype == torch.bfloat16 and ah.dtype == torch.float16
    assert (ah.float() - want).abs().max() <= 2 ** -10 * s and (a.float() - want).abs().max() <= 2 ** -7 * s
    wt = torch.randn(Cin, Cout, 4, 4, device="cuda", generator=g) * 0.05
    wth = ops.conv_weight_f16(wt, 1)
    st = torch.zeros(2, Cout, device="cuda")
    y = ops.conv_fwd(x, wth, None, ops.KIND_CONVT_K4S2, 2 * H, 2 * H, stats=st, fp16_in=True, out_mode="f32")
    wt16 = wth.float().view(Cout, 4, 4, Cin).permute(3, 0, 1, 2)
    want = torch.nn.functional.conv_transpose2d(x.float().permute(0, 3, 1, 2), wt16, None, stride=2, padding=1).permute(0, 2, 3, 1)
    assert (y - want).abs().max() <= 1e-4 * want.abs().max()
    assert torch.allclose(st[0], want.sum(dim=(0, 1, 2)), rtol=1e-3, atol=1e-2)
    assert torch.allclose(st[1], (want * want).sum(dim=(0, 1, 2)), rtol=1e-3, atol=1e-2)
