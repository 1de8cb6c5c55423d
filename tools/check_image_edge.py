"""Fused image-edge kernels (csrc/image_edge.cu) against torch in true fp32 and against the column-buffer path they
replace, then timed at the benched size with CUDA events. Run on a B200:  python tools/check_image_edge.py [--time]"""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gan_playground_b200 import ops  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
FAILED = []
ONLY = None
for _i, _a in enumerate(sys.argv):
    if _a == "--only":
        ONLY = sys.argv[_i + 1]      # fwd | wgrad | convt: one kernel family per process (a faulting kernel poisons the context)


def want(what):
    return ONLY is None or ONLY == what


def rel(a, b):
    return ((a.float() - b.float()).abs().max() / (b.float().abs().max() + 1e-30)).item()


def report(name, err, tol):
    ok = err <= tol and err == err
    print("%-66s max-rel-err %.3e (tol %.0e) %s" % (name, err, tol, "ok" if ok else "FAILED"))
    if not ok:
        FAILED.append(name)


def nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


def case(NB, res, C, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    img = (torch.rand(NB, 3, res, res, generator=g) * 2 - 1).to(dev)
    w = (torch.randn(C, 3, 4, 4, generator=g) * 0.05).to(dev)
    b = (torch.randn(C, generator=g) * 0.1).to(dev)
    ref = F.leaky_relu(F.conv2d(img, w, b, stride=2, padding=1), 0.2)
    tag = "B%d r%d C%d" % (NB, res, C)
    for fmt, name, tol in ((ops.COMP_NONE, "bf16", 1.5e-2), (ops.COMP_LO, "bf16x3", 2e-4), (ops.COMP_F16, "fp16", 2e-3)):
        if not want("fwd"):
            break
        out, comp = ops.image_conv_fwd(img, w, b, ops.ACT_LRELU, fmt)
        val = out.float() + comp.float() if fmt == ops.COMP_LO else (comp.float() if fmt == ops.COMP_F16 else out.float())
        report("image_conv_fwd %s %s" % (name, tag), rel(val, nhwc(ref)), tol)
        report("image_conv_fwd %s %s (bf16 tensor)" % (name, tag), rel(out, nhwc(ref)), 1.5e-2)
    # data-gradient use: mul = tanh output, no bias, no activation
    t = torch.tanh(torch.randn(NB, 3, res, res, generator=g)).to(dev)
    if want("fwd"):
        ref = F.conv2d(img * (1 - t * t), w, None, stride=2, padding=1)
        out, _ = ops.image_conv_fwd(img, w, None, ops.ACT_NONE, ops.COMP_NONE, mul=t)
        report("image_conv_fwd mul=tanh' %s" % tag, rel(out, nhwc(ref)), 1.5e-2)
    if not want("wgrad"):
        return
    # weight gradient: dense = bf16 dy
    dy = (torch.randn(NB, res // 2, res // 2, C, generator=g) * 0.1).to(dev).to(torch.bfloat16)
    for mul in (None, t):
        x = img if mul is None else img * (1 - mul * mul)
        cols = F.unfold(x.to(torch.bfloat16).float(), 4, stride=2, padding=1)          # (NB, 48, P)
        ref_dw = torch.einsum("npm,njp->mj", dy.float().view(NB, -1, C), cols).view(C, 3, 4, 4)
        dw = torch.zeros(C, 3, 4, 4, device=dev)
        db = torch.zeros(C, device=dev)
        ops.image_conv_wgrad(dy, img, mul, dw, db)
        report("image_conv_wgrad%s %s" % (" mul" if mul is not None else "", tag), rel(dw, ref_dw), 2e-3)
        report("image_conv_wgrad dbias%s %s" % (" mul" if mul is not None else "", tag), rel(db, dy.float().sum((0, 1, 2))), 2e-3)
    # accumulation into an existing buffer
    dw2 = dw.clone()
    ops.image_conv_wgrad(dy, img, t, dw2, None)
    report("image_conv_wgrad accumulates %s" % tag, rel(dw2, 2 * dw), 1e-5)


def case_t(NB, res, C, seed):
    """transposed direction: x NHWC (NB, res/2, res/2, C) -> image (NB, 3, res, res)"""
    g = torch.Generator(device="cpu").manual_seed(seed)
    Hs = res // 2
    x = (torch.randn(NB, C, Hs, Hs, generator=g)).to(dev)
    w = (torch.randn(C, 3, 4, 4, generator=g) * 0.05).to(dev)
    b = (torch.randn(3, generator=g) * 0.1).to(dev)
    ref = torch.tanh(F.conv_transpose2d(x, w, b, stride=2, padding=1))
    tag = "B%d r%d C%d" % (NB, res, C)
    if not want("convt"):
        return
    xn = nhwc(x)
    hi = xn.to(torch.bfloat16)
    lo = (xn - hi.float()).to(torch.bfloat16)
    report("image_convt_fwd bf16 %s" % tag, rel(ops.image_convt_fwd(hi, None, w, b, 3, ops.ACT_TANH), ref), 1.5e-2)
    report("image_convt_fwd bf16x3 %s" % tag, rel(ops.image_convt_fwd(hi, lo, w, b, 3, ops.ACT_TANH), ref), 2e-4)
    report("image_convt_fwd fp16 %s" % tag, rel(ops.image_convt_fwd(xn.to(torch.float16), None, w, b, 3, ops.ACT_TANH), ref), 2e-3)
    ref = F.conv_transpose2d(x, w, None, stride=2, padding=1)
    report("image_convt_fwd no bias / act %s" % tag, rel(ops.image_convt_fwd(hi, None, w, None, 3, ops.ACT_NONE), ref), 1.5e-2)


def timeit(fn, reps=20):
    if "--time-only" in sys.argv:
        reps = 1
    for _ in range(0 if "--time-only" in sys.argv else 3):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


def timing(NB=1024, res=64, C=128):
    """The benched shape: DCGAN-64's D block 0 is 3 -> 128 at 32x32 output, G's last layer 128 -> 3."""
    img = torch.rand(NB, 3, res, res, device=dev) * 2 - 1
    t = torch.tanh(torch.randn(NB, 3, res, res, device=dev))
    w = torch.randn(C, 3, 4, 4, device=dev) * 0.05
    b = torch.zeros(C, device=dev)
    b3 = torch.zeros(3, device=dev)
    P = NB * (res // 2) ** 2
    img_b = img.numel() * 4
    act_b = P * C * 2
    print("\ntiming at B=%d res=%d C=%d (us per launch; GB/s of algorithmic bytes; 6551 GB/s measured peak)" % (NB, res, C))
    for fmt, name in ((ops.COMP_NONE, "bf16"), (ops.COMP_LO, "bf16x3"), (ops.COMP_F16, "fp16")):
        us = timeit(lambda: ops.image_conv_fwd(img, w, b, ops.ACT_LRELU, fmt))
        by = img_b + act_b * (1 if fmt == ops.COMP_NONE else 2)
        print("  image_conv_fwd %-7s %7.1f us  %6.0f GB/s" % (name, us, by / us / 1e3))
    us = timeit(lambda: ops.image_conv_fwd(img, w, None, ops.ACT_NONE, ops.COMP_NONE, mul=t))
    print("  image_conv_fwd mul     %7.1f us  %6.0f GB/s" % (us, (2 * img_b + act_b) / us / 1e3))
    dy = torch.randn(NB, res // 2, res // 2, C, device=dev).to(torch.bfloat16)
    dw = torch.zeros(C, 3, 4, 4, device=dev)
    db = torch.zeros(C, device=dev)
    us = timeit(lambda: ops.image_conv_wgrad(dy, img, None, dw, db))
    print("  image_conv_wgrad       %7.1f us  %6.0f GB/s" % (us, (img_b + act_b) / us / 1e3))
    us = timeit(lambda: ops.image_conv_wgrad(dy, img, t, dw, None))
    print("  image_conv_wgrad mul   %7.1f us  %6.0f GB/s" % (us, (2 * img_b + act_b) / us / 1e3))
    lo = torch.zeros_like(dy)
    for xa, xl, name in ((dy, None, "bf16"), (dy, lo, "bf16x3"), (dy.to(torch.float16), None, "fp16")):
        us = timeit(lambda: ops.image_convt_fwd(xa, xl, w, b3, 3, ops.ACT_TANH))
        by = img_b + act_b * (2 if xl is not None else 1)
        print("  image_convt_fwd %-6s %7.1f us  %6.0f GB/s" % (name, us, by / us / 1e3))
    # the column-buffer path they replace
    us1 = timeit(lambda: ops.im2col_k4s2(img))
    col = ops.im2col_k4s2(img)
    wp = ops.pack_matrix(w, C, 48, C, 64, 48, 1)
    us2 = timeit(lambda: ops.conv_fwd(col, wp, b, ops.KIND_CONV_K1S1, res // 2, res // 2, ops.ACT_LRELU))
    print("  (old) im2col %.1f us + K=64 GEMM %.1f us = %.1f us" % (us1, us2, us1 + us2))
    us3 = timeit(lambda: ops.conv_wgrad(dy, col, ops.KIND_CONV_K1S1, 1))
    print("  (old) wgrad from the column buffer %.1f us" % us3)
    wpt = ops.pack_matrix(w, 48, C, 64, C, 1, 48)
    us4 = timeit(lambda: ops.conv_fwd(dy, wpt, None, ops.KIND_CONV_K1S1, res // 2, res // 2))
    ycol = ops.conv_fwd(dy, wpt, None, ops.KIND_CONV_K1S1, res // 2, res // 2)
    us5 = timeit(lambda: ops.col2im_k4s2(ycol, b3, 3, ops.ACT_TANH))
    print("  (old) K=64 GEMM %.1f us + col2im %.1f us = %.1f us" % (us4, us5, us4 + us5))


def unsupported():
    """Shapes outside the fused kernels' set must be refused loudly (the Python side then takes the column-buffer path)."""
    img = torch.zeros(2, 3, 64, 64, device=dev)
    w = torch.zeros(8, 3, 4, 4, device=dev)
    try:
        ops.image_conv_fwd(img, w, None, ops.ACT_NONE)
        ok = False
    except Exception as e:  # GpError
        ok = "unsupported geometry" in str(e)
    print("%-66s %s" % ("image_conv_fwd refuses Cout=8", "ok" if ok else "FAILED"))
    if not ok:
        FAILED.append("refusal")
    assert not ops.image_edge_ok(3, 64, 64, 8) and ops.image_edge_ok(3, 64, 64, 128)
    assert ops.image_edge_ok(3, 64, 64, 128, transposed=True) and not ops.image_edge_ok(3, 128, 128, 64, transposed=True)


if __name__ == "__main__":
    if "--time-only" in sys.argv:   # one launch set for ncu
        timing()
        sys.exit(0)
    case(4, 64, 64, 0)
    case(3, 64, 128, 1)      # odd number of images: tiles never straddle images
    case(2, 32, 128, 2)      # Wo = 16: eight output rows per tile
    case(2, 128, 64, 4)      # Wo = 64: two output rows per tile
    case(160, 64, 128, 6)    # more tiles than resident CTAs: the persistent loops, ring reuse, stage wrap-around
    case(40, 32, 64, 7)
    case_t(4, 64, 64, 10)
    case_t(3, 32, 128, 11)   # Ws = 16
    case_t(3, 64, 128, 12)
    case_t(160, 64, 128, 14)
    case_t(300, 32, 64, 15)
    unsupported()
    if "--time" in sys.argv:
        timing()
    print("\nimage edge: %s" % ("ALL OK" if not FAILED else "FAILED: %s" % FAILED))
    sys.exit(1 if FAILED else 0)
