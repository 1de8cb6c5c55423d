"""On-GPU bring-up check of the tcgen05 implicit-GEMM kernels against torch fp32 convolutions.

Run on a B200:  python tools/selftest_conv.py            (runs every case in its own subprocess with a timeout)
                python tools/selftest_conv.py --case fwd_k1_small
Diagnostic output is verbose on purpose: one GPU round trip should tell what is wrong.
"""
import argparse
import ctypes
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch
import torch.nn.functional as F

from gan_playground_b200 import _lib

# the references below are torch convolutions on the GPU: they must be TRUE fp32 (torch's default lets cuDNN / cuBLAS use
# TF32, i.e. 10-bit operands — SURVEY.md D8), or the check would be TF32 judging bf16
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False

KIND = {"k4s2": 0, "convt": 1, "k3s1": 2, "k1s1": 3}


def bf(x):
    return x.to(torch.bfloat16)


def report(name, got, ref, tol=2e-2):
    got = got.float()
    ref = ref.float()
    err = (got - ref).abs()
    denom = ref.abs().max().item() + 1e-20
    rel = err.max().item() / denom
    ok = rel < tol and torch.isfinite(got).all().item()
    print("%-28s %s  max_abs_err=%.4e  max_ref=%.4e  rel=%.3e  shape=%s" %
          (name, "OK  " if ok else "FAIL", err.max().item(), denom, rel, tuple(got.shape)))
    if not ok:
        flat = err.flatten()
        bad = (flat > tol * denom).nonzero().flatten()
        print("   mismatches: %d / %d ; nonfinite=%d" % (bad.numel(), flat.numel(), (~torch.isfinite(got)).sum().item()))
        g2 = got.reshape(-1, got.shape[-1])
        r2 = ref.reshape(-1, ref.shape[-1])
        e2 = (g2 - r2).abs() > tol * denom
        rows_bad = e2.any(dim=1).nonzero().flatten()
        cols_bad = e2.any(dim=0).nonzero().flatten()
        print("   bad rows: n=%d first=%s" % (rows_bad.numel(), rows_bad[:24].tolist()))
        print("   bad cols: n=%d first=%s" % (cols_bad.numel(), cols_bad[:24].tolist()))
        print("   got[0,:8]=%s" % g2[0, :8].tolist())
        print("   ref[0,:8]=%s" % r2[0, :8].tolist())
        if rows_bad.numel():
            r = rows_bad[0].item()
            print("   got[%d,:8]=%s" % (r, g2[r, :8].tolist()))
            print("   ref[%d,:8]=%s" % (r, r2[r, :8].tolist()))
        ratio = (g2.abs().sum() / (r2.abs().sum() + 1e-20)).item()
        print("   sum|got|/sum|ref| = %.4f" % ratio)
    return ok


def conv_fwd(x_nhwc, w_packed, bias, kind, Hout, Wout, act=0, stats=False):
    NB, Hin, Win, Cin = x_nhwc.shape
    Nout = w_packed.shape[0]
    out = torch.full((NB, Hout, Wout, Nout), float("nan"), device="cuda", dtype=torch.bfloat16)
    s = ss = None
    if stats:
        s = torch.zeros(Nout, device="cuda")
        ss = torch.zeros(Nout, device="cuda")
    p = _lib.ConvFwd(x_nhwc.data_ptr(), w_packed.data_ptr(), bias.data_ptr() if bias is not None else None,
                     out.data_ptr(), s.data_ptr() if stats else None, ss.data_ptr() if stats else None,
                     NB, Hin, Win, Cin, Hout, Wout, Nout, KIND[kind], act, None)
    rc = _lib.lib().gp_conv_fwd(ctypes.byref(p), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    _lib.check(rc, "gp_conv_fwd")
    torch.cuda.synchronize()
    return (out, s, ss) if stats else out


def conv_wgrad(dense, gath, kind, taps):
    NB, Hs, Ws, Cd = dense.shape
    _, Hg, Wg, Cg = gath.shape
    dw = torch.zeros(Cd, taps, Cg, device="cuda")
    p = _lib.ConvWgrad(dense.data_ptr(), gath.data_ptr(), dw.data_ptr(), NB, Hs, Ws, Cd, Hg, Wg, Cg, KIND[kind])
    rc = _lib.lib().gp_conv_wgrad(ctypes.byref(p), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    _lib.check(rc, "gp_conv_wgrad")
    torch.cuda.synchronize()
    return dw


def case_fwd_k1(M, N, K, seed=0, structured=False):
    g = torch.Generator(device="cuda").manual_seed(seed)
    if structured:
        # A[m,k] = 1 if k == m % K ; W[n,k] = n + k/1000  -> out[m,n] = W[n, m%K]: reveals row/col/swizzle permutations
        a = torch.zeros(M, K, device="cuda")
        a[torch.arange(M), torch.arange(M) % K] = 1.0
        w = (torch.arange(N, device="cuda").float()[:, None] + torch.arange(K, device="cuda").float()[None, :] / 64.0)
    else:
        a = torch.randn(M, K, device="cuda", generator=g)
        w = torch.randn(N, K, device="cuda", generator=g) * 0.1
    a, w = bf(a), bf(w)
    bias = torch.randn(N, device="cuda", generator=g)
    out = conv_fwd(a.view(M, 1, 1, K).contiguous(), w.contiguous(), bias, "k1s1", 1, 1)
    ref = a.float() @ w.float().t() + bias
    return report("fwd_k1 M%d N%d K%d%s" % (M, N, K, " struct" if structured else ""), out.view(M, N), ref)


def case_fwd_conv(kind, NB, Hin, Cin, Cout, act=0, stats=False, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = bf(torch.randn(NB, Cin, Hin, Hin, device="cuda", generator=g))
    bias = torch.randn(Cout, device="cuda", generator=g)
    if kind == "k4s2":
        w = bf(torch.randn(Cout, Cin, 4, 4, device="cuda", generator=g) * 0.05)
        ref = F.conv2d(x.float(), w.float(), bias, stride=2, padding=1)
        wp = w.permute(0, 2, 3, 1).contiguous().view(Cout, -1)
        Hout = Hin // 2
    elif kind == "convt":
        w = bf(torch.randn(Cin, Cout, 4, 4, device="cuda", generator=g) * 0.05)
        ref = F.conv_transpose2d(x.float(), w.float(), bias, stride=2, padding=1)
        wp = w.permute(1, 2, 3, 0).contiguous().view(Cout, -1)
        Hout = Hin * 2
    elif kind == "k3s1":
        w = bf(torch.randn(Cout, Cin, 3, 3, device="cuda", generator=g) * 0.05)
        ref = F.conv2d(x.float(), w.float(), bias, stride=1, padding=1)
        wp = w.permute(0, 2, 3, 1).contiguous().view(Cout, -1)
        Hout = Hin
    else:
        raise ValueError(kind)
    pre = ref
    if act == 2:
        ref = F.leaky_relu(ref, 0.2)
    elif act == 1:
        ref = F.relu(ref)
    elif act == 3:
        ref = torch.tanh(ref)
    x_nhwc = x.permute(0, 2, 3, 1).contiguous()
    res = conv_fwd(x_nhwc, wp, bias, kind, Hout, Hout, act=act, stats=stats)
    name = "fwd_%s B%d H%d C%d->%d" % (kind, NB, Hin, Cin, Cout)
    if stats:
        out, s, ss = res
        ok = report(name, out, ref.permute(0, 2, 3, 1))
        ok &= report(name + " sum", s[None], pre.sum(dim=(0, 2, 3))[None], tol=1e-3)
        ok &= report(name + " sumsq", ss[None], (pre * pre).sum(dim=(0, 2, 3))[None], tol=1e-3)
        return ok
    return report(name, res, ref.permute(0, 2, 3, 1))


def case_wgrad(kind, NB, Hs, Cd, Cg, seed=0):
    """kind k4s2: dense on (Hs,Hs) grid, gath on (2Hs,2Hs): dW[m,kh,kw,n] = sum dense[p,m]*gath[2p-1+k, n]."""
    g = torch.Generator(device="cuda").manual_seed(seed)
    if kind == "k4s2":
        Hg, k, stride, pad = 2 * Hs, 4, 2, 1
    elif kind == "k3s1":
        Hg, k, stride, pad = Hs, 3, 1, 1
    else:
        Hg, k, stride, pad = Hs, 1, 1, 0
    dy = bf(torch.randn(NB, Cd, Hs, Hs, device="cuda", generator=g))
    x = bf(torch.randn(NB, Cg, Hg, Hg, device="cuda", generator=g))
    w = torch.zeros(Cd, Cg, k, k, device="cuda", requires_grad=True)
    y = F.conv2d(x.float(), w, None, stride=stride, padding=pad)
    (gw,) = torch.autograd.grad(y, w, dy.float())
    ref = gw.permute(0, 2, 3, 1).contiguous()  # [Cd][kh][kw][Cg]
    dw = conv_wgrad(dy.permute(0, 2, 3, 1).contiguous(), x.permute(0, 2, 3, 1).contiguous(), kind, k * k)
    return report("wgrad_%s B%d Hs%d Cd%d Cg%d" % (kind, NB, Hs, Cd, Cg), dw.view(Cd, k * k * Cg), ref.view(Cd, -1),
                  tol=5e-3)


CASES = {
    "fwd_k1_struct": lambda: case_fwd_k1(128, 64, 64, structured=True),
    "fwd_k1_small": lambda: case_fwd_k1(128, 64, 64),
    "fwd_k1_k256": lambda: case_fwd_k1(128, 128, 256),
    "fwd_k1_multi": lambda: case_fwd_k1(1000, 256, 512),
    "fwd_k1_big": lambda: case_fwd_k1(40000, 512, 1024),
    "fwd_k4s2_a": lambda: case_fwd_conv("k4s2", 8, 8, 64, 64),
    "fwd_k4s2_b": lambda: case_fwd_conv("k4s2", 32, 32, 128, 256, act=2),
    "fwd_k4s2_c": lambda: case_fwd_conv("k4s2", 20, 8, 512, 1024, stats=True),
    "fwd_convt_a": lambda: case_fwd_conv("convt", 8, 4, 64, 64),
    "fwd_convt_b": lambda: case_fwd_conv("convt", 24, 8, 512, 256, act=1),
    "fwd_convt_c": lambda: case_fwd_conv("convt", 8, 16, 256, 128, stats=True),
    "fwd_k3s1_a": lambda: case_fwd_conv("k3s1", 8, 8, 128, 128),
    "wgrad_k1_a": lambda: case_wgrad("k1s1", 64, 1, 128, 64),
    "wgrad_k1_b": lambda: case_wgrad("k1s1", 512, 2, 256, 256),
    "wgrad_k4s2_a": lambda: case_wgrad("k4s2", 8, 4, 128, 64),
    "wgrad_k4s2_b": lambda: case_wgrad("k4s2", 64, 8, 512, 256),
    "wgrad_k4s2_c": lambda: case_wgrad("k4s2", 16, 16, 256, 128),
    "wgrad_k3s1_a": lambda: case_wgrad("k3s1", 16, 8, 128, 128),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", default=None)
    ap.add_argument("--timeout", type=int, default=90)
    args = ap.parse_args()
    if args.case:
        torch.manual_seed(0)
        ok = CASES[args.case]()
        sys.exit(0 if ok else 1)
    results = {}
    for name in CASES:
        try:
            pr = subprocess.run([sys.executable, os.path.abspath(__file__), "--case", name], timeout=args.timeout,
                                stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            out = pr.stdout
            rc = pr.returncode
        except subprocess.TimeoutExpired as e:
            out = (e.stdout or b"").decode() if isinstance(e.stdout, bytes) else (e.stdout or "")
            rc = "TIMEOUT"
        lines = [l for l in out.splitlines() if l.strip()]
        print("==== %s rc=%s" % (name, rc))
        for l in lines[-30:]:
            print("   " + l)
        results[name] = rc
        sys.stdout.flush()
    print("SUMMARY:", {k: v for k, v in results.items()})
    nfail = sum(1 for v in results.values() if v != 0)
    print("FAILED: %d / %d" % (nfail, len(results)))
    sys.exit(1 if nfail else 0)


if __name__ == "__main__":
    main()
