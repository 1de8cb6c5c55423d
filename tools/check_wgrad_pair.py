"""CTA-pair wgrad (conv_gemm_kernel<MODE_WGRAD, 256, 1, false, true>): correctness against torch's fp32 weight gradient
on shapes that take the pair path (GP_WGRAD_2CTA=1), then the DCGAN-64 batch-1024 wgrad layers timed with and without it.
   python tools/check_wgrad_pair.py"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

import torch

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def timed(st, kind, NB, Hs, Cd, Cg, taps, iters=20):
    from gan_playground_b200 import _lib

    Hg = 2 * Hs if kind == "k4s2" else Hs
    dense = torch.randn(NB, Hs, Hs, Cd, device="cuda").bfloat16()
    gath = torch.randn(NB, Hg, Hg, Cg, device="cuda").bfloat16()
    dw = torch.zeros(Cd, taps, Cg, device="cuda")
    p = _lib.ConvWgrad(dense.data_ptr(), gath.data_ptr(), dw.data_ptr(), NB, Hs, Hs, Cd, Hg, Hg, Cg, st.KIND[kind])
    s = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    out = {}
    for flag in ("0", "1"):
        os.environ["GP_WGRAD_2CTA"] = flag
        for _ in range(3):
            _lib.check(_lib.lib().gp_conv_wgrad(ctypes.byref(p), s), "gp_conv_wgrad")
        tot = 0.0
        for _ in range(iters):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _lib.check(_lib.lib().gp_conv_wgrad(ctypes.byref(p), s), "gp_conv_wgrad")
            e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        out[flag] = tot / iters * 1e3
    print("wgrad %s B%d Hs%d Cd%d Cg%d: lone %.1f us, pair %.1f us" % (kind, NB, Hs, Cd, Cg, out["0"], out["1"]))


def main():
    import selftest_conv as st

    if len(sys.argv) > 1 and sys.argv[1] == "--one":      # a short run for ncu
        timed(st, "k4s2", 1024, 8, 512, 256, 16, iters=1)
        return 0
    os.environ["GP_WGRAD_2CTA"] = "1"
    ok = True
    ok &= st.case_wgrad("k4s2", 256, 8, 512, 256)
    ok &= st.case_wgrad("k4s2", 128, 16, 256, 128)
    ok &= st.case_wgrad("k4s2", 96, 8, 384, 320)        # dW rows not a multiple of 256: the peer's rows are masked
    ok &= st.case_wgrad("k3s1", 64, 16, 256, 256)
    ok &= st.case_wgrad("k1s1", 4096, 2, 512, 512)
    print("CHECK_WGRAD_PAIR", "OK" if ok else "FAILED")
    if ok:
        for (Hs, Cd, Cg) in ((16, 256, 128), (8, 512, 256), (4, 1024, 512), (32, 128, 128)):
            timed(st, "k4s2", 1024, Hs, Cd, Cg, 16)
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
