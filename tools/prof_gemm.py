"""Launch a few representative tensor-core GEMM calls of the DCGAN-64 step in isolation (for `ncu --set full`):
   python tools/prof_gemm.py [--batch 1024] [--cases img_fwd,d1_fwd_stats,d2_wgrad,...] [--reps 3]
Prints CUDA-event times per case (not under ncu) so the same script doubles as a micro-benchmark."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

from gan_playground_b200 import ops


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--cases", default="img_fwd,img_dgrad,d1_fwd,d1_fwd_stats,d2_fwd,d3_fwd,g0_fwd,g1_fwd,g2_fwd,g2_fwd_stats,d1_wgrad,d2_wgrad,d3_wgrad")
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    B = args.batch
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)

    def act(n, h, w, c):
        return torch.randn(n, h, w, c, device=dev).to(torch.bfloat16)

    def wt(n, k):
        return (torch.randn(n, k, device=dev) * 0.02).to(torch.bfloat16)

    cases = {}
    # D blocks.0 as an im2col GEMM (K = 64), with bias + LeakyReLU
    cases["img_fwd"] = lambda: ops.conv_fwd(X["col"], W["img"], Bv[128], ops.KIND_CONV_K1S1, 32, 32, ops.ACT_LRELU)
    cases["img_dgrad"] = lambda: ops.conv_fwd(X["a128"], W["imgT"], None, ops.KIND_CONV_K1S1, 32, 32)
    cases["d1_fwd"] = lambda: ops.conv_fwd(X["a128"], W["d1"], Bv[256], ops.KIND_CONV_K4S2, 16, 16)
    cases["d1_fwd_x3"] = lambda: ops.conv_fwd(X["a128"], W["d1x3"], Bv[256], ops.KIND_CONV_K4S2, 16, 16, x_lo=X["a128_lo"],
                                              out_mode="f32", stats=torch.zeros(2, 256, device=dev))
    # the same layer on fp16 operands (one MMA, fp32 output + fused statistics): the "fp16" forward mode
    cases["d1_fwd_f16"] = lambda: ops.conv_fwd(X["a128_h"], W["d1_h"], Bv[256],
                                               ops.KIND_CONV_K4S2, 16, 16, fp16_in=True, out_mode="f32",
                                               stats=torch.zeros(2, 256, device=dev))
    cases["d1_fwd_stats"] = lambda: ops.conv_fwd(X["a128"], W["d1"], Bv[256], ops.KIND_CONV_K4S2, 16, 16,
                                                 stats=torch.zeros(2, 256, device=dev))
    cases["g2_fwd_stats"] = lambda: ops.conv_fwd(X["a256_16"], W["g2"], Bv[128], ops.KIND_CONVT_K4S2, 32, 32,
                                                 stats=torch.zeros(2, 128, device=dev))
    cases["d2_fwd"] = lambda: ops.conv_fwd(X["a256_16"], W["d2"], Bv[512], ops.KIND_CONV_K4S2, 8, 8)
    cases["g0_fwd"] = lambda: ops.conv_fwd(X["a1024_4"], W["g0"], Bv[512], ops.KIND_CONVT_K4S2, 8, 8)
    cases["g1_fwd"] = lambda: ops.conv_fwd(X["a512_8"], W["g1"], Bv[256], ops.KIND_CONVT_K4S2, 16, 16)
    cases["g2_fwd"] = lambda: ops.conv_fwd(X["a256_16"], W["g2"], Bv[128], ops.KIND_CONVT_K4S2, 32, 32)
    cases["d3_fwd"] = lambda: ops.conv_fwd(X["a512_8"], W["d3"], Bv[1024], ops.KIND_CONV_K4S2, 4, 4)
    cases["d1_wgrad"] = lambda: ops.conv_wgrad(X["a256_16"], X["a128"], ops.KIND_CONV_K4S2, 16)
    cases["d2_wgrad"] = lambda: ops.conv_wgrad(X["a512_8"], X["a256_16"], ops.KIND_CONV_K4S2, 16)
    cases["d3_wgrad"] = lambda: ops.conv_wgrad(X["a1024_4"], X["a512_8"], ops.KIND_CONV_K4S2, 16)

    X = {"col": act(B, 32, 32, 64), "a128": act(B, 32, 32, 128), "a256_16": act(B, 16, 16, 256),
         "a512_8": act(B, 8, 8, 512), "a1024_4": act(B, 4, 4, 1024)}
    X["a128_lo"] = (X["a128"].float() * 2.0 ** -9).to(torch.bfloat16)
    W = {"d1x3": wt(256, 2 * 16 * 128), "img": wt(128, 64), "imgT": wt(64, 128), "d1": wt(256, 16 * 128), "g2": wt(128, 16 * 256),
         "d3": wt(1024, 16 * 512),
         "d2": wt(512, 16 * 256), "g0": wt(512, 16 * 1024), "g1": wt(256, 16 * 512)}
    X["a128_h"], W["d1_h"] = X["a128"].to(torch.float16), W["d1"].to(torch.float16)
    Bv = {n: torch.randn(n, device=dev) * 0.1 for n in (128, 256, 512, 1024)}

    for name in args.cases.split(","):
        fn = cases[name]
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        print("%-14s %8.1f us/call" % (name, e0.elapsed_time(e1) / args.reps * 1e3))


if __name__ == "__main__":
    main()
