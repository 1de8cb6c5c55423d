"""Do an HBM-bound streaming kernel and a tensor-bound persistent GEMM overlap when launched on two streams?
The GEMM CTAs (192 threads, ~200 KB of shared memory, 114-248 registers per thread) leave room on every SM for the
256-thread blocks of the BatchNorm kernels. Times N launches of each alone, then both concurrently.
   python tools/overlap_probe.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch


def timed(fn, reps=5):
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def main():
    from gan_playground_b200 import ops

    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    NB, N = 1024, 8
    # d2 layer of DCGAN-64: dgrad-like forward GEMM and its wgrad
    x = torch.randn(NB, 16, 16, 128, device=dev).bfloat16()
    w = torch.randn(256, 128, 4, 4, device=dev) * 0.05
    wp = ops.pack_conv_weight(w, 0)
    dy = torch.randn(NB, 8, 8, 256, device=dev).bfloat16()
    # BatchNorm backward of the 32x32x64 layer (fp32 pre-BN tensor)
    y = torch.randn(NB, 32, 32, 64, device=dev)
    da = torch.randn(NB, 32, 32, 64, device=dev).bfloat16()
    st = ops.bn_stats_f32(y)
    fin = ops.bn_finalize(st, y.numel() // 64, torch.ones(64, device=dev), torch.zeros(64, device=dev), None, None, None)
    red = ops.bn_bwd_reduce_f32(da, y, fin, ops.ACT_LRELU)
    s_main, s_side = torch.cuda.current_stream(), torch.cuda.Stream()

    def gemm_fwd():
        for _ in range(N):
            ops.conv_fwd(x, wp, None, ops.KIND_CONV_K4S2, 8, 8)

    def gemm_wgrad():
        for _ in range(N):
            ops.conv_wgrad(dy, x, ops.KIND_CONV_K4S2, 16)

    def ew():
        for _ in range(N):
            ops.bn_bwd_reduce_f32(da, y, fin, ops.ACT_LRELU)
            ops.bn_bwd_apply_f32(da, y, fin, red, y.numel() // 64, ops.ACT_LRELU)

    def both(g, gemm_first=False):
        def run():
            s_side.wait_stream(s_main)
            if gemm_first:
                g()
            with torch.cuda.stream(s_side):
                ew()
            if not gemm_first:
                g()
            s_main.wait_stream(s_side)
        return run

    for _ in range(2):
        gemm_fwd(), gemm_wgrad(), ew()
    t_f, t_w, t_e = timed(gemm_fwd), timed(gemm_wgrad), timed(ew)
    t_fe, t_we = timed(both(gemm_fwd)), timed(both(gemm_wgrad))
    print("%d launches each: fwd GEMM %.3f ms, wgrad %.3f ms, BN backward (reduce + apply) %.3f ms" % (N, t_f, t_w, t_e))
    print("concurrent: fwd GEMM + BN backward %.3f ms (serial %.3f, ideal %.3f)" % (t_fe, t_f + t_e, max(t_f, t_e)))
    print("concurrent: wgrad    + BN backward %.3f ms (serial %.3f, ideal %.3f)" % (t_we, t_w + t_e, max(t_w, t_e)))
    t_fe, t_we = timed(both(gemm_fwd, True)), timed(both(gemm_wgrad, True))
    print("GEMMs enqueued first: fwd + BN backward %.3f ms, wgrad + BN backward %.3f ms" % (t_fe, t_we))


if __name__ == "__main__":
    main()
