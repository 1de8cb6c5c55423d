"""Per-kernel breakdown of one eager training step of a bench.py configuration on the GPU: every C-ABI call is bracketed by CUDA events
(ops.CallProfiler); tensor-core GEMMs are reported as algorithmic TFLOP/s per launch, HBM-bound kernels as algorithmic
GB/s (each input read once + each output written once) against MEASURED_PEAKS.json.

    python tools/step_breakdown.py [--batch 1024] [--precision bf16|bf16x3] [--gemms]
"""
import argparse
import collections
import contextlib
import io
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch


def algorithmic_bytes(name, a):
    """Compulsory HBM bytes of one call from its raw C-ABI arguments (see include/gpb200.h for the signatures)."""
    if name in ("gp_bn_stats", "gp_colsum"):
        return a[1] * a[2] * 2
    if name == "gp_bn_stats_f32":
        return a[1] * a[2] * 4
    if name == "gp_bn_apply_act":
        return a[2] * a[3] * 4
    if name in ("gp_bn_apply_act_split", "gp_bn_apply_act_pair"):
        return a[3] * a[4] * 8
    if name == "gp_pair_to_f16":      # (hi, lo, ld_in, out, ld_out, rows, cols): two bf16 reads + one fp16 write
        return a[5] * a[6] * 6
    if name == "gp_bn_bwd_reduce":
        return a[2] * a[3] * 4
    if name == "gp_bn_bwd_reduce_f32":
        return a[2] * a[3] * 6
    if name == "gp_bn_bwd_apply":
        return a[3] * a[4] * 6
    if name == "gp_bn_bwd_apply_f32":
        return a[3] * a[4] * 8
    if name == "gp_act_bwd":
        return a[3] * 6
    if name in ("gp_im2col_k4s2", "gp_im2col_k4s2_split"):
        if name == "gp_im2col_k4s2":
            img, mul, col, NB, ch, Hi, Wi = a[:7]
            return NB * ch * Hi * Wi * 4 * (2 if mul else 1) + NB * (Hi // 2) * (Wi // 2) * 64 * 2
        img, hi, lo, NB, ch, Hi, Wi = a[:7]
        return NB * ch * Hi * Wi * 4 + NB * (Hi // 2) * (Wi // 2) * 64 * 4
    if name in ("gp_col2im_k4s2", "gp_col2im_k4s2_f32"):
        col, bias, img, NB, ch, Hi, Wi = a[:7]
        return NB * ch * Hi * Wi * 4 + NB * (Hi // 2) * (Wi // 2) * 64 * (2 if name == "gp_col2im_k4s2" else 4)
    if name == "gp_image_conv_k4s2_fwd":   # fused image-edge kernels (csrc/image_edge.cu): image (+ tanh factor) in, activation(s) out
        img, mul, w, bias, out, comp, fmt, NB, ch, Hi, Wi, C = a[:12]
        return NB * ch * Hi * Wi * 4 * (2 if mul else 1) + NB * (Hi // 2) * (Wi // 2) * C * 2 * (2 if comp else 1)
    if name == "gp_image_conv_k4s2_wgrad":
        dense, img, mul, dw, db, NB, ch, Hi, Wi, M = a[:10]
        return NB * ch * Hi * Wi * 4 * (2 if mul else 1) + NB * (Hi // 2) * (Wi // 2) * M * 2
    if name == "gp_image_convt_k4s2_fwd":
        x, x_lo, fmt, w, bias, img, NB, Hs, Ws, C, ch = a[:11]
        return NB * Hs * Ws * C * 2 * (2 if x_lo else 1) + NB * ch * 4 * Hs * Ws * 4
    if name == "gp_image_bias_grad":
        dout, mul, db, NB, ch, HW = a[:6]
        return NB * ch * HW * 4 * (2 if mul else 1)
    if name in ("gp_head_fwd", "gp_head_fwd_split"):
        off = 0 if name == "gp_head_fwd" else 1
        NB, HW, C = a[4 + off], a[5 + off], a[6 + off]
        return NB * HW * C * 2 * (1 + off)
    if name == "gp_head_bwd":
        NB, HW, C = a[6], a[7], a[8]
        return NB * HW * C * 2 * ((1 if a[3] else 0) + (1 if a[4] else 0))
    if name in ("gp_pack_conv_weight", "gp_split_conv_weight"):
        n = a[2] * a[3] * a[4]
        return n * (4 + (2 if name == "gp_pack_conv_weight" else 4))
    if name == "gp_unpack_conv_wgrad":
        return a[2] * a[3] * a[4] * 8
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--precision", default=None)
    ap.add_argument("--gemms", action="store_true", help="list every GEMM launch")
    ap.add_argument("--out", default=None)
    ap.add_argument("--config", default="cfg2", help="bench.py configuration (cfg2 DCGAN-64, cfg3 SN-DCGAN-32, cfg4 SNGAN "
                                                     "projection, cfg5 ACGAN-64): same nets / step driver / optimiser as the bench")
    args = ap.parse_args()
    from gan_playground_b200 import config, ops
    if args.precision:
        config.set_precision(args.precision)
    import bench

    dev = torch.device("cuda", 0)
    runner, inputs = bench.build_config(args.config, dev, args.batch, 1, False, False)
    for _ in range(3):
        runner.step_eager(*inputs)
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(5):
        runner.step_eager(*inputs)
    t1.record()
    torch.cuda.synchronize()
    step_ms = t0.elapsed_time(t1) / 5
    prof = ops.CallProfiler()
    with prof:
        runner.step_eager(*inputs)
    torch.cuda.synchronize()

    peaks = {"hbm_gbs": 6550.7, "bf16_tflops_sustained": 1361.6}
    try:
        peaks.update(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))))
    except OSError:
        pass
    agg = collections.OrderedDict()
    gemms = []
    tot = 0.0
    for name, a, e0, e1, note in prof.records:
        ms = e0.elapsed_time(e1)
        tot += ms
        if note is not None and not name.startswith("gp_image_conv"):
            gemms.append((note[0], ms, note[1]))
            key = name + " [tensor]"
            ent = agg.setdefault(key, [0, 0.0, 0.0, 0.0])
            ent[0] += 1; ent[1] += ms; ent[3] += note[1]
        else:
            b = algorithmic_bytes(name, a)
            ent = agg.setdefault(name, [0, 0.0, 0.0, 0.0])
            ent[0] += 1; ent[1] += ms; ent[2] += (b or 0)
    lines = []
    lines.append("%s batch %d precision %s: eager step %.3f ms (5-step mean); profiled own kernels %.3f ms in %d calls"
                 % (args.config, args.batch, bench.precision_label(runner), step_ms, tot, len(prof.records)))
    lines.append("%-28s %5s %10s %7s  %s" % ("entry point", "calls", "total us", "share", "achieved (of measured peak)"))
    for k, (n, ms, by, fl) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if fl:
            perf = "%7.1f TFLOP/s (%.2f of %.0f sustained)" % (fl / ms / 1e9, fl / ms / 1e9 / peaks["bf16_tflops_sustained"], peaks["bf16_tflops_sustained"])
        elif by:
            perf = "%7.0f GB/s    (%.2f of %.0f)" % (by / ms / 1e6, by / ms / 1e6 / peaks["hbm_gbs"], peaks["hbm_gbs"])
        else:
            perf = ""
        lines.append("%-28s %5d %10.1f %6.1f%%  %s" % (k, n, ms * 1e3, 100 * ms / step_ms, perf))
    lines.append("%-28s %5s %10.1f %6.1f%%  (torch host ops: Adam, fills, grad accumulation, launch gaps)"
                 % ("not in own kernels", "", (step_ms - tot) * 1e3, 100 * (step_ms - tot) / step_ms))
    if args.gemms:
        # HBM-bound calls by (entry point, problem size): where the per-kernel averages above come from
        lines.append("")
        shp = collections.OrderedDict()
        for name, a, e0, e1, note in prof.records:
            if note is not None:
                continue
            b = algorithmic_bytes(name, a)
            if not b:
                continue
            ent = shp.setdefault((name, b), [0, 0.0])
            ent[0] += 1; ent[1] += e0.elapsed_time(e1)
        for (name, b), (n, ms) in sorted(shp.items(), key=lambda kv: -kv[1][1])[:40]:
            lines.append("    %-26s %7.1f MB x%2d %9.1f us/call %7.0f GB/s" % (name, b / 1e6, n, ms / n * 1e3, b * n / ms / 1e6))
        lines.append("")
        for i, (label, ms, fl) in enumerate(gemms):
            lines.append("%3d %-52s %8.1f us %7.1f TFLOP/s" % (i, label, ms * 1e3, fl / ms / 1e9))
    text = "\n".join(lines)
    print(text)
    if args.out:
        with open(args.out, "w") as f:
            f.write(text + "\n")


if __name__ == "__main__":
    main()
