"""Small, fast cases for compute-sanitizer (memcheck / racecheck / synccheck) over the hand-rolled pipelines:

  * gp::conv_gemm_kernel — TMA producer / tcgen05 issuer / epilogue warps talking through mbarriers and TMEM — in every
    tile shape (BN 64 / 128 / 256 x MT 1 / 2, forced with GP_TILE_FWD / GP_TILE_WGRAD), forward (bf16, bf16x3, fp16 with the
    fused fp16 residual) and wgrad (hybrid stream-K with partial tiles);
  * the streaming BatchNorm / companion-format kernels and the loss kernel (shared-memory reductions);
  * `--peer` (run under torchrun with 2 ranks): the one-shot NVLink peer exchange of csrc/peer_sync.cu (flags + slots).

    compute-sanitizer --tool memcheck python tools/sanitize_cases.py
    compute-sanitizer --tool racecheck python tools/sanitize_cases.py --quick

Every case also checks its result against torch (a sanitizer-clean wrong answer is still wrong). Exit code 0 = all ok."""
import argparse
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

import torch  # noqa: E402

# torch references on the GPU must be true fp32 (the default lets cuDNN / cuBLAS use TF32: 10-bit operands)
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def gemm_cases(quick):
    import selftest_conv as st

    ok = True
    # 256 output channels / dW columns: every forced tile width (64 / 128 / 256) is a natural fit; ragged pixel counts
    ok &= st.case_fwd_k1(300, 256, 128)
    ok &= st.case_fwd_conv("k4s2", 5, 8, 64, 256, act=2, stats=True)
    ok &= st.case_fwd_conv("convt", 5, 4, 64, 256, act=1, stats=True)
    ok &= st.case_fwd_conv("k3s1", 3, 8, 64, 256)
    ok &= st.case_wgrad("k4s2", 9, 4, 256, 256)
    ok &= st.case_wgrad("k3s1", 5, 8, 256, 256)
    if not quick:
        ok &= st.case_fwd_conv("k4s2", 16, 16, 128, 256, stats=True)
        ok &= st.case_wgrad("k1s1", 130, 1, 256, 256)
    return ok


def mode_cases():
    """bf16x3 / fp16 forward modes, fused fp16 residual, and the companion-format element-wise kernels."""
    import torch.nn.functional as F

    from gan_playground_b200 import ops

    torch.manual_seed(0)
    ok = True
    NB, H, C, N = 4, 8, 64, 128
    x = torch.randn(NB, H, H, C, device="cuda")
    w = torch.randn(N, C, 3, 3, device="cuda") * 0.05
    b = torch.randn(N, device="cuda")
    ref = F.conv2d(x.permute(0, 3, 1, 2), w, b, padding=1).permute(0, 2, 3, 1)
    xb = x.bfloat16()
    lo = (x - xb.float()).bfloat16()
    hi_o, lo_o = ops.conv_fwd(xb, ops.split_conv_weight(w, 0), b, ops.KIND_CONV_K3S1, H, H, x_lo=lo, out_mode="split")
    e3 = ((hi_o.float() + lo_o.float()) - ref).abs().max().item() / ref.abs().max().item()
    res = torch.randn(NB, H, H, N, device="cuda").half()
    ob, oh = ops.conv_fwd(x.half(), ops.conv_weight_f16(w, 0), b, ops.KIND_CONV_K3S1, H, H, ops.ACT_RELU, residual=res,
                          fp16_in=True, out_mode="pair")
    eh = (oh.float() - F.relu(ref + res.float())).abs().max().item() / ref.abs().max().item()
    print("bf16x3 k3s1 rel err %.2e | fp16 k3s1 + fp16 residual + relu rel err %.2e" % (e3, eh))
    ok &= e3 < 1e-4 and eh < 3e-3
    # the two outputs of a "pair" store are roundings of the same fp32 value
    ok &= (ob.float() - oh.float()).abs().max().item() <= 2 ** -8 * oh.float().abs().max().item()
    # BatchNorm on an activation with a companion, pooling, blur, upsample
    st = ops.bn_stats_comp(ob, oh)
    fin = ops.bn_finalize(st, NB * H * H, None, None, None, None, None)
    a, ac = ops.bn_apply_act_comp(ob, oh, fin, ops.ACT_LRELU, ops.COMP_F16)
    want = F.leaky_relu(F.batch_norm(oh.float().permute(0, 3, 1, 2), None, None, None, None, True, 0.1, 1e-5), 0.2)
    checks = {"x3/fp16 conv": bool(ok)}
    checks["bn apply (companion)"] = (ac.float().permute(0, 3, 1, 2) - want).abs().max().item() < 5e-3
    p, pc = ops.pool2x(a, 0.25, comp=ac, out_fmt=ops.COMP_F16)
    checks["pool2x"] = (pc.float().permute(0, 3, 1, 2) - F.avg_pool2d(ac.float().permute(0, 3, 1, 2), 2)).abs().max().item() < 2e-3
    bl, blc = ops.blur3x3_fwd(a, 2, comp=ac, out_fmt=ops.COMP_F16)
    checks["blur"] = bl.shape == (NB, H // 2, H // 2, N) and torch.isfinite(blc.float()).all().item()
    up = ops.upsample2x(ac, 1.0)
    checks["upsample"] = up.dtype == torch.float16 and torch.equal(up[:, ::2, ::2], ac)
    loss, dpred = ops.gan_loss(torch.randn(33, 1, device="cuda"), ops.LOSS_BCE, 0.9)
    checks["loss"] = bool(torch.isfinite(loss).item())
    print("mode cases:", checks)
    ok = all(checks.values())
    torch.cuda.synchronize()
    return bool(ok)


def tile_sweep(quick):
    """Re-run the GEMM cases in a child per forced tile shape (the choice is cached per process in static state)."""
    ok = True
    shapes = ["64x1", "128x1", "128x2", "256x1"] + ([] if quick else ["64x2", "256x2"])
    for t in shapes:
        env = dict(os.environ, GP_TILE_FWD=t, GP_TILE_WGRAD=t if t != "256x2" else "256x1")
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--only-gemm"] + (["--quick"] if quick else []), env=env)
        print("tile %s rc=%d" % (t, r.returncode), flush=True)
        ok &= r.returncode == 0
    return ok


def peer_case():
    from gan_playground_b200 import ops, parallel

    rank, world = parallel.init()
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    assert parallel.init_peer_sync(dev), "peer exchange unavailable"
    ok = True
    for it in range(20):
        t = torch.full((2, 256), float(rank + 1 + it), device=dev)
        ops.peer_allreduce_sum_(parallel.peer_ctx(), t)
        want = sum(r + 1 + it for r in range(world))
        ok &= bool((t == want).all().item())
    torch.cuda.synchronize()
    torch.distributed.barrier()
    print("rank %d peer exchange x20: %s" % (rank, "ok" if ok else "MISMATCH"), flush=True)
    return ok


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--only-gemm", action="store_true")
    ap.add_argument("--peer", action="store_true")
    a = ap.parse_args()
    if a.peer:
        good = peer_case()
        sys.stdout.flush()
        os._exit(0 if good else 1)          # skip NCCL teardown under the sanitizer
    if a.only_gemm:
        sys.exit(0 if gemm_cases(a.quick) else 1)
    good = tile_sweep(a.quick)
    good &= mode_cases()
    print("SANITIZE CASES:", "ALL OK" if good else "FAILED")
    sys.exit(0 if good else 1)
