"""CTA-pair (tcgen05 cta_group::2) forward GEMM: correctness against torch fp32 convolutions on shapes large enough for
gp_conv_fwd to take the pair path (run with GP_FWD_2CTA=1), bf16 / fp16 / bf16x3 operands, odd tile counts, fused
statistics; then the timing of the DCGAN-64 layers with and without it.   python tools/check_pair.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
os.environ.setdefault("GP_FWD_2CTA", "1")

import torch
import torch.nn.functional as F

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def main():
    import selftest_conv as st
    from gan_playground_b200 import ops

    ok = True
    ok &= st.case_fwd_conv("k4s2", 160, 32, 128, 256, act=2, stats=True)
    ok &= st.case_fwd_conv("convt", 128, 16, 256, 256, act=1, stats=True)
    ok &= st.case_fwd_conv("k3s1", 75, 16, 64, 512)          # odd number of 128-pixel tiles: the peer's last tile is empty
    ok &= st.case_fwd_k1(20000, 512, 256)
    ok &= st.case_fwd_conv("k4s2", 1024, 16, 256, 512, stats=True)
    # bf16x3 and fp16 operands through ops.conv_fwd
    torch.manual_seed(0)
    NB, H, C, N = 256, 16, 128, 256
    x = torch.randn(NB, H, H, C, device="cuda")
    w = torch.randn(N, C, 3, 3, device="cuda") * 0.05
    b = torch.randn(N, device="cuda")
    ref = F.conv2d(x.permute(0, 3, 1, 2), w, b, padding=1).permute(0, 2, 3, 1)
    xb = x.bfloat16()
    lo = (x - xb.float()).bfloat16()
    y = ops.conv_fwd(xb, ops.split_conv_weight(w, 0), b, ops.KIND_CONV_K3S1, H, H, x_lo=lo, out_mode="f32")
    e3 = (y - ref).abs().max().item() / ref.abs().max().item()
    yh = ops.conv_fwd(x.half(), ops.conv_weight_f16(w, 0), b, ops.KIND_CONV_K3S1, H, H, fp16_in=True, out_mode="f32")
    eh = (yh - ref).abs().max().item() / ref.abs().max().item()
    print("pair path: bf16x3 rel err %.2e, fp16 rel err %.2e" % (e3, eh))
    ok &= e3 < 1e-4 and eh < 3e-3
    print("CHECK_PAIR", "OK" if ok else "FAILED")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
