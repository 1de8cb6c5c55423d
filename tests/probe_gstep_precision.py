"""Probe (not a collected test): gradient cosine of the G step of DCGAN-64 (width 64) against the fp32 CPU oracle as a
function of the batch size and of the forward operand format of the G step (fp16 single-MMA vs bf16x3).
    python tests/probe_gstep_precision.py [B ...]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from parity import global_cos, prebn_biases, quiet  # noqa: E402


def main():
    from gan_playground_b200 import config
    from gan_playground_b200.criterion import GANLoss
    from gan_playground_b200.models import dcgan
    from oracle import gan_oracle as O

    sizes = [int(a) for a in sys.argv[1:]] or [128, 256, 512, 1024]
    torch.set_num_threads(os.cpu_count())
    for B in sizes:
        for seed in (0, 1):
            torch.manual_seed(seed)
            netG, netD = quiet(lambda: dcgan.Generator()), quiet(lambda: dcgan.Discriminator())
            sd_g = {k: v.clone() for k, v in netG.state_dict().items()}
            sd_d = {k: v.clone() for k, v in netD.state_dict().items()}
            gen = torch.Generator().manual_seed(100 + B + seed)
            x = torch.rand(B, 3, 64, 64, generator=gen) * 2 - 1
            z = torch.randn(2, B, 100, generator=gen)
            ref = O.dcgan_step_grads(sd_g, sd_d, x, z[0], z[1])
            netG.cuda(), netD.cuda()
            crit = GANLoss("vanilla", 0.9, 0.1, 0.9).cuda()
            zd = z.cuda()
            skip = tuple(prebn_biases(netG))
            for mode in ("fp16", "bf16x3", "bf16"):
                netG.zero_grad(), netD.zero_grad()
                with config.precision_scope(mode):
                    out = netD(netG(zd[1]))
                crit(out, False, True).backward()
                err = ((out.detach().cpu().view(-1) - ref["d_g"].view(-1)).abs().max() / ref["d_g"].abs().max()).item()
                cos = global_cos(netG.named_parameters(), ref["g_grads"], skip)
                print("B=%4d seed %d G-step %-6s: D(G(z)) rel err %.2e, cos G-step %.6f" % (B, seed, mode, err, cos), flush=True)


if __name__ == "__main__":
    main()
