"""Run-to-run noise of a module-level G step at tiny batch (the configuration of tests/test_gpu_weight_stage.py):
per-parameter relative gradient difference between two runs with batched staging, two without, and across."""
import contextlib, io, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from gan_playground_b200 import config
from gan_playground_b200.criterion import GANLoss
from gan_playground_b200.models import dcgan


def run(batched, batch=8):
    config.set_batch_stage(batched)
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        netG, netD = dcgan.Generator(ngf=32, resolution=32).cuda(), dcgan.Discriminator(ndf=32, resolution=32).cuda()
    crit = GANLoss("vanilla", 0.9, 0.1, 0.9).cuda()
    z = torch.randn(batch, 100, generator=torch.Generator().manual_seed(1)).cuda()
    crit(netD(netG(z)), False, True).backward()
    with torch.no_grad():
        for p in list(netG.parameters()) + list(netD.parameters()):
            p.grad = None
            p.mul_(1.0)
    img = netG(z)
    img.retain_grad()
    loss = crit(netD(img), False, True)
    loss.backward()
    torch.cuda.synchronize()
    named = [("G." + n, p) for n, p in netG.named_parameters()] + [("D." + n, p) for n, p in netD.named_parameters()]
    out = {n: (p.grad.clone() if p.grad is not None else torch.zeros_like(p)) for n, p in named}
    out["image"] = img.detach().clone()
    out["d_image"] = img.grad.clone()
    return loss.item(), out


def rel(a, b):
    return float((a - b).norm() / (b.norm() + 1e-30))


for batch in (8, 64):
    runs = {k: run(k[0] == "T", batch) for k in ("T1", "T2", "F1", "F2")}
    print(f"batch {batch}: losses", {k: v[0] for k, v in runs.items()})
    names = list(runs["T1"][1])
    print(f"{'tensor':34s} {'T1-T2':>10s} {'F1-F2':>10s} {'T1-F1':>10s}")
    for n in names:
        print(f"{n:34s} {rel(runs['T1'][1][n], runs['T2'][1][n]):10.2e} {rel(runs['F1'][1][n], runs['F2'][1][n]):10.2e} "
              f"{rel(runs['T1'][1][n], runs['F1'][1][n]):10.2e}")
