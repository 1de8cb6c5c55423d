"""On-GPU bring-up check of the DCGAN modules against the golden fixtures (generated from the reference) and the
CPU oracle at wider configs. Verbose diagnostics; each case in a subprocess with a timeout."""
import argparse
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch

from gan_playground_b200.criterion import GANLoss
from gan_playground_b200.models import dcgan


def unpack(d):
    return {k: v["q"].float() * v["scale"] for k, v in d.items()}


def cos(a, b):
    a, b = a.flatten().double(), b.flatten().double()
    return (a @ b / (a.norm() * b.norm() + 1e-30)).item()


def relerr(a, b):
    return ((a - b).abs().max() / (b.abs().max() + 1e-30)).item()


def cmp_grads(name, got, ref, skip=()):
    worst = 1.0
    num = den_a = den_b = 0.0
    for k, r in ref.items():
        g = got.get(k)
        if g is None:
            print("   %s: MISSING grad %s" % (name, k))
            worst = -1
            continue
        g = g.detach().float().cpu()
        c = cos(g, r)
        tag = ""
        if any(s in k for s in skip):
            tag = " (excluded: analytically-zero pre-BN bias)"
        else:
            worst = min(worst, c)
            num += (g.double() * r.double()).sum().item()
            den_a += (g.double() ** 2).sum().item()
            den_b += (r.double() ** 2).sum().item()
        print("   %-32s cos=%.6f  |got|=%.3e |ref|=%.3e%s" % (k, c, g.norm().item(), r.norm().item(), tag))
    glob = num / ((den_a ** 0.5) * (den_b ** 0.5) + 1e-30)
    print("   %s: worst per-tensor cos %.6f, global cos %.6f" % (name, worst, glob))
    return worst, glob


def prebn_bias_keys(net_prefix, sd):
    # conv biases that feed a BatchNorm: blocks.{i}.0.bias where blocks.{i}.1.weight exists
    return [k for k in sd if k.endswith(".0.bias") and k.replace(".0.bias", ".1.weight") in sd]


def run_fixture(path):
    fx = torch.load(path, weights_only=False)
    res, width = fx["res"], fx["width"]
    zd = fx.get("z_dim", 100)
    netG = dcgan.Generator(z_dim=zd, ngf=width, resolution=res).cuda()
    netD = dcgan.Discriminator(ndf=width, resolution=res).cuda()
    netG.load_state_dict(fx["sd_g"])
    netD.load_state_dict(fx["sd_d"])
    crit = GANLoss(fx["mode"], *fx["labels"]).cuda()
    x, z1, z2 = fx["x"].cuda(), fx["z1"].cuda(), fx["z2"].cuda()
    ok = True
    out = netD(x)
    loss = crit(out, True)
    loss.backward()
    print("d_real relerr %.3e  loss %.5f vs %.5f" % (relerr(out.cpu(), fx["d_real"]), loss.item(), fx["loss_real"].item()))
    skipD = prebn_bias_keys("d", fx["sd_d"])
    w, g = cmp_grads("D-real", {k: p.grad for k, p in netD.named_parameters()}, unpack(fx["d_grads_real"]), skipD)
    ok &= g > 0.99
    fake1 = netG(z1)
    print("fake1 relerr %.3e" % relerr(fake1.detach().cpu(), fx["fake1"]))
    netD.zero_grad()
    out = netD(fake1.detach())
    loss = crit(out, False)
    loss.backward()
    print("d_fake relerr %.3e  loss %.5f vs %.5f" % (relerr(out.cpu(), fx["d_fake"]), loss.item(), fx["loss_fake"].item()))
    w, g = cmp_grads("D-fake", {k: p.grad for k, p in netD.named_parameters()}, unpack(fx["d_grads_fake"]), skipD)
    ok &= g > 0.98
    netG.zero_grad(), netD.zero_grad()
    fake2 = netG(z2)
    out = netD(fake2)
    loss = crit(out, False, True)
    loss.backward()
    print("fake2 relerr %.3e d_g relerr %.3e loss %.5f vs %.5f" % (relerr(fake2.detach().cpu(), fx["fake2"]),
          relerr(out.detach().cpu(), fx["d_g"]), loss.item(), fx["loss_g"].item()))
    skipG = prebn_bias_keys("g", fx["sd_g"])
    w, g = cmp_grads("G-step", {k: p.grad for k, p in netG.named_parameters()}, unpack(fx["g_grads"]), skipG)
    ok &= g > 0.95
    # buffers
    for k, v in fx["buf_d_after"].items():
        got = netD.state_dict()[k].cpu()
        print("   buf D %-36s relerr %.3e" % (k, relerr(got.float(), v.float())))
    for k, v in fx["buf_g_after"].items():
        got = netG.state_dict()[k].cpu()
        print("   buf G %-36s relerr %.3e" % (k, relerr(got.float(), v.float())))
    return ok


def run_oracle(res, width, batch):
    from oracle import gan_oracle as O
    torch.manual_seed(0)
    netG = dcgan.Generator(ngf=width, resolution=res)
    netD = dcgan.Discriminator(ndf=width, resolution=res)
    sd_g = {k: v.clone() for k, v in netG.state_dict().items()}
    sd_d = {k: v.clone() for k, v in netD.state_dict().items()}
    gen = torch.Generator().manual_seed(1)
    x = torch.rand(batch, 3, res, res, generator=gen) * 2 - 1
    z1 = torch.randn(batch, 100, generator=gen)
    z2 = torch.randn(batch, 100, generator=gen)
    torch.set_num_threads(os.cpu_count())
    ref = O.dcgan_step_grads(sd_g, sd_d, x, z1, z2)
    netG.cuda(), netD.cuda()
    crit = GANLoss("vanilla", 0.9, 0.1, 0.9).cuda()
    xc, z1c, z2c = x.cuda(), z1.cuda(), z2.cuda()
    ok = True
    out = netD(xc)
    loss = crit(out, True)
    loss.backward()
    print("d_real relerr %.3e loss %.5f vs %.5f" % (relerr(out.cpu(), ref["d_real"]), loss.item(), ref["loss_real"].item()))
    skipD = prebn_bias_keys("d", sd_d)
    w, g = cmp_grads("D-real", {k: p.grad for k, p in netD.named_parameters()}, ref["d_grads_real"], skipD)
    ok &= g > 0.999
    fake1 = netG(z1c)
    print("fake1 relerr %.3e" % relerr(fake1.detach().cpu(), ref["fake1"]))
    netD.zero_grad()
    out = netD(fake1.detach())
    loss = crit(out, False)
    loss.backward()
    print("d_fake relerr %.3e loss %.5f vs %.5f" % (relerr(out.cpu(), ref["d_fake"]), loss.item(), ref["loss_fake"].item()))
    w, g = cmp_grads("D-fake", {k: p.grad for k, p in netD.named_parameters()}, ref["d_grads_fake"], skipD)
    netG.zero_grad(), netD.zero_grad()
    out = netD(netG(z2c))
    loss = crit(out, False, True)
    loss.backward()
    print("d_g relerr %.3e loss %.5f vs %.5f" % (relerr(out.detach().cpu(), ref["d_g"]), loss.item(), ref["loss_g"].item()))
    skipG = prebn_bias_keys("g", sd_g)
    w, g = cmp_grads("G-step", {k: p.grad for k, p in netG.named_parameters()}, ref["g_grads"], skipG)
    return ok


def run_timing(batch, steps=5):
    torch.manual_seed(0)
    netG = dcgan.Generator().cuda()
    netD = dcgan.Discriminator().cuda()
    crit = GANLoss("vanilla", 0.9, 0.1, 0.9).cuda()
    optG = torch.optim.Adam(netG.parameters(), lr=4e-4, betas=(0.5, 0.999))
    optD = torch.optim.Adam(netD.parameters(), lr=1e-4, betas=(0.5, 0.999))
    x = torch.rand(batch, 3, 64, 64, device="cuda") * 2 - 1

    def step():
        optD.zero_grad()
        out = netD(x)
        l1 = crit(out, True)
        l1.backward()
        z = torch.randn(batch, 100, device="cuda")
        fake = netG(z)
        out = netD(fake.detach())
        l2 = crit(out, False)
        l2.backward()
        optD.step()
        optG.zero_grad()
        z = torch.randn(batch, 100, device="cuda")
        out = netD(netG(z))
        l3 = crit(out, False, True)
        l3.backward()
        optG.step()
        return l1, l2, l3

    for _ in range(3):
        ls = step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        ls = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print("B=%d: %.3f ms/step  -> %.1f steps/s, %.0f img/s, %.1f TFLOP/s (minimal-step FLOPs)  losses %s" % (
        batch, ms, 1e3 / ms, batch * 1e3 / ms, 9.7994e9 * batch / ms / 1e9, [round(l.item(), 4) for l in ls]))
    print("peak mem %.2f GB" % (torch.cuda.max_memory_allocated() / 2**30))
    return True


CASES = {
    "golden_r32_w4": lambda: run_fixture(os.path.join(ROOT, "tests/golden/dcgan_r32_w4.pt")),
    "golden_r64_w4": lambda: run_fixture(os.path.join(ROOT, "tests/golden/dcgan_r64_w4.pt")),
    "oracle_r32_w16_b16": lambda: run_oracle(32, 16, 16),
    "oracle_r64_w64_b32": lambda: run_oracle(64, 64, 32),
    "timing_b128": lambda: run_timing(128),
    "timing_b1024": lambda: run_timing(1024),
}

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", default=None)
    ap.add_argument("--timeout", type=int, default=240)
    ap.add_argument("--only", default=None)
    args = ap.parse_args()
    if args.case:
        ok = CASES[args.case]()
        sys.exit(0 if ok else 1)
    results = {}
    for name in CASES:
        if args.only and args.only not in name:
            continue
        try:
            pr = subprocess.run([sys.executable, os.path.abspath(__file__), "--case", name], timeout=args.timeout,
                                stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            out, rc = pr.stdout, pr.returncode
        except subprocess.TimeoutExpired as e:
            out = e.stdout.decode() if isinstance(e.stdout, bytes) else (e.stdout or "")
            rc = "TIMEOUT"
        print("==== %s rc=%s" % (name, rc))
        for l in [l for l in out.splitlines() if l.strip() and not l.startswith("Param count")][-80:]:
            print("   " + l)
        results[name] = rc
        sys.stdout.flush()
    print("SUMMARY:", results)
