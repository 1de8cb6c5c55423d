"""Library bar: the same DCGAN-64 training step written with stock torch.nn layers (cuDNN / cuBLAS / ATen kernels — what
the reference's `python main_dcgan.py` executes on a CUDA device), timed on the same B200 for context.

Not part of the product and not used by bench.py. The networks below restate the reference's architecture
(models/dcgan.py:21-57,94-124) with torch.nn modules; the loop body is main_dcgan.py:68-95 (Adam 4e-4 / 1e-4,
betas (0.5, 0.999), GANLoss('vanilla', 0.9, 0.1, 0.9) == BCE-with-logits against soft labels).

    python tools/torch_library_bar.py [--batch 1024] [--steps 10]
Modes: tf32  = fp32 modules, cudnn.benchmark=True, TF32 allowed (the reference's own CUDA configuration)
       bf16  = torch.autocast(bfloat16) + channels_last (the usual "fast" stock configuration)
       fp32  = TF32 disabled (true fp32, the numerics the parity tests compare against)"""
import argparse
import json

import torch
import torch.nn as nn
import torch.nn.functional as F


class G(nn.Module):
    def __init__(self, z=100, ngf=64):
        super().__init__()
        c = [ngf * 16, ngf * 8, ngf * 4, ngf * 2]
        self.linear = nn.Linear(z, c[0] * 16)
        self.blocks = nn.ModuleList(nn.Sequential(nn.ConvTranspose2d(i, o, 4, 2, 1), nn.BatchNorm2d(o), nn.ReLU(True))
                                    for i, o in zip(c[:-1], c[1:]))
        self.out_layer = nn.Sequential(nn.ConvTranspose2d(c[-1], 3, 4, 2, 1), nn.Tanh())

    def forward(self, z):
        h = F.relu(self.linear(z), True).view(z.size(0), -1, 4, 4)
        for b in self.blocks:
            h = b(h)
        return self.out_layer(h)


class D(nn.Module):
    def __init__(self, ndf=64):
        super().__init__()
        cin, cout = [3, ndf * 2, ndf * 4, ndf * 8], [ndf * 2, ndf * 4, ndf * 8, ndf * 16]
        self.blocks = nn.ModuleList()
        for k, (i, o) in enumerate(zip(cin, cout)):
            layers = [nn.Conv2d(i, o, 4, 2, 1)] + ([nn.BatchNorm2d(o)] if k else []) + [nn.LeakyReLU(0.2, True)]
            self.blocks.append(nn.Sequential(*layers))
        self.out_layer = nn.Linear(cout[-1], 1)

    def forward(self, x):
        for b in self.blocks:
            x = b(x)
        return self.out_layer(x.sum(dim=[2, 3]))


def run(mode, batch, steps, warmup=3):
    torch.backends.cudnn.benchmark = True
    torch.backends.cudnn.allow_tf32 = mode != "fp32"
    torch.backends.cuda.matmul.allow_tf32 = mode != "fp32"
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    netG, netD = G().to(dev), D().to(dev)
    if mode == "bf16":
        netG, netD = netG.to(memory_format=torch.channels_last), netD.to(memory_format=torch.channels_last)
    optG = torch.optim.Adam(netG.parameters(), lr=4e-4, betas=(0.5, 0.999))
    optD = torch.optim.Adam(netD.parameters(), lr=1e-4, betas=(0.5, 0.999))
    bce = nn.BCEWithLogitsLoss()
    x = torch.rand(batch, 3, 64, 64, device=dev) * 2 - 1
    if mode == "bf16":
        x = x.contiguous(memory_format=torch.channels_last)
    ctx = torch.autocast("cuda", dtype=torch.bfloat16) if mode == "bf16" else torch.autocast("cuda", enabled=False)

    def step():
        optD.zero_grad()
        with ctx:
            out = netD(x)
            l1 = bce(out.float(), torch.full_like(out, 0.9, dtype=torch.float32))
        dx = out.mean().item()
        l1.backward()
        z = torch.randn(batch, 100, device=dev)
        with ctx:
            fake = netG(z)
            out = netD(fake.detach())
            l2 = bce(out.float(), torch.full_like(out, 0.1, dtype=torch.float32))
        d1 = out.mean().item()
        l2.backward()
        optD.step()
        optG.zero_grad()
        z = torch.randn(batch, 100, device=dev)
        with ctx:
            out = netD(netG(z))
            l3 = bce(out.float(), torch.full_like(out, 0.9, dtype=torch.float32))
        d2 = out.mean().item()
        l3.backward()
        optG.step()
        return dx + d1 + d2

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"mode": mode, "batch": batch, "ms_per_step": ms, "img_per_s": batch * 1e3 / ms}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--modes", default="tf32,bf16,fp32")
    a = ap.parse_args()
    for m in a.modes.split(","):
        print(json.dumps(run(m, a.batch, a.steps)))
