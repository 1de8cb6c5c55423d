"""Throughput of the other BASELINE.json configurations through their step drivers (one CUDA-graph replay per step,
device-resident synthetic inputs, CUDA-event timing). bench.py stays the headline (config 1: DCGAN-64 batch 1024);
these are the parity-test configurations of SURVEY.md §8(d) timed for context:

  cfg3  SN-DCGAN 32x32 (models/dcgan_specnorm.py), hinge, loop of main_dcgan.py:68-95        -> engine.DcganStep
  cfg4  SNGAN projection 32x32, 10 classes, loop of main_sngan.py:65-100, n_disc_update=1   -> engine.SnganStep
  cfg5  ACGAN 64x64 (models/acgan.py), batch 512, loop of main_acgan.py:84-133              -> engine.AcganStep

    python tools/bench_loops.py [--steps 20] [--warmup 5] [--only cfg5] > gpurun_out/bench_loops.jsonl

One JSON line per configuration; algorithmic FLOPs per image are SURVEY.md §8(d)'s minimal-step figures. A configuration
that fails prints {"config": ..., "error": ...} and the others still run."""
import argparse
import contextlib
import io
import json
import os
import sys
import traceback

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def quiet(fn):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn()


def timed(step, steps, warmup):
    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        last = step()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, last


def cfg3(dev, batch):
    from gan_playground_b200.criterion import GANLoss
    from gan_playground_b200.engine import DcganStep
    from gan_playground_b200.models import dcgan_specnorm as M
    from gan_playground_b200.optim import FusedAdam

    torch.manual_seed(0)
    netG = quiet(lambda: M.Generator(resolution=32)).to(dev)
    netD = quiet(lambda: M.Discriminator(resolution=32)).to(dev)
    oG = FusedAdam(netG.parameters(), lr=2e-4, betas=(0.0, 0.999))
    oD = FusedAdam(netD.parameters(), lr=2e-4, betas=(0.0, 0.999))
    run = DcganStep(netG, netD, GANLoss("hinge").to(dev), oG, oD, batch, 100, dev, use_graph=True)
    x = torch.rand(batch, 3, 32, 32, device=dev) * 2 - 1
    return (lambda: run.step(x)), 1.647e9, "SN-DCGAN 32x32 hinge, main_dcgan loop"


def cfg4(dev, batch):
    from gan_playground_b200.criterion import GANLoss
    from gan_playground_b200.engine import SnganStep
    from gan_playground_b200.models import sngan_projection as M
    from gan_playground_b200.optim import FusedAdam

    torch.manual_seed(0)
    netG = quiet(lambda: M.ResNetGenerator(ch=64, dim_z=128, bottom_width=2, img_dim=3, n_classes=10)).to(dev)
    netD = quiet(lambda: M.SNResNetProjectionDiscriminator(ch=64, n_classes=10, img_dim=3)).to(dev)
    oG = FusedAdam(netG.parameters(), lr=2e-4, betas=(0.0, 0.999))
    oD = FusedAdam(netD.parameters(), lr=2e-4, betas=(0.0, 0.999))
    run = SnganStep(netG, netD, GANLoss("hinge").to(dev), oG, oD, batch, 128, dev, n_classes=10, n_disc_update=1,
                    use_graph=True)
    x = torch.rand(batch, 3, 32, 32, device=dev) * 2 - 1
    y = torch.randint(10, (batch,), device=dev)
    return (lambda: run.step(x, y)), 5.506e9, "SNGAN projection 32x32, 10 classes, main_sngan loop (n_disc_update=1)"


def cfg5(dev, batch):
    from gan_playground_b200.criterion import GANLoss
    from gan_playground_b200.engine import AcganStep
    from gan_playground_b200.models import acgan as M
    from gan_playground_b200.optim import FusedAdam

    torch.manual_seed(0)
    netG = quiet(lambda: M.Generator(z_dim=100, ngf=64, n_class=10)).to(dev)
    netD = quiet(lambda: M.Discriminator(ndf=64, n_class=10)).to(dev)
    oG = FusedAdam(netG.parameters(), lr=4e-4, betas=(0.5, 0.999))
    oD = FusedAdam(netD.parameters(), lr=1e-4, betas=(0.5, 0.999))
    run = AcganStep(netG, netD, GANLoss("vanilla", 0.9, 0.1, 0.9).to(dev), oG, oD, batch, 100, dev, use_graph=True,
                    mixed_precision=MIXED["acgan"])
    x = torch.rand(batch, 3, 64, 64, device=dev) * 2 - 1
    y = torch.randint(0, 2, (batch, 10), device=dev).float()
    return (lambda: run.step(x, y)), 8.979e9, "ACGAN 64x64 two-head, main_acgan loop"


MIXED = {"acgan": False}      # --acgan-mixed: AcganStep's opt-in per-pass precision policy (real bf16 / D-fake fp16)
CONFIGS = {"cfg3": (cfg3, 512), "cfg4": (cfg4, 256), "cfg5": (cfg5, 512)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--only", default="")
    ap.add_argument("--batch", type=int, default=0, help="override the per-configuration batch")
    ap.add_argument("--acgan-mixed", action="store_true", help="cfg5 with AcganStep(mixed_precision=True)")
    args = ap.parse_args()
    MIXED["acgan"] = args.acgan_mixed
    from gan_playground_b200 import _lib, config

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    for name, (make, batch) in CONFIGS.items():
        if args.only and name not in args.only.split(","):
            continue
        batch = args.batch or batch
        try:
            step, flops_img, what = make(dev, batch)
            l0 = _lib.launch_count()
            ms, last = timed(step, args.steps, max(args.warmup, 3))
            print(json.dumps({"config": name, "workload": what, "batch": batch, "ms_per_step": ms,
                              "images_per_sec": batch / ms * 1e3, "steps_per_sec": 1e3 / ms,
                              "tflops_minimal_step": flops_img * batch / ms * 1e-9,
                              "precision": config.precision() + (" (mixed per-pass policy)" if name == "cfg5" and args.acgan_mixed else ""),
                              "cuda_graph": True, "steps": args.steps, "gpu_launches_incl_capture": _lib.launch_count() - l0,
                              "last_logged": last}), flush=True)
        except Exception as e:  # keep going: the other configurations are independent
            print(json.dumps({"config": name, "error": "%s: %s" % (type(e).__name__, e),
                              "trace": traceback.format_exc().splitlines()[-6:]}), flush=True)
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
