#!/usr/bin/env python
"""bench.py — benchmark of the hot path: one adversarial G+D train step through the step drivers of engine.py.

Headline (default, `--config cfg2`): DCGAN-64 (loop body of the reference's main_dcgan.py:68-95, nz=100, ngf=ndf=64,
GANLoss('vanilla', .9, .1, .9), Adam 4e-4 / 1e-4, betas (0.5, 0.999)) at global batch 1024 on N B200s (batch sharded
over ranks; gradient exchange over NCCL + synchronised BatchNorm over NVLink peer memory). The other BASELINE.json
configurations run under the same JSON contract:

    --config cfg3   SN-DCGAN 32x32 hinge (models/dcgan_specnorm.py), main_dcgan.py loop, global batch 1024
    --config cfg4   SNGAN projection 32x32, 10 classes (models/sngan_projection.py), main_sngan.py:65-100 loop
                    (n_disc_update=1), global batch 1024
    --config cfg5   ACGAN 64x64 two-head (models/acgan.py), main_acgan.py:84-133 loop, 512 images PER GPU (weak scaling)

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...   # the reference's own CPU path on the box's host cores: the UNMODIFIED reference
                                           # modules from baseline/_ref when present (kind "reference"), else the oracle port

Prints ONE JSON line on rank 0 (contract in the task statement): metric/value (device-resident inputs), e2e (pinned
host inputs + host reads inside the timed region), roofline of the dominant kernel (tcgen05 implicit-GEMM convs) against
BOTH measured peaks, cpu_baseline, clocks, gpu_launches.
"""
import argparse
import contextlib
import gc
import io
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

REF_DIR = os.path.join(ROOT, "baseline", "_ref")

# minimal algorithmic FLOPs of one step per image: SURVEY.md §8(d)
CONFIGS = {
    "cfg2": dict(metric="dcgan64_train_images_per_sec", flops_img=9.7994e9, batch=1024, scaling="strong", res=64, z=100,
                 workload="DCGAN-64 G+D adversarial train step (main_dcgan.py:68-95 loop body incl. Adam steps and the "
                          "three .item() reads), nz=100 ngf=ndf=64, synthetic 64x64 images"),
    "cfg3": dict(metric="sn_dcgan32_train_images_per_sec", flops_img=1.647e9, batch=1024, scaling="strong", res=32, z=100,
                 workload="SN-DCGAN 32x32 (models/dcgan_specnorm.py, spectral-norm power iteration per forward) hinge loss, "
                          "main_dcgan.py:68-95 loop body incl. Adam(2e-4, betas (0, 0.999)), width 64, synthetic 32x32 images"),
    "cfg4": dict(metric="sngan_projection32_train_images_per_sec", flops_img=5.506e9, batch=1024, scaling="strong", res=32,
                 z=128,
                 workload="SNGAN projection 32x32 class-conditional, 10 classes (models/sngan_projection.py, ch=64, "
                          "bottom_width=2), hinge, main_sngan.py:65-100 loop body with n_disc_update=1 incl. Adam steps"),
    "cfg5": dict(metric="acgan64_train_images_per_sec", flops_img=8.979e9, batch=512, scaling="weak", res=64, z=100,
                 workload="ACGAN 64x64 two-head (models/acgan.py, width 64, 10 attributes), main_acgan.py:84-133 loop body "
                          "(adversarial + 0.5 x MSE auxiliary objective) incl. Adam steps, 512 images per GPU"),
}


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return ({"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0},
                "fallback (B200_PROFILING.md)")


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])), mx.append(float(f[1])), pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def quiet(fn):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn()


# ----------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's own CPU implementation of the step on the box's host cores.
#   kind "reference": the UNMODIFIED reference modules (models/*.py, utils/criterion.py) imported from baseline/_ref — a
#       git-ignored copy of /root/reference made by oracle/install_ref.py, which travels to the GPU box — driven by the
#       loop body of the matching main_*.py (the scripts bury it inside main() behind a dataset download and a CLI);
#   kind "port": oracle/gan_oracle.py's restatement of the same loops, when baseline/_ref is absent.
# None of this repo's models, kernels or engine is on this path.
# ----------------------------------------------------------------------------------------------------------------
def _synthetic_cpu(cfg_name, batch, gen):
    c = CONFIGS[cfg_name]
    x = torch.rand(batch, 3, c["res"], c["res"], generator=gen) * 2 - 1
    if cfg_name == "cfg4":
        return x, torch.randint(10, (batch,), generator=gen)
    if cfg_name == "cfg5":
        return x, torch.randint(0, 2, (batch, 10), generator=gen).float()
    return (x,)


def _reference_stepper(cfg_name):
    """step(inputs, gen) running the unmodified reference modules on the CPU in fp32, following the script's loop."""
    import importlib

    sys.path.insert(0, REF_DIR)
    try:
        for m in [k for k in sys.modules if k == "models" or k.startswith("models.") or k == "utils" or k.startswith("utils.")]:
            del sys.modules[m]
        GANLoss = importlib.import_module("utils.criterion").GANLoss
        c = CONFIGS[cfg_name]
        torch.manual_seed(0)
        if cfg_name in ("cfg2", "cfg3"):
            M = importlib.import_module("models.dcgan" if cfg_name == "cfg2" else "models.dcgan_specnorm")
            netG = quiet(lambda: M.Generator(z_dim=c["z"], resolution=c["res"]))
            netD = quiet(lambda: M.Discriminator(resolution=c["res"]))
            if cfg_name == "cfg2":
                crit, lrg, lrd, betas = GANLoss('vanilla', 0.9, 0.1, 0.9), 4e-4, 1e-4, (0.5, 0.999)
            else:
                crit, lrg, lrd, betas = GANLoss('hinge'), 2e-4, 2e-4, (0.0, 0.999)
        elif cfg_name == "cfg4":
            M = importlib.import_module("models.sngan_projection")
            netG = M.ResNetGenerator(ch=64, dim_z=128, bottom_width=2, img_dim=3, n_classes=10)
            netD = M.SNResNetProjectionDiscriminator(ch=64, n_classes=10, img_dim=3)
            crit, lrg, lrd, betas = GANLoss('hinge'), 2e-4, 2e-4, (0.0, 0.999)
        else:
            M = importlib.import_module("models.acgan")
            netG = quiet(lambda: M.Generator(z_dim=100, ngf=64, n_class=10))
            netD = quiet(lambda: M.Discriminator(ndf=64, n_class=10))
            crit, lrg, lrd, betas = GANLoss('vanilla', 0.9, 0.1, 0.9), 4e-4, 1e-4, (0.5, 0.999)
    finally:
        sys.path.remove(REF_DIR)
    optG = torch.optim.Adam(netG.parameters(), lr=lrg, betas=betas)
    optD = torch.optim.Adam(netD.parameters(), lr=lrd, betas=betas)
    netG.train(), netD.train()
    mse = torch.nn.MSELoss()

    def step(inputs, gen):
        B = inputs[0].shape[0]
        if cfg_name in ("cfg2", "cfg3"):          # main_dcgan.py:68-95
            (x,) = inputs
            optD.zero_grad()
            outD = netD(x)
            outD.mean().item()
            criterion_real = crit(outD, True)
            criterion_real.backward()
            outG = netG(torch.randn(B, c["z"], generator=gen))
            outD = netD(outG.detach())
            outD.mean().item()
            crit(outD, False).backward()
            optD.step()
            optG.zero_grad()
            outD = netD(netG(torch.randn(B, c["z"], generator=gen)))
            outD.mean().item()
            lossG = crit(outD, False, True)
            lossG.backward()
            optG.step()
            return lossG.item()
        if cfg_name == "cfg4":                    # main_sngan.py:72-99 with n_disc_update = 1
            x, y = inputs
            optD.zero_grad()
            outD = netD(x, y)
            outD.mean().item()
            crit(outD, True).backward()
            z, cl = torch.randn(B, c["z"], generator=gen), torch.randint(10, (B,), generator=gen)
            outG = netG(z, cl)
            outD = netD(outG.detach(), cl)
            outD.mean().item()
            crit(outD, False).backward()
            optD.step()
            optG.zero_grad()
            outD = netD(outG, cl)
            outD.mean().item()
            lossG = crit(outD, False, True)
            lossG.backward()
            optG.step()
            return lossG.item()
        x, y = inputs                             # main_acgan.py:90-131
        optD.zero_grad()
        adv, cls = netD(x)
        torch.sigmoid(adv).mean().item()
        (crit(adv, True) + mse(cls, y) * 0.5).backward()
        outG = netG(torch.randn(B, c["z"], generator=gen), y)
        adv, cls = netD(outG.detach())
        torch.sigmoid(adv).mean().item()
        (crit(adv, False) + mse(cls, y) * 0.5).backward()
        optD.step()
        optG.zero_grad()
        adv, cls = netD(outG)
        torch.sigmoid(adv).mean().item()
        lossG = crit(adv, False, True) + mse(cls, y) * 0.5
        lossG.backward()
        optG.step()
        return lossG.item()

    return step


def _port_stepper(cfg_name):
    """The same loops through oracle/gan_oracle.py's trainers (used when baseline/_ref is absent)."""
    from oracle import gan_oracle as O

    c = CONFIGS[cfg_name]
    if cfg_name == "cfg2":
        tr = O.CpuDcganTrainer(*O.init_dcgan_state(seed=0))
        return lambda inputs, gen: tr.step(inputs[0], torch.randn(inputs[0].shape[0], c["z"], generator=gen),
                                           torch.randn(inputs[0].shape[0], c["z"], generator=gen))[2]
    # initial weights only (never timed): state dicts in the reference's layout from the parameter-holder mirrors
    torch.manual_seed(0)
    if cfg_name == "cfg3":
        raise RuntimeError("cfg3 has no oracle-port trainer: run oracle/install_ref.py so baseline/_ref exists")
    if cfg_name == "cfg4":
        from gan_playground_b200.models import sngan_projection as M

        netG = M.ResNetGenerator(ch=64, dim_z=128, bottom_width=2, img_dim=3, n_classes=10)
        netD = M.SNResNetProjectionDiscriminator(ch=64, n_classes=10, img_dim=3)
        tr = O.CpuSnganTrainer(netG.state_dict(), netD.state_dict(), n_disc_update=1, bottom_width=2)
        return lambda inputs, gen: tr.step(inputs[0], inputs[1], torch.randn(inputs[0].shape[0], c["z"], generator=gen),
                                           torch.randint(10, (inputs[0].shape[0],), generator=gen))[2]
    from gan_playground_b200.models import acgan as M

    netG, netD = quiet(lambda: M.Generator(n_class=10)), quiet(lambda: M.Discriminator(n_class=10))
    tr = O.CpuAcganTrainer(netG.state_dict(), netD.state_dict())
    return lambda inputs, gen: tr.step(inputs[0], inputs[1], torch.randn(inputs[0].shape[0], c["z"], generator=gen))[2]


def cpu_reference(cfg_name, batch, steps, warmup):
    """Times `steps` whole steps of the reference's CPU path at `batch` images on all host cores."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    have_ref = os.path.exists(os.path.join(REF_DIR, "models", "dcgan.py"))
    step = _reference_stepper(cfg_name) if have_ref else _port_stepper(cfg_name)
    gen = torch.Generator().manual_seed(1234)
    inputs = _synthetic_cpu(cfg_name, batch, gen)
    for _ in range(warmup):
        step(inputs, gen)
    t0 = time.perf_counter()
    for _ in range(steps):
        last = step(inputs, gen)
    dt = (time.perf_counter() - t0) / steps
    assert last == last, "reference step produced NaN"
    what = ("unmodified reference modules (baseline/_ref) driven by the loop body of the matching main_*.py"
            if have_ref else "oracle/gan_oracle.py port of the reference loop")
    return {"value": batch / dt, "unit": "img/s", "cores": cores, "kind": "reference" if have_ref else "port",
            "sample": "%d timed step(s) of the full %d-image step after %d warm-up, fp32 on %d host threads, %s; %.2f s per "
                      "step" % (steps, batch, warmup, cores, what, dt),
            "ms_per_step": dt * 1e3}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    c = CONFIGS[args.config]
    world = max(1, args.gpus)
    batch = args.global_batch or (c["batch"] * world if c["scaling"] == "weak" else c["batch"])
    steps = max(1, min(args.steps, 3))       # a full-batch CPU step takes seconds: bounded so the run ends within minutes
    warm = max(1, min(args.warmup, 1))
    cb = cpu_reference(args.config, batch, steps, warm)
    line = {"impl": "reference", "metric": c["metric"], "value": cb["value"], "unit": "img/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": cb["ms_per_step"],
            "higher_is_better": True, "scaling": c["scaling"], "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": c["workload"], "global_batch": batch, "parallelism": "cpu x%d threads" % cb["cores"],
                       "name": args.config, "steps_note": "steps / warmup clamped to <= 3 / 1 (seconds per CPU step)"},
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------------------
# product arm: nets + step driver per configuration
# ----------------------------------------------------------------------------------------------------------------
def build_config(name, dev, per_gpu, world, use_graph, torch_adam):
    """Returns (runner, device inputs of one step). Synthetic inputs per SURVEY.md §8(d)."""
    from gan_playground_b200 import parallel
    from gan_playground_b200.criterion import GANLoss
    from gan_playground_b200.engine import AcganStep, DcganStep, SnganStep
    from gan_playground_b200.optim import FusedAdam

    c = CONFIGS[name]
    torch.manual_seed(0)
    if name == "cfg2":
        from gan_playground_b200.models import dcgan as M

        netG, netD = quiet(lambda: M.Generator()).to(dev), quiet(lambda: M.Discriminator()).to(dev)
        lrg, lrd, betas, crit = 4e-4, 1e-4, (0.5, 0.999), GANLoss('vanilla', 0.9, 0.1, 0.9)
    elif name == "cfg3":
        from gan_playground_b200.models import dcgan_specnorm as M

        netG = quiet(lambda: M.Generator(resolution=32)).to(dev)
        netD = quiet(lambda: M.Discriminator(resolution=32)).to(dev)
        lrg, lrd, betas, crit = 2e-4, 2e-4, (0.0, 0.999), GANLoss('hinge')
    elif name == "cfg4":
        from gan_playground_b200.models import sngan_projection as M

        netG = M.ResNetGenerator(ch=64, dim_z=128, bottom_width=2, img_dim=3, n_classes=10).to(dev)
        netD = M.SNResNetProjectionDiscriminator(ch=64, n_classes=10, img_dim=3).to(dev)
        lrg, lrd, betas, crit = 2e-4, 2e-4, (0.0, 0.999), GANLoss('hinge')
    else:
        from gan_playground_b200.models import acgan as M

        netG = quiet(lambda: M.Generator(z_dim=100, ngf=64, n_class=10)).to(dev)
        netD = quiet(lambda: M.Discriminator(ndf=64, n_class=10)).to(dev)
        lrg, lrd, betas, crit = 4e-4, 1e-4, (0.5, 0.999), GANLoss('vanilla', 0.9, 0.1, 0.9)
    parallel.broadcast_module(netG)
    parallel.broadcast_module(netD)
    if torch_adam:    # the optimiser the reference scripts build themselves (main_dcgan.py:55-56)
        optG = torch.optim.Adam(netG.parameters(), lr=lrg, betas=betas, capturable=use_graph)
        optD = torch.optim.Adam(netD.parameters(), lr=lrd, betas=betas, capturable=use_graph)
    else:             # same update as one kernel per network over flat buffers (+ ZeRO-1 sharding across ranks)
        optG = FusedAdam(netG.parameters(), lr=lrg, betas=betas, shard=world > 1)
        optD = FusedAdam(netD.parameters(), lr=lrd, betas=betas, shard=world > 1)
    crit = crit.to(dev)
    netG.train(), netD.train()
    rank = parallel.rank()
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    x = torch.rand(per_gpu, 3, c["res"], c["res"], device=dev, generator=gen) * 2 - 1
    if name in ("cfg2", "cfg3"):
        runner = DcganStep(netG, netD, crit, optG, optD, per_gpu, c["z"], dev, use_graph=use_graph)
        inputs = [x]
    elif name == "cfg4":
        runner = SnganStep(netG, netD, crit, optG, optD, per_gpu, c["z"], dev, n_classes=10, n_disc_update=1,
                           use_graph=use_graph)
        inputs = [x, torch.randint(10, (per_gpu,), device=dev, generator=gen)]
    else:
        runner = AcganStep(netG, netD, crit, optG, optD, per_gpu, c["z"], dev, use_graph=use_graph)
        inputs = [x, torch.randint(0, 2, (per_gpu, 10), device=dev, generator=gen).float()]
    return runner, inputs


def precision_label(runner):
    from gan_playground_b200 import config

    label = config.precision()
    real, fake = getattr(runner, "real_precision", None), getattr(runner, "fake_precision", None)
    if type(runner).__name__ == "SnganStep" and label == "bf16x3":
        label = "fp16 (single-MMA fp16 operands: the ResNet nodes' mapping of the bf16x3 default)"
    if real:
        label += " (real-image D pass: %s)" % real
    if fake:
        label += " (D-fake chain: %s)" % fake
    return label


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="cfg2", choices=sorted(CONFIGS),
                    help="BASELINE.json configuration (default cfg2 = the headline DCGAN-64 batch-1024 step)")
    ap.add_argument("--scaling", default=None, choices=["strong", "weak"],
                    help="strong: the global batch sharded over ranks; weak: that batch per GPU (default: per config)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--global-batch", type=int, default=0, help="diagnostics only: override the configuration's batch")
    ap.add_argument("--torch-adam", action="store_true", help="torch.optim.Adam + flat gradient buckets instead of optim.FusedAdam")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel eagerly instead of one CUDA graph replay per step")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    args.warmup = max(args.warmup, 3)
    cfg = CONFIGS[args.config]
    scaling = args.scaling or cfg["scaling"]

    from gan_playground_b200 import _lib, ops, parallel

    rank, world = parallel.init()
    peer_sync = False
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    if world > 1:
        peer_sync = parallel.init_peer_sync(dev)   # SyncBN sums over NVLink peer memory (falls back to NCCL)
    base_batch = args.global_batch or cfg["batch"]
    per_gpu = base_batch if scaling == "weak" else base_batch // world
    global_batch = per_gpu * world
    use_graph = not args.no_graph
    # loop body of the script (+ DP gradient exchange before each optimiser step), eager or as one CUDA graph per step
    runner, x_dev = build_config(args.config, dev, per_gpu, world, use_graph, args.torch_adam)
    x_host = [t.cpu().pin_memory() for t in x_dev]

    def step(inputs):
        return runner.step(*inputs)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
        return ms.item()

    # clocks are sampled from the first warm-up step to the end of the timed region (nvidia-smi needs ~0.5 s to start;
    # warm-up and timed steps are the same continuous load)
    sampler = ClockSampler(dev.index or 0)
    if rank == 0:
        sampler.start()
        time.sleep(0.5)
    for _ in range(args.warmup):
        step(x_dev)
    total_ms = timed(lambda: step(x_dev), args.steps)
    ms_per_step = total_ms / args.steps

    # ---- end-to-end: inputs start in pinned host memory, results are read back on the host, every step.
    # Input pipeline: the inputs of step i+1 travel host->device on a copy stream while step i computes (two device
    # buffers); every timed step issues exactly one H2D copy of a full batch and reads its logged scalars back.
    last = {}
    copy_stream = torch.cuda.Stream(device=dev)
    bufs = [[torch.empty_like(t) for t in x_dev], [torch.empty_like(t) for t in x_dev]]
    arrived = [torch.cuda.Event(), torch.cuda.Event()]
    e2e_i = [0]

    def prefetch(slot):
        with torch.cuda.stream(copy_stream):
            for d, h in zip(bufs[slot], x_host):
                d.copy_(h, non_blocking=True)
            arrived[slot].record(copy_stream)

    def e2e_step():
        slot = e2e_i[0] & 1
        e2e_i[0] += 1
        torch.cuda.current_stream().wait_event(arrived[slot])
        # the other buffer was consumed by the previous step, which has fully completed (step() ends with the host
        # read of its scalars), so the next batch may overwrite it now
        prefetch(slot ^ 1)
        r = step(bufs[slot])
        last["logged"] = tuple(r)

    prefetch(0)
    for _ in range(2):
        e2e_step()
    e2e_ms = timed(e2e_step, args.steps) / args.steps
    copy_stream.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    h2d = sum(t.numel() * t.element_size() for t in x_host)
    d2h = runner.N_SCALARS * 4  # the script's logged scalars (losses, D(x) / D(G(z)) means), one read per step

    # ---- roofline of the dominant kernel family: tcgen05 implicit-GEMM convs, timed per launch with CUDA events
    # (one eager step: same kernels as the graph replays, launched one by one so they can be bracketed by events)
    runner.step_eager(*x_dev)       # un-timed: the eager path allocates outside the graph's private pool the first time
    torch.cuda.synchronize()
    prof = ops.GemmProfiler()
    l0 = _lib.launch_count()
    with prof:
        runner.step_eager(*x_dev)
    torch.cuda.synchronize()
    own_per_step = _lib.launch_count() - l0
    launches = own_per_step * args.steps   # own kernels per step x timed steps
    gemm = prof.summary()
    edge = prof.summary(hbm=True)   # fused image-edge launches (csrc/image_edge.cu): HBM-bound, accounted in bytes
    peaks, peaks_src = load_peaks()
    burst, sustained = peaks["bf16_tflops"], peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"])
    label = precision_label(runner)
    # DRAM bytes moved by the GEMM launches of one step, from the committed ncu launch list of this command
    # (profiles/ncu_gemm_traffic.json, written by tools/ncu_traffic.py; only valid for the configuration and precision
    # policy it was captured on — anything else reports null)
    traffic, traffic_src = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_gemm_traffic.json")))
        tj = tj.get(args.config, tj if "precision" in tj else {})
        if world == 1 and tj.get("global_batch") == global_batch and tj.get("precision", "") == label:
            traffic, traffic_src = tj["dram_bytes_per_step_gemm"], tj["source"]
    except (OSError, ValueError, KeyError):
        pass
    roofline = {"bound": "tensor", "achieved": gemm["tflops"], "peak": burst, "unit": "TFLOP/s",
                "frac": gemm["tflops"] / burst,
                "peak_sustained": sustained, "frac_sustained": gemm["tflops"] / sustained,
                "traffic": traffic, "traffic_unit": "bytes of DRAM traffic per step (all GEMM launches)",
                "traffic_source": traffic_src,
                "kernel": "gp::conv_gemm_kernel<MODE,BN,MT> (all %d launches of one step)" % gemm["launches"],
                "peak_source": "%s: frac = vs bf16_tflops (burst, the strict denominator); frac_sustained = vs "
                               "bf16_tflops_sustained (each launch is bracketed by CUDA events inside one eager step that "
                               "follows %d graph-replayed steps, i.e. at sustained clocks)" % (peaks_src, args.steps * 2),
                "gemm_ms_per_step": gemm["ms"], "gemm_share_of_step": gemm["ms"] / ms_per_step,
                "algorithmic_flops_per_step": gemm["flops"],
                "image_edge": None if edge["launches"] == 0 else {
                    "kernel": "gp::image_conv_fwd / image_conv_wgrad / image_convt_fwd kernels (fp32 NCHW image <-> first / last "
                              "NHWC activation without a column buffer; %d launches of one step)" % edge["launches"],
                    "bound": "hbm", "achieved": edge["gbs"], "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": edge["gbs"] / peaks["hbm_gbs"], "ms_per_step": edge["ms"],
                    "algorithmic_bytes_per_step": edge["bytes"], "algorithmic_flops_per_step": edge["flops"]},
                "whole_step_tflops": cfg["flops_img"] * per_gpu / (ms_per_step * 1e-3) / 1e12,
                "whole_step_frac": cfg["flops_img"] * per_gpu / (ms_per_step * 1e-3) / 1e12 / burst}

    line = None
    if rank == 0:
        value = global_batch * 1e3 / ms_per_step
        line = {
            "metric": cfg["metric"], "value": value, "unit": "img/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": scaling,
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": cfg["workload"], "name": args.config, "global_batch": global_batch,
                       "per_gpu_batch": per_gpu, "parallelism": "dp%d" % world, "precision": label,
                       "cuda_graph": use_graph,
                       "small_shard_overlap": {"wgrad_side_stream": getattr(runner, "_wg_stream", None) is not None,
                                               "g_forward_next_to_real_pass": getattr(runner, "_g_stream", None) is not None},
                       "optimizer": "torch.optim.Adam" if args.torch_adam else "FusedAdam(flat%s)" % (", zero1" if world > 1 else ""),
                       "syncbn": "n/a (1 rank)" if world == 1 else ("one-shot NVLink peer exchange fused with the statistics finalize"
                                                                   if peer_sync else "NCCL all-reduce"),
                       "l2": "no flush needed: the per-step working set (GBs of activations) >> 126 MB L2"},
            "steps_per_sec": 1e3 / ms_per_step,
            "tflops_minimal_step": cfg["flops_img"] * global_batch / (ms_per_step * 1e-3) / 1e12,
            "e2e": {"value": global_batch * 1e3 / e2e_ms, "unit": "img/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "last_logged": last.get("logged"),
                    "input_pipeline": "pinned host batch -> device on a copy stream, overlapped with the previous step (2 buffers)"},
            "gpu_launches": launches, "gpu_launches_per_step": own_per_step, "clocks": clocks, "roofline": roofline,
        }
    # ---- teardown BEFORE the CPU baseline: graphs that captured NCCL kernels must die before their communicator
    # (destroy_process_group() hung at N=8 with live graphs), and the other ranks must not spin while rank 0 uses the cores
    sys.stdout.flush()
    torch.cuda.synchronize()
    def emit():
        if rank == 0:
            print(json.dumps(line))
            sys.stdout.flush()

    if world > 1:
        # a hung NCCL teardown must not hang the bench: after 60 s print what there is (without the CPU baseline) and leave
        watchdog = threading.Timer(60.0, lambda: (line is not None and line.update(teardown="watchdog"), emit(), os._exit(0)))
        watchdog.daemon = True
        watchdog.start()
        runner.graphs.clear()
        del runner, step
        gc.collect()
        torch.cuda.synchronize()
        torch.distributed.barrier()
        parallel.shutdown()
        watchdog.cancel()
    if rank == 0 and not args.no_cpu_baseline:
        # bounded sample: one timed step of (at most 1024 images of) the same step on the host cores
        cb = cpu_reference(args.config, min(global_batch, 1024), 1, 1)
        line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
    emit()


if __name__ == "__main__":
    main()
