#!/usr/bin/env python
"""bench.py — headline benchmark of the hot path: DCGAN-64 G+D adversarial train steps (loop body of the reference's
main_dcgan.py:68-95, nz=100, ngf=ndf=64, GANLoss('vanilla', .9, .1, .9), Adam 4e-4 / 1e-4, betas (0.5, 0.999))
at global batch 1024 on N B200s (batch sharded over ranks; NCCL grad all-reduce + synchronised BatchNorm).

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference's CPU path (oracle port) on the box's host cores

Prints ONE JSON line on rank 0 (contract in the task statement): metric/value (device-resident inputs), e2e (pinned
host inputs + host reads inside the timed region), roofline of the dominant kernel (tcgen05 implicit-GEMM convs),
cpu_baseline, clocks, gpu_launches.
"""
import argparse
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

GLOBAL_BATCH = 1024
Z_DIM = 100
FLOPS_PER_IMG = 9.7994e9  # minimal algorithmic FLOPs of one step per image
METRIC = "dcgan64_train_images_per_sec"
WORKLOAD = ("DCGAN-64 G+D adversarial train step (main_dcgan.py:68-95 loop body incl. Adam steps and the three "
            ".item() reads), nz=100 ngf=ndf=64, synthetic 64x64 images")


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


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


# ----------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's CPU implementation (oracle port; the reference is pure Python on
# torch and cannot travel to the GPU box)
# ----------------------------------------------------------------------------------------------------------------
def cpu_reference(sample_batch, steps, warmup):
    from oracle import gan_oracle as O

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd_g, sd_d = O.init_dcgan_state(seed=0)
    tr = O.CpuDcganTrainer(sd_g, sd_d)
    gen = torch.Generator().manual_seed(1234)
    x = torch.rand(sample_batch, 3, 64, 64, generator=gen) * 2 - 1
    for _ in range(warmup):
        tr.step(x, torch.randn(sample_batch, Z_DIM, generator=gen), torch.randn(sample_batch, Z_DIM, generator=gen))
    t0 = time.perf_counter()
    for _ in range(steps):
        tr.step(x, torch.randn(sample_batch, Z_DIM, generator=gen), torch.randn(sample_batch, Z_DIM, generator=gen))
    dt = (time.perf_counter() - t0) / steps
    return {"value": sample_batch / dt, "unit": "img/s", "cores": cores, "kind": "port",
            "sample": "%d-image slices of the batch-1024 step (same nets, fp32, oracle/gan_oracle.CpuDcganTrainer, "
                      "%d timed steps after %d warm-up), %.2f s per slice-step" % (sample_batch, steps, warmup, dt),
            "ms_per_sample_step": dt * 1e3}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 3))
    warm = max(1, min(args.warmup, 1))
    cb = cpu_reference(64, steps, warm)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "img/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": cb["ms_per_sample_step"] * (GLOBAL_BATCH / 64.0),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "global_batch": GLOBAL_BATCH, "parallelism": "cpu"},
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------------------
# product arm
# ----------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong: global batch 1024 sharded over ranks; weak: 1024 images per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--global-batch", type=int, default=GLOBAL_BATCH,
                    help="diagnostics only: the benchmark configuration is the default 1024")
    ap.add_argument("--torch-adam", action="store_true", help="torch.optim.Adam + flat gradient buckets instead of optim.FusedAdam")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel eagerly instead of one CUDA graph replay per step")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    args.warmup = max(args.warmup, 3)

    from gan_playground_b200 import _lib, config, ops, parallel
    from gan_playground_b200.criterion import GANLoss
    from gan_playground_b200.engine import DcganStep
    from gan_playground_b200.optim import FusedAdam
    from gan_playground_b200.models import dcgan

    rank, world = parallel.init()
    peer_sync = False
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    if world > 1:
        peer_sync = parallel.init_peer_sync(dev)   # SyncBN sums over NVLink peer memory (falls back to NCCL)
    per_gpu = args.global_batch if args.scaling == "weak" else args.global_batch // world
    global_batch = per_gpu * world

    torch.manual_seed(0)
    import contextlib
    import io

    with contextlib.redirect_stdout(io.StringIO()):
        netG = dcgan.Generator().to(dev)
        netD = dcgan.Discriminator().to(dev)
    parallel.broadcast_module(netG)
    parallel.broadcast_module(netD)
    use_graph = not args.no_graph
    if args.torch_adam:   # the optimiser the reference scripts build themselves (main_dcgan.py:55-56)
        optG = torch.optim.Adam(netG.parameters(), lr=4e-4, betas=(0.5, 0.999), capturable=use_graph)
        optD = torch.optim.Adam(netD.parameters(), lr=1e-4, betas=(0.5, 0.999), capturable=use_graph)
    else:                 # same update as one kernel per network over flat buffers (+ ZeRO-1 sharding across ranks)
        optG = FusedAdam(netG.parameters(), lr=4e-4, betas=(0.5, 0.999), shard=world > 1)
        optD = FusedAdam(netD.parameters(), lr=1e-4, betas=(0.5, 0.999), shard=world > 1)
    crit = GANLoss('vanilla', target_real_label=0.9, target_fake_label=0.1, target_fake_G_label=0.9).to(dev)
    netG.train(), netD.train()
    # loop body of main_dcgan.py:68-95 (+ DP gradient all-reduce before each optimiser step), eager or as one CUDA graph
    runner = DcganStep(netG, netD, crit, optG, optD, per_gpu, Z_DIM, dev, use_graph=use_graph)

    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    x_dev = torch.rand(per_gpu, 3, 64, 64, device=dev, generator=gen) * 2 - 1
    x_host = x_dev.cpu().pin_memory()
    z_host = torch.randn(2, per_gpu, Z_DIM).pin_memory()

    def step(inputs):
        return runner.step(inputs)

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
    # Input pipeline: the images of step i+1 travel host->device on a copy stream while step i computes (two device
    # buffers); every timed step issues exactly one H2D copy of a full batch and reads its six scalars back.
    last = {}
    copy_stream = torch.cuda.Stream(device=dev)
    bufs = [torch.empty_like(x_dev), torch.empty_like(x_dev)]
    arrived = [torch.cuda.Event(), torch.cuda.Event()]
    e2e_i = [0]

    def prefetch(slot):
        with torch.cuda.stream(copy_stream):
            bufs[slot].copy_(x_host, non_blocking=True)
            arrived[slot].record(copy_stream)

    def e2e_step():
        slot = e2e_i[0] & 1
        e2e_i[0] += 1
        torch.cuda.current_stream().wait_event(arrived[slot])
        # the other buffer was consumed by the previous step, which has fully completed (step() ends with the host
        # read of its scalars), so the next batch may overwrite it now
        prefetch(slot ^ 1)
        r = step(bufs[slot])
        last["losses"] = tuple(r[:3])

    prefetch(0)
    for _ in range(2):
        e2e_step()
    e2e_ms = timed(e2e_step, args.steps) / args.steps
    copy_stream.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    h2d = x_host.numel() * 4
    d2h = 6 * 4  # three D-output means + three loss scalars

    # ---- roofline of the dominant kernel family: tcgen05 implicit-GEMM convs, timed per launch with CUDA events
    # (one eager step: same kernels as the graph replays, launched one by one so they can be bracketed by events)
    prof = ops.GemmProfiler()
    l0 = _lib.launch_count()
    with prof:
        runner.step_eager(x_dev)
    torch.cuda.synchronize()
    launches = (_lib.launch_count() - l0) * args.steps   # own kernels per step x timed steps
    gemm = prof.summary()
    peaks, peaks_src = load_peaks()
    peak_tf = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"])
    # DRAM bytes moved by the same 53 GEMM launches of one step, from the committed ncu launch list of this command
    # (profiles/ncu_gemm_traffic.json; only valid for the configuration it was captured on)
    precision_label = config.precision() + (" (real-image D pass: bf16)" if runner.real_precision else "") + \
        (" (D-fake chain: %s)" % runner.fake_precision if runner.fake_precision else "")
    traffic, traffic_src = None, None
    try:
        tj = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "ncu_gemm_traffic.json")))
        if world == 1 and tj.get("global_batch") == global_batch and tj.get("precision", "") == precision_label:
            traffic, traffic_src = tj["dram_bytes_per_step_gemm"], tj["source"]
    except (OSError, ValueError, KeyError):
        pass
    roofline = {"bound": "tensor", "achieved": gemm["tflops"], "peak": peak_tf, "unit": "TFLOP/s",
                "frac": gemm["tflops"] / peak_tf, "traffic": traffic, "traffic_unit": "bytes of DRAM traffic per step (all GEMM launches)",
                "traffic_source": traffic_src,
                "kernel": "gp::conv_gemm_kernel<MODE,BN> (all %d launches of one step)" % gemm["launches"],
                "peak_source": "%s bf16_tflops_sustained (kernel timed inside a long step)" % peaks_src,
                "gemm_ms_per_step": gemm["ms"], "gemm_share_of_step": gemm["ms"] / ms_per_step,
                "algorithmic_flops_per_step": gemm["flops"]}

    if rank == 0:
        value = global_batch * 1e3 / ms_per_step
        line = {
            "metric": METRIC, "value": value, "unit": "img/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "global_batch": global_batch, "per_gpu_batch": per_gpu,
                       "parallelism": "dp%d" % world, "precision": precision_label,
                       "cuda_graph": use_graph,
                       "optimizer": "torch.optim.Adam" if args.torch_adam else "FusedAdam(flat%s)" % (", zero1" if world > 1 else ""),
                       "syncbn": "n/a (1 rank)" if world == 1 else ("one-shot NVLink peer exchange fused with the statistics finalize"
                                                                   if peer_sync else "NCCL all-reduce"),
                       "l2": "no flush needed: per-step working set (~3.4 GB of activations at 1024 img/GPU) >> 126 MB L2"},
            "steps_per_sec": 1e3 / ms_per_step,
            "tflops_minimal_step": FLOPS_PER_IMG * global_batch / (ms_per_step * 1e-3) / 1e12,
            "e2e": {"value": global_batch * 1e3 / e2e_ms, "unit": "img/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "last_losses": last.get("losses"),
                    "input_pipeline": "pinned host batch -> device on a copy stream, overlapped with the previous step (2 buffers)"},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline,
        }
        if world == 1 and not args.no_cpu_baseline:
            cb = cpu_reference(64, 2, 1)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line))
    sys.stdout.flush()
    sys.stderr.flush()
    if world > 1:
        # Tearing down an NCCL communicator while CUDA graphs that captured its kernels are alive can block forever
        # (seen at N=8: the JSON line was out, destroy_process_group() never returned). All ranks are past their last
        # collective here, so drain the device and leave without the destructor path.
        torch.cuda.synchronize()
        os._exit(0)
    parallel.shutdown()


if __name__ == "__main__":
    main()
