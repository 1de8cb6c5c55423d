"""Multi-GPU equivalence check (run under torchrun with N ranks): N ranks on B/N samples each must reproduce the
single-device step at global batch B (SyncBN statistics + averaged gradients), up to bf16 rounding.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dp_check.py
"""
import contextlib
import io
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

from gan_playground_b200 import parallel
from gan_playground_b200.criterion import GANLoss
from gan_playground_b200.models import dcgan


def step(netG, netD, crit, x, z1, z2, sync):
    bD, bG = parallel.GradBucket(netD), parallel.GradBucket(netG)
    bD.attach()
    l1 = crit(netD(x), True)
    l1.backward()
    l2 = crit(netD(netG(z1).detach()), False)
    l2.backward()
    if sync:
        bD.all_reduce_mean()
    gD = {k: p.grad.clone() for k, p in netD.named_parameters()}
    bG.attach()
    netD.zero_grad()
    l3 = crit(netD(netG(z2)), False, True)
    l3.backward()
    if sync:
        bG.all_reduce_mean()
    gG = {k: p.grad.clone() for k, p in netG.named_parameters()}
    return (l1.item(), l2.item(), l3.item()), gD, gG


def cos(a, b):
    num = sum((a[k].double() * b[k].double()).sum() for k in a)
    da = sum((a[k].double() ** 2).sum() for k in a) ** 0.5
    db = sum((b[k].double() ** 2).sum() for k in a) ** 0.5
    return (num / (da * db)).item()


def main():
    rank, world = parallel.init()
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    peer = parallel.init_peer_sync(dev)
    if rank == 0:
        print("SyncBN exchange: %s" % ("one-shot NVLink peer kernel" if peer else "NCCL all-reduce"))
    B = 64
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        netG, netD = dcgan.Generator(ngf=32).to(dev), dcgan.Discriminator(ndf=32).to(dev)
    parallel.broadcast_module(netG), parallel.broadcast_module(netD)
    sdG = {k: v.clone() for k, v in netG.state_dict().items()}
    sdD = {k: v.clone() for k, v in netD.state_dict().items()}
    crit = GANLoss("vanilla", 0.9, 0.1, 0.9).to(dev)
    gen = torch.Generator().manual_seed(7)
    x = (torch.rand(B, 3, 64, 64, generator=gen) * 2 - 1).to(dev)
    z1, z2 = torch.randn(B, 100, generator=gen).to(dev), torch.randn(B, 100, generator=gen).to(dev)
    losses, gD, gG = step(netG, netD, crit, parallel.shard(x), parallel.shard(z1), parallel.shard(z2), True)
    lt = torch.tensor(losses, device=dev)
    torch.distributed.all_reduce(lt)
    lt /= world
    rmD = netD.blocks[1][1].running_mean.clone()
    # single-device run of the full batch on every rank (identical work), with the collectives switched off
    parallel._state["enabled"] = False
    parallel._state["world"] = 1
    netG.load_state_dict(sdG), netD.load_state_dict(sdD)
    losses1, gD1, gG1 = step(netG, netD, crit, x, z1, z2, False)

    # ---- full optimiser steps: engine.DcganStep + FusedAdam (ZeRO-1 sharded over the ranks) vs one device
    from gan_playground_b200.engine import DcganStep
    from gan_playground_b200.optim import FusedAdam

    def train(n_steps, dp):
        parallel._state["enabled"], parallel._state["world"] = dp, (world if dp else 1)
        torch.manual_seed(0)
        with contextlib.redirect_stdout(io.StringIO()):
            nG, nD = dcgan.Generator(ngf=32).to(dev), dcgan.Discriminator(ndf=32).to(dev)
        nG.load_state_dict(sdG), nD.load_state_dict(sdD)
        oG = FusedAdam(nG.parameters(), lr=4e-4, betas=(0.5, 0.999), shard=dp)
        oD = FusedAdam(nD.parameters(), lr=1e-4, betas=(0.5, 0.999), shard=dp)
        b = B // world if dp else B
        run = DcganStep(nG, nD, crit, oG, oD, b, 100, dev, use_graph=False)
        out = []
        for i in range(n_steps):
            xi, zi = (x.roll(i, 0), torch.stack([z1.roll(i, 0), z2.roll(i, 0)]))
            if dp:
                xi, zi = parallel.shard(xi), parallel.shard(zi, 1)
            out.append(run.step_eager(xi.contiguous(), zi.contiguous())[:3])
        return out, {("G." + k): v.detach().clone() for k, v in nG.state_dict().items()} | \
            {("D." + k): v.detach().clone() for k, v in nD.state_dict().items()}

    tr_dp, sd_dp = train(3, True)
    tr_1, sd_1 = train(3, False)
    parallel._state["enabled"], parallel._state["world"] = True, world
    # Adam normalises every element's step to ~lr, so element-wise weight differences amplify rounding noise wherever a
    # gradient is near zero; compare the UPDATE DIRECTION of the conv / linear weights instead (global cosine)
    sd0 = {("G." + k): v for k, v in sdG.items()} | {("D." + k): v for k, v in sdD.items()}
    keys = [k for k in sd_1 if sd_1[k].dim() >= 2]
    upd_cos = cos({k: sd_dp[k] - sd0[k] for k in keys}, {k: sd_1[k] - sd0[k] for k in keys})
    # replicas must stay bit-identical: every rank applied the same all-gathered parameters
    flat = torch.cat([sd_dp[k].float().reshape(-1) for k in sorted(sd_dp) if sd_dp[k].dtype.is_floating_point])
    hi, lo = flat.clone(), flat.clone()
    torch.distributed.all_reduce(hi, op=torch.distributed.ReduceOp.MAX)
    torch.distributed.all_reduce(lo, op=torch.distributed.ReduceOp.MIN)
    replica_diff = (hi - lo).abs().max().item()
    ldp = torch.tensor(tr_dp[-1], device=dev)
    torch.distributed.all_reduce(ldp)
    ldp /= world
    if rank == 0:
        print("world %d: DP mean losses %s vs single-device %s" % (world, [round(v, 5) for v in lt.tolist()],
                                                                   [round(v, 5) for v in losses1]))
        print("D grad cosine DP vs single: %.6f   G grad cosine: %.6f" % (cos(gD, gD1), cos(gG, gG1)))
        print("running_mean max abs diff: %.3e" % (rmD - netD.blocks[1][1].running_mean).abs().max().item())
        ok = cos(gD, gD1) > 0.999 and cos(gG, gG1) > 0.99 and all(abs(a - b) < 0.01 * abs(b) + 1e-3 for a, b in zip(lt.tolist(), losses1))
        print("3 optimiser steps, ZeRO-1 FusedAdam on %d ranks vs one device: step-3 losses %s vs %s; weight-update cosine %.5f"
              % (world, [round(v, 4) for v in ldp.tolist()], [round(v, 4) for v in tr_1[-1]], upd_cos))
        print("replica max |difference| across ranks after 3 steps: %.3e (must be 0)" % replica_diff)
        # (Adam turns near-zero gradient elements into +-lr steps, so bf16-level gradient differences flip a few percent
        # of the update signs: the cosine is informative, the losses and the replica identity are the check)
        ok = ok and replica_diff == 0.0 and upd_cos > 0.9 and all(abs(a - b) < 0.01 * abs(b) + 1e-3 for a, b in zip(ldp.tolist(), tr_1[-1]))
        print("DP_CHECK", "OK" if ok else "FAIL")
    sys.stdout.flush()
    torch.distributed.barrier()
    torch.cuda.synchronize()
    os._exit(0)


if __name__ == "__main__":
    main()
