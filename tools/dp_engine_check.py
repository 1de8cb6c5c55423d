"""Data-parallel check of the step DRIVER (run under torchrun with N ranks): engine.DcganStep as one CUDA graph per step
with every small-shard overlap on (weight gradients on a side stream, G(z1) next to the real-image pass on the second
peer lane) must (1) keep the replicas bit-identical — parameters and BatchNorm buffers — and (2) log the same first-step
losses as the in-order run.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 tools/dp_engine_check.py
"""
import contextlib
import io
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch
import torch.distributed as dist

from gan_playground_b200 import config, parallel
from gan_playground_b200.criterion import GANLoss
from gan_playground_b200.engine import DcganStep
from gan_playground_b200.models import dcgan
from gan_playground_b200.optim import FusedAdam


def run(mode, dev, rank, world, steps=4, per_gpu=32):
    config.set_wgrad_stream_mode(mode)
    config.set_g_ahead_mode(mode)
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        netG, netD = dcgan.Generator(ngf=32, resolution=32).to(dev), dcgan.Discriminator(ndf=32, resolution=32).to(dev)
    parallel.broadcast_module(netG)
    parallel.broadcast_module(netD)
    oG = FusedAdam(netG.parameters(), lr=4e-4, betas=(0.5, 0.999), shard=True)
    oD = FusedAdam(netD.parameters(), lr=1e-4, betas=(0.5, 0.999), shard=True)
    crit = GANLoss("vanilla", 0.9, 0.1, 0.9).to(dev)
    runner = DcganStep(netG, netD, crit, oG, oD, per_gpu, 100, dev, use_graph=True)
    gen = torch.Generator().manual_seed(11)
    xs = (torch.rand(steps, world * per_gpu, 3, 32, 32, generator=gen) * 2 - 1)
    zs = torch.randn(steps, 2, world * per_gpu, 100, generator=gen)
    sl = slice(rank * per_gpu, (rank + 1) * per_gpu)
    losses = [runner.step(xs[i, sl].to(dev), zs[i, :, sl].to(dev).contiguous()) for i in range(steps)]
    torch.cuda.synchronize()
    state = torch.cat([t.detach().float().flatten() for n in (netG, netD) for t in list(n.parameters()) + list(n.buffers())])
    return losses, state


def main():
    parallel.init()
    rank, world = parallel.rank(), parallel.world_size()
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
    peer = parallel.init_peer_sync(dev)
    la, sa = run("1", dev, rank, world)
    lb, sb = run("0", dev, rank, world)
    ok = True
    for name, st in (("overlap on", sa), ("overlap off", sb)):
        ref = st.clone()
        dist.broadcast(ref, src=0)
        same = bool(torch.equal(ref, st))
        ok &= same
        if rank == 0 or not same:
            print("rank %d: replicas identical to rank 0 (%s): %s" % (rank, name, same))
    d0 = max(abs(x - y) for x, y in zip(la[0], lb[0]))
    ok &= d0 < 1e-3
    drift = float((sa - sb).abs().max())
    if rank == 0:
        print("peer SyncBN:", peer, "| first-step losses on / off:", [round(v, 5) for v in la[0]], [round(v, 5) for v in lb[0]],
              "| max diff %.2e" % d0, "| max parameter / buffer difference after %d steps %.3e" % (len(la), drift))
        print("DP ENGINE CHECK:", "OK" if ok else "FAILED")
    parallel.shutdown()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
