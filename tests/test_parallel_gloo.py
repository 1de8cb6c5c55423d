"""Host-side data-parallel logic on CPU with the gloo backend, world_size 2 (the N>1 path of bench.py):
shard(), all_reduce_sum_ (SyncBN partial sums), GradBucket averaging, broadcast of replicas."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from gan_playground_b200 import parallel

    r, w = parallel.init(backend="gloo")
    assert (r, w) == (rank, world) and parallel.enabled()
    # shard: rank r owns the contiguous slice [r*B/N, (r+1)*B/N)
    full = torch.arange(8.0).view(8, 1)
    mine = parallel.shard(full)
    assert mine.flatten().tolist() == [4.0 * rank + i for i in range(4)]
    # SyncBN partial sums: global statistics == statistics of the concatenated batch
    st = torch.stack([mine.sum(0), (mine * mine).sum(0)])
    parallel.all_reduce_sum_(st)
    assert st[0].item() == full.sum().item() and st[1].item() == (full * full).sum().item()
    # replicas: broadcast from rank 0
    lin = torch.nn.Linear(3, 2)
    with torch.no_grad():
        lin.weight.fill_(float(rank + 1))
    parallel.broadcast_module(lin)
    assert lin.weight.eq(1.0).all()
    # gradient averaging through the flat bucket; accumulation over two backwards first (D-real + D-fake)
    bucket = parallel.GradBucket(lin)
    bucket.attach()
    x = torch.ones(1, 3) * (rank + 1)
    lin(x).sum().backward()
    lin(x).sum().backward()
    bucket.all_reduce_mean()
    expect = 2 * (1 + 2) / 2.0                      # two backwards, mean over ranks of x = 1 and x = 2
    assert torch.allclose(lin.weight.grad, torch.full((2, 3), expect))
    assert lin.weight.grad.data_ptr() == bucket.views[0].data_ptr()
    with parallel.no_sync():
        pass
    parallel.shutdown()
    q.put((rank, "ok"))


def _adam_reference(p, g, m, v, step, lr, beta1, beta2, eps, grad_scale=1.0):
    """torch restatement of csrc/optim.cu (test double for the CUDA kernel in the CPU-only gloo test)."""
    t = step.item() + 1.0
    g = g * grad_scale
    m.add_((g - m) * (1.0 - beta1))
    v.mul_(beta2).add_((1.0 - beta2) * g * g)
    p.sub_((lr / (1.0 - beta1 ** t)) * m / (v.sqrt() / (1.0 - beta2 ** t) ** 0.5 + eps))


def _zero1_worker(rank, world, port, q):
    """optim.FusedAdam(shard=True): reduce-scatter of the flat gradient, 1/N of the update per rank, all-gather of the
    parameters == torch.optim.Adam on the rank-averaged gradients (host logic only: the kernel is replaced)."""
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from gan_playground_b200 import ops, parallel
    from gan_playground_b200.optim import FusedAdam

    ops.adam_flat = _adam_reference
    parallel.init(backend="gloo")
    torch.manual_seed(0)
    shapes = [(5, 3), (7,), (4, 2, 2, 2), (1,), (33,)]
    ref = [torch.nn.Parameter(torch.randn(s)) for s in shapes]
    got = [torch.nn.Parameter(p.detach().clone()) for p in ref]
    o_ref = torch.optim.Adam(ref, lr=1e-2, betas=(0.5, 0.999))
    o_got = FusedAdam(got, lr=1e-2, betas=(0.5, 0.999), shard=True)
    assert o_got.shard and o_got._flat[0]["m"].numel() * world == o_got._flat[0]["n"]
    gen = torch.Generator().manual_seed(5)
    for step in range(4):
        o_ref.zero_grad()
        o_got.zero_grad()
        for a, b in zip(ref, got):
            g_ranks = [torch.randn(a.shape, generator=gen) for _ in range(world)]   # same stream on every rank
            a.grad = sum(g_ranks) / world
            b.grad.add_(g_ranks[rank])                                           # two accumulating backwards
            b.grad.add_(torch.zeros_like(b.grad))
        o_ref.step()
        o_got.step()
        for a, b in zip(ref, got):
            assert torch.allclose(a, b, atol=1e-6, rtol=1e-5), (step, a.shape)
    parallel.shutdown()
    q.put((rank, "ok"))


def test_zero1_fused_adam_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_zero1_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert sorted(q.get(timeout=5) for _ in range(2)) == [(0, "ok"), (1, "ok")]


def test_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    got = sorted(q.get(timeout=5) for _ in range(2))
    assert got == [(0, "ok"), (1, "ok")]


def test_single_process_is_a_noop():
    from gan_playground_b200 import parallel

    assert parallel.world_size() == 1 and not parallel.enabled()
    t = torch.ones(3)
    assert parallel.all_reduce_sum_(t) is t
    assert parallel.shard(t) is t


# ---- the step drivers under data parallelism (host logic): BatchNorm-free stand-in nets, so that two ranks on half
# ---- batches with averaged gradients must reproduce the single-process run on the full batch exactly
def _toy_nets(seed):
    import torch.nn as nn

    torch.manual_seed(seed)

    class G(nn.Module):
        def __init__(self, cond):
            super().__init__()
            self.fc = nn.Linear(4 + (3 if cond else 0), 3 * 4 * 4)
            self.cond = cond

        def forward(self, z, y=None):
            if self.cond:
                y = y if y.is_floating_point() else torch.nn.functional.one_hot(y, 3).float()
                z = torch.cat([z, y], 1)
            return torch.tanh(self.fc(z)).view(-1, 3, 4, 4)

    class D(nn.Module):
        def __init__(self, heads):
            super().__init__()
            self.body = nn.Linear(3 * 4 * 4, 8)
            self.out = nn.Linear(8, heads)
            self.img_dim, self.resolution = 3, 4
            self.block1 = nn.Module()
            self.block1.c1 = nn.Module()
            self.block1.c1.in_channels = 3

        def packed_logits(self, x):
            return self.out(torch.nn.functional.leaky_relu(self.body(x.flatten(1)), 0.2))

        def forward(self, x, y=None):
            return self.packed_logits(x)[:, :1]

    return G, D


def _drive(kind, rank, world, steps=3, B=8):
    """Run `steps` iterations of one driver on this rank's shard (or on the full batch when world == 1)."""
    from gan_playground_b200 import engine, parallel
    from gan_playground_b200.criterion import ACGANLoss, GANLoss
    from oracle import gan_oracle as O

    G, D = _toy_nets(0)
    netG, netD = G(kind != "dcgan"), D(4 if kind == "acgan" else 1)
    parallel.broadcast_module(netG), parallel.broadcast_module(netD)
    optG = torch.optim.SGD(netG.parameters(), lr=0.1)
    optD = torch.optim.SGD(netD.parameters(), lr=0.1)
    gen = torch.Generator().manual_seed(3)
    xs = torch.rand(steps, B, 3, 4, 4, generator=gen) * 2 - 1
    zs = torch.randn(steps, 2, B, 4, generator=gen)
    ys = torch.randint(3, (steps, B), generator=gen)
    cs = torch.randint(3, (steps, B), generator=gen)
    yf = torch.randint(0, 2, (steps, B, 3), generator=gen).float()
    b = B // world
    sl = slice(rank * b, (rank + 1) * b)
    dev = torch.device("cpu")

    def crit(pred, is_real, is_generator=False):
        return O.gan_loss("hinge", pred, is_real, is_generator)

    class CpuACGANLoss(ACGANLoss):
        def forward(self, packed, labels, is_real, is_generator=False):
            adv, cls = packed[:, :1], packed[:, 1:]
            l_adv = O.gan_loss("vanilla", adv, is_real, is_generator, 0.9, 0.1, 0.9)
            l_aux = torch.nn.functional.mse_loss(cls, labels)
            return torch.stack([l_adv, l_aux, l_adv + self.aux_weight * l_aux, torch.sigmoid(adv).mean()])

    if kind == "dcgan":
        run = engine.DcganStep(netG, netD, crit, optG, optD, b, 4, dev, overlap=False)
        step = lambda i: run.step(xs[i, sl], zs[i, :, sl])
    elif kind == "sngan":
        run = engine.SnganStep(netG, netD, crit, optG, optD, b, 4, dev, n_classes=3, n_disc_update=2, resolution=4,
                               overlap=False)
        step = lambda i: run.step(xs[i, sl], ys[i, sl], zs[i, 0, sl], cs[i, sl])
    else:
        run = engine.AcganStep(netG, netD, CpuACGANLoss(GANLoss("vanilla", 0.9, 0.1, 0.9)), optG, optD, b, 4, dev, n_class=3,
                               overlap=False)
        step = lambda i: run.step(xs[i, sl], yf[i, sl], zs[i, 0, sl])
    for i in range(steps):
        step(i)
    return [p.detach().clone() for p in list(netG.parameters()) + list(netD.parameters())]


def _driver_worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from gan_playground_b200 import parallel

    parallel.init(backend="gloo")
    out = {kind: [t.tolist() for t in _drive(kind, rank, world)] for kind in ("dcgan", "sngan", "acgan")}   # plain lists:
    parallel.shutdown()                                                   # tensors in a Queue need the sender alive
    q.put((rank, out))


def test_step_drivers_data_parallel_equals_full_batch_gloo():
    """engine.DcganStep / SnganStep / AcganStep on 2 gloo ranks (half batch each, gradients averaged through the flat
    buckets before every optimiser step) == the same driver in one process on the full batch: the mean losses of the
    scripts make the averaged shard gradients the full-batch gradients. Replicas must also be identical to each other."""
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    single = {kind: _drive(kind, 0, 1) for kind in ("dcgan", "sngan", "acgan")}
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_driver_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=180) for _ in range(2))
    for p in procs:
        p.join(60)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    for kind, ref in single.items():
        for a, b0, b1 in zip(ref, got[0][kind], got[1][kind]):
            assert b0 == b1, kind                                              # replicas stay identical
            b0 = torch.tensor(b0)
            assert torch.allclose(a, b0, atol=1e-6, rtol=1e-5), (kind, (a - b0).abs().max())


def _checkpoint_worker(rank, world, port, q, tmp):
    """checkpoint.save_model / sample_images under data parallelism: only rank 0 writes, every rank runs the sampling
    forward (train-mode BatchNorm statistics must stay replicated), the file has the scripts' layout."""
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from gan_playground_b200 import checkpoint, parallel

    parallel.init(backend="gloo")
    torch.manual_seed(rank)          # different initial replicas on purpose: broadcast makes them equal
    netG = torch.nn.Sequential(torch.nn.Linear(4, 3 * 4 * 4), torch.nn.Unflatten(1, (3, 4, 4)), torch.nn.BatchNorm2d(3))
    netD = torch.nn.Linear(5, 1)
    parallel.broadcast_module(netG), parallel.broadcast_module(netD)
    optG, optD = torch.optim.Adam(netG.parameters(), lr=1e-3), torch.optim.Adam(netD.parameters(), lr=1e-3)
    netG(torch.randn(6, 4)).sum().backward()
    optG.step()
    netG.train()
    out = checkpoint.sample_images(netG, torch.ones(8, 4), os.path.join(tmp, "res", "fake.jpg"))
    assert out.shape == (8, 3, 4, 4) and int(netG[2].num_batches_tracked) == 2      # sampled in train mode on EVERY rank
    path = checkpoint.save_model((netG, netD), (optG, optD), 0, os.path.join(tmp, "ckpt"))
    assert path.endswith("checkpoint_001.pth") and os.path.exists(path) and os.path.exists(os.path.join(tmp, "res", "fake.jpg"))
    q.put((rank, sorted(os.listdir(os.path.join(tmp, "ckpt")))))
    parallel.shutdown()


def test_checkpoint_and_sampling_write_on_rank_zero_only(tmp_path):
    ctx = mp.get_context("spawn")
    q, port = ctx.Queue(), _free_port()
    procs = [ctx.Process(target=_checkpoint_worker, args=(r, 2, port, q, str(tmp_path))) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert sorted(q.get(timeout=5)[1] for _ in range(2)) == [["checkpoint_001.pth"]] * 2
    ck = torch.load(tmp_path / "ckpt" / "checkpoint_001.pth", weights_only=False)
    assert sorted(ck) == ["epoch", "optimizer", "state_dict"] and ck["epoch"] == 0
    assert sorted(ck["state_dict"]) == sorted(ck["optimizer"]) == ["discriminator", "generator"]
    assert ck["optimizer"]["generator"]["state"][0]["exp_avg"].abs().sum() > 0
