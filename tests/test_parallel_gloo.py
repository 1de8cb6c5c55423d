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
