"""engine.DcganStep: one CUDA-graph replay per step must be the same training step as the eager loop body of
main_dcgan.py:68-95 — same losses, same BatchNorm bookkeeping, and capture (which needs warm-up steps) must not train."""
import contextlib
import io

import pytest
import torch

pytestmark = pytest.mark.gpu


def _build(use_graph, torch_adam=False):
    from gan_playground_b200.criterion import GANLoss
    from gan_playground_b200.engine import DcganStep
    from gan_playground_b200.models import dcgan
    from gan_playground_b200.optim import FusedAdam

    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        netG = dcgan.Generator(ngf=16, resolution=32).cuda()
        netD = dcgan.Discriminator(ndf=16, resolution=32).cuda()
    if torch_adam:
        oG = torch.optim.Adam(netG.parameters(), lr=4e-4, betas=(0.5, 0.999), capturable=use_graph)
        oD = torch.optim.Adam(netD.parameters(), lr=1e-4, betas=(0.5, 0.999), capturable=use_graph)
    else:
        oG = FusedAdam(netG.parameters(), lr=4e-4, betas=(0.5, 0.999))
        oD = FusedAdam(netD.parameters(), lr=1e-4, betas=(0.5, 0.999))
    crit = GANLoss("vanilla", 0.9, 0.1, 0.9).cuda()
    return netG, netD, DcganStep(netG, netD, crit, oG, oD, 16, 100, torch.device("cuda", 0), use_graph=use_graph)


@pytest.mark.parametrize("torch_adam", [False, True])
def test_graph_replay_is_the_eager_step(torch_adam):
    gen = torch.Generator().manual_seed(5)
    steps = 3
    xs = (torch.rand(steps, 16, 3, 32, 32, generator=gen) * 2 - 1).cuda()
    zs = torch.randn(steps, 2, 16, 100, generator=gen).cuda()
    runs = []
    for use_graph in (False, True):
        netG, netD, runner = _build(use_graph, torch_adam)
        w0 = netG.linear.weight.detach().clone()
        losses = [runner.step(xs[i], zs[i]) for i in range(steps)]
        runs.append((losses, netG, netD, w0))
    (le, gE, dE, _), (lg, gG, dG, _) = runs
    # first step: identical weights -> the D losses agree to rounding, the G loss after one (identical) D update
    assert max(abs(a - b) for a, b in zip(le[0][:2], lg[0][:2])) < 1e-4, (le[0], lg[0])
    assert max(abs(a - b) for a, b in zip(le[0], lg[0])) < 5e-3, (le[0], lg[0])
    for a, b in zip(le[1:], lg[1:]):
        assert max(abs(x - y) for x, y in zip(a, b)) < 3e-2, (a, b)
    # capture's warm-up steps were undone: BatchNorm saw exactly 3 steps x (3 D | 2 G) forwards in both runs
    assert int(dE.blocks[1][1].num_batches_tracked) == int(dG.blocks[1][1].num_batches_tracked) == 3 * steps
    assert int(gE.blocks[0][1].num_batches_tracked) == int(gG.blocks[0][1].num_batches_tracked) == 2 * steps
    assert torch.allclose(dE.blocks[1][1].running_mean, dG.blocks[1][1].running_mean, atol=2e-3)
    # weights moved by at most the Adam step bound and mostly together
    for pe, pg in zip(list(gE.parameters()) + list(dE.parameters()), list(gG.parameters()) + list(dG.parameters())):
        assert (pe - pg).abs().max().item() <= 4 * 4e-4 * steps + 1e-4


def test_eager_step_after_replays_restages_the_weights():
    """Replays update the weights without running Python, so the operand caches filled by earlier eager calls would be
    stale (the generator's by one optimiser step): step() drops them, and an eager step afterwards re-stages and works."""
    import math

    gen = torch.Generator().manual_seed(6)
    x = (torch.rand(16, 3, 32, 32, generator=gen) * 2 - 1).cuda()
    z = torch.randn(2, 16, 100, generator=gen).cuda()
    netG, netD, runner = _build(True)
    for _ in range(3):
        out = runner.step(x, z)
    assert all(math.isfinite(v) for v in out)
    assert len(netG._gp_cache._d) == 0 and len(netD._gp_cache._d) == 0
    vals = runner.step_eager(x, z)
    assert all(math.isfinite(v) for v in vals)
    assert len(netG._gp_cache._d) > 0                      # staged again from the current parameters
    # and replaying again after an eager step still works (the graph owns its own staging kernels)
    out2 = runner.step(x, z)
    assert all(math.isfinite(v) for v in out2)


@pytest.mark.parametrize("use_graph", [False, True])
def test_weight_gradients_on_the_side_stream_are_the_same_step(use_graph):
    """config.wgrad_side: the step drivers fork the weight-gradient GEMMs off the backward chain and join before the
    optimiser. Same inputs, same seeds: losses and updated weights must agree with the in-chain run to atomics noise."""
    from gan_playground_b200 import config

    gen = torch.Generator().manual_seed(7)
    steps = 3
    xs = (torch.rand(steps, 16, 3, 32, 32, generator=gen) * 2 - 1).cuda()
    zs = torch.randn(steps, 2, 16, 100, generator=gen).cuda()
    runs = []
    try:
        for mode in ("1", "0"):
            config.set_wgrad_stream_mode(mode)
            netG, netD, runner = _build(use_graph)
            assert (runner._wg_stream is not None) == (mode == "1")
            losses = [runner.step(xs[i], zs[i]) for i in range(steps)]
            torch.cuda.synchronize()
            runs.append((losses, [p.detach().clone() for p in list(netG.parameters()) + list(netD.parameters())]))
    finally:
        config.set_wgrad_stream_mode("auto")
    (la, pa), (lb, pb) = runs
    assert max(abs(x - y) for x, y in zip(la[0], lb[0])) < 1e-4, (la[0], lb[0])
    for a, b in zip(la[1:], lb[1:]):
        assert max(abs(x - y) for x, y in zip(a, b)) < 2e-2, (a, b)
    for x, y in zip(pa, pb):
        assert (x - y).abs().max().item() <= 4 * 4e-4 * steps + 1e-4


@pytest.mark.parametrize("use_graph", [False, True])
def test_generator_forward_next_to_the_real_pass_is_the_same_step(use_graph):
    """config.g_ahead: G(z1) of the D-fake chain is issued on its own stream before the real-image pass. Same inputs,
    same seeds: the step must agree with the in-order run to atomics noise."""
    from gan_playground_b200 import config

    gen = torch.Generator().manual_seed(8)
    steps = 3
    xs = (torch.rand(steps, 16, 3, 32, 32, generator=gen) * 2 - 1).cuda()
    zs = torch.randn(steps, 2, 16, 100, generator=gen).cuda()
    runs = []
    try:
        for mode in ("1", "0"):
            config.set_g_ahead_mode(mode)
            netG, netD, runner = _build(use_graph)
            assert (runner._g_stream is not None) == (mode == "1")
            losses = [runner.step(xs[i], zs[i]) for i in range(steps)]
            torch.cuda.synchronize()
            runs.append((losses, [p.detach().clone() for p in list(netG.parameters()) + list(netD.parameters())],
                         netG.blocks[0][1].running_mean.clone()))
    finally:
        config.set_g_ahead_mode("auto")
    (la, pa, ra), (lb, pb, rb) = runs
    assert max(abs(x - y) for x, y in zip(la[0], lb[0])) < 1e-4, (la[0], lb[0])
    for a, b in zip(la[1:], lb[1:]):
        assert max(abs(x - y) for x, y in zip(a, b)) < 2e-2, (a, b)
    for x, y in zip(pa, pb):
        assert (x - y).abs().max().item() <= 4 * 4e-4 * steps + 1e-4
    assert torch.allclose(ra, rb, atol=2e-3)
