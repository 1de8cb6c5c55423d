"""SN-DCGAN (models/dcgan_specnorm.py) and ACGAN (models/acgan.py) mirrors on the GPU vs golden fixtures from the
reference and vs the CPU oracle. Spectral-norm buffers (u, v) are fp32 GEMV results: compared at 1e-4."""
import contextlib
import io
import os

import pytest
import torch

from conftest import load_golden, unpack_grads
from test_gpu_dcgan import global_cos, quiet, relerr

pytestmark = pytest.mark.gpu


def test_spectral_norm_kernels_match_torch_hook():
    """gp_sn_* vs torch.nn.utils.spectral_norm itself (the third-party code the reference calls), Conv2d (dim 0) and
    ConvTranspose2d (dim 1): sigma-normalised weight, in-place u/v after 3 forwards, and weight_orig gradient."""
    from gan_playground_b200 import functional as GF

    for ctor, dim in ((lambda: torch.nn.Conv2d(24, 40, 4, 2, 1), 0), (lambda: torch.nn.ConvTranspose2d(24, 40, 4, 2, 1), 1),
                      (lambda: torch.nn.Linear(96, 1), 0), (lambda: torch.nn.Embedding(10, 64), 0)):
        torch.manual_seed(0)
        ref = torch.nn.utils.spectral_norm(ctor()).cuda()
        mine_w = ref.weight_orig.detach().clone().requires_grad_(True)
        u, v = ref.weight_u.detach().clone(), ref.weight_v.detach().clone()
        ref.train()
        for it in range(3):
            # run torch's hook (it fires in the module's forward pre-hook) by touching forward on a dummy input
            hook = next(iter(ref._forward_pre_hooks.values()))
            hook(ref, None)
            w_ref = ref.weight
            w_mine = GF.SpectralNormFn.apply(mine_w, u, v, dim, True)
            assert torch.allclose(w_mine, w_ref, rtol=1e-4, atol=1e-6), (dim, it)
        assert torch.allclose(u, ref.weight_u, atol=1e-5) and torch.allclose(v, ref.weight_v, atol=1e-5)
        g = torch.randn_like(w_ref)
        (gr,) = torch.autograd.grad(w_ref, ref.weight_orig, g)
        (gm,) = torch.autograd.grad(w_mine, mine_w, g)
        assert torch.allclose(gm, gr, rtol=1e-3, atol=1e-5 * gr.abs().max().item())
        # eval mode: no power iteration, sigma from the stored u, v
        ref.eval()
        hook(ref, None)
        u0 = u.clone()
        w_eval = GF.SpectralNormFn.apply(mine_w, u, v, dim, False)
        assert torch.equal(u, u0) and torch.allclose(w_eval, ref.weight, rtol=1e-4, atol=1e-6)


def test_sn_dcgan_golden_step():
    from gan_playground_b200.criterion import GANLoss
    from gan_playground_b200.models import dcgan_specnorm as M
    from parity import Bars

    fx = load_golden("snd_r32_w4.pt")
    netG = quiet(lambda: M.Generator(z_dim=fx["z_dim"], ngf=fx["width"], resolution=32)).cuda()
    netD = quiet(lambda: M.Discriminator(ndf=fx["width"], resolution=32)).cuda()
    netG.load_state_dict(fx["sd_g"])
    netD.load_state_dict(fx["sd_d"])
    crit = GANLoss("hinge").cuda()
    x, z1, z2 = fx["x"].cuda(), fx["z1"].cuda(), fx["z2"].cuda()
    bars = Bars("golden snd_r32_w4 (unmodified reference, SN-DCGAN width %d, batch %d)" % (fx["width"], x.shape[0]))
    out = netD(x)
    loss = crit(out, True)
    loss.backward()
    bars.act("D(x)", out, fx["d_real"]), bars.loss("loss_real", loss.item(), fx["loss_real"])
    bars.cos("D-real", global_cos(netD.named_parameters(), unpack_grads(fx["d_grads_real"])))
    fake1 = netG(z1)
    bars.act("G(z)", fake1, fx["fake1"])
    netD.zero_grad()
    out = netD(fx["fake1"].cuda())
    crit(out, False).backward()
    bars.act("D(G(z))", out, fx["d_fake"])
    bars.cos("D-fake", global_cos(netD.named_parameters(), unpack_grads(fx["d_grads_fake"])))
    netG.zero_grad(), netD.zero_grad()
    loss = crit(netD(netG(z2)), False, True)
    loss.backward()
    assert abs(loss.item() - fx["loss_g"].item()) < 0.02 * abs(fx["loss_g"].item()) + 1e-3   # -mean D(G(z)) sits near zero
    bars.cos("G-step", global_cos(netG.named_parameters(), unpack_grads(fx["g_grads"])))
    bars.finish()
    # u / v after 3 D forwards and 2 G forwards (one in-place power iteration per train-mode forward)
    for net, key in ((netD, "buf_d_after"), (netG, "buf_g_after")):
        sd = net.state_dict()
        for k, v in fx[key].items():
            if k.endswith(("weight_u", "weight_v")):
                assert torch.allclose(sd[k].cpu(), v, atol=2e-4), k
            if k.endswith("num_batches_tracked"):
                assert int(sd[k]) == int(v)
    out, hid = netD(x, out_hidden=True)
    assert hid.shape == (x.shape[0], fx["width"] * 8, 4, 4) and hid.dtype == torch.float32


def test_acgan_golden_step():
    from gan_playground_b200.criterion import GANLoss
    from gan_playground_b200.models import acgan as M
    from parity import Bars

    fx = load_golden("acgan_r64_w4.pt")
    netG = quiet(lambda: M.Generator(z_dim=16, ngf=fx["width"], n_class=10)).cuda()
    netD = quiet(lambda: M.Discriminator(ndf=fx["width"], n_class=10)).cuda()
    netG.load_state_dict(fx["sd_g"])
    netD.load_state_dict(fx["sd_d"])
    crit = GANLoss("vanilla", 0.9, 0.1, 0.9).cuda()
    mse = torch.nn.MSELoss()
    x, y, z = fx["x"].cuda(), fx["y"].cuda(), fx["z"].cuda()
    bars = Bars("golden acgan_r64_w4 (unmodified reference, ACGAN width %d, batch %d)" % (fx["width"], x.shape[0]))
    adv, cls = netD(x)
    loss = crit(adv, True) + mse(cls, y) * 0.5       # main_acgan.py:95-97
    loss.backward()
    bars.act("D(x) adv", adv, fx["d_real"]), bars.act("D(x) aux", cls, fx["d_real_cls"])
    bars.loss("loss_real", loss.item(), fx["loss_real"])
    bars.cos("D-real", global_cos(netD.named_parameters(), unpack_grads(fx["d_grads_real"])))
    fake = netG(z, y)
    bars.act("G(z,y)", fake, fx["fake"])
    netG.zero_grad(), netD.zero_grad()
    adv, cls = netD(fake)
    loss = crit(adv, False, True) + mse(cls, y) * 0.5
    loss.backward()
    bars.loss("loss_g", loss.item(), fx["loss_g"])
    bars.cos("G-step", global_cos(netG.named_parameters(), unpack_grads(fx["g_grads"])))
    bars.finish()


def test_batched_spectral_norm_matches_torch_hooks():
    """functional.spectral_norm_all — every hook of a forward in one batched call — vs torch.nn.utils.spectral_norm itself on
    a mixed set of layers (Conv2d 3x3 / 1x1, ConvTranspose2d (dim 1), Linear with a length-1 u, Embedding): W / sigma, the
    in-place u / v after 3 training forwards, the gradient through sigma (some outputs unused), and eval mode."""
    from gan_playground_b200 import functional as GF

    ctors = [(lambda: torch.nn.Conv2d(24, 40, 3, 1, 1), 0), (lambda: torch.nn.Conv2d(16, 64, 1), 0),
             (lambda: torch.nn.ConvTranspose2d(24, 40, 4, 2, 1), 1), (lambda: torch.nn.Linear(96, 1), 0),
             (lambda: torch.nn.Embedding(10, 64), 0), (lambda: torch.nn.Conv2d(3, 8, 3, 1, 1), 0)]
    torch.manual_seed(0)
    refs = [torch.nn.utils.spectral_norm(c()).cuda() for c, _ in ctors]
    dims = [d for _, d in ctors]
    mine = [torch.nn.utils.spectral_norm(c()).cuda() for c, _ in ctors]
    for m, r in zip(mine, refs):
        m.load_state_dict(r.state_dict())
    for r in refs:
        r.train()
    for it in range(3):
        for r in refs:
            next(iter(r._forward_pre_hooks.values()))(r, None)
        table = GF.spectral_norm_all(mine, dims, True)
        for m, r in zip(mine, refs):
            assert torch.allclose(table[m], r.weight, rtol=1e-4, atol=1e-6), (type(r).__name__, it)
    for m, r in zip(mine, refs):
        assert torch.allclose(m.weight_u, r.weight_u, atol=1e-5) and torch.allclose(m.weight_v, r.weight_v, atol=1e-5)
    # gradient through sigma: outputs 0, 2, 3 receive one, the others are unused
    used = (0, 2, 3)
    gs = {i: torch.randn_like(refs[i].weight) for i in used}
    loss_ref = sum((refs[i].weight * gs[i]).sum() for i in used)
    loss_mine = sum((table[mine[i]] * gs[i]).sum() for i in used)
    g_ref = torch.autograd.grad(loss_ref, [refs[i].weight_orig for i in used])
    g_mine = torch.autograd.grad(loss_mine, [mine[i].weight_orig for i in used])
    for a, b in zip(g_mine, g_ref):
        assert torch.allclose(a, b, rtol=1e-3, atol=1e-5 * b.abs().max().item())
    # eval mode: no power iteration
    u0 = [m.weight_u.clone() for m in mine]
    for r in refs:
        r.eval()
        next(iter(r._forward_pre_hooks.values()))(r, None)
    table = GF.spectral_norm_all(mine, dims, False)
    for m, r, u in zip(mine, refs, u0):
        assert torch.equal(m.weight_u, u) and torch.allclose(table[m], r.weight, rtol=1e-4, atol=1e-6)
