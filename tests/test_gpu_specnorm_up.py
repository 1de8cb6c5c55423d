"""dcgan_specnorm_up (models/dcgan_specnorm_up.py: upsample + spectral-normed Conv3x3 generator, dcgan_specnorm's
discriminator — SURVEY.md §8f row 4) on the GPU: the golden step of the unmodified reference, and the north_star bars at
width 64 against the CPU oracle."""
import os

import pytest
import torch

from conftest import load_golden, unpack_grads
from parity import Bars, global_cos, prebn_biases, quiet
from test_gpu_parity_bars import _clone_sd, _three_passes

pytestmark = pytest.mark.gpu


def prebn_biases_up_g(net):
    """Generator blocks are [Upsample, SN Conv3x3, BatchNorm, ReLU]: `blocks.i.1.bias` feeds the BatchNorm `blocks.i.2`
    (analytically-zero gradient, rounding noise in the reference)."""
    names = dict(net.named_parameters())
    return [k for k in names if k.endswith(".1.bias") and k.replace(".1.bias", ".2.weight") in names]


def test_specnorm_up_golden_step():
    from gan_playground_b200.criterion import GANLoss
    from gan_playground_b200.models import dcgan_specnorm_up as M

    fx = load_golden("snd_up_r32_w4.pt")
    netG = quiet(lambda: M.Generator(z_dim=fx["z_dim"], ngf=fx["width"], resolution=32)).cuda()
    netD = quiet(lambda: M.Discriminator(ndf=fx["width"], resolution=32)).cuda()
    netG.load_state_dict(fx["sd_g"])
    netD.load_state_dict(fx["sd_d"])
    crit = GANLoss("hinge").cuda()
    x, z1, z2 = fx["x"].cuda(), fx["z1"].cuda(), fx["z2"].cuda()
    bars = Bars("golden snd_up_r32_w4 (unmodified reference, dcgan_specnorm_up width %d, batch %d)" % (fx["width"], x.shape[0]))
    skip_g, skip_d = prebn_biases_up_g(netG), prebn_biases(netD)
    out = netD(x)
    loss = crit(out, True)
    loss.backward()
    bars.act("D(x)", out, fx["d_real"]), bars.loss("loss_real", loss.item(), fx["loss_real"])
    bars.cos("D-real", global_cos(netD.named_parameters(), unpack_grads(fx["d_grads_real"]), skip_d))
    fake1 = netG(z1)
    assert fake1.shape == (x.shape[0], 3, 32, 32) and fake1.dtype == torch.float32
    bars.act("G(z)", fake1, fx["fake1"])
    netD.zero_grad()
    out = netD(fx["fake1"].cuda())
    crit(out, False).backward()
    bars.act("D(G(z))", out, fx["d_fake"])
    bars.cos("D-fake", global_cos(netD.named_parameters(), unpack_grads(fx["d_grads_fake"]), skip_d))
    netG.zero_grad(), netD.zero_grad()
    loss = crit(netD(netG(z2)), False, True)
    loss.backward()
    assert abs(loss.item() - fx["loss_g"].item()) < 0.02 * abs(fx["loss_g"].item()) + 1e-3   # -mean D(G(z)) sits near zero
    bars.cos("G-step", global_cos(netG.named_parameters(), unpack_grads(fx["g_grads"]), skip_g))
    bars.finish()
    # u / v after 3 D forwards and 2 G forwards (one in-place power iteration per train-mode forward), BatchNorm counters
    for net, key in ((netD, "buf_d_after"), (netG, "buf_g_after")):
        sd = net.state_dict()
        for k, v in fx[key].items():
            if k.endswith(("weight_u", "weight_v")):
                assert torch.allclose(sd[k].cpu(), v, atol=2e-4), k
            if k.endswith("num_batches_tracked"):
                assert int(sd[k]) == int(v)
            if k.endswith(("running_mean", "running_var")):
                assert torch.allclose(sd[k].cpu(), v, rtol=2e-2, atol=2e-3), k


def test_specnorm_up32_width64_meets_the_bars():
    from gan_playground_b200.criterion import GANLoss
    from gan_playground_b200.models import dcgan_specnorm_up as M
    from oracle import gan_oracle as O

    torch.manual_seed(0)
    netG, netD = quiet(lambda: M.Generator(resolution=32)), quiet(lambda: M.Discriminator(resolution=32))
    sd_g, sd_d = _clone_sd(netG), _clone_sd(netD)
    gen = torch.Generator().manual_seed(5)
    B = 32
    x = torch.rand(B, 3, 32, 32, generator=gen) * 2 - 1
    z = torch.randn(2, B, 100, generator=gen)
    torch.set_num_threads(os.cpu_count())
    ref = O.dcgan_step_grads(sd_g, sd_d, x, z[0], z[1], labels=(1.0, 0.0, 1.0), mode="hinge", up=True)
    netG.cuda(), netD.cuda()
    crit = GANLoss("hinge").cuda()
    bars = Bars("f4 dcgan_specnorm_up-32 hinge w64 B=%d (default mode)" % B, loss_abs=0.05)    # hinge G loss sits near zero
    xd, zd = x.cuda(), z.cuda()
    _three_passes(bars, netG, netD, crit, ref, lambda: netD(xd), lambda i: netG(zd[i]), lambda img: netD(img), None,
                  skip_g=prebn_biases_up_g(netG))
    bars.finish()
