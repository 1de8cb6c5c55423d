"""dcgan_blur (models/dcgan_blur.py + models/ops.py::BlurPool2d — the networks main_dcgan.py:52-53 builds) on the GPU:
the BlurPool kernels vs torch's reflect-pad depth-wise conv and its autograd adjoint, and one main_dcgan.py step of the
mirror vs the golden fixture produced by the unmodified reference, at the north_star bars (default bf16x3 mode)."""
import pytest
import torch
import torch.nn.functional as F

from conftest import load_golden, unpack_grads
from test_gpu_dcgan import global_cos, quiet, relerr

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("stride", [1, 2])
@pytest.mark.parametrize("shape", [(3, 8, 8, 16), (2, 16, 12, 8), (1, 2, 2, 8), (2, 5, 7, 24)])
def test_blurpool_kernels_match_torch(stride, shape):
    from gan_playground_b200 import ops

    NB, H, W, C = shape
    torch.manual_seed(0)
    x = torch.randn(NB, H, W, C, device="cuda").to(torch.bfloat16)
    a = torch.tensor([1.0, 2.0, 1.0], device="cuda")
    filt = ((a[:, None] * a[None, :]) / 16.0)[None, None].repeat(C, 1, 1, 1)
    xr = x.float().permute(0, 3, 1, 2).contiguous().requires_grad_(True)
    ref = F.conv2d(F.pad(xr, (1, 1, 1, 1), mode="reflect"), filt, stride=stride, groups=C)
    out = ops.blur3x3_fwd(x, stride)
    assert out.shape == (NB, ref.shape[2], ref.shape[3], C)
    assert (out.float().permute(0, 3, 1, 2) - ref).abs().max().item() <= 2e-2 * ref.abs().max().item()
    g = torch.randn_like(ref).to(torch.bfloat16)
    (gr,) = torch.autograd.grad(ref, xr, g.float())
    gm = ops.blur3x3_bwd(g.permute(0, 2, 3, 1).contiguous(), H, W, stride)
    assert (gm.float().permute(0, 3, 1, 2) - gr).abs().max().item() <= 2e-2 * gr.abs().max().item()


def test_dcgan_blur_golden_step():
    from gan_playground_b200.criterion import GANLoss
    from gan_playground_b200.models import dcgan_blur as M

    fx = load_golden("dcgan_blur_r32_w8.pt")
    netG = quiet(lambda: M.Generator(z_dim=fx["z_dim"], ngf=fx["width"], resolution=fx["res"])).cuda()
    netD = quiet(lambda: M.Discriminator(ndf=fx["width"], resolution=fx["res"])).cuda()
    netG.load_state_dict(fx["sd_g"])
    netD.load_state_dict(fx["sd_d"])
    crit = GANLoss(fx["mode"], *fx["labels"]).cuda()
    x, z1, z2 = fx["x"].cuda(), fx["z1"].cuda(), fx["z2"].cuda()
    out = netD(x)
    loss = crit(out, True)
    loss.backward()
    from parity import Bars, prebn_biases_blur_g

    bars = Bars("golden dcgan_blur_r32_w8 (unmodified reference, batch %d)" % x.shape[0])
    bars.act("D(x)", out, fx["d_real"]), bars.loss("loss_real", loss.item(), fx["loss_real"])
    bars.cos("D-real", global_cos(netD.named_parameters(), unpack_grads(fx["d_grads_real"])))
    fake1 = netG(z1)
    assert fake1.shape == fx["fake1"].shape
    bars.act("G(z)", fake1, fx["fake1"])
    netD.zero_grad()
    out = netD(fx["fake1"].cuda())
    crit(out, False).backward()
    bars.act("D(G(z))", out, fx["d_fake"])
    bars.cos("D-fake", global_cos(netD.named_parameters(), unpack_grads(fx["d_grads_fake"])))
    netG.zero_grad(), netD.zero_grad()
    loss = crit(netD(netG(z2)), False, True)
    loss.backward()
    bars.loss("loss_g", loss.item(), fx["loss_g"])
    skip = set(prebn_biases_blur_g(netG))
    bars.cos("G-step", global_cos(netG.named_parameters(), {k: v for k, v in unpack_grads(fx["g_grads"]).items() if k not in skip},
                                  skip_prebn=False))
    bars.finish()
    # BatchNorm bookkeeping after 3 D forwards / 2 G forwards of the loop body
    for net, key in ((netD, "buf_d_after"), (netG, "buf_g_after")):
        sd = net.state_dict()
        for k, v in fx[key].items():
            if k.endswith("num_batches_tracked"):
                assert int(sd[k]) == int(v), k
            elif k.endswith(("running_mean", "running_var")):
                assert torch.allclose(sd[k].cpu(), v, atol=2e-2 * max(1.0, v.abs().max().item())), k


def test_dcgan_blur_full_width_vs_oracle():
    """Full width (ngf = ndf = 64), 64x64, batch 16: D(x) logits and D-real gradients against the CPU oracle."""
    import os

    from gan_playground_b200.criterion import GANLoss
    from gan_playground_b200.models import dcgan_blur as M
    from oracle import gan_oracle as O

    torch.manual_seed(0)
    netG, netD = quiet(lambda: M.Generator()), quiet(lambda: M.Discriminator())
    sd_g = {k: v.clone() for k, v in netG.state_dict().items()}
    sd_d = {k: v.clone() for k, v in netD.state_dict().items()}
    gen = torch.Generator().manual_seed(1)
    B = 16
    x = torch.rand(B, 3, 64, 64, generator=gen) * 2 - 1
    z1, z2 = torch.randn(B, 100, generator=gen), torch.randn(B, 100, generator=gen)
    torch.set_num_threads(os.cpu_count())
    ref = O.dcgan_step_grads(sd_g, sd_d, x, z1, z2, blur=True)
    netG.cuda(), netD.cuda()
    crit = GANLoss("vanilla", 0.9, 0.1, 0.9).cuda()
    out = netD(x.cuda())
    crit(out, True).backward()
    assert relerr(out, ref["d_real"]) < 1e-2
    assert global_cos(netD.named_parameters(), ref["d_grads_real"]) > 0.999
    fake = netG(z1.cuda())
    assert relerr(fake, ref["fake1"]) < 1e-2
