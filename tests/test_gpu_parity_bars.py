"""North-star parity bars at BASELINE width for EVERY configuration of BASELINE.json, against the fp32 CPU oracle on the
same seeded inputs: per-boundary activation max-rel-error <= 1e-2, gradient cosine >= 0.999 for each of the three
backward passes of a step (D-real, D-fake, G step), losses within 2 %. No threshold here is looser than north_star's.

  cfg 1/2  DCGAN-64 (models/dcgan.py), ngf = ndf = 64, through engine.DcganStep's per-pass precision policy at the
           benched batch sizes: 64 (cfg 1), 128 (the 8-GPU shard of cfg 2) and 1024 (cfg 2 on one GPU)
  cfg 3    SN-DCGAN-32 hinge (models/dcgan_specnorm.py), width 64
  cfg 4    SNGAN projection pair (models/sngan_projection.py), ch = 64, 32x32, 10 classes, bottom_width 2
  cfg 5    ACGAN-64 (models/acgan.py), width 64, 10 attributes
  f1       dcgan_blur (models/dcgan_blur.py — what main_dcgan.py:52-53 instantiates), width 64

Every measured number is printed and appended to gpurun_out/parity_table.jsonl (tests/parity.py) — the table of
DESIGN.md §5 is that file."""
import os

import pytest
import torch

from parity import Bars, global_cos, prebn_biases, prebn_biases_blur_g, quiet, resnet_g_zero_biases

pytestmark = pytest.mark.gpu


def _clone_sd(net):
    return {k: v.clone() for k, v in net.state_dict().items()}


def _grads(net):
    return {k: p.grad.detach().clone() for k, p in net.named_parameters() if p.grad is not None}


class _G:   # global_cos reads .grad
    def __init__(self, g):
        self.grad = g


def _minus(net, before):
    return {k: _G(p.grad.detach() - before[k]) for k, p in net.named_parameters() if k in before}


def _three_passes(bars, netG, netD, crit, ref, d_real, gen_fake, d_on, scopes, skip_g=(), fake_key="fake1", second_z=True):
    """The three backward passes of one iteration on the GPU mirrors, recorded against the oracle's `ref`.
    d_real() -> logits of the real batch; gen_fake(i) -> generator output for noise i; d_on(img) -> logits.
    scopes = (real, fake, gstep) precision scopes (engine policy) or None. second_z: main_dcgan.py draws fresh noise
    for the G step (two generator forwards); main_sngan.py / main_acgan.py re-use the fake batch."""
    import contextlib

    sc = scopes or (contextlib.nullcontext(),) * 3
    skip_d = prebn_biases(netD)
    with sc[0]:
        out = d_real()
    loss = crit(out, True)
    loss.backward()
    bars.act("D(x)", out, ref["d_real"]), bars.loss("loss_real", loss.item(), ref["loss_real"])
    bars.cos("D-real", global_cos(netD.named_parameters(), ref["d_grads_real"], skip_d))
    g_real = _grads(netD)
    with sc[1]:
        fake = gen_fake(0)
        out = d_on(fake.detach())
    lf = crit(out, False)
    lf.backward()
    bars.act("G(z)", fake, ref[fake_key]), bars.act("D(G(z))", out, ref["d_fake"])
    bars.loss("loss_fake", lf.item(), ref["loss_fake"])
    bars.cos("D-fake", global_cos(_minus(netD, g_real).items(), ref["d_grads_fake"], skip_d))
    both = {k: ref["d_grads_real"][k] + ref["d_grads_fake"][k] for k in ref["d_grads_real"] if k in ref["d_grads_fake"]}
    bars.cos("D-accum", global_cos(netD.named_parameters(), both, skip_d))
    netG.zero_grad(), netD.zero_grad()
    with sc[2]:
        if second_z:
            fake = gen_fake(1)
        out = d_on(fake)
    lg = crit(out, False, True)
    lg.backward()
    bars.act("D(G(z))_g", out, ref["d_g"]), bars.loss("loss_g", lg.item(), ref["loss_g"])
    bars.cos("G-step", global_cos(netG.named_parameters(), ref["g_grads"], tuple(skip_g) + tuple(prebn_biases(netG))))


# ------------------------------------------------------------------------------------------------ cfg 1 / 2: DCGAN-64
@pytest.mark.parametrize("B", [64, 128, 1024])
def test_dcgan64_engine_policy_meets_the_bars_at_benched_batches(B):
    """The precision policy bench.py runs (engine.DcganStep: real pass bf16, D-fake chain fp16, G step bf16x3), pass by
    pass, at the batch sizes the bench uses, plus one whole engine step against the oracle trainer's iteration."""
    from gan_playground_b200 import config
    from gan_playground_b200.criterion import GANLoss
    from gan_playground_b200.engine import DcganStep
    from gan_playground_b200.models import dcgan
    from gan_playground_b200.optim import FusedAdam
    from oracle import gan_oracle as O

    torch.manual_seed(0)
    netG, netD = quiet(lambda: dcgan.Generator()), quiet(lambda: dcgan.Discriminator())
    sd_g, sd_d = _clone_sd(netG), _clone_sd(netD)
    gen = torch.Generator().manual_seed(100 + B)
    x = torch.rand(B, 3, 64, 64, generator=gen) * 2 - 1
    z = torch.randn(2, B, 100, generator=gen)
    torch.set_num_threads(os.cpu_count())
    ref = O.dcgan_step_grads(sd_g, sd_d, x, z[0], z[1])
    netG.cuda(), netD.cuda()
    crit = GANLoss("vanilla", 0.9, 0.1, 0.9).cuda()
    oG = FusedAdam(netG.parameters(), lr=4e-4, betas=(0.5, 0.999))
    oD = FusedAdam(netD.parameters(), lr=1e-4, betas=(0.5, 0.999))
    run = DcganStep(netG, netD, crit, oG, oD, B, 100, torch.device("cuda", 0))
    assert (run.real_precision, run.fake_precision, config.precision()) == ("bf16", "fp16", "bf16x3")
    scopes = (config.precision_scope(run.real_precision), config.precision_scope(run.fake_precision),
              config.precision_scope(config.precision()))
    bars = Bars("cfg2 DCGAN-64 w64 B=%d (engine policy: real bf16 / fake fp16 / G-step bf16x3)" % B)
    xd, zd = x.cuda(), z.cuda()
    _three_passes(bars, netG, netD, crit, ref, lambda: netD(xd), lambda i: netG(zd[i]), lambda img: netD(img), scopes)
    # one whole engine step (both optimiser steps) on fresh copies against the oracle trainer's iteration
    netG.load_state_dict(sd_g), netD.load_state_dict(sd_d)
    netG.zero_grad(), netD.zero_grad()
    tr = O.CpuDcganTrainer(sd_g, sd_d)
    want = tr.step(x, z[0], z[1])
    got = run.step(xd, zd)
    for j, name in enumerate(("lossD_real", "lossD_fake", "lossG")):
        bars.loss("step " + name, got[j], want[j])
    bars.finish()


# ------------------------------------------------------------------------------------------------ cfg 3: SN-DCGAN-32
def test_sn_dcgan32_width64_meets_the_bars():
    from gan_playground_b200.criterion import GANLoss
    from gan_playground_b200.models import dcgan_specnorm as M
    from oracle import gan_oracle as O

    torch.manual_seed(0)
    netG, netD = quiet(lambda: M.Generator(resolution=32)), quiet(lambda: M.Discriminator(resolution=32))
    sd_g, sd_d = _clone_sd(netG), _clone_sd(netD)
    gen = torch.Generator().manual_seed(3)
    B = 64
    x = torch.rand(B, 3, 32, 32, generator=gen) * 2 - 1
    z = torch.randn(2, B, 100, generator=gen)
    torch.set_num_threads(os.cpu_count())
    # same order of forwards as main_dcgan.py's loop: D(x), G(z1), D(fake1), G(z2), D(fake2) — one power iteration each
    ref = O.dcgan_step_grads(sd_g, sd_d, x, z[0], z[1], labels=(1.0, 0.0, 1.0), mode="hinge", sn=True, flatten_head=True)
    netG.cuda(), netD.cuda()
    crit = GANLoss("hinge").cuda()
    bars = Bars("cfg3 SN-DCGAN-32 hinge w64 B=%d (default mode)" % B, loss_abs=0.05)    # hinge G loss sits near zero
    xd, zd = x.cuda(), z.cuda()
    _three_passes(bars, netG, netD, crit, ref, lambda: netD(xd), lambda i: netG(zd[i]), lambda img: netD(img), None,
                  skip_g=())
    bars.finish()


# ------------------------------------------------------------------------------------------------ cfg 4: SNGAN projection
def test_sngan_projection_ch64_meets_the_bars():
    from gan_playground_b200.criterion import GANLoss
    from gan_playground_b200.models import sngan_projection as M
    from oracle import gan_oracle as O

    torch.manual_seed(0)
    netG = quiet(lambda: M.ResNetGenerator(ch=64, dim_z=128, bottom_width=2, img_dim=3, n_classes=10))
    netD = quiet(lambda: M.SNResNetProjectionDiscriminator(ch=64, n_classes=10, img_dim=3))
    sd_g, sd_d = _clone_sd(netG), _clone_sd(netD)
    gen = torch.Generator().manual_seed(4)
    B = 64
    x = torch.rand(B, 3, 32, 32, generator=gen) * 2 - 1
    y = torch.randint(10, (B,), generator=gen)
    z = torch.randn(B, 128, generator=gen)
    c = torch.randint(10, (B,), generator=gen)
    torch.set_num_threads(os.cpu_count())
    ref = O.sngan_step_grads(sd_g, sd_d, x, y, z, c, bottom_width=2)
    netG.cuda(), netD.cuda()
    crit = GANLoss("hinge").cuda()
    bars = Bars("cfg4 SNGAN-projection ch64 32x32 B=%d (default mode -> fp16 operands)" % B, loss_abs=0.05)
    xd, yd, zd, cd = x.cuda(), y.cuda(), z.cuda(), c.cuda()
    _three_passes(bars, netG, netD, crit, ref, lambda: netD(xd, yd), lambda i: netG(zd, cd), lambda img: netD(img, cd), None,
                  skip_g=resnet_g_zero_biases(netG), fake_key="fake", second_z=False)
    bars.finish()


def test_conditional_batchnorm_kernels_match_torch():
    """ConditionalBatchNorm2d (models/sngan_projection.py:6-19) on its own: forward (with and without the fused nearest
    upsample, in every companion format) and backward (dx, embedding gradient) against torch autograd in fp32."""
    import torch.nn.functional as F

    from gan_playground_b200 import ops

    torch.manual_seed(0)
    NB, H, W, C, ncls = 6, 4, 4, 64, 10
    x32 = torch.randn(NB, C, H, W, device="cuda") * 1.7 + 0.3
    emb = torch.randn(ncls, 2 * C, device="cuda") * 0.1
    emb[:, :C] += 1.0
    labels = torch.randint(ncls, (NB,), device="cuda")
    for up in (False, True):
        xr = x32.clone().requires_grad_(True)
        er = emb.clone().requires_grad_(True)
        xh = F.batch_norm(xr, None, None, None, None, True, 0.1, 1e-5)
        e = F.embedding(labels, er)
        ref = F.relu(e[:, :C, None, None] * xh + e[:, C:, None, None])
        if up:
            ref = F.interpolate(ref, scale_factor=2)
        g = torch.randn_like(ref)
        gx, ge = torch.autograd.grad(ref, (xr, er), g)
        x = x32.permute(0, 2, 3, 1).contiguous()
        xb, xhf = x.bfloat16(), x.half()
        for comp, fmt, tol in ((None, ops.COMP_NONE, 1e-2), (xhf, ops.COMP_F16, 2e-3),
                               ((x - xb.float()).bfloat16(), ops.COMP_LO, 1e-4)):
            st = ops.bn_stats_comp(xb, comp) if comp is not None else ops.bn_stats(xb)
            fin = ops.bn_finalize(st, NB * H * W, None, None, None, None, None)
            if fmt:
                out, oc = ops.cbn_apply_act(xb, fin, emb, labels, ops.ACT_RELU, up, comp=comp, out_fmt=fmt)
                val = oc.float() if fmt == ops.COMP_F16 else out.float() + oc.float()
            else:
                val = ops.cbn_apply_act(xb, fin, emb, labels, ops.ACT_RELU, up).float()
            err = (val.permute(0, 3, 1, 2) - ref).abs().max().item() / ref.abs().max().item()
            assert err < tol, (up, fmt, err)
        # backward reads the bf16 tensor in every mode
        st = ops.bn_stats(xb)
        fin = ops.bn_finalize(st, NB * H * W, None, None, None, None, None)
        da = g.permute(0, 2, 3, 1).contiguous().bfloat16()
        S, demb = ops.cbn_bwd_reduce(da, xb, fin, emb, labels, ops.ACT_RELU, up, ncls)
        dx = ops.cbn_bwd_apply(da, xb, fin, emb, labels, S, NB * H * W, ops.ACT_RELU, up)
        cx = F.cosine_similarity(dx.float().permute(0, 3, 1, 2).flatten(), gx.flatten(), dim=0).item()
        ce = F.cosine_similarity(demb.flatten(), ge.flatten(), dim=0).item()
        assert cx > 0.999 and ce > 0.999, (up, cx, ce)


# ------------------------------------------------------------------------------------------------ cfg 5: ACGAN-64
def test_acgan64_width64_meets_the_bars():
    from gan_playground_b200.criterion import GANLoss
    from gan_playground_b200.models import acgan as M
    from oracle import gan_oracle as O

    torch.manual_seed(0)
    netG, netD = quiet(lambda: M.Generator(n_class=10)), quiet(lambda: M.Discriminator(n_class=10))
    sd_g, sd_d = _clone_sd(netG), _clone_sd(netD)
    gen = torch.Generator().manual_seed(5)
    B = 32
    x = torch.rand(B, 3, 64, 64, generator=gen) * 2 - 1
    y = torch.randint(0, 2, (B, 10), generator=gen).float()
    z = torch.randn(B, 100, generator=gen)
    torch.set_num_threads(os.cpu_count())
    ref = O.acgan_step_grads(sd_g, sd_d, x, y, z)
    netG.cuda(), netD.cuda()
    adv_crit = GANLoss("vanilla", 0.9, 0.1, 0.9).cuda()
    mse = torch.nn.MSELoss()
    xd, yd, zd = x.cuda(), y.cuda(), z.cuda()
    aux = {}

    def crit(out, is_real, is_gen=False):          # main_acgan.py:95-97,114-116,129-131 on the two-head output
        return adv_crit(out, is_real, is_gen) + 0.5 * mse(aux["cls"], yd)

    def d_on(img):
        out, aux["cls"] = netD(img)
        return out

    bars = Bars("cfg5 ACGAN-64 w64 B=%d (default mode)" % B)
    _three_passes(bars, netG, netD, crit, ref, lambda: d_on(xd), lambda i: netG(zd, yd), d_on, None, fake_key="fake",
                  second_z=False)
    bars.finish()


# ------------------------------------------------------------------------------------------------ f1: dcgan_blur
def test_dcgan_blur64_width64_meets_the_bars():
    from gan_playground_b200.criterion import GANLoss
    from gan_playground_b200.models import dcgan_blur as M
    from oracle import gan_oracle as O

    torch.manual_seed(0)
    netG, netD = quiet(lambda: M.Generator()), quiet(lambda: M.Discriminator())
    sd_g, sd_d = _clone_sd(netG), _clone_sd(netD)
    gen = torch.Generator().manual_seed(6)
    B = 16
    x = torch.rand(B, 3, 64, 64, generator=gen) * 2 - 1
    z = torch.randn(2, B, 100, generator=gen)
    torch.set_num_threads(os.cpu_count())
    ref = O.dcgan_step_grads(sd_g, sd_d, x, z[0], z[1], blur=True)
    netG.cuda(), netD.cuda()
    crit = GANLoss("vanilla", 0.9, 0.1, 0.9).cuda()
    bars = Bars("f1 dcgan_blur-64 w64 B=%d (default mode bf16x3)" % B)
    xd, zd = x.cuda(), z.cuda()
    _three_passes(bars, netG, netD, crit, ref, lambda: netD(xd), lambda i: netG(zd[i]), lambda img: netD(img), None,
                  skip_g=prebn_biases_blur_g(netG))
    bars.finish()
