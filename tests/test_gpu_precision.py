"""The two forward precision modes against the fp32 CPU oracle at the BASELINE width (DCGAN-64, ngf=ndf=64, batch 32),
with BASELINE.json's north_star bars: per-boundary activation max-rel-error <= 1e-2, gradient cosine >= 0.999, losses
within 2 %.  "bf16x3" (default) must meet all of them; "bf16" is the fast mode whose G-step cosine is bounded near 0.97
by forward operand rounding — its thresholds document what it reaches."""
import os

import pytest
import torch

from test_gpu_dcgan import global_cos, quiet, relerr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def reference_step():
    from gan_playground_b200.models import dcgan
    from oracle import gan_oracle as O

    torch.manual_seed(0)
    netG, netD = quiet(lambda: dcgan.Generator()), quiet(lambda: dcgan.Discriminator())
    sd_g = {k: v.clone() for k, v in netG.state_dict().items()}
    sd_d = {k: v.clone() for k, v in netD.state_dict().items()}
    gen = torch.Generator().manual_seed(1)
    B = 32
    x = torch.rand(B, 3, 64, 64, generator=gen) * 2 - 1
    z1, z2 = torch.randn(B, 100, generator=gen), torch.randn(B, 100, generator=gen)
    torch.set_num_threads(os.cpu_count())
    ref = O.dcgan_step_grads(sd_g, sd_d, x, z1, z2)
    return sd_g, sd_d, x, z1, z2, ref


BARS = {
    #            act    cos D-real  cos D-fake  cos G-step
    "bf16x3": (1e-2, 0.999, 0.999, 0.999),
    "bf16": (2e-2, 0.999, 0.99, 0.95),
    # one MMA on fp16 operands: meets every bar except the G step's (emulated 0.99677, tools/precision_study.py)
    "fp16": (5e-3, 0.999, 0.999, 0.99),
}


@pytest.mark.parametrize("mode", ["bf16x3", "bf16", "fp16"])
def test_dcgan64_step_meets_precision_bars(mode, reference_step):
    from gan_playground_b200 import config
    from gan_playground_b200.criterion import GANLoss
    from gan_playground_b200.models import dcgan

    sd_g, sd_d, x, z1, z2, ref = reference_step
    act_bar, c_real, c_fake, c_g = BARS[mode]
    prev = config.precision()
    config.set_precision(mode)
    try:
        netG, netD = quiet(lambda: dcgan.Generator()).cuda(), quiet(lambda: dcgan.Discriminator()).cuda()
        netG.load_state_dict(sd_g), netD.load_state_dict(sd_d)
        crit = GANLoss("vanilla", 0.9, 0.1, 0.9).cuda()
        out = netD(x.cuda())
        loss = crit(out, True)
        loss.backward()
        e1 = relerr(out, ref["d_real"])
        g1 = global_cos(netD.named_parameters(), ref["d_grads_real"])
        l1 = abs(loss.item() - ref["loss_real"].item()) / ref["loss_real"].item()
        fake = netG(z1.cuda())
        e2 = relerr(fake, ref["fake1"])
        netD.zero_grad()
        out = netD(fake.detach())
        lf = crit(out, False)
        lf.backward()
        e3 = relerr(out, ref["d_fake"])
        g2 = global_cos(netD.named_parameters(), ref["d_grads_fake"])
        netG.zero_grad(), netD.zero_grad()
        out = netD(netG(z2.cuda()))
        lg = crit(out, False, True)
        lg.backward()
        e4 = relerr(out, ref["d_g"])
        g3 = global_cos(netG.named_parameters(), ref["g_grads"])
        l3 = abs(lg.item() - ref["loss_g"].item()) / ref["loss_g"].item()
        print("\n[%s] act err: D(x) %.2e  G(z) %.2e  D(G(z1)) %.2e  D(G(z2)) %.2e | cos: D-real %.6f  D-fake %.6f  G-step %.6f"
              " | loss rel: %.2e %.2e" % (mode, e1, e2, e3, e4, g1, g2, g3, l1, l3))
        assert max(e1, e2, e3, e4) < act_bar
        assert g1 > c_real and g2 > c_fake and g3 > c_g
        assert l1 < 0.02 and l3 < 0.02
    finally:
        config.set_precision(prev)


def test_mixed_policy_real_pass_in_bf16_meets_the_bars(reference_step):
    """engine.DcganStep's policy: the real-image discriminator pass runs plain bf16 operands
    (config.precision_scope("bf16")), every pass that sees generated images runs bf16x3. The real pass alone and the
    accumulated D gradient of the step (real + fake, what optD.step() consumes) must both meet the north_star bars."""
    from gan_playground_b200 import config
    from gan_playground_b200.criterion import GANLoss
    from gan_playground_b200.models import dcgan

    sd_g, sd_d, x, z1, z2, ref = reference_step
    prev = config.precision()
    config.set_precision("bf16x3")
    try:
        netG, netD = quiet(lambda: dcgan.Generator()).cuda(), quiet(lambda: dcgan.Discriminator()).cuda()
        netG.load_state_dict(sd_g), netD.load_state_dict(sd_d)
        crit = GANLoss("vanilla", 0.9, 0.1, 0.9).cuda()
        with config.precision_scope("bf16"):
            out = netD(x.cuda())
        assert config.precision() == "bf16x3"
        loss = crit(out, True)
        loss.backward()
        e1 = relerr(out, ref["d_real"])
        g1 = global_cos(netD.named_parameters(), ref["d_grads_real"])
        out = netD(netG(z1.cuda()).detach())
        crit(out, False).backward()                      # accumulates onto the real-pass gradients
        both = {k: ref["d_grads_real"][k] + ref["d_grads_fake"][k] for k in ref["d_grads_real"]}
        g12 = global_cos(netD.named_parameters(), both)
        print("\n[mixed] D(x) act err %.2e | cos: D-real (bf16 pass) %.6f  D real+fake accumulated %.6f" % (e1, g1, g12))
        assert e1 < 1e-2 and g1 > 0.999 and g12 > 0.999
        assert abs(loss.item() - ref["loss_real"].item()) < 0.02 * ref["loss_real"].item()
    finally:
        config.set_precision(prev)


def test_fp16_fake_chain_policy_meets_the_bars(reference_step):
    """engine.DcganStep(fake_precision="fp16"): the D-fake chain G(z1) -> D(G(z1).detach()) runs ONE MMA on fp16 operands,
    the real pass plain bf16, only the G step bf16x3. What optD.step() consumes — the accumulated real + fake D gradient —
    and the D-fake pass alone must meet the north_star bars; the G step is untouched by the policy."""
    from gan_playground_b200 import config
    from gan_playground_b200.criterion import GANLoss
    from gan_playground_b200.models import dcgan

    sd_g, sd_d, x, z1, z2, ref = reference_step
    prev = config.precision()
    config.set_precision("bf16x3")
    try:
        netG, netD = quiet(lambda: dcgan.Generator()).cuda(), quiet(lambda: dcgan.Discriminator()).cuda()
        netG.load_state_dict(sd_g), netD.load_state_dict(sd_d)
        crit = GANLoss("vanilla", 0.9, 0.1, 0.9).cuda()
        with config.precision_scope("bf16"):
            out = netD(x.cuda())
        crit(out, True).backward()
        g_real = {k: p.grad.detach().clone() for k, p in netD.named_parameters() if p.grad is not None}
        with config.precision_scope("fp16"):
            fake = netG(z1.cuda())
            out = netD(fake.detach())
        lf = crit(out, False)
        lf.backward()                                     # accumulates onto the real-pass gradients
        e2, e3 = relerr(fake, ref["fake1"]), relerr(out, ref["d_fake"])
        both = {k: ref["d_grads_real"][k] + ref["d_grads_fake"][k] for k in ref["d_grads_real"]}
        g12 = global_cos(netD.named_parameters(), both)
        only_fake = {k: p.grad.detach() - g_real[k] for k, p in netD.named_parameters() if k in g_real}

        class _P:   # global_cos reads .grad
            def __init__(self, g):
                self.grad = g
        g2 = global_cos({k: _P(v) for k, v in only_fake.items()}.items(), ref["d_grads_fake"])
        l2 = abs(lf.item() - ref["loss_fake"].item()) / ref["loss_fake"].item()
        print("\n[fp16 fake chain] act err G(z) %.2e D(G(z)) %.2e | cos: D-fake %.6f  D real+fake accumulated %.6f | loss rel %.2e"
              % (e2, e3, g2, g12, l2))
        assert max(e2, e3) < 1e-2 and g2 > 0.999 and g12 > 0.999 and l2 < 0.02
        # the G step afterwards runs in the global bf16x3 mode on the same nets (fp16-staged weights do not leak into it)
        netG.zero_grad(), netD.zero_grad()
        out = netD(netG(z2.cuda()))
        crit(out, False, True).backward()
        g3 = global_cos(netG.named_parameters(), ref["g_grads"])
        print("[fp16 fake chain] G step after it (bf16x3): cos %.6f" % g3)
        assert g3 > 0.999
    finally:
        config.set_precision(prev)
