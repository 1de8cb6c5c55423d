"""Host-side plumbing of every network family in every forward precision mode WITHOUT a GPU: the C-ABI entry points are
replaced by a checker that validates each call's argument list against the binding table (count and scalar / pointer
kinds, i.e. what ctypes would marshal) and launches nothing; outputs are whatever torch.empty allocated. What this
covers is everything between `module(x)` and the kernel calls — autograd node signatures (number of gradients returned),
companion-tensor hand-over between nodes (functional_resnet.py), mode scopes, weight-staging caches, the step drivers —
so that a GPU run is not spent on a Python arity error. Numerics are the GPU tests' business."""
import contextlib
import ctypes
import io
import os
import re

import pytest
import torch

from conftest import ROOT


def quiet(fn):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn()


@pytest.fixture()
def fake_kernels(monkeypatch):
    from gan_playground_b200 import criterion, ops
    from gan_playground_b200.models import _common

    calls = []
    monkeypatch.setattr(criterion, "_require_cuda", lambda t, what: None)

    def fake_fn(name):
        sig = ops._SIGS[name]

        def call(*args):
            assert len(args) == len(sig), "%s: %d arguments for a %d-argument entry point" % (name, len(args), len(sig))
            for i, (a, t) in enumerate(zip(args, sig)):
                if t is ctypes.c_void_p:
                    assert a is None or isinstance(a, int), "%s arg %d: pointer expected, got %r" % (name, i, type(a))
                elif t in (ctypes.c_int, ctypes.c_longlong):
                    assert isinstance(a, int) and not isinstance(a, bool), "%s arg %d: int expected, got %r" % (name, i, a)
                else:
                    assert isinstance(a, (int, float)), "%s arg %d: float expected, got %r" % (name, i, a)
            calls.append(name)
            return 0

        return call

    def chk(t, dtype, name):
        if t.dtype != dtype:
            raise ops._lib.GpError("%s must be %s, got %s" % (name, dtype, t.dtype))
        if not t.is_contiguous():
            raise ops._lib.GpError("%s must be contiguous" % name)

    monkeypatch.setattr(ops, "_fn", fake_fn)
    monkeypatch.setattr(ops, "_chk", chk)
    monkeypatch.setattr(ops, "_stream", lambda: 0)
    monkeypatch.setattr(_common, "require_cuda", lambda t, what: None)
    for mod in ("dcgan", "dcgan_specnorm", "dcgan_specnorm_up", "dcgan_blur", "acgan", "sngan_projection"):
        m = __import__("gan_playground_b200.models." + mod, fromlist=["x"])
        if hasattr(m, "require_cuda"):
            monkeypatch.setattr(m, "require_cuda", lambda t, what: None)
    return calls


def _all_grads(net):
    missing = [k for k, p in net.named_parameters() if p.requires_grad and p.grad is None]
    wrong = [k for k, p in net.named_parameters() if p.grad is not None and p.grad.shape != p.shape]
    return missing, wrong


def _step(netG, netD, crit, d_args, g_args, second_g=True):
    """D-real, D-fake on detached fakes, G step: every parameter of D then of G must receive a gradient of its shape."""
    netG.train(), netD.train()
    out = netD(*d_args)
    out = out[0] if isinstance(out, tuple) else out
    crit(out, True).backward()
    fake = netG(*g_args)
    assert fake.dtype == torch.float32 and fake.dim() == 4
    out = netD(fake.detach(), *d_args[1:])
    out = out[0] if isinstance(out, tuple) else out
    crit(out, False).backward()
    missing, wrong = _all_grads(netD)
    netG.zero_grad(), netD.zero_grad()
    if second_g:
        fake = netG(*g_args)
    out = netD(fake, *d_args[1:])
    out = out[0] if isinstance(out, tuple) else out
    crit(out, False, True).backward()
    mg, wg = _all_grads(netG)
    return missing + mg, wrong + wg


@pytest.mark.parametrize("mode", ["bf16", "fp16", "bf16x3"])
@pytest.mark.parametrize("family", ["dcgan", "dcgan_specnorm", "dcgan_specnorm_up", "dcgan_blur", "sngan_projection",
                                    "sngan_unconditional"])
def test_forward_backward_plumbing(fake_kernels, mode, family):
    from gan_playground_b200 import config
    from gan_playground_b200.criterion import GANLoss
    from gan_playground_b200.models import dcgan, dcgan_blur, dcgan_specnorm, dcgan_specnorm_up, sngan_projection

    torch.manual_seed(0)
    B = 4
    prev = config.precision()
    config.set_precision(mode)
    try:
        if family == "sngan_projection":
            netG = sngan_projection.ResNetGenerator(ch=8, dim_z=16, bottom_width=2, n_classes=10)
            netD = sngan_projection.SNResNetProjectionDiscriminator(ch=8, n_classes=10)
            y = torch.randint(10, (B,))
            missing, wrong = _step(netG, netD, GANLoss("hinge"), (torch.randn(B, 3, 32, 32), y), (torch.randn(B, 16), y),
                                   second_g=False)
            # the zero-gradient conv biases of the generator still get (noise) gradients; l_y rows are gathered
        elif family == "sngan_unconditional":
            netG = sngan_projection.ResNetGenerator(ch=8, dim_z=16, bottom_width=2, n_classes=0)
            netD = sngan_projection.SNResNetProjectionDiscriminator(ch=8, n_classes=0)
            missing, wrong = _step(netG, netD, GANLoss("hinge"), (torch.randn(B, 3, 32, 32),), (torch.randn(B, 16), None),
                                   second_g=False)
        else:
            M = {"dcgan": dcgan, "dcgan_specnorm": dcgan_specnorm, "dcgan_specnorm_up": dcgan_specnorm_up,
                 "dcgan_blur": dcgan_blur}[family]
            netG = quiet(lambda: M.Generator(z_dim=16, ngf=8, resolution=32))
            netD = quiet(lambda: M.Discriminator(ndf=8, resolution=32))
            missing, wrong = _step(netG, netD, GANLoss("vanilla", 0.9, 0.1, 0.9), (torch.randn(B, 3, 32, 32),),
                                   (torch.randn(B, 16),))
        # biases in front of a BatchNorm get no gradient tensor at all in the DCGAN-family nodes (analytically zero)
        missing = [k for k in missing if not k.endswith(".0.bias")]
        assert not missing, "no gradient for %s" % missing
        assert not wrong, wrong
        assert "gp_conv_fwd" in fake_kernels and "gp_conv_wgrad" in fake_kernels
        assert config.precision() == mode                      # every scope was left
    finally:
        config.set_precision(prev)


def test_acgan_two_heads_plumbing(fake_kernels):
    from gan_playground_b200.criterion import GANLoss
    from gan_playground_b200.models import acgan

    netG = quiet(lambda: acgan.Generator(z_dim=16, ngf=8, n_class=10))
    netD = quiet(lambda: acgan.Discriminator(ndf=8, n_class=10))
    y = torch.randint(0, 2, (4, 10)).float()
    crit, mse = GANLoss("vanilla", 0.9, 0.1, 0.9), torch.nn.MSELoss()
    adv, cls = netD(torch.randn(4, 3, 64, 64))
    (crit(adv, True) + 0.5 * mse(cls, y)).backward()
    fake = netG(torch.randn(4, 16), y)
    adv, cls = netD(fake)
    (crit(adv, False, True) + 0.5 * mse(cls, y)).backward()
    missing, wrong = _all_grads(netG)
    assert not [k for k in missing if not k.endswith(".0.bias")] and not wrong


@pytest.mark.parametrize("driver", ["dcgan", "sngan", "acgan"])
def test_step_drivers_plumbing(fake_kernels, driver, monkeypatch):
    """One eager step of each step driver on CPU stand-in kernels (FusedAdam's update kernel included)."""
    from gan_playground_b200.criterion import GANLoss
    from gan_playground_b200.engine import AcganStep, DcganStep, SnganStep
    from gan_playground_b200.models import acgan, dcgan, sngan_projection
    from gan_playground_b200.optim import FusedAdam

    dev, B = torch.device("cpu"), 4
    if driver == "dcgan":
        netG, netD = quiet(lambda: dcgan.Generator(z_dim=16, ngf=8, resolution=32)), quiet(lambda: dcgan.Discriminator(ndf=8, resolution=32))
        oG, oD = FusedAdam(netG.parameters(), lr=4e-4, betas=(0.5, 0.999)), FusedAdam(netD.parameters(), lr=1e-4, betas=(0.5, 0.999))
        run = DcganStep(netG, netD, GANLoss("vanilla", 0.9, 0.1, 0.9), oG, oD, B, 16, dev)
        assert (run.real_precision, run.fake_precision) == ("bf16", "fp16")
        out = run.step(torch.randn(B, 3, 32, 32))
    elif driver == "sngan":
        netG = sngan_projection.ResNetGenerator(ch=8, dim_z=16, bottom_width=2, n_classes=10)
        netD = sngan_projection.SNResNetProjectionDiscriminator(ch=8, n_classes=10)
        oG, oD = FusedAdam(netG.parameters(), lr=2e-4, betas=(0.0, 0.999)), FusedAdam(netD.parameters(), lr=2e-4, betas=(0.0, 0.999))
        run = SnganStep(netG, netD, GANLoss("hinge"), oG, oD, B, 16, dev, n_classes=10, n_disc_update=1)
        out = run.step(torch.randn(B, 3, 32, 32), torch.randint(10, (B,)))
    else:
        netG, netD = quiet(lambda: acgan.Generator(z_dim=16, ngf=8, n_class=10)), quiet(lambda: acgan.Discriminator(ndf=8, n_class=10))
        oG, oD = FusedAdam(netG.parameters(), lr=4e-4, betas=(0.5, 0.999)), FusedAdam(netD.parameters(), lr=1e-4, betas=(0.5, 0.999))
        run = AcganStep(netG, netD, GANLoss("vanilla", 0.9, 0.1, 0.9), oG, oD, B, 16, dev)
        out = run.step(torch.randn(B, 3, 64, 64), torch.randint(0, 2, (B, 10)).float())
    assert len(out) == run.N_SCALARS
    assert "gp_adam_flat" in fake_kernels and fake_kernels.count("gp_adam_flat") == 2


def test_binding_table_matches_header_prototypes():
    """Every entry of ops._SIGS has as many arguments as its prototype in include/gpb200.h, pointers where the header has
    pointers and scalars where it has scalars."""
    from gan_playground_b200 import ops

    text = open(os.path.join(ROOT, "include", "gpb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = dict(re.findall(r"\bint\s+(gp_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S))
    assert len(protos) >= 55
    for name, sig in ops._SIGS.items():
        assert name in protos, name
        params = [p.strip() for p in protos[name].split(",")]
        assert len(params) == len(sig), "%s: header has %d parameters, binding table %d" % (name, len(params), len(sig))
        for p, t in zip(params, sig):
            is_ptr = "*" in p
            assert is_ptr == (t is ctypes.c_void_p), "%s: parameter `%s` vs binding %s" % (name, p, t)
            if not is_ptr:
                want = (ctypes.c_longlong if "long long" in p else ctypes.c_double if "double" in p else
                        ctypes.c_float if "float" in p else ctypes.c_int)
                assert t is want, "%s: parameter `%s` bound as %s" % (name, p, t)


def test_backward_links_replace_the_standalone_passes(fake_kernels):
    """functional.BwdLink (config.bwd_fusion, off by default — it measured slower): with the fusion on, the data-gradient GEMMs take over the activation-derivative pass of the
    blocks without BatchNorm and the BatchNorm-backward reduction of every block that feeds another GEMM block; only the
    block in front of the head keeps its own reduction. Off: the standalone kernels run."""
    from gan_playground_b200 import config
    from gan_playground_b200.criterion import GANLoss
    from gan_playground_b200.models import dcgan

    prev = config.bwd_fusion()

    def counts(on):
        config.set_bwd_fusion(on)
        try:
            torch.manual_seed(0)
            netG = quiet(lambda: dcgan.Generator(z_dim=16, ngf=8, resolution=32))
            netD = quiet(lambda: dcgan.Discriminator(ndf=8, resolution=32))
            del fake_kernels[:]
            missing, wrong = _step(netG, netD, GANLoss("vanilla", 0.9, 0.1, 0.9), (torch.randn(4, 3, 32, 32),),
                                   (torch.randn(4, 16),))
            assert not [k for k in missing if not k.endswith(".0.bias")] and not wrong
            red = sum(fake_kernels.count(k) for k in ("gp_bn_bwd_reduce", "gp_bn_bwd_reduce_f32", "gp_bn_bwd_reduce_comp"))
            return fake_kernels.count("gp_act_bwd"), red
        finally:
            config.set_bwd_fusion(prev)

    act_off, red_off = counts(False)
    act_on, red_on = counts(True)
    # resolution 32: D = image block + 2 BN blocks (3 passes), G = linear + 2 BN blocks + image layer (G step only)
    assert act_off == 3 + 1 and red_off == 3 * 2 + 2          # D block 0 x 3 passes + G's linear; every BN block
    assert act_on == 0 and red_on == 3 * 1                     # only D's last block (in front of the head) reduces itself
