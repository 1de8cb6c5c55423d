"""Host-side mirror of the reference's module API: state_dict layout, RNG order, attributes, error behaviour."""
import json
import os

import pytest
import torch

from conftest import GOLDEN
from gan_playground_b200 import _lib
from gan_playground_b200.criterion import GANLoss
from gan_playground_b200.models import dcgan

KEYS = json.load(open(os.path.join(GOLDEN, "state_dict_keys.json")))


def describe(net):
    return [[k, list(v.shape), str(v.dtype)] for k, v in net.state_dict().items()]


@pytest.mark.parametrize("name,ctor", [
    ("dcgan.Generator", lambda: dcgan.Generator()),
    ("dcgan.Discriminator", lambda: dcgan.Discriminator()),
    ("dcgan.Generator@32", lambda: dcgan.Generator(resolution=32)),
    ("dcgan.Discriminator@32", lambda: dcgan.Discriminator(resolution=32)),
])
def test_state_dict_layout_matches_reference(name, ctor, capsys):
    assert describe(ctor()) == KEYS[name]


def test_same_seed_gives_same_weights_as_reference(capsys):
    probe = KEYS["seed0_probe"]
    torch.manual_seed(0)
    g = dcgan.Generator(ngf=8, resolution=32)
    d = dcgan.Discriminator(ndf=8, resolution=32)
    assert g.linear.weight.flatten()[:8].tolist() == probe["g_linear_w0"]
    assert g.blocks[0][0].weight.flatten()[:8].tolist() == probe["g_blocks0_w0"]
    assert d.blocks[0][0].weight.flatten()[:8].tolist() == probe["d_blocks0_w0"]
    assert d.out_layer.weight.flatten()[:8].tolist() == probe["d_out_w0"]
    assert g.param_count == probe["g_param_count"] and d.param_count == probe["d_param_count"]


def test_constructor_prints_param_count(capsys):
    dcgan.Generator(ngf=8, resolution=32)
    out = capsys.readouterr().out
    assert out.startswith("Param count for Gs initialized parameters: ")


def test_unknown_resolution_is_keyerror():
    with pytest.raises(KeyError):
        dcgan.Generator(resolution=48)
    with pytest.raises(KeyError):
        dcgan.Discriminator(resolution=48)


def test_unknown_init_only_prints(capsys):
    dcgan.Discriminator(ndf=8, resolution=32, init="bogus")
    assert "Init style not recognized..." in capsys.readouterr().out


def test_skip_init_keeps_torch_defaults():
    torch.manual_seed(0)
    a = dcgan.Generator(ngf=8, resolution=32, skip_init=True)
    assert not hasattr(a, "param_count")


def test_attributes():
    g = dcgan.Generator(ngf=16, resolution=64)
    assert g.arch == {"in_channels": [256, 128, 64], "out_channels": [128, 64, 32]}
    assert (g.z_dim, g.ngf, g.img_dim, g.resolution, g.bottom_width, g.init) == (100, 16, 3, 64, 4, "N02")
    d = dcgan.Discriminator(ndf=16, resolution=64)
    assert d.arch == {"in_channels": [3, 32, 64, 128], "out_channels": [32, 64, 128, 256]}
    assert dcgan.G_arch(64)[128]["out_channels"] == [512, 256, 128, 64]
    assert dcgan.D_arch(64, 1)[32]["in_channels"] == [1, 128, 256]


def test_cpu_input_fails_loudly_no_fallback():
    g = dcgan.Generator(ngf=8, resolution=32)
    with pytest.raises(_lib.GpError):
        g(torch.randn(2, 100))
    d = dcgan.Discriminator(ndf=8, resolution=32)
    with pytest.raises(_lib.GpError):
        d(torch.randn(2, 3, 32, 32))
    with pytest.raises(_lib.GpError):
        GANLoss("hinge")(torch.randn(4, 1), True)


def test_ganloss_api():
    fx = torch.load(os.path.join(GOLDEN, "ganloss.pt"), weights_only=False)
    crit = GANLoss("vanilla", 0.9, 0.1, 0.9)
    assert sorted(crit.state_dict().keys()) == fx["buffers"]
    assert crit.gan_mode == "vanilla"
    with pytest.raises(NotImplementedError):
        GANLoss("wgan")
    sd = GANLoss("vanilla", 0.7, 0.2, 0.6).state_dict()
    crit.load_state_dict(sd)
    assert crit._host_labels == pytest.approx((0.7, 0.2, 0.6))


def test_parameters_are_fp32_leaf_params_usable_by_adam():
    g = dcgan.Generator(ngf=8, resolution=32)
    ps = list(g.parameters())
    assert all(p.dtype == torch.float32 and p.is_leaf and p.requires_grad for p in ps)
    torch.optim.Adam(ps, lr=4e-4, betas=(0.5, 0.999))
    # only reference keys are in the state dict (the weight cache is not serialised)
    assert not any("cache" in k for k in g.state_dict())
