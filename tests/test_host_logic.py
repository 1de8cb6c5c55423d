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


def _family_ctors():
    from gan_playground_b200.models import acgan, dcgan_specnorm, dcgan_specnorm_up, sngan_projection as S

    return {
        "dcgan_specnorm_up.Generator": lambda: dcgan_specnorm_up.Generator(ngf=8),
        "dcgan_specnorm_up.Discriminator": lambda: dcgan_specnorm_up.Discriminator(ndf=8),
        "dcgan_specnorm_up.Generator@32": lambda: dcgan_specnorm_up.Generator(ngf=8, resolution=32),
        "dcgan_specnorm_up.Discriminator@128": lambda: dcgan_specnorm_up.Discriminator(ndf=4, resolution=128),
        "dcgan.Generator@128": lambda: dcgan.Generator(ngf=8, resolution=128),
        "dcgan.Discriminator@128": lambda: dcgan.Discriminator(ndf=8, resolution=128),
        "dcgan_specnorm.Generator": lambda: dcgan_specnorm.Generator(ngf=8),
        "dcgan_specnorm.Discriminator": lambda: dcgan_specnorm.Discriminator(ndf=8),
        "dcgan_specnorm.Generator@32": lambda: dcgan_specnorm.Generator(resolution=32),
        "dcgan_specnorm.Discriminator@32": lambda: dcgan_specnorm.Discriminator(resolution=32),
        "sngan_projection.ResNetGenerator": lambda: S.ResNetGenerator(n_classes=10, bottom_width=2),
        "sngan_projection.SNResNetProjectionDiscriminator": lambda: S.SNResNetProjectionDiscriminator(n_classes=10),
        "sngan_projection.ResNetGenerator@uncond": lambda: S.ResNetGenerator(ch=8, n_classes=0),
        "sngan_projection.SNResNetProjectionDiscriminator@uncond": lambda: S.SNResNetProjectionDiscriminator(ch=8, n_classes=0),
        "acgan.Generator": lambda: acgan.Generator(),
        "acgan.Discriminator": lambda: acgan.Discriminator(),
        "acgan.Generator@32c5": lambda: acgan.Generator(ngf=8, resolution=32, n_class=5),
        "acgan.Discriminator@32c5": lambda: acgan.Discriminator(ndf=8, resolution=32, n_class=5),
    }


@pytest.mark.parametrize("name", sorted(_family_ctors()))
def test_state_dict_layout_of_every_family_matches_reference(name, capsys):
    """Keys, shapes and dtypes in state_dict() order for the other constructors / variants of SURVEY.md §8b: resolution 128,
    spectral-norm nets (weight_orig / weight_u / weight_v), the projection pair with and without classes, ACGAN with a
    non-default class count."""
    assert describe(_family_ctors()[name]()) == KEYS[name]


@pytest.mark.parametrize("name,res", [("dcgan_blur.Generator", 64), ("dcgan_blur.Discriminator", 64),
                                      ("dcgan_blur.Generator@32", 32), ("dcgan_blur.Discriminator@32", 32)])
def test_dcgan_blur_state_dict_layout_matches_reference(name, res, capsys):
    """models/dcgan_blur.py (what main_dcgan.py:52-53 instantiates): same keys / shapes / dtypes, incl. the `filt` buffers."""
    from gan_playground_b200.models import dcgan_blur

    net = dcgan_blur.Generator(resolution=res) if "Generator" in name else dcgan_blur.Discriminator(resolution=res)
    assert describe(net) == KEYS[name]


def test_dcgan_blur_same_seed_construction_matches_reference_fixture(capsys):
    """The golden fixture holds the reference's seed-6 state_dicts: the mirror built from the same seed is identical
    (same torch layers in the same order; G's init touches only Linear, as upstream)."""
    from conftest import load_golden
    from gan_playground_b200.models import dcgan_blur

    fx = load_golden("dcgan_blur_r32_w8.pt")
    torch.manual_seed(6)
    g = dcgan_blur.Generator(z_dim=fx["z_dim"], ngf=fx["width"], resolution=fx["res"])
    d = dcgan_blur.Discriminator(ndf=fx["width"], resolution=fx["res"])
    for net, sd in ((g, fx["sd_g"]), (d, fx["sd_d"])):
        mine = net.state_dict()
        assert list(mine.keys()) == list(sd.keys())
        assert all(torch.equal(mine[k], sd[k]) for k in sd)
    with pytest.raises(KeyError):
        dcgan_blur.Generator(resolution=48)


def test_root_level_shims_export_dcgan_blur_and_blurpool():
    """main_dcgan.py:11 does `from models import dcgan, dcgan_specnorm, dcgan_blur`; dcgan_blur.py:5 `from models.ops import BlurPool2d`."""
    import importlib

    m = importlib.import_module("models.dcgan_blur")
    o = importlib.import_module("models.ops")
    assert hasattr(m, "Generator") and hasattr(m, "Discriminator") and hasattr(o, "BlurPool2d")
    b = o.BlurPool2d(channels=8, stride=1)
    assert b.filt.shape == (8, 1, 3, 3) and abs(b.filt.sum().item() - 8.0) < 1e-6


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


def test_forward_tile_cost_model_fills_the_machine(built_lib):
    """Host-side tile selection of gp_conv_fwd (no GPU needed): full-batch DCGAN-64 layers get the 256-wide tiles,
    the long-K layer the 256x256 one, narrow outputs 256-row tiles, and 128-image shards (8-GPU strong scaling) the
    narrower tiles that still give every SM work."""
    from gan_playground_b200 import ops

    K4, CT, K1 = ops.KIND_CONV_K4S2, ops.KIND_CONVT_K4S2, ops.KIND_CONV_K1S1
    assert ops.conv_fwd_plan(1024, 32, 32, 128, 16, 16, 256, K4) == (256, 1, 2048)       # D block 1
    assert ops.conv_fwd_plan(1024, 8, 8, 512, 4, 4, 1024, K4) == (256, 2, 256)           # D block 3: K = 8192
    assert ops.conv_fwd_plan(1024, 8, 8, 512, 4, 4, 1024, K4, x3=True)[:2] == (256, 1)   # no 256x256 bf16x3 variant
    assert ops.conv_fwd_plan(1024, 16, 16, 256, 32, 32, 128, CT) == (128, 2, 4096)       # G block 2: N = 128
    assert ops.conv_fwd_plan(1024, 32, 32, 64, 32, 32, 128, K1)[:2] == (128, 2)          # image-side GEMM
    bn, mt, tiles = ops.conv_fwd_plan(128, 8, 8, 512, 4, 4, 1024, K4)                    # D block 3 at 128 images
    assert (bn, mt) == (128, 1) and 100 <= tiles <= 148
    for shape in ((128, 32, 32, 128, 16, 16, 256, K4), (128, 4, 4, 1024, 8, 8, 512, CT), (16, 16, 16, 64, 8, 8, 64, K4)):
        bn, mt, tiles = ops.conv_fwd_plan(*shape)
        assert bn in (64, 128, 256) and mt in (1, 2) and tiles >= 1
        assert bn // 2 < shape[6] or bn == 64                                            # never mostly padding


@pytest.mark.parametrize("family", ["sngan_projection", "sngan_projection@uncond", "acgan", "dcgan_specnorm",
                                    "dcgan_specnorm_up"])
def test_same_seed_construction_of_every_family_matches_reference(family, capsys):
    """The mirrors create their torch parameter holders, apply the initialisers and wrap spectral norm in the reference's
    order, so a construction under the same seed consumes the same RNG stream: every parameter AND buffer (spectral-norm
    u / v included) is identical to the unmodified reference's — (sum, first element) probes of each state_dict entry."""
    from gan_playground_b200.models import acgan, dcgan_specnorm, dcgan_specnorm_up, sngan_projection as S

    build = {
        "dcgan_specnorm_up": lambda: (dcgan_specnorm_up.Generator(z_dim=16, ngf=8, resolution=32),
                                      dcgan_specnorm_up.Discriminator(ndf=8, resolution=32)),
        "sngan_projection": lambda: (S.ResNetGenerator(ch=8, dim_z=16, bottom_width=2, n_classes=10),
                                     S.SNResNetProjectionDiscriminator(ch=8, n_classes=10)),
        "sngan_projection@uncond": lambda: (S.ResNetGenerator(ch=8, dim_z=16, n_classes=0),
                                            S.SNResNetProjectionDiscriminator(ch=8, n_classes=0)),
        "acgan": lambda: (acgan.Generator(z_dim=16, ngf=8, n_class=10), acgan.Discriminator(ndf=8, n_class=10)),
        "dcgan_specnorm": lambda: (dcgan_specnorm.Generator(z_dim=16, ngf=8, resolution=32),
                                   dcgan_specnorm.Discriminator(ndf=8, resolution=32)),
    }[family]
    torch.manual_seed(7)
    nets = build()
    want = KEYS["same_seed_probes"][family]
    for net, probes in zip(nets, want):
        sd = net.state_dict()
        assert list(sd.keys()) == [k for k, _, _ in probes]
        for k, total, first in probes:
            assert float(sd[k].double().sum()) == total and float(sd[k].flatten()[0]) == first, k
