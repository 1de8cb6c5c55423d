"""Checkpoint compatibility: a checkpoint written by the UNMODIFIED reference's save_model
(main_dcgan.py:107-123 / main_sngan.py:107-123 — {'state_dict': {'generator', 'discriminator'}, 'optimizer': {...},
'epoch'}), stored in tests/golden/checkpoint_ref_layout.pt by oracle/make_golden.py, loads into the mirrors and their
optimisers, training resumes from it with the reference's next losses, and a checkpoint written from the mirrors has the
same layout (so the reference loads it back). Construction / state_dict handling is host logic: no GPU needed; the
resumed GPU iteration is checked in tests/test_gpu_zz_loops.py."""
import contextlib
import io

import pytest
import torch

from conftest import load_golden
from oracle import gan_oracle as O


def quiet(fn):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn()


def _pairs(fx):
    from gan_playground_b200.models import dcgan, sngan_projection as S

    return {
        "dcgan": (quiet(lambda: dcgan.Generator(z_dim=16, ngf=4, resolution=32)),
                  quiet(lambda: dcgan.Discriminator(ndf=4, resolution=32)), fx["dcgan"]["checkpoint"]),
        "sngan": (S.ResNetGenerator(ch=4, dim_z=16, bottom_width=2, img_dim=3, n_classes=10),
                  S.SNResNetProjectionDiscriminator(ch=4, n_classes=10, img_dim=3), fx["sngan"]["checkpoint"]),
    }


def _layout(obj):
    """Nested key / shape / dtype structure of a checkpoint object (values dropped)."""
    if torch.is_tensor(obj):
        return (tuple(obj.shape), str(obj.dtype))
    if isinstance(obj, dict):
        return {k: _layout(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return [_layout(v) for v in obj]
    return type(obj).__name__


@pytest.mark.parametrize("family", ["dcgan", "sngan"])
def test_reference_checkpoint_loads_and_round_trips(family, tmp_path):
    netG, netD, ck = _pairs(load_golden("checkpoint_ref_layout.pt"))[family]
    assert sorted(ck) == ["epoch", "optimizer", "state_dict"]
    for net, key in ((netG, "generator"), (netD, "discriminator")):
        missing, unexpected = net.load_state_dict(ck["state_dict"][key], strict=True)
        assert not missing and not unexpected
        for k, v in net.state_dict().items():
            assert torch.equal(v, ck["state_dict"][key][k]), k
    # the reference's torch.optim.Adam state maps onto the mirrors' parameters one to one, in order
    optG = torch.optim.Adam(netG.parameters(), lr=1.0)
    optD = torch.optim.Adam(netD.parameters(), lr=1.0)
    optG.load_state_dict(ck["optimizer"]["generator"])
    optD.load_state_dict(ck["optimizer"]["discriminator"])
    for opt, net in ((optG, netG), (optD, netD)):
        ps = list(net.parameters())
        assert len(opt.state) == len(ps)
        for p in ps:
            assert opt.state[p]["exp_avg"].shape == p.shape and float(opt.state[p]["step"]) >= 1
    assert optG.param_groups[0]["lr"] == ck["optimizer"]["generator"]["param_groups"][0]["lr"]
    # written back in the reference's layout: identical structure (keys, shapes, dtypes), loadable with weights_only
    mine = {"state_dict": {"generator": netG.state_dict(), "discriminator": netD.state_dict()},
            "optimizer": {"generator": optG.state_dict(), "discriminator": optD.state_dict()}, "epoch": ck["epoch"]}
    path = tmp_path / "checkpoint_001.pth"
    torch.save(mine, path)
    back = torch.load(path, weights_only=True)
    assert _layout(back) == _layout(ck)


def test_oracle_trainer_resumes_from_the_reference_checkpoint():
    """The CPU restatement continues the reference's run: parameters, BatchNorm buffers and Adam moments from the
    checkpoint give the reference's next-iteration losses."""
    import os
    import sys

    from conftest import ROOT

    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from make_golden import trace_data

    fx = load_golden("checkpoint_ref_layout.pt")["dcgan"]
    ck = fx["checkpoint"]
    tr = O.CpuDcganTrainer(ck["state_dict"]["generator"], ck["state_dict"]["discriminator"])
    tr.opt_g.load_state_dict(ck["optimizer"]["generator"])
    tr.opt_d.load_state_dict(ck["optimizer"]["discriminator"])
    xs, zs = trace_data(fx["seed"], 3, 8, 32, 16)
    got = tr.step(xs[2], zs[2, 0], zs[2, 1])[:3]
    assert got == pytest.approx(fx["next_losses"], abs=2e-5)
