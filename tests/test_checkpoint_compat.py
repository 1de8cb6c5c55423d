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


def test_fused_adam_checkpoint_steps_in_torch_adam_and_back():
    """a checkpoint written by optim.FusedAdam must load into the reference scripts' torch.optim.Adam AND
    step there (torch's load_state_dict replaces the param groups wholesale, so every torch.optim.Adam group key has to
    be in the file; `step` is a CPU scalar like torch's own), and the other way round. Host logic only: no kernel runs."""
    from gan_playground_b200.optim import FusedAdam

    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Linear(5, 4))
    ref_opt = torch.optim.Adam(net.parameters(), lr=4e-4, betas=(0.5, 0.999))
    net(torch.randn(3, 6)).sum().backward()
    ref_opt.step()                                              # real Adam state: step 1, non-zero moments
    fused = FusedAdam(net.parameters(), lr=1e-3, betas=(0.9, 0.999))
    fused.load_state_dict(ref_opt.state_dict())                 # torch -> fused
    assert fused.param_groups[0]["lr"] == 4e-4 and tuple(fused.param_groups[0]["betas"]) == (0.5, 0.999)
    sd = fused.state_dict()
    want = set(torch.optim.Adam([torch.zeros(1)]).param_groups[0]) - {"params"}
    assert want <= set(sd["param_groups"][0]), "missing torch.optim.Adam group keys: %s" % (want - set(sd["param_groups"][0]))
    for i, e in ref_opt.state_dict()["state"].items():
        assert torch.allclose(sd["state"][i]["exp_avg"], e["exp_avg"]) and torch.allclose(sd["state"][i]["exp_avg_sq"], e["exp_avg_sq"])
        assert sd["state"][i]["step"].device.type == "cpu" and float(sd["state"][i]["step"]) == float(e["step"]) == 1.0
    # fused -> torch: loads and STEPS (this raised KeyError: 'weight_decay' before the groups carried every key)
    net2 = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Linear(5, 4))
    net2.load_state_dict({k: v.clone() for k, v in net.state_dict().items()})
    opt2 = torch.optim.Adam(net2.parameters(), lr=1.0)
    opt2.load_state_dict(sd)
    x = torch.randn(3, 6)
    net2(x).sum().backward()
    before = [p.detach().clone() for p in net2.parameters()]
    opt2.step()
    assert all(not torch.equal(a, b) for a, b in zip(before, net2.parameters()))
    # and it moved exactly like the torch optimiser the state came from
    net.zero_grad()
    net(x).sum().backward()
    ref_opt.step()
    for a, b in zip(net.parameters(), net2.parameters()):
        assert torch.allclose(a, b, atol=1e-7)
