"""Host logic of engine.DcganStep / SnganStep / AcganStep on the CPU: the ORDER of forward / backward / zero_grad / step
calls, the detach and graph re-use, the G step every n_disc_update iterations, the frozen-D G step and the positions of
the logged scalars must be the reference scripts' (main_dcgan.py:68-95, main_sngan.py:65-100, main_acgan.py:84-133).

The networks here are CPU stand-ins — nn.Modules whose parameters are the entries of a reference-layout state_dict and
whose forward is the oracle's functional restatement — so the drivers' eager loop bodies run without a GPU and must
reproduce the traces the UNMODIFIED reference produced (tests/golden/*_loop_*.pt, dcgan_trace_r32_w4.pt). The CUDA
networks / losses the drivers normally carry are compared with the same oracle in the `-m gpu` tests."""
import os
import sys

import pytest
import torch
import torch.nn as nn

from conftest import ROOT, load_golden
from oracle import gan_oracle as O

sys.path.insert(0, os.path.join(ROOT, "oracle"))


class OracleNet(nn.Module):
    """A reference-layout state_dict as a module: float entries become parameters, BN / SN buffers stay buffers (updated in
    place by the oracle's functional nets), forward = `fn(state, *inputs)`."""

    def __init__(self, sd, fn, **attrs):
        super().__init__()
        params, buffers = O.split_state({k: v.clone() for k, v in sd.items()})
        self._names = {}
        for k, v in params.items():
            self._names[k] = ("p", k.replace(".", "__"))
            self.register_parameter(k.replace(".", "__"), nn.Parameter(v))
        self._bufs = buffers
        self._fn = fn
        for k, v in attrs.items():
            setattr(self, k, v)

    def state(self):
        sd = {k: getattr(self, n) for k, (_, n) in self._names.items()}
        sd.update(self._bufs)
        return sd

    def forward(self, *inputs):
        return self._fn(self.state(), self._bufs, *inputs)

    # engine._AdversarialStep snapshots / restores state_dict() only in graph mode; eager mode never calls it


def _drive(runner_cls, *args, **kw):
    from gan_playground_b200 import engine

    return getattr(engine, runner_cls)(*args, device=torch.device("cpu"), use_graph=False, **kw)


def test_dcgan_step_loop_is_main_dcgan(monkeypatch):
    from make_golden import trace_data

    fx = load_golden("dcgan_trace_r32_w4.pt")
    xs, zs = trace_data(fx["seed"], fx["steps"], fx["batch"], fx["res"], fx["z_dim"])
    netG = OracleNet(fx["sd_g"], lambda sd, b, z: O.dcgan_generator(sd, z, buffers=b))
    netD = OracleNet(fx["sd_d"], lambda sd, b, x: O.dcgan_discriminator(sd, x, buffers=b), img_dim=3, resolution=fx["res"])
    optG = torch.optim.Adam(netG.parameters(), lr=4e-4, betas=(0.5, 0.999))
    optD = torch.optim.Adam(netD.parameters(), lr=1e-4, betas=(0.5, 0.999))

    def crit(pred, is_real, is_generator=False):
        return O.gan_loss("vanilla", pred, is_real, is_generator, 0.9, 0.1, 0.9)

    run = _drive("DcganStep", netG, netD, crit, optG, optD, fx["batch"], fx["z_dim"])
    ref = O.CpuDcganTrainer(fx["sd_g"], fx["sd_d"])
    for i in range(6):
        got = run.step(xs[i], zs[i])
        want = ref.step(xs[i], zs[i, 0], zs[i, 1])
        assert got == pytest.approx(list(want), abs=2e-5), i          # same arithmetic in the same order
        assert got[:3] == pytest.approx(fx["trace"][i].tolist(), rel=2e-2, abs=2e-3)
    assert int(netD._bufs["blocks.1.1.num_batches_tracked"]) == 18 and int(netG._bufs["blocks.0.1.num_batches_tracked"]) == 12
    assert all(p.requires_grad for p in netD.parameters())           # the frozen-D G step un-freezes on exit
    assert run.iteration == 6


def test_sngan_step_loop_is_main_sngan():
    from make_golden import sngan_loop_data

    fx = load_golden("sngan_loop_ch8.pt")
    xs, ys, zs, cs = sngan_loop_data(fx["seed"], fx["steps"], fx["batch"], fx["z_dim"])
    netG = OracleNet(fx["sd_g"], lambda sd, b, z, c: O.sngan_generator(sd, z, c, bottom_width=2, buffers=b))
    netD = OracleNet(fx["sd_d"], lambda sd, b, x, y=None: O.sngan_discriminator(sd, x, y))
    netD.block1 = nn.Module()
    netD.block1.c1 = nn.Module()
    netD.block1.c1.in_channels = 3
    optG = torch.optim.Adam(netG.parameters(), lr=2e-4, betas=(0.0, 0.999))
    optD = torch.optim.Adam(netD.parameters(), lr=2e-4, betas=(0.0, 0.999))

    def crit(pred, is_real, is_generator=False):
        return O.gan_loss("hinge", pred, is_real, is_generator)

    n = fx["n_disc_update"]
    run = _drive("SnganStep", netG, netD, crit, optG, optD, fx["batch"], fx["z_dim"], n_disc_update=n)
    ref = O.CpuSnganTrainer(fx["sd_g"], fx["sd_d"], n_disc_update=n, bottom_width=2)
    for i in range(5):
        got = run.step(xs[i], ys[i], zs[i], cs[i])
        want = ref.step(xs[i], ys[i], zs[i], cs[i])
        for j, (g, w) in enumerate(zip(got, want)):
            if w is None:
                assert g != g and i % n != 0 and j in (2, 5)          # NaN: no G step on this iteration
            else:
                assert g == pytest.approx(w, abs=2e-5), (i, j)
        row = fx["trace"][i]
        ok = ~torch.isnan(row)
        assert torch.tensor(got)[ok].tolist() == pytest.approx(row[ok].tolist(), rel=2e-2, abs=2e-3)
    # the generator ran once per iteration (its BatchNorm counters say so), D three or two times
    assert int(netG._bufs["b6.num_batches_tracked"]) == 5


def test_acgan_step_loop_is_main_acgan():
    from gan_playground_b200.criterion import ACGANLoss, GANLoss
    from make_golden import acgan_loop_data

    fx = load_golden("acgan_loop_r64_w4.pt")
    xs, ys, zs = acgan_loop_data(fx["seed"], fx["steps"], fx["batch"], fx["z_dim"])

    class CpuACGANLoss(ACGANLoss):
        """gp_acgan_loss's contract in torch ops: [adv, aux, adv + w * aux, mean sigmoid(adv)] of the packed logits."""

        def forward(self, packed_logits, labels, is_real, is_generator=False):
            adv, cls = packed_logits[:, :1], packed_logits[:, 1:]
            l_adv = O.gan_loss("vanilla", adv, is_real, is_generator, 0.9, 0.1, 0.9)
            l_aux = torch.nn.functional.mse_loss(cls, labels)
            return torch.stack([l_adv, l_aux, l_adv + self.aux_weight * l_aux, torch.sigmoid(adv).mean()])

    netG = OracleNet(fx["sd_g"], lambda sd, b, z, y: O.dcgan_generator(sd, z, y, acgan=True, buffers=b))
    netD = OracleNet(fx["sd_d"], None, img_dim=3, resolution=64)
    netD.packed_logits = lambda x: torch.cat(O.dcgan_discriminator(netD.state(), x, acgan=True, buffers=netD._bufs), 1)
    optG = torch.optim.Adam(netG.parameters(), lr=4e-4, betas=(0.5, 0.999))
    optD = torch.optim.Adam(netD.parameters(), lr=1e-4, betas=(0.5, 0.999))
    crit = CpuACGANLoss(GANLoss("vanilla", 0.9, 0.1, 0.9))
    run = _drive("AcganStep", netG, netD, crit, optG, optD, fx["batch"], fx["z_dim"])
    ref = O.CpuAcganTrainer(fx["sd_g"], fx["sd_d"])
    for i in range(5):
        got = run.step(xs[i], ys[i], zs[i])
        want = ref.step(xs[i], ys[i], zs[i])
        assert len(got) == 7
        assert got == pytest.approx(list(want), abs=2e-5), i
        assert got == pytest.approx(fx["trace"][i].tolist(), rel=2e-2, abs=2e-3)
    assert int(netD._bufs["blocks.1.1.num_batches_tracked"]) == 15 and int(netG._bufs["blocks.0.1.num_batches_tracked"]) == 5


def test_graph_mode_rejects_mixed_noise_sources():
    """Host-side guard of graph mode (no GPU needed to reach it): a driver captured with caller-supplied noise cannot
    silently switch to on-device noise, the captured graph reads the static noise buffers."""
    from gan_playground_b200 import engine

    run = engine.DcganStep.__new__(engine.DcganStep)
    run.use_graph, run.graphs, run.fixed_noise, run.iteration = True, {0: object()}, True, 3
    with pytest.raises(ValueError):
        run._run([torch.zeros(1)], None)


def test_weight_cache_only_caches_real_parameters():
    """A derived weight (W / sigma) is re-staged every forward even when it looks like an unchanged leaf: under frozen
    parameters / no_grad it is a version-0 leaf whose address the allocator may hand out again."""
    from gan_playground_b200.functional import WeightCache

    cache, calls = WeightCache(), []
    p = nn.Parameter(torch.zeros(4))
    make = lambda: calls.append(1) or len(calls)
    assert cache.get("k", p, make) == 1 and cache.get("k", p, make) == 1          # cached
    with torch.no_grad():
        p.add_(1.0)                                                               # optimiser step: version bump
    assert cache.get("k", p, make) == 2
    p._gp_epoch = 1                                                               # FusedAdam's out-of-band update
    assert cache.get("k", p, make) == 3
    p.requires_grad_(False)                                                       # frozen (G step): still the same parameter
    assert cache.get("k", p, make) == 3
    derived = torch.zeros(4)                                                      # leaf, version 0, not a Parameter
    assert cache.get("d", derived, make) == 4 and cache.get("d", derived, make) == 5
