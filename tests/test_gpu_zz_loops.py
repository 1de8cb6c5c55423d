"""engine.SnganStep / engine.AcganStep on the GPU against the CPU oracle's restatements of main_sngan.py:65-100 and
main_acgan.py:84-133 (oracle.CpuSnganTrainer / CpuAcganTrainer, themselves pinned to traces of the unmodified
reference), plus the fused ACGAN objective kernel (gp_acgan_loss) and the packed two-head pass against torch formulas.

GAN training amplifies rounding, so every trace comparison prints an fp32 control beside it: the
oracle re-run with a 1e-6 perturbation of the first image batch."""
import contextlib
import io
import math
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = "cuda"


def quiet(fn):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn()


def rel(a, b):
    return (a - b).abs() / b.abs().clamp_min(1e-3)


# ------------------------------------------------------------------------------------------------ kernels
@pytest.mark.parametrize("is_real,is_gen", [(True, False), (False, False), (False, True)])
@pytest.mark.parametrize("mode", ["vanilla", "hinge", "lsgan"])
def test_acgan_loss_kernel_matches_torch(mode, is_real, is_gen):
    from gan_playground_b200.criterion import ACGANLoss, GANLoss
    from oracle import gan_oracle as O

    gen = torch.Generator().manual_seed(3)
    NB, K = 37, 10
    logits = (torch.randn(NB, K + 1, generator=gen) * 2).requires_grad_(True)
    labels = torch.randint(0, 2, (NB, K), generator=gen).float()
    l_adv = O.gan_loss(mode, logits[:, :1], is_real, is_gen, 0.9, 0.1, 0.9)
    l_aux = torch.nn.functional.mse_loss(logits[:, 1:], labels)
    (l_adv + 0.5 * l_aux).backward()
    crit = ACGANLoss(GANLoss(mode, 0.9, 0.1, 0.9), 0.5).to(DEV)
    lg = logits.detach().to(DEV).requires_grad_(True)
    out = crit(lg, labels.to(DEV), is_real, is_gen)
    out[crit.TOTAL].backward()
    want = torch.stack([l_adv, l_aux, l_adv + 0.5 * l_aux, torch.sigmoid(logits[:, 0]).mean()]).detach()
    assert torch.allclose(out.detach().cpu(), want, rtol=1e-5, atol=1e-6), (out, want)
    assert torch.allclose(lg.grad.cpu(), logits.grad, rtol=1e-5, atol=1e-7)


def test_packed_heads_equal_two_separate_heads():
    """One pass over the features with the stacked (1 + n_class, C) weight == the two Head nodes it replaces: logits,
    feature gradient and all four parameter gradients."""
    from gan_playground_b200 import functional as GF

    gen = torch.Generator().manual_seed(4)
    NB, H, W, C, K = 8, 4, 4, 64, 10
    a = torch.randn(NB, H, W, C, generator=gen).to(DEV).to(torch.bfloat16)
    ws = [torch.randn(o, C, generator=gen).to(DEV) * 0.1 for o in (1, K)]
    bs = [torch.randn(o, generator=gen).to(DEV) for o in (1, K)]
    g = torch.randn(NB, 1 + K, generator=gen).to(DEV)

    def leaves():
        return (a.clone().requires_grad_(True), [w.clone().requires_grad_(True) for w in ws],
                [b.clone().requires_grad_(True) for b in bs])

    a1, w1, b1 = leaves()
    packed = GF.PackedHeads.apply(a1, None, w1[0], b1[0], w1[1], b1[1], False)
    packed.backward(g)
    a2, w2, b2 = leaves()
    o0 = GF.Head.apply(a2, None, w2[0], b2[0], False)
    o1 = GF.Head.apply(a2, None, w2[1], b2[1], False)
    torch.cat([o0, o1], 1).backward(g)
    assert packed.shape == (NB, 1 + K)
    assert torch.allclose(packed, torch.cat([o0, o1], 1), rtol=1e-5, atol=1e-5)
    assert torch.allclose(a1.grad.float(), a2.grad.float(), rtol=2e-2, atol=2e-2)     # bf16 sum of two vs one rounding
    for x, y in zip(w1 + b1, w2 + b2):
        assert torch.allclose(x.grad, y.grad, rtol=1e-4, atol=1e-4)
    # frozen parameters (the G step): only the feature gradient is produced
    a3, w3, b3 = leaves()
    for t in w3 + b3:
        t.requires_grad_(False)
    GF.PackedHeads.apply(a3, None, w3[0], b3[0], w3[1], b3[1], False).backward(g)
    assert torch.allclose(a3.grad.float(), a1.grad.float())


# ------------------------------------------------------------------------------------------------ ACGAN loop
def _acgan(width, z_dim, use_graph, fused_adam=True):
    from gan_playground_b200.criterion import GANLoss
    from gan_playground_b200.engine import AcganStep
    from gan_playground_b200.models import acgan
    from gan_playground_b200.optim import FusedAdam

    torch.manual_seed(0)
    netG = quiet(lambda: acgan.Generator(z_dim=z_dim, ngf=width, n_class=10))
    netD = quiet(lambda: acgan.Discriminator(ndf=width, n_class=10))
    sd = ({k: v.clone() for k, v in netG.state_dict().items()}, {k: v.clone() for k, v in netD.state_dict().items()})
    netG.to(DEV), netD.to(DEV)
    if fused_adam:
        oG = FusedAdam(netG.parameters(), lr=4e-4, betas=(0.5, 0.999))
        oD = FusedAdam(netD.parameters(), lr=1e-4, betas=(0.5, 0.999))
    else:
        oG = torch.optim.Adam(netG.parameters(), lr=4e-4, betas=(0.5, 0.999), capturable=use_graph)
        oD = torch.optim.Adam(netD.parameters(), lr=1e-4, betas=(0.5, 0.999), capturable=use_graph)
    crit = GANLoss("vanilla", 0.9, 0.1, 0.9).to(DEV)
    return netG, netD, sd, lambda batch: AcganStep(netG, netD, crit, oG, oD, batch, z_dim, torch.device(DEV, 0),
                                                   use_graph=use_graph)


def test_acgan_step_tracks_the_oracle_loop():
    from oracle import gan_oracle as O

    steps, B, width, z_dim = 8, 16, 16, 32
    gen = torch.Generator().manual_seed(21)
    xs = torch.rand(steps, B, 3, 64, 64, generator=gen) * 2 - 1
    ys = torch.randint(0, 2, (steps, B, 10), generator=gen).float()
    zs = torch.randn(steps, B, z_dim, generator=gen)
    netG, netD, (sd_g, sd_d), make = _acgan(width, z_dim, use_graph=False, fused_adam=False)
    torch.set_num_threads(os.cpu_count())

    def oracle(perturb):
        tr = O.CpuAcganTrainer(sd_g, sd_d)
        return torch.tensor([tr.step(xs[i] + (perturb if i == 0 else 0.0), ys[i], zs[i]) for i in range(steps)]), tr

    ref, tr = oracle(0.0)
    ctl, tr_ctl = oracle(1e-6)
    run = make(B)
    got = torch.tensor([run.step(xs[i].to(DEV), ys[i].to(DEV), zs[i].to(DEV)) for i in range(steps)])
    assert torch.isfinite(got).all() and got.shape == (steps, 7)
    dev0, dev, c = rel(got[0], ref[0]).max().item(), rel(got, ref).max().item(), rel(ctl, ref).max().item()
    print("\nACGAN loop, %d steps: first step %.3f%%, all steps %.3f%% (fp32 1e-6-perturbation control %.3f%%)"
          % (steps, 100 * dev0, 100 * dev, 100 * c))
    assert dev0 < 0.005                   # teacher-forced: identical weights (measured 8e-5 in bf16x3)
    assert dev < max(0.03, 3 * c)
    # bookkeeping of the loop: D saw 3 forwards per iteration, G one; the auxiliary head moved like the oracle's
    assert int(netD.blocks[1][1].num_batches_tracked) == 3 * steps and int(netG.blocks[0][1].num_batches_tracked) == steps
    upd_ref = (tr.pd["out_aux.weight"].detach() - sd_d["out_aux.weight"]).flatten()
    upd = (netD.out_aux.weight.detach().cpu() - sd_d["out_aux.weight"]).flatten()
    upd_ctl = (tr_ctl.pd["out_aux.weight"].detach() - sd_d["out_aux.weight"]).flatten()
    cos = torch.nn.functional.cosine_similarity(upd, upd_ref, dim=0).item()
    cos_ctl = torch.nn.functional.cosine_similarity(upd_ctl, upd_ref, dim=0).item()
    print("auxiliary head: cosine of the %d-step weight update vs the oracle's %.4f (fp32 1e-6-perturbation control %.4f)"
          % (steps, cos, cos_ctl))
    # Adam turns every gradient into a +-lr step, so the accumulated update is chaotic even in fp32: the bar is the
    # fp32 control's own agreement with the oracle (minus 0.02), capped at the north_star cosine
    assert cos > min(0.999, cos_ctl) - 0.02


def test_acgan_graph_replay_is_the_eager_step():
    steps, B, width, z_dim = 3, 16, 16, 32
    gen = torch.Generator().manual_seed(22)
    xs = (torch.rand(steps, B, 3, 64, 64, generator=gen) * 2 - 1).to(DEV)
    ys = torch.randint(0, 2, (steps, B, 10), generator=gen).float().to(DEV)
    zs = torch.randn(steps, B, z_dim, generator=gen).to(DEV)
    runs = []
    for use_graph in (False, True):
        netG, netD, _, make = _acgan(width, z_dim, use_graph)
        run = make(B)
        runs.append(([run.step(xs[i], ys[i], zs[i]) for i in range(steps)], netG, netD))
    (le, gE, dE), (lg, gG, dG) = runs
    assert max(abs(a - b) for a, b in zip(le[0][:2], lg[0][:2])) < 1e-3, (le[0], lg[0])
    for a, b in zip(le, lg):
        assert all(math.isfinite(v) for v in b)
        assert max(abs(x - y) for x, y in zip(a, b)) < 3e-2, (a, b)
    # capture's warm-up steps were undone
    assert int(dE.blocks[1][1].num_batches_tracked) == int(dG.blocks[1][1].num_batches_tracked) == 3 * steps
    assert int(gE.blocks[0][1].num_batches_tracked) == int(gG.blocks[0][1].num_batches_tracked) == steps


# ------------------------------------------------------------------------------------------------ SNGAN loop
def _sngan(ch, z_dim, use_graph, n_disc_update):
    from gan_playground_b200.criterion import GANLoss
    from gan_playground_b200.engine import SnganStep
    from gan_playground_b200.models import sngan_projection as M
    from gan_playground_b200.optim import FusedAdam

    torch.manual_seed(0)
    netG = quiet(lambda: M.ResNetGenerator(ch=ch, dim_z=z_dim, bottom_width=2, img_dim=3, n_classes=10))
    netD = quiet(lambda: M.SNResNetProjectionDiscriminator(ch=ch, n_classes=10, img_dim=3))
    sd = ({k: v.clone() for k, v in netG.state_dict().items()}, {k: v.clone() for k, v in netD.state_dict().items()})
    netG.to(DEV), netD.to(DEV)
    netG.train(), netD.train()
    oG = FusedAdam(netG.parameters(), lr=2e-4, betas=(0.0, 0.999))
    oD = FusedAdam(netD.parameters(), lr=2e-4, betas=(0.0, 0.999))
    crit = GANLoss("hinge").to(DEV)
    return netG, netD, sd, lambda batch: SnganStep(netG, netD, crit, oG, oD, batch, z_dim, torch.device(DEV, 0),
                                                   n_classes=10, n_disc_update=n_disc_update, use_graph=use_graph)


def _sngan_data(steps, B, z_dim, seed):
    gen = torch.Generator().manual_seed(seed)
    return (torch.rand(steps, B, 3, 32, 32, generator=gen) * 2 - 1, torch.randint(10, (steps, B), generator=gen),
            torch.randn(steps, B, z_dim, generator=gen), torch.randint(10, (steps, B), generator=gen))


def test_sngan_step_tracks_the_oracle_loop():
    from oracle import gan_oracle as O

    steps, B, ch, z_dim, n = 6, 16, 16, 32, 2
    xs, ys, zs, cs = _sngan_data(steps, B, z_dim, 23)
    netG, netD, (sd_g, sd_d), make = _sngan(ch, z_dim, False, n)
    torch.set_num_threads(os.cpu_count())

    def oracle(perturb):
        tr = O.CpuSnganTrainer(sd_g, sd_d, n_disc_update=n, bottom_width=2)
        rows = [tr.step(xs[i] + (perturb if i == 0 else 0.0), ys[i], zs[i], cs[i]) for i in range(steps)]
        return torch.tensor([[float("nan") if v is None else v for v in r] for r in rows]), tr

    ref, tr = oracle(0.0)
    ctl, _ = oracle(1e-6)
    run = make(B)
    got = torch.tensor([run.step(xs[i].to(DEV), ys[i].to(DEV), zs[i].to(DEV), cs[i].to(DEV)) for i in range(steps)])
    ok = ~torch.isnan(ref)
    assert torch.equal(ok, ~torch.isnan(got)), got           # G-step numbers exist exactly on iterations 0, 2, 4
    assert ok[:, 2].tolist() == [i % n == 0 for i in range(steps)]
    # the two discriminator hinge losses (columns 0-1, ~1 while the logits are small) relative; the generator loss
    # -mean D(G(z)) and the D(.) means (columns 2-5) sit near zero, so those are measured against the largest of them
    scale = ref[:, 2:][ok[:, 2:]].abs().max().item()
    dl0 = rel(got[0, :2], ref[0, :2]).max().item()
    dl = rel(got[:, :2], ref[:, :2]).max().item()
    cl = rel(ctl[:, :2], ref[:, :2]).max().item()
    dm = ((got[:, 2:] - ref[:, 2:]).abs()[ok[:, 2:]].max() / scale).item()
    cm = ((ctl[:, 2:] - ref[:, 2:]).abs()[ok[:, 2:]].max() / scale).item()
    print("\nSNGAN loop, %d steps (bf16 operands): D losses first step %.3f%%, all steps %.3f%% (fp32 control %.3f%%); "
          "G loss and D(.) means %.3f%% of their scale %.4f (control %.3f%%)"
          % (steps, 100 * dl0, 100 * dl, 100 * cl, 100 * dm, scale, 100 * cm))
    assert dl0 < 0.02
    assert dl < max(0.02, 3 * cl)
    assert dm < max(0.05, 3 * cm)       # measured 3e-3
    # one generator forward per iteration, spectral-norm vectors advanced 2 or 3 times per iteration like the oracle's
    assert int(netG.b6.num_batches_tracked) == steps
    u_ref = tr.bd["l6.weight_u"]
    assert torch.nn.functional.cosine_similarity(netD.l6.weight_u.detach().cpu(), u_ref, dim=0).abs().item() > 0.999


def test_sngan_graph_variants_replay_the_eager_loop():
    """Graph mode captures two variants (with / without the G step) and replays the one iteration i needs."""
    steps, B, ch, z_dim, n = 5, 16, 16, 32, 2
    xs, ys, zs, cs = (t.to(DEV) for t in _sngan_data(steps, B, z_dim, 24))
    runs = []
    for use_graph in (False, True):
        netG, netD, _, make = _sngan(ch, z_dim, use_graph, n)
        run = make(B)
        runs.append(([run.step(xs[i], ys[i], zs[i], cs[i]) for i in range(steps)], netG, netD, run))
    (le, gE, dE, _), (lg, gG, dG, rg) = runs
    assert sorted(rg.graphs) == [0, 1]
    for i, (a, b) in enumerate(zip(le, lg)):
        for j, (x, y) in enumerate(zip(a, b)):
            if math.isnan(x):
                assert math.isnan(y) and i % n != 0 and j in (2, 5)
            else:
                assert abs(x - y) < 5e-2, (i, j, a, b)
    assert int(gE.b6.num_batches_tracked) == int(gG.b6.num_batches_tracked) == steps
    assert torch.allclose(dE.l6.weight_u, dG.l6.weight_u, atol=5e-2)


# ------------------------------------------------------------------------------------------------ checkpoints
@pytest.mark.parametrize("fused", [False, True])
def test_gpu_training_resumes_from_the_reference_checkpoint(fused):
    """Mirrors + optimiser (torch.optim.Adam as in the scripts, or optim.FusedAdam) restored from the reference's
    checkpoint run the reference's next iteration: same three losses (<= 2 %, the north_star bar)."""
    import os
    import sys

    from conftest import ROOT
    from gan_playground_b200.criterion import GANLoss
    from gan_playground_b200.engine import DcganStep
    from gan_playground_b200.optim import FusedAdam

    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from make_golden import trace_data

    from conftest import load_golden
    from test_checkpoint_compat import _pairs

    fx = load_golden("checkpoint_ref_layout.pt")
    netG, netD, ck = _pairs(fx)["dcgan"]
    netG.load_state_dict(ck["state_dict"]["generator"])
    netD.load_state_dict(ck["state_dict"]["discriminator"])
    netG.cuda(), netD.cuda()
    if fused:
        optG = FusedAdam(netG.parameters(), lr=4e-4, betas=(0.5, 0.999))
        optD = FusedAdam(netD.parameters(), lr=1e-4, betas=(0.5, 0.999))
    else:
        optG = torch.optim.Adam(netG.parameters(), lr=4e-4, betas=(0.5, 0.999))
        optD = torch.optim.Adam(netD.parameters(), lr=1e-4, betas=(0.5, 0.999))
    optG.load_state_dict(ck["optimizer"]["generator"])
    optD.load_state_dict(ck["optimizer"]["discriminator"])
    if fused:   # same layout out as in
        sd = optG.state_dict()
        ref = ck["optimizer"]["generator"]
        assert sorted(sd["state"]) == sorted(ref["state"])
        for i, e in ref["state"].items():
            assert torch.allclose(sd["state"][i]["exp_avg"].cpu(), e["exp_avg"]) and float(sd["state"][i]["step"]) == float(e["step"])
    xs, zs = trace_data(fx["dcgan"]["seed"], 3, 8, 32, 16)
    run = DcganStep(netG, netD, GANLoss("vanilla", 0.9, 0.1, 0.9).cuda(), optG, optD, 8, 16, torch.device("cuda", 0))
    got = run.step(xs[2].cuda(), zs[2].cuda())[:3]
    print("resumed iteration (fused=%s): ours %s vs reference %s" % (fused, got, fx["dcgan"]["next_losses"]))
    assert got == pytest.approx(fx["dcgan"]["next_losses"], rel=2e-2, abs=2e-3)
