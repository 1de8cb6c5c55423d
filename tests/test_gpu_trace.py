"""200-step G/D loss traces (main_dcgan.py:68-95 with Adam) against the fp32 CPU oracle on the same data stream.

GAN training is chaotic: SURVEY.md §7.3 measured that the fp32 reference itself, re-run with a 1e-6 perturbation of
the first input batch, deviates point-wise by up to 7 % within 200 steps. So the comparison protocol is the one
SURVEY prescribes: (a) the 200-step mean and the 50-step moving average of each loss must agree within 2 %
(north_star's bar), and (b) the point-wise deviation over the first 20 steps must not exceed the fp32-perturbation
control by more than a small factor; the control is computed and printed next to every number."""
import os

import pytest
import torch

from test_gpu_dcgan import quiet

pytestmark = pytest.mark.gpu

STEPS, B, RES, W = 200, 32, 32, 16


def moving_avg(t, k=50):
    c = torch.cumsum(torch.cat([torch.zeros(1, t.shape[1]), t]), 0)
    return (c[k:] - c[:-k]) / k


def test_loss_trace_200_steps_within_2_percent():
    from gan_playground_b200.criterion import GANLoss
    from gan_playground_b200.models import dcgan
    from oracle import gan_oracle as O

    torch.manual_seed(0)
    netG = quiet(lambda: dcgan.Generator(ngf=W, resolution=RES))
    netD = quiet(lambda: dcgan.Discriminator(ndf=W, resolution=RES))
    sd_g = {k: v.clone() for k, v in netG.state_dict().items()}
    sd_d = {k: v.clone() for k, v in netD.state_dict().items()}
    gen = torch.Generator().manual_seed(11)
    xs = torch.rand(STEPS, B, 3, RES, RES, generator=gen) * 2 - 1
    zs = torch.randn(STEPS, 2, B, 100, generator=gen)
    torch.set_num_threads(os.cpu_count())

    def run_oracle(perturb):
        tr = O.CpuDcganTrainer(sd_g, sd_d)
        out = []
        for i in range(STEPS):
            x = xs[i] + (perturb if i == 0 else 0.0)
            out.append(tr.step(x, zs[i, 0], zs[i, 1])[:3])
        return torch.tensor(out)

    ref = run_oracle(0.0)
    ctl = run_oracle(1e-6)                       # fp32 reference vs itself under a 1e-6 input perturbation

    netG.cuda(), netD.cuda()
    crit = GANLoss("vanilla", 0.9, 0.1, 0.9).cuda()
    optG = torch.optim.Adam(netG.parameters(), lr=4e-4, betas=(0.5, 0.999))
    optD = torch.optim.Adam(netD.parameters(), lr=1e-4, betas=(0.5, 0.999))
    got = []
    for i in range(STEPS):
        x, z1, z2 = xs[i].cuda(), zs[i, 0].cuda(), zs[i, 1].cuda()
        optD.zero_grad()
        l1 = crit(netD(x), True)
        l1.backward()
        l2 = crit(netD(netG(z1).detach()), False)
        l2.backward()
        optD.step()
        optG.zero_grad()
        l3 = crit(netD(netG(z2)), False, True)
        l3.backward()
        optG.step()
        got.append([l1.item(), l2.item(), l3.item()])
    got = torch.tensor(got)
    assert torch.isfinite(got).all()

    def rel(a, b):
        return ((a - b).abs() / b.abs().clamp_min(1e-6))

    mean_dev = rel(got.mean(0), ref.mean(0)).max().item()
    mean_ctl = rel(ctl.mean(0), ref.mean(0)).max().item()
    ma_dev = rel(moving_avg(got), moving_avg(ref)).max().item()
    ma_ctl = rel(moving_avg(ctl), moving_avg(ref)).max().item()
    pt_dev = rel(got[:20], ref[:20]).max().item()
    pt_ctl = rel(ctl[:20], ref[:20]).max().item()
    pt200_dev = rel(got, ref).max().item()
    pt200_ctl = rel(ctl, ref).max().item()
    print("\n200-step mean: ours %.3f%%  (fp32 1e-6-perturbation control %.3f%%)" % (100 * mean_dev, 100 * mean_ctl))
    print("50-step moving average max: ours %.3f%%  (control %.3f%%)" % (100 * ma_dev, 100 * ma_ctl))
    print("point-wise, first 20 steps: ours %.3f%%  (control %.3f%%)" % (100 * pt_dev, 100 * pt_ctl))
    print("point-wise, all 200 steps : ours %.3f%%  (control %.3f%%)" % (100 * pt200_dev, 100 * pt200_ctl))
    assert mean_dev < 0.02
    assert ma_dev < max(0.02, 2 * ma_ctl)
    assert pt_dev < max(0.02, 3 * pt_ctl)


def test_engine_mixed_policy_loss_trace_200_steps():
    """The same 200-step protocol through engine.DcganStep with its per-pass precision policy (real pass bf16, D-fake chain
    single-MMA fp16, G step bf16x3), eager and as CUDA-graph replays. tools/precision_study.py --trace emulates this on the
    CPU at 0.29 % (200-step mean) / 1.6 % (50-step moving average) against an fp32 control of 0.09 % / 2.1 %."""
    from gan_playground_b200.criterion import GANLoss
    from gan_playground_b200.engine import DcganStep
    from gan_playground_b200.models import dcgan
    from gan_playground_b200.optim import FusedAdam
    from oracle import gan_oracle as O

    torch.manual_seed(0)
    netG0 = quiet(lambda: dcgan.Generator(ngf=W, resolution=RES))
    netD0 = quiet(lambda: dcgan.Discriminator(ndf=W, resolution=RES))
    sd_g = {k: v.clone() for k, v in netG0.state_dict().items()}
    sd_d = {k: v.clone() for k, v in netD0.state_dict().items()}
    gen = torch.Generator().manual_seed(11)
    xs = torch.rand(STEPS, B, 3, RES, RES, generator=gen) * 2 - 1
    zs = torch.randn(STEPS, 2, B, 100, generator=gen)
    torch.set_num_threads(os.cpu_count())

    def run_oracle(perturb):
        tr = O.CpuDcganTrainer(sd_g, sd_d)
        return torch.tensor([tr.step(xs[i] + (perturb if i == 0 else 0.0), zs[i, 0], zs[i, 1])[:3] for i in range(STEPS)])

    ref, ctl = run_oracle(0.0), run_oracle(1e-6)

    def rel(a, b):
        return ((a - b).abs() / b.abs().clamp_min(1e-6)).max().item()

    for use_graph in (False, True):
        netG = quiet(lambda: dcgan.Generator(ngf=W, resolution=RES))
        netD = quiet(lambda: dcgan.Discriminator(ndf=W, resolution=RES))
        netG.load_state_dict(sd_g), netD.load_state_dict(sd_d)
        netG.cuda(), netD.cuda()
        oG = FusedAdam(netG.parameters(), lr=4e-4, betas=(0.5, 0.999))
        oD = FusedAdam(netD.parameters(), lr=1e-4, betas=(0.5, 0.999))
        run = DcganStep(netG, netD, GANLoss("vanilla", 0.9, 0.1, 0.9).cuda(), oG, oD, B, 100, torch.device("cuda", 0),
                        use_graph=use_graph)
        assert run.real_precision == "bf16" and run.fake_precision == "fp16"
        xd, zd = xs.cuda(), zs.cuda()
        got = torch.tensor([run.step(xd[i], zd[i])[:3] for i in range(STEPS)])
        assert torch.isfinite(got).all()
        mean_dev, ma_dev = rel(got.mean(0), ref.mean(0)), rel(moving_avg(got), moving_avg(ref))
        ma_ctl, pt_dev, pt_ctl = rel(moving_avg(ctl), moving_avg(ref)), rel(got[:20], ref[:20]), rel(ctl[:20], ref[:20])
        print("\nengine mixed policy (%s): 200-step mean %.3f%%, 50-step moving average %.3f%% (control %.3f%%), first 20 "
              "steps %.3f%% (control %.3f%%)" % ("graph" if use_graph else "eager", 100 * mean_dev, 100 * ma_dev,
                                                 100 * ma_ctl, 100 * pt_dev, 100 * pt_ctl))
        assert mean_dev < 0.02
        assert ma_dev < max(0.02, 2 * ma_ctl)
        assert pt_dev < max(0.03, 3 * pt_ctl)
