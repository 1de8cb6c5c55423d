"""Pins the CPU oracle (oracle/gan_oracle.py) to the reference: against the committed golden vectors produced by the
unmodified reference (oracle/make_golden.py), and against the live reference when /root/reference is present
(build container only). fp32 CPU vs fp32 CPU: tolerances are rounding-level (forward 1e-5; gradients 2e-3 relative
because the fixtures store them as fp16-normalised values)."""
import os
import sys

import pytest
import torch

from conftest import REFERENCE, ROOT, load_golden, unpack_grads
from oracle import gan_oracle as O


def close(a, b, tol):
    return ((a.float() - b.float()).abs().max() <= tol * (b.float().abs().max() + 1e-12)).item()


def check_grads(got, ref, tol=2e-3, floor=1e-5):
    for k, r in ref.items():
        if r.abs().max() < floor:        # analytically-zero gradients (pre-BN biases): rounding noise only
            continue
        assert k in got, k
        assert close(got[k], r, tol), (k, (got[k] - r).abs().max().item(), r.abs().max().item())


@pytest.mark.parametrize("name,kw", [
    ("dcgan_r32_w4.pt", {}), ("dcgan_r64_w4.pt", {}), ("snd_r32_w4.pt", {"sn": True, "flatten_head": True}),
    ("dcgan_blur_r32_w8.pt", {"blur": True}), ("snd_up_r32_w4.pt", {"up": True}),
])
def test_dcgan_family_step_matches_golden(name, kw):
    fx = load_golden(name)
    sd_g = {k: v.clone() for k, v in fx["sd_g"].items()}
    sd_d = {k: v.clone() for k, v in fx["sd_d"].items()}
    r = O.dcgan_step_grads(sd_g, sd_d, fx["x"], fx["z1"], fx["z2"], labels=fx["labels"], mode=fx["mode"], **kw)
    for key in ("fake1", "fake2", "d_real", "d_fake", "d_g", "loss_real", "loss_fake", "loss_g"):
        assert close(r[key], fx[key], 1e-5), key
    check_grads(r["d_grads_real"], unpack_grads(fx["d_grads_real"]))
    check_grads(r["d_grads_fake"], unpack_grads(fx["d_grads_fake"]))
    check_grads(r["g_grads"], unpack_grads(fx["g_grads"]))


def test_sngan_projection_forward_and_sn_buffers_match_golden():
    fx = load_golden("sngan_proj_ch8.pt")
    sd_g = {k: v.clone() for k, v in fx["sd_g"].items()}
    sd_d = {k: v.clone() for k, v in fx["sd_d"].items()}
    with torch.no_grad():
        d_real = O.sngan_discriminator(sd_d, fx["x"], fx["y"])
        fake = O.sngan_generator(sd_g, fx["z"], fx["c"], bottom_width=2)
        d_fake = O.sngan_discriminator(sd_d, fake, fx["c"])
        d_g = O.sngan_discriminator(sd_d, fake, fx["c"])
    assert close(d_real, fx["d_real"], 1e-5) and close(fake, fx["fake"], 1e-5)
    assert close(d_fake, fx["d_fake"], 1e-5) and close(d_g, fx["d_g"], 1e-5)
    # u / v after the three train-mode forwards of one main_sngan.py iteration (one power iteration each)
    for k, v in fx["buf_d_after"].items():
        if k.endswith(("weight_u", "weight_v")):
            assert close(sd_d[k], v, 1e-5), k


def test_sngan_projection_gradients_match_golden():
    fx = load_golden("sngan_proj_ch8.pt")
    sd_d = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and not k.endswith(("_u", "_v")) else v.clone())
            for k, v in fx["sd_d"].items()}
    out = O.sngan_discriminator(sd_d, fx["x"], fx["y"])
    loss = O.gan_loss("hinge", out, True)
    assert close(loss, fx["loss_real"], 1e-5)
    leaves = [k for k, v in sd_d.items() if v.requires_grad]
    grads = torch.autograd.grad(loss, [sd_d[k] for k in leaves], allow_unused=True)
    check_grads({k: g for k, g in zip(leaves, grads) if g is not None}, unpack_grads(fx["d_grads_real"]))


def test_acgan_forward_matches_golden():
    fx = load_golden("acgan_r64_w4.pt")
    with torch.no_grad():
        fake = O.dcgan_generator({k: v.clone() for k, v in fx["sd_g"].items()}, fx["z"], fx["y"], acgan=True)
        adv, cls = O.dcgan_discriminator({k: v.clone() for k, v in fx["sd_d"].items()}, fx["x"], acgan=True)
    assert close(fake, fx["fake"], 1e-5) and close(adv, fx["d_real"], 1e-5) and close(cls, fx["d_real_cls"], 1e-5)


def test_ganloss_values_and_gradients_match_golden():
    fx = load_golden("ganloss.pt")
    for row in fx["rows"]:
        p = fx["pred"].clone().requires_grad_(True)
        l = O.gan_loss(row["mode"], p, row["is_real"], row["is_generator"], *row["labels"])
        l.backward()
        assert abs(l.item() - row["loss"].item()) < 1e-6
        assert torch.allclose(p.grad, row["dpred"], atol=1e-7)
    with pytest.raises(NotImplementedError):
        O.gan_loss("wgan", fx["pred"], True)


def test_cpu_trainer_reproduces_reference_loss_trace():
    """20 iterations of main_dcgan.py:68-95 with Adam: the oracle trainer tracks the reference's trace. Adam turns the
    rounding-noise gradients of the pre-BN biases into +-lr steps, so agreement is 5e-3, not rounding-level."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from make_golden import trace_data

    fx = load_golden("dcgan_trace_r32_w4.pt")
    xs, zs = trace_data(fx["seed"], fx["steps"], fx["batch"], fx["res"], fx["z_dim"])
    assert torch.equal(xs[0, 0, 0, 0, :4], fx["x0_probe"])
    tr = O.CpuDcganTrainer(fx["sd_g"], fx["sd_d"])
    trace = torch.tensor([tr.step(xs[i], zs[i, 0], zs[i, 1])[:3] for i in range(fx["steps"])])
    assert (trace - fx["trace"]).abs().max() < 5e-3
    assert int(tr.bd["blocks.1.1.num_batches_tracked"]) == int(fx["buf_d_after"]["blocks.1.1.num_batches_tracked"]) == 60
    assert int(tr.bg["blocks.0.1.num_batches_tracked"]) == 40


def test_cpu_sngan_trainer_reproduces_reference_loop():
    """12 iterations of main_sngan.py:65-100 (n_disc_update=2: the G step runs on even iterations only and re-uses the
    fake batch's graph) against the unmodified reference's trace. The first iterations agree to rounding; later ones
    drift as Adam amplifies rounding-level gradient differences (bar: the north_star's 2 %)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from make_golden import sngan_loop_data

    fx = load_golden("sngan_loop_ch8.pt")
    xs, ys, zs, cs = sngan_loop_data(fx["seed"], fx["steps"], fx["batch"], fx["z_dim"])
    assert torch.equal(xs[0, 0, 0, 0, :4], fx["x0_probe"])
    tr = O.CpuSnganTrainer(fx["sd_g"], fx["sd_d"], n_disc_update=fx["n_disc_update"], bottom_width=2)
    rows = [tr.step(xs[i], ys[i], zs[i], cs[i]) for i in range(fx["steps"])]
    for i, row in enumerate(rows):
        assert (row[2] is None) == (i % fx["n_disc_update"] != 0) == bool(torch.isnan(fx["trace"][i, 2]))
    got = torch.tensor([[float("nan") if v is None else v for v in row] for row in rows])
    ok = ~torch.isnan(fx["trace"])
    assert torch.equal(ok, ~torch.isnan(got))
    assert (got[:2] - fx["trace"][:2])[ok[:2]].abs().max() < 1e-4
    assert ((got - fx["trace"])[ok].abs() <= 0.02 * fx["trace"][ok].abs() + 2e-3).all()
    sd_g, sd_d = tr.state_dicts()
    # bookkeeping: D ran 2 forwards per iteration + 1 per G step, G one forward per iteration
    assert int(sd_g["b6.num_batches_tracked"]) == int(fx["buf_g_after"]["b6.num_batches_tracked"]) == fx["steps"]
    assert torch.allclose(sd_d["l6.weight_u"], fx["buf_d_after"]["l6.weight_u"], atol=2e-3)
    assert torch.allclose(sd_g["l1.weight"][:4], fx["g_l1_w_after"], atol=5e-4)


def test_cpu_acgan_trainer_reproduces_reference_loop():
    """12 iterations of main_acgan.py:84-133 (adversarial BCE + 0.5 x MSE on the auxiliary head, fake batch conditioned
    on the real labels, one generator forward per iteration) against the unmodified reference's seven logged numbers."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from make_golden import acgan_loop_data

    fx = load_golden("acgan_loop_r64_w4.pt")
    xs, ys, zs = acgan_loop_data(fx["seed"], fx["steps"], fx["batch"], fx["z_dim"])
    assert torch.equal(xs[0, 0, 0, 0, :4], fx["x0_probe"])
    tr = O.CpuAcganTrainer(fx["sd_g"], fx["sd_d"])
    got = torch.tensor([tr.step(xs[i], ys[i], zs[i]) for i in range(fx["steps"])])
    assert got.shape == fx["trace"].shape == (fx["steps"], 7)
    assert (got[:2] - fx["trace"][:2]).abs().max() < 1e-4
    assert ((got - fx["trace"]).abs() <= 0.02 * fx["trace"].abs() + 2e-3).all()
    sd_g, sd_d = tr.state_dicts()
    assert int(sd_d["blocks.1.1.num_batches_tracked"]) == int(fx["buf_d_after"]["blocks.1.1.num_batches_tracked"]) == 3 * fx["steps"]
    assert int(sd_g["blocks.0.1.num_batches_tracked"]) == int(fx["buf_g_after"]["blocks.0.1.num_batches_tracked"]) == fx["steps"]
    assert torch.allclose(sd_d["out_aux.weight"], fx["d_aux_w_after"], atol=5e-4)


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="live reference only exists in the build container")
def test_oracle_matches_live_reference_full_width():
    import importlib.util
    import contextlib
    import io

    spec = importlib.util.spec_from_file_location("ref_dcgan_live", os.path.join(REFERENCE, "models", "dcgan.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    torch.manual_seed(3)
    with contextlib.redirect_stdout(io.StringIO()):
        netG, netD = ref.Generator(ngf=32), ref.Discriminator(ndf=32)
    z = torch.randn(4, 100)
    with torch.no_grad():
        sd_g = {k: v.clone() for k, v in netG.state_dict().items()}
        sd_d = {k: v.clone() for k, v in netD.state_dict().items()}
        fake_ref = netG(z)
        out_ref = netD(fake_ref)
        fake = O.dcgan_generator(sd_g, z)
        out = O.dcgan_discriminator(sd_d, fake)
    assert close(fake, fake_ref, 1e-5) and close(out, out_ref, 1e-5)
