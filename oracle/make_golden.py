"""Generate tests/golden/*.pt from the UNMODIFIED reference (imported from /root/reference) — run in the build
container only (the GPU box has no /root/reference). Also cross-checks oracle/gan_oracle.py against the live
reference on the same inputs (max abs difference printed; must be ~1e-6 or below).

    python oracle/make_golden.py            # writes tests/golden/
    python oracle/make_golden.py dcgan_blur_r32_w8.pt   # only the named fixtures (+ the key lists)
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("GP_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from oracle import gan_oracle as O  # noqa: E402


def ref_modules():
    """Import the reference's model modules without polluting sys.path for the product's own `models` shim."""
    import importlib.util

    mods = {}
    for name in ("dcgan", "dcgan_specnorm", "dcgan_specnorm_up", "sngan_projection", "acgan"):
        spec = importlib.util.spec_from_file_location("ref_models_" + name, os.path.join(REF, "models", name + ".py"))
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        mods[name] = m
    # models/dcgan_blur.py does `from models.ops import BlurPool2d`: give it the REFERENCE's models/ops.py (not this
    # repository's `models` shim) while it is being imported
    import types

    spec = importlib.util.spec_from_file_location("ref_models_ops", os.path.join(REF, "models", "ops.py"))
    ref_ops = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_ops)
    saved = {k: sys.modules.get(k) for k in ("models", "models.ops")}
    pkg = types.ModuleType("models")
    pkg.ops, pkg.__path__ = ref_ops, []
    sys.modules["models"], sys.modules["models.ops"] = pkg, ref_ops
    try:
        spec = importlib.util.spec_from_file_location("ref_models_dcgan_blur", os.path.join(REF, "models", "dcgan_blur.py"))
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        mods["dcgan_blur"] = m
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    spec = importlib.util.spec_from_file_location("ref_criterion", os.path.join(REF, "utils", "criterion.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    mods["criterion"] = m
    return mods


def clone_sd(net):
    return {k: v.detach().clone() for k, v in net.state_dict().items()}


def pack_grads(d):
    """Store gradients compactly: per-tensor fp32 scale + fp16 normalised values (relative error 5e-4 per element,
    far inside the cosine >= 0.999 / 1e-3 bars the tests apply). Forward outputs, losses and weights stay fp32."""
    out = {}
    for k, g in d.items():
        s = g.abs().max().clamp_min(1e-30)
        out[k] = {"scale": s.clone(), "q": (g / s).to(torch.float16)}
    return out


def unpack_grads(d):
    return {k: v["q"].float() * v["scale"] for k, v in d.items()}


def grads_of(net):
    return {k: p.grad.detach().clone() for k, p in net.named_parameters() if p.grad is not None}


def buffers_of(net):
    return {k: v.detach().clone() for k, v in net.state_dict().items()
            if k.endswith(("running_mean", "running_var", "num_batches_tracked", "weight_u", "weight_v"))}


def maxdiff(a, b):
    return max((a[k].float() - b[k].float()).abs().max().item() for k in a if k in b)


def dcgan_like_fixture(mods, modname, res, width, batch, mode, labels, seed, z_dim=100):
    """One step of the main_dcgan.py loop body (:68-95) on the reference, recording everything."""
    M = mods[modname]
    torch.manual_seed(seed)
    netG = M.Generator(z_dim=z_dim, ngf=width, resolution=res)
    netD = M.Discriminator(ndf=width, resolution=res)
    crit = mods["criterion"].GANLoss(mode, *labels)
    netG.train(), netD.train()
    sd_g0, sd_d0 = clone_sd(netG), clone_sd(netD)
    gen = torch.Generator().manual_seed(seed + 1)
    x = torch.rand(batch, 3, res, res, generator=gen) * 2 - 1
    z1 = torch.randn(batch, z_dim, generator=gen)
    z2 = torch.randn(batch, z_dim, generator=gen)
    fx = {"sd_g": sd_g0, "sd_d": sd_d0, "x": x, "z1": z1, "z2": z2, "mode": mode, "labels": labels,
          "res": res, "width": width, "model": modname, "z_dim": z_dim}
    # D real
    netD.zero_grad()
    out = netD(x)
    loss_real = crit(out, True)
    loss_real.backward()
    fx["d_real"], fx["loss_real"], fx["d_grads_real"] = out.detach().clone(), loss_real.detach().clone(), grads_of(netD)
    # D fake
    fake1 = netG(z1)
    netD.zero_grad()
    out = netD(fake1.detach())
    loss_fake = crit(out, False)
    loss_fake.backward()
    fx["fake1"], fx["d_fake"], fx["loss_fake"] = fake1.detach().clone(), out.detach().clone(), loss_fake.detach().clone()
    fx["d_grads_fake"] = grads_of(netD)
    # G step
    netG.zero_grad(), netD.zero_grad()
    fake2 = netG(z2)
    out = netD(fake2)
    loss_g = crit(out, False, True)
    loss_g.backward()
    fx["fake2"], fx["d_g"], fx["loss_g"] = fake2.detach().clone(), out.detach().clone(), loss_g.detach().clone()
    fx["g_grads"] = grads_of(netG)
    fx["d_grads_gstep"] = grads_of(netD)
    # buffers after the 3 D forwards / 2 G forwards (running stats, num_batches_tracked, SN u/v)
    fx["buf_g_after"], fx["buf_d_after"] = buffers_of(netG), buffers_of(netD)

    # ---- cross-check the oracle restatement on the same inputs
    kw = dict(sn=(modname == "dcgan_specnorm"), flatten_head=(modname == "dcgan_specnorm"))
    if modname == "dcgan_blur":
        kw = dict(blur=True)
    if modname == "dcgan_specnorm_up":
        kw = dict(up=True)
    og = {k: v.clone() for k, v in sd_g0.items()}
    od = {k: v.clone() for k, v in sd_d0.items()}
    r = O.dcgan_step_grads(og, od, x, z1, z2, labels=labels, mode=mode, **kw)
    print("[%s r%d w%d] oracle vs reference: fake1 %.2e d_real %.2e loss_g %.2e  dgrad_real %.2e dgrad_fake %.2e ggrad %.2e" % (
        modname, res, width, (r["fake1"] - fx["fake1"]).abs().max(), (r["d_real"] - fx["d_real"]).abs().max(),
        (r["loss_g"] - fx["loss_g"]).abs().max(), maxdiff(r["d_grads_real"], fx["d_grads_real"]),
        maxdiff(r["d_grads_fake"], fx["d_grads_fake"]), maxdiff(r["g_grads"], fx["g_grads"])))
    for k in ("d_grads_real", "d_grads_fake", "g_grads", "d_grads_gstep"):
        fx[k] = pack_grads(fx[k])
    return fx


def trace_data(seed, steps, batch, res, z_dim):
    """Deterministic data stream of the trace fixture (CPU generator), regenerated by the tests from the seed."""
    gen = torch.Generator().manual_seed(seed + 7)
    xs = torch.rand(steps, batch, 3, res, res, generator=gen) * 2 - 1
    zs = torch.randn(steps, 2, batch, z_dim, generator=gen)
    return xs, zs


def dcgan_trace_fixture(mods, res, width, batch, steps, seed, z_dim=16):
    """`steps` iterations of main_dcgan.py:68-95 with Adam (lr 4e-4 / 1e-4, betas (0.5, 0.999): literals of :55-56)."""
    M = mods["dcgan"]
    torch.manual_seed(seed)
    netG = M.Generator(z_dim=z_dim, ngf=width, resolution=res)
    netD = M.Discriminator(ndf=width, resolution=res)
    crit = mods["criterion"].GANLoss("vanilla", 0.9, 0.1, 0.9)
    optG = torch.optim.Adam(netG.parameters(), lr=4e-4, betas=(0.5, 0.999))
    optD = torch.optim.Adam(netD.parameters(), lr=1e-4, betas=(0.5, 0.999))
    sd_g0, sd_d0 = clone_sd(netG), clone_sd(netD)
    xs, zs = trace_data(seed, steps, batch, res, z_dim)
    trace = []
    for i in range(steps):
        optD.zero_grad()
        out = netD(xs[i])
        l_real = crit(out, True)
        l_real.backward()
        fake = netG(zs[i, 0])
        out = netD(fake.detach())
        l_fake = crit(out, False)
        l_fake.backward()
        optD.step()
        optG.zero_grad()
        fake = netG(zs[i, 1])
        out = netD(fake)
        l_g = crit(out, False, True)
        l_g.backward()
        optG.step()
        trace.append([l_real.item(), l_fake.item(), l_g.item()])
    return {"sd_g": sd_g0, "sd_d": sd_d0, "seed": seed, "steps": steps, "batch": batch, "z_dim": z_dim,
            "trace": torch.tensor(trace), "res": res, "width": width,
            "x0_probe": xs[0, 0, 0, 0, :4].clone(), "buf_g_after": buffers_of(netG), "buf_d_after": buffers_of(netD),
            "g_linear_w_after": netG.state_dict()["linear.weight"][:4].clone()}


def sngan_fixture(mods, ch, batch, seed):
    M = mods["sngan_projection"]
    torch.manual_seed(seed)
    netG = M.ResNetGenerator(ch=ch, dim_z=16, bottom_width=2, img_dim=3, n_classes=10)
    netD = M.SNResNetProjectionDiscriminator(ch=ch, n_classes=10, img_dim=3)
    crit = mods["criterion"].GANLoss("hinge")
    netG.train(), netD.train()
    sd_g0, sd_d0 = clone_sd(netG), clone_sd(netD)
    gen = torch.Generator().manual_seed(seed + 1)
    x = torch.rand(batch, 3, 32, 32, generator=gen) * 2 - 1
    y = torch.randint(10, (batch,), generator=gen)
    z = torch.randn(batch, 16, generator=gen)
    c = torch.randint(10, (batch,), generator=gen)
    fx = {"sd_g": sd_g0, "sd_d": sd_d0, "x": x, "y": y, "z": z, "c": c, "ch": ch}
    # loop body of main_sngan.py:72-100
    netD.zero_grad()
    out = netD(x, y)
    l = crit(out, True)
    l.backward()
    fx["d_real"], fx["loss_real"], fx["d_grads_real"] = out.detach().clone(), l.detach().clone(), grads_of(netD)
    fake = netG(z, c)
    netD.zero_grad()
    out = netD(fake.detach(), c)
    l = crit(out, False)
    l.backward()
    fx["fake"], fx["d_fake"], fx["loss_fake"], fx["d_grads_fake"] = fake.detach().clone(), out.detach().clone(), l.detach().clone(), grads_of(netD)
    netG.zero_grad(), netD.zero_grad()
    out = netD(fake, c)
    l = crit(out, False, True)
    l.backward()
    fx["d_g"], fx["loss_g"], fx["g_grads"] = out.detach().clone(), l.detach().clone(), grads_of(netG)
    fx["buf_g_after"], fx["buf_d_after"] = buffers_of(netG), buffers_of(netD)
    for k in ("d_grads_real", "d_grads_fake", "g_grads"):
        fx[k] = pack_grads(fx[k])
    # oracle cross-check (forward)
    og = {k: v.clone() for k, v in sd_g0.items()}
    od = {k: v.clone() for k, v in sd_d0.items()}
    with torch.no_grad():
        o_real = O.sngan_discriminator(od, x, y)
        o_fake_img = O.sngan_generator(og, z, c, bottom_width=2)
        o_fake = O.sngan_discriminator(od, o_fake_img, c)
        o_g = O.sngan_discriminator(od, o_fake_img, c)  # third D forward of the loop (G step)
    print("[sngan ch%d] oracle vs reference: d_real %.2e fake %.2e d_fake %.2e d_g %.2e u-buffer(3 fwd) %.2e" % (
        ch, (o_real - fx["d_real"]).abs().max(), (o_fake_img - fx["fake"]).abs().max(),
        (o_fake - fx["d_fake"]).abs().max(), (o_g - fx["d_g"]).abs().max(),
        (od["block2.c1.weight_u"] - netD.state_dict()["block2.c1.weight_u"]).abs().max()))
    return fx


def acgan_fixture(mods, width, batch, seed):
    M = mods["acgan"]
    torch.manual_seed(seed)
    netG = M.Generator(z_dim=16, ngf=width, n_class=10)
    netD = M.Discriminator(ndf=width, n_class=10)
    crit = mods["criterion"].GANLoss("vanilla", 0.9, 0.1, 0.9)
    mse = torch.nn.MSELoss()
    sd_g0, sd_d0 = clone_sd(netG), clone_sd(netD)
    gen = torch.Generator().manual_seed(seed + 1)
    x = torch.rand(batch, 3, 64, 64, generator=gen) * 2 - 1
    y = torch.randint(0, 2, (batch, 10), generator=gen).float()
    z = torch.randn(batch, 16, generator=gen)
    fx = {"sd_g": sd_g0, "sd_d": sd_d0, "x": x, "y": y, "z": z, "width": width}
    # loop body of main_acgan.py:91-133
    netD.zero_grad()
    adv, cls = netD(x)
    l = crit(adv, True) + mse(cls, y) * 0.5
    l.backward()
    fx["d_real"], fx["d_real_cls"], fx["loss_real"], fx["d_grads_real"] = adv.detach().clone(), cls.detach().clone(), l.detach().clone(), grads_of(netD)
    fake = netG(z, y)
    netD.zero_grad()
    adv, cls = netD(fake.detach())
    l = crit(adv, False) + mse(cls, y) * 0.5
    l.backward()
    fx["fake"], fx["d_fake"], fx["loss_fake"], fx["d_grads_fake"] = fake.detach().clone(), adv.detach().clone(), l.detach().clone(), grads_of(netD)
    netG.zero_grad(), netD.zero_grad()
    adv, cls = netD(fake)
    l = crit(adv, False, True) + mse(cls, y) * 0.5
    l.backward()
    fx["d_g"], fx["loss_g"], fx["g_grads"] = adv.detach().clone(), l.detach().clone(), grads_of(netG)
    with torch.no_grad():
        o_fake = O.dcgan_generator({k: v.clone() for k, v in sd_g0.items()}, z, y, acgan=True)
        o_adv, o_cls = O.dcgan_discriminator({k: v.clone() for k, v in sd_d0.items()}, x, acgan=True)
    for k in ("d_grads_real", "d_grads_fake", "g_grads"):
        fx[k] = pack_grads(fx[k])
    print("[acgan w%d] oracle vs reference: fake %.2e d_real %.2e cls %.2e" % (
        width, (o_fake - fx["fake"]).abs().max(), (o_adv - fx["d_real"]).abs().max(), (o_cls - fx["d_real_cls"]).abs().max()))
    return fx


def sngan_loop_data(seed, steps, batch, z_dim, n_class=10):
    """Deterministic data stream of the main_sngan loop fixture, regenerated by the tests from the seed."""
    gen = torch.Generator().manual_seed(seed + 11)
    xs = torch.rand(steps, batch, 3, 32, 32, generator=gen) * 2 - 1
    ys = torch.randint(n_class, (steps, batch), generator=gen)
    zs = torch.randn(steps, batch, z_dim, generator=gen)
    cs = torch.randint(n_class, (steps, batch), generator=gen)
    return xs, ys, zs, cs


def sngan_loop_fixture(mods, ch, batch, steps, seed, n_disc_update=2, z_dim=16):
    """`steps` iterations of main_sngan.py:65-100 (Adam lr 2e-4, betas (0, 0.999): argparse defaults :19-21; hinge loss;
    the G step every `n_disc_update` iterations re-using the fake batch's graph) on the unmodified reference."""
    M = mods["sngan_projection"]
    torch.manual_seed(seed)
    netG = M.ResNetGenerator(ch=ch, dim_z=z_dim, bottom_width=2, img_dim=3, n_classes=10)
    netD = M.SNResNetProjectionDiscriminator(ch=ch, n_classes=10, img_dim=3)
    crit = mods["criterion"].GANLoss("hinge")
    optG = torch.optim.Adam(netG.parameters(), lr=2e-4, betas=(0.0, 0.999))
    optD = torch.optim.Adam(netD.parameters(), lr=2e-4, betas=(0.0, 0.999))
    netG.train(), netD.train()
    sd_g0, sd_d0 = clone_sd(netG), clone_sd(netD)
    xs, ys, zs, cs = sngan_loop_data(seed, steps, batch, z_dim)
    trace = []
    for i in range(steps):
        optD.zero_grad()
        outD = netD(xs[i], ys[i])
        Dx = outD.mean().item()
        lossD_real = crit(outD, True)
        lossD_real.backward()
        outG = netG(zs[i], cs[i])
        outD = netD(outG.detach(), cs[i])
        Dgz1 = outD.mean().item()
        lossD_fake = crit(outD, False)
        lossD_fake.backward()
        optD.step()
        row = [lossD_real.item(), lossD_fake.item(), float("nan"), Dx, Dgz1, float("nan")]
        if i % n_disc_update == 0:
            optG.zero_grad()
            outD = netD(outG, cs[i])
            row[5] = outD.mean().item()
            lossG = crit(outD, False, True)
            lossG.backward()
            optG.step()
            row[2] = lossG.item()
        trace.append(row)
    fx = {"sd_g": sd_g0, "sd_d": sd_d0, "seed": seed, "steps": steps, "batch": batch, "z_dim": z_dim, "ch": ch,
          "n_disc_update": n_disc_update, "trace": torch.tensor(trace), "x0_probe": xs[0, 0, 0, 0, :4].clone(),
          "buf_g_after": buffers_of(netG), "buf_d_after": buffers_of(netD),
          "g_l1_w_after": netG.state_dict()["l1.weight"][:4].clone(),
          "d_l6_w_after": netD.state_dict()["l6.weight_orig"].clone()}
    tr = O.CpuSnganTrainer(sd_g0, sd_d0, n_disc_update=n_disc_update, bottom_width=2)
    ot = []
    for i in range(steps):
        r = tr.step(xs[i], ys[i], zs[i], cs[i])
        ot.append([float("nan") if v is None else v for v in r])
    d = (torch.tensor(ot) - fx["trace"])
    print("[sngan loop ch%d] oracle trainer vs reference over %d steps: max |d| %.2e" % (ch, steps, d[~d.isnan()].abs().max()))
    return fx


def acgan_loop_data(seed, steps, batch, z_dim, n_class=10):
    gen = torch.Generator().manual_seed(seed + 13)
    xs = torch.rand(steps, batch, 3, 64, 64, generator=gen) * 2 - 1
    ys = torch.randint(0, 2, (steps, batch, n_class), generator=gen).float()   # CelebA-style float attribute vectors
    zs = torch.randn(steps, batch, z_dim, generator=gen)
    return xs, ys, zs


def acgan_loop_fixture(mods, width, batch, steps, seed, z_dim=16):
    """`steps` iterations of main_acgan.py:84-133 (Adam lr 4e-4 / 1e-4, betas (0.5, 0.999): :59-60; vanilla GANLoss with
    0.9 / 0.1 / 0.9 labels + 0.5 x MSELoss on the auxiliary head) on the unmodified reference."""
    M = mods["acgan"]
    torch.manual_seed(seed)
    netG = M.Generator(z_dim=z_dim, ngf=width, n_class=10)
    netD = M.Discriminator(ndf=width, n_class=10)
    criterion_adv = mods["criterion"].GANLoss("vanilla", 0.9, 0.1, 0.9)
    criterion_aux = torch.nn.MSELoss()
    optG = torch.optim.Adam(netG.parameters(), lr=4e-4, betas=(0.5, 0.999))
    optD = torch.optim.Adam(netD.parameters(), lr=1e-4, betas=(0.5, 0.999))
    netG.train(), netD.train()
    sd_g0, sd_d0 = clone_sd(netG), clone_sd(netD)
    xs, ys, zs = acgan_loop_data(seed, steps, batch, z_dim)
    trace = []
    for i in range(steps):
        img_real, lbl_real = xs[i], ys[i]
        optD.zero_grad()
        outD_adv, outD_cls = netD(img_real)
        Dx = torch.sigmoid(outD_adv).mean().item()
        lossD_real_adv = criterion_adv(outD_adv, True)
        lossD_real_aux = criterion_aux(outD_cls, lbl_real)
        (lossD_real_adv + lossD_real_aux * 0.5).backward()
        c = lbl_real.float()
        outG = netG(zs[i], c)
        outD_adv, outD_cls = netD(outG.detach())
        Dgz1 = torch.sigmoid(outD_adv).mean().item()
        lossD_fake_adv = criterion_adv(outD_adv, False)
        lossD_fake_aux = criterion_aux(outD_cls, c)
        (lossD_fake_adv + lossD_fake_aux * 0.5).backward()
        optD.step()
        optG.zero_grad()
        outD_adv, outD_cls = netD(outG)
        Dgz2 = torch.sigmoid(outD_adv).mean().item()
        lossG_adv = criterion_adv(outD_adv, False, True)
        lossG_aux = criterion_aux(outD_cls, c)
        (lossG_adv + lossG_aux * 0.5).backward()
        optG.step()
        trace.append([(lossD_real_adv + lossD_fake_adv).item(), (lossD_real_aux + lossD_fake_aux).item(),
                      lossG_adv.item(), lossG_aux.item(), Dx, Dgz1, Dgz2])
    fx = {"sd_g": sd_g0, "sd_d": sd_d0, "seed": seed, "steps": steps, "batch": batch, "z_dim": z_dim, "width": width,
          "trace": torch.tensor(trace), "x0_probe": xs[0, 0, 0, 0, :4].clone(),
          "buf_g_after": buffers_of(netG), "buf_d_after": buffers_of(netD),
          "g_linear_w_after": netG.state_dict()["linear.weight"][:4].clone(),
          "d_aux_w_after": netD.state_dict()["out_aux.weight"].clone()}
    tr = O.CpuAcganTrainer(sd_g0, sd_d0)
    ot = torch.tensor([tr.step(xs[i], ys[i], zs[i]) for i in range(steps)])
    print("[acgan loop w%d] oracle trainer vs reference over %d steps: max |d| %.2e" % (width, steps, (ot - fx["trace"]).abs().max()))
    return fx


def checkpoint_fixture(mods, seed=12):
    """A checkpoint exactly as the reference writes it (save_model, main_dcgan.py:107-123 / main_sngan.py:107-123:
    {'state_dict': {'generator', 'discriminator'}, 'optimizer': {'generator', 'discriminator'}, 'epoch'}) after two
    optimiser steps, for dcgan (res 32, width 4) and the SNGAN projection pair (ch 4) — so the mirrors can be checked to
    load it, continue from it and write the same layout back. The losses of the NEXT iteration on the reference are stored
    as the resume check."""
    out = {}
    M = mods["dcgan"]
    torch.manual_seed(seed)
    netG, netD = M.Generator(z_dim=16, ngf=4, resolution=32), M.Discriminator(ndf=4, resolution=32)
    crit = mods["criterion"].GANLoss("vanilla", 0.9, 0.1, 0.9)
    optG = torch.optim.Adam(netG.parameters(), lr=4e-4, betas=(0.5, 0.999))
    optD = torch.optim.Adam(netD.parameters(), lr=1e-4, betas=(0.5, 0.999))
    xs, zs = trace_data(seed, 3, 8, 32, 16)

    def it(i):
        optD.zero_grad()
        l1 = crit(netD(xs[i]), True)
        l1.backward()
        l2 = crit(netD(netG(zs[i, 0]).detach()), False)
        l2.backward()
        optD.step()
        optG.zero_grad()
        l3 = crit(netD(netG(zs[i, 1])), False, True)
        l3.backward()
        optG.step()
        return [l1.item(), l2.item(), l3.item()]

    it(0), it(1)
    out["dcgan"] = {"checkpoint": {"state_dict": {"generator": clone_sd(netG), "discriminator": clone_sd(netD)},
                                   "optimizer": {"generator": optG.state_dict(), "discriminator": optD.state_dict()},
                                   "epoch": 0},
                    "seed": seed, "next_losses": it(2)}
    S = mods["sngan_projection"]
    torch.manual_seed(seed + 1)
    sG = S.ResNetGenerator(ch=4, dim_z=16, bottom_width=2, img_dim=3, n_classes=10)
    sD = S.SNResNetProjectionDiscriminator(ch=4, n_classes=10, img_dim=3)
    oG = torch.optim.Adam(sG.parameters(), lr=2e-4, betas=(0.0, 0.999))
    oD = torch.optim.Adam(sD.parameters(), lr=2e-4, betas=(0.0, 0.999))
    hinge = mods["criterion"].GANLoss("hinge")
    x, y, z, c = (t[0] for t in sngan_loop_data(seed, 1, 4, 16))
    oD.zero_grad()
    hinge(sD(x, y), True).backward()
    fake = sG(z, c)
    hinge(sD(fake.detach(), c), False).backward()
    oD.step()
    oG.zero_grad()
    hinge(sD(fake, c), False, True).backward()
    oG.step()
    out["sngan"] = {"checkpoint": {"state_dict": {"generator": clone_sd(sG), "discriminator": clone_sd(sD)},
                                   "optimizer": {"generator": oG.state_dict(), "discriminator": oD.state_dict()},
                                   "epoch": 0}}
    import copy

    return copy.deepcopy(out)


def ganloss_fixture(mods):
    G = mods["criterion"].GANLoss
    gen = torch.Generator().manual_seed(5)
    pred = torch.randn(16, 1, generator=gen) * 2
    rows = []
    for mode, labels in (("vanilla", (0.9, 0.1, 0.9)), ("vanilla", (1.0, 0.0, 1.0)), ("lsgan", (1.0, 0.0, 1.0)), ("hinge", (1.0, 0.0, 1.0))):
        crit = G(mode, *labels)
        for is_real, is_gen in ((True, False), (False, False), (False, True)):
            p = pred.clone().requires_grad_(True)
            l = crit(p, is_real, is_gen)
            l.backward()
            o = O.gan_loss(mode, pred, is_real, is_gen, *labels)
            assert abs(o.item() - l.item()) < 1e-6, (mode, is_real, is_gen)
            rows.append({"mode": mode, "labels": labels, "is_real": is_real, "is_generator": is_gen,
                         "loss": l.detach().clone(), "dpred": p.grad.clone()})
    return {"pred": pred, "rows": rows, "buffers": sorted(G("vanilla").state_dict().keys())}


def key_lists(mods):
    out = {}

    def desc(net):
        return [[k, list(v.shape), str(v.dtype)] for k, v in net.state_dict().items()]

    import contextlib
    import io

    with contextlib.redirect_stdout(io.StringIO()):
        out["dcgan.Generator"] = desc(mods["dcgan"].Generator())
        out["dcgan.Discriminator"] = desc(mods["dcgan"].Discriminator())
        out["dcgan.Generator@32"] = desc(mods["dcgan"].Generator(resolution=32))
        out["dcgan.Discriminator@32"] = desc(mods["dcgan"].Discriminator(resolution=32))
        out["dcgan.Generator@128"] = desc(mods["dcgan"].Generator(ngf=8, resolution=128))
        out["dcgan.Discriminator@128"] = desc(mods["dcgan"].Discriminator(ndf=8, resolution=128))
        out["dcgan_specnorm.Generator"] = desc(mods["dcgan_specnorm"].Generator(ngf=8))
        out["dcgan_specnorm.Discriminator"] = desc(mods["dcgan_specnorm"].Discriminator(ndf=8))
        out["dcgan_specnorm.Generator@32"] = desc(mods["dcgan_specnorm"].Generator(resolution=32))
        out["dcgan_specnorm.Discriminator@32"] = desc(mods["dcgan_specnorm"].Discriminator(resolution=32))
        out["sngan_projection.ResNetGenerator"] = desc(mods["sngan_projection"].ResNetGenerator(n_classes=10, bottom_width=2))
        out["sngan_projection.SNResNetProjectionDiscriminator"] = desc(mods["sngan_projection"].SNResNetProjectionDiscriminator(n_classes=10))
        out["acgan.Generator"] = desc(mods["acgan"].Generator())
        out["acgan.Discriminator"] = desc(mods["acgan"].Discriminator())
        out["acgan.Generator@32c5"] = desc(mods["acgan"].Generator(ngf=8, resolution=32, n_class=5))
        out["acgan.Discriminator@32c5"] = desc(mods["acgan"].Discriminator(ndf=8, resolution=32, n_class=5))
        out["sngan_projection.ResNetGenerator@uncond"] = desc(mods["sngan_projection"].ResNetGenerator(ch=8, n_classes=0))
        out["sngan_projection.SNResNetProjectionDiscriminator@uncond"] = desc(mods["sngan_projection"].SNResNetProjectionDiscriminator(ch=8, n_classes=0))
        U = mods["dcgan_specnorm_up"]
        out["dcgan_specnorm_up.Generator"] = desc(U.Generator(ngf=8))
        out["dcgan_specnorm_up.Discriminator"] = desc(U.Discriminator(ndf=8))
        out["dcgan_specnorm_up.Generator@32"] = desc(U.Generator(ngf=8, resolution=32))
        out["dcgan_specnorm_up.Discriminator@128"] = desc(U.Discriminator(ndf=4, resolution=128))
        out["dcgan_blur.Generator"] = desc(mods["dcgan_blur"].Generator())
        out["dcgan_blur.Discriminator"] = desc(mods["dcgan_blur"].Discriminator())
        out["dcgan_blur.Generator@32"] = desc(mods["dcgan_blur"].Generator(resolution=32))
        out["dcgan_blur.Discriminator@32"] = desc(mods["dcgan_blur"].Discriminator(resolution=32))
        # RNG-order parity of every family: (sum, first element) of every state_dict entry of a seed-7 construction
        def probes(build):
            torch.manual_seed(7)
            nets = build()
            return [[[k, float(v.double().sum()), float(v.flatten()[0])] for k, v in n.state_dict().items()] for n in nets]

        S, A, N = mods["sngan_projection"], mods["acgan"], mods["dcgan_specnorm"]
        out["same_seed_probes"] = {
            "sngan_projection": probes(lambda: (S.ResNetGenerator(ch=8, dim_z=16, bottom_width=2, n_classes=10),
                                                S.SNResNetProjectionDiscriminator(ch=8, n_classes=10))),
            "sngan_projection@uncond": probes(lambda: (S.ResNetGenerator(ch=8, dim_z=16, n_classes=0),
                                                       S.SNResNetProjectionDiscriminator(ch=8, n_classes=0))),
            "acgan": probes(lambda: (A.Generator(z_dim=16, ngf=8, n_class=10), A.Discriminator(ndf=8, n_class=10))),
            "dcgan_specnorm": probes(lambda: (N.Generator(z_dim=16, ngf=8, resolution=32), N.Discriminator(ndf=8, resolution=32))),
            "dcgan_specnorm_up": probes(lambda: (mods["dcgan_specnorm_up"].Generator(z_dim=16, ngf=8, resolution=32),
                                                 mods["dcgan_specnorm_up"].Discriminator(ndf=8, resolution=32))),
        }
        # RNG-order parity: first weights of a seed-0 construction
        torch.manual_seed(0)
        g = mods["dcgan"].Generator(ngf=8, resolution=32)
        d = mods["dcgan"].Discriminator(ndf=8, resolution=32)
        out["seed0_probe"] = {"g_linear_w0": g.linear.weight.flatten()[:8].tolist(),
                              "g_blocks0_w0": g.blocks[0][0].weight.flatten()[:8].tolist(),
                              "d_blocks0_w0": d.blocks[0][0].weight.flatten()[:8].tolist(),
                              "d_out_w0": d.out_layer.weight.flatten()[:8].tolist(),
                              "g_param_count": g.param_count, "d_param_count": d.param_count}
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    mods = ref_modules()
    torch.set_num_threads(4)
    import contextlib
    import io

    quiet = io.StringIO()
    with contextlib.redirect_stdout(quiet):
        pass
    fixtures = {
        "dcgan_r32_w4.pt": lambda: dcgan_like_fixture(mods, "dcgan", 32, 4, 8, "vanilla", (0.9, 0.1, 0.9), 0, z_dim=100),
        "dcgan_r64_w4.pt": lambda: dcgan_like_fixture(mods, "dcgan", 64, 4, 2, "vanilla", (0.9, 0.1, 0.9), 1, z_dim=16),
        "snd_r32_w4.pt": lambda: dcgan_like_fixture(mods, "dcgan_specnorm", 32, 4, 8, "hinge", (1.0, 0.0, 1.0), 2, z_dim=16),
        "snd_up_r32_w4.pt": lambda: dcgan_like_fixture(mods, "dcgan_specnorm_up", 32, 4, 8, "hinge", (1.0, 0.0, 1.0), 7, z_dim=16),
        "dcgan_blur_r32_w8.pt": lambda: dcgan_like_fixture(mods, "dcgan_blur", 32, 8, 4, "vanilla", (0.9, 0.1, 0.9), 6, z_dim=16),
        "dcgan_trace_r32_w4.pt": lambda: dcgan_trace_fixture(mods, 32, 4, 8, 20, 3),
        "sngan_proj_ch8.pt": lambda: sngan_fixture(mods, 8, 2, 4),
        "acgan_r64_w4.pt": lambda: acgan_fixture(mods, 4, 2, 5),
        "ganloss.pt": lambda: ganloss_fixture(mods),
        "checkpoint_ref_layout.pt": lambda: checkpoint_fixture(mods),
        "sngan_loop_ch8.pt": lambda: sngan_loop_fixture(mods, 8, 8, 12, 8),
        "acgan_loop_r64_w4.pt": lambda: acgan_loop_fixture(mods, 4, 8, 12, 9),
    }
    only = [a for a in sys.argv[1:] if not a.startswith("-")]   # optional: fixture file names to (re)generate
    for name, fn in fixtures.items():
        if only and name not in only:
            continue
        fx = fn()
        path = os.path.join(OUT, name)
        torch.save(fx, path)
        print("wrote %s (%.1f KB)" % (path, os.path.getsize(path) / 1024))
    with open(os.path.join(OUT, "state_dict_keys.json"), "w") as f:
        json.dump(key_lists(mods), f, indent=0)
    print("wrote state_dict_keys.json")


if __name__ == "__main__":
    main()
