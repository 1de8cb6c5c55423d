"""CPU emulation of the forward-precision policies of the DCGAN-64 step (no GPU needed): which operand / storage
roundings cost the north_star gradient-cosine bar, and what is the cheapest policy that keeps it?

The real kernels differ from the fp32 reference only through bf16 roundings:
  operands   conv / linear GEMMs read bf16 tiles; the `bf16x3` mode reads hi+lo pairs of BOTH operands (3 MMAs);
  y storage  the pre-BatchNorm conv output is stored bf16 (`bf16`) or fp32 (`bf16x3`); batch statistics always come
             from the fp32 accumulators;
  a storage  post-activation tensors are stored as one bf16 (`bf16`) or a hi+lo pair (`bf16x3`);
  backward   always single bf16: da / dy tensors are bf16, dgrad reads (dy, W_hi), wgrad reads (dy, x_hi), fp32 accumulate.
This script restates those roundings on fp32 CPU tensors (custom autograd nodes with explicit backward operands) for a
per-layer POLICY = (x bits, w bits, y storage, a storage), calibrates the two shipped modes against the numbers measured
on the B200 (tests/test_gpu_precision.py), and then scores cheaper candidates with the same metrics:
activation max-rel-error and the D-real / D-fake / G-step global gradient cosines against the fp32 oracle.

    python tools/precision_study.py [--batch 32] [--only substring|substring] [--sweep]      # a few seconds per policy

It is an analysis tool (imports oracle/: test infrastructure); nothing in the product path depends on it."""
import argparse
import contextlib
import io
import os
import sys
import time

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import gan_oracle as O  # noqa: E402

EPS = 1e-5


def q(x):
    return x.bfloat16().float()


def split16(x):
    hi = q(x)
    return hi + q(x - hi)


def operand(x, fmt):
    """GEMM operand formats: 8 = one bf16, 16 = bf16 hi + bf16 lo, "h" = one fp16 (11 significant bits, narrower range:
    values below 6e-8 flush), "hh" = fp16 hi + fp16 lo."""
    if fmt == 8:
        return q(x)
    if fmt == 16:
        return split16(x)
    if fmt == "h+8":               # storage of an activation whose consumers split it themselves: ~16+ bits
        fmt = "hh"
    hi = x.half().float()
    return hi if fmt == "h" else hi + (x - hi).half().float()


def ste(x, xq):
    """value xq, gradient of x"""
    return x + (xq - x).detach()


class RoundGrad(torch.autograd.Function):
    """identity whose incoming gradient is stored as bf16 (the da / dy tensors between nodes)"""

    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return q(g)


def q8(x):
    """e4m3 with a per-tensor power-of-two scale that puts the largest magnitude near the top of the format's range (a
    separate accumulator / block scale would undo it exactly); 4 significant bits."""
    m = x.abs().max().item()
    if m == 0.0:
        return x
    s = 2.0 ** (8 - int(torch.ceil(torch.log2(torch.tensor(m))).item()))      # max |x| * s in (128, 256]
    return (x * s).to(torch.float8_e4m3fn).float() / s


def product(X, W, op, fmt):
    """fmt "h+8": main product on fp16 operands + the two first-order correction products x_lo*w_hi and x_hi*w_lo on
    e4m3 operands (kind::f8f6f4 runs at twice the 16-bit rate, so the three products cost two 16-bit MMAs)."""
    xh, wh = X.half().float(), W.half().float()
    return op(xh, wh) + op(q8(X - xh), q8(wh)) + op(q8(xh), q8(W - wh))


class ConvEmu(torch.autograd.Function):
    """conv / conv-transpose (k4 s2 p1) or linear with operand precisions (xbits, wbits) in {8, 16}; backward with single
    bf16 operands: dx = dgrad(q(dy), q(w)), dw = wgrad(q(dy), q(x)) in fp32 accumulation."""

    @staticmethod
    def forward(ctx, x, w, kind, xbits, wbits):
        ctx.kind = kind
        ctx.save_for_backward(q(x), q(w))
        if xbits == "h+8":
            return product(x, w, lambda a, b: ConvEmu.op(a, b, kind), xbits)
        X, W = operand(x, xbits), operand(w, wbits)
        return ConvEmu.op(X, W, kind)

    @staticmethod
    def op(X, W, kind):
        if kind == "conv":
            return F.conv2d(X, W, None, stride=2, padding=1)
        if kind == "convT":
            return F.conv_transpose2d(X, W, None, stride=2, padding=1)
        if kind == "c3":
            return F.conv2d(X, W, None, stride=1, padding=1)
        if kind == "c1":
            return F.conv2d(X, W, None)
        return F.linear(X, W)

    @staticmethod
    def backward(ctx, dy):
        xh, wh = ctx.saved_tensors
        with torch.enable_grad():
            xh = xh.detach().requires_grad_(True)
            wh = wh.detach().requires_grad_(True)
            y = ConvEmu.op(xh, wh, ctx.kind)
            dx, dw = torch.autograd.grad(y, (xh, wh), q(dy))
        return dx, dw, None, None, None


def conv_emu(x, w, b, kind, pol):
    y = ConvEmu.apply(x, w, kind, pol["x"], pol["w"])
    if b is not None:
        y = y + (b[None, :, None, None] if y.dim() == 4 else b)
    return y


def store_act(a, pol):
    """post-activation storage (hi or hi+lo) + bf16 gradient storage on the way back"""
    aq = operand(a, pol["a"])
    return RoundGrad.apply(ste(a, aq))


def bn_act(y, gamma, beta, act, pol):
    """BatchNorm(train) + activation as the kernels do it: statistics from the fp32 accumulators, normalisation applied to
    the STORED y (bf16 or fp32); the incoming gradient of y is stored as bf16."""
    mean = y.mean(dim=(0, 2, 3))
    var = y.var(dim=(0, 2, 3), unbiased=False)
    ys = RoundGrad.apply(ste(y, q(y)) if pol["y"] == "bf16" else (ste(y, y.half().float()) if pol["y"] == "f16" else y))
    xhat = (ys - mean[None, :, None, None]) * torch.rsqrt(var[None, :, None, None] + EPS)
    z = xhat * gamma[None, :, None, None] + beta[None, :, None, None]
    return store_act(F.relu(z) if act == "relu" else F.leaky_relu(z, 0.2), pol)


def convT_image(x, w, b, pol):
    """last generator layer: 1-tap GEMM per kernel tap into a column buffer (bf16 in `bf16`, fp32 in `bf16x3`), then
    col2im + bias + tanh in fp32"""
    if pol["col"] == "f32":
        return torch.tanh(conv_emu(x, w, b, "convT", pol))
    out = None
    for kh in range(4):
        for kw in range(4):
            m = torch.zeros_like(w)
            m[:, :, kh, kw] = 1.0
            part = ConvEmu.apply(x, w * m, "convT", pol["x"], pol["w"])
            part = RoundGrad.apply(ste(part, q(part)))          # the bf16 column buffer (zeros stay zeros)
            out = part if out is None else out + part
    return torch.tanh(out + b[None, :, None, None])


def generator(sd, z, P):
    h = F.relu(conv_emu(z, sd["linear.weight"], sd["linear.bias"], "linear", P["g_lin"]))
    h = store_act(h, P["g_lin"]).view(h.size(0), -1, 4, 4)
    i = 0
    while ("blocks.%d.0.bias" % i) in sd:
        p, pol = "blocks.%d." % i, P["g%d" % i]
        y = conv_emu(h, sd[p + "0.weight"], sd[p + "0.bias"], "convT", pol)
        h = bn_act(y, sd[p + "1.weight"], sd[p + "1.bias"], "relu", pol)
        i += 1
    return convT_image(h, sd["out_layer.0.weight"], sd["out_layer.0.bias"], P["g_out"])


def discriminator(sd, x, P):
    h, i = x, 0
    while ("blocks.%d.0.bias" % i) in sd:
        p, pol = "blocks.%d." % i, P["d%d" % i]
        y = conv_emu(h, sd[p + "0.weight"], sd[p + "0.bias"], "conv", pol)
        if (p + "1.weight") in sd:
            h = bn_act(y, sd[p + "1.weight"], sd[p + "1.bias"], "lrelu", pol)
        else:
            h = store_act(F.leaky_relu(y, 0.2), pol)
        i += 1
    return F.linear(h.sum(dim=(2, 3)), sd["out_layer.weight"], sd["out_layer.bias"])


LAYERS = ["g_lin", "g0", "g1", "g2", "g_out", "d0", "d1", "d2", "d3"]
BF16 = dict(x=8, w=8, y="bf16", a=8, col="bf16")
X3 = dict(x=16, w=16, y="f32", a=16, col="f32")


def policy(base, **over):
    P = {k: dict(base) for k in LAYERS}
    for k, v in over.items():
        for layer in (LAYERS if k == "all" else [k]):
            P[layer] = dict(P[layer], **v)
    return P


def mmas(P, fake_passes_only=True):
    """relative forward MMA work of the generated-image passes (bf16 = 1.0), weighting every layer equally (the DCGAN-64
    layers have equal FLOPs, the image-side ones are negligible)"""
    big = ["g0", "g1", "g2", "d1", "d2", "d3"]
    n = {8: 1, "h": 1, 16: 2, "hh": 2}
    return sum(2 if P[k]["x"] == "h+8" else n[P[k]["x"]] * n[P[k]["w"]] - (n[P[k]["x"]] * n[P[k]["w"]] == 4)
               for k in big) / len(big)


POLICIES = {
    "bf16": policy(BF16),
    "bf16x3": policy(X3),
    # which single ingredient matters?
    "bf16 + y fp32": policy(BF16, all=dict(y="f32")),
    "bf16 + y fp32 + a hi/lo": policy(BF16, all=dict(y="f32", a=16, col="f32")),
    "x 16 only (2 MMA), y fp32, a hi/lo": policy(X3, all=dict(w=8)),
    "w 16 only (2 MMA), y fp32": policy(BF16, all=dict(w=16, y="f32", col="f32")),
    # x3 only where it matters?
    "x3 in D only": policy(BF16, d0=X3, d1=X3, d2=X3, d3=X3, g_out=dict(a=16, col="f32")),
    "x3 in G only": policy(X3, d0=BF16, d1=BF16, d2=BF16, d3=BF16),
    # fp16 forward operands (tcgen05 kind::f16 takes either): 11 significant bits in ONE MMA; backward stays bf16
    "fp16 operands (1 MMA), y fp32, a fp16": policy(BF16, all=dict(x="h", w="h", y="f32", a="h", col="f32")),
    "fp16 operands (1 MMA), y bf16, a fp16": policy(BF16, all=dict(x="h", w="h", y="bf16", a="h", col="f32")),
    "fp16 x, fp16 hi+lo w (2 MMA), y fp32": policy(BF16, all=dict(x="h", w="hh", y="f32", a="h", col="f32")),
    "fp16 hi+lo x, fp16 w (2 MMA), y fp32": policy(BF16, all=dict(x="hh", w="h", y="f32", a="hh", col="f32")),
    "fp16x3 (3 MMA), y fp32": policy(BF16, all=dict(x="hh", w="hh", y="f32", a="hh", col="f32")),
    "fp16 operands (1 MMA), y fp16, a fp16": policy(BF16, all=dict(x="h", w="h", y="f16", a="h", col="f32")),
    # main product fp16 + both correction products on e4m3 operands at the doubled fp8 rate (2 MMA-equivalents)
    "fp16 + 2 x e4m3 corrections (2 MMA-eq)": policy(BF16, all=dict(x="h+8", w="h+8", y="f32", a="h+8", col="f32")),
}


def cos(ga, gb, skip_prebn_of):
    num = da = db = 0.0
    for k, r in gb.items():
        if k.endswith(".0.bias") and k.replace(".0.bias", ".1.weight") in skip_prebn_of:
            continue
        g = ga[k].double()
        r = r.double()
        num += (g * r).sum().item()
        da += (g * g).sum().item()
        db += (r * r).sum().item()
    return num / (da ** 0.5 * db ** 0.5 + 1e-30)


def relerr(a, b):
    return ((a - b).abs().max() / (b.abs().max() + 1e-30)).item()


def run(P, sd_g, sd_d, x, z1, z2, ref):
    pg = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v) for k, v in sd_g.items()}
    pd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v) for k, v in sd_d.items()}
    dl = [k for k, v in pd.items() if torch.is_tensor(v) and v.requires_grad]
    gl = [k for k, v in pg.items() if torch.is_tensor(v) and v.requires_grad]
    out = {}
    d_real = discriminator(pd, x, P)
    gr = torch.autograd.grad(O.gan_loss("vanilla", d_real, True, False, 0.9, 0.1, 0.9), [pd[k] for k in dl], allow_unused=True)
    out["act D(x)"] = relerr(d_real.detach(), ref["d_real"])
    out["cos D-real"] = cos({k: g for k, g in zip(dl, gr) if g is not None}, ref["d_grads_real"], pd)
    with torch.no_grad():
        fake1 = generator(pg, z1, P)
    out["act G(z)"] = relerr(fake1, ref["fake1"])
    d_fake = discriminator(pd, fake1, P)
    gf = torch.autograd.grad(O.gan_loss("vanilla", d_fake, False, False, 0.9, 0.1, 0.9), [pd[k] for k in dl], allow_unused=True)
    out["act D(G(z))"] = relerr(d_fake.detach(), ref["d_fake"])
    out["cos D-fake"] = cos({k: g for k, g in zip(dl, gf) if g is not None}, ref["d_grads_fake"], pd)
    d_g = discriminator(pd, generator(pg, z2, P), P)
    gg = torch.autograd.grad(O.gan_loss("vanilla", d_g, False, True, 0.9, 0.1, 0.9), [pg[k] for k in gl], allow_unused=True)
    out["cos G-step"] = cos({k: g for k, g in zip(gl, gg) if g is not None}, ref["g_grads"], pg)
    return out


def acgan_study(args):
    """The same one-step study for the ACGAN loop (main_acgan.py:84-133: one generator forward per iteration, adversarial
    BCE + 0.5 x MSE on the auxiliary head): does AcganStep's opt-in per-pass policy — real pass bf16, D's pass over the
    DETACHED fake batch fp16 (1 MMA), generator forward and G step bf16x3 — keep the bars? Reported per pass and for the
    accumulated real + fake discriminator gradient that optD.step() consumes."""
    from gan_playground_b200.models import acgan

    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        netG, netD = acgan.Generator(ngf=args.width, n_class=10), acgan.Discriminator(ndf=args.width, n_class=10)
    sd_g = {k: v.clone() for k, v in netG.state_dict().items()}
    sd_d = {k: v.clone() for k, v in netD.state_dict().items()}
    gen = torch.Generator().manual_seed(1)
    B = args.batch
    x = torch.rand(B, 3, 64, 64, generator=gen) * 2 - 1
    y = torch.randint(0, 2, (B, 10), generator=gen).float()
    z = torch.randn(B, 100, generator=gen)

    def gen_emu(sd, z, y, P):
        h = store_act(conv_emu(torch.cat([z, y], 1), sd["linear.weight"], sd["linear.bias"], "linear", P["g_lin"]), P["g_lin"])
        h = h.view(h.size(0), -1, 4, 4)
        for i in range(3):
            p, pol = "blocks.%d." % i, P["g%d" % i]
            h = bn_act(conv_emu(h, sd[p + "0.weight"], sd[p + "0.bias"], "convT", pol), sd[p + "1.weight"], sd[p + "1.bias"],
                       "relu", pol)
        return convT_image(h, sd["out_layer.0.weight"], sd["out_layer.0.bias"], P["g_out"])

    def dis_emu(sd, x, P):
        h = store_act(F.leaky_relu(conv_emu(x, sd["blocks.0.0.weight"], sd["blocks.0.0.bias"], "conv", P["d0"]), 0.2), P["d0"])
        for i in range(1, 4):
            p, pol = "blocks.%d." % i, P["d%d" % i]
            h = bn_act(conv_emu(h, sd[p + "0.weight"], sd[p + "0.bias"], "conv", pol), sd[p + "1.weight"], sd[p + "1.bias"],
                       "lrelu", pol)
        h = h.sum(dim=(2, 3))
        return F.linear(h, sd["out_layer.weight"], sd["out_layer.bias"]), F.linear(h, sd["out_aux.weight"], sd["out_aux.bias"])

    def objective(adv, cls, real, g=False):
        return O.gan_loss("vanilla", adv, real, g, 0.9, 0.1, 0.9) + 0.5 * F.mse_loss(cls, y)

    def leaves(sd):
        return {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v) for k, v in sd.items()}

    def grads(loss, p):
        ks = [k for k, v in p.items() if torch.is_tensor(v) and v.requires_grad]
        return {k: g for k, g in zip(ks, torch.autograd.grad(loss, [p[k] for k in ks], allow_unused=True)) if g is not None}

    def step(p_real, p_gen, p_fake, p_g, exact=False):
        pg, pd = leaves(sd_g), leaves(sd_d)
        G = (lambda z: O.dcgan_generator(pg, z, y, acgan=True)) if exact else (lambda z: gen_emu(pg, z, y, p_gen))
        D = (lambda x, P: O.dcgan_discriminator(pd, x, acgan=True)) if exact else (lambda x, P: dis_emu(pd, x, P))
        g_real = grads(objective(*D(x, p_real), True), pd)
        fake = G(z)
        g_fake = grads(objective(*D(fake.detach(), p_fake), False), pd)
        g_gen = grads(objective(*D(fake, p_g), False, True), pg)
        return g_real, g_fake, g_gen, pd, pg

    ref = step(None, None, None, None, exact=True)
    H = dict(x="h", w="h", y="f32", a="h", col="f32")
    print("ACGAN-64 width %d batch %d, emulated: gradient cosines against the fp32 oracle" % (args.width, B))
    print("%-58s %9s %9s %9s %9s" % ("policy", "D-real", "D-fake", "D r+f", "G-step"))
    for name, pol in (("bf16x3 on every pass", (policy(X3),) * 4),
                      ("mixed: real bf16 / D(fake.detach()) fp16 / G, G step bf16x3", (policy(BF16), policy(X3), policy(BF16, all=H), policy(X3))),
                      ("bf16 on every pass", (policy(BF16),) * 4)):
        r = step(*pol)
        both = {k: r[0][k] + r[1][k] for k in r[0]}
        both_ref = {k: ref[0][k] + ref[1][k] for k in ref[0]}
        print("%-58s %9.6f %9.6f %9.6f %9.6f" % (name, cos(r[0], ref[0], r[3]), cos(r[1], ref[1], r[3]),
                                                 cos(both, both_ref, r[3]), cos(r[2], ref[2], r[4])), flush=True)


def sndcgan_study(args):
    """One-step study for SN-DCGAN 32x32 with the hinge loss (cfg 3: models/dcgan_specnorm.py through the main_dcgan.py
    loop): spectral-normed conv weights W / sigma (sigma from fp32 GEMV kernels) are what the GEMMs round."""
    from gan_playground_b200.models import dcgan_specnorm as M

    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        netG, netD = M.Generator(ngf=args.width, resolution=32), M.Discriminator(ndf=args.width, resolution=32)
    sd_g = {k: v.clone() for k, v in netG.state_dict().items()}
    sd_d = {k: v.clone() for k, v in netD.state_dict().items()}
    gen = torch.Generator().manual_seed(1)
    B = args.batch
    x = torch.rand(B, 3, 32, 32, generator=gen) * 2 - 1
    z1, z2 = torch.randn(B, 100, generator=gen), torch.randn(B, 100, generator=gen)

    def snw(sd, p, dim):
        return O.spectral_norm_weight(sd, p, dim, True, update_buffers=False)

    def gen_emu(sd, z, P):
        h = store_act(F.relu(conv_emu(z, sd["linear.weight"], sd["linear.bias"], "linear", P["g_lin"])), P["g_lin"])
        h = h.view(h.size(0), -1, 4, 4)
        for i in range(2):
            p, pol = "blocks.%d." % i, P["g%d" % i]
            h = bn_act(conv_emu(h, snw(sd, p + "0.", 1), sd[p + "0.bias"], "convT", pol), sd[p + "1.weight"],
                       sd[p + "1.bias"], "relu", pol)
        return convT_image(h, snw(sd, "out_layer.0.", 1), sd["out_layer.0.bias"], P["g_out"])

    def dis_emu(sd, x, P):
        h = store_act(F.leaky_relu(conv_emu(x, snw(sd, "blocks.0.0.", 0), sd["blocks.0.0.bias"], "conv", P["d0"]), 0.2), P["d0"])
        for i in range(1, 3):
            p, pol = "blocks.%d." % i, P["d%d" % i]
            h = bn_act(conv_emu(h, snw(sd, p + "0.", 0), sd[p + "0.bias"], "conv", pol), sd[p + "1.weight"],
                       sd[p + "1.bias"], "lrelu", pol)
        return F.linear(h.flatten(1), sd["out_layer.weight"], sd["out_layer.bias"])

    def leaves(sd):
        return {k: (v.clone().requires_grad_(True) if v.is_floating_point() and not k.endswith(("_u", "_v", "running_mean", "running_var"))
                    else v.clone()) for k, v in sd.items()}

    def grads(loss, p):
        ks = [k for k, v in p.items() if torch.is_tensor(v) and v.requires_grad]
        return {k: g for k, g in zip(ks, torch.autograd.grad(loss, [p[k] for k in ks], allow_unused=True)) if g is not None}

    def step(p_real, p_fake, p_g):
        pg, pd = leaves(sd_g), leaves(sd_d)
        if p_real is None:
            G = lambda z, P: O.dcgan_generator(dict(pg), z, sn=True, training=True)
            D = lambda xx, P: O.dcgan_discriminator(dict(pd), xx, sn=True, flatten_head=True, training=True)
            keep = [{k: v.clone() for k, v in d.items() if k.endswith(("_u", "_v"))} for d in (pg, pd)]

            def restore():      # the oracle's hook advances u / v in place: every pass must start from the same vectors
                for d, kp in zip((pg, pd), keep):
                    for k, v in kp.items():
                        d[k].copy_(v)
        else:
            G, D, restore = (lambda z, P: gen_emu(pg, z, P)), (lambda xx, P: dis_emu(pd, xx, P)), (lambda: None)
        g_real = grads(O.gan_loss("hinge", D(x, p_real), True), pd)
        restore()
        with torch.no_grad():
            fake1 = G(z1, p_fake)
        restore()
        g_fake = grads(O.gan_loss("hinge", D(fake1, p_fake), False), pd)
        restore()
        fake2 = G(z2, p_g)
        restore()
        g_gen = grads(O.gan_loss("hinge", D(fake2, p_g), False, True), pg)
        restore()
        return g_real, g_fake, g_gen

    ref = step(None, None, None)
    H = dict(x="h", w="h", y="f32", a="h", col="f32")
    print("SN-DCGAN 32x32 width %d batch %d, hinge loss, emulated: gradient cosines against the fp32 oracle" % (args.width, B))
    print("%-58s %9s %9s %9s %9s" % ("policy (real pass / D-fake chain / G step)", "D-real", "D-fake", "D r+f", "G-step"))
    for name, pol in (("bf16x3 / bf16x3 / bf16x3", (policy(X3),) * 3),
                      ("bf16 / bf16x3 / bf16x3 (DcganStep today for this family)", (policy(BF16), policy(X3), policy(X3))),
                      ("bf16 / fp16 / bf16x3", (policy(BF16), policy(BF16, all=H), policy(X3))),
                      ("fp16 / fp16 / fp16", (policy(BF16, all=H),) * 3),
                      ("bf16 / bf16 / bf16", (policy(BF16),) * 3)):
        r = step(*pol)
        both = {k: r[0][k] + r[1][k] for k in r[0]}
        both_ref = {k: ref[0][k] + ref[1][k] for k in ref[0]}
        nb = {}
        print("%-58s %9.6f %9.6f %9.6f %9.6f" % (name, cos(r[0], ref[0], nb), cos(r[1], ref[1], nb), cos(both, both_ref, nb),
                                                 cos(r[2], ref[2], nb)), flush=True)


def sngan_study(args):
    """One-step study for the SNGAN projection pair (models/sngan_projection.py, loop of main_sngan.py:65-100; cfg 4:
    ch 64, 32x32, bottom_width 2, 10 classes). The ResNet nodes (functional_resnet.py) have no BatchNorm between most
    convs, so every conv output is STORED in the activation format and feeds the next GEMM directly: one policy =
    (operand format, storage format) for all layers. Which one do the nodes need to meet the bars?"""
    from gan_playground_b200.models import sngan_projection as M

    torch.manual_seed(0)
    ch, B = args.width, args.batch
    with contextlib.redirect_stdout(io.StringIO()):
        netG = M.ResNetGenerator(ch=ch, dim_z=128, bottom_width=2, img_dim=3, n_classes=10)
        netD = M.SNResNetProjectionDiscriminator(ch=ch, n_classes=10, img_dim=3)
    sd_g = {k: v.clone() for k, v in netG.state_dict().items()}
    sd_d = {k: v.clone() for k, v in netD.state_dict().items()}
    gen = torch.Generator().manual_seed(1)
    x = torch.rand(B, 3, 32, 32, generator=gen) * 2 - 1
    y = torch.randint(10, (B,), generator=gen)
    z = torch.randn(B, 128, generator=gen)
    c = torch.randint(10, (B,), generator=gen)

    def leaves(sd):
        return {k: (v.clone().requires_grad_(True) if v.is_floating_point() and not k.endswith(("_u", "_v", "running_mean", "running_var"))
                    else v.clone()) for k, v in sd.items()}

    def st(a, pol):                     # a stored activation (no fp32 copy exists anywhere)
        return store_act(a, pol)

    def cbn(sd, p, h, lab):
        mean, var = h.mean(dim=(0, 2, 3)), h.var(dim=(0, 2, 3), unbiased=False)
        xhat = (h - mean[None, :, None, None]) * torch.rsqrt(var[None, :, None, None] + EPS)
        e = F.embedding(lab, sd[p + "embed.weight"])
        C = h.shape[1]
        return e[:, :C, None, None] * xhat + e[:, C:, None, None]

    def gen_emu(sd, z, lab, pol):
        h = st(conv_emu(z, sd["l1.weight"], sd["l1.bias"], "linear", pol), pol).view(z.size(0), -1, 2, 2)
        for b in ("block2.", "block3.", "block4.", "block5."):
            a = st(F.interpolate(F.relu(cbn(sd, b + "b1.", h, lab)), scale_factor=2), pol)
            t = st(conv_emu(a, sd[b + "c1.weight"], sd[b + "c1.bias"], "c3", pol), pol)
            a2 = st(F.relu(cbn(sd, b + "b2.", t, lab)), pol)
            up = st(F.interpolate(h, scale_factor=2), pol)
            sc = st(conv_emu(up, sd[b + "c_sc.weight"], sd[b + "c_sc.bias"], "c1", pol), pol)
            h = st(conv_emu(a2, sd[b + "c2.weight"], sd[b + "c2.bias"], "c3", pol) + sc, pol)
        mean, var = h.mean(dim=(0, 2, 3)), h.var(dim=(0, 2, 3), unbiased=False)
        a = (h - mean[None, :, None, None]) * torch.rsqrt(var[None, :, None, None] + EPS)
        a = st(F.relu(a * sd["b6.weight"][None, :, None, None] + sd["b6.bias"][None, :, None, None]), pol)
        return torch.tanh(conv_emu(a, sd["l6.weight"], sd["l6.bias"], "c3", dict(pol, col="f32")))

    def snw(sd, p):
        return O.spectral_norm_weight(sd, p, 0, True, update_buffers=False)

    def dis_emu(sd, x, lab, pol):
        h1 = st(F.relu(conv_emu(x, snw(sd, "block1.c1."), sd["block1.c1.bias"], "c3", pol)), pol)
        s = st(conv_emu(x, snw(sd, "block1.c_sc."), sd["block1.c_sc.bias"], "c1", pol), pol)
        h = st(conv_emu(h1, snw(sd, "block1.c2."), sd["block1.c2.bias"], "c3", pol) + s, pol)
        h = st(F.avg_pool2d(h, 2), pol)
        for b in ("block2.", "block3.", "block4.", "block5."):
            a = st(F.relu(h), pol)
            t = st(F.relu(conv_emu(a, snw(sd, b + "c1."), sd[b + "c1.bias"], "c3", pol)), pol)
            sc = st(conv_emu(h, snw(sd, b + "c_sc."), sd[b + "c_sc.bias"], "c1", pol), pol)
            h = st(conv_emu(t, snw(sd, b + "c2."), sd[b + "c2.bias"], "c3", pol) + sc, pol)
            h = st(F.avg_pool2d(h, 2), pol)
        f = F.relu(h).sum(dim=(2, 3))
        out = F.linear(f, snw(sd, "l6."), sd["l6.bias"])
        return out + (F.embedding(lab, snw(sd, "l_y.")) * f).sum(dim=1, keepdim=True)

    def grads(loss, p):
        ks = [k for k, v in p.items() if torch.is_tensor(v) and v.requires_grad]
        return {k: g for k, g in zip(ks, torch.autograd.grad(loss, [p[k] for k in ks], allow_unused=True)) if g is not None}

    def step(pol):
        pg, pd = leaves(sd_g), leaves(sd_d)
        if pol is None:
            G = lambda: O.sngan_generator(pg, z, c, bottom_width=2)
            D = lambda xx, lab: O.sngan_discriminator({k: v for k, v in pd.items()}, xx, lab, training=True)
            # the oracle's hook advances u / v in place; every D call below must see the same u, v as the emulation
            def D(xx, lab, _pd=pd):
                keep = {k: v.clone() for k, v in _pd.items() if k.endswith(("_u", "_v"))}
                out = O.sngan_discriminator(_pd, xx, lab, training=True)
                for k, v in keep.items():
                    _pd[k].copy_(v)
                return out
        else:
            G = lambda: gen_emu(pg, z, c, pol)
            D = lambda xx, lab: dis_emu(pd, xx, lab, pol)
        d_real = D(x, y)
        g_real = grads(O.gan_loss("hinge", d_real, True), pd)
        fake = G()
        d_fake = D(fake.detach(), c)
        g_fake = grads(O.gan_loss("hinge", d_fake, False), pd)
        g_gen = grads(O.gan_loss("hinge", D(fake, c), False, True), pg)
        return dict(d_real=d_real.detach(), fake=fake.detach(), d_fake=d_fake.detach(), g_real=g_real, g_fake=g_fake,
                    g_gen=g_gen, pd=pd, pg=pg)

    ref = step(None)
    print("SNGAN projection ch %d batch %d, 32x32, emulated (operand format, storage format) for every layer" % (ch, B))
    print("%-44s | %9s %9s %9s | %9s %9s %9s" % ("policy", "D(x)", "G(z)", "D(G(z))", "cos Dreal", "cos Dfake", "cos Gstep"))
    pols = {"bf16 operands, bf16 storage (today)": dict(x=8, w=8, a=8, col="bf16"),
            "bf16 operands, hi/lo storage": dict(x=8, w=8, a=16, col="f32"),
            "fp16 operands (1 MMA), fp16 storage": dict(x="h", w="h", a="h", col="f32"),
            "fp16 operands (1 MMA), hi/lo storage": dict(x="h", w="h", a="hh", col="f32"),
            "bf16x3, hi/lo storage": dict(x=16, w=16, a=16, col="f32")}
    for name, pol in pols.items():
        if args.only and not any(tok in name for tok in args.only.split("|")):
            continue
        t = time.time()
        r = step(pol)
        skip = {}      # no analytically-zero bias gradients to exclude here: cosine over every parameter with a gradient
        print("%-44s | %9.2e %9.2e %9.2e | %9.6f %9.6f %9.6f   (%.0f s)" % (
            name, relerr(r["d_real"], ref["d_real"]), relerr(r["fake"], ref["fake"]), relerr(r["d_fake"], ref["d_fake"]),
            cos(r["g_real"], ref["g_real"], skip), cos(r["g_fake"], ref["g_fake"], skip), cos(r["g_gen"], ref["g_gen"], skip),
            time.time() - t), flush=True)


def trace(args):
    """200-step loss traces (loop of main_dcgan.py:68-95 with Adam, as tests/test_gpu_trace.py: width 16, 32x32, batch 32)
    under EMULATED per-pass precision policies against the fp32 oracle, with the fp32 1e-6-perturbation control — the
    protocol of SURVEY.md §7.3: 200-step mean and 50-step moving average within 2 %, early point-wise deviation against
    the control. `mixed` is engine.DcganStep's policy: real pass bf16, D-fake chain fp16 (1 MMA), G step bf16x3."""
    from gan_playground_b200.models import dcgan

    STEPS, B, RES, W = args.steps, 32, 32, 16
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        netG, netD = dcgan.Generator(ngf=W, resolution=RES), dcgan.Discriminator(ndf=W, resolution=RES)
    sd_g = {k: v.clone() for k, v in netG.state_dict().items()}
    sd_d = {k: v.clone() for k, v in netD.state_dict().items()}
    gen = torch.Generator().manual_seed(11)
    xs = torch.rand(STEPS, B, 3, RES, RES, generator=gen) * 2 - 1
    zs = torch.randn(STEPS, 2, B, 100, generator=gen)

    def run_oracle(perturb):
        tr = O.CpuDcganTrainer(sd_g, sd_d)
        return torch.tensor([tr.step(xs[i] + (perturb if i == 0 else 0.0), zs[i, 0], zs[i, 1])[:3] for i in range(STEPS)])

    def run_emulated(p_real, p_fake, p_g):
        pg = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v) for k, v in sd_g.items()}
        pd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v) for k, v in sd_d.items()}
        og = torch.optim.Adam([v for v in pg.values() if v.requires_grad], lr=4e-4, betas=(0.5, 0.999))
        od = torch.optim.Adam([v for v in pd.values() if v.requires_grad], lr=1e-4, betas=(0.5, 0.999))
        loss = lambda p, real, g=False: O.gan_loss("vanilla", p, real, g, 0.9, 0.1, 0.9)
        out = []
        for i in range(STEPS):
            od.zero_grad()
            l1 = loss(discriminator(pd, xs[i], p_real), True)
            l1.backward()
            with torch.no_grad():
                fake = generator(pg, zs[i, 0], p_fake)
            l2 = loss(discriminator(pd, fake, p_fake), False)
            l2.backward()
            od.step()
            og.zero_grad()
            l3 = loss(discriminator(pd, generator(pg, zs[i, 1], p_g), p_g), False, True)
            l3.backward()
            og.step()
            out.append([l1.item(), l2.item(), l3.item()])
        return torch.tensor(out)

    def moving_avg(t, k=min(50, STEPS)):
        c = torch.cumsum(torch.cat([torch.zeros(1, t.shape[1]), t]), 0)
        return (c[k:] - c[:-k]) / k

    rel = lambda a, b: ((a - b).abs() / b.abs().clamp_min(1e-6)).max().item()
    ref, ctl = run_oracle(0.0), run_oracle(1e-6)
    H = dict(x="h", w="h", y="f32", a="h", col="f32")
    runs = {"fp32 control (1e-6 perturbation)": ctl,
            "mixed: real bf16 / fake chain fp16 / G step bf16x3": run_emulated(policy(BF16), policy(BF16, all=H), policy(X3)),
            "bf16x3 on every pass": run_emulated(policy(X3), policy(X3), policy(X3)),
            "bf16 on every pass": run_emulated(policy(BF16), policy(BF16), policy(BF16))}
    print("%d-step loss traces vs the fp32 oracle, max over (lossD_real, lossD_fake, lossG) of the relative deviation" % STEPS)
    print("%-52s %10s %14s %16s %12s" % ("run (emulated roundings)", "mean", "50-step avg", "first 20 steps", "all steps"))
    for name, t in runs.items():
        print("%-52s %9.3f%% %13.3f%% %15.3f%% %11.3f%%" % (
            name, 100 * rel(t.mean(0), ref.mean(0)), 100 * rel(moving_avg(t), moving_avg(ref)),
            100 * rel(t[:20], ref[:20]), 100 * rel(t, ref)), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--width", type=int, default=64)
    ap.add_argument("--only", default="", help="policies whose name contains one of these |-separated substrings")
    ap.add_argument("--sweep", action="store_true", help="per-layer sensitivity: one layer fp16 1-MMA / rest bf16x3 and back")
    ap.add_argument("--trace", action="store_true", help="200-step loss traces under the per-pass policies instead of the one-step study")
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--sndcgan", action="store_true", help="the one-step study for SN-DCGAN 32x32 with the hinge loss (cfg 3)")
    ap.add_argument("--sngan", action="store_true", help="the one-step study for the SNGAN projection pair (ResNet nodes)")
    ap.add_argument("--acgan", action="store_true", help="the one-step study for the main_acgan.py loop and AcganStep's opt-in policy")
    args = ap.parse_args()
    torch.set_num_threads(os.cpu_count())
    if args.trace:
        return trace(args)
    if args.acgan:
        return acgan_study(args)
    if args.sngan:
        return sngan_study(args)
    if args.sndcgan:
        return sndcgan_study(args)
    from gan_playground_b200.models import dcgan

    torch.set_num_threads(os.cpu_count())
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        netG, netD = dcgan.Generator(ngf=args.width), dcgan.Discriminator(ndf=args.width)
    sd_g = {k: v.clone() for k, v in netG.state_dict().items()}
    sd_d = {k: v.clone() for k, v in netD.state_dict().items()}
    gen = torch.Generator().manual_seed(1)
    B = args.batch
    x = torch.rand(B, 3, 64, 64, generator=gen) * 2 - 1
    z1, z2 = torch.randn(B, 100, generator=gen), torch.randn(B, 100, generator=gen)
    ref = O.dcgan_step_grads(sd_g, sd_d, x, z1, z2)
    print("DCGAN-64 width %d batch %d (same seeds as tests/test_gpu_precision.py); measured on the B200: bf16 -> act 6.2e-3 /"
          " 1.3e-2 / 1.3e-2, cos 0.999911 / 0.9938 / 0.9707;  bf16x3 -> act 1.2e-5 / 2.5e-5 / 2.6e-5, cos 1.000000 /"
          " 0.999977 / 0.999893" % (args.width, B))
    print("%-40s %5s | %9s %9s %9s | %9s %9s %9s" % ("policy (emulated)", "MMAs", "D(x)", "G(z)", "D(G(z))", "cos Dreal",
                                                    "cos Dfake", "cos Gstep"))
    for name, P in POLICIES.items():
        if args.only and not any(tok in name for tok in args.only.split("|")):
            continue
        t = time.time()
        r = run(P, sd_g, sd_d, x, z1, z2, ref)
        print("%-40s %5.2f | %9.2e %9.2e %9.2e | %9.6f %9.6f %9.6f   (%.0f s)" % (
            name, mmas(P), r["act D(x)"], r["act G(z)"], r["act D(G(z))"], r["cos D-real"], r["cos D-fake"],
            r["cos G-step"], time.time() - t), flush=True)
    if args.sweep:
        H = dict(x="h", w="h", y="f32", a="h", col="f32")
        for title, make in (("one layer fp16 1-MMA, rest bf16x3", lambda L: policy(X3, **{L: H})),
                            ("one layer bf16x3, rest fp16 1-MMA", lambda L: policy(BF16, all=H, **{L: X3}))):
            print(title + ":")
            for L in LAYERS:
                r = run(make(L), sd_g, sd_d, x, z1, z2, ref)
                print("  %-6s cos D-fake %.6f  G-step %.6f" % (L, r["cos D-fake"], r["cos G-step"]), flush=True)


if __name__ == "__main__":
    main()
