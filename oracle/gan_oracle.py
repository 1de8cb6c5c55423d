"""ORACLE — test infrastructure only. NOT part of the product path.

A CPU, fp32, *functional* restatement of the reference's one-training-step hot path (hermanprawiro/gan-playground):
networks are pure functions of a reference-layout ``state_dict`` (name -> tensor), evaluated with plain
``torch.nn.functional`` ops on the CPU. Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this file; the product package
(``gan_playground_b200``) never does and fails loudly when its CUDA extension is missing.

The arithmetic of the reference lives in PyTorch (third-party; container-pinned torch 2.11.0); this file restates
the reference's *composition* of those ops, each function citing the reference file:line it follows, and torch's
own ``spectral_norm`` hook (``torch:nn/utils/spectral_norm.py``) as explicit formulas.

Pinning: the reference ships no tests or golden vectors, so the oracle is pinned against the
reference itself: ``oracle/make_golden.py`` imports the unmodified reference from /root/reference, runs it on seeded
inputs and commits the outputs under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks this file against
those fixtures (and against the live reference when /root/reference is present).
"""
import math

import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------------------------------
# Architecture tables
# --------------------------------------------------------------------------------------------------------------


def g_arch(ngf=64):
    """Channel tables of the DCGAN generator. Reference: models/dcgan.py:5-19 (same in dcgan_specnorm.py, acgan.py)."""
    mult = {32: ([8, 4], [4, 2]), 64: ([16, 8, 4], [8, 4, 2]), 128: ([16, 8, 4, 2], [8, 4, 2, 1])}
    return {r: {"in_channels": [ngf * m for m in i], "out_channels": [ngf * m for m in o]} for r, (i, o) in mult.items()}


def d_arch(ndf=64, img_dim=3):
    """Channel tables of the DCGAN discriminator. Reference: models/dcgan.py:78-92."""
    mult = {32: ([2, 4], [2, 4, 8]), 64: ([2, 4, 8], [2, 4, 8, 16]), 128: ([1, 2, 4, 8], [1, 2, 4, 8, 16])}
    return {r: {"in_channels": [img_dim] + [ndf * m for m in i], "out_channels": [ndf * m for m in o]}
            for r, (i, o) in mult.items()}


# --------------------------------------------------------------------------------------------------------------
# Building blocks
# --------------------------------------------------------------------------------------------------------------

BN_EPS = 1e-5
BN_MOMENTUM = 0.1


def batch_norm_train(x, weight, bias, buffers=None, prefix=None):
    """Training-mode BatchNorm2d as torch's nn.BatchNorm2d does it (torch: aten::native_batch_norm):
    normalise with the biased batch variance, update running stats with the unbiased one (momentum 0.1),
    bump num_batches_tracked. `buffers` (a dict) is updated in place when given."""
    mean = x.mean(dim=(0, 2, 3))
    var = x.var(dim=(0, 2, 3), unbiased=False)
    if buffers is not None:
        n = x.numel() // x.shape[1]
        rm, rv = buffers[prefix + "running_mean"], buffers[prefix + "running_var"]
        rm.mul_(1 - BN_MOMENTUM).add_(BN_MOMENTUM * mean.detach())
        rv.mul_(1 - BN_MOMENTUM).add_(BN_MOMENTUM * var.detach() * n / max(n - 1, 1))
        buffers[prefix + "num_batches_tracked"] += 1
    xhat = (x - mean[None, :, None, None]) * torch.rsqrt(var[None, :, None, None] + BN_EPS)
    if weight is not None:
        xhat = xhat * weight[None, :, None, None] + bias[None, :, None, None]
    return xhat


def spectral_norm_weight(sd, prefix, dim, training=True, eps=1e-12, update_buffers=True):
    """torch.nn.utils.spectral_norm's pre-forward hook as formulas (torch:nn/utils/spectral_norm.py:62-114):
    W_mat = W_orig permuted so `dim` is first, flattened; one power iteration under no_grad, in place on u and v
    (training only); sigma = u . (W_mat v) differentiable w.r.t. W_orig; weight = W_orig / sigma.
    `dim` is 0 for Conv2d/Linear/Embedding and 1 for ConvTranspose2d (torch:...spectral_norm.py:322-334)."""
    w = sd[prefix + "weight_orig"]
    u, v = sd[prefix + "weight_u"], sd[prefix + "weight_v"]
    wm = w
    if dim != 0:
        wm = wm.permute(dim, *[d for d in range(wm.dim()) if d != dim])
    wm = wm.reshape(wm.size(0), -1)
    if training:
        with torch.no_grad():
            v_new = F.normalize(torch.mv(wm.t(), u), dim=0, eps=eps)
            u_new = F.normalize(torch.mv(wm, v_new), dim=0, eps=eps)
            if update_buffers:
                v.copy_(v_new)
                u.copy_(u_new)
            u, v = u_new.clone(), v_new.clone()
    sigma = torch.dot(u, torch.mv(wm, v))
    return w / sigma


def _weight(sd, prefix, sn, dim, training):
    if sn:
        return spectral_norm_weight(sd, prefix, dim, training)
    return sd[prefix + "weight"]


# --------------------------------------------------------------------------------------------------------------
# DCGAN family (models/dcgan.py, models/dcgan_specnorm.py, models/acgan.py)
# --------------------------------------------------------------------------------------------------------------


def dcgan_generator(sd, z, y=None, *, sn=False, acgan=False, bottom_width=4, training=True, buffers=None, acts=None):
    """Generator.forward. Reference: models/dcgan.py:49-57 (relu(linear(z)) -> view -> [ConvT4s2p1+BN+ReLU]* ->
    ConvT4s2p1 -> Tanh); models/dcgan_specnorm.py:49-57 (same with spectral-normed ConvT, dim=1);
    models/acgan.py:49-57 (linear(cat[z, y]) with NO ReLU)."""
    if acgan:
        h = F.linear(torch.cat([z, y], 1), sd["linear.weight"], sd["linear.bias"])
    else:
        h = F.relu(F.linear(z, sd["linear.weight"], sd["linear.bias"]))
    h = h.view(h.size(0), -1, bottom_width, bottom_width)
    if acts is not None:
        acts["linear"] = h
    i = 0
    while ("blocks.%d.0.bias" % i) in sd:
        p = "blocks.%d." % i
        h = F.conv_transpose2d(h, _weight(sd, p + "0.", sn, 1, training), sd[p + "0.bias"], stride=2, padding=1)
        if acts is not None:
            acts[p + "0"] = h
        h = F.relu(batch_norm_train(h, sd[p + "1.weight"], sd[p + "1.bias"], buffers, p + "1."))
        if acts is not None:
            acts[p + "2"] = h
        i += 1
    h = F.conv_transpose2d(h, _weight(sd, "out_layer.0.", sn, 1, training), sd["out_layer.0.bias"], stride=2, padding=1)
    return torch.tanh(h)


def dcgan_discriminator(sd, x, *, sn=False, acgan=False, flatten_head=False, training=True, buffers=None, acts=None):
    """Discriminator.forward. Reference: models/dcgan.py:117-124 ([Conv4s2p1 (+BN from block 1) + LeakyReLU(0.2)]*
    -> sum over (H, W) -> Linear); models/dcgan_specnorm.py:120-131 (spectral-normed convs, flatten head);
    models/acgan.py:118-126 (extra out_aux Linear head, returns a pair)."""
    h = x
    i = 0
    while ("blocks.%d.0.bias" % i) in sd:
        p = "blocks.%d." % i
        h = F.conv2d(h, _weight(sd, p + "0.", sn, 0, training), sd[p + "0.bias"], stride=2, padding=1)
        if acts is not None:
            acts[p + "0"] = h
        if (p + "1.weight") in sd:
            h = batch_norm_train(h, sd[p + "1.weight"], sd[p + "1.bias"], buffers, p + "1.")
        h = F.leaky_relu(h, 0.2)
        if acts is not None:
            acts[p + "act"] = h
        i += 1
    h = h.flatten(1) if flatten_head else h.sum(dim=(2, 3))
    out = F.linear(h, sd["out_layer.weight"], sd["out_layer.bias"])
    if acgan:
        return out, F.linear(h, sd["out_aux.weight"], sd["out_aux.bias"])
    return out


def dcgan_up_generator(sd, z, *, bottom_width=4, training=True, buffers=None):
    """Generator.forward of models/dcgan_specnorm_up.py:49-58: relu(linear(z)) -> view -> [Upsample x2 (nearest) ->
    spectral-normed Conv3x3 s1 p1 (dim 0) -> BN -> ReLU]* (blocks at :36-42) -> Upsample x2 -> spectral-normed Conv3x3 ->
    Tanh (:43-47). The discriminator of that file (:111-135) is dcgan_specnorm's: `dcgan_discriminator(sn=True,
    flatten_head=True)`."""
    h = F.relu(F.linear(z, sd["linear.weight"], sd["linear.bias"]))
    h = h.view(h.size(0), -1, bottom_width, bottom_width)
    i = 0
    while ("blocks.%d.1.bias" % i) in sd:
        p = "blocks.%d." % i
        h = F.interpolate(h, scale_factor=2)
        h = F.conv2d(h, spectral_norm_weight(sd, p + "1.", 0, training), sd[p + "1.bias"], stride=1, padding=1)
        h = F.relu(batch_norm_train(h, sd[p + "2.weight"], sd[p + "2.bias"], buffers, p + "2."))
        i += 1
    h = F.interpolate(h, scale_factor=2)
    h = F.conv2d(h, spectral_norm_weight(sd, "out_layer.1.", 0, training), sd["out_layer.1.bias"], stride=1, padding=1)
    return torch.tanh(h)


# --------------------------------------------------------------------------------------------------------------
# dcgan_blur (models/dcgan_blur.py + models/ops.py::BlurPool2d) — what main_dcgan.py:52-53 instantiates
# --------------------------------------------------------------------------------------------------------------


def blurpool2d(x, stride):
    """BlurPool2d(filt_size=3, pad_type='reflect'). Reference: models/ops.py:7-47 — reflection pad 1 on every side,
    then a depth-wise 3x3 convolution with the fixed kernel outer([1,2,1],[1,2,1]) / 16 at the given stride."""
    a = torch.tensor([1.0, 2.0, 1.0], dtype=x.dtype)
    filt = (a[:, None] * a[None, :]) / 16.0
    c = x.shape[1]
    return F.conv2d(F.pad(x, (1, 1, 1, 1), mode="reflect"), filt[None, None].repeat(c, 1, 1, 1), stride=stride, groups=c)


def dcgan_blur_generator(sd, z, *, bottom_width=4, training=True, buffers=None):
    """Generator.forward of models/dcgan_blur.py:53-60: relu(linear(z)) -> view -> [Upsample x2 (nearest) ->
    Conv3x3 s1 p1 -> BlurPool2d(stride 1) -> BN -> LeakyReLU(0.2)]* -> Conv3x3 -> Tanh (blocks at :37-45, out :46-49)."""
    h = F.relu(F.linear(z, sd["linear.weight"], sd["linear.bias"]))
    h = h.view(h.size(0), -1, bottom_width, bottom_width)
    i = 0
    while ("blocks.%d.1.bias" % i) in sd:
        p = "blocks.%d." % i
        h = F.interpolate(h, scale_factor=2)
        h = F.conv2d(h, sd[p + "1.weight"], sd[p + "1.bias"], stride=1, padding=1)
        h = blurpool2d(h, 1)
        h = F.leaky_relu(batch_norm_train(h, sd[p + "3.weight"], sd[p + "3.bias"], buffers, p + "3."), 0.2)
        i += 1
    h = F.conv2d(h, sd["out_layer.0.weight"], sd["out_layer.0.bias"], stride=1, padding=1)
    return torch.tanh(h)


def dcgan_blur_discriminator(sd, x, *, training=True, buffers=None):
    """Discriminator.forward of models/dcgan_blur.py:124-131: [Conv3x3 s1 p1 (+BN from block 1) -> LeakyReLU(0.2) ->
    BlurPool2d(stride 2) except after the last block]* -> sum over (H, W) -> Linear (blocks at :108-118)."""
    h = x
    n_blocks = 0
    while ("blocks.%d.0.bias" % n_blocks) in sd:
        n_blocks += 1
    for i in range(n_blocks):
        p = "blocks.%d." % i
        h = F.conv2d(h, sd[p + "0.weight"], sd[p + "0.bias"], stride=1, padding=1)
        if i != 0:
            h = batch_norm_train(h, sd[p + "1.weight"], sd[p + "1.bias"], buffers, p + "1.")
        h = F.leaky_relu(h, 0.2)
        if i < n_blocks - 1:
            h = blurpool2d(h, 2)
    h = h.sum(dim=(2, 3))
    return F.linear(h, sd["out_layer.weight"], sd["out_layer.bias"])


# --------------------------------------------------------------------------------------------------------------
# SNGAN projection (models/sngan_projection.py)
# --------------------------------------------------------------------------------------------------------------


def _cbn(sd, p, x, y, buffers):
    """ConditionalBatchNorm2d.forward. Reference: models/sngan_projection.py:15-19."""
    c = x.shape[1]
    out = batch_norm_train(x, None, None, buffers, p + "bn.")
    emb = F.embedding(y, sd[p + "embed.weight"])
    gamma, beta = emb[:, :c], emb[:, c:]
    return gamma[:, :, None, None] * out + beta[:, :, None, None]


def _res_gen_block(sd, p, x, y, buffers):
    """ResGenBlock.forward (upsample=True, learnable shortcut). Reference: models/sngan_projection.py:48-66."""
    h = F.relu(_cbn(sd, p + "b1.", x, y, buffers))
    h = F.interpolate(h, scale_factor=2)
    h = F.conv2d(h, sd[p + "c1.weight"], sd[p + "c1.bias"], padding=1)
    h = F.relu(_cbn(sd, p + "b2.", h, y, buffers))
    h = F.conv2d(h, sd[p + "c2.weight"], sd[p + "c2.bias"], padding=1)
    sc = F.conv2d(F.interpolate(x, scale_factor=2), sd[p + "c_sc.weight"], sd[p + "c_sc.bias"])
    return h + sc


def sngan_generator(sd, z, y, *, bottom_width=4, buffers=None):
    """ResNetGenerator.forward. Reference: models/sngan_projection.py:83-97."""
    h = F.linear(z, sd["l1.weight"], sd["l1.bias"])
    h = h.reshape(h.shape[0], -1, bottom_width, bottom_width)
    for b in ("block2.", "block3.", "block4.", "block5."):
        h = _res_gen_block(sd, b, h, y, buffers)
    h = F.relu(batch_norm_train(h, sd["b6.weight"], sd["b6.bias"], buffers, "b6."))
    return torch.tanh(F.conv2d(h, sd["l6.weight"], sd["l6.bias"], padding=1))


def _sn_conv(sd, p, x, padding, training):
    return F.conv2d(x, spectral_norm_weight(sd, p, 0, training), sd[p + "bias"], padding=padding)


def sngan_discriminator(sd, x, y=None, *, training=True):
    """SNResNetProjectionDiscriminator.forward. Reference: models/sngan_projection.py:183-196, blocks :125-136
    (ResDisBlock) and :156-164 (ResDisOptimizedBlock). Every conv / l6 / l_y is spectral-normed (dim 0)."""
    # block1 (optimized block): c1 -> relu -> c2 -> avgpool ; shortcut c_sc(x) -> avgpool
    h = _sn_conv(sd, "block1.c1.", x, 1, training)
    h = _sn_conv(sd, "block1.c2.", F.relu(h), 1, training)
    h = F.avg_pool2d(h, 2)
    sc = F.avg_pool2d(_sn_conv(sd, "block1.c_sc.", x, 0, training), 2)
    h = h + sc
    for b in ("block2.", "block3.", "block4.", "block5."):
        xin = h
        t = _sn_conv(sd, b + "c1.", F.relu(xin), 1, training)
        t = _sn_conv(sd, b + "c2.", F.relu(t), 1, training)
        t = F.avg_pool2d(t, 2)
        sc = F.avg_pool2d(_sn_conv(sd, b + "c_sc.", xin, 0, training), 2)
        h = t + sc
    h = F.relu(h).sum(dim=(2, 3))
    out = F.linear(h, spectral_norm_weight(sd, "l6.", 0, training), sd["l6.bias"])
    if y is not None:
        w_y = F.embedding(y, spectral_norm_weight(sd, "l_y.", 0, training))
        out = out + (w_y * h).sum(dim=1, keepdim=True)
    return out


# --------------------------------------------------------------------------------------------------------------
# Losses (utils/criterion.py)
# --------------------------------------------------------------------------------------------------------------


def gan_loss(mode, pred, is_real, is_generator=False, real_label=1.0, fake_label=0.0, fake_g_label=1.0):
    """GANLoss.forward. Reference: utils/criterion.py:21-41 — vanilla = BCE-with-logits against a constant soft label
    (mean), lsgan = MSE, hinge = relu(1 -/+ p).mean() for D and -p.mean() for G."""
    if mode in ("vanilla", "lsgan"):
        t = real_label if is_real else (fake_g_label if is_generator else fake_label)
        target = torch.full_like(pred, float(t))
        if mode == "vanilla":
            return F.binary_cross_entropy_with_logits(pred, target)
        return F.mse_loss(pred, target)
    if mode == "hinge":
        if is_real:
            return F.relu(1.0 - pred).mean()
        if is_generator:
            return -pred.mean()
        return F.relu(1.0 + pred).mean()
    raise NotImplementedError("GAN mode %s is not implemented" % mode)


# --------------------------------------------------------------------------------------------------------------
# One training step (main_dcgan.py:68-95) as a pure function of parameter dicts — used for gradients / loss traces
# --------------------------------------------------------------------------------------------------------------


def split_state(sd):
    """Split a state_dict into differentiable leaves (float params) and buffers (BN stats, SN u/v)."""
    params, buffers = {}, {}
    for k, v in sd.items():
        is_buf = k.endswith(("running_mean", "running_var", "num_batches_tracked", "weight_u", "weight_v"))
        (buffers if is_buf else params)[k] = v
    return params, buffers


def dcgan_step_grads(sd_g, sd_d, x_real, z1, z2, labels=(0.9, 0.1, 0.9), mode="vanilla", **net_kw):
    """Gradients of the three backward passes of one step, following main_dcgan.py:68-95:
    D-real backward, D-fake backward (on G(z1).detach()), G-step backward through D (on G(z2)).
    Returns dict with losses, D grads (accumulated real+fake), G grads, and the logits."""
    rl, fl, gl = labels
    g_kw = {k: v for k, v in net_kw.items() if k in ("sn", "bottom_width")}
    d_kw = {k: v for k, v in net_kw.items() if k in ("sn", "flatten_head")}
    if net_kw.get("up"):     # models/dcgan_specnorm_up.py: upsample + SN conv3x3 generator, dcgan_specnorm's discriminator
        g_kw = {k: v for k, v in net_kw.items() if k == "bottom_width"}
        return _step_grads(dcgan_up_generator, dcgan_discriminator, g_kw, dict(sn=True, flatten_head=True), sd_g, sd_d,
                           x_real, z1, z2, labels, mode)
    if net_kw.get("blur"):   # models/dcgan_blur.py instead of models/dcgan.py
        g_kw = {k: v for k, v in net_kw.items() if k == "bottom_width"}
        d_kw = {}
        return _step_grads(dcgan_blur_generator, dcgan_blur_discriminator, g_kw, d_kw, sd_g, sd_d, x_real, z1, z2,
                           labels, mode)
    return _step_grads(dcgan_generator, dcgan_discriminator, g_kw, d_kw, sd_g, sd_d, x_real, z1, z2, labels, mode)


def _step_grads(dcgan_generator, dcgan_discriminator, g_kw, d_kw, sd_g, sd_d, x_real, z1, z2, labels, mode):
    rl, fl, gl = labels
    pg = {k: (v.detach().clone().requires_grad_(True) if v.is_floating_point() else v.clone()) for k, v in sd_g.items()}
    pd = {k: (v.detach().clone().requires_grad_(True) if v.is_floating_point() else v.clone()) for k, v in sd_d.items()}
    out = {}
    d_real = dcgan_discriminator(pd, x_real, **d_kw)
    loss_real = gan_loss(mode, d_real, True, False, rl, fl, gl)
    fake1 = dcgan_generator(pg, z1, **g_kw).detach()
    d_fake = dcgan_discriminator(pd, fake1, **d_kw)
    loss_fake = gan_loss(mode, d_fake, False, False, rl, fl, gl)
    d_leaves = [k for k, v in pd.items() if v.requires_grad]
    gr = torch.autograd.grad(loss_real, [pd[k] for k in d_leaves], allow_unused=True)
    gf = torch.autograd.grad(loss_fake, [pd[k] for k in d_leaves], allow_unused=True)
    out["d_grads_real"] = {k: g for k, g in zip(d_leaves, gr) if g is not None}
    out["d_grads_fake"] = {k: g for k, g in zip(d_leaves, gf) if g is not None}
    fake2 = dcgan_generator(pg, z2, **g_kw)
    d_g = dcgan_discriminator(pd, fake2, **d_kw)
    loss_g = gan_loss(mode, d_g, False, True, rl, fl, gl)
    g_leaves = [k for k, v in pg.items() if v.requires_grad]
    gg = torch.autograd.grad(loss_g, [pg[k] for k in g_leaves], allow_unused=True)
    out["g_grads"] = {k: g for k, g in zip(g_leaves, gg) if g is not None}
    out.update(loss_real=loss_real.detach(), loss_fake=loss_fake.detach(), loss_g=loss_g.detach(),
               d_real=d_real.detach(), d_fake=d_fake.detach(), d_g=d_g.detach(), fake1=fake1, fake2=fake2.detach())
    return out


def _leaves(sd):
    return {k: (v.detach().clone().requires_grad_(True) if v.is_floating_point() and not k.endswith(
        ("running_mean", "running_var", "weight_u", "weight_v")) else v.clone()) for k, v in sd.items()}


def _grads(loss, p):
    keys = [k for k, v in p.items() if v.requires_grad]
    gs = torch.autograd.grad(loss, [p[k] for k in keys], allow_unused=True, retain_graph=True)
    return {k: g for k, g in zip(keys, gs) if g is not None}


def sngan_step_grads(sd_g, sd_d, x_real, y_real, z, c, bottom_width=4):
    """The three backward passes of one main_sngan.py iteration (:72-99) at FIXED weights (no optimiser step in
    between): D-real (:73-77), D-fake on G(z, c).detach() (:82-87), G step through D on the SAME fake batch, re-using the
    generator graph of :82 (:94-98). Hinge losses (:57). Spectral-norm u / v advance once per discriminator forward, as
    in the script. Returns losses, logits, the fake batch and the per-pass gradients."""
    pg, pd = _leaves(sd_g), _leaves(sd_d)
    d_real = sngan_discriminator(pd, x_real, y_real)
    loss_real = gan_loss("hinge", d_real, True)
    g_real = _grads(loss_real, pd)
    fake = sngan_generator(pg, z, c, bottom_width=bottom_width)
    d_fake = sngan_discriminator(pd, fake.detach(), c)
    loss_fake = gan_loss("hinge", d_fake, False)
    g_fake = _grads(loss_fake, pd)
    d_g = sngan_discriminator(pd, fake, c)
    loss_g = gan_loss("hinge", d_g, False, True)
    g_g = _grads(loss_g, pg)
    return dict(loss_real=loss_real.detach(), loss_fake=loss_fake.detach(), loss_g=loss_g.detach(), d_real=d_real.detach(),
                d_fake=d_fake.detach(), d_g=d_g.detach(), fake=fake.detach(), d_grads_real=g_real, d_grads_fake=g_fake,
                g_grads=g_g)


def acgan_step_grads(sd_g, sd_d, x_real, y_real, z, labels=(0.9, 0.1, 0.9), aux_weight=0.5):
    """The three backward passes of one main_acgan.py iteration (:90-131) at FIXED weights: objective
    criterion_adv + 0.5 * MSELoss(aux head, labels) on the real batch (:95-97), on G(z, y).detach() (:107-116) and, for
    the generator, on the same fake batch through D (:123-131). Returns losses (the total each pass back-propagates),
    both heads' outputs, the fake batch and the per-pass gradients."""
    rl, fl, gl = labels
    pg, pd = _leaves(sd_g), _leaves(sd_d)

    def objective(x, is_real, is_gen):
        adv, cls = dcgan_discriminator(pd, x, acgan=True)
        return gan_loss("vanilla", adv, is_real, is_gen, rl, fl, gl) + aux_weight * F.mse_loss(cls, y_real), adv, cls

    loss_real, adv_r, cls_r = objective(x_real, True, False)
    g_real = _grads(loss_real, pd)
    fake = dcgan_generator(pg, z, y_real, acgan=True)
    loss_fake, adv_f, cls_f = objective(fake.detach(), False, False)
    g_fake = _grads(loss_fake, pd)
    loss_g, adv_g, _ = objective(fake, False, True)
    g_g = _grads(loss_g, pg)
    return dict(loss_real=loss_real.detach(), loss_fake=loss_fake.detach(), loss_g=loss_g.detach(), d_real=adv_r.detach(),
                d_real_cls=cls_r.detach(), d_fake=adv_f.detach(), d_fake_cls=cls_f.detach(), d_g=adv_g.detach(),
                fake=fake.detach(), d_grads_real=g_real, d_grads_fake=g_fake, g_grads=g_g)


def step_flops_dcgan64(batch):
    """Minimal algorithmic FLOPs of one DCGAN-64 step: 9.7994 GFLOP per image."""
    return 9.7994e9 * batch


# --------------------------------------------------------------------------------------------------------------
# CPU training loop of the reference (used by bench.py's cpu_baseline / --impl reference legs and by trace tests)
# --------------------------------------------------------------------------------------------------------------


def init_dcgan_state(seed=0, ngf=64, ndf=64, res=64, z_dim=100, img_dim=3, bottom_width=4):
    """Fresh reference-layout state dicts with the reference's 'N02' initialisation (models/dcgan.py:61-69,128-136):
    N(0, 0.02) on ConvT/Conv/Linear weights, torch defaults (here: zeros) for biases, BN weight 1 / bias 0.
    (Biases use zeros instead of torch's uniform default: irrelevant for timing, and tests load explicit weights.)"""
    g = torch.Generator().manual_seed(seed)

    def n02(*shape):
        return torch.randn(*shape, generator=g) * 0.02

    ga, da = g_arch(ngf)[res], d_arch(ndf, img_dim)[res]
    sd_g = {"linear.weight": n02(ga["in_channels"][0] * bottom_width ** 2, z_dim),
            "linear.bias": torch.zeros(ga["in_channels"][0] * bottom_width ** 2)}
    for i, (ci, co) in enumerate(zip(ga["in_channels"], ga["out_channels"])):
        p = "blocks.%d." % i
        sd_g.update({p + "0.weight": n02(ci, co, 4, 4), p + "0.bias": torch.zeros(co), p + "1.weight": torch.ones(co),
                     p + "1.bias": torch.zeros(co), p + "1.running_mean": torch.zeros(co),
                     p + "1.running_var": torch.ones(co), p + "1.num_batches_tracked": torch.tensor(0)})
    sd_g.update({"out_layer.0.weight": n02(ga["out_channels"][-1], img_dim, 4, 4), "out_layer.0.bias": torch.zeros(img_dim)})
    sd_d = {}
    for i, (ci, co) in enumerate(zip(da["in_channels"], da["out_channels"])):
        p = "blocks.%d." % i
        sd_d.update({p + "0.weight": n02(co, ci, 4, 4), p + "0.bias": torch.zeros(co)})
        if i != 0:
            sd_d.update({p + "1.weight": torch.ones(co), p + "1.bias": torch.zeros(co), p + "1.running_mean": torch.zeros(co),
                         p + "1.running_var": torch.ones(co), p + "1.num_batches_tracked": torch.tensor(0)})
    sd_d.update({"out_layer.weight": n02(1, da["out_channels"][-1]), "out_layer.bias": torch.zeros(1)})
    return sd_g, sd_d


class CpuDcganTrainer:
    """The loop body of main_dcgan.py:68-95 on the CPU in fp32: Adam(lr 4e-4 / 1e-4, betas (0.5, 0.999)) (:55-56),
    GANLoss('vanilla', 0.9, 0.1, 0.9) (:58), D-real / D-fake backward accumulate, G step through D."""

    def __init__(self, sd_g, sd_d, labels=(0.9, 0.1, 0.9), mode="vanilla", lr_g=4e-4, lr_d=1e-4, betas=(0.5, 0.999)):
        self.pg, self.bg = split_state({k: v.clone() for k, v in sd_g.items()})
        self.pd, self.bd = split_state({k: v.clone() for k, v in sd_d.items()})
        for d in (self.pg, self.pd):
            for v in d.values():
                v.requires_grad_(True)
        self.opt_g = torch.optim.Adam(list(self.pg.values()), lr=lr_g, betas=betas)
        self.opt_d = torch.optim.Adam(list(self.pd.values()), lr=lr_d, betas=betas)
        self.labels, self.mode = labels, mode

    def _g(self, z):
        return dcgan_generator({**self.pg, **self.bg}, z, buffers=self.bg)

    def _d(self, x):
        return dcgan_discriminator({**self.pd, **self.bd}, x, buffers=self.bd)

    def step(self, x, z1, z2):
        rl, fl, gl = self.labels
        self.opt_d.zero_grad()
        out = self._d(x)
        dx = out.mean().item()
        l_real = gan_loss(self.mode, out, True, False, rl, fl, gl)
        l_real.backward()
        fake = self._g(z1)
        out = self._d(fake.detach())
        dgz1 = out.mean().item()
        l_fake = gan_loss(self.mode, out, False, False, rl, fl, gl)
        l_fake.backward()
        self.opt_d.step()
        self.opt_g.zero_grad()
        fake = self._g(z2)
        out = self._d(fake)
        dgz2 = out.mean().item()
        l_g = gan_loss(self.mode, out, False, True, rl, fl, gl)
        l_g.backward()
        self.opt_g.step()
        return l_real.item(), l_fake.item(), l_g.item(), dx, dgz1, dgz2


class _CpuTrainer:
    """Shared plumbing of the CPU loop restatements: parameters become leaves, buffers are updated in place."""

    def __init__(self, sd_g, sd_d, lr_g, lr_d, betas):
        self.pg, self.bg = split_state({k: v.clone() for k, v in sd_g.items()})
        self.pd, self.bd = split_state({k: v.clone() for k, v in sd_d.items()})
        for d in (self.pg, self.pd):
            for v in d.values():
                v.requires_grad_(True)
        self.opt_g = torch.optim.Adam(list(self.pg.values()), lr=lr_g, betas=betas)
        self.opt_d = torch.optim.Adam(list(self.pd.values()), lr=lr_d, betas=betas)

    def state_dicts(self):
        """(generator, discriminator) state in the reference's state_dict layout (detached copies)."""
        return ({k: v.detach().clone() for k, v in {**self.pg, **self.bg}.items()},
                {k: v.detach().clone() for k, v in {**self.pd, **self.bd}.items()})


class CpuSnganTrainer(_CpuTrainer):
    """The loop body of main_sngan.py:65-100 on the CPU in fp32: Adam(lr 2e-4, betas (0, 0.999)) on both nets (:54-55,
    argparse defaults :20-21), GANLoss('hinge') (:57); ONE generator forward per iteration — the G step (:92-99) re-uses
    the graph of the fake batch the discriminator was just trained on and runs only when `i % n_disc_update == 0`."""

    def __init__(self, sd_g, sd_d, n_disc_update=5, bottom_width=4, lr=2e-4, betas=(0.0, 0.999)):
        super().__init__(sd_g, sd_d, lr, lr, betas)
        self.n_disc_update, self.bottom_width = n_disc_update, bottom_width
        self.i = 0

    def _d(self, x, y):
        sd = {**self.pd, **self.bd}
        return sngan_discriminator(sd, x, y)          # u / v buffers advance in place (shared tensors)

    def step(self, x, y, z, c):
        """Returns (lossD_real, lossD_fake, lossG or None, D(x), D(G(z))_1, D(G(z))_2 or None)."""
        self.opt_d.zero_grad()
        out = self._d(x, y)
        dx = out.mean().item()
        l_real = gan_loss("hinge", out, True)
        l_real.backward()
        fake = sngan_generator({**self.pg, **self.bg}, z, c, bottom_width=self.bottom_width, buffers=self.bg)
        out = self._d(fake.detach(), c)
        dgz1 = out.mean().item()
        l_fake = gan_loss("hinge", out, False)
        l_fake.backward()
        self.opt_d.step()
        l_g = dgz2 = None
        if self.i % self.n_disc_update == 0:
            self.opt_g.zero_grad()
            out = self._d(fake, c)
            dgz2 = out.mean().item()
            l_g = gan_loss("hinge", out, False, True)
            l_g.backward()
            self.opt_g.step()
            l_g = l_g.item()
        self.i += 1
        return l_real.item(), l_fake.item(), l_g, dx, dgz1, dgz2


class CpuAcganTrainer(_CpuTrainer):
    """The loop body of main_acgan.py:84-133 on the CPU in fp32: Adam(lr 4e-4 / 1e-4, betas (0.5, 0.999)) (:59-60),
    adversarial GANLoss('vanilla', 0.9, 0.1, 0.9) (:62) plus 0.5 x MSELoss between the auxiliary head and the float
    attribute vector (:64,95-97,114-116,129-131). The fake batch is conditioned on the real batch's labels (:106-107);
    one generator forward per iteration (the G step re-uses outG, :123); D(x) / D(G(z)) are sigmoid means (:94,112,127)."""

    def __init__(self, sd_g, sd_d, labels=(0.9, 0.1, 0.9), lr_g=4e-4, lr_d=1e-4, betas=(0.5, 0.999), aux_weight=0.5):
        super().__init__(sd_g, sd_d, lr_g, lr_d, betas)
        self.labels, self.aux_weight = labels, aux_weight

    def _d(self, x):
        return dcgan_discriminator({**self.pd, **self.bd}, x, acgan=True, buffers=self.bd)

    def step(self, x, y, z):
        """Returns (lossD_adv, lossD_aux, lossG_adv, lossG_aux, D(x), D(G(z))_1, D(G(z))_2) — the seven numbers of the
        reference's progress line (:136-137)."""
        rl, fl, gl = self.labels
        self.opt_d.zero_grad()
        adv, cls = self._d(x)
        dx = torch.sigmoid(adv).mean().item()
        l_real_adv, l_real_aux = gan_loss("vanilla", adv, True, False, rl, fl, gl), F.mse_loss(cls, y)
        (l_real_adv + l_real_aux * self.aux_weight).backward()
        fake = dcgan_generator({**self.pg, **self.bg}, z, y, acgan=True, buffers=self.bg)
        adv, cls = self._d(fake.detach())
        dgz1 = torch.sigmoid(adv).mean().item()
        l_fake_adv, l_fake_aux = gan_loss("vanilla", adv, False, False, rl, fl, gl), F.mse_loss(cls, y)
        (l_fake_adv + l_fake_aux * self.aux_weight).backward()
        self.opt_d.step()
        self.opt_g.zero_grad()
        adv, cls = self._d(fake)
        dgz2 = torch.sigmoid(adv).mean().item()
        l_g_adv, l_g_aux = gan_loss("vanilla", adv, False, True, rl, fl, gl), F.mse_loss(cls, y)
        (l_g_adv + l_g_aux * self.aux_weight).backward()
        self.opt_g.step()
        return ((l_real_adv + l_fake_adv).item(), (l_real_aux + l_fake_aux).item(), l_g_adv.item(), l_g_aux.item(),
                dx, dgz1, dgz2)
