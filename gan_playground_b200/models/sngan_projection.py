"""SNGAN with projection discriminator: API mirror of the reference's models/sngan_projection.py on B200 kernels.

Class names, constructor signatures, sub-module names (l1, block2..5, b6, l6 / block1..5, l6, l_y; c1, c2, c_sc, b1, b2,
bn, embed) and therefore state_dict keys follow upstream exactly; torch layers are fp32 parameter holders created in the
same order (same RNG stream) and wrapped with torch.nn.utils.spectral_norm where upstream does. forward() runs the
NHWC-bf16 tensor-core path (functional_resnet.py); spectral norm runs as gp_sn_* GEMV kernels."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import config
from .. import functional as GF
from .. import functional_resnet as GR
from .. import ops
from ._common import bn_buffers, require_cuda


def _node(fn, h, *args):
    """Apply a node to an activation carrying its companion tensor (`_gp_lo`); re-attach the output's."""
    return GR.attach(fn.apply(h, GR.comp_of(h), *args))


def _conv(h, weight, bias, residual, act, cache, key):
    return GR.attach(GR.Conv2dNHWC.apply(h, GR.comp_of(h), weight, bias, residual, GR.comp_of(residual), act, cache, key))


def _act_code(activation):
    if activation is F.relu or activation is torch.relu:
        return ops.ACT_RELU
    raise NotImplementedError("only activation=F.relu is implemented on the B200 path (reference default)")


def _sn(module, training, table=None):
    """W / sigma of a spectral-normed holder: from the forward's batched table when there is one, else on its own."""
    if table is not None:
        return table[module]
    return GF.SpectralNormFn.apply(module.weight_orig, module.weight_u, module.weight_v, 0, training)


def _holder(layer, gain=1.0, spectral=False):
    """Finish a freshly built torch layer the way every layer of this file is finished upstream: xavier-uniform weight
    (`gain`), zero bias when there is one, optional torch spectral-norm wrapper. The calls happen per layer, in creation
    order — that order is what makes a same-seed construction consume the reference's RNG stream."""
    nn.init.xavier_uniform_(layer.weight, gain=gain)
    if getattr(layer, "bias", None) is not None:
        nn.init.zeros_(layer.bias)
    return nn.utils.spectral_norm(layer) if spectral else layer


_SQRT2 = 2 ** 0.5


def _register(module, layers):
    """Attach (name, layer) pairs in order: sub-module registration order is the state_dict key order."""
    for name, layer in layers:
        setattr(module, name, layer)


class ConditionalBatchNorm2d(nn.Module):
    """Parameter holder (reference: models/sngan_projection.py:6-19): BatchNorm2d(affine=False) + Embedding(n, 2C)
    with gamma ~ N(1, 0.02), beta = 0. The computation is fused into GR.CondBNAct by the owning block."""

    def __init__(self, num_features, num_classes):
        super().__init__()
        C = self.num_features = num_features
        self.bn = nn.BatchNorm2d(C, affine=False)
        self.embed = nn.Embedding(num_classes, 2 * C)          # row = [gamma(C) | beta(C)] of one class
        gamma, beta = self.embed.weight.data.split(C, dim=1)
        gamma.normal_(1, 0.02)
        beta.zero_()

    def forward(self, x, y, act=ops.ACT_NONE, upsample=False):
        """x: NHWC bf16 feature map (internal layout), y: int64 labels."""
        return _node(GR.CondBNAct, x, self.embed.weight, y, bn_buffers(self.bn), act, upsample, self.training)


class ResGenBlock(nn.Module):
    def __init__(self, in_channels, out_channels, hidden_channels=None, ksize=3, pad=1,
                 activation=F.relu, upsample=False, n_classes=0):
        super().__init__()
        hidden = out_channels if hidden_channels is None else hidden_channels
        self.activation, self.upsample, self.n_classes = activation, upsample, n_classes
        self.learnable_sc = upsample or in_channels != out_channels

        def norm(C):
            return ConditionalBatchNorm2d(C, n_classes) if n_classes > 0 else nn.BatchNorm2d(C)

        _register(self, [("c1", _holder(nn.Conv2d(in_channels, hidden, ksize, padding=pad), _SQRT2)),
                         ("c2", _holder(nn.Conv2d(hidden, out_channels, ksize, padding=pad), _SQRT2)),
                         ("b1", norm(in_channels)),
                         ("b2", norm(hidden))])
        if self.learnable_sc:
            self.c_sc = _holder(nn.Conv2d(in_channels, out_channels, 1))
        self._gp_cache = GF.WeightCache()

    def _norm_act(self, bn, h, y, upsample):
        act = _act_code(self.activation)
        if y is not None:
            return bn(h, y, act=act, upsample=upsample)
        h = _node(GR.BNAct, h, bn.weight, bn.bias, bn_buffers(bn), act, self.training)
        return _node(GR.Upsample2x, h) if upsample else h

    def forward(self, x, y=None):
        """x: NHWC bf16. cBN -> ReLU -> nearest x2 -> c1 -> cBN -> ReLU -> c2, plus shortcut c_sc(nearest x2(x))
        (reference :48-66). The 1x1 shortcut conv commutes with nearest upsampling, so it runs on the small grid."""
        h = self._norm_act(self.b1, x, y, self.upsample)
        h = _conv(h, self.c1.weight, self.c1.bias, None, ops.ACT_NONE, self._gp_cache, "c1")
        h = self._norm_act(self.b2, h, y, False)
        if self.learnable_sc:
            sc = _conv(x, self.c_sc.weight, self.c_sc.bias, None, ops.ACT_NONE, self._gp_cache, "c_sc")
            if self.upsample:
                sc = _node(GR.Upsample2x, sc)
        else:
            sc = x
        return _conv(h, self.c2.weight, self.c2.bias, sc, ops.ACT_NONE, self._gp_cache, "c2")


class ResNetGenerator(nn.Module):
    def __init__(self, ch=64, dim_z=128, bottom_width=4, img_dim=3, activation=F.relu, n_classes=0):
        super().__init__()
        self.bottom_width, self.activation, self.dim_z, self.n_classes = bottom_width, activation, dim_z, n_classes
        widths = [ch * m for m in (16, 8, 4, 2, 1)]                 # 16ch at the bottom, halved by every block
        self.l1 = _holder(nn.Linear(dim_z, bottom_width ** 2 * widths[0]))
        for i, (cin, cout) in enumerate(zip(widths, widths[1:]), start=2):
            setattr(self, "block%d" % i, ResGenBlock(cin, cout, activation=activation, upsample=True, n_classes=n_classes))
        self.b6 = nn.BatchNorm2d(ch)
        self.l6 = _holder(nn.Conv2d(ch, img_dim, 3, stride=1, padding=1))
        self._gp_cache = GF.WeightCache()

    def forward(self, z, y):
        require_cuda(z, "sngan_projection.ResNetGenerator")
        with config.resnet_scope():
            h = GF.linear_to_nhwc(z, self.l1.weight, self.l1.bias, self.bottom_width, ops.ACT_NONE, self._gp_cache, "l1")
            if y is not None:
                y = y.contiguous()
            for block in (self.block2, self.block3, self.block4, self.block5):
                h = block(h, y)
            h = _node(GR.BNAct, h, self.b6.weight, self.b6.bias, bn_buffers(self.b6), _act_code(self.activation),
                      self.training)
            return GR.ImageOut3.apply(h, GR.comp_of(h), self.l6.weight, self.l6.bias)


class ResDisBlock(nn.Module):
    def __init__(self, in_channels, out_channels, hidden_channels=None, ksize=3, pad=1,
                 activation=F.relu, downsample=False):
        super().__init__()
        hidden = in_channels if hidden_channels is None else hidden_channels
        self.activation, self.downsample = activation, downsample
        self.learnable_sc = downsample or in_channels != out_channels
        # one layer at a time — create, initialise, wrap — so the RNG draws interleave as upstream's do
        for name, cin, cout in (("c1", in_channels, hidden), ("c2", hidden, out_channels)):
            setattr(self, name, _holder(nn.Conv2d(cin, cout, ksize, padding=pad), _SQRT2, spectral=True))
        if self.learnable_sc:
            self.c_sc = _holder(nn.Conv2d(in_channels, out_channels, 1), spectral=True)
        self._gp_cache = GF.WeightCache()

    def sn_modules(self):
        return [self.c1, self.c2] + ([self.c_sc] if self.learnable_sc else [])

    def forward(self, x, sn=None):
        """x: NHWC bf16. relu -> c1 -> relu -> c2 (-> avgpool) + shortcut c_sc(x) (-> avgpool) (reference :125-136).
        Pooling is linear, so the shortcut is added in c2's epilogue and the sum is pooled once.
        sn: {module: W / sigma} when the owning network ran all its spectral-norm hooks in one batched call."""
        _act_code(self.activation)
        t = self.training
        h = _node(GR.ReluFn, x)
        h = _conv(h, _sn(self.c1, t, sn), self.c1.bias, None, ops.ACT_RELU, self._gp_cache, "c1")
        sc = _conv(x, _sn(self.c_sc, t, sn), self.c_sc.bias, None, ops.ACT_NONE, self._gp_cache, "c_sc") \
            if self.learnable_sc else x
        h = _conv(h, _sn(self.c2, t, sn), self.c2.bias, sc, ops.ACT_NONE, self._gp_cache, "c2")
        return _node(GR.Pool2x, h) if self.downsample else h


class ResDisOptimizedBlock(nn.Module):
    def __init__(self, in_channels, out_channels, ksize=3, pad=1, activation=F.relu):
        super().__init__()
        self.activation = activation
        for name, cin, k, gain in (("c1", in_channels, ksize, _SQRT2), ("c2", out_channels, ksize, _SQRT2),
                                   ("c_sc", in_channels, 1, 1.0)):
            setattr(self, name, _holder(nn.Conv2d(cin, out_channels, k, padding=pad if k == ksize else 0), gain, spectral=True))
        self._gp_cache = GF.WeightCache()

    def sn_modules(self):
        return [self.c1, self.c2, self.c_sc]

    def forward(self, x, sn=None):
        """x: fp32 NCHW image. c1 -> relu -> c2 -> avgpool, plus avgpool(c_sc(x)) (reference :156-164)."""
        _act_code(self.activation)
        t = self.training
        h1, h1c, sc, scc = GR.ImageConv3.apply(x, _sn(self.c1, t, sn), self.c1.bias, _sn(self.c_sc, t, sn), self.c_sc.bias)
        h1, sc = GR.attach((h1, h1c)), GR.attach((sc, scc))
        h = _conv(h1, _sn(self.c2, t, sn), self.c2.bias, sc, ops.ACT_NONE, self._gp_cache, "c2")
        return _node(GR.Pool2x, h)


class SNResNetProjectionDiscriminator(nn.Module):
    def __init__(self, ch=64, n_classes=0, img_dim=3, activation=F.relu):
        super().__init__()
        self.activation = activation
        widths = [ch * m for m in (1, 2, 4, 8, 16)]                 # doubled by every down-sampling block
        self.block1 = ResDisOptimizedBlock(img_dim, widths[0])
        for i, (cin, cout) in enumerate(zip(widths, widths[1:]), start=2):
            setattr(self, "block%d" % i, ResDisBlock(cin, cout, activation=activation, downsample=True))
        self.l6 = _holder(nn.Linear(widths[-1], 1), spectral=True)
        if n_classes > 0:                                           # projection: one spectral-normed embedding row per class
            self.l_y = _holder(nn.Embedding(n_classes, widths[-1]), spectral=True)

    def forward(self, x, y=None):
        require_cuda(x, "sngan_projection.SNResNetProjectionDiscriminator")
        _act_code(self.activation)
        with config.resnet_scope():
            t = self.training
            # every spectral-norm hook this forward fires (l_y only when labels are given, as upstream's hooks), batched
            blocks = (self.block1, self.block2, self.block3, self.block4, self.block5)
            mods = [m for b in blocks for m in b.sn_modules()] + [self.l6] + ([self.l_y] if y is not None else [])
            sn = GF.spectral_norm_all(mods, [0] * len(mods), t) if config.batched_sn() else None
            h = self.block1(x, sn)
            for block in blocks[1:]:
                h = block(h, sn)
            Ey = _sn(self.l_y, t, sn) if (y is not None) else None
            return GR.ProjHead.apply(h, GR.comp_of(h), _sn(self.l6, t, sn), self.l6.bias, Ey,
                                     y.contiguous() if y is not None else None)


if __name__ == "__main__":
    netG = ResNetGenerator(n_classes=10).cuda()
    netD = SNResNetProjectionDiscriminator(n_classes=10).cuda()
    z = torch.randn(4, 128, device="cuda")
    c = torch.randint(10, (4,), device="cuda")
    fake = netG(z, c)
    logits = netD(fake, c)
    print(fake.shape)
    print(logits.shape)
