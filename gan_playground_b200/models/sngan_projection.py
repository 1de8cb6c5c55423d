"""SNGAN with projection discriminator: API mirror of the reference's models/sngan_projection.py on B200 kernels.

Class names, constructor signatures, sub-module names (l1, block2..5, b6, l6 / block1..5, l6, l_y; c1, c2, c_sc, b1, b2,
bn, embed) and therefore state_dict keys follow upstream exactly; torch layers are fp32 parameter holders created in the
same order (same RNG stream) and wrapped with torch.nn.utils.spectral_norm where upstream does. forward() runs the
NHWC-bf16 tensor-core path (functional_resnet.py); spectral norm runs as gp_sn_* GEMV kernels."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import functional as GF
from .. import functional_resnet as GR
from .. import ops
from ._common import bn_buffers, require_cuda


def _act_code(activation):
    if activation is F.relu or activation is torch.relu:
        return ops.ACT_RELU
    raise NotImplementedError("only activation=F.relu is implemented on the B200 path (reference default)")


def _sn(module, training):
    return GF.SpectralNormFn.apply(module.weight_orig, module.weight_u, module.weight_v, 0, training)


class ConditionalBatchNorm2d(nn.Module):
    """Parameter holder (reference: models/sngan_projection.py:6-19): BatchNorm2d(affine=False) + Embedding(n, 2C)
    with gamma ~ N(1, 0.02), beta = 0. The computation is fused into GR.CondBNAct by the owning block."""

    def __init__(self, num_features, num_classes):
        super().__init__()
        self.num_features = num_features
        self.bn = nn.BatchNorm2d(num_features, affine=False)
        self.embed = nn.Embedding(num_classes, num_features * 2)
        self.embed.weight.data[:, :num_features].normal_(1, 0.02)
        self.embed.weight.data[:, num_features:].zero_()

    def forward(self, x, y, act=ops.ACT_NONE, upsample=False):
        """x: NHWC bf16 feature map (internal layout), y: int64 labels."""
        return GR.CondBNAct.apply(x, self.embed.weight, y, bn_buffers(self.bn), act, upsample, self.training)


class ResGenBlock(nn.Module):
    def __init__(self, in_channels, out_channels, hidden_channels=None, ksize=3, pad=1,
                 activation=F.relu, upsample=False, n_classes=0):
        super().__init__()
        self.activation = activation
        self.upsample = upsample
        self.learnable_sc = in_channels != out_channels or upsample
        hidden_channels = out_channels if hidden_channels is None else hidden_channels
        self.n_classes = n_classes
        self.c1 = nn.Conv2d(in_channels, hidden_channels, ksize, padding=pad)
        nn.init.xavier_uniform_(self.c1.weight, gain=(2 ** 0.5))
        nn.init.zeros_(self.c1.bias)
        self.c2 = nn.Conv2d(hidden_channels, out_channels, ksize, padding=pad)
        nn.init.xavier_uniform_(self.c2.weight, gain=(2 ** 0.5))
        nn.init.zeros_(self.c2.bias)
        if n_classes > 0:
            self.b1 = ConditionalBatchNorm2d(in_channels, n_classes)
            self.b2 = ConditionalBatchNorm2d(hidden_channels, n_classes)
        else:
            self.b1 = nn.BatchNorm2d(in_channels)
            self.b2 = nn.BatchNorm2d(hidden_channels)
        if self.learnable_sc:
            self.c_sc = nn.Conv2d(in_channels, out_channels, 1, padding=0)
            nn.init.xavier_uniform_(self.c_sc.weight)
            nn.init.zeros_(self.c_sc.bias)
        self._gp_cache = GF.WeightCache()

    def _norm_act(self, bn, h, y, upsample):
        act = _act_code(self.activation)
        if y is not None:
            return bn(h, y, act=act, upsample=upsample)
        h = GR.BNAct.apply(h, bn.weight, bn.bias, bn_buffers(bn), act, self.training)
        return GR.Upsample2x.apply(h) if upsample else h

    def forward(self, x, y=None):
        """x: NHWC bf16. cBN -> ReLU -> nearest x2 -> c1 -> cBN -> ReLU -> c2, plus shortcut c_sc(nearest x2(x))
        (reference :48-66). The 1x1 shortcut conv commutes with nearest upsampling, so it runs on the small grid."""
        h = self._norm_act(self.b1, x, y, self.upsample)
        h = GR.Conv2dNHWC.apply(h, self.c1.weight, self.c1.bias, None, ops.ACT_NONE, self._gp_cache, "c1")
        h = self._norm_act(self.b2, h, y, False)
        if self.learnable_sc:
            sc = GR.Conv2dNHWC.apply(x, self.c_sc.weight, self.c_sc.bias, None, ops.ACT_NONE, self._gp_cache, "c_sc")
            if self.upsample:
                sc = GR.Upsample2x.apply(sc)
        else:
            sc = x
        return GR.Conv2dNHWC.apply(h, self.c2.weight, self.c2.bias, sc, ops.ACT_NONE, self._gp_cache, "c2")


class ResNetGenerator(nn.Module):
    def __init__(self, ch=64, dim_z=128, bottom_width=4, img_dim=3, activation=F.relu, n_classes=0):
        super().__init__()
        self.bottom_width = bottom_width
        self.activation = activation
        self.dim_z = dim_z
        self.n_classes = n_classes
        self.l1 = nn.Linear(dim_z, (bottom_width ** 2) * ch * 16)
        nn.init.xavier_uniform_(self.l1.weight)
        nn.init.zeros_(self.l1.bias)
        self.block2 = ResGenBlock(ch * 16, ch * 8, activation=activation, upsample=True, n_classes=n_classes)
        self.block3 = ResGenBlock(ch * 8, ch * 4, activation=activation, upsample=True, n_classes=n_classes)
        self.block4 = ResGenBlock(ch * 4, ch * 2, activation=activation, upsample=True, n_classes=n_classes)
        self.block5 = ResGenBlock(ch * 2, ch, activation=activation, upsample=True, n_classes=n_classes)
        self.b6 = nn.BatchNorm2d(ch)
        self.l6 = nn.Conv2d(ch, img_dim, 3, stride=1, padding=1)
        nn.init.xavier_uniform_(self.l6.weight)
        nn.init.zeros_(self.l6.bias)
        self._gp_cache = GF.WeightCache()

    def forward(self, z, y):
        require_cuda(z, "sngan_projection.ResNetGenerator")
        h = GF.linear_to_nhwc(z, self.l1.weight, self.l1.bias, self.bottom_width, ops.ACT_NONE, self._gp_cache, "l1")
        if y is not None:
            y = y.contiguous()
        for block in (self.block2, self.block3, self.block4, self.block5):
            h = block(h, y)
        h = GR.BNAct.apply(h, self.b6.weight, self.b6.bias, bn_buffers(self.b6), _act_code(self.activation), self.training)
        return GR.ImageOut3.apply(h, self.l6.weight, self.l6.bias)


class ResDisBlock(nn.Module):
    def __init__(self, in_channels, out_channels, hidden_channels=None, ksize=3, pad=1,
                 activation=F.relu, downsample=False):
        super().__init__()
        self.activation = activation
        self.downsample = downsample
        self.learnable_sc = (in_channels != out_channels) or downsample
        hidden_channels = in_channels if hidden_channels is None else hidden_channels
        self.c1 = nn.Conv2d(in_channels, hidden_channels, ksize, padding=pad)
        nn.init.xavier_uniform_(self.c1.weight, gain=(2 ** 0.5))
        nn.init.zeros_(self.c1.bias)
        nn.utils.spectral_norm(self.c1)
        self.c2 = nn.Conv2d(hidden_channels, out_channels, ksize, padding=pad)
        nn.init.xavier_uniform_(self.c2.weight, gain=(2 ** 0.5))
        nn.init.zeros_(self.c2.bias)
        nn.utils.spectral_norm(self.c2)
        if self.learnable_sc:
            self.c_sc = nn.Conv2d(in_channels, out_channels, 1, padding=0)
            nn.init.xavier_uniform_(self.c_sc.weight)
            nn.init.zeros_(self.c_sc.bias)
            nn.utils.spectral_norm(self.c_sc)
        self._gp_cache = GF.WeightCache()

    def forward(self, x):
        """x: NHWC bf16. relu -> c1 -> relu -> c2 (-> avgpool) + shortcut c_sc(x) (-> avgpool) (reference :125-136).
        Pooling is linear, so the shortcut is added in c2's epilogue and the sum is pooled once."""
        _act_code(self.activation)
        t = self.training
        h = GR.ReluFn.apply(x)
        h = GR.Conv2dNHWC.apply(h, _sn(self.c1, t), self.c1.bias, None, ops.ACT_RELU, self._gp_cache, "c1")
        sc = GR.Conv2dNHWC.apply(x, _sn(self.c_sc, t), self.c_sc.bias, None, ops.ACT_NONE, self._gp_cache, "c_sc") \
            if self.learnable_sc else x
        h = GR.Conv2dNHWC.apply(h, _sn(self.c2, t), self.c2.bias, sc, ops.ACT_NONE, self._gp_cache, "c2")
        return GR.Pool2x.apply(h) if self.downsample else h


class ResDisOptimizedBlock(nn.Module):
    def __init__(self, in_channels, out_channels, ksize=3, pad=1, activation=F.relu):
        super().__init__()
        self.activation = activation
        self.c1 = nn.Conv2d(in_channels, out_channels, ksize, padding=pad)
        nn.init.xavier_uniform_(self.c1.weight, gain=(2 ** 0.5))
        nn.init.zeros_(self.c1.bias)
        nn.utils.spectral_norm(self.c1)
        self.c2 = nn.Conv2d(out_channels, out_channels, ksize, padding=pad)
        nn.init.xavier_uniform_(self.c2.weight, gain=(2 ** 0.5))
        nn.init.zeros_(self.c2.bias)
        nn.utils.spectral_norm(self.c2)
        self.c_sc = nn.Conv2d(in_channels, out_channels, 1, padding=0)
        nn.init.xavier_uniform_(self.c_sc.weight)
        nn.init.zeros_(self.c_sc.bias)
        nn.utils.spectral_norm(self.c_sc)
        self._gp_cache = GF.WeightCache()

    def forward(self, x):
        """x: fp32 NCHW image. c1 -> relu -> c2 -> avgpool, plus avgpool(c_sc(x)) (reference :156-164)."""
        _act_code(self.activation)
        t = self.training
        h1, sc = GR.ImageConv3.apply(x, _sn(self.c1, t), self.c1.bias, _sn(self.c_sc, t), self.c_sc.bias)
        h = GR.Conv2dNHWC.apply(h1, _sn(self.c2, t), self.c2.bias, sc, ops.ACT_NONE, self._gp_cache, "c2")
        return GR.Pool2x.apply(h)


class SNResNetProjectionDiscriminator(nn.Module):
    def __init__(self, ch=64, n_classes=0, img_dim=3, activation=F.relu):
        super().__init__()
        self.activation = activation
        self.block1 = ResDisOptimizedBlock(img_dim, ch)
        self.block2 = ResDisBlock(ch, ch * 2, activation=activation, downsample=True)
        self.block3 = ResDisBlock(ch * 2, ch * 4, activation=activation, downsample=True)
        self.block4 = ResDisBlock(ch * 4, ch * 8, activation=activation, downsample=True)
        self.block5 = ResDisBlock(ch * 8, ch * 16, activation=activation, downsample=True)
        self.l6 = nn.Linear(ch * 16, 1)
        nn.init.xavier_uniform_(self.l6.weight)
        nn.init.zeros_(self.l6.bias)
        nn.utils.spectral_norm(self.l6)
        if n_classes > 0:
            self.l_y = nn.Embedding(n_classes, ch * 16)
            nn.init.xavier_uniform_(self.l_y.weight)
            nn.utils.spectral_norm(self.l_y)

    def forward(self, x, y=None):
        require_cuda(x, "sngan_projection.SNResNetProjectionDiscriminator")
        _act_code(self.activation)
        h = self.block1(x)
        for block in (self.block2, self.block3, self.block4, self.block5):
            h = block(h)
        t = self.training
        Ey = _sn(self.l_y, t) if (y is not None) else None
        return GR.ProjHead.apply(h, _sn(self.l6, t), self.l6.bias, Ey, y.contiguous() if y is not None else None)


if __name__ == "__main__":
    netG = ResNetGenerator(n_classes=10).cuda()
    netD = SNResNetProjectionDiscriminator(n_classes=10).cuda()
    z = torch.randn(4, 128, device="cuda")
    c = torch.randint(10, (4,), device="cuda")
    fake = netG(z, c)
    logits = netD(fake, c)
    print(fake.shape)
    print(logits.shape)
