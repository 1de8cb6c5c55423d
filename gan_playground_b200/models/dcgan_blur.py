"""dcgan_blur generator / discriminator — the networks main_dcgan.py:52-53 actually instantiates — as API mirrors of
the reference's models/dcgan_blur.py on B200 kernels.

G: relu(Linear) -> view -> n x [nearest x2 -> Conv3x3 -> BlurPool(stride 1) -> BN -> LeakyReLU 0.2] -> Conv3x3 -> Tanh
D: n x [Conv3x3 (+BN from block 1) -> LeakyReLU 0.2 (-> BlurPool(stride 2) except after the last block)] -> sum -> Linear
Same constructor arguments, sub-module indices and state_dict keys (`blocks.i.1.*` conv, `blocks.i.2.filt`, `blocks.i.3.*`
BN in G; `blocks.i.0.*`, `blocks.i.1.*`, `blocks.i.{2|3}.filt` in D); torch layers are fp32 parameter holders built in the
reference's order (same RNG stream). All three forward precision modes of config.py run here (functional_resnet.py)."""
import torch.nn as nn

from .. import functional as GF
from .. import functional_resnet as GR
from .. import ops
from ._common import bn_buffers, init_and_count, require_cuda
from .ops import BlurPool2d


def G_arch(ngf=64, img_dim=3):
    """models/dcgan_blur.py:7-21 (one more block than models/dcgan.py: the last conv keeps the resolution)."""
    plan = {32: ([8, 4, 2], [4, 2, 1]), 64: ([16, 8, 4, 2], [8, 4, 2, 1]), 128: ([16, 8, 8, 4, 2], [8, 8, 4, 2, 1])}
    return {r: {'in_channels': [ngf * m for m in i], 'out_channels': [ngf * m for m in o]} for r, (i, o) in plan.items()}


def D_arch(ndf=64, img_dim=3):
    """models/dcgan_blur.py:80-94."""
    plan = {32: ([1, 2, 4], [1, 2, 4, 8]), 64: ([1, 2, 4, 8], [1, 2, 4, 8, 16]), 128: ([1, 2, 4, 8, 8], [1, 2, 4, 8, 8, 16])}
    return {r: {'in_channels': [img_dim] + [ndf * m for m in i], 'out_channels': [ndf * m for m in o]}
            for r, (i, o) in plan.items()}


def _node(fn, h, *args):
    return GR.attach(fn.apply(h, GR.comp_of(h), *args))


def _conv(h, conv, cache, key):
    return GR.attach(GR.Conv2dNHWC.apply(h, GR.comp_of(h), conv.weight, conv.bias, None, None, ops.ACT_NONE, cache, key))


class Generator(nn.Module):
    def __init__(self, z_dim=100, ngf=64, img_dim=3, resolution=64, bottom_width=4, init='N02', skip_init=False):
        super().__init__()
        self.z_dim, self.ngf, self.img_dim = z_dim, ngf, img_dim
        self.resolution, self.bottom_width, self.init = resolution, bottom_width, init
        self.arch = G_arch(ngf=ngf, img_dim=img_dim)[resolution]
        cin, cout = self.arch['in_channels'], self.arch['out_channels']
        self.linear = nn.Linear(z_dim, cin[0] * (bottom_width ** 2))
        self.blocks = nn.ModuleList()
        for i, o in zip(cin, cout):
            self.blocks.append(nn.Sequential(nn.Upsample(scale_factor=2), nn.Conv2d(i, o, 3, stride=1, padding=1),
                                             BlurPool2d(channels=o, stride=1), nn.BatchNorm2d(o), nn.LeakyReLU(0.2, True)))
        self.out_layer = nn.Sequential(nn.Conv2d(cout[-1], img_dim, 3, stride=1, padding=1), nn.Tanh())
        self._gp_cache = GF.WeightCache()
        if not skip_init:
            self.init_weights()

    def init_weights(self):
        # upstream initialises ConvTranspose2d and Linear only (models/dcgan_blur.py:62-78): the 3x3 convs keep torch's
        # default initialisation
        init_and_count(self, (nn.ConvTranspose2d, nn.Linear), "G")

    def forward(self, z):
        require_cuda(z, "dcgan_blur.Generator")
        h = GF.linear_to_nhwc(z, self.linear.weight, self.linear.bias, self.bottom_width, ops.ACT_RELU,
                              self._gp_cache, "linear")
        for i, block in enumerate(self.blocks):
            conv, blur, bn = block[1], block[2], block[3]
            h = _node(GR.Upsample2x, h)
            h = _conv(h, conv, self._gp_cache, "blocks.%d" % i)
            h = blur.forward_nhwc(h)
            h = _node(GR.BNAct, h, bn.weight, bn.bias, bn_buffers(bn), ops.ACT_LRELU, self.training)
        last = self.out_layer[0]
        return GR.ImageOut3.apply(h, GR.comp_of(h), last.weight, last.bias)


class Discriminator(nn.Module):
    def __init__(self, ndf=64, img_dim=3, resolution=64, output_dim=1, init='N02', skip_init=False):
        super().__init__()
        self.ndf, self.img_dim, self.resolution, self.init = ndf, img_dim, resolution, init
        self.arch = D_arch(ndf=ndf, img_dim=img_dim)[resolution]
        n_blocks = len(self.arch['in_channels'])
        self.blocks = nn.ModuleList()
        for idx, (i, o) in enumerate(zip(self.arch['in_channels'], self.arch['out_channels'])):
            block = [nn.Conv2d(i, o, 3, stride=1, padding=1)]
            if idx != 0:
                block.append(nn.BatchNorm2d(o))
            block.append(nn.LeakyReLU(0.2, True))
            if idx < n_blocks - 1:
                block.append(BlurPool2d(channels=o))
            self.blocks.append(nn.Sequential(*block))
        self.out_layer = nn.Linear(self.arch['out_channels'][-1], output_dim)
        self._gp_cache = GF.WeightCache()
        if not skip_init:
            self.init_weights()

    def init_weights(self):
        init_and_count(self, (nn.Conv2d, nn.Linear), "D")

    def forward(self, x):
        require_cuda(x, "dcgan_blur.Discriminator")
        n_blocks = len(self.blocks)
        for idx, block in enumerate(self.blocks):
            conv = block[0]
            if idx == 0:
                h = GR.attach(GR.ImageConv3Act.apply(x, conv.weight, conv.bias, ops.ACT_LRELU, self._gp_cache, "blocks.0"))
            else:
                bn = block[1]
                h = _conv(h, conv, self._gp_cache, "blocks.%d" % idx)
                h = _node(GR.BNAct, h, bn.weight, bn.bias, bn_buffers(bn), ops.ACT_LRELU, self.training)
            if idx < n_blocks - 1:
                h = block[-1].forward_nhwc(h)
        return GF.Head.apply(h, GR.comp_of(h), self.out_layer.weight, self.out_layer.bias, False)
