"""DCGAN generator / discriminator: API mirror of the reference's models/dcgan.py on B200 kernels.

Same constructor arguments, attributes (`arch`, `param_count`, ...), sub-module names and therefore the same
state_dict keys/shapes; the torch layers are kept as fp32 *parameter holders* (so the RNG order
at construction, torch.optim.Adam and torch.save behave as in the reference) while forward() runs the fused
NHWC-bf16 tensor-core path of gan_playground_b200.functional."""
import torch
import torch.nn as nn

from .. import functional as GF
from .. import ops
from ._common import bn_buffers, d_channels, g_channels, init_and_count, require_cuda


def G_arch(ngf=64, img_dim=3):
    return g_channels(ngf)


def D_arch(ndf=64, img_dim=3):
    return d_channels(ndf, img_dim)


class Generator(nn.Module):
    """z -> relu(Linear) -> view(B, C, bw, bw) -> n x [ConvT k4s2p1 + BN + ReLU] -> ConvT k4s2p1 -> Tanh
    (reference: models/dcgan.py:21-57)."""

    def __init__(self, z_dim=100, ngf=64, img_dim=3, resolution=64, bottom_width=4, init='N02', skip_init=False):
        super().__init__()
        self.z_dim, self.ngf, self.img_dim = z_dim, ngf, img_dim
        self.resolution, self.bottom_width, self.init = resolution, bottom_width, init
        self.arch = G_arch(ngf=ngf, img_dim=img_dim)[resolution]  # KeyError on an unknown resolution, as upstream
        cin, cout = self.arch['in_channels'], self.arch['out_channels']
        self.linear = nn.Linear(z_dim, cin[0] * bottom_width ** 2)
        self.blocks = nn.ModuleList(
            nn.Sequential(nn.ConvTranspose2d(i, o, 4, stride=2, padding=1), nn.BatchNorm2d(o), nn.ReLU(True))
            for i, o in zip(cin, cout))
        self.out_layer = nn.Sequential(nn.ConvTranspose2d(cout[-1], img_dim, 4, stride=2, padding=1), nn.Tanh())
        self._gp_cache = GF.WeightCache()
        if not skip_init:
            self.init_weights()

    def init_weights(self):
        init_and_count(self, (nn.ConvTranspose2d, nn.Linear), "G")

    def _trunk(self, h, link=None):
        # a strictly sequential chain: every activation has ONE consumer, so each block's data-gradient GEMM can do the
        # first backward pass of the block before it (functional.BwdLink)
        for i, block in enumerate(self.blocks):
            conv, bn = block[0], block[1]
            nxt = GF.BwdLink()
            h = GF.with_lo(GF.ConvBlock, h, conv.weight, conv.bias, bn.weight, bn.bias, bn_buffers(bn), True,
                           ops.ACT_RELU, self._gp_cache, "blocks.%d" % i, self.training, False, link, nxt)
            link = nxt
        last = self.out_layer[0]
        return GF.with_lo(GF.ImageConvT, h, last.weight, last.bias, ops.ACT_TANH, self._gp_cache, "out_layer", link)

    def forward(self, z):
        require_cuda(z, "dcgan.Generator")
        link = GF.BwdLink()
        h = GF.linear_to_nhwc(z, self.linear.weight, self.linear.bias, self.bottom_width, ops.ACT_RELU,
                              self._gp_cache, "linear", link)
        return self._trunk(h, link)


class Discriminator(nn.Module):
    """x -> n x [Conv k4s2p1 (+BN from block 1) + LeakyReLU 0.2] -> sum over (H, W) -> Linear
    (reference: models/dcgan.py:94-124)."""

    def __init__(self, ndf=64, img_dim=3, resolution=64, output_dim=1, init='N02', skip_init=False):
        super().__init__()
        self.ndf, self.img_dim, self.resolution, self.init = ndf, img_dim, resolution, init
        self.arch = D_arch(ndf=ndf, img_dim=img_dim)[resolution]
        self.blocks = nn.ModuleList()
        for idx, (i, o) in enumerate(zip(self.arch['in_channels'], self.arch['out_channels'])):
            layers = [nn.Conv2d(i, o, 4, stride=2, padding=1)]
            if idx != 0:
                layers.append(nn.BatchNorm2d(o))
            layers.append(nn.LeakyReLU(0.2, True))
            self.blocks.append(nn.Sequential(*layers))
        self.out_layer = nn.Linear(self.arch['out_channels'][-1], output_dim)
        self._gp_cache = GF.WeightCache()
        if not skip_init:
            self.init_weights()

    def init_weights(self):
        init_and_count(self, (nn.Conv2d, nn.Linear), "D")

    def _features(self, x):
        first = self.blocks[0][0]
        link = GF.BwdLink()
        h = GF.image_conv(x, first.weight, first.bias, ops.ACT_LRELU, self._gp_cache, "blocks.0", link)
        for i in range(1, len(self.blocks)):
            conv, bn = self.blocks[i][0], self.blocks[i][1]
            last = i == len(self.blocks) - 1
            nxt = None if last else GF.BwdLink()      # the head is not a GEMM: the last block keeps its own reduction
            h = GF.with_lo(GF.ConvBlock, h, conv.weight, conv.bias, bn.weight, bn.bias, bn_buffers(bn), False,
                           ops.ACT_LRELU, self._gp_cache, "blocks.%d" % i, self.training, last, link, nxt)
            link = nxt
        return h

    def forward(self, x):
        require_cuda(x, "dcgan.Discriminator")
        h = self._features(x)
        return GF.with_lo(GF.Head, h, self.out_layer.weight, self.out_layer.bias, False)
