"""Spectral-norm DCGAN: API mirror of the reference's models/dcgan_specnorm.py on B200 kernels.

Every ConvTranspose2d (generator, SN dim=1) and Conv2d (discriminator, SN dim=0) is wrapped with
torch.nn.utils.spectral_norm exactly as upstream (reference: models/dcgan_specnorm.py:37,42,107), so the parameter
holders expose `weight_orig`, `weight_u`, `weight_v` with torch's state_dict layout and RNG draws; the wrapper's
forward pre-hook never runs because forward() below calls the kernels directly: one power iteration + sigma + W/sigma
are gp_sn_* GEMV kernels (functional.SpectralNormFn). The discriminator head is flatten + Linear (:113,125-126)."""
import torch
import torch.nn as nn

from .. import config
from .. import functional as GF
from .. import ops
from ._common import bn_buffers, d_channels, g_channels, init_and_count, require_cuda


def G_arch(ngf=64, img_dim=3):
    return g_channels(ngf)


def D_arch(ndf=64, img_dim=3):
    return d_channels(ndf, img_dim)


def sn_weight(module, dim, training, table=None):
    if table is not None:
        return table[module]
    return GF.SpectralNormFn.apply(module.weight_orig, module.weight_u, module.weight_v, dim, training)


def _sn_table(convs, dim, training):
    """Every spectral-norm hook of one forward in a single batched call (functional.spectral_norm_all)."""
    return GF.spectral_norm_all(convs, [dim] * len(convs), training) if config.batched_sn() else None


class Generator(nn.Module):
    def __init__(self, z_dim=100, ngf=64, img_dim=3, resolution=64, bottom_width=4, init='N02', skip_init=False):
        super().__init__()
        self.z_dim, self.ngf, self.img_dim = z_dim, ngf, img_dim
        self.resolution, self.bottom_width, self.init = resolution, bottom_width, init
        self.arch = G_arch(ngf=ngf, img_dim=img_dim)[resolution]
        cin, cout = self.arch['in_channels'], self.arch['out_channels']
        self.linear = nn.Linear(z_dim, cin[0] * bottom_width ** 2)
        self.blocks = nn.ModuleList(
            nn.Sequential(nn.utils.spectral_norm(nn.ConvTranspose2d(i, o, 4, stride=2, padding=1)),
                          nn.BatchNorm2d(o), nn.ReLU(True))
            for i, o in zip(cin, cout))
        self.out_layer = nn.Sequential(
            nn.utils.spectral_norm(nn.ConvTranspose2d(cout[-1], img_dim, 4, stride=2, padding=1)), nn.Tanh())
        self._gp_cache = GF.WeightCache()
        if not skip_init:
            self.init_weights()

    def init_weights(self):
        init_and_count(self, (nn.ConvTranspose2d, nn.Linear), "G")

    def forward(self, z):
        require_cuda(z, "dcgan_specnorm.Generator")
        sn = _sn_table([b[0] for b in self.blocks] + [self.out_layer[0]], 1, self.training)
        link = GF.BwdLink()      # sequential chain: see models/dcgan.py
        h = GF.linear_to_nhwc(z, self.linear.weight, self.linear.bias, self.bottom_width, ops.ACT_RELU,
                              self._gp_cache, "linear", link)
        for i, block in enumerate(self.blocks):
            conv, bn = block[0], block[1]
            w = sn_weight(conv, 1, self.training, sn)
            nxt = GF.BwdLink()
            h = GF.with_lo(GF.ConvBlock, h, w, conv.bias, bn.weight, bn.bias, bn_buffers(bn), True, ops.ACT_RELU,
                           self._gp_cache, "blocks.%d" % i, self.training, False, link, nxt)
            link = nxt
        last = self.out_layer[0]
        return GF.with_lo(GF.ImageConvT, h, sn_weight(last, 1, self.training, sn), last.bias, ops.ACT_TANH, self._gp_cache,
                          "out_layer", link)


class Discriminator(nn.Module):
    def __init__(self, ndf=64, img_dim=3, resolution=64, bottom_width=4, output_dim=1, init='N02', skip_init=False):
        super().__init__()
        self.ndf, self.img_dim, self.resolution = ndf, img_dim, resolution
        self.bottom_width, self.init = bottom_width, init
        self.arch = D_arch(ndf=ndf, img_dim=img_dim)[resolution]
        self.blocks = nn.ModuleList()
        for idx, (i, o) in enumerate(zip(self.arch['in_channels'], self.arch['out_channels'])):
            layers = [nn.utils.spectral_norm(nn.Conv2d(i, o, 4, stride=2, padding=1))]
            if idx != 0:
                layers.append(nn.BatchNorm2d(o))
            layers.append(nn.LeakyReLU(0.2, True))
            self.blocks.append(nn.Sequential(*layers))
        self.out_layer = nn.Linear(self.arch['out_channels'][-1] * bottom_width ** 2, output_dim)
        self._gp_cache = GF.WeightCache()
        if not skip_init:
            self.init_weights()

    def init_weights(self):
        init_and_count(self, (nn.Conv2d, nn.Linear), "D")

    def forward(self, x, out_hidden=False):
        require_cuda(x, "dcgan_specnorm.Discriminator")
        first = self.blocks[0][0]
        sn = _sn_table([b[0] for b in self.blocks], 0, self.training)
        link = GF.BwdLink()
        h = GF.image_conv(x, sn_weight(first, 0, self.training, sn), first.bias, ops.ACT_LRELU, self._gp_cache, "blocks.0",
                          link)
        for i in range(1, len(self.blocks)):
            conv, bn = self.blocks[i][0], self.blocks[i][1]
            last = i == len(self.blocks) - 1
            nxt = None if last else GF.BwdLink()
            h = GF.with_lo(GF.ConvBlock, h, sn_weight(conv, 0, self.training, sn), conv.bias, bn.weight, bn.bias,
                           bn_buffers(bn), False, ops.ACT_LRELU, self._gp_cache, "blocks.%d" % i, self.training, last,
                           link, nxt)
            link = nxt
        out = GF.with_lo(GF.Head, h, self.out_layer.weight, self.out_layer.bias, True)
        if out_hidden:
            # upstream hands back the NCHW fp32 feature map; this is a layout change at the API boundary only
            return out, h.permute(0, 3, 1, 2).float()
        return out
