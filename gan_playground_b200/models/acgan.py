"""ACGAN generator / discriminator: API mirror of the reference's models/acgan.py on B200 kernels.
Differences from dcgan: the generator's Linear consumes cat[z, y] and has NO ReLU (reference: models/acgan.py:32,49-50);
the discriminator has a second Linear head `out_aux` and returns (out, out_aux) (:112-126) — here both heads are one
pass over the features (functional.PackedHeads)."""
import torch
import torch.nn as nn

from .. import functional as GF
from .. import ops
from ._common import d_channels, g_channels, init_and_count, require_cuda
from . import dcgan as _dcgan


def G_arch(ngf=64, img_dim=3):
    return g_channels(ngf)


def D_arch(ndf=64, img_dim=3):
    return d_channels(ndf, img_dim)


class Generator(_dcgan.Generator):
    def __init__(self, z_dim=100, ngf=64, img_dim=3, resolution=64, n_class=10, bottom_width=4, init='N02',
                 skip_init=False):
        nn.Module.__init__(self)
        self.z_dim, self.ngf, self.img_dim = z_dim, ngf, img_dim
        self.resolution, self.bottom_width, self.init = resolution, bottom_width, init
        self.arch = G_arch(ngf=ngf, img_dim=img_dim)[resolution]
        cin, cout = self.arch['in_channels'], self.arch['out_channels']
        self.linear = nn.Linear(z_dim + n_class, cin[0] * bottom_width ** 2)
        self.blocks = nn.ModuleList(
            nn.Sequential(nn.ConvTranspose2d(i, o, 4, stride=2, padding=1), nn.BatchNorm2d(o), nn.ReLU(True))
            for i, o in zip(cin, cout))
        self.out_layer = nn.Sequential(nn.ConvTranspose2d(cout[-1], img_dim, 4, stride=2, padding=1), nn.Tanh())
        self._gp_cache = GF.WeightCache()
        if not skip_init:
            self.init_weights()

    def forward(self, z, y):
        require_cuda(z, "acgan.Generator")
        zy = torch.cat([z, y], 1)   # host-side glue on a (B, z_dim + n_class) tensor, as upstream
        h = GF.linear_to_nhwc(zy, self.linear.weight, self.linear.bias, self.bottom_width, ops.ACT_NONE,
                              self._gp_cache, "linear")
        return self._trunk(h)


class Discriminator(_dcgan.Discriminator):
    def __init__(self, ndf=64, img_dim=3, resolution=64, n_class=10, output_dim=1, init='N02', skip_init=False):
        nn.Module.__init__(self)
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
        self.out_aux = nn.Linear(self.arch['out_channels'][-1], n_class)
        self._gp_cache = GF.WeightCache()
        if not skip_init:
            self.init_weights()

    def packed_logits(self, x):
        """(B, output_dim + n_class) fp32: column block 0 = out_layer, block 1 = out_aux — both heads from one pass over
        the features (functional.PackedHeads). `forward` returns its two column blocks; the fused ACGAN objective
        (criterion.ACGANLoss) consumes it whole."""
        require_cuda(x, "acgan.Discriminator")
        h = self._features(x)
        return GF.with_lo(GF.PackedHeads, h, self.out_layer.weight, self.out_layer.bias, self.out_aux.weight,
                          self.out_aux.bias, False)

    def forward(self, x):
        p = self.packed_logits(x)
        od = self.out_layer.out_features
        return p[:, :od], p[:, od:]
