"""Spectral-norm DCGAN with an upsample + conv generator: API mirror of the reference's models/dcgan_specnorm_up.py on B200
kernels (a variant that reuses the conv + BN kernels, no new kernel class).

G: relu(Linear) -> view -> n x [nearest x2 -> SN Conv3x3 -> BN -> ReLU] -> nearest x2 -> SN Conv3x3 -> Tanh
   (models/dcgan_specnorm_up.py:36-47,49-58; the convs are nn.Conv2d, so spectral norm runs with dim 0)
D: dcgan_specnorm's discriminator — SN Conv4x4 s2 (+BN from block 1) + LeakyReLU 0.2, flatten head (:111-135); same channel
   tables (:83-97 == models/dcgan.py:78-92).
Same constructor arguments, sub-module indices and state_dict keys (`blocks.i.1.weight_orig/_u/_v`, `blocks.i.2.*` BN,
`out_layer.1.*` in G); torch layers are fp32 parameter holders built in the reference's order (same RNG stream: the
reference's init loop draws N(0, 0.02) into the spectral-norm wrappers' plain `weight` attribute, which leaves
`weight_orig` at torch's default initialisation — reproduced by init_and_count). The 3x3 nodes are the ones of
dcgan_blur / the ResNet pair (functional_resnet.py), so all three forward precision modes of config.py run here."""
import torch.nn as nn

from .. import functional as GF
from .. import functional_resnet as GR
from .. import ops
from . import dcgan_specnorm as _snd
from ._common import bn_buffers, g_channels, init_and_count, require_cuda

D_arch = _snd.D_arch


def G_arch(ngf=64, img_dim=3):
    """models/dcgan_specnorm_up.py:5-19 (the tables of models/dcgan.py)."""
    return g_channels(ngf)


def _node(fn, h, *args):
    return GR.attach(fn.apply(h, GR.comp_of(h), *args))


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
            self.blocks.append(nn.Sequential(nn.Upsample(scale_factor=2),
                                             nn.utils.spectral_norm(nn.Conv2d(i, o, 3, padding=1)),
                                             nn.BatchNorm2d(o), nn.ReLU(True)))
        self.out_layer = nn.Sequential(nn.Upsample(scale_factor=2),
                                       nn.utils.spectral_norm(nn.Conv2d(cout[-1], img_dim, 3, padding=1)), nn.Tanh())
        self._gp_cache = GF.WeightCache()
        if not skip_init:
            self.init_weights()

    def init_weights(self):
        init_and_count(self, (nn.Conv2d, nn.Linear), "G")

    def forward(self, z):
        require_cuda(z, "dcgan_specnorm_up.Generator")
        convs = [b[1] for b in self.blocks] + [self.out_layer[1]]
        sn = _snd._sn_table(convs, 0, self.training)
        h = GF.linear_to_nhwc(z, self.linear.weight, self.linear.bias, self.bottom_width, ops.ACT_RELU,
                              self._gp_cache, "linear")
        for i, block in enumerate(self.blocks):
            conv, bn = block[1], block[2]
            h = _node(GR.Upsample2x, h)
            h = GR.attach(GR.Conv2dNHWC.apply(h, GR.comp_of(h), _snd.sn_weight(conv, 0, self.training, sn), conv.bias,
                                              None, None, ops.ACT_NONE, self._gp_cache, "blocks.%d" % i))
            h = _node(GR.BNAct, h, bn.weight, bn.bias, bn_buffers(bn), ops.ACT_RELU, self.training)
        last = self.out_layer[1]
        h = _node(GR.Upsample2x, h)
        return GR.ImageOut3.apply(h, GR.comp_of(h), _snd.sn_weight(last, 0, self.training, sn), last.bias)


class Discriminator(_snd.Discriminator):
    """models/dcgan_specnorm_up.py:99-135 is dcgan_specnorm's discriminator line for line (constructor, `out_hidden`,
    flatten head, init over Conv2d / Linear)."""

    def forward(self, x, out_hidden=False):
        require_cuda(x, "dcgan_specnorm_up.Discriminator")
        return super().forward(x, out_hidden)
