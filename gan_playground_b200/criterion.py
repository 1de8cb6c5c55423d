"""GANLoss: API mirror of the reference's utils/criterion.py:4-41 with the loss value and its gradient computed by one
fused reduction kernel (gp_gan_loss). Labels are registered buffers exactly as upstream (so .to(device) and
state_dict() behave the same: keys real_label / fake_label / fake_G_label)."""
import torch
import torch.nn as nn

from . import functional as GF
from . import ops
from ._lib import GpError


def _require_cuda(t, what):
    if not t.is_cuda:
        raise GpError("%s on %s — CUDA only, no CPU fallback" % (what, t.device))


class GANLoss(nn.Module):
    def __init__(self, gan_mode, target_real_label=1.0, target_fake_label=0.0, target_fake_G_label=1.0):
        super().__init__()
        self.gan_mode = gan_mode
        self.register_buffer('real_label', torch.tensor(target_real_label))
        self.register_buffer('fake_label', torch.tensor(target_fake_label))
        self.register_buffer('fake_G_label', torch.tensor(target_fake_G_label))
        if gan_mode not in ('vanilla', 'lsgan', 'hinge'):
            raise NotImplementedError('GAN mode %s is not implemented' % gan_mode)
        self.loss = None  # upstream stores an nn loss module here; the fused kernel needs none
        # host copies of the labels: reading the device buffers every call would be a sync in the hot loop
        self._host_labels = (float(target_real_label), float(target_fake_label), float(target_fake_G_label))

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)
        self._host_labels = (float(self.real_label), float(self.fake_label), float(self.fake_G_label))

    def _mode_target(self, is_real, is_generator):
        r, f, g = self._host_labels
        if self.gan_mode in ('vanilla', 'lsgan'):
            target = r if is_real else (g if is_generator else f)
            return (ops.LOSS_BCE if self.gan_mode == 'vanilla' else ops.LOSS_MSE), target
        return (ops.LOSS_HINGE_REAL if is_real else (ops.LOSS_NEG_MEAN if is_generator else ops.LOSS_HINGE_FAKE)), 0.0

    def forward(self, prediction, is_real, is_generator=False):
        _require_cuda(prediction, "GANLoss: prediction")
        mode, target = self._mode_target(is_real, is_generator)
        return GF.GanLossFn.apply(prediction, mode, target)


class ACGANLoss(nn.Module):
    """The objective of the reference's main_acgan.py as one fused kernel (gp_acgan_loss):
    `criterion_adv(outD_adv, is_real, is_generator) + aux_weight * nn.MSELoss()(outD_cls, labels)` (:95-97,114-116,
    129-131) evaluated on the packed two-head logits of `acgan.Discriminator.packed_logits`.

    forward(...) returns a 4-vector [adversarial term, auxiliary term, adv + aux_weight * aux, mean sigmoid(adv)] —
    element 2 is what the script back-propagates, elements 0 / 1 / 3 are the numbers it logs. The script itself keeps
    working unchanged with `GANLoss` + `torch.nn.MSELoss` on `Discriminator.forward`'s pair; this class is the fused
    route `engine.AcganStep` takes."""

    ADV, AUX, TOTAL, SIGMOID_MEAN = 0, 1, 2, 3

    def __init__(self, criterion_adv, aux_weight=0.5):
        super().__init__()
        if not isinstance(criterion_adv, GANLoss):
            raise TypeError("ACGANLoss wraps a GANLoss, got %s" % type(criterion_adv).__name__)
        self.criterion_adv, self.aux_weight = criterion_adv, float(aux_weight)

    def forward(self, packed_logits, labels, is_real, is_generator=False):
        _require_cuda(packed_logits, "ACGANLoss: logits")
        mode, target = self.criterion_adv._mode_target(is_real, is_generator)
        return GF.AcganLossFn.apply(packed_logits, labels, mode, target, self.aux_weight)
