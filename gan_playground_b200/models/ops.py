"""`models.ops` of the reference (models/ops.py): the anti-aliasing BlurPool layers. Only what the hot path uses is
computed natively: BlurPool2d(filt_size=3, pad_type='reflect', stride 1 | 2) — the configuration of models/dcgan_blur.py —
runs as the gp_blur3x3 kernels; the module keeps the reference's constructor, attributes and `filt` buffer (state_dict key
`...filt`, shape (channels, 1, 3, 3)). Other filter sizes / pad types and BlurPool1d are off the hot path and raise."""
import numpy as np
import torch
import torch.nn as nn

from .. import _lib
from .. import functional_resnet as GR

_BINOMIAL = {1: [1.], 2: [1., 1.], 3: [1., 2., 1.], 4: [1., 3., 3., 1.], 5: [1., 4., 6., 4., 1.],
             6: [1., 5., 10., 10., 5., 1.], 7: [1., 6., 15., 20., 15., 6., 1.]}


class BlurPool2d(nn.Module):
    def __init__(self, pad_type='reflect', filt_size=3, stride=2, channels=None, pad_off=0):
        super().__init__()
        self.filt_size, self.pad_off, self.stride, self.channels, self.pad_type = filt_size, pad_off, stride, channels, pad_type
        lo, hi = int(1. * (filt_size - 1) / 2), int(np.ceil(1. * (filt_size - 1) / 2))
        self.pad_sizes = [p + pad_off for p in (lo, hi, lo, hi)]
        self.off = int((stride - 1) / 2.)
        a = torch.tensor(_BINOMIAL[filt_size])
        filt = a[:, None] * a[None, :]
        filt = filt / filt.sum()
        self.register_buffer('filt', filt[None, None, :, :].repeat((self.channels, 1, 1, 1)))

    def native(self):
        return self.filt_size == 3 and self.pad_off == 0 and self.pad_type in ('refl', 'reflect') and self.stride in (1, 2)

    def forward_nhwc(self, h):
        """h: NHWC bf16 (internal layout of the model mirrors)."""
        if not self.native():
            raise _lib.GpError("BlurPool2d(filt_size=%d, pad_type=%r, stride=%d, pad_off=%d) is not on the B200 hot path; "
                               "only filt_size=3 / reflect / stride 1|2 (models/dcgan_blur.py) is implemented"
                               % (self.filt_size, self.pad_type, self.stride, self.pad_off))
        return GR.attach(GR.BlurPool.apply(h, GR.comp_of(h), self.stride))

    def forward(self, inp):
        """Stand-alone use on an NCHW tensor, as in the reference (fp32 in / fp32 out, computed in bf16 NHWC)."""
        if not inp.is_cuda:
            raise _lib.GpError("BlurPool2d: input is on %s — this implementation runs only on CUDA" % (inp.device,))
        if inp.shape[1] % 8 != 0:
            raise _lib.GpError("BlurPool2d: channels must be a multiple of 8 on the B200 path")
        from .. import config

        h = inp.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)
        with config.precision_scope("bf16"):
            return self.forward_nhwc(h).permute(0, 3, 1, 2).to(inp.dtype)


class BlurPool1d(nn.Module):
    def __init__(self, pad_type='reflect', filt_size=3, stride=2, channels=None, pad_off=0):
        super().__init__()
        raise _lib.GpError("BlurPool1d is not used by any network on the B200 hot path (models/ops.py:61-101 upstream)")
