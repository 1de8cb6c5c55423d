"""Shared pieces of the model mirrors: architecture tables, weight initialisation, input checks."""
import torch
import torch.nn as nn

from .. import _lib


def g_channels(ngf):
    """Generator channel plan per output resolution (reference: models/dcgan.py:5-19)."""
    plan = {32: (8, 4, 2), 64: (16, 8, 4, 2), 128: (16, 8, 4, 2, 1)}
    return {res: {"in_channels": [ngf * m for m in ms[:-1]], "out_channels": [ngf * m for m in ms[1:]]}
            for res, ms in plan.items()}


def d_channels(ndf, img_dim):
    """Discriminator channel plan per input resolution (reference: models/dcgan.py:78-92)."""
    plan = {32: (2, 4, 8), 64: (2, 4, 8, 16), 128: (1, 2, 4, 8, 16)}
    return {res: {"in_channels": [img_dim] + [ndf * m for m in ms[:-1]], "out_channels": [ndf * m for m in ms]}
            for res, ms in plan.items()}


_INITS = {
    "ortho": nn.init.orthogonal_,
    "N02": lambda w: nn.init.normal_(w, 0, 0.02),
    "glorot": nn.init.xavier_uniform_,
    "xavier": nn.init.xavier_uniform_,
}


def init_and_count(net, layer_types, tag):
    """Weight init + the reference's `param_count` bookkeeping (reference: models/dcgan.py:59-76,126-143).
    The count walks self.modules(), so containers are counted again — kept because the printed number and the
    attribute are part of the observable API."""
    fn = _INITS.get(net.init)
    total = 0
    for module in net.modules():
        if isinstance(module, layer_types):
            if fn is not None:
                fn(module.weight)
            else:
                print('Init style not recognized...')
        total += sum(p.data.nelement() for p in module.parameters())
    net.param_count = total
    print("Param count for %ss initialized parameters: %d" % (tag, total))


def require_cuda(t, what):
    if not t.is_cuda:
        raise _lib.GpError("%s: input is on %s — this implementation runs only on CUDA (sm_100a); there is no CPU "
                           "fallback. Move the module and its inputs to the GPU." % (what, t.device))


def bn_buffers(bn):
    return (bn.running_mean, bn.running_var, bn.num_batches_tracked)
