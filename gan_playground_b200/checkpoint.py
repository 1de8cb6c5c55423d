"""Checkpoint I/O and the sampling path of the reference's scripts for data-parallel runs.

`save_model` writes the dict layout of main_dcgan.py:106-123 / main_sngan.py:111-128 / main_acgan.py (`state_dict` ->
generator / discriminator, `optimizer` -> generator / discriminator, `epoch`; file `checkpoint_%03d.pth`), so a file
written here loads in the unmodified scripts and the other way round (tests/test_checkpoint_compat.py). Under data
parallelism replicas are identical, so ONLY RANK 0 writes and every rank leaves through a barrier; with
optim.FusedAdam(shard=True) the Adam moments live sharded over the ranks and are gathered first.

`sample_images` is the scripts' `netG(fixed_noise).detach()` + `save_image(..., normalize=True)` (main_dcgan.py:101-103):
the generator runs in whatever mode it is in — the scripts sample in train mode, which updates BatchNorm running
statistics, so EVERY rank runs the forward (replicas must stay identical) and rank 0 alone writes the file."""
import os

import torch
import torch.distributed as dist

from . import parallel


def _optimizer_state(opt):
    sd = opt.state_dict()
    if getattr(opt, "shard", False) and parallel.enabled():
        # ZeRO-1: each rank holds the moments of its slice of the flat buffer (zeros elsewhere): sum over ranks
        for e in sd["state"].values():
            for k in ("exp_avg", "exp_avg_sq"):
                dist.all_reduce(e[k], op=dist.ReduceOp.SUM)
    return sd


def save_model(models, optimizers, epoch, checkpoint_path):
    """Same signature and file as the scripts' save_model; returns the path (on every rank)."""
    netG, netD = models
    optG, optD = optimizers
    path = '%s/checkpoint_%03d.pth' % (checkpoint_path, epoch + 1)
    checkpoint = {
        'state_dict': {'generator': netG.state_dict(), 'discriminator': netD.state_dict()},
        'optimizer': {'generator': _optimizer_state(optG), 'discriminator': _optimizer_state(optD)},
        'epoch': epoch,
    }
    if parallel.rank() == 0:
        os.makedirs(checkpoint_path, exist_ok=True)
        torch.save(checkpoint, path)
    if parallel.enabled():
        dist.barrier()
    return path


def sample_images(netG, fixed_noise, path, *labels, nrow=8):
    """`outG = netG(fixed_noise[, fixed_label]).detach(); save_image(outG, path, normalize=True)`; returns outG."""
    with torch.no_grad():
        out = netG(fixed_noise, *labels).detach()
    if parallel.rank() == 0:
        from torchvision.utils import save_image

        os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
        save_image(out.float().cpu(), path, normalize=True, nrow=nrow)
    if parallel.enabled():
        dist.barrier()
    return out
