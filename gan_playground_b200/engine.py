"""Step driver for the adversarial loop of the reference's main_dcgan.py:68-95.

`DcganStep` runs exactly that loop body (D-real backward, D-fake backward on G(z).detach(), optD.step, G step through D,
optG.step) either eagerly — one launch per kernel, three host reads per step like the reference's `.item()` calls — or
as ONE CUDA graph replay per step: the whole step (forward, autograd backward, NCCL all-reduces, Adam) is captured on a
side stream after a warm-up, the six logged scalars (losses, D(x), D(G(z))) are written to a device buffer inside the
graph and read back once per step. Graph replay removes the per-launch host overhead (~250 launches per step), which
is what bounds small per-GPU batches (strong scaling at 128 images/GPU).

Requirements for graph mode: torch optimisers built with `capturable=True` (FusedAdam is capturable as is); fixed batch
size. The weight-staging caches of the networks are cleared before capture so the staging kernels are part of the graph,
and the warm-up steps capture needs are undone afterwards (parameters, BatchNorm buffers and optimiser state are restored
in place), so the first replay is the first training step."""
import torch

from . import config, ops, parallel
from .optim import FusedAdam


def _clear_caches(*nets):
    for net in nets:
        for m in net.modules():
            c = getattr(m, "_gp_cache", None)
            if c is not None:
                c.clear()


class DcganStep:
    def __init__(self, netG, netD, criterion, optG, optD, batch, z_dim, device, use_graph=False, warmup=3, overlap=True,
                 mixed_precision=True):
        self.netG, self.netD, self.crit, self.optG, self.optD = netG, netD, criterion, optG, optD
        self.batch, self.z_dim, self.dev = batch, z_dim, device
        # FusedAdam owns flat parameter / gradient buffers and does the gradient collective itself; any other optimiser
        # (torch.optim.Adam as in the reference scripts) gets a flat gradient bucket per network for the all-reduce
        self.bucketD = None if isinstance(optD, FusedAdam) else parallel.GradBucket(netD)
        self.bucketG = None if isinstance(optG, FusedAdam) else parallel.GradBucket(netG)
        self.use_graph = use_graph
        self.graph = None
        self.x_static = torch.zeros(batch, netD.img_dim, netD.resolution, netD.resolution, device=device)
        self.z_static = torch.zeros(2, batch, z_dim, device=device)
        self.scalars = torch.zeros(6, device=device)   # lossD_real, lossD_fake, lossG, D(x), D(G(z))1, D(G(z))2
        self.fixed_z = False
        self._warmup = warmup
        self.overlap = overlap
        self.arena = ops.ZeroArena(device)
        self._d_params = [p for p in netD.parameters() if p.requires_grad]
        # mixed forward precision: the real-image D pass needs no 3-MMA forward (config.precision_scope)
        self.real_precision = "bf16" if (mixed_precision and config.x3()) else None
        self._side = torch.cuda.Stream(device=device) if device.type == "cuda" else None

    # ---- the loop body; `log(i, t)` receives the six scalars as 0-dim device tensors
    def _body(self, inputs, z1, z2, log):
        netG, netD, crit = self.netG, self.netD, self.crit
        self.arena.begin()                                       # one memset for every zero-initialised workspace
        ops.ZeroArena.active = self.arena
        try:
            self._body_steps(netG, netD, crit, inputs, z1, z2, log)
        finally:
            ops.ZeroArena.active = None

    def _body_steps(self, netG, netD, crit, inputs, z1, z2, log):
        self._zero(self.optD, self.bucketD)                      # optD.zero_grad()
        with config.precision_scope(self.real_precision or config.precision()):
            outD = netD(inputs)
        log(3, outD.mean())
        lossD_real = crit(outD, True)
        lossD_real.backward()
        outG = netG(z1)
        outD = netD(outG.detach())
        log(4, outD.mean())
        lossD_fake = crit(outD, False)
        lossD_fake.backward()
        if parallel.enabled() and self.overlap:
            # D's gradient exchange + update run on a side stream while the generator forward of the G step (which does
            # not read D) runs on the main one; D's forward below waits for the join
            cur = torch.cuda.current_stream()
            self._side.wait_stream(cur)
            with torch.cuda.stream(self._side):
                self._step(self.optD, self.bucketD)
            self._zero(self.optG, self.bucketG)
            outG = netG(z2)
            cur.wait_stream(self._side)
        else:
            self._step(self.optD, self.bucketD)                  # (gradient all-reduce +) optD.step()
            self._zero(self.optG, self.bucketG)                  # optG.zero_grad()
            outG = netG(z2)
        # G step: D only relays the gradient to G. The reference also computes D's weight gradients here and throws
        # them away at the next optD.zero_grad() (main_dcgan.py:68,87-94; SURVEY.md §8d "minimal step"): skip them
        for p in self._d_params:
            p.requires_grad_(False)
        try:
            outD = netD(outG)
            log(5, outD.mean())
            lossG = crit(outD, False, True)
            lossG.backward()
        finally:
            for p in self._d_params:
                p.requires_grad_(True)
        self._step(self.optG, self.bucketG)                      # (gradient all-reduce +) optG.step()
        log(0, lossD_real.detach()), log(1, lossD_fake.detach()), log(2, lossG.detach())

    @staticmethod
    def _zero(opt, bucket):
        if bucket is None:
            opt.zero_grad()
        else:
            bucket.attach()

    @staticmethod
    def _step(opt, bucket):
        if bucket is not None:
            bucket.all_reduce_mean()
        opt.step()

    def _noise(self):
        return torch.randn(self.batch, self.z_dim, device=self.dev), torch.randn(self.batch, self.z_dim, device=self.dev)

    def step_eager(self, inputs, z=None):
        """Reference-faithful: every logged scalar is read on the host as soon as it exists (`.item()`)."""
        vals = [0.0] * 6

        def log(i, t):
            vals[i] = t.item()

        z1, z2 = (z[0], z[1]) if z is not None else self._noise()
        self._body(inputs, z1, z2, log)
        return vals

    # ---- capture must not train: the warm-up steps run real optimiser steps, so everything they touch is restored
    def _snapshot(self):
        snap = {"nets": [{k: v.detach().clone() for k, v in n.state_dict().items()} for n in (self.netG, self.netD)], "opts": []}
        for opt in (self.optG, self.optD):
            if isinstance(opt, FusedAdam):
                snap["opts"].append([None if st is None else {k: st[k].clone() for k in ("m", "v", "step")} for st in opt._flat])
            else:
                snap["opts"].append({id(p): {k: v.clone() for k, v in stt.items() if torch.is_tensor(v)}
                                     for p, stt in opt.state.items()})
        return snap

    @torch.no_grad()
    def _restore(self, snap):
        for net, sd in zip((self.netG, self.netD), snap["nets"]):
            cur = net.state_dict()
            for k, v in sd.items():
                cur[k].copy_(v)                      # in place: parameters stay views of the flat optimiser buffers
        for opt, saved in zip((self.optG, self.optD), snap["opts"]):
            if isinstance(opt, FusedAdam):
                for st, sv in zip(opt._flat, saved):
                    if st is not None:
                        for k in ("m", "v", "step"):
                            st[k].copy_(sv[k])
                for group in opt.param_groups:
                    for p in group["params"]:
                        p._gp_epoch = getattr(p, "_gp_epoch", 0) + 1
            else:
                for p, stt in opt.state.items():     # in place too: the captured graph references these tensors
                    old = saved.get(id(p))
                    for k, v in stt.items():
                        if torch.is_tensor(v):
                            if old is not None and k in old:
                                v.copy_(old[k])
                            else:
                                v.zero_()            # state created by the warm-up: back to torch.optim.Adam's initial zeros

    def _capture(self):
        snap = self._snapshot()
        _clear_caches(self.netG, self.netD)
        s = torch.cuda.Stream(device=self.dev)
        s.wait_stream(torch.cuda.current_stream())
        scal = self.scalars

        def log(i, t):
            scal[i].copy_(t)

        def body():
            if self.fixed_z:
                z1, z2 = self.z_static[0], self.z_static[1]
            else:
                z1, z2 = self._noise()
            self._body(self.x_static, z1, z2, log)

        with torch.cuda.stream(s):
            for _ in range(self._warmup):
                body()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        _clear_caches(self.netG, self.netD)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            body()
        self.graph = g
        self._restore(snap)
        torch.cuda.synchronize()

    def step(self, inputs, z=None):
        """One training step. Returns [lossD_real, lossD_fake, lossG, D(x), D(G(z))_1, D(G(z))_2] as Python floats."""
        if not self.use_graph:
            return self.step_eager(inputs, z)
        if self.graph is None:
            self.fixed_z = z is not None
            self._capture()
        if inputs.data_ptr() != self.x_static.data_ptr():
            self.x_static.copy_(inputs, non_blocking=True)
        if z is not None:
            self.z_static.copy_(z, non_blocking=True)
        self.graph.replay()
        # the replay updated the weights behind Python's back: operand copies cached by earlier eager calls are stale
        _clear_caches(self.netG, self.netD)
        return self.scalars.tolist()      # one device->host read per step
