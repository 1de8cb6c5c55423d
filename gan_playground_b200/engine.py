"""Step drivers for the adversarial loops of the reference's training scripts:

* `DcganStep` — main_dcgan.py:68-95 (two generator forwards per iteration),
* `SnganStep` — main_sngan.py:65-100 (class-conditional, ONE generator forward whose graph the G step re-uses, G step
  only every `n_disc_update` iterations),
* `AcganStep` — main_acgan.py:84-133 (two-head discriminator, adversarial + 0.5 x MSE auxiliary objective).

Each runs exactly the script's loop body (D-real backward, D-fake backward on G(z).detach(), optD.step, G step through
D, optG.step) either eagerly — one launch per kernel, host reads of the logged scalars like the reference's `.item()`
calls — or as ONE CUDA graph replay per step: the whole step (forward, autograd backward, NCCL collectives, Adam) is
captured on a side stream after a warm-up, the logged scalars are written to a device buffer inside the graph and read
back once per step. Graph replay removes the per-launch host overhead (~250 launches per step), which is what bounds
small per-GPU batches (strong scaling at 128 images/GPU).

Requirements for graph mode: torch optimisers built with `capturable=True` (FusedAdam is capturable as is); fixed batch
size. The weight-staging caches of the networks are cleared before capture so the staging kernels are part of the graph,
and the warm-up steps capture needs are undone afterwards (parameters, BatchNorm / spectral-norm buffers and optimiser
state are restored in place), so the first replay is the first training step."""
import os

import torch

from . import config, ops, parallel
from .criterion import ACGANLoss
from .optim import FusedAdam


def _clear_caches(*nets):
    for net in nets:
        for m in net.modules():
            c = getattr(m, "_gp_cache", None)
            if c is not None:
                c.clear()


class _AdversarialStep:
    """Shared machinery: gradient buckets / flat optimisers, the zero arena, eager stepping, graph capture + replay.
    A subclass supplies
      N_SCALARS                    how many numbers its script logs per iteration,
      _data_static / _noise_static lists of device tensors a replay reads its inputs from,
      _draw_noise()                fresh on-device noise when the caller supplies none,
      _loop(data, noise, log, i)   the script's loop body; `log(j, t)` receives logged scalar j as a 0-dim device tensor,
      _graph_key(i)                which captured variant iteration i replays (scripts whose body depends on `i`)."""

    N_SCALARS = 6

    def __init__(self, netG, netD, optG, optD, batch, device, use_graph=False, warmup=3, overlap=True):
        self.netG, self.netD, self.optG, self.optD = netG, netD, optG, optD
        self.batch, self.dev = batch, device
        # FusedAdam owns flat parameter / gradient buffers and does the gradient collective itself; any other optimiser
        # (torch.optim.Adam as in the reference scripts) gets a flat gradient bucket per network for the all-reduce
        self.bucketD = None if isinstance(optD, FusedAdam) else parallel.GradBucket(netD)
        self.bucketG = None if isinstance(optG, FusedAdam) else parallel.GradBucket(netG)
        self.use_graph = use_graph
        self.graphs = {}
        self.scalars = torch.zeros(self.N_SCALARS, device=device)
        self.fixed_noise = False
        self.iteration = 0
        self._warmup = warmup
        self.overlap = overlap
        self.arena = ops.ZeroArena(device)
        self._d_params = [p for p in netD.parameters() if p.requires_grad]
        self._side = torch.cuda.Stream(device=device) if device.type == "cuda" else None
        # weight-gradient GEMMs leave the backward chain for this stream when a launch cannot fill the GPU (config.py)
        self._wg_stream = (torch.cuda.Stream(device=device)
                           if device.type == "cuda" and config.wgrad_stream_wanted(batch) else None)

    # kept for callers that look at the single-variant graph (DcganStep / AcganStep)
    @property
    def graph(self):
        return next(iter(self.graphs.values()), None)

    def _graph_key(self, i):
        return 0

    # ---- one loop body inside the zero arena (one memset for every zero-initialised workspace of the step)
    def _body(self, data, noise, log, i):
        self.arena.begin()
        ops.ZeroArena.active = self.arena
        try:
            with config.wgrad_side_scope(self._wg_stream):
                self._loop(data, noise, log, i)
        finally:
            ops.ZeroArena.active = None

    @staticmethod
    def _zero(opt, bucket):
        if bucket is None:
            opt.zero_grad()
        else:
            bucket.attach()

    @staticmethod
    def _step(opt, bucket):
        config.join_wgrad_side()          # every weight gradient of this network has landed in its .grad buffer
        if bucket is not None:
            bucket.all_reduce_mean()
        opt.step()

    def _d_step_then(self, g_work):
        """optD.step() (with its gradient exchange), then `g_work()` — generator-side work that does not read D. Under
        data parallelism D's exchange + update run on a side stream while `g_work` runs on the main one; whatever reads D
        next waits for the join."""
        if parallel.enabled() and self.overlap:
            # join on the MAIN stream first: the join releases the tensors the weight gradients read, and the main stream's
            # allocator must not reuse them before those GEMMs are done
            config.join_wgrad_side()
            cur = torch.cuda.current_stream()
            self._side.wait_stream(cur)
            with torch.cuda.stream(self._side):
                self._step(self.optD, self.bucketD)
            out = g_work()
            cur.wait_stream(self._side)
            return out
        self._step(self.optD, self.bucketD)
        return g_work()

    class _frozen_d:
        """G step: D only relays the gradient to G. The reference also computes D's weight gradients there and throws
        them away at the next optD.zero_grad() (main_dcgan.py:68,87-94; SURVEY.md §8d "minimal step"): skip them."""

        def __init__(self, params):
            self.params = params

        def __enter__(self):
            for p in self.params:
                p.requires_grad_(False)

        def __exit__(self, *exc):
            for p in self.params:
                p.requires_grad_(True)

    # ---- eager: every logged scalar is read on the host as soon as it exists, like the reference's `.item()` calls
    def _run_eager(self, data, noise):
        vals = [float("nan")] * self.N_SCALARS

        def log(j, t):
            vals[j] = t.item()

        if noise is None:
            noise = self._draw_noise()
        self._body(data, noise, log, self.iteration)
        self.iteration += 1
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

    def _capture(self, key, i):
        snap = self._snapshot()
        _clear_caches(self.netG, self.netD)
        s = torch.cuda.Stream(device=self.dev)
        s.wait_stream(torch.cuda.current_stream())
        scal = self.scalars

        def log(j, t):
            scal[j].copy_(t)

        def body():
            noise = self._noise_static if self.fixed_noise else self._draw_noise()
            self._body(self._data_static, noise, log, i)

        with torch.cuda.stream(s):
            for _ in range(self._warmup):
                body()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        _clear_caches(self.netG, self.netD)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            body()
        self.graphs[key] = g
        self._restore(snap)
        torch.cuda.synchronize()

    def _run(self, data, noise):
        """One training step. Returns the script's logged scalars as Python floats (NaN for a number the script does
        not produce on this iteration)."""
        if not self.use_graph:
            return self._run_eager(data, noise)
        i = self.iteration
        key = self._graph_key(i)
        if not self.graphs:
            self.fixed_noise = noise is not None
        elif self.fixed_noise != (noise is not None):
            raise ValueError("graph mode: either every step supplies its noise or none does (captured with %s)"
                             % ("caller-supplied noise" if self.fixed_noise else "on-device noise"))
        if key not in self.graphs:
            self._capture(key, i)
        for st, t in zip(self._data_static, data):
            if t.data_ptr() != st.data_ptr():
                st.copy_(t, non_blocking=True)
        if noise is not None:
            for st, t in zip(self._noise_static, noise):
                st.copy_(t, non_blocking=True)
        if len(self.graphs) > 1:
            self.scalars.fill_(float("nan"))     # a variant that skips the G step leaves those entries untouched
        self.graphs[key].replay()
        self.iteration += 1
        # the replay updated the weights behind Python's back: operand copies cached by earlier eager calls are stale
        _clear_caches(self.netG, self.netD)
        return self.scalars.tolist()      # one device->host read per step


class DcganStep(_AdversarialStep):
    """main_dcgan.py:68-95. Logged scalars: [lossD_real, lossD_fake, lossG, D(x), D(G(z))_1, D(G(z))_2]."""

    def __init__(self, netG, netD, criterion, optG, optD, batch, z_dim, device, use_graph=False, warmup=3, overlap=True,
                 mixed_precision=True, fake_precision=None):
        super().__init__(netG, netD, optG, optD, batch, device, use_graph, warmup, overlap)
        self.crit, self.z_dim = criterion, z_dim
        self.x_static = torch.zeros(batch, netD.img_dim, netD.resolution, netD.resolution, device=device)
        self.z_static = torch.zeros(2, batch, z_dim, device=device)
        self._data_static = [self.x_static]
        self._noise_static = [self.z_static[0], self.z_static[1]]
        # mixed forward precision (config.precision_scope): only the G step needs the 3-MMA forward. The real-image D pass
        # meets every parity bar with plain bf16 operands (D-real cosine 0.99991), the D-fake chain G(z1) ->
        # D(G(z1).detach()) with single-MMA fp16 operands (D-fake cosine 0.99944, accumulated D gradient 0.99991:
        # tests/test_gpu_precision.py). GP_FAKE_PRECISION / fake_precision= override ("bf16x3" = no special case).
        self.real_precision = "bf16" if (mixed_precision and config.x3()) else None
        if fake_precision is None:
            # default only for the family it was measured on (models/dcgan.py nets); SN-DCGAN / others keep the global mode
            measured = all(type(n).__module__.endswith("models.dcgan") for n in (netG, netD))
            fake_precision = os.environ.get("GP_FAKE_PRECISION", "") or ("fp16" if measured else "bf16x3")
        self.fake_precision = fake_precision if (mixed_precision and config.x3() and fake_precision != "bf16x3") else None
        # G(z1) of the D-fake chain runs on its own stream next to the real-image pass at small shard sizes (config.py)
        self._g_stream = (torch.cuda.Stream(device=device)
                          if device.type == "cuda" and config.g_ahead_wanted(batch) else None)

    def _draw_noise(self):
        return torch.randn(self.batch, self.z_dim, device=self.dev), torch.randn(self.batch, self.z_dim, device=self.dev)

    def _loop(self, data, noise, log, i):
        netG, netD, crit = self.netG, self.netD, self.crit
        (inputs,), (z1, z2) = data, noise
        self._zero(self.optD, self.bucketD)                      # optD.zero_grad()
        ahead = self._g_stream is not None
        if ahead:
            # main_dcgan.py:77 issued early on a side stream: G(z1) reads only G, the real-image pass only D
            cur = torch.cuda.current_stream()
            self._g_stream.wait_stream(cur)
            with torch.cuda.stream(self._g_stream), parallel.peer_lane(1), \
                    config.precision_scope(self.fake_precision or config.precision()):
                outG = netG(z1)
        with config.precision_scope(self.real_precision or config.precision()):
            outD = netD(inputs)
        log(3, outD.mean())
        lossD_real = crit(outD, True)
        lossD_real.backward()
        with config.precision_scope(self.fake_precision or config.precision()):
            if ahead:
                cur.wait_stream(self._g_stream)
            else:
                outG = netG(z1)
            outD = netD(outG.detach())
        log(4, outD.mean())
        lossD_fake = crit(outD, False)
        lossD_fake.backward()

        def g_forward():
            self._zero(self.optG, self.bucketG)                  # optG.zero_grad()
            return netG(z2)

        outG = self._d_step_then(g_forward)                      # (gradient all-reduce +) optD.step()
        with self._frozen_d(self._d_params):
            outD = netD(outG)
            log(5, outD.mean())
            lossG = crit(outD, False, True)
            lossG.backward()
        self._step(self.optG, self.bucketG)                      # (gradient all-reduce +) optG.step()
        log(0, lossD_real.detach()), log(1, lossD_fake.detach()), log(2, lossG.detach())

    def step_eager(self, inputs, z=None):
        """Reference-faithful: every logged scalar is read on the host as soon as it exists (`.item()`)."""
        return self._run_eager([inputs], None if z is None else (z[0], z[1]))

    def step(self, inputs, z=None):
        """One training step. Returns [lossD_real, lossD_fake, lossG, D(x), D(G(z))_1, D(G(z))_2] as Python floats."""
        return self._run([inputs], None if z is None else (z[0], z[1]))


class SnganStep(_AdversarialStep):
    """main_sngan.py:65-100: class-conditional hinge GAN. The generator runs ONCE per iteration — the G step (:92-99)
    sends the batch the discriminator was just trained on through the updated D again and back-propagates into the
    generator graph of :82 — and only when `i % n_disc_update == 0` (`--n_disc_update`, default 5, :24).
    Logged scalars: [lossD_real, lossD_fake, lossG, D(x), D(G(z))_1, D(G(z))_2]; entries 2 and 5 are NaN on iterations
    without a G step (the script keeps printing its stale values there).
    Graph mode captures two variants (with / without the G step) and replays the one iteration `i` needs."""

    def __init__(self, netG, netD, criterion, optG, optD, batch, z_dim, device, n_classes=10, n_disc_update=5,
                 resolution=32, use_graph=False, warmup=3, overlap=True):
        super().__init__(netG, netD, optG, optD, batch, device, use_graph, warmup, overlap)
        if n_disc_update < 1:
            raise ValueError("n_disc_update must be >= 1")
        self.crit, self.z_dim, self.n_classes, self.n_disc_update = criterion, z_dim, n_classes, int(n_disc_update)
        img_dim = netD.block1.c1.in_channels
        self.x_static = torch.zeros(batch, img_dim, resolution, resolution, device=device)
        self.y_static = torch.zeros(batch, dtype=torch.long, device=device)
        self.z_static = torch.zeros(batch, z_dim, device=device)
        self.c_static = torch.zeros(batch, dtype=torch.long, device=device)
        self._data_static = [self.x_static, self.y_static]
        self._noise_static = [self.z_static, self.c_static]

    def _graph_key(self, i):
        return int(i % self.n_disc_update == 0)

    def _draw_noise(self):
        return (torch.randn(self.batch, self.z_dim, device=self.dev),
                torch.randint(self.n_classes, (self.batch,), device=self.dev))

    def _loop(self, data, noise, log, i):
        netG, netD, crit = self.netG, self.netD, self.crit
        (img_real, lbl_real), (z, c) = data, noise
        g_step = i % self.n_disc_update == 0
        self._zero(self.optD, self.bucketD)
        outD = netD(img_real, lbl_real)
        log(3, outD.mean())
        lossD_real = crit(outD, True)
        lossD_real.backward()
        outG = netG(z, c)
        outD = netD(outG.detach(), c)
        log(4, outD.mean())
        lossD_fake = crit(outD, False)
        lossD_fake.backward()
        self._step(self.optD, self.bucketD)      # the G step reads the updated D right away: nothing to overlap with
        log(0, lossD_real.detach()), log(1, lossD_fake.detach())
        if g_step:
            self._zero(self.optG, self.bucketG)
            with self._frozen_d(self._d_params):
                outD = netD(outG, c)
                log(5, outD.mean())
                lossG = crit(outD, False, True)
                lossG.backward()
            self._step(self.optG, self.bucketG)
            log(2, lossG.detach())

    def step_eager(self, img_real, lbl_real, z=None, c=None):
        return self._run_eager([img_real, lbl_real], None if z is None else (z, c))

    def step(self, img_real, lbl_real, z=None, c=None):
        """One iteration. Returns [lossD_real, lossD_fake, lossG, D(x), D(G(z))_1, D(G(z))_2] (NaN where not produced)."""
        return self._run([img_real, lbl_real], None if z is None else (z, c))


class AcganStep(_AdversarialStep):
    """main_acgan.py:84-133: two-head discriminator, objective `criterion_adv + 0.5 * MSELoss(aux head, labels)` on the
    real batch, on G(z, labels).detach() and — for the generator — on the same fake batch again (one generator forward
    per iteration, :107,123). Both heads come from one pass over the features and each of the three objectives is one
    fused value+gradient kernel (criterion.ACGANLoss on acgan.Discriminator.packed_logits).
    Logged scalars, in the order of the script's progress line (:136-137):
    [lossD_adv, lossD_aux, lossG_adv, lossG_aux, D(x), D(G(z))_1, D(G(z))_2] (the D(.) numbers are sigmoid means)."""

    N_SCALARS = 7

    def __init__(self, netG, netD, criterion_adv, optG, optD, batch, z_dim, device, n_class=10, aux_weight=0.5,
                 use_graph=False, warmup=3, overlap=True, mixed_precision=False):
        super().__init__(netG, netD, optG, optD, batch, device, use_graph, warmup, overlap)
        self.crit = criterion_adv if isinstance(criterion_adv, ACGANLoss) else ACGANLoss(criterion_adv, aux_weight)
        # per-pass forward precision as in DcganStep — OPT-IN here (not yet measured on the GPU for this loop): real pass
        # bf16, D's pass over the detached fake batch single-MMA fp16; the one generator forward feeds the G step, so it
        # and the G step's D pass stay in the global mode
        mixed = mixed_precision and config.x3()
        self.real_precision, self.fake_precision = ("bf16", "fp16") if mixed else (None, None)
        self.z_dim, self.n_class = z_dim, n_class
        self.x_static = torch.zeros(batch, netD.img_dim, netD.resolution, netD.resolution, device=device)
        self.y_static = torch.zeros(batch, n_class, device=device)
        self.z_static = torch.zeros(batch, z_dim, device=device)
        self._data_static = [self.x_static, self.y_static]
        self._noise_static = [self.z_static]

    def _draw_noise(self):
        return (torch.randn(self.batch, self.z_dim, device=self.dev),)

    def _loop(self, data, noise, log, i):
        netG, netD, crit = self.netG, self.netD, self.crit
        (img_real, lbl_real), (z,) = data, noise
        self._zero(self.optD, self.bucketD)
        with config.precision_scope(self.real_precision or config.precision()):
            real = crit(netD.packed_logits(img_real), lbl_real, True)
        real[crit.TOTAL].backward()
        log(4, real[crit.SIGMOID_MEAN].detach())
        c = lbl_real
        outG = netG(z, c)
        with config.precision_scope(self.fake_precision or config.precision()):
            fake = crit(netD.packed_logits(outG.detach()), c, False)
        fake[crit.TOTAL].backward()
        log(5, fake[crit.SIGMOID_MEAN].detach())
        self._step(self.optD, self.bucketD)
        d_terms = (real + fake).detach()         # lossD_adv = real_adv + fake_adv, lossD_aux likewise (:119-120)
        log(0, d_terms[crit.ADV]), log(1, d_terms[crit.AUX])
        self._zero(self.optG, self.bucketG)
        with self._frozen_d(self._d_params):
            gen = crit(netD.packed_logits(outG), c, False, True)
            gen[crit.TOTAL].backward()
        self._step(self.optG, self.bucketG)
        gen = gen.detach()
        log(2, gen[crit.ADV]), log(3, gen[crit.AUX]), log(6, gen[crit.SIGMOID_MEAN])

    def step_eager(self, img_real, lbl_real, z=None):
        return self._run_eager([img_real, lbl_real], None if z is None else (z,))

    def step(self, img_real, lbl_real, z=None):
        """One iteration. Returns [lossD_adv, lossD_aux, lossG_adv, lossG_aux, D(x), D(G(z))_1, D(G(z))_2]."""
        return self._run([img_real, lbl_real], None if z is None else (z,))
