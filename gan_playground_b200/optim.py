"""Optimiser edge of the hot path: `FusedAdam` = torch.optim.Adam as the reference scripts
configure it (main_dcgan.py:55-56: lr 4e-4 / 1e-4, betas (0.5, 0.999); main_sngan.py:55-56: 2e-4, betas (0, 0.999);
eps 1e-8, no weight decay, no amsgrad) running as ONE kernel per network over flat fp32 buffers.

The parameters stay ordinary `nn.Parameter`s (state_dict / torch.save / the nets' forward are unaffected): their
storage is re-pointed into one flat buffer, their `.grad`s are views into a second flat buffer, so

  * `zero_grad()` is one memset,
  * the data-parallel gradient exchange is one collective on the flat gradient buffer (no bucket copies), and
  * with `shard=True` (ZeRO-1 style) each rank reduce-scatters the gradients, updates only its 1/N slice of
    (p, m, v) and all-gathers the parameters — same bytes on the wire as the all-reduce, 1/N of the optimiser work.

`torch.optim.Adam(net.parameters())` keeps working on the same nets (the reference scripts construct it themselves);
this class is what bench.py / engine.DcganStep use."""
import torch
import torch.distributed as dist

from . import ops, parallel


def _bump(p):
    # the bf16 operand caches (functional.WeightCache) key on this counter: the kernel updates p behind autograd's back
    p._gp_epoch = getattr(p, "_gp_epoch", 0) + 1


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, shard=False):
        if lr < 0 or eps < 0 or not (0 <= betas[0] < 1) or not (0 <= betas[1] < 1):
            raise ValueError("invalid Adam hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps))
        self.shard = bool(shard) and parallel.enabled()
        self._flat = []
        for group in self.param_groups:
            self._flat.append(self._flatten(group))

    # ---- flat storage ------------------------------------------------------------------------------------------
    def _flatten(self, group):
        ps = [p for p in group["params"] if p.requires_grad]
        if not ps:
            return None
        dev = ps[0].device   # the update kernel (ops.adam_flat) refuses CPU tensors: build after module.to(cuda)
        world = parallel.world_size() if self.shard else 1
        offs, n = [], 0
        for p in ps:
            if p.dtype != torch.float32:
                raise ops._lib.GpError("FusedAdam handles fp32 parameters only")
            offs.append(n)
            n += (p.numel() + 3) // 4 * 4          # every parameter starts 16-byte aligned
        quantum = 4 * world
        n = (n + quantum - 1) // quantum * quantum
        fp = torch.zeros(n, device=dev, dtype=torch.float32)
        fg = torch.zeros(n, device=dev, dtype=torch.float32)
        gviews = []
        with torch.no_grad():
            for p, o in zip(ps, offs):
                v = fp[o:o + p.numel()].view_as(p)
                v.copy_(p.data)
                p.data = v
                gviews.append(fg[o:o + p.numel()].view_as(p))
                _bump(p)
        sh = n // world
        r = parallel.rank() if self.shard else 0
        st = {"params": ps, "offs": offs, "n": n, "p": fp, "g": fg, "gviews": gviews, "lo": r * sh, "hi": (r + 1) * sh,
              "m": torch.zeros(sh, device=dev, dtype=torch.float32), "v": torch.zeros(sh, device=dev, dtype=torch.float32),
              "step": torch.zeros((), device=dev, dtype=torch.float32)}
        return st

    def zero_grad(self, set_to_none=False):
        """One memset of the flat gradient buffer; every .grad is (re)pointed at its slice of it."""
        for st in self._flat:
            if st is None:
                continue
            st["g"].zero_()
            for p, v in zip(st["params"], st["gviews"]):
                p.grad = v

    def _collect(self, st):
        # a backward that ran while .grad was None created a fresh tensor: fold it back into the flat buffer.
        # NOTE: a parameter that received NO gradient is updated with a zero gradient (its moments decay and it moves by
        # lr * m / (sqrt(v) + eps)), whereas torch.optim.Adam skips it: FusedAdam requires every parameter to take part
        # in every step — true for all networks / loops of the hot path (each step back-propagates through the whole net).
        for p, v in zip(st["params"], st["gviews"]):
            g = p.grad
            if g is None:
                v.zero_()
            elif g.data_ptr() != v.data_ptr():
                v.copy_(g)
            p.grad = v

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        for group, st in zip(self.param_groups, self._flat):
            if st is None:
                continue
            self._collect(st)
            b1, b2 = group["betas"]
            scale = 1.0
            lo, hi = st["lo"], st["hi"]
            if parallel.enabled() and parallel.sync_grads():
                scale = 1.0 / parallel.world_size()
                if self.shard:
                    parallel.reduce_scatter_sum_(st["g"], lo, hi)
                else:
                    dist.all_reduce(st["g"], op=dist.ReduceOp.SUM)
            ops.adam_flat(st["p"][lo:hi], st["g"][lo:hi], st["m"], st["v"], st["step"], group["lr"], b1, b2, group["eps"],
                          scale)
            st["step"] += 1
            if self.shard:
                parallel.all_gather_(st["p"], lo, hi)
            for p in st["params"]:
                _bump(p)
        return loss

    # ---- torch.optim.Adam-shaped state for checkpoints ------------------------------------------------------------
    @staticmethod
    def _torch_adam_group_defaults():
        """Every key torch.optim.Adam keeps in a param group (weight_decay, amsgrad, maximize, foreach, capturable,
        differentiable, fused, ... — whatever this torch version has), at the values the reference scripts run with:
        torch's load_state_dict REPLACES the groups wholesale, so a checkpoint that lacks them breaks the next step()."""
        d = dict(torch.optim.Adam([torch.zeros(1)]).defaults)
        d.pop("lr", None), d.pop("betas", None), d.pop("eps", None)
        return d

    def state_dict(self):
        """Same layout as torch.optim.Adam.state_dict(): per-parameter step (a CPU fp32 scalar, as torch writes it for a
        non-capturable Adam) / exp_avg / exp_avg_sq, and param groups carrying every torch.optim.Adam key, so the
        checkpoint loads into the reference scripts' torch.optim.Adam and steps. With shard=True the moments of other
        ranks' slices are zeros (gather the ranks' state dicts to checkpoint a sharded run)."""
        state, idx = {}, 0
        groups = []
        extra = self._torch_adam_group_defaults()
        for group, st in zip(self.param_groups, self._flat):
            ids = []
            for p in group["params"]:
                ids.append(idx)
                idx += 1
            groups.append({**extra, **{k: v for k, v in group.items() if k != "params"}, "params": ids})
            if st is None:
                continue
            full_m = torch.zeros(st["n"], device=st["m"].device)
            full_v = torch.zeros(st["n"], device=st["m"].device)
            full_m[st["lo"]:st["hi"]] = st["m"]
            full_v[st["lo"]:st["hi"]] = st["v"]
            pid = {id(p): i for i, p in zip(ids, group["params"])}
            for p, o in zip(st["params"], st["offs"]):
                state[pid[id(p)]] = {"step": st["step"].detach().cpu().clone(), "exp_avg": full_m[o:o + p.numel()].view_as(p).clone(),
                                     "exp_avg_sq": full_v[o:o + p.numel()].view_as(p).clone()}
        return {"state": state, "param_groups": groups}

    def load_state_dict(self, sd):
        idx = 0
        for group, st, g_sd in zip(self.param_groups, self._flat, sd["param_groups"]):
            for k, v in g_sd.items():
                if k in ("lr", "betas", "eps"):          # the other torch.optim.Adam keys have no meaning here
                    group[k] = v
            ids = list(range(idx, idx + len(group["params"])))
            idx += len(group["params"])
            if st is None:
                continue
            pid = {id(p): i for i, p in zip(ids, group["params"])}
            full_m = torch.zeros(st["n"], device=st["m"].device)
            full_v = torch.zeros(st["n"], device=st["m"].device)
            for p, o in zip(st["params"], st["offs"]):
                e = sd["state"].get(pid[id(p)])
                if e is None:
                    continue
                full_m[o:o + p.numel()] = e["exp_avg"].reshape(-1).to(full_m)
                full_v[o:o + p.numel()] = e["exp_avg_sq"].reshape(-1).to(full_v)
                st["step"].fill_(float(e["step"]))
            st["m"].copy_(full_m[st["lo"]:st["hi"]])
            st["v"].copy_(full_v[st["lo"]:st["hi"]])
