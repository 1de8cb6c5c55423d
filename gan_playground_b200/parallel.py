"""Batch-sharded data parallelism for the adversarial step: one process per GPU, NCCL over NVLink.

The reference has no parallelism at all; the semantics defined here are "N ranks on B/N samples each
== the single-device reference at global batch B in exact arithmetic":
  * BatchNorm statistics (sum, sum of squares; and the two backward sums) are all-reduced -> global-batch statistics;
  * losses are means over the *global* batch, so parameter gradients are averaged (all-reduce sum / world);
  * parameters, Adam state, SN u/v and running statistics are replicated (broadcast once from rank 0).
"""
import contextlib
import os
import sys

import torch
import torch.distributed as dist

_state = {"enabled": False, "world": 1, "rank": 0, "group": None, "sync_grads": True}


def init(backend=None, device=None):
    """Initialise from torchrun-style environment variables (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return 0, 1
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        torch.cuda.set_device(local)
    if not dist.is_initialized():
        kw = {}
        if backend == "nccl":
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    _state.update(enabled=True, world=world, rank=rank)
    return rank, world


def shutdown():
    if dist.is_initialized():
        dist.destroy_process_group()
    _state.update(enabled=False, world=1, rank=0, peer=None)


def world_size():
    return _state["world"]


def rank():
    return _state["rank"]


def enabled():
    return _state["enabled"] and _state["world"] > 1


def sync_grads():
    """False inside `no_sync()` (gradient exchange skipped for an accumulating backward)."""
    return _state["sync_grads"]


PEER_LANES = 2


def init_peer_sync(device=None):
    """Set up the one-shot NVLink exchange used for the SyncBN reductions (csrc/peer_sync.cu): a symmetric buffer per
    rank, mapped into every peer through torch's symmetric-memory rendezvous (plumbing only: the exchange itself is our
    kernel). Returns True when active; on any failure (no peer access, > 8 ranks, GP_PEER_SYNC=0) the SyncBN sums keep
    going through NCCL all-reduce."""
    if not enabled() or _state.get("peer") is not None:
        return _state.get("peer") is not None
    if os.environ.get("GP_PEER_SYNC", "1") == "0" or dist.get_backend() != "nccl" or _state["world"] > 8:
        return False
    # A rank-local failure (symmetric allocation, rendezvous) is caught into a flag; the decision is then taken by a MIN
    # all-reduce that EVERY rank reaches, outside the try block — a rank must never fall back on its own while its peers
    # wait in a barrier or later spin in the peer kernel.
    from . import ops

    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else device
    local_ok, why, pending = 1, "", []
    try:
        import torch.distributed._symmetric_memory as symm_mem

        # PEER_LANES independent contexts (buffer + epoch each): exchanges issued on two streams that run side by side (the
        # step drivers' generator forward next to the real-image pass) must not share slots, flags or epochs
        for _ in range(PEER_LANES):
            n = ops.peer_buffer_bytes() // 4
            buf = symm_mem.empty(n, dtype=torch.float32, device=dev)
            buf.zero_()
            torch.cuda.synchronize()
            hdl = symm_mem.rendezvous(buf, dist.group.WORLD)
            ptrs = [int(p) for p in hdl.buffer_ptrs]
            epoch = torch.zeros(1, dtype=torch.int32, device=dev)
            torch.cuda.synchronize()
            if len(ptrs) != _state["world"] or not all(ptrs):
                local_ok, why = 0, "incomplete peer mapping"
                break
            pending.append({"ctx": ops.make_peer_ctx(ptrs, _state["rank"], epoch), "buf": buf, "hdl": hdl, "epoch": epoch})
    except Exception as exc:  # noqa: BLE001 — any rendezvous problem means "use NCCL", never a crash
        local_ok, why = 0, str(exc)
    ok = torch.tensor([local_ok], device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)         # doubles as the barrier after the buffers were zeroed
    if ok.item() != 1:
        if _state["rank"] == 0 or why:
            print("gan_playground_b200.parallel: peer SyncBN exchange unavailable on rank %d (%s); all ranks use NCCL "
                  "all-reduce" % (_state["rank"], why or "a peer failed"), file=sys.stderr)
        return False
    _state["peer"] = pending
    return True


def peer_ctx():
    """The peer-exchange context (ops.PeerCtx) of the current lane, or None when SyncBN sums go through NCCL."""
    p = _state.get("peer")
    return None if p is None else p[_state.get("peer_lane", 0)]["ctx"]


class peer_lane:
    """`with peer_lane(1): ...` — SyncBN exchanges issued inside use the second peer context. Every rank must issue the
    same exchanges in the same order PER LANE; lanes are independent of each other (different streams may interleave)."""

    def __init__(self, lane):
        if not 0 <= lane < PEER_LANES:
            raise ValueError("peer lane out of range")
        self.lane = lane

    def __enter__(self):
        self.prev = _state.get("peer_lane", 0)
        _state["peer_lane"] = self.lane

    def __exit__(self, *exc):
        _state["peer_lane"] = self.prev


def all_reduce_sum_(t):
    """In-place sum over ranks (SyncBN partial sums). No-op on a single rank."""
    if enabled():
        ctx = peer_ctx()
        if ctx is not None and t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.numel() <= 4096:
            from . import ops

            ops.peer_allreduce_sum_(ctx, t)
        else:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def reduce_scatter_sum_(flat, lo, hi):
    """flat[lo:hi] (this rank's slice) <- sum over ranks of flat[lo:hi]; the rest of `flat` is left unspecified.
    NCCL: one in-place reduce-scatter. gloo (CPU tests) has no reduce-scatter: all-reduce the whole buffer."""
    if not enabled():
        return
    if dist.get_backend() == "nccl":
        dist.reduce_scatter_tensor(flat[lo:hi], flat, op=dist.ReduceOp.SUM)
    else:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)


def all_gather_(flat, lo, hi):
    """Every rank's flat[lo:hi] slice -> the full `flat` on every rank (in place)."""
    if not enabled():
        return
    if dist.get_backend() == "nccl":
        dist.all_gather_into_tensor(flat, flat[lo:hi])
    else:
        parts = [torch.empty(hi - lo, dtype=flat.dtype, device=flat.device) for _ in range(_state["world"])]
        dist.all_gather(parts, flat[lo:hi].clone())
        flat.copy_(torch.cat(parts))


def broadcast_module(module, src=0):
    """Replicate parameters and buffers from rank `src` (done once after construction)."""
    if not enabled():
        return
    with torch.no_grad():
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t, src=src)       # in place on the parameter itself: bumps ._version for the operand caches
    for m in module.modules():               # and drop whatever an earlier forward staged from the old values
        c = getattr(m, "_gp_cache", None)
        if c is not None:
            c.clear()


def shard(t, dim=0):
    """Rank r's contiguous shard [r*B/N, (r+1)*B/N) of a global-batch tensor."""
    if not enabled():
        return t
    n = t.shape[dim] // _state["world"]
    return t.narrow(dim, _state["rank"] * n, n)


@contextlib.contextmanager
def no_sync():
    """Skip gradient all-reduces inside the block (use for the first of several accumulating backwards)."""
    prev = _state["sync_grads"]
    _state["sync_grads"] = False
    try:
        yield
    finally:
        _state["sync_grads"] = prev


class GradBucket:
    """Flat fp32 bucket for one network's gradients: one all-reduce per optimiser step.

    Usage:  bucket = GradBucket(net);  ...backward()...;  bucket.all_reduce_mean();  opt.step()
    Parameters' .grad become views into the flat buffer so no copy is needed before or after the collective."""

    def __init__(self, module):
        self.params = [p for p in module.parameters() if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(n, device=dev, dtype=torch.float32)
        self.views = []
        off = 0
        for p in self.params:
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()

    def attach(self):
        """Point every .grad at its slice of the flat buffer (zeroed). Call instead of optimizer.zero_grad()."""
        self.flat.zero_()
        for p, v in zip(self.params, self.views):
            p.grad = v

    def all_reduce_mean(self):
        if not enabled():
            return None
        # a backward may have replaced .grad with a fresh tensor (set_to_none semantics); gather those back
        for p, v in zip(self.params, self.views):
            if p.grad is not None and p.grad.data_ptr() != v.data_ptr():
                v.copy_(p.grad)
                p.grad = v
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)     # stream-ordered: the scale below runs after it
        self.flat.mul_(1.0 / _state["world"])
        return None


def all_reduce_grads_mean(module):
    """Simple (unbucketed-per-call) variant: average every parameter gradient across ranks."""
    if not enabled():
        return
    grads = [p.grad for p in module.parameters() if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat.mul_(1.0 / _state["world"])
    off = 0
    for g in grads:
        g.copy_(flat[off:off + g.numel()].view_as(g))
        off += g.numel()
