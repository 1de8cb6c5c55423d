"""Autograd nodes of the hot path. Each node is one fused block of the reference networks
(conv / conv-transpose [+ BatchNorm] + activation, the first Linear, the image-side layers, the heads, the loss) whose
forward and backward are sequences of C-ABI kernel calls (ops.py). Activations between nodes are NHWC bf16 tensors;
parameters stay fp32 nn.Parameters in torch's layout and receive fp32 gradients, so torch.optim.Adam and
state_dict()/torch.save work unchanged.

Precision modes (config.py): in "bf16x3" every activation travels as a hi/lo bf16 pair — node outputs are
(a_hi, a_lo) with a_lo non-differentiable, and the next node takes (x, x_lo); autograd only ever sees the hi tensor
(gradients are w.r.t. the real-valued activation). In "bf16" the lo half is None. In "fp16" the companion tensor is an
fp16 COPY of the activation (dtype torch.float16 tells it apart from a bf16 lo half): forward GEMMs read it as their
single-MMA operand, everything on the backward side reads the bf16 tensor."""
import torch

from . import config, ops, parallel

BN_EPS = 1e-5
BN_MOMENTUM = 0.1


class WeightCache:
    """GEMM-operand copies of fp32 parameters, re-staged only when the parameter changes (optimizer.step bumps
    tensor._version). One D forward per step packs; the other two reuse (main_dcgan.py:70,80,91).

    4x4 conv weights go through `get_conv`, which remembers (parameter, format, orientation) of every request: the first
    stale request after an optimiser step re-stages EVERY remembered copy of the network in one launch
    (ops.stage_conv_weights — each fp32 weight is read once for all its formats) instead of one launch per layer, format
    and orientation. `clear()` drops the copies, not the list."""

    def __init__(self):
        self._d = {}
        self._recipes = {}

    @staticmethod
    def _ver(param):
        # _version: bumped by torch in-place updates (torch.optim.Adam); _gp_epoch: bumped by optim.FusedAdam, whose
        # kernel updates the storage behind autograd's back
        return (param._version, getattr(param, "_gp_epoch", 0))

    def _fresh(self, key, param):
        ent = self._d.get(key)
        return ent is not None and ent[0] == self._ver(param) and ent[1] == param.data_ptr()

    def get(self, key, param, make):
        # only real parameters are cached. A derived weight (W / sigma of spectral norm) is a new tensor every forward;
        # under no_grad / frozen parameters it is even a *leaf* with version 0 whose address the allocator may hand out
        # again next forward, so neither is_leaf nor (version, data_ptr) can tell two of them apart
        if not isinstance(param, torch.nn.Parameter) or not param.is_leaf:
            return make()
        if self._fresh(key, param):
            return self._d[key][2]
        val = make()
        self._d[key] = (self._ver(param), param.data_ptr(), val)
        return val

    def get_conv(self, key, param, fmt, n_dim, make):
        """A conv weight as [N][tap][C] operand in format `fmt` (ops.STAGE_*), orientation n_dim; `make` is the
        per-tensor fallback (shapes / devices the batched kernel does not take)."""
        if not isinstance(param, torch.nn.Parameter) or not param.is_leaf:
            return make()
        if not (config.batch_stage() and ops.stage_conv_ok(param)):
            return self.get(key, param, make)
        self._recipes[key] = (param, fmt, n_dim)
        if self._fresh(key, param):
            return self._d[key][2]
        stale = {}
        for k, (p, f, nd) in self._recipes.items():
            if not self._fresh(k, p):
                stale.setdefault(id(p), (p, []))[1].append((k, f, nd))
        groups = list(stale.values())
        outs = ops.stage_conv_weights([(p, [(f, nd) for _, f, nd in reqs]) for p, reqs in groups])
        for (p, reqs), tensors in zip(groups, outs):
            for (k, _, _), t in zip(reqs, tensors):
                self._d[k] = (self._ver(p), p.data_ptr(), t)
        return self._d[key][2]

    def clear(self):
        self._d.clear()


def with_lo(fn, h, *args):
    """Apply a node to an activation that may carry its low half as the `_gp_lo` attribute; re-attach the output's."""
    out = fn.apply(h, getattr(h, "_gp_lo", None), *args)
    if isinstance(out, tuple):
        a, lo = out
        if lo is not None:
            a._gp_lo = lo
        return a
    return out


def _f16_operand(x, x_lo):
    """The fp16 forward operand of an activation: its fp16 companion, or (input produced in another mode) a cast."""
    if x_lo is not None and x_lo.dtype == torch.float16:
        return x_lo
    return x.to(torch.float16) if x_lo is None else (x.float() + x_lo.float()).to(torch.float16)


class BwdLink:
    """Backward-side coupling of two CONSECUTIVE blocks of a sequential net (created by the model's forward, which knows
    that the producer's activation has exactly one consumer). The producer describes what its backward starts with —
    the activation derivative (kind "act": a, act) or the BatchNorm-backward reduction (kind "bn": y, fin, act) — and the
    consumer's data-gradient GEMM does that work in its epilogue (ops.conv_fwd(bwd=...)): one pass over dA less per layer.
    The consumer records the dA tensor it produced; the producer uses the fused result only when autograd hands it that
    very tensor."""
    __slots__ = ("kind", "a", "y", "fin", "act", "dx", "red")

    def __init__(self):
        self.kind = self.a = self.y = self.fin = self.act = self.dx = self.red = None

    def offer_act(self, a, act):
        if config.bwd_fusion() and act in ops._ACT_SLOPE:
            self.kind, self.a, self.act = "act", a, act

    def offer_bn(self, y, fin, act, training):
        if (config.bwd_fusion() and training and act in ops._ACT_SLOPE and y.dtype in (torch.float32, torch.bfloat16)
                and y.shape[-1] <= ops.MAX_STAT_COLS):
            self.kind, self.y, self.fin, self.act = "bn", y, fin, act

    def conv_bwd_arg(self, shape, device):
        """The bwd= argument for the consumer's data-gradient GEMM whose output has `shape` (None: nothing to fuse)."""
        if self.kind == "act" and tuple(self.a.shape) == tuple(shape):
            return ("mask", self.a, self.act)
        if self.kind == "bn" and tuple(self.y.shape) == tuple(shape):
            self.red = ops.zeros((2, shape[-1]), device)
            return ("bn", self.y, self.fin, self.act, self.red)
        return None

    def produced(self, dx):
        self.dx = dx

    def take(self, da):
        """Producer side: ("act", None) if da already carries the activation derivative, ("bn", red) if the sums of
        this da are ready, else (None, None). The link is consumed either way."""
        kind, dx, red = self.kind, self.dx, self.red
        self.dx = self.red = self.a = self.y = self.fin = None
        if dx is None:
            return None, None
        same = da.data_ptr() == dx.data_ptr() and da.shape == dx.shape
        if kind == "act":
            if not same:
                raise ops._lib.GpError("BwdLink: the activation derivative was fused into one consumer's data gradient, "
                                       "but the activation has further consumers (build the model without the link)")
            return "act", None
        return ("bn", red) if same else (None, None)


def _link_bwd(link, shape, device):
    return link.conv_bwd_arg(shape, device) if link is not None else None


def _bn_forward(y, gamma, beta, bufs, act, training=True, st=None, pair=False, comp=None, out_fmt=0):
    """y: pre-BN conv output, NHWC bf16 (fp32 in bf16x3 mode). Returns (a, a_lo, fin[4,C], count).
    training: batch statistics (all-reduced over ranks) + running-stat update, as nn.BatchNorm2d.train();
    eval: normalise with the running statistics. st: [2, C] sums already produced by the conv epilogue.
    comp / out_fmt: y is an existing activation with a companion tensor (csrc/act_io.cuh) / the companion format to
    produce (the ResNet and blur nodes, which normalise something other than a conv's fp32 output)."""
    C = y.shape[-1]
    count = (y.numel() // C) * parallel.world_size()
    rm, rv, nbt = bufs if bufs is not None else (None, None, None)
    f32 = y.dtype == torch.float32
    if y.dtype == torch.float16:          # 2-byte pre-BN storage of the fp16 mode: the tensor is its own fp16 companion;
        comp, out_fmt = y, ops.COMP_F16   # the activation leaves as a (bf16, fp16) pair (heads read the fp16 copy too)
    if training or rm is None:
        if st is None:
            st = ops.bn_stats_f32(y) if f32 else (ops.bn_stats_comp(y, comp) if comp is not None else ops.bn_stats(y))
        ctx = parallel.peer_ctx() if parallel.enabled() else None
        if ctx is not None and 2 * C <= ops.PEER_MAX_FLOATS:
            # SyncBN: one-shot NVLink exchange of the partial sums fused with the finalize (one launch, no NCCL)
            fin = ops.bn_finalize_peer(ctx, st, count, gamma, beta, rm, rv, nbt, BN_EPS, BN_MOMENTUM)
        else:
            parallel.all_reduce_sum_(st)
            fin = ops.bn_finalize(st, count, gamma, beta, rm, rv, nbt, BN_EPS, BN_MOMENTUM)
    else:
        fin = ops.bn_eval_params(rm, rv, gamma, beta, BN_EPS)
    if f32:
        a, a_lo = ops.bn_apply_act_pair(y, fin, act) if pair else ops.bn_apply_act_split(y, fin, act)
    elif comp is not None or out_fmt:
        a, a_lo = ops.bn_apply_act_comp(y, comp, fin, act, out_fmt)
    else:
        a, a_lo = ops.bn_apply_act(y, fin, act), None
    return a, a_lo, fin, count


def _bn_backward(da, y, fin, count, act, training=True, need_affine=True, comp=None, affine=None, red=None):
    """Returns (dy, dgamma, dbeta); dgamma / dbeta are None when the affine parameters need no gradient — or when their
    gradients were delivered straight into the parameters' .grad buffers by the apply kernel (`affine` = (gamma, beta)
    parameters that own such a buffer, ops.grad_target). comp: companion tensor of y — the backward must see the value
    the forward normalised, not its bf16 rounding (activation mask, xhat)."""
    f32 = y.dtype == torch.float32
    if y.dtype == torch.float16:          # 2-byte pre-BN storage of the fp16 mode (its own companion)
        comp = y
    if red is not None:
        pass            # accumulated by the epilogue of the GEMM that produced da (BwdLink)
    elif f32:
        red = ops.bn_bwd_reduce_f32(da, y, fin, act)
    elif comp is not None:
        red = ops.bn_bwd_reduce_comp(da, y, comp, fin, act)
    else:
        red = ops.bn_bwd_reduce(da, y, fin, act)
    parallel.all_reduce_sum_(red)
    # eval mode: statistics are constants, so the two batch-coupling terms vanish
    red_used = red if training else torch.zeros_like(red)
    # dbeta = sum dz, dgamma = sum dz * xhat. After the all-reduce `red` holds global sums of rank-local-mean-loss
    # gradients; parameter gradients are averaged over ranks later, so what is delivered is global / world.
    w = parallel.world_size()
    acc = None
    if need_affine and training and affine is not None:
        tg, tb = ops.grad_target(affine[0]), ops.grad_target(affine[1])
        if tg is not None and tb is not None:
            acc = (tb, tg)
    if f32:
        dy = ops.bn_bwd_apply_f32(da, y, fin, red_used, count, act, acc, 1.0 / w)
    elif comp is not None:
        dy = ops.bn_bwd_apply_comp(da, y, comp, fin, red_used, count, act, acc, 1.0 / w)
    else:
        dy = ops.bn_bwd_apply(da, y, fin, red_used, count, act, acc, 1.0 / w)
    if not need_affine or acc is not None:
        return dy, None, None
    dgamma, dbeta = red[1], red[0]
    if w > 1:
        dgamma, dbeta = dgamma / w, dbeta / w
    return dy, dgamma.clone(), dbeta.clone()


def _params(*ts):
    """The nn.Parameter objects among a node's inputs (None for derived tensors such as a spectral-normalised weight):
    kept on the ctx so that backward can deliver gradients straight into their .grad buffers (ops.grad_target)."""
    return tuple(t if isinstance(t, torch.nn.Parameter) else None for t in ts)


def _deliver_conv_wgrad(dwp, shape, param):
    """Packed fp32 weight gradient -> the parameter: accumulated straight into param.grad when it owns a buffer (returns
    None: autograd has nothing left to add), else returned as a fresh tensor in torch's layout."""
    tgt = ops.grad_target(param)
    if tgt is not None:
        ops.unpack_conv_wgrad(dwp, shape, into=tgt)
        return None
    return ops.unpack_conv_wgrad(dwp, shape)


def _wgrad_off_chain(param, fn, *reads):
    """Run `fn` (a weight-gradient GEMM + its delivery into param.grad) on the step driver's side stream when there is
    one (config.wgrad_side) and the gradient is delivered in place; `reads` are the tensors it reads, kept alive until the
    join so that the allocator cannot hand their memory to main-stream work that runs concurrently."""
    side = config.wgrad_side()
    if side is None or ops.grad_target(param) is None:
        return fn()
    stream, hold = side[0], side[1]
    stream.wait_stream(torch.cuda.current_stream())
    side[2] = True
    with torch.cuda.stream(stream):
        out = fn()
    hold.extend(reads)
    return out


def _deliver_matrix_grad(src, shape, param, *args, **kw):
    tgt = ops.grad_target(param)
    if tgt is not None:
        ops.unpack_matrix(src, shape, *args, into=tgt, **kw)
        return None
    return ops.unpack_matrix(src, shape, *args, **kw)


def _deliver_colsum(dy, bias):
    tgt = ops.grad_target(bias)
    if tgt is not None:
        ops.colsum(dy, into=tgt)
        return None
    return ops.colsum(dy)


class ConvBlock(torch.autograd.Function):
    """[Conv2d | ConvTranspose2d](k4 s2 p1, bias) [+ BatchNorm2d(train)] + activation on NHWC bf16.
    Reference: models/dcgan.py:35-40 (G blocks) and :104-110 (D blocks)."""

    @staticmethod
    def forward(ctx, x, x_lo, weight, bias, gamma, beta, bufs, transposed, act, cache, key, training=True,
                feeds_head=False, in_link=None, out_link=None):
        NB, H, W, Cin = x.shape
        ctx.in_link, ctx.out_link = in_link, out_link
        x3, fp16 = config.x3(), config.fp16()
        n_dim = 1 if transposed else 0
        if transposed:
            Ho, Wo, kind = 2 * H, 2 * W, ops.KIND_CONVT_K4S2
        else:
            Ho, Wo, kind = H // 2, W // 2, ops.KIND_CONV_K4S2
        has_bn = gamma is not None or bufs is not None
        b = bias.detach() if bias is not None else None
        ctx.set_materialize_grads(False)
        # BatchNorm batch statistics come out of the GEMM epilogue (fp32 accumulators), not from a second pass over y
        Cout = weight.shape[1] if transposed else weight.shape[0]
        st = None
        if has_bn and training and config.fused_stats() and Cout <= 2048:
            st = ops.zeros((2, Cout), x.device)
        if x3:
            wp = cache.get_conv((key, "fwd3"), weight, ops.STAGE_SPLIT, n_dim, lambda: ops.split_conv_weight(weight.detach(), n_dim))
            if x_lo is None or x_lo.dtype != torch.bfloat16:   # produced by a bf16 / fp16 pass: no low half to add
                x_lo = torch.zeros_like(x)
            y = ops.conv_fwd(x, wp, b, kind, Ho, Wo, ops.ACT_NONE if has_bn else act, stats=st, x_lo=x_lo,
                             out_mode="f32" if has_bn else "split")
        elif fp16:
            # one MMA on fp16 operands; the output pair is (bf16, fp16) — or a (hi, lo) bf16 pair when the head, which is
            # not a GEMM, consumes it
            wp = cache.get_conv((key, "fwdh"), weight, ops.STAGE_F16, n_dim, lambda: ops.conv_weight_f16(weight.detach(), n_dim))
            # pre-BatchNorm output: 2-byte fp16 storage when the statistics come from the epilogue (fp32 accumulators)
            prebn = "f16" if (st is not None and config.fp16_prebn() == "f16") else "f32"
            y = ops.conv_fwd(_f16_operand(x, x_lo), wp, b, kind, Ho, Wo, ops.ACT_NONE if has_bn else act, stats=st,
                             fp16_in=True, out_mode=prebn if has_bn else ("split" if feeds_head else "pair"))
        else:
            wp = cache.get_conv((key, "fwd"), weight, ops.STAGE_BF16, n_dim, lambda: ops.pack_conv_weight(weight.detach(), n_dim))
            y = ops.conv_fwd(x, wp, b, kind, Ho, Wo, ops.ACT_NONE if has_bn else act, stats=st)
        ctx.transposed, ctx.act, ctx.has_bn, ctx.cache, ctx.key = transposed, act, has_bn, cache, key
        ctx.params = _params(weight, bias, gamma, beta)   # the Parameter objects: gradients go straight into their .grad
        if has_bn:
            a, a_lo, fin, count = _bn_forward(y, gamma.detach() if gamma is not None else None,
                                              beta.detach() if beta is not None else None, bufs, act, training, st,
                                              pair=fp16 and not feeds_head)
            ctx.count, ctx.training = count, training
            ctx.save_for_backward(x, weight, y, fin)
            if out_link is not None:
                out_link.offer_bn(y, fin, act, training)
        else:
            a, a_lo = y if (x3 or fp16) else (y, None)
            ctx.save_for_backward(x, weight, a)
            if out_link is not None:
                out_link.offer_act(a, act)
        if a_lo is not None:
            ctx.mark_non_differentiable(a_lo)
        return a, a_lo

    @staticmethod
    def backward(ctx, da, _unused=None):
        if da is None:
            return (None,) * 15
        da = da.contiguous()
        fused, red = ctx.out_link.take(da) if ctx.out_link is not None else (None, None)
        if ctx.has_bn:
            x, weight, y, fin = ctx.saved_tensors
            dy, dgamma, dbeta = _bn_backward(da, y, fin, ctx.count, ctx.act, ctx.training,
                                             need_affine=ctx.needs_input_grad[4] or ctx.needs_input_grad[5],
                                             affine=ctx.params[2:4], red=red if fused == "bn" else None)
            # a bias in front of BatchNorm has an analytically zero gradient (the reference's is fp32 rounding noise,
            # SURVEY.md §2.2): hand autograd no tensor at all instead of a zero fill plus an accumulation pass
            dbias = None
        else:
            x, weight, a = ctx.saved_tensors
            dy = da if (ctx.act == ops.ACT_NONE or fused == "act") else ops.act_bwd(da, a, ctx.act)
            dgamma = dbeta = None
            dbias = _deliver_colsum(dy, ctx.params[1]) if ctx.needs_input_grad[3] else None
        dweight = dx = None
        if ctx.needs_input_grad[2]:
            def wgrad():
                if ctx.transposed:   # dW[Cin][tap][Cout]: dense = x (input grid), gathered = dy (output grid)
                    dwp = ops.conv_wgrad(x, dy, ops.KIND_CONV_K4S2, 16)
                else:                # dW[Cout][tap][Cin]: dense = dy (output grid), gathered = x (input grid)
                    dwp = ops.conv_wgrad(dy, x, ops.KIND_CONV_K4S2, 16)
                return _deliver_conv_wgrad(dwp, weight.shape, ctx.params[0])
            dweight = _wgrad_off_chain(ctx.params[0], wgrad, x, dy)
        if ctx.needs_input_grad[0]:
            NB, H, W, _ = x.shape
            bwd = _link_bwd(ctx.in_link, x.shape, x.device)   # the producing block's backward work, in this epilogue
            if ctx.transposed:   # dgrad of ConvT == strided conv over dy with weights [Cin][tap][Cout]
                wpd = ctx.cache.get_conv((ctx.key, "dgrad"), weight, ops.STAGE_BF16, 0, lambda: ops.pack_conv_weight(weight.detach(), 0))
                dx = ops.conv_fwd(dy, wpd, None, ops.KIND_CONV_K4S2, H, W, bwd=bwd)
            else:                # dgrad of Conv == 4-phase transposed conv over dy with weights [Cin][tap][Cout]
                wpd = ctx.cache.get_conv((ctx.key, "dgrad"), weight, ops.STAGE_BF16, 1, lambda: ops.pack_conv_weight(weight.detach(), 1))
                dx = ops.conv_fwd(dy, wpd, None, ops.KIND_CONVT_K4S2, H, W, bwd=bwd)
            if bwd is not None:
                ctx.in_link.produced(dx)
        if not ctx.needs_input_grad[3]:
            dbias = None
        return dx, None, dweight, dbias, dgamma, dbeta, None, None, None, None, None, None, None, None, None


class LinearToNHWC(torch.autograd.Function):
    """z (B, K) fp32 -> act(Linear(z)).view(B, C, bw, bw) delivered as NHWC bf16 (B, bw, bw, C).
    Reference: models/dcgan.py:32,50-51 (ReLU) and models/acgan.py:49-50 (no activation)."""

    @staticmethod
    def forward(ctx, z, weight, bias, bw, act, cache, key, out_link=None):
        B, K = z.shape
        ctx.out_link = out_link
        O = weight.shape[0]
        HW = bw * bw
        C = O // HW
        Kp = (K + 7) // 8 * 8
        ctx.set_materialize_grads(False)
        zc = z.detach().contiguous()
        # bias in NHWC-flatten order: dst[(s % HW)*C + s // HW] = bias[s]
        bp = cache.get((key, "bias"), bias, lambda: ops.unpack_matrix(bias.detach(), (O,), O, 1, 1, 1, 1, perm=C))
        fl = 2.0 * B * O * K
        if config.x3():
            zb, z_lo = ops.split_rows(zc, B, K, Kp, K, 1)
            wp = cache.get((key, "fwd3"), weight,
                           lambda: ops.split_weight_matrix(weight.detach(), O, K, O, Kp, K, 1, perm=HW))
            a, a_lo = ops.conv_fwd(zb.view(B, 1, 1, Kp), wp, bp, ops.KIND_CONV_K1S1, 1, 1, act, flops=fl,
                                   x_lo=z_lo.view(B, 1, 1, Kp), out_mode="split")
            a_lo = a_lo.view(B, bw, bw, C)
        elif config.fp16():
            zb, z_lo = ops.split_rows(zc, B, K, Kp, K, 1)          # zb (bf16) is what wgrad reads
            zh = ops.pair_to_f16(zb, z_lo, B, Kp, Kp)
            wp = cache.get((key, "fwdh"), weight,
                           lambda: ops.weight_matrix_f16(weight.detach(), O, K, O, Kp, K, 1, perm=HW))
            a, a_lo = ops.conv_fwd(zh.view(B, 1, 1, Kp), wp, bp, ops.KIND_CONV_K1S1, 1, 1, act, flops=fl, fp16_in=True,
                                   out_mode="pair")
            a_lo = a_lo.view(B, bw, bw, C)
        else:
            zb = ops.pack_matrix(zc, B, K, B, Kp, K, 1)
            wp = cache.get((key, "fwd"), weight, lambda: ops.pack_matrix(weight.detach(), O, K, O, Kp, K, 1, perm=HW))
            a, a_lo = ops.conv_fwd(zb.view(B, 1, 1, Kp), wp, bp, ops.KIND_CONV_K1S1, 1, 1, act, flops=fl), None
        ctx.save_for_backward(zb, a, weight)
        ctx.dims = (B, K, O, HW, C, Kp, act)
        ctx.params = _params(weight, bias)
        if a_lo is not None:
            ctx.mark_non_differentiable(a_lo)
        if out_link is not None:
            out_link.offer_act(a.view(B, bw, bw, C), act)
        return a.view(B, bw, bw, C), a_lo

    @staticmethod
    def backward(ctx, da, _unused=None):
        if da is None:
            return (None,) * 8
        zb, a, weight = ctx.saved_tensors
        B, K, O, HW, C, Kp, act = ctx.dims
        da = da.contiguous()
        fused, _ = ctx.out_link.take(da) if ctx.out_link is not None else (None, None)
        da = da.view(B, 1, 1, O)
        dy = da if (act == ops.ACT_NONE or fused == "act") else ops.act_bwd(da, a, act)
        dweight = dbias = None
        if ctx.needs_input_grad[1]:
            dwp = ops.conv_wgrad(dy, zb.view(B, 1, 1, Kp), ops.KIND_CONV_K1S1, 1, flops=2.0 * B * O * K)  # [O][1][Kp]
            dweight = _deliver_matrix_grad(dwp.view(O, Kp), weight.shape, ctx.params[0], O, K, Kp, K, 1, perm=HW)
        if ctx.needs_input_grad[2]:
            db = ops.colsum(dy)
            dbias = _deliver_matrix_grad(db, (O,), ctx.params[1], O, 1, 1, 1, 1, perm=HW)
        if ctx.needs_input_grad[0]:
            raise ops._lib.GpError("gradient w.r.t. the latent input of the first Linear is not implemented")
        return None, dweight, dbias, None, None, None, None, None


def linear_to_nhwc(z, weight, bias, bw, act, cache, key, out_link=None):
    a, lo = LinearToNHWC.apply(z, weight, bias, bw, act, cache, key, out_link)
    if lo is not None:
        a._gp_lo = lo
    return a


class ImageConv(torch.autograd.Function):
    """D's first layer: Conv2d(img_dim -> C, k4 s2 p1) + LeakyReLU(0.2) reading the fp32 NCHW image directly.
    Reference: models/dcgan.py:106-109 (block 0 has no BatchNorm). im2col -> 1-tap tensor-core GEMM (K = 64)."""

    @staticmethod
    def forward(ctx, x, weight, bias, act, cache, key, out_link=None):
        x = x.contiguous()
        NB, ch, H, W = x.shape
        Cout = weight.shape[0]
        fl = 2.0 * NB * (H // 2) * (W // 2) * Cout * ch * 16
        ctx.set_materialize_grads(False)
        ctx.out_link = out_link
        ctx.fused = ops.image_edge_ok(ch, H, W, Cout) and weight.is_contiguous()
        if ctx.fused:
            # csrc/image_edge.cu: the column tile lives in shared memory only; the image itself is what wgrad re-reads
            fmt = ops.COMP_LO if config.x3() else (ops.COMP_F16 if config.fp16() else ops.COMP_NONE)
            a, a_lo = _timed_edge("image_conv_fwd", fl, lambda: ops.image_conv_fwd(x.detach(), weight.detach(), bias.detach(), act, fmt),
                                  x, _act_bytes(NB, H // 2, W // 2, Cout, 1 if fmt == ops.COMP_NONE else 2))
            if a_lo is not None:
                ctx.mark_non_differentiable(a_lo)
            ctx.save_for_backward(x, weight, a)
            ctx.misc = (act, cache, key, (NB, ch, H, W))
            ctx.params = _params(weight, bias)
            if out_link is not None:
                out_link.offer_act(a, act)
            return a, a_lo
        if config.x3():
            col, col_lo = ops.im2col_k4s2_split(x.detach())
            wp = cache.get((key, "fwd3"), weight,
                           lambda: ops.split_weight_matrix(weight.detach(), Cout, ch * 16, Cout, 64, ch * 16, 1))
            a, a_lo = ops.conv_fwd(col, wp, bias.detach(), ops.KIND_CONV_K1S1, H // 2, W // 2, act, flops=fl,
                                   x_lo=col_lo, out_mode="split")
            ctx.mark_non_differentiable(a_lo)
        elif config.fp16():
            col, col_lo = ops.im2col_k4s2_split(x.detach())        # col (bf16) is kept for wgrad
            colh = ops.pair_to_f16(col, col_lo, col.numel() // 64, 64, 64).view(col.shape)
            wp = cache.get((key, "fwdh"), weight,
                           lambda: ops.weight_matrix_f16(weight.detach(), Cout, ch * 16, Cout, 64, ch * 16, 1))
            a, a_lo = ops.conv_fwd(colh, wp, bias.detach(), ops.KIND_CONV_K1S1, H // 2, W // 2, act, flops=fl,
                                   fp16_in=True, out_mode="pair")
            ctx.mark_non_differentiable(a_lo)
        else:
            col = ops.im2col_k4s2(x.detach())
            wp = cache.get((key, "fwd"), weight,
                           lambda: ops.pack_matrix(weight.detach(), Cout, ch * 16, Cout, 64, ch * 16, 1))
            a, a_lo = ops.conv_fwd(col, wp, bias.detach(), ops.KIND_CONV_K1S1, H // 2, W // 2, act, flops=fl), None
        ctx.save_for_backward(col, weight, a)   # the im2col buffer (hi half) is kept for wgrad instead of rebuilt
        ctx.misc = (act, cache, key, (NB, ch, H, W))
        ctx.params = _params(weight, bias)
        if out_link is not None:
            out_link.offer_act(a, act)
        return a, a_lo

    @staticmethod
    def backward(ctx, da, _unused=None):
        if da is None:
            return (None,) * 7
        col, weight, a = ctx.saved_tensors
        act, cache, key, (NB, ch, H, W) = ctx.misc
        Cout = weight.shape[0]
        fl = 2.0 * NB * (H // 2) * (W // 2) * Cout * ch * 16
        da = da.contiguous()
        fused, _ = ctx.out_link.take(da) if ctx.out_link is not None else (None, None)
        dy = da if (act == ops.ACT_NONE or fused == "act") else ops.act_bwd(da, a, act)
        dweight = dbias = dx = None
        if ctx.fused:
            x = col  # the fused forward saved the image, not a column buffer
            if ctx.needs_input_grad[1] or ctx.needs_input_grad[2]:
                # one pass over dy and the image: weight gradient in torch's layout and the bias gradient (the column
                # sums of dy, as the constant-1 column of the tile), added straight into the .grad buffers when they exist
                tw, tb = ops.grad_target(ctx.params[0]), ops.grad_target(ctx.params[1])
                dw = tw if tw is not None else ops.zeros(tuple(weight.shape), dy.device)
                db = None
                if ctx.needs_input_grad[2]:
                    db = tb if tb is not None else ops.zeros((Cout,), dy.device)
                _timed_edge("image_conv_wgrad", fl, lambda: ops.image_conv_wgrad(dy, x, None, dw, db), dy, x)
                dweight = None if tw is not None else dw
                dbias = None if (tb is not None or db is None) else db
            if ctx.needs_input_grad[0]:
                dx = _image_dgrad(dy, weight, ch, cache, key, H, W, fl)
            return dx, dweight, dbias, None, None, None, None
        if ctx.needs_input_grad[1]:
            dwp = ops.conv_wgrad(dy, col, ops.KIND_CONV_K1S1, 1, flops=fl)  # [Cout][1][64]
            dweight = _deliver_matrix_grad(dwp.view(Cout, 64), weight.shape, ctx.params[0], Cout, ch * 16, 64, ch * 16, 1)
        if ctx.needs_input_grad[2]:
            dbias = _deliver_colsum(dy, ctx.params[1])
        if ctx.needs_input_grad[0]:
            dx = _image_dgrad(dy, weight, ch, cache, key, H, W, fl)
        return dx, dweight, dbias, None, None, None, None


def _timed_edge(name, flops, fn, *tensors):
    """The fused image-edge launches are HBM-bound: bench.py accounts them by the bytes of the tensors they stream."""
    return ops._timed(name, flops, fn, nbytes=sum(t.numel() * t.element_size() for t in tensors if t is not None))


def _act_bytes(NB, H, W, C, copies=1):
    """a stand-in with .numel() / .element_size() for an NHWC 2-byte activation (or `copies` of them) not yet allocated"""
    class _B:
        def numel(self):
            return NB * H * W * C * copies

        def element_size(self):
            return 2
    return _B()


def _image_dgrad(dy, weight, ch, cache, key, H, W, fl):
    """Image gradient of D's first conv: dx = col2im(dy * W) — fused (csrc/image_edge.cu) when the shape allows."""
    Cout = weight.shape[0]
    if ops.image_edge_ok(ch, H, W, Cout, transposed=True) and weight.is_contiguous():
        return _timed_edge("image_convt_fwd (dgrad)", fl, lambda: ops.image_convt_fwd(dy, None, weight.detach(), None, ch, ops.ACT_NONE),
                           dy, _act_bytes(dy.shape[0], H, W, ch, 2))   # fp32 image gradient out
    # dcol[px][j] = sum_o dy[px][o] * W[o][j]  -> weights [64][Cout] = W^T (rows j >= ch*16 are zero)
    wpt = cache.get((key, "dgrad"), weight,
                    lambda: ops.pack_matrix(weight.detach(), ch * 16, Cout, 64, Cout, 1, ch * 16))
    dcol = ops.conv_fwd(dy, wpt, None, ops.KIND_CONV_K1S1, H // 2, W // 2, flops=fl)
    return ops.col2im_k4s2(dcol, None, ch, ops.ACT_NONE)


def image_conv(x, weight, bias, act, cache, key, out_link=None):
    a, lo = ImageConv.apply(x, weight, bias, act, cache, key, out_link)
    if lo is not None:
        a._gp_lo = lo
    return a


class ImageConvT(torch.autograd.Function):
    """G's last layer: ConvTranspose2d(C -> img_dim, k4 s2 p1) + Tanh writing the fp32 NCHW image.
    Reference: models/dcgan.py:41-44. 1-tap tensor-core GEMM (N = 64) -> col2im + bias + tanh."""

    @staticmethod
    def forward(ctx, x, x_lo, weight, bias, act, cache, key, in_link=None):
        NB, H, W, Cin = x.shape
        ch = weight.shape[1]
        fl = 2.0 * NB * H * W * Cin * ch * 16
        ctx.in_link = in_link
        if ops.image_edge_ok(ch, 2 * H, 2 * W, Cin, transposed=True) and weight.is_contiguous():
            # csrc/image_edge.cu: GEMM + col2im + bias + tanh in one kernel, the column values stay on the SM
            if config.x3():
                xa, xl = x, (x_lo if x_lo is not None else torch.zeros_like(x))
            elif config.fp16():
                xa, xl = _f16_operand(x, x_lo), None
            else:
                xa, xl = x, None
            out = _timed_edge("image_convt_fwd", fl, lambda: ops.image_convt_fwd(xa, xl, weight.detach(), bias.detach(), ch, act),
                              xa, xl, _act_bytes(NB, 2 * H, 2 * W, ch, 2))   # fp32 image out
            ctx.save_for_backward(x, weight, out)
            ctx.misc = (act, cache, key)
            ctx.params = _params(weight, bias)
            return out
        if config.x3():
            wp = cache.get((key, "fwd3"), weight,
                           lambda: ops.split_weight_matrix(weight.detach(), ch * 16, Cin, 64, Cin, 1, ch * 16))
            if x_lo is None:
                x_lo = torch.zeros_like(x)
            ycol = ops.conv_fwd(x, wp, None, ops.KIND_CONV_K1S1, H, W, flops=fl, x_lo=x_lo, out_mode="f32")
            out = ops.col2im_k4s2_f32(ycol, bias.detach(), ch, act)
        elif config.fp16():
            wp = cache.get((key, "fwdh"), weight,
                           lambda: ops.weight_matrix_f16(weight.detach(), ch * 16, Cin, 64, Cin, 1, ch * 16))
            ycol = ops.conv_fwd(_f16_operand(x, x_lo), wp, None, ops.KIND_CONV_K1S1, H, W, flops=fl, fp16_in=True,
                                out_mode="f32")
            out = ops.col2im_k4s2_f32(ycol, bias.detach(), ch, act)
        else:
            wp = cache.get((key, "fwd"), weight,
                           lambda: ops.pack_matrix(weight.detach(), ch * 16, Cin, 64, Cin, 1, ch * 16))
            ycol = ops.conv_fwd(x, wp, None, ops.KIND_CONV_K1S1, H, W, flops=fl)
            out = ops.col2im_k4s2(ycol, bias.detach(), ch, act)
        ctx.save_for_backward(x, weight, out)
        ctx.misc = (act, cache, key)
        ctx.params = _params(weight, bias)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, weight, out = ctx.saved_tensors
        act, cache, key = ctx.misc
        NB, H, W, Cin = x.shape
        ch = weight.shape[1]
        fl = 2.0 * NB * H * W * Cin * ch * 16
        dout = dout.contiguous()
        dweight = dbias = dx = None
        mul = out if act == ops.ACT_TANH else None
        bwd = _link_bwd(ctx.in_link, x.shape, x.device) if ctx.needs_input_grad[0] else None
        if ops.image_edge_ok(ch, 2 * H, 2 * W, Cin) and weight.is_contiguous() and bwd is None:
            # fused: the column tile of dout * tanh' is rebuilt in shared memory by both kernels instead of stored
            if ctx.needs_input_grad[2]:
                tw = ops.grad_target(ctx.params[0])
                dw = tw if tw is not None else ops.zeros(tuple(weight.shape), x.device)
                _timed_edge("image_conv_wgrad", fl, lambda: ops.image_conv_wgrad(x, dout, mul, dw), x, dout, mul)
                dweight = None if tw is not None else dw
            if ctx.needs_input_grad[3]:
                tgt = ops.grad_target(ctx.params[1])
                dbias = ops.image_bias_grad(dout, mul, into=tgt)
                if tgt is not None:
                    dbias = None
            if ctx.needs_input_grad[0]:
                dx, _ = _timed_edge("image_conv_fwd (dgrad)", fl,
                                    lambda: ops.image_conv_fwd(dout, weight.detach(), None, ops.ACT_NONE, ops.COMP_NONE, mul=mul),
                                    dout, mul, x)
            return dx, None, dweight, dbias, None, None, None, None
        # dcol[(n,ih,iw)][(co,kh,kw)] = dpre[n, co, 2ih-1+kh, 2iw-1+kw], dpre = dout * (1 - out^2) fused into the gather
        dcol = ops.im2col_k4s2(dout, out if act == ops.ACT_TANH else None)
        if ctx.needs_input_grad[2]:
            dwp = ops.conv_wgrad(x, dcol, ops.KIND_CONV_K1S1, 1, flops=fl)  # [Cin][1][64]
            dweight = _deliver_matrix_grad(dwp.view(Cin, 64), weight.shape, ctx.params[0], Cin, ch * 16, 64, ch * 16, 1)
        if ctx.needs_input_grad[3]:
            tgt = ops.grad_target(ctx.params[1])
            dbias = ops.image_bias_grad(dout, out if act == ops.ACT_TANH else None, into=tgt)
            if tgt is not None:
                dbias = None
        if ctx.needs_input_grad[0]:
            wpd = cache.get((key, "dgrad"), weight,
                            lambda: ops.pack_matrix(weight.detach(), Cin, ch * 16, Cin, 64, ch * 16, 1))
            dx = ops.conv_fwd(dcol, wpd, None, ops.KIND_CONV_K1S1, H, W, flops=fl, bwd=bwd)
            if bwd is not None:
                ctx.in_link.produced(dx)
        return dx, None, dweight, dbias, None, None, None, None


class Head(torch.autograd.Function):
    """Discriminator head on NHWC bf16 features: out[b][o] = bias[o] + sum_{hw,c} a[b,hw,c] * w[o, c, hw].
    flatten=False: global sum pooling + Linear (models/dcgan.py:121-122); flatten=True: NCHW flatten + Linear
    (models/dcgan_specnorm.py:125-126)."""

    @staticmethod
    def forward(ctx, a, a_lo, weight, bias, flatten):
        NB, H, W, C = a.shape
        O = weight.shape[0]
        w = weight.detach()
        if flatten and H * W > 1:
            # torch's NCHW flatten order (c, hw) -> the features' own order (hw, c): the kernels then read the weight
            # with 16-byte loads instead of one 4-byte load per 64-byte stride (a few KB, restaged per call)
            w = w.view(O, C, H * W).transpose(1, 2).contiguous()
            strides = (C * H * W, 1, C)
        else:
            strides = (C * H * W, H * W, 1) if flatten else (C, 1, 0)
        b = bias.detach() if bias is not None else None
        if a_lo is not None:            # bf16 low half or fp16 copy: read the most precise view of the features
            out = ops.head_fwd_comp(a, a_lo, w, b, O, *strides)
        else:
            out = ops.head_fwd(a, w, b, O, *strides)
        ctx.save_for_backward(a, w)
        ctx.strides, ctx.O, ctx.has_bias, ctx.restaged = strides, O, bias is not None, w.shape != weight.shape
        ctx.wshape = tuple(weight.shape)
        return out

    @staticmethod
    def backward(ctx, dout):
        a, w = ctx.saved_tensors
        da, dw, db = ops.head_bwd(dout.contiguous(), a, w, ctx.O, *ctx.strides,
                                  need_da=ctx.needs_input_grad[0], need_dw=ctx.needs_input_grad[2], need_db=ctx.has_bias)
        if dw is not None and ctx.restaged:     # [O][HW][C] -> torch's (O, C * HW)
            dw = dw.transpose(1, 2).reshape(ctx.wshape)
        return da, None, dw, (db if ctx.has_bias else None), None


class PackedHeads(torch.autograd.Function):
    """Both heads of the ACGAN discriminator (out_layer, out_aux: models/acgan.py:122-126) in ONE pass over the
    features: the two Linear weights are stacked into one (O1 + O2, C) head operand, the result is the packed logits
    (NB, O1 + O2), and one backward pass produces the feature gradient and both weight / bias gradients."""

    @staticmethod
    def forward(ctx, a, a_lo, w1, b1, w2, b2, flatten):
        NB, H, W, C = a.shape
        O1, O2 = w1.shape[0], w2.shape[0]
        w = torch.cat([w1.detach(), w2.detach()], 0)    # (O1 + O2, C) fp32, a few tens of KB
        b = torch.cat([b1.detach(), b2.detach()], 0)
        strides = (C * H * W, H * W, 1) if flatten else (C, 1, 0)
        if a_lo is not None:
            out = ops.head_fwd_comp(a, a_lo, w, b, O1 + O2, *strides)
        else:
            out = ops.head_fwd(a, w, b, O1 + O2, *strides)
        ctx.save_for_backward(a, w)
        ctx.strides, ctx.O1, ctx.O2 = strides, O1, O2
        return out

    @staticmethod
    def backward(ctx, dout):
        a, w = ctx.saved_tensors
        O1, O2 = ctx.O1, ctx.O2
        need_w = any(ctx.needs_input_grad[2:6])
        da, dw, db = ops.head_bwd(dout.contiguous(), a, w, O1 + O2, *ctx.strides,
                                  need_da=ctx.needs_input_grad[0], need_dw=need_w, need_db=need_w)
        if not need_w:
            return da, None, None, None, None, None, None
        return da, None, dw[:O1], db[:O1], dw[O1:], db[O1:], None


class AcganLossFn(torch.autograd.Function):
    """main_acgan.py:95-97,114-116,129-131 on the packed logits: returns the 4-vector [adversarial term, auxiliary MSE
    term, adv + aux_weight * aux, mean sigmoid(adv)]; only element 2 (the quantity the script calls .backward() on)
    carries a gradient, produced by the same kernel."""

    @staticmethod
    def forward(ctx, logits, labels, mode, target, aux_weight):
        out4, dlogits = ops.acgan_loss(logits.detach().contiguous(), labels.contiguous(), mode, target, aux_weight)
        ctx.save_for_backward(dlogits)
        return out4

    @staticmethod
    def backward(ctx, g):
        (dlogits,) = ctx.saved_tensors
        return dlogits * g[2], None, None, None, None


class GanLossFn(torch.autograd.Function):
    """GANLoss value + gradient in one fused reduction (utils/criterion.py:30-41)."""

    @staticmethod
    def forward(ctx, pred, mode, target):
        loss, dpred = ops.gan_loss(pred.detach().contiguous(), mode, target)
        ctx.save_for_backward(dpred)
        return loss

    @staticmethod
    def backward(ctx, g):
        (dpred,) = ctx.saved_tensors
        return dpred * g, None, None


class SpectralNormFn(torch.autograd.Function):
    """weight_orig -> weight_orig / sigma with one in-place power iteration on (u, v) per training forward
    (torch:nn/utils/spectral_norm.py:92-114). dim = 0 for Conv2d / Linear / Embedding, 1 for ConvTranspose2d.
    Backward differentiates through sigma with u, v constant: dW = (G - <G, W_sn> u v^T) / sigma."""

    @staticmethod
    def forward(ctx, weight_orig, u, v, dim, training):
        w = weight_orig.detach().contiguous()
        sigma = ops.sn_sigma(w, u, v, dim, training)
        w_sn = ops.sn_scale(w, sigma)
        ctx.save_for_backward(w_sn, u.clone(), v.clone(), sigma)   # clones: later forwards update u, v in place
        ctx.dim = dim
        return w_sn

    @staticmethod
    def backward(ctx, g):
        w_sn, u, v, sigma = ctx.saved_tensors
        return ops.sn_grad(g.contiguous(), w_sn, ctx.dim, u, v, sigma), None, None, None, None


class SpectralNormAllFn(torch.autograd.Function):
    """SpectralNormFn for EVERY spectral-normed weight a forward uses, in five launches instead of ~7 per hook
    (ops.sn_batched): forward(training, dims, n, w_0..w_{n-1}, u_0.., v_0..) -> (w_sn_0, ..., w_sn_{n-1}). Same semantics
    per hook as torch.nn.utils.spectral_norm's pre-forward hook (one in-place power iteration in training mode, sigma
    differentiated with u, v held constant)."""

    @staticmethod
    def forward(ctx, training, dims, n, *tensors):
        ws = [t.detach().contiguous() for t in tensors[:n]]
        us, vs = list(tensors[n:2 * n]), list(tensors[2 * n:3 * n])
        outs, sigma, keep, offs = ops.sn_batched(ws, us, vs, dims, training)
        ctx.save_for_backward(sigma, keep, *outs)
        ctx.meta = (tuple(dims), offs, n)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *gs):
        sigma, keep = ctx.saved_tensors[:2]
        w_sns = ctx.saved_tensors[2:]
        dims, offs, n = ctx.meta
        gs = [g.contiguous() if (g is not None and ctx.needs_input_grad[3 + i]) else None for i, g in enumerate(gs)]
        douts = ops.sn_grad_batched(gs, w_sns, sigma, keep, offs, dims)
        return (None, None, None) + tuple(douts) + (None,) * (2 * n)


def spectral_norm_all(modules, dims, training):
    """{module: W / sigma} for the spectral-normed `modules` (torch.nn.utils.spectral_norm holders: weight_orig, weight_u,
    weight_v) with one batched call; dims[i] = 0 (Conv2d / Linear / Embedding) or 1 (ConvTranspose2d)."""
    out = {}
    for s in range(0, len(modules), ops.SN_MAX):
        ms, ds = modules[s:s + ops.SN_MAX], tuple(dims[s:s + ops.SN_MAX])
        res = SpectralNormAllFn.apply(training, ds, len(ms), *[m.weight_orig for m in ms], *[m.weight_u for m in ms],
                                      *[m.weight_v for m in ms])
        out.update(zip(ms, res))
    return out
