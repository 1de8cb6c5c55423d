"""Autograd nodes for the SNGAN projection networks (reference: models/sngan_projection.py): stride-1 convs with fused
residual add, conditional BatchNorm with fused nearest-upsample, pooling / upsampling, the 3-channel image-side layers
and the projection head. Same conventions as functional.py: NHWC bf16 activations, fp32 torch-layout parameters.

Forward precision (config.py): every node returns (a, a_comp) — the bf16 activation autograd sees and its
non-differentiable COMPANION tensor (csrc/act_io.cuh): None in "bf16"; the fp16 copy of the value in "fp16" (forward GEMMs
run ONE MMA on fp16 operands); the bf16 low half in "bf16x3" (three MMAs on hi/lo pairs). Forward element-wise kernels read
the most precise view and write both tensors; everything on the backward side reads the bf16 tensor, so the backward is
the same in all modes. A fused residual (the ResNet blocks) is implemented for "bf16" / "fp16" only; the projection
networks therefore run the global "bf16x3" default as "fp16" (config.resnet_scope) — they have no BatchNorm in D, and
11-bit operands meet every north_star bar there. The dcgan_blur networks (BatchNorm in D, no residuals)
run all three modes."""
import torch

from . import config, ops, parallel
from .functional import (BN_EPS, BN_MOMENTUM, _bn_backward, _bn_forward, _deliver_colsum, _deliver_conv_wgrad,
                         _f16_operand, _params)


def _fmt():
    """Companion format the forward nodes produce in the current mode."""
    return ops.COMP_LO if config.x3() else (ops.COMP_F16 if config.fp16() else ops.COMP_NONE)


def _f16(comp):
    return comp if (comp is not None and comp.dtype == torch.float16) else None


def _lo(x, comp):
    """The bf16 low half of an activation for a bf16x3 GEMM (zeros when it was produced in another mode)."""
    return comp if (comp is not None and comp.dtype == torch.bfloat16) else torch.zeros_like(x)


def attach(out):
    """(a, a_comp) -> a with the companion riding along as `_gp_lo` (the convention of functional.with_lo)."""
    a, comp = out
    if comp is not None:
        a._gp_lo = comp
    return a


def comp_of(h):
    return getattr(h, "_gp_lo", None) if h is not None else None


class Conv2dNHWC(torch.autograd.Function):
    """nn.Conv2d(k=3, p=1) / nn.Conv2d(k=1) (+ bias) [+ residual] + activation, no normalisation.
    Reference: ResGenBlock / ResDisBlock convs c1, c2, c_sc (models/sngan_projection.py:30-44,105-119)."""

    @staticmethod
    def forward(ctx, x, x_comp, weight, bias, residual, res_comp, act, cache, key):
        ctx.set_materialize_grads(False)   # no zero tensors for the companion outputs
        NB, H, W, _ = x.shape
        ksize = weight.shape[2]
        kind = ops.KIND_CONV_K3S1 if ksize == 3 else ops.KIND_CONV_K1S1
        b = bias.detach() if bias is not None else None
        if _fmt() == ops.COMP_F16:
            wp = cache.get((key, "fwdh"), weight, lambda: ops.conv_weight_f16(weight.detach(), 0))
            res = _f16(res_comp) if _f16(res_comp) is not None else residual
            a, a_comp = ops.conv_fwd(_f16_operand(x, x_comp), wp, b, kind, H, W, act, residual=res, fp16_in=True,
                                     out_mode="pair")
            ctx.mark_non_differentiable(a_comp)
        elif _fmt() == ops.COMP_LO:
            if residual is not None:
                raise ops._lib.GpError("Conv2dNHWC: the fused residual is implemented for the bf16 / fp16 modes "
                                       "(run the ResNet blocks inside config.resnet_scope())")
            wp = cache.get((key, "fwd3"), weight, lambda: ops.split_conv_weight(weight.detach(), 0))
            a, a_comp = ops.conv_fwd(x, wp, b, kind, H, W, act, x_lo=_lo(x, x_comp), out_mode="split")
            ctx.mark_non_differentiable(a_comp)
        else:
            wp = cache.get((key, "fwd"), weight, lambda: ops.pack_conv_weight(weight.detach(), 0))
            a, a_comp = ops.conv_fwd(x, wp, b, kind, H, W, act, residual=residual), None
        ctx.save_for_backward(x, weight, a)
        ctx.misc = (kind, ksize * ksize, act, cache, key, residual is not None, bias is not None)
        ctx.params = _params(weight, bias)
        return a, a_comp

    @staticmethod
    def backward(ctx, da, _unused=None):
        if da is None:
            return (None,) * 9
        x, weight, a = ctx.saved_tensors
        kind, taps, act, cache, key, has_res, has_bias = ctx.misc
        da = da.contiguous()
        dy = ops.act_bwd(da, a, act) if act != ops.ACT_NONE else da
        NB, H, W, _ = x.shape
        dx = dweight = dbias = None
        if ctx.needs_input_grad[2]:
            dweight = _deliver_conv_wgrad(ops.conv_wgrad(dy, x, kind, taps), weight.shape, ctx.params[0])
        if has_bias and ctx.needs_input_grad[3]:
            dbias = _deliver_colsum(dy, ctx.params[1])
        if ctx.needs_input_grad[0]:
            # dgrad of a stride-1 conv: the same conv with channels swapped and the tap order reversed
            wpd = cache.get((key, "dgrad"), weight, lambda: ops.pack_conv_weight(weight.detach(), 1 | 2))
            dx = ops.conv_fwd(dy, wpd, None, kind, H, W)
        dres = dy if (has_res and ctx.needs_input_grad[4]) else None
        return dx, None, dweight, dbias, dres, None, None, None, None


class BNAct(torch.autograd.Function):
    """nn.BatchNorm2d (affine) + activation on an existing NHWC tensor (generator's b6 + ReLU, :92-93)."""

    @staticmethod
    def forward(ctx, y, y_comp, gamma, beta, bufs, act, training):
        ctx.set_materialize_grads(False)   # no zero tensors for the companion outputs
        a, a_comp, fin, count = _bn_forward(y, gamma.detach(), beta.detach(), bufs, act, training, comp=y_comp,
                                            out_fmt=_fmt())
        # the backward normalises the SAME value: keep y's companion (activation mask / xhat from the bf16 rounding of y
        # cost the dcgan_blur G step 1.6e-3 of gradient cosine)
        ctx.has_comp = y_comp is not None
        ctx.save_for_backward(y, fin, *((y_comp,) if y_comp is not None else ()))
        ctx.misc = (count, act, training)
        ctx.params = _params(gamma, beta)
        if a_comp is not None:
            ctx.mark_non_differentiable(a_comp)
        return a, a_comp

    @staticmethod
    def backward(ctx, da, _unused=None):
        if da is None:
            return (None,) * 7
        y, fin = ctx.saved_tensors[:2]
        y_comp = ctx.saved_tensors[2] if ctx.has_comp else None
        count, act, training = ctx.misc
        dy, dgamma, dbeta = _bn_backward(da.contiguous(), y, fin, count, act, training, comp=y_comp, affine=ctx.params)
        return dy, None, dgamma, dbeta, None, None, None


class CondBNAct(torch.autograd.Function):
    """ConditionalBatchNorm2d (models/sngan_projection.py:6-19) + activation [+ nearest x2 upsample (:53)]:
    batch statistics (synchronised over ranks) -> xhat * embed(y)[:C] + embed(y)[C:] -> act -> upsample, one pass."""

    @staticmethod
    def forward(ctx, x, x_comp, emb, labels, bufs, act, upsample, training):
        ctx.set_materialize_grads(False)   # no zero tensors for the companion outputs
        C = x.shape[-1]
        count = (x.numel() // C) * parallel.world_size()
        rm, rv, nbt = bufs
        fmt = _fmt()
        if training:
            st = ops.bn_stats_comp(x, x_comp) if x_comp is not None else ops.bn_stats(x)
            parallel.all_reduce_sum_(st)
            fin = ops.bn_finalize(st, count, None, None, rm, rv, nbt, BN_EPS, BN_MOMENTUM)
        else:
            fin = ops.bn_eval_params(rm, rv, None, None, BN_EPS)
        e = emb.detach()
        if fmt or x_comp is not None:
            out, out_comp = ops.cbn_apply_act(x, fin, e, labels, act, upsample, comp=x_comp, out_fmt=fmt or ops.comp_fmt_of(x_comp))
            if not fmt:
                out_comp = None
        else:
            out, out_comp = ops.cbn_apply_act(x, fin, e, labels, act, upsample), None
        ctx.has_comp = x_comp is not None
        ctx.save_for_backward(x, fin, e, labels, *((x_comp,) if x_comp is not None else ()))
        ctx.misc = (count, act, upsample, training, emb.shape[0])
        if out_comp is not None:
            ctx.mark_non_differentiable(out_comp)
        return out, out_comp

    @staticmethod
    def backward(ctx, da, _unused=None):
        if da is None:
            return (None,) * 8
        x, fin, e, labels = ctx.saved_tensors[:4]
        x_comp = ctx.saved_tensors[4] if ctx.has_comp else None
        count, act, upsample, training, ncls = ctx.misc
        da = da.contiguous()
        S, demb = ops.cbn_bwd_reduce(da, x, fin, e, labels, act, upsample, ncls, comp=x_comp)
        parallel.all_reduce_sum_(S)
        if not training:
            S = torch.zeros_like(S)
        dx = ops.cbn_bwd_apply(da, x, fin, e, labels, S, count, act, upsample, comp=x_comp)
        return dx, None, demb, None, None, None, None, None


class Pool2x(torch.autograd.Function):
    """F.avg_pool2d(x, 2) (:128,132)."""

    @staticmethod
    def forward(ctx, x, x_comp):
        ctx.set_materialize_grads(False)   # no zero tensors for the companion outputs
        fmt = _fmt()
        if not fmt:
            return ops.pool2x(x, 0.25), None
        out, out_comp = ops.pool2x(x, 0.25, comp=x_comp, out_fmt=fmt)
        ctx.mark_non_differentiable(out_comp)
        return out, out_comp

    @staticmethod
    def backward(ctx, g, _unused=None):
        return (ops.upsample2x(g.contiguous(), 0.25) if g is not None else None), None


class Upsample2x(torch.autograd.Function):
    """F.interpolate(x, scale_factor=2) (nearest) (:60)."""

    @staticmethod
    def forward(ctx, x, x_comp):
        ctx.set_materialize_grads(False)   # no zero tensors for the companion outputs
        out = ops.upsample2x(x, 1.0)
        if not _fmt() or x_comp is None:
            return out, None
        out_comp = ops.upsample2x(x_comp, 1.0)      # a pure copy: exact for either companion format
        ctx.mark_non_differentiable(out_comp)
        return out, out_comp

    @staticmethod
    def backward(ctx, g, _unused=None):
        return (ops.pool2x(g.contiguous(), 1.0) if g is not None else None), None


class ReluFn(torch.autograd.Function):
    """F.relu on a block input whose raw value is still needed by the shortcut (:122)."""

    @staticmethod
    def forward(ctx, x, x_comp):
        ctx.set_materialize_grads(False)   # no zero tensors for the companion outputs
        fmt = _fmt()
        if fmt:
            a, a_comp = ops.act_fwd(x, ops.ACT_RELU, comp=x_comp, out_fmt=fmt)
            ctx.mark_non_differentiable(a_comp)
        else:
            a, a_comp = ops.act_fwd(x, ops.ACT_RELU), None
        ctx.save_for_backward(a)
        return a, a_comp

    @staticmethod
    def backward(ctx, g, _unused=None):
        if g is None:
            return None, None
        (a,) = ctx.saved_tensors
        return ops.act_bwd(g.contiguous(), a, ops.ACT_RELU), None


def _center_tap(wsc):
    """(Cout, ch, 1, 1) 1x1 kernel embedded as the centre tap of a 3x3 kernel: lets the shortcut share c1's im2col."""
    full = torch.zeros((wsc.shape[0], wsc.shape[1], 3, 3), device=wsc.device, dtype=wsc.dtype)
    full[:, :, 1, 1] = wsc[:, :, 0, 0]
    return full


class ImageConv3(torch.autograd.Function):
    """First block of the projection discriminator on the fp32 NCHW image: h1 = relu(c1(x)) (3x3) and s = c_sc(x) (1x1)
    from ONE im2col (models/sngan_projection.py:156-163; avg-pooling commutes with the sum, so both are produced at
    full resolution and pooled once after c2). 1-tap tensor-core GEMMs with K = 32."""

    @staticmethod
    def forward(ctx, x, w1, b1, wsc, bsc):
        ctx.set_materialize_grads(False)   # no zero tensors for the companion outputs
        x = x.contiguous()
        NB, ch, H, W = x.shape
        Cout = w1.shape[0]
        K = ch * 9
        fl = 2.0 * NB * H * W * Cout
        if _fmt() == ops.COMP_F16:
            _, colh = ops.im2col_k3s1(x.detach(), out_fmt=ops.COMP_F16)
            wp1 = ops.weight_matrix_f16(w1.detach().contiguous(), Cout, K, Cout, 32, K, 1)
            wps = ops.weight_matrix_f16(_center_tap(wsc.detach()), Cout, K, Cout, 32, K, 1)
            h1, h1c = ops.conv_fwd(colh, wp1, b1.detach(), ops.KIND_CONV_K1S1, H, W, ops.ACT_RELU, flops=fl * K,
                                   fp16_in=True, out_mode="pair")
            s, sc = ops.conv_fwd(colh, wps, bsc.detach(), ops.KIND_CONV_K1S1, H, W, flops=fl * ch, fp16_in=True,
                                 out_mode="pair")
            ctx.mark_non_differentiable(h1c, sc)
        else:
            col = ops.im2col_k3s1(x.detach())
            wp1 = ops.pack_matrix(w1.detach().contiguous(), Cout, K, Cout, 32, K, 1)
            wps = ops.pack_matrix(_center_tap(wsc.detach()), Cout, K, Cout, 32, K, 1)
            h1 = ops.conv_fwd(col, wp1, b1.detach(), ops.KIND_CONV_K1S1, H, W, ops.ACT_RELU, flops=fl * K)
            s = ops.conv_fwd(col, wps, bsc.detach(), ops.KIND_CONV_K1S1, H, W, flops=fl * ch)
            h1c = sc = None
        ctx.save_for_backward(x, w1, wsc, h1)
        return h1, h1c, s, sc

    @staticmethod
    def backward(ctx, dh1, _u1, ds, _u2):
        x, w1, wsc, h1 = ctx.saved_tensors
        if dh1 is None and ds is None:
            return (None,) * 5
        NB, ch, H, W = x.shape
        Cout, K = w1.shape[0], ch * 9
        # one of the two branches unused by the caller's graph: its gradient is zero
        dh1 = dh1 if dh1 is not None else torch.zeros_like(h1)
        ds = ds if ds is not None else torch.zeros_like(h1)
        dy1 = ops.act_bwd(dh1.contiguous(), h1, ops.ACT_RELU)
        ds = ds.contiguous()
        col = ops.im2col_k3s1(x)
        fl = 2.0 * NB * H * W * Cout
        dw1 = ops.unpack_matrix(ops.conv_wgrad(dy1, col, ops.KIND_CONV_K1S1, 1, flops=fl * K).view(Cout, 32), w1.shape,
                                Cout, K, 32, K, 1)
        dws_full = ops.unpack_matrix(ops.conv_wgrad(ds, col, ops.KIND_CONV_K1S1, 1, flops=fl * ch).view(Cout, 32),
                                     (Cout, ch, 3, 3), Cout, K, 32, K, 1)
        dwsc = dws_full[:, :, 1:2, 1:2].contiguous()
        db1, dbsc = ops.colsum(dy1), ops.colsum(ds)
        dx = None
        if ctx.needs_input_grad[0]:
            wt1 = ops.pack_matrix(w1.detach().contiguous(), K, Cout, 32, Cout, 1, K)
            wts = ops.pack_matrix(_center_tap(wsc.detach()), K, Cout, 32, Cout, 1, K)
            dcol = ops.conv_fwd(dy1, wt1, None, ops.KIND_CONV_K1S1, H, W, flops=fl * K)
            dcol = ops.conv_fwd(ds, wts, None, ops.KIND_CONV_K1S1, H, W, residual=dcol, flops=fl * ch)
            dx = ops.col2im_k3s1(dcol, ch)
        return dx, dw1, db1, dwsc, dbsc


class BlurPool(torch.autograd.Function):
    """BlurPool2d(filt_size=3, reflect, stride) on NHWC bf16 (models/ops.py:7-47; dcgan_blur.py:41 stride 1, :116 stride 2)."""

    @staticmethod
    def forward(ctx, x, x_comp, stride):
        ctx.set_materialize_grads(False)   # no zero tensors for the companion outputs
        ctx.dims = (x.shape[1], x.shape[2], stride)
        fmt = _fmt()
        if not fmt:
            return ops.blur3x3_fwd(x, stride), None
        out, out_comp = ops.blur3x3_fwd(x, stride, comp=x_comp, out_fmt=fmt)
        ctx.mark_non_differentiable(out_comp)
        return out, out_comp

    @staticmethod
    def backward(ctx, g, _unused=None):
        if g is None:
            return None, None, None
        H, W, stride = ctx.dims
        return ops.blur3x3_bwd(g.contiguous(), H, W, stride), None, None


class ImageConv3Act(torch.autograd.Function):
    """Conv2d(img_dim -> C, 3x3, p=1) + activation reading the fp32 NCHW image (first block of the dcgan_blur
    discriminator, models/dcgan_blur.py:111-115): im2col (K = 32) -> 1-tap tensor-core GEMM with fused bias + act."""

    @staticmethod
    def forward(ctx, x, weight, bias, act, cache, key):
        ctx.set_materialize_grads(False)   # no zero tensors for the companion outputs
        x = x.contiguous()
        NB, ch, H, W = x.shape
        Cout, K = weight.shape[0], ch * 9
        fl = 2.0 * NB * H * W * Cout * K
        if _fmt() == ops.COMP_F16:
            col, colh = ops.im2col_k3s1(x.detach(), out_fmt=ops.COMP_F16)      # col (bf16) is kept for wgrad
            wp = cache.get((key, "fwdh"), weight,
                           lambda: ops.weight_matrix_f16(weight.detach().contiguous(), Cout, K, Cout, 32, K, 1))
            a, a_comp = ops.conv_fwd(colh, wp, bias.detach(), ops.KIND_CONV_K1S1, H, W, act, flops=fl, fp16_in=True,
                                     out_mode="pair")
            ctx.mark_non_differentiable(a_comp)
        elif _fmt() == ops.COMP_LO:
            col, col_lo = ops.im2col_k3s1(x.detach(), out_fmt=ops.COMP_LO)
            wp = cache.get((key, "fwd3"), weight,
                           lambda: ops.split_weight_matrix(weight.detach().contiguous(), Cout, K, Cout, 32, K, 1))
            a, a_comp = ops.conv_fwd(col, wp, bias.detach(), ops.KIND_CONV_K1S1, H, W, act, flops=fl, x_lo=col_lo,
                                     out_mode="split")
            ctx.mark_non_differentiable(a_comp)
        else:
            col = ops.im2col_k3s1(x.detach())
            wp = cache.get((key, "fwd"), weight, lambda: ops.pack_matrix(weight.detach().contiguous(), Cout, K, Cout, 32, K, 1))
            a, a_comp = ops.conv_fwd(col, wp, bias.detach(), ops.KIND_CONV_K1S1, H, W, act, flops=fl), None
        ctx.save_for_backward(col, weight, a)
        ctx.misc = (act, cache, key, (NB, ch, H, W))
        return a, a_comp

    @staticmethod
    def backward(ctx, da, _unused=None):
        if da is None:
            return (None,) * 6
        col, weight, a = ctx.saved_tensors
        act, cache, key, (NB, ch, H, W) = ctx.misc
        Cout, K = weight.shape[0], ch * 9
        fl = 2.0 * NB * H * W * Cout * K
        dy = ops.act_bwd(da.contiguous(), a, act) if act != ops.ACT_NONE else da.contiguous()
        dweight = dbias = dx = None
        if ctx.needs_input_grad[1]:
            dweight = ops.unpack_matrix(ops.conv_wgrad(dy, col, ops.KIND_CONV_K1S1, 1, flops=fl).view(Cout, 32),
                                        weight.shape, Cout, K, 32, K, 1)
        if ctx.needs_input_grad[2]:
            dbias = ops.colsum(dy)
        if ctx.needs_input_grad[0]:
            wt = cache.get((key, "dgrad"), weight, lambda: ops.pack_matrix(weight.detach().contiguous(), K, Cout, 32, Cout, 1, K))
            dx = ops.col2im_k3s1(ops.conv_fwd(dy, wt, None, ops.KIND_CONV_K1S1, H, W, flops=fl), ch)
        return dx, dweight, dbias, None, None, None


class ImageOut3(torch.autograd.Function):
    """Last layer of the ResNet generator: tanh(Conv2d(ch -> img_dim, 3x3)) written as the fp32 NCHW image (:95).
    The GEMM runs with Nout padded to 8 (rows >= img_dim are zero)."""

    @staticmethod
    def forward(ctx, x, x_comp, weight, bias):
        NB, H, W, C = x.shape
        ch = weight.shape[0]
        w8 = torch.zeros((8,) + tuple(weight.shape[1:]), device=x.device, dtype=torch.float32)
        w8[:ch] = weight.detach()
        b8 = torch.zeros((8,), device=x.device, dtype=torch.float32)
        b8[:ch] = bias.detach()
        fl = 2.0 * NB * H * W * ch * C * 9
        if _fmt() == ops.COMP_F16:
            # the pre-tanh image stays fp32 between the GEMM and the tanh / layout pass
            y8 = ops.conv_fwd(_f16_operand(x, x_comp), ops.conv_weight_f16(w8, 0), b8, ops.KIND_CONV_K3S1, H, W, flops=fl,
                              fp16_in=True, out_mode="f32")
        elif _fmt() == ops.COMP_LO:
            y8 = ops.conv_fwd(x, ops.split_conv_weight(w8, 0), b8, ops.KIND_CONV_K3S1, H, W, flops=fl, x_lo=_lo(x, x_comp),
                              out_mode="f32")
        else:
            y8 = ops.conv_fwd(x, ops.pack_conv_weight(w8, 0), b8, ops.KIND_CONV_K3S1, H, W, flops=fl)
        out = ops.nhwc8_to_image(y8, ch, True)
        ctx.save_for_backward(x, w8, out)
        ctx.ch = ch
        return out

    @staticmethod
    def backward(ctx, dout):
        x, w8, out = ctx.saved_tensors
        ch = ctx.ch
        NB, H, W, C = x.shape
        dy8 = ops.image_to_nhwc8_grad(dout.contiguous(), out, True)
        fl = 2.0 * NB * H * W * ch * C * 9
        dw8 = ops.unpack_conv_wgrad(ops.conv_wgrad(dy8, x, ops.KIND_CONV_K3S1, 9, flops=fl), w8.shape)
        db8 = ops.colsum(dy8)
        dx = None
        if ctx.needs_input_grad[0]:
            wpd = ops.pack_conv_weight(w8, 1 | 2)
            dx = ops.conv_fwd(dy8, wpd, None, ops.KIND_CONV_K3S1, H, W, flops=fl)
        return dx, None, dw8[:ch].contiguous(), db8[:ch].contiguous()


class ProjHead(torch.autograd.Function):
    """relu -> sum over (H, W) -> l6(h) + sum_c l_y(y)_c * h_c (models/sngan_projection.py:190-195)."""

    @staticmethod
    def forward(ctx, a, a_comp, w6, b6, Ey, labels):
        h = ops.relu_sumpool(a, a_comp)
        e = Ey.detach().contiguous() if Ey is not None else None
        out = ops.proj_head_fwd(h, w6.detach().contiguous(), b6.detach(), e, labels)
        ctx.save_for_backward(a, h, w6, e if e is not None else torch.empty(0, device=a.device), labels
                              if labels is not None else torch.empty(0, device=a.device))
        ctx.has_proj = Ey is not None
        return out

    @staticmethod
    def backward(ctx, dout):
        a, h, w6, e, labels = ctx.saved_tensors
        E = e if ctx.has_proj else None
        lb = labels if ctx.has_proj else None
        dh, dw, db, dE = ops.proj_head_bwd(dout.contiguous(), h, w6.detach().contiguous(), E, lb,
                                           e.shape[0] if ctx.has_proj else 0)
        da = ops.relu_sumpool_bwd(dh, a)
        return da, None, dw.view_as(w6), db, dE, None
