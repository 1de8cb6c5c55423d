"""Thin torch-tensor wrappers over the C-ABI (include/gpb200.h). PyTorch is the host: it owns device memory and the
current stream; every function here only validates shapes/dtypes, allocates outputs with torch and passes raw pointers.
No computation happens in Python or in torch ops on this path."""
import ctypes

import torch

from . import _lib
from ._lib import ConvFwd, ConvWgrad, check

ACT_NONE, ACT_RELU, ACT_LRELU, ACT_TANH = 0, 1, 2, 3
KIND_CONV_K4S2, KIND_CONVT_K4S2, KIND_CONV_K3S1, KIND_CONV_K1S1 = 0, 1, 2, 3
LOSS_BCE, LOSS_MSE, LOSS_HINGE_REAL, LOSS_HINGE_FAKE, LOSS_NEG_MEAN = 0, 1, 2, 3, 4

_vp, _i, _ll, _f, _d = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_float, ctypes.c_double
_SIGS = {
    "gp_conv_fwd": [_vp, _vp],
    "gp_conv_wgrad": [_vp, _vp],
    "gp_pack_conv_weight": [_vp, _vp, _i, _i, _i, _i, _vp, _vp],
    "gp_unpack_conv_wgrad": [_vp, _vp, _i, _i, _i, _vp],
    "gp_pack_matrix": [_vp, _vp, _i, _i, _i, _i, _ll, _ll, _i, _vp, _vp],
    "gp_unpack_matrix": [_vp, _vp, _i, _i, _i, _ll, _ll, _i, _vp],
    "gp_bn_stats": [_vp, _ll, _i, _vp, _vp, _vp],
    "gp_bn_finalize": [_vp, _vp, _d, _i, _f, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "gp_bn_eval_params": [_vp, _vp, _vp, _vp, _i, _f, _vp, _vp, _vp, _vp, _vp],
    "gp_bn_apply_act": [_vp, _vp, _ll, _i, _vp, _vp, _i, _vp],
    "gp_bn_bwd_reduce": [_vp, _vp, _ll, _i, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp],
    "gp_bn_bwd_apply": [_vp, _vp, _vp, _ll, _i, _vp, _vp, _vp, _vp, _vp, _vp, _d, _i, _vp, _vp, _f, _vp],
    "gp_act_bwd": [_vp, _vp, _vp, _ll, _i, _vp],
    "gp_colsum": [_vp, _ll, _i, _vp, _vp],
    "gp_im2col_k4s2": [_vp, _vp, _vp, _i, _i, _i, _i, _vp],
    "gp_col2im_k4s2": [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp],
    "gp_image_bias_grad": [_vp, _vp, _vp, _i, _i, _i, _vp],
    "gp_head_fwd": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _ll, _ll, _ll, _vp],
    "gp_head_bwd": [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _ll, _ll, _ll, _vp],
    "gp_gan_loss": [_vp, _i, _i, _f, _vp, _vp, _vp],
    "gp_acgan_loss": [_vp, _vp, _i, _i, _i, _f, _f, _vp, _vp, _vp],
    "gp_sn_sigma": [_vp, _i, _i, _i, _i, _vp, _vp, _f, _i, _vp, _vp, _vp],
    "gp_sn_scale": [_vp, _vp, _vp, _ll, _vp],
    "gp_sn_grad": [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp],
}
_bound = {}


class CallProfiler:
    """Times EVERY C-ABI call with CUDA events on the launching stream (tools/step_breakdown.py): records
    (entry point, raw argument tuple, start event, end event). Eager steps only — never active inside a graph capture."""

    active = None
    note = None   # (label, algorithmic FLOPs) of the next GEMM call, set by conv_fwd / conv_wgrad

    def __init__(self):
        self.records = []

    def __enter__(self):
        CallProfiler.active = self
        return self

    def __exit__(self, *exc):
        CallProfiler.active = None


def _fn(name):
    f = _bound.get(name)
    if f is None:
        f = getattr(_lib.lib(), name)
        f.argtypes = _SIGS[name]
        f.restype = ctypes.c_int
        _bound[name] = f
    prof = CallProfiler.active
    if prof is None:
        return f

    def timed(*args):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        rc = f(*args)
        b.record()
        prof.records.append((name, args, a, b, CallProfiler.note))
        CallProfiler.note = None
        return rc

    return timed


def exported_symbols():
    """Names of every compute entry point declared in include/gpb200.h (used by the CPU-side ABI test)."""
    return sorted(_SIGS) + ["gp_version", "gp_last_error", "gp_launch_count", "gp_peer_buffer_bytes", "gp_conv_fwd_plan"]


def _p(t):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _chk(t, dtype, name):
    if not t.is_cuda:
        raise _lib.GpError("%s must be a CUDA tensor (the hot path has no CPU fallback)" % name)
    if t.dtype != dtype:
        raise _lib.GpError("%s must be %s, got %s" % (name, dtype, t.dtype))
    if not t.is_contiguous():
        raise _lib.GpError("%s must be contiguous" % name)


# ------------------------------------------------------------------------------------------------ zeroed workspaces
class ZeroArena:
    """One memset per training step instead of ~100 small fills: the zero-initialised fp32 workspaces the kernels
    accumulate into (split-K weight gradients, BatchNorm partial sums, bias-gradient sums) are carved out of one buffer
    that `begin()` clears at the start of a step. Slices are valid until the next `begin()`, which every consumer on the
    training path satisfies (the sums are finalised / unpacked / accumulated into .grad within the same step).
    The three step drivers of engine.py (DcganStep / SnganStep / AcganStep) activate it — their parameters always own a
    .grad view (FusedAdam / GradBucket), so autograd never adopts a slice as .grad; everywhere else `zeros()` is
    torch.zeros."""

    active = None

    def __init__(self, device):
        self.dev, self.buf, self.off, self.need = device, None, 0, 0

    def begin(self):
        if self.buf is None or self.need > self.buf.numel():
            if self.need:
                self.buf = torch.zeros(int(self.need * 1.05) + 1024, device=self.dev, dtype=torch.float32)
        elif self.off:
            self.buf[:self.off].zero_()
        self.off, self.need = 0, 0

    def take(self, numel):
        n = (numel + 63) // 64 * 64          # 256-byte granules keep every slice aligned for 16-byte vector atomics
        self.need += n
        if self.buf is None or self.off + n > self.buf.numel():
            return None
        t = self.buf[self.off:self.off + numel]
        self.off += n
        return t


def zeros(shape, device):
    """fp32 zeros: a slice of the active ZeroArena when there is one, else torch.zeros."""
    a = ZeroArena.active
    if a is not None:
        n = 1
        for d in shape:
            n *= d
        t = a.take(n)
        if t is not None:
            return t.view(shape)
    return torch.zeros(shape, device=device, dtype=torch.float32)


# ------------------------------------------------------------------------------------------------ conv GEMMs
class GemmProfiler:
    """Times every tensor-core GEMM launch (gp::conv_gemm_kernel) with CUDA events on the launching stream and
    accumulates its ALGORITHMIC FLOPs — bench.py's roofline numbers come from here (measured live, no profiler).
    The fused image-edge launches (csrc/image_edge.cu) are HBM-bound: they are recorded with their algorithmic BYTES and
    summarised separately (`summary(hbm=True)`), not against the tensor-core peak."""

    active = None

    def __init__(self):
        self.records = []

    def __enter__(self):
        GemmProfiler.active = self
        return self

    def __exit__(self, *exc):
        GemmProfiler.active = None

    def summary(self, hbm=False):
        torch.cuda.synchronize()
        recs = [r for r in self.records if (r[4] is not None) == hbm]
        ms = sum(a.elapsed_time(b) for a, b, _, _, _ in recs)
        fl = sum(f for _, _, f, _, _ in recs)
        out = {"ms": ms, "flops": fl, "tflops": fl / (ms * 1e-3) / 1e12 if ms > 0 else 0.0, "launches": len(recs)}
        if hbm:
            by = sum(nb for _, _, _, _, nb in recs)
            out.update(bytes=by, gbs=by / (ms * 1e-3) / 1e9 if ms > 0 else 0.0)
        return out

    def per_launch(self):
        torch.cuda.synchronize()
        return [(name, a.elapsed_time(b), f) for a, b, f, name, _ in self.records]


def _timed(name, flops, fn, nbytes=None):
    if CallProfiler.active is not None:
        CallProfiler.note = (name, flops)
    prof = GemmProfiler.active
    if prof is None:
        return fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    r = fn()
    b.record()
    prof.records.append((a, b, flops, name, nbytes))
    return r


_TAPS_PER_OUT = {KIND_CONV_K4S2: 16, KIND_CONVT_K4S2: 4, KIND_CONV_K3S1: 9, KIND_CONV_K1S1: 1}


CONV_IN_F16, CONV_LO_F16, CONV_RES_F16, CONV_OUT_F16 = 1, 2, 4, 8


CONV_BWD_MASK, CONV_BWD_BN_F32, CONV_BWD_BN_BF16 = 1, 2, 3
MAX_STAT_COLS = 2048     # kMaxStatCols of csrc/conv_gemm.cuh: widest output with epilogue-accumulated column sums
_ACT_SLOPE = {ACT_RELU: 0.0, ACT_LRELU: 0.2}


def conv_fwd(x, wp, bias, kind, Hout, Wout, act=ACT_NONE, stats=None, flops=None, residual=None, x_lo=None,
             out_mode="bf16", fp16_in=False, bwd=None):
    """x: bf16 (NB, Hin, Win, Cin); wp: bf16 [Nout, taps*Cin]; returns bf16 (NB, Hout, Wout, Nout).
    stats: optional fp32 [2, Nout] zeroed tensor receiving per-channel sum / sum of squares.
    flops: algorithmic FLOPs of this call when they differ from the padded GEMM shape (image layers).
    bf16x3 mode: x_lo = low halves of x and wp = [Nout, 2*taps*Cin] (hi | lo).
    fp16 mode: fp16_in=True, x and wp are fp16 (one MMA; same layouts as the bf16 operands, x_lo must be None).
    out_mode: "bf16" -> bf16 tensor; "split" -> (hi, lo) bf16 pair; "f32" -> fp32 tensor; "pair" -> (bf16, fp16) copies
    of the same value (backward operand, next forward operand of the fp16 mode); "f16" -> one fp16 tensor (2-byte
    pre-BatchNorm storage of the fp16 mode).
    bwd: this GEMM is the data gradient of the next layer and does backward work of the PRODUCING layer in its epilogue
    (gp_conv_fwd_t::bwd_mode): ("mask", a, act) multiplies the result by act'(a); ("bn", y, fin, act, red) accumulates
    the BatchNorm-backward sums of the producer into the zeroed fp32 [2, Nout] tensor red."""
    op_dtype = torch.float16 if fp16_in else torch.bfloat16
    _chk(x, op_dtype, "x")
    _chk(wp, op_dtype, "wp")
    if fp16_in and x_lo is not None:
        raise _lib.GpError("conv_fwd: fp16 operands run as one MMA (x_lo must be None)")
    NB, Hin, Win, Cin = x.shape
    Nout = wp.shape[0]
    shape = (NB, Hout, Wout, Nout)
    out = out_lo = out_f32 = None
    if out_mode == "f32":
        out_f32 = torch.empty(shape, device=x.device, dtype=torch.float32)
    else:
        out = torch.empty(shape, device=x.device, dtype=torch.float16 if out_mode == "f16" else torch.bfloat16)
        if out_mode == "split":
            out_lo = torch.empty(shape, device=x.device, dtype=torch.bfloat16)
        elif out_mode == "pair":
            out_lo = torch.empty(shape, device=x.device, dtype=torch.float16)
    flags = (CONV_IN_F16 if fp16_in else 0) | (CONV_LO_F16 if out_mode == "pair" else 0) | \
        (CONV_OUT_F16 if out_mode == "f16" else 0)
    if residual is not None and residual.dtype == torch.float16:   # the fp16 companion of the shortcut activation
        flags |= CONV_RES_F16
    if bias is not None:
        _chk(bias, torch.float32, "bias")
    if x_lo is not None:
        _chk(x_lo, torch.bfloat16, "x_lo")
    bwd_src = bwd_fin = None
    bwd_slope, bwd_mode = 0.0, 0
    if bwd is not None:
        if bwd[0] == "mask":
            _, bwd_src, bact = bwd
            _chk(bwd_src, torch.bfloat16, "bwd mask source")
            bwd_mode = CONV_BWD_MASK
        else:
            _, bwd_src, bwd_fin, bact, stats = bwd
            _chk(bwd_fin, torch.float32, "bwd fin")
            _chk(stats, torch.float32, "bwd red")
            if bwd_src.dtype not in (torch.float32, torch.bfloat16) or not bwd_src.is_contiguous():
                raise _lib.GpError("conv_fwd: the fused BatchNorm-backward reduction reads a contiguous fp32 / bf16 y")
            bwd_mode = CONV_BWD_BN_F32 if bwd_src.dtype == torch.float32 else CONV_BWD_BN_BF16
        if tuple(bwd_src.shape) != shape:
            raise _lib.GpError("conv_fwd: bwd source %s does not have the output's shape %s" % (tuple(bwd_src.shape), shape))
        bwd_slope = _ACT_SLOPE[bact]
    p = ConvFwd(_p(x), _p(wp), _p(bias), _p(out), _p(stats[0]) if stats is not None else None,
                _p(stats[1]) if stats is not None else None, NB, Hin, Win, Cin, Hout, Wout, Nout, kind, act, _p(residual),
                _p(x_lo), _p(out_lo), _p(out_f32), flags, _p(bwd_src), _p(bwd_fin), bwd_slope, bwd_mode)
    if flops is None:
        flops = 2.0 * NB * Hout * Wout * Nout * Cin * _TAPS_PER_OUT[kind]
    _timed("conv_fwd kind%d %dx%dx%d C%d->%d%s" % (kind, NB, Hout, Wout, Cin, Nout,
                                                    " x3" if x_lo is not None else (" fp16" if fp16_in else "")),
           flops, lambda: check(_fn("gp_conv_fwd")(ctypes.addressof(p), _stream()), "gp_conv_fwd"))
    if out_mode == "f32":
        return out_f32
    if out_mode in ("split", "pair"):
        return out, out_lo
    return out


def conv_fwd_plan(NB, Hin, Win, Cin, Hout, Wout, Nout, kind, x3=False):
    """(BN, MT, tiles) gp_conv_fwd would choose for this shape (host-side cost model; runs without a GPU)."""
    p = ConvFwd(None, None, None, None, None, None, NB, Hin, Win, Cin, Hout, Wout, Nout, kind, 0, None,
                1 if x3 else None, None, None)
    bn, mt, tiles = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    f = _lib.lib().gp_conv_fwd_plan
    f.argtypes = [_vp, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]
    f.restype = ctypes.c_int
    check(f(ctypes.addressof(p), ctypes.byref(bn), ctypes.byref(mt), ctypes.byref(tiles)), "gp_conv_fwd_plan")
    return bn.value, mt.value, tiles.value


def conv_wgrad(dense, gath, kind, taps, flops=None):
    """dense: bf16 (NB, Hs, Ws, Cd); gath: bf16 (NB, Hg, Wg, Cg); returns fp32 [Cd, taps, Cg]."""
    _chk(dense, torch.bfloat16, "dense")
    _chk(gath, torch.bfloat16, "gath")
    NB, Hs, Ws, Cd = dense.shape
    _, Hg, Wg, Cg = gath.shape
    dw = zeros((Cd, taps, Cg), dense.device)
    p = ConvWgrad(_p(dense), _p(gath), _p(dw), NB, Hs, Ws, Cd, Hg, Wg, Cg, kind)
    if flops is None:
        flops = 2.0 * NB * Hs * Ws * Cd * Cg * taps
    _timed("conv_wgrad kind%d %dx%dx%d Cd%d Cg%d" % (kind, NB, Hs, Ws, Cd, Cg), flops,
           lambda: check(_fn("gp_conv_wgrad")(ctypes.addressof(p), _stream()), "gp_conv_wgrad"))
    return dw


# ------------------------------------------------------------------------------------------------ weight staging
def pack_conv_weight(w, n_dim, inv_scale=None):
    """w: fp32 (D0, D1, kh, kw) -> bf16 [N, taps*C], (N, C) = (D0, D1) if n_dim == 0 else (D1, D0)."""
    _chk(w, torch.float32, "w")
    D0, D1 = w.shape[0], w.shape[1]
    taps = w.shape[2] * w.shape[3]
    N, C = (D0, D1) if n_dim == 0 else (D1, D0)
    dst = torch.empty((N, taps * C), device=w.device, dtype=torch.bfloat16)
    check(_fn("gp_pack_conv_weight")(_p(w), _p(dst), D0, D1, taps, n_dim, _p(inv_scale), _stream()), "gp_pack_conv_weight")
    return dst


# ---- batched staging of 4x4 conv weights (csrc/weight_stage.cu)
STAGE_BF16, STAGE_F16, STAGE_SPLIT = 0, 1, 2
STAGE_MAX_LAYERS, STAGE_MAX_DST = 12, 4
_SIGS.update({"gp_stage_conv_weights": [_vp, _vp]})


class StageDst(ctypes.Structure):
    """gp_stage_dst_t of include/gpb200.h."""
    _fields_ = [("ptr", ctypes.c_void_p), ("ld", ctypes.c_longlong), ("n_dim", ctypes.c_int32), ("fmt", ctypes.c_int32)]


class StageLayer(ctypes.Structure):
    """gp_stage_layer_t of include/gpb200.h."""
    _fields_ = [("src", ctypes.c_void_p), ("D0", ctypes.c_int32), ("D1", ctypes.c_int32), ("tile0", ctypes.c_int32),
                ("ndst", ctypes.c_int32), ("dst", StageDst * STAGE_MAX_DST)]


class StageTable(ctypes.Structure):
    """gp_stage_table_t of include/gpb200.h."""
    _fields_ = [("count", ctypes.c_int32), ("total_tiles", ctypes.c_int32), ("layer", StageLayer * STAGE_MAX_LAYERS)]


def stage_conv_ok(w):
    """Weights the batched staging kernel takes: fp32 CUDA (D0, D1, 4, 4) with D0, D1 multiples of 32."""
    return (w.is_cuda and w.dtype == torch.float32 and w.dim() == 4 and w.shape[2] * w.shape[3] == 16 and w.is_contiguous()
            and w.shape[0] % 32 == 0 and w.shape[1] % 32 == 0 and w.data_ptr() % 16 == 0)


def stage_conv_weights(items):
    """items: [(w fp32 (D0, D1, 4, 4), [(fmt, n_dim), ...]), ...] -> [[operand tensor per request], ...]: rows [N][tap][C] as
    pack_conv_weight (STAGE_BF16), conv_weight_f16 (STAGE_F16) or split_conv_weight (STAGE_SPLIT: hi | lo) write them, all
    layers and formats from one launch per STAGE_MAX_LAYERS layers."""
    outs = []
    for i0 in range(0, len(items), STAGE_MAX_LAYERS):
        chunk = items[i0:i0 + STAGE_MAX_LAYERS]
        tb = StageTable()
        tb.count = len(chunk)
        for li, (w, reqs) in enumerate(chunk):
            _chk(w, torch.float32, "w")
            if not stage_conv_ok(w) or not 0 < len(reqs) <= STAGE_MAX_DST:
                raise _lib.GpError("stage_conv_weights: weight %s with %d requests is outside the batched kernel's set"
                                   % (tuple(w.shape), len(reqs)))
            D0, D1 = w.shape[0], w.shape[1]
            L = tb.layer[li]
            L.src, L.D0, L.D1, L.ndst = _p(w), D0, D1, len(reqs)
            res = []
            for di, (fmt, n_dim) in enumerate(reqs):
                N, C = (D0, D1) if n_dim == 0 else (D1, D0)
                ld = (32 if fmt == STAGE_SPLIT else 16) * C
                t = torch.empty((N, ld), device=w.device, dtype=torch.float16 if fmt == STAGE_F16 else torch.bfloat16)
                L.dst[di].ptr, L.dst[di].ld, L.dst[di].n_dim, L.dst[di].fmt = _p(t), ld, n_dim, fmt
                res.append(t)
            outs.append(res)
        check(_fn("gp_stage_conv_weights")(ctypes.addressof(tb), _stream()), "gp_stage_conv_weights")
    return outs


UNPACK_ACCUMULATE = 1 << 30


def grad_target(param):
    """The parameter's existing .grad buffer when gradients can be delivered straight into it (fp32, contiguous, a leaf
    that already owns a .grad — the step drivers' flat gradient buffers), else None. The node then returns no tensor
    for that parameter, so autograd runs no AccumulateGrad pass: `grad += new` happens inside the producing kernel
    (gradients of the D-real and D-fake passes accumulate, main_dcgan.py:73,84)."""
    if param is None or not isinstance(param, torch.nn.Parameter) or not param.is_leaf:
        return None
    g = param.grad
    if g is None or g.dtype != torch.float32 or not g.is_contiguous() or g.shape != param.shape or not g.is_cuda == param.is_cuda:
        return None
    return g


def unpack_conv_wgrad(dwp, shape, into=None):
    """dwp: fp32 [M, taps, N] -> fp32 tensor of `shape` = (M, N, kh, kw); into: accumulate into this tensor instead."""
    _chk(dwp, torch.float32, "dwp")
    M, taps, N = dwp.shape
    dst = torch.empty(shape, device=dwp.device, dtype=torch.float32) if into is None else into
    check(_fn("gp_unpack_conv_wgrad")(_p(dwp), _p(dst), M, N, taps | (UNPACK_ACCUMULATE if into is not None else 0),
                                      _stream()), "gp_unpack_conv_wgrad")
    return dst


def pack_matrix(src, R, K, Rpad, ld, s_r, s_k, perm=1, inv_scale=None):
    """Generic fp32 -> bf16 [Rpad, ld] staging with strides, zero padding and optional NCHW->NHWC row permutation."""
    _chk(src, torch.float32, "src")
    dst = torch.empty((Rpad, ld), device=src.device, dtype=torch.bfloat16)
    check(_fn("gp_pack_matrix")(_p(src), _p(dst), R, K, Rpad, ld, s_r, s_k, perm, _p(inv_scale), _stream()), "gp_pack_matrix")
    return dst


def unpack_matrix(src, out_shape, R, K, ld_src, s_r, s_k, perm=1, into=None):
    _chk(src, torch.float32, "src")
    dst = torch.empty(out_shape, device=src.device, dtype=torch.float32) if into is None else into   # every element is written
    check(_fn("gp_unpack_matrix")(_p(src), _p(dst), R, K, ld_src, s_r, s_k, perm | (UNPACK_ACCUMULATE if into is not None else 0),
                                  _stream()), "gp_unpack_matrix")
    return dst


# ------------------------------------------------------------------------------------------------ BatchNorm
def bn_stats(y):
    """y: bf16 (..., C) -> fp32 [2, C] (sum, sum of squares)."""
    _chk(y, torch.bfloat16, "y")
    C = y.shape[-1]
    P = y.numel() // C
    st = zeros((2, C), y.device)
    check(_fn("gp_bn_stats")(_p(y), P, C, _p(st[0]), _p(st[1]), _stream()), "gp_bn_stats")
    return st


def bn_finalize(st, count, gamma, beta, running_mean, running_var, nbt, eps=1e-5, momentum=0.1):
    """Returns fp32 [4, C]: mean, rstd, scale, shift. Updates running buffers in place when given."""
    C = st.shape[1]
    out = torch.empty((4, C), device=st.device, dtype=torch.float32)
    check(_fn("gp_bn_finalize")(_p(st[0]), _p(st[1]), float(count), C, eps, momentum, _p(gamma), _p(beta), _p(out[0]),
                                _p(out[1]), _p(out[2]), _p(out[3]), _p(running_mean), _p(running_var), _p(nbt), _stream()),
          "gp_bn_finalize")
    return out


def bn_eval_params(running_mean, running_var, gamma, beta, eps=1e-5):
    """Eval-mode BatchNorm: fp32 [4, C] (mean, rstd, scale, shift) from the running statistics."""
    C = running_mean.shape[0]
    out = torch.empty((4, C), device=running_mean.device, dtype=torch.float32)
    check(_fn("gp_bn_eval_params")(_p(running_mean), _p(running_var), _p(gamma), _p(beta), C, eps, _p(out[0]), _p(out[1]),
                                   _p(out[2]), _p(out[3]), _stream()), "gp_bn_eval_params")
    return out


def bn_apply_act(y, fin, act):
    _chk(y, torch.bfloat16, "y")
    C = y.shape[-1]
    out = torch.empty_like(y)
    check(_fn("gp_bn_apply_act")(_p(y), _p(out), y.numel() // C, C, _p(fin[2]), _p(fin[3]), act, _stream()), "gp_bn_apply_act")
    return out


def bn_bwd_reduce(da, y, fin, act):
    _chk(da, torch.bfloat16, "da")
    _chk(y, torch.bfloat16, "y")
    C = y.shape[-1]
    red = zeros((2, C), y.device)
    check(_fn("gp_bn_bwd_reduce")(_p(da), _p(y), y.numel() // C, C, _p(fin[2]), _p(fin[3]), _p(fin[0]), _p(fin[1]), act,
                                  _p(red[0]), _p(red[1]), _stream()), "gp_bn_bwd_reduce")
    return red


def bn_bwd_apply(da, y, fin, red, count, act, acc=None, acc_scale=1.0):
    """acc = (dbeta_grad_buffer, dgamma_grad_buffer): the affine gradients are added into them by the same launch."""
    C = y.shape[-1]
    dy = torch.empty_like(y)
    check(_fn("gp_bn_bwd_apply")(_p(da), _p(y), _p(dy), y.numel() // C, C, _p(fin[2]), _p(fin[3]), _p(fin[0]), _p(fin[1]),
                                 _p(red[0]), _p(red[1]), float(count), act, _p(acc[0]) if acc else None,
                                 _p(acc[1]) if acc else None, float(acc_scale), _stream()), "gp_bn_bwd_apply")
    return dy


def act_bwd(da, a, act):
    _chk(da, torch.bfloat16, "da")
    _chk(a, torch.bfloat16, "a")
    dy = torch.empty_like(a)
    check(_fn("gp_act_bwd")(_p(da), _p(a), _p(dy), a.numel(), act, _stream()), "gp_act_bwd")
    return dy


def colsum(x, into=None):
    """out[c] (+)= sum over rows; into: an existing fp32 [C] buffer to accumulate into (a bias' .grad)."""
    _chk(x, torch.bfloat16, "x")
    C = x.shape[-1]
    out = zeros((C,), x.device) if into is None else into
    check(_fn("gp_colsum")(_p(x), x.numel() // C, C, _p(out), _stream()), "gp_colsum")
    return out


# ------------------------------------------------------------------------------------------------ image layers
def im2col_k4s2(img, mul=None):
    """img: fp32 NCHW (NB, ch, Hi, Wi) -> bf16 (NB, Hi/2, Wi/2, 64)."""
    _chk(img, torch.float32, "img")
    NB, ch, Hi, Wi = img.shape
    col = torch.empty((NB, Hi // 2, Wi // 2, 64), device=img.device, dtype=torch.bfloat16)
    if mul is not None:
        _chk(mul, torch.float32, "mul")
    check(_fn("gp_im2col_k4s2")(_p(img), _p(mul), _p(col), NB, ch, Hi, Wi, _stream()), "gp_im2col_k4s2")
    return col


def col2im_k4s2(col, bias, ch, act):
    """col: bf16 (NB, Ho, Wo, 64) -> fp32 NCHW (NB, ch, 2Ho, 2Wo)."""
    _chk(col, torch.bfloat16, "col")
    NB, Ho, Wo, _ = col.shape
    img = torch.empty((NB, ch, 2 * Ho, 2 * Wo), device=col.device, dtype=torch.float32)
    check(_fn("gp_col2im_k4s2")(_p(col), _p(bias), _p(img), NB, ch, 2 * Ho, 2 * Wo, act, _stream()), "gp_col2im_k4s2")
    return img


def image_bias_grad(dout, mul=None, into=None):
    _chk(dout, torch.float32, "dout")
    NB, ch, H, W = dout.shape
    db = zeros((ch,), dout.device) if into is None else into
    check(_fn("gp_image_bias_grad")(_p(dout), _p(mul), _p(db), NB, ch, H * W, _stream()), "gp_image_bias_grad")
    return db


# ------------------------------------------------------------------------------------------------ heads / losses
def head_fwd(a, w, bias, O, s_o, s_c, s_hw):
    """a: bf16 (NB, H, W, C); returns fp32 (NB, O)."""
    _chk(a, torch.bfloat16, "a")
    _chk(w, torch.float32, "w")
    NB, H, W, C = a.shape
    out = torch.empty((NB, O), device=a.device, dtype=torch.float32)
    check(_fn("gp_head_fwd")(_p(a), _p(w), _p(bias), _p(out), NB, H * W, C, O, s_o, s_c, s_hw, _stream()), "gp_head_fwd")
    return out


def head_bwd(dout, a, w, O, s_o, s_c, s_hw, need_da=True, need_dw=True, need_db=True):
    _chk(dout, torch.float32, "dout")
    NB, H, W, C = a.shape
    da = torch.empty_like(a) if need_da else None
    dw = zeros(tuple(w.shape), w.device) if need_dw else None
    db = torch.empty((O,), device=a.device, dtype=torch.float32) if (need_db and need_dw) else None
    check(_fn("gp_head_bwd")(_p(dout), _p(a), _p(w), _p(da), _p(dw), _p(db), NB, H * W, C, O, s_o, s_c, s_hw, _stream()),
          "gp_head_bwd")
    return da, dw, db


def gan_loss(pred, mode, target):
    """pred: fp32 (n, ...) -> (loss scalar fp32 tensor, dpred fp32 like pred) in one kernel."""
    _chk(pred, torch.float32, "pred")
    loss = torch.empty((), device=pred.device, dtype=torch.float32)
    dpred = torch.empty_like(pred)
    check(_fn("gp_gan_loss")(_p(pred), pred.numel(), mode, float(target), _p(loss), _p(dpred), _stream()), "gp_gan_loss")
    return loss, dpred


def acgan_loss(logits, labels, mode, target, aux_weight):
    """logits: fp32 (NB, 1 + K) packed two-head output, labels: fp32 (NB, K) -> (out4, dlogits): out4 = [adversarial
    term, auxiliary MSE term, adv + aux_weight * aux, mean sigmoid(adv)], dlogits = d out4[2] / d logits. One kernel."""
    _chk(logits, torch.float32, "logits")
    _chk(labels, torch.float32, "labels")
    NB, K = labels.shape
    if tuple(logits.shape) != (NB, K + 1):
        raise _lib.GpError("acgan_loss: logits %s do not match labels %s (+1 adversarial column)"
                           % (tuple(logits.shape), tuple(labels.shape)))
    out4 = torch.empty((4,), device=logits.device, dtype=torch.float32)
    dlogits = torch.empty_like(logits)
    check(_fn("gp_acgan_loss")(_p(logits), _p(labels), NB, K, mode, float(target), float(aux_weight), _p(out4), _p(dlogits),
                               _stream()), "gp_acgan_loss")
    return out4, dlogits


# ------------------------------------------------------------------------------------------------ spectral norm
def _sn_dims(w):
    A, B = w.shape[0], (w.shape[1] if w.dim() > 1 else 1)
    T = w.numel() // (A * B)
    return A, B, T


def sn_sigma(w, u, v, dim, training, eps=1e-12):
    """One power iteration (in place on u, v when training) and sigma; returns the fp32 scalar tensor sigma."""
    _chk(w, torch.float32, "w")
    _chk(u, torch.float32, "u")
    _chk(v, torch.float32, "v")
    A, B, T = _sn_dims(w)
    scratch = torch.empty((u.numel() + v.numel(),), device=w.device, dtype=torch.float32)
    sigma = torch.empty((), device=w.device, dtype=torch.float32)
    check(_fn("gp_sn_sigma")(_p(w), A, B, T, dim, _p(u), _p(v), eps, 1 if training else 0, _p(scratch), _p(sigma), _stream()),
          "gp_sn_sigma")
    return sigma


def sn_scale(w, sigma):
    out = torch.empty_like(w)
    check(_fn("gp_sn_scale")(_p(w), _p(sigma), _p(out), w.numel(), _stream()), "gp_sn_scale")
    return out


def sn_grad(g, w_sn, dim, u, v, sigma):
    _chk(g, torch.float32, "g")
    A, B, T = _sn_dims(w_sn)
    dot = torch.empty((), device=g.device, dtype=torch.float32)
    out = torch.empty_like(w_sn)
    check(_fn("gp_sn_grad")(_p(g), _p(w_sn), A, B, T, dim, _p(u), _p(v), _p(sigma), _p(dot), _p(out), _stream()), "gp_sn_grad")
    return out


SN_MAX = 24
_SIGS.update({"gp_sn_batched": [_vp, _vp], "gp_sn_grad_batched": [_vp, _vp]})


class SnBatch(ctypes.Structure):
    """gp_sn_batch_t of include/gpb200.h."""
    _fields_ = [(n, ctypes.c_void_p * SN_MAX) for n in ("w", "u", "v", "out", "sigma")] + \
               [(n, ctypes.c_int32 * SN_MAX) for n in ("A", "B", "T", "dim")] + \
               [("count", ctypes.c_int32), ("training", ctypes.c_int32), ("eps", ctypes.c_float),
                ("scratch", ctypes.c_void_p), ("keep", ctypes.c_void_p)]


class SnGradBatch(ctypes.Structure):
    """gp_sn_grad_batch_t of include/gpb200.h."""
    _fields_ = [(n, ctypes.c_void_p * SN_MAX) for n in ("g", "w_sn", "u", "v", "sigma", "out")] + \
               [(n, ctypes.c_int32 * SN_MAX) for n in ("A", "B", "T", "dim")] + \
               [("count", ctypes.c_int32), ("dot", ctypes.c_void_p)]


def sn_batched(ws, us, vs, dims, training, eps=1e-12):
    """Spectral norm of every hook of one forward in five launches. ws / us / vs: lists of fp32 tensors (weight_orig, u,
    v), dims: 0 | 1 per hook. u, v are advanced in place when training. Returns (w_sn list — views into one flat buffer,
    sigma fp32 [n], keep fp32 flat buffer of the (u | v) used, offsets of each hook inside keep)."""
    n = len(ws)
    if not 0 < n <= SN_MAX:
        raise _lib.GpError("sn_batched: 1..%d hooks per call, got %d" % (SN_MAX, n))
    dev = ws[0].device
    numels = [w.numel() for w in ws]
    flat = torch.empty(sum((m + 3) // 4 * 4 for m in numels), device=dev, dtype=torch.float32)
    sigma = torch.empty(n, device=dev, dtype=torch.float32)
    b = SnBatch()
    outs, offs, off, pos = [], [], 0, 0
    for i, (w, u, v, d) in enumerate(zip(ws, us, vs, dims)):
        _chk(w, torch.float32, "w")
        _chk(u, torch.float32, "u")
        _chk(v, torch.float32, "v")
        A, B, T = _sn_dims(w)
        o = flat[pos:pos + numels[i]].view(w.shape)
        pos += (numels[i] + 3) // 4 * 4
        outs.append(o)
        b.w[i], b.u[i], b.v[i], b.out[i], b.sigma[i] = _p(w), _p(u), _p(v), _p(o), sigma.data_ptr() + 4 * i
        b.A[i], b.B[i], b.T[i], b.dim[i] = A, B, T, d
        offs.append((off, u.numel(), v.numel()))
        off += u.numel() + v.numel()
    scratch = torch.empty(off, device=dev, dtype=torch.float32)
    keep = torch.empty(off, device=dev, dtype=torch.float32)
    b.count, b.training, b.eps, b.scratch, b.keep = n, 1 if training else 0, eps, _p(scratch), _p(keep)
    check(_fn("gp_sn_batched")(ctypes.addressof(b), _stream()), "gp_sn_batched")
    return outs, sigma, keep, offs


def sn_grad_batched(gs, w_sns, sigma, keep, offs, dims):
    """Gradient through W / sigma for every hook that received one (gs[i] may be None). Returns a list (None where gs[i] is)."""
    n = len(gs)
    dev = sigma.device
    b = SnGradBatch()
    outs = []
    for i, g in enumerate(gs):
        if g is None:
            outs.append(None)
            continue
        _chk(g, torch.float32, "g")
        A, B, T = _sn_dims(w_sns[i])
        o = torch.empty_like(w_sns[i])
        off, nu, nv = offs[i]
        b.g[i], b.w_sn[i], b.out[i] = _p(g), _p(w_sns[i]), _p(o)
        b.u[i], b.v[i], b.sigma[i] = keep.data_ptr() + 4 * off, keep.data_ptr() + 4 * (off + nu), sigma.data_ptr() + 4 * i
        b.A[i], b.B[i], b.T[i], b.dim[i] = A, B, T, dims[i]
        outs.append(o)
    dot = torch.empty(n, device=dev, dtype=torch.float32)
    b.count, b.dot = n, _p(dot)
    check(_fn("gp_sn_grad_batched")(ctypes.addressof(b), _stream()), "gp_sn_grad_batched")
    return outs


# ------------------------------------------------------------------------------------------------ SNGAN projection
_SIGS.update({
    "gp_cbn_apply_act": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _i, _i, _vp],
    "gp_cbn_bwd_reduce": [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _i, _vp],
    "gp_cbn_bwd_apply": [_vp, _vp, _vp, _i, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _d, _i, _i, _vp],
    "gp_upsample2x": [_vp, _vp, _i, _i, _i, _i, _f, _vp],
    "gp_pool2x": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _vp],
    "gp_act_fwd": [_vp, _vp, _vp, _vp, _i, _ll, _i, _vp],
    "gp_im2col_k3s1": [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp],
    "gp_col2im_k3s1": [_vp, _vp, _i, _i, _i, _i, _vp],
    "gp_nhwc8_to_image": [_vp, _i, _vp, _i, _i, _i, _i, _vp],
    "gp_image_to_nhwc8_grad": [_vp, _vp, _vp, _i, _i, _i, _i, _vp],
    "gp_relu_sumpool": [_vp, _vp, _i, _vp, _i, _i, _i, _vp],
    "gp_relu_sumpool_bwd": [_vp, _vp, _vp, _i, _i, _i, _vp],
    "gp_proj_head_fwd": [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp],
    "gp_proj_head_bwd": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp],
})


COMP_NONE, COMP_LO, COMP_F16 = 0, 1, 2
_COMP_DTYPE = {COMP_LO: torch.bfloat16, COMP_F16: torch.float16}


def comp_fmt_of(comp):
    """Companion format of a companion tensor (act_io.cuh): fp16 = copy of the value, bf16 = low half, None = none."""
    if comp is None:
        return COMP_NONE
    return COMP_F16 if comp.dtype == torch.float16 else COMP_LO


def _comp_io(comp, out_fmt, shape, device):
    """(input companion to pass, kernel format, output companion tensor or None). One format per call: an input
    companion of another format than the requested output's is dropped (the bf16 tensor alone is read)."""
    fmt_in = comp_fmt_of(comp)
    if comp is not None and out_fmt not in (COMP_NONE, fmt_in):
        comp, fmt_in = None, COMP_NONE
    fmt = fmt_in or out_fmt
    out_comp = torch.empty(shape, device=device, dtype=_COMP_DTYPE[out_fmt]) if out_fmt else None
    return comp, fmt, out_comp


def _ret(out, out_comp, out_fmt):
    return (out, out_comp) if out_fmt else out


def cbn_apply_act(y, fin, emb, labels, act, upsample, comp=None, out_fmt=COMP_NONE):
    """y: bf16 (NB, H, W, C); fin: [4, C] from bn_finalize (mean, rstd, ...); emb: fp32 [ncls, 2C] or None.
    comp: companion of y; out_fmt != 0: returns (out, out_comp)."""
    _chk(y, torch.bfloat16, "y")
    NB, H, W, C = y.shape
    s = 2 if upsample else 1
    shape = (NB, s * H, s * W, C)
    out = torch.empty(shape, device=y.device, dtype=torch.bfloat16)
    comp, fmt, out_comp = _comp_io(comp, out_fmt, shape, y.device)
    check(_fn("gp_cbn_apply_act")(_p(y), _p(comp), _p(out), _p(out_comp), fmt, NB, H, W, C, _p(fin[0]), _p(fin[1]), _p(emb),
                                  _p(labels), act, 1 if upsample else 0, _stream()), "gp_cbn_apply_act")
    return _ret(out, out_comp, out_fmt)


def cbn_bwd_reduce(da, y, fin, emb, labels, act, upsample, n_classes, comp=None):
    """Returns (S fp32 [2, C], demb fp32 [ncls, 2C] or None). comp: companion of y (the forward's view of it)."""
    _chk(da, torch.bfloat16, "da")
    NB, H, W, C = y.shape
    part = torch.empty((NB, 2, C), device=y.device, dtype=torch.float32)
    S = torch.empty((2, C), device=y.device, dtype=torch.float32)
    demb = torch.empty((n_classes, 2 * C), device=y.device, dtype=torch.float32) if emb is not None else None
    check(_fn("gp_cbn_bwd_reduce")(_p(da), _p(y), _p(comp), comp_fmt_of(comp), NB, H, W, C, _p(fin[0]), _p(fin[1]), _p(emb),
                                   _p(labels), act, 1 if upsample else 0, _p(part), _p(S), _p(demb), n_classes, _stream()),
          "gp_cbn_bwd_reduce")
    return S, demb


def cbn_bwd_apply(da, y, fin, emb, labels, S, count, act, upsample, comp=None):
    NB, H, W, C = y.shape
    dy = torch.empty_like(y)
    check(_fn("gp_cbn_bwd_apply")(_p(da), _p(y), _p(comp), comp_fmt_of(comp), _p(dy), NB, H, W, C, _p(fin[0]), _p(fin[1]),
                                  _p(emb), _p(labels), _p(S), float(count), act, 1 if upsample else 0, _stream()),
          "gp_cbn_bwd_apply")
    return dy


def upsample2x(x, scale=1.0):
    """Nearest x2. scale == 1 is a pure 16-bit copy, so an fp16 companion tensor goes through the same kernel."""
    if x.dtype == torch.float16 and scale == 1.0:
        _chk(x, torch.float16, "x")
    else:
        _chk(x, torch.bfloat16, "x")
    NB, H, W, C = x.shape
    out = torch.empty((NB, 2 * H, 2 * W, C), device=x.device, dtype=x.dtype)
    check(_fn("gp_upsample2x")(_p(x), _p(out), NB, H, W, C, scale, _stream()), "gp_upsample2x")
    return out


def pool2x(x, scale, comp=None, out_fmt=COMP_NONE):
    _chk(x, torch.bfloat16, "x")
    NB, H, W, C = x.shape
    shape = (NB, H // 2, W // 2, C)
    out = torch.empty(shape, device=x.device, dtype=torch.bfloat16)
    comp, fmt, out_comp = _comp_io(comp, out_fmt, shape, x.device)
    check(_fn("gp_pool2x")(_p(x), _p(comp), _p(out), _p(out_comp), fmt, NB, H // 2, W // 2, C, scale, _stream()), "gp_pool2x")
    return _ret(out, out_comp, out_fmt)


def act_fwd(x, act, comp=None, out_fmt=COMP_NONE):
    _chk(x, torch.bfloat16, "x")
    out = torch.empty_like(x)
    comp, fmt, out_comp = _comp_io(comp, out_fmt, x.shape, x.device)
    check(_fn("gp_act_fwd")(_p(x), _p(comp), _p(out), _p(out_comp), fmt, x.numel(), act, _stream()), "gp_act_fwd")
    return _ret(out, out_comp, out_fmt)


def im2col_k3s1(img, out_fmt=COMP_NONE):
    _chk(img, torch.float32, "img")
    NB, ch, H, W = img.shape
    shape = (NB, H, W, 32)
    col = torch.empty(shape, device=img.device, dtype=torch.bfloat16)
    _, fmt, col_comp = _comp_io(None, out_fmt, shape, img.device)
    check(_fn("gp_im2col_k3s1")(_p(img), _p(col), _p(col_comp), fmt, NB, ch, H, W, _stream()), "gp_im2col_k3s1")
    return _ret(col, col_comp, out_fmt)


def col2im_k3s1(col, ch):
    _chk(col, torch.bfloat16, "col")
    NB, H, W, _ = col.shape
    img = torch.empty((NB, ch, H, W), device=col.device, dtype=torch.float32)
    check(_fn("gp_col2im_k3s1")(_p(col), _p(img), NB, ch, H, W, _stream()), "gp_col2im_k3s1")
    return img


def nhwc8_to_image(x, ch, tanh_act):
    """x: bf16 or fp32 (NB, H, W, 8) -> fp32 NCHW image of the first `ch` channels."""
    if x.dtype != torch.float32:
        _chk(x, torch.bfloat16, "x")
    else:
        _chk(x, torch.float32, "x")
    NB, H, W, _ = x.shape
    img = torch.empty((NB, ch, H, W), device=x.device, dtype=torch.float32)
    check(_fn("gp_nhwc8_to_image")(_p(x), 1 if x.dtype == torch.float32 else 0, _p(img), NB, ch, H * W,
                                   1 if tanh_act else 0, _stream()), "gp_nhwc8_to_image")
    return img


def image_to_nhwc8_grad(dout, out, tanh_act):
    _chk(dout, torch.float32, "dout")
    NB, ch, H, W = dout.shape
    dy = torch.empty((NB, H, W, 8), device=dout.device, dtype=torch.bfloat16)
    check(_fn("gp_image_to_nhwc8_grad")(_p(dout), _p(out), _p(dy), NB, ch, H * W, 1 if tanh_act else 0, _stream()),
          "gp_image_to_nhwc8_grad")
    return dy


def relu_sumpool(a, comp=None):
    _chk(a, torch.bfloat16, "a")
    NB, H, W, C = a.shape
    h = torch.empty((NB, C), device=a.device, dtype=torch.float32)
    check(_fn("gp_relu_sumpool")(_p(a), _p(comp), comp_fmt_of(comp), _p(h), NB, H * W, C, _stream()), "gp_relu_sumpool")
    return h


def relu_sumpool_bwd(dh, a):
    _chk(dh, torch.float32, "dh")
    NB, H, W, C = a.shape
    da = torch.empty_like(a)
    check(_fn("gp_relu_sumpool_bwd")(_p(dh), _p(a), _p(da), NB, H * W, C, _stream()), "gp_relu_sumpool_bwd")
    return da


def proj_head_fwd(h, w, b, E, labels):
    _chk(h, torch.float32, "h")
    NB, C = h.shape
    out = torch.empty((NB, 1), device=h.device, dtype=torch.float32)
    check(_fn("gp_proj_head_fwd")(_p(h), _p(w), _p(b), _p(E), _p(labels), _p(out), NB, C, _stream()), "gp_proj_head_fwd")
    return out


def proj_head_bwd(dout, h, w, E, labels, n_classes):
    NB, C = h.shape
    dh = torch.empty_like(h)
    dw = torch.empty((1, C), device=h.device, dtype=torch.float32)
    db = torch.empty((1,), device=h.device, dtype=torch.float32)
    dE = torch.empty((n_classes, C), device=h.device, dtype=torch.float32) if E is not None else None
    check(_fn("gp_proj_head_bwd")(_p(dout), _p(h), _p(w), _p(E), _p(labels), _p(dh), _p(dw), _p(db), _p(dE), NB, C,
                                  n_classes, _stream()), "gp_proj_head_bwd")
    return dh, dw, db, dE


# ------------------------------------------------------------------------------------------------ bf16x3 precision mode
_SIGS.update({
    "gp_split_matrix": [_vp, _vp, _i, _i, _i, _i, _i, _ll, _ll, _i, _ll, _vp],
    "gp_split_conv_weight": [_vp, _vp, _i, _i, _i, _i, _vp],
    "gp_bn_stats_f32": [_vp, _ll, _i, _vp, _vp, _vp],
    "gp_bn_apply_act_split": [_vp, _vp, _vp, _ll, _i, _vp, _vp, _i, _vp],
    "gp_bn_apply_act_pair": [_vp, _vp, _vp, _ll, _i, _vp, _vp, _i, _vp],
    "gp_pair_to_f16": [_vp, _vp, _ll, _vp, _ll, _ll, _i, _vp],
    "gp_bn_stats_comp": [_vp, _vp, _i, _ll, _i, _vp, _vp, _vp],
    "gp_bn_apply_act_comp": [_vp, _vp, _vp, _vp, _i, _ll, _i, _vp, _vp, _i, _vp],
    "gp_bn_bwd_reduce_f32": [_vp, _vp, _ll, _i, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp],
    "gp_bn_bwd_apply_f32": [_vp, _vp, _vp, _ll, _i, _vp, _vp, _vp, _vp, _vp, _vp, _d, _i, _vp, _vp, _f, _vp],
    "gp_bn_bwd_reduce_comp": [_vp, _vp, _vp, _i, _ll, _i, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp],
    "gp_bn_bwd_apply_comp": [_vp, _vp, _vp, _i, _vp, _ll, _i, _vp, _vp, _vp, _vp, _vp, _vp, _d, _i, _vp, _vp, _f, _vp],
    "gp_im2col_k4s2_split": [_vp, _vp, _vp, _i, _i, _i, _i, _vp],
    "gp_col2im_k4s2_f32": [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp],
    "gp_head_fwd_split": [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _ll, _ll, _ll, _vp],
    "gp_head_fwd_comp": [_vp, _vp, _i, _vp, _vp, _vp, _i, _i, _i, _i, _ll, _ll, _ll, _vp],
})


def split_rows(src, R, K, Kp, s_r, s_k, perm=1):
    """fp32 matrix -> (hi, lo) bf16 pair, each [R, Kp] (activation-side operand of a bf16x3 GEMM)."""
    _chk(src, torch.float32, "src")
    dst = torch.empty((2, R, Kp), device=src.device, dtype=torch.bfloat16)
    check(_fn("gp_split_matrix")(_p(src), _p(dst), R, K, R, Kp, Kp, s_r, s_k, perm, R * Kp, _stream()), "gp_split_matrix")
    return dst[0], dst[1]


def split_weight_matrix(src, R, K, Rpad, Kp, s_r, s_k, perm=1):
    """fp32 matrix -> bf16 [Rpad, 2*Kp]: hi block | lo block per row (weight-side operand of a bf16x3 GEMM)."""
    _chk(src, torch.float32, "src")
    dst = torch.empty((Rpad, 2 * Kp), device=src.device, dtype=torch.bfloat16)
    check(_fn("gp_split_matrix")(_p(src), _p(dst), R, K, Rpad, 2 * Kp, Kp, s_r, s_k, perm, Kp, _stream()), "gp_split_matrix")
    return dst


def split_conv_weight(w, n_dim):
    """w fp32 (D0, D1, kh, kw) -> bf16 [N, 2*taps*C] (hi | lo)."""
    _chk(w, torch.float32, "w")
    D0, D1 = w.shape[0], w.shape[1]
    taps = w.shape[2] * w.shape[3]
    N, C = (D0, D1) if n_dim == 0 else (D1, D0)
    dst = torch.empty((N, 2 * taps * C), device=w.device, dtype=torch.bfloat16)
    check(_fn("gp_split_conv_weight")(_p(w), _p(dst), D0, D1, taps, n_dim, _stream()), "gp_split_conv_weight")
    return dst


def bn_stats_f32(y):
    _chk(y, torch.float32, "y")
    C = y.shape[-1]
    st = zeros((2, C), y.device)
    check(_fn("gp_bn_stats_f32")(_p(y), y.numel() // C, C, _p(st[0]), _p(st[1]), _stream()), "gp_bn_stats_f32")
    return st


def bn_apply_act_split(y, fin, act):
    _chk(y, torch.float32, "y")
    C = y.shape[-1]
    hi = torch.empty(y.shape, device=y.device, dtype=torch.bfloat16)
    lo = torch.empty(y.shape, device=y.device, dtype=torch.bfloat16)
    check(_fn("gp_bn_apply_act_split")(_p(y), _p(hi), _p(lo), y.numel() // C, C, _p(fin[2]), _p(fin[3]), act, _stream()),
          "gp_bn_apply_act_split")
    return hi, lo


def bn_bwd_reduce_f32(da, y, fin, act):
    _chk(da, torch.bfloat16, "da")
    _chk(y, torch.float32, "y")
    C = y.shape[-1]
    red = zeros((2, C), y.device)
    check(_fn("gp_bn_bwd_reduce_f32")(_p(da), _p(y), y.numel() // C, C, _p(fin[2]), _p(fin[3]), _p(fin[0]), _p(fin[1]), act,
                                      _p(red[0]), _p(red[1]), _stream()), "gp_bn_bwd_reduce_f32")
    return red


def bn_bwd_apply_f32(da, y, fin, red, count, act, acc=None, acc_scale=1.0):
    C = y.shape[-1]
    dy = torch.empty(y.shape, device=y.device, dtype=torch.bfloat16)
    check(_fn("gp_bn_bwd_apply_f32")(_p(da), _p(y), _p(dy), y.numel() // C, C, _p(fin[2]), _p(fin[3]), _p(fin[0]), _p(fin[1]),
                                     _p(red[0]), _p(red[1]), float(count), act, _p(acc[0]) if acc else None,
                                     _p(acc[1]) if acc else None, float(acc_scale), _stream()), "gp_bn_bwd_apply_f32")
    return dy


def bn_bwd_reduce_comp(da, y, comp, fin, act):
    """bn_bwd_reduce with y read through its companion tensor (the value the forward normalised)."""
    _chk(da, torch.bfloat16, "da")
    if y.dtype == torch.float16:
        y, comp = _as_hi(y)
    _chk(y, torch.bfloat16, "y")
    C = y.shape[-1]
    red = zeros((2, C), y.device)
    check(_fn("gp_bn_bwd_reduce_comp")(_p(da), _p(y), _p(comp), comp_fmt_of(comp), y.numel() // C, C, _p(fin[2]), _p(fin[3]),
                                       _p(fin[0]), _p(fin[1]), act, _p(red[0]), _p(red[1]), _stream()), "gp_bn_bwd_reduce_comp")
    return red


def bn_bwd_apply_comp(da, y, comp, fin, red, count, act, acc=None, acc_scale=1.0):
    if y.dtype == torch.float16:
        y, comp = _as_hi(y)
    C = y.shape[-1]
    dy = torch.empty_like(y)
    check(_fn("gp_bn_bwd_apply_comp")(_p(da), _p(y), _p(comp), comp_fmt_of(comp), _p(dy), y.numel() // C, C, _p(fin[2]),
                                      _p(fin[3]), _p(fin[0]), _p(fin[1]), _p(red[0]), _p(red[1]), float(count), act,
                                      _p(acc[0]) if acc else None, _p(acc[1]) if acc else None, float(acc_scale), _stream()),
          "gp_bn_bwd_apply_comp")
    return dy


def bn_stats_comp(x, comp):
    """Per-channel sum / sum of squares of an activation read through its companion (most precise view)."""
    _chk(x, torch.bfloat16, "x")
    C = x.shape[-1]
    st = zeros((2, C), x.device)
    check(_fn("gp_bn_stats_comp")(_p(x), _p(comp), comp_fmt_of(comp), x.numel() // C, C, _p(st[0]), _p(st[1]), _stream()),
          "gp_bn_stats_comp")
    return st


def _as_hi(y):
    """A pure-fp16 tensor (2-byte pre-BatchNorm storage of the fp16 mode) handed to a companion-format kernel: it is its
    own fp16 companion; the bf16 slot gets the same storage and is never read (act_io.cuh: GP_COMP_F16 reads the
    companion only)."""
    return (y.view(torch.bfloat16), y) if y.dtype == torch.float16 else (y, None)


def bn_apply_act_comp(y, comp, fin, act, out_fmt):
    """act(y * scale + shift) of an activation with a companion; returns (out, out_comp)."""
    if y.dtype == torch.float16:
        y, comp = _as_hi(y)
    _chk(y, torch.bfloat16, "y")
    C = y.shape[-1]
    out = torch.empty_like(y)
    comp, fmt, out_comp = _comp_io(comp, out_fmt, y.shape, y.device)
    check(_fn("gp_bn_apply_act_comp")(_p(y), _p(comp), _p(out), _p(out_comp), fmt, y.numel() // C, C, _p(fin[2]), _p(fin[3]),
                                      act, _stream()), "gp_bn_apply_act_comp")
    return out, out_comp


def pair_to_f16(hi, lo, rows, cols, ld_in, out=None):
    """fp16(hi + lo) of a hi/lo bf16 pair stored with row pitch ld_in -> fp16 [rows, cols] (the single-MMA operand of the
    fp16 mode). hi / lo may be views into one staging buffer (e.g. the hi | lo halves of a packed weight row)."""
    for t, name in ((hi, "hi"), (lo, "lo")):       # may be strided views into one staging buffer: no contiguity check
        if t.dtype != torch.bfloat16:
            raise _lib.GpError("pair_to_f16: %s must be bf16, got %s" % (name, t.dtype))
    _chk(hi if hi.is_contiguous() else hi.new_empty(1), torch.bfloat16, "hi")
    if out is None:
        out = torch.empty((rows, cols), device=hi.device, dtype=torch.float16)
    check(_fn("gp_pair_to_f16")(_p(hi), _p(lo), ld_in, _p(out), cols, rows, cols, _stream()), "gp_pair_to_f16")
    return out


def conv_weight_f16(w, n_dim):
    """w fp32 (D0, D1, kh, kw) -> fp16 [N, taps*C] (same row layout as pack_conv_weight)."""
    wp = split_conv_weight(w, n_dim)                     # [N, 2*K]: hi block | lo block per row
    K = wp.shape[1] // 2
    return pair_to_f16(wp, wp[:, K:], wp.shape[0], K, 2 * K)


def weight_matrix_f16(src, R, K, Rpad, Kp, s_r, s_k, perm=1):
    """fp32 matrix -> fp16 [Rpad, Kp] (same arguments as split_weight_matrix / pack_matrix)."""
    wp = split_weight_matrix(src, R, K, Rpad, Kp, s_r, s_k, perm)
    return pair_to_f16(wp, wp[:, Kp:], Rpad, Kp, 2 * Kp)


def bn_apply_act_pair(y, fin, act):
    """act(y * scale + shift) stored as (bf16, fp16) copies of the same value."""
    _chk(y, torch.float32, "y")
    C = y.shape[-1]
    b = torch.empty(y.shape, device=y.device, dtype=torch.bfloat16)
    h = torch.empty(y.shape, device=y.device, dtype=torch.float16)
    check(_fn("gp_bn_apply_act_pair")(_p(y), _p(b), _p(h), y.numel() // C, C, _p(fin[2]), _p(fin[3]), act, _stream()),
          "gp_bn_apply_act_pair")
    return b, h


def im2col_k4s2_split(img):
    _chk(img, torch.float32, "img")
    NB, ch, Hi, Wi = img.shape
    col = torch.empty((2, NB, Hi // 2, Wi // 2, 64), device=img.device, dtype=torch.bfloat16)
    check(_fn("gp_im2col_k4s2_split")(_p(img), _p(col[0]), _p(col[1]), NB, ch, Hi, Wi, _stream()), "gp_im2col_k4s2_split")
    return col[0], col[1]


def col2im_k4s2_f32(col, bias, ch, act):
    _chk(col, torch.float32, "col")
    NB, Ho, Wo, _ = col.shape
    img = torch.empty((NB, ch, 2 * Ho, 2 * Wo), device=col.device, dtype=torch.float32)
    check(_fn("gp_col2im_k4s2_f32")(_p(col), _p(bias), _p(img), NB, ch, 2 * Ho, 2 * Wo, act, _stream()), "gp_col2im_k4s2_f32")
    return img


def head_fwd_split(a_hi, a_lo, w, bias, O, s_o, s_c, s_hw):
    _chk(a_hi, torch.bfloat16, "a_hi")
    _chk(a_lo, torch.bfloat16, "a_lo")
    NB, H, W, C = a_hi.shape
    out = torch.empty((NB, O), device=a_hi.device, dtype=torch.float32)
    check(_fn("gp_head_fwd_split")(_p(a_hi), _p(a_lo), _p(w), _p(bias), _p(out), NB, H * W, C, O, s_o, s_c, s_hw, _stream()),
          "gp_head_fwd_split")
    return out


def head_fwd_comp(a, comp, w, bias, O, s_o, s_c, s_hw):
    """head_fwd reading the features through their companion tensor (fp16 copy or bf16 low half)."""
    _chk(a, torch.bfloat16, "a")
    NB, H, W, C = a.shape
    out = torch.empty((NB, O), device=a.device, dtype=torch.float32)
    check(_fn("gp_head_fwd_comp")(_p(a), _p(comp), comp_fmt_of(comp), _p(w), _p(bias), _p(out), NB, H * W, C, O, s_o, s_c,
                                  s_hw, _stream()), "gp_head_fwd_comp")
    return out


# ------------------------------------------------------------------------------------------------ fused image edge
_SIGS.update({
    "gp_image_conv_k4s2_fwd": [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp],
    "gp_image_conv_k4s2_wgrad": [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp],
    "gp_image_convt_k4s2_fwd": [_vp, _vp, _i, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp],
})


def image_edge_ok(ch, Hi, Wi, C, transposed=False):
    """Shapes the fused image-edge kernels (csrc/image_edge.cu) take; anything else goes through im2col / col2im.
    (Hi, Wi): the IMAGE side; C: channels of the NHWC side (whole 64-channel TMA boxes: 64 or 128 — every full-width
    DCGAN-family net of the reference). GP_IMAGE_EDGE=0 forces the column-buffer path."""
    import os

    if os.environ.get("GP_IMAGE_EDGE", "1") == "0" or ch != 3 or Hi % 2 or Wi % 4 or C not in (64, 128):
        return False
    Ho, Wo = Hi // 2, Wi // 2
    if (Ho * Wo) % 128:
        return False
    return Wo in (16, 32) if transposed else (Wo <= 64 and 128 % Wo == 0)


def image_conv_fwd(img, w, bias, act, comp_fmt=COMP_NONE, mul=None):
    """Fused im2col + GEMM of the 4x4 stride-2 conv on the fp32 NCHW image (no column buffer). w: fp32 (Cout, ch, 4, 4)
    or any contiguous [Cout][ch*16] view. Returns (out bf16 NHWC, companion or None)."""
    _chk(img, torch.float32, "img")
    _chk(w, torch.float32, "w")
    NB, ch, Hi, Wi = img.shape
    Cout = w.shape[0]
    shape = (NB, Hi // 2, Wi // 2, Cout)
    out = torch.empty(shape, device=img.device, dtype=torch.bfloat16)
    comp = torch.empty(shape, device=img.device, dtype=_COMP_DTYPE[comp_fmt]) if comp_fmt != COMP_NONE else None
    if bias is not None:
        _chk(bias, torch.float32, "bias")
    if mul is not None:
        _chk(mul, torch.float32, "mul")
    check(_fn("gp_image_conv_k4s2_fwd")(_p(img), _p(mul), _p(w), _p(bias), _p(out), _p(comp), comp_fmt, NB, ch, Hi, Wi, Cout,
                                        act, _stream()), "gp_image_conv_k4s2_fwd")
    return out, comp


def image_conv_wgrad(dense, img, mul, dw, dbias=None):
    """dw[m][j] += sum_px dense[px][m] * im2col(img * (1 - mul^2))[px][j] accumulated into dw (fp32, shape (M, ch, 4, 4));
    dbias (fp32 [M], optional) += column sums of dense."""
    _chk(dense, torch.bfloat16, "dense")
    _chk(img, torch.float32, "img")
    _chk(dw, torch.float32, "dw")
    NB, ch, Hi, Wi = img.shape
    M = dense.shape[-1]
    check(_fn("gp_image_conv_k4s2_wgrad")(_p(dense), _p(img), _p(mul), _p(dw), _p(dbias), NB, ch, Hi, Wi, M, _stream()),
          "gp_image_conv_k4s2_wgrad")
    return dw


def image_convt_fwd(x, x_lo, w, bias, ch, act):
    """Fused GEMM + col2im (+ bias + activation) of the 4x4 stride-2 transposed conv onto the fp32 NCHW image.
    x: bf16 or fp16 NHWC (NB, Hs, Ws, C); x_lo: bf16 low halves (bf16x3 operands) or None; w: fp32 [C][ch*16]."""
    if x.dtype == torch.float16:
        fmt = COMP_F16
        _chk(x, torch.float16, "x")
    else:
        _chk(x, torch.bfloat16, "x")
        fmt = COMP_LO if x_lo is not None else COMP_NONE
    if x_lo is not None:
        _chk(x_lo, torch.bfloat16, "x_lo")
    _chk(w, torch.float32, "w")
    NB, Hs, Ws, C = x.shape
    img = torch.empty((NB, ch, 2 * Hs, 2 * Ws), device=x.device, dtype=torch.float32)
    check(_fn("gp_image_convt_k4s2_fwd")(_p(x), _p(x_lo), fmt, _p(w), _p(bias), _p(img), NB, Hs, Ws, C, ch, act, _stream()),
          "gp_image_convt_k4s2_fwd")
    return img


# ------------------------------------------------------------------------------------------------ optimiser edge
_SIGS.update({
    "gp_adam_flat": [_vp, _vp, _vp, _vp, _ll, _d, _d, _d, _d, _vp, _d, _vp],
})


def adam_flat(p, g, m, v, step, lr, beta1, beta2, eps, grad_scale=1.0):
    """In-place Adam update of the flat fp32 buffer p (torch.optim.Adam semantics); step: fp32 device scalar holding
    the number of steps taken so far (not modified here)."""
    for t, name in ((p, "p"), (g, "g"), (m, "m"), (v, "v")):
        _chk(t, torch.float32, name)
    check(_fn("gp_adam_flat")(_p(p), _p(g), _p(m), _p(v), p.numel(), lr, beta1, beta2, eps, _p(step), grad_scale, _stream()),
          "gp_adam_flat")


# ------------------------------------------------------------------------------------------------ SyncBN over peer memory
class PeerCtx(ctypes.Structure):
    """gp_peer_t of include/gpb200.h: peer-mapped device pointers of every rank's symmetric buffer."""
    _fields_ = [("bufs", ctypes.c_void_p * 8), ("world", ctypes.c_int32), ("rank", ctypes.c_int32), ("epoch", ctypes.c_void_p)]


_SIGS.update({
    "gp_peer_allreduce_sum": [_vp, _vp, _i, _vp],
    "gp_bn_finalize_peer": [_vp, _vp, _d, _i, _f, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
})
PEER_MAX_FLOATS = 4096


def peer_buffer_bytes():
    f = _lib.lib().gp_peer_buffer_bytes
    f.restype = ctypes.c_longlong
    return int(f())


def make_peer_ctx(buffer_ptrs, rank, epoch_tensor):
    ctx = PeerCtx()
    for i, ptr in enumerate(buffer_ptrs):
        ctx.bufs[i] = int(ptr)
    ctx.world, ctx.rank, ctx.epoch = len(buffer_ptrs), rank, epoch_tensor.data_ptr()
    return ctx


def peer_allreduce_sum_(ctx, t):
    """In-place sum over ranks of a small contiguous fp32 tensor (<= 4096 elements): one kernel, no NCCL."""
    _chk(t, torch.float32, "t")
    check(_fn("gp_peer_allreduce_sum")(ctypes.addressof(ctx), _p(t), t.numel(), _stream()), "gp_peer_allreduce_sum")
    return t


def bn_finalize_peer(ctx, st, count, gamma, beta, running_mean, running_var, nbt, eps=1e-5, momentum=0.1):
    """bn_finalize with the cross-rank exchange of st = [sum | sumsq] fused in; st ends up holding the global sums."""
    _chk(st, torch.float32, "st")
    C = st.shape[1]
    out = torch.empty((4, C), device=st.device, dtype=torch.float32)
    check(_fn("gp_bn_finalize_peer")(ctypes.addressof(ctx), _p(st), float(count), C, eps, momentum, _p(gamma), _p(beta),
                                     _p(out[0]), _p(out[1]), _p(out[2]), _p(out[3]), _p(running_mean), _p(running_var),
                                     _p(nbt), _stream()), "gp_bn_finalize_peer")
    return out


# ------------------------------------------------------------------------------------------------ BlurPool2d (dcgan_blur)
_SIGS.update({
    "gp_blur3x3_fwd": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp],
    "gp_blur3x3_bwd": [_vp, _vp, _i, _i, _i, _i, _i, _vp],
})


def blur3x3_fwd(x, stride, comp=None, out_fmt=COMP_NONE):
    """x: bf16 (NB, H, W, C) -> reflect-padded [1,2,1]x[1,2,1]/16 depth-wise blur at `stride` (models/ops.py:7-47)."""
    _chk(x, torch.bfloat16, "x")
    NB, H, W, C = x.shape
    shape = (NB, (H - 1) // stride + 1, (W - 1) // stride + 1, C)
    out = torch.empty(shape, device=x.device, dtype=torch.bfloat16)
    comp, fmt, out_comp = _comp_io(comp, out_fmt, shape, x.device)
    check(_fn("gp_blur3x3_fwd")(_p(x), _p(comp), _p(out), _p(out_comp), fmt, NB, H, W, C, stride, _stream()), "gp_blur3x3_fwd")
    return _ret(out, out_comp, out_fmt)


def blur3x3_bwd(dout, H, W, stride):
    """Adjoint of blur3x3_fwd: dout on the output grid -> gradient on the (H, W) input grid."""
    _chk(dout, torch.bfloat16, "dout")
    NB, _, _, C = dout.shape
    din = torch.empty((NB, H, W, C), device=dout.device, dtype=torch.bfloat16)
    check(_fn("gp_blur3x3_bwd")(_p(dout), _p(din), NB, H, W, C, stride, _stream()), "gp_blur3x3_bwd")
    return din
