"""Run-time switches of the hot path.

precision:
  "bf16x3" (default) — forward GEMMs use hi/lo split bf16 operands (x_hi*w_hi + x_lo*w_hi + x_hi*w_lo, fp32 accumulate)
            and pre-BatchNorm outputs are stored in fp32: fp32-class forward activations on bf16 tensor cores, needed for
            the north_star gradient-cosine bar (>= 0.999 on the G step). Backward GEMMs stay single bf16.
  "bf16"   — single bf16 operands everywhere, bf16 activation storage: ~1.6x faster, G-step gradient cosine ~0.97.
  "fp16"   — ONE MMA on fp16 operands (11 significant bits; tcgen05 kind::f16 takes fp16 as well as bf16), fp32
            pre-BatchNorm storage, activations stored twice: bf16 (what the backward GEMMs read — the backward stays bf16,
            fp16 would underflow its 1e-8-scale gradients) and fp16 (the next forward operand). Emulated cosines
            (tools/precision_study.py): D-real 0.99999, D-fake 0.9993, G-step 0.9968 — good for every pass except the G step.
Select with gan_playground_b200.config.set_precision(...) or the GP_PRECISION environment variable.
The DCGAN-family nodes (dcgan / acgan / dcgan_specnorm) implement all three; the ResNet and blur nodes implement "bf16"
and "fp16" and run the default "bf16x3" as "fp16" (resnet_scope).
"""
import os

_VALID = ("bf16", "bf16x3", "fp16")
_precision = os.environ.get("GP_PRECISION", "bf16x3")
if _precision not in _VALID:
    raise ValueError("GP_PRECISION must be one of %s" % (_VALID,))


def precision():
    return _precision


def set_precision(p):
    global _precision
    if p not in _VALID:
        raise ValueError("precision must be one of %s" % (_VALID,))
    _precision = p


def x3():
    return _precision == "bf16x3"


def fp16():
    return _precision == "fp16"


class precision_scope:
    """`with precision_scope("bf16"): out = netD(x)` — run the forward passes issued inside the block in the given
    mode (their backward follows the formats the forward saved). Used by engine.DcganStep for the real-image
    discriminator pass: plain bf16 operands already give D-real gradient cosine 0.9999 and logits within 6e-3 of the
    fp32 reference (tests/test_gpu_precision.py), so the 3-MMA forward is spent only on passes that see generated images."""

    def __init__(self, p):
        if p not in _VALID:
            raise ValueError("precision must be one of %s" % (_VALID,))
        self.p = p

    def __enter__(self):
        global _precision
        self.prev = _precision
        _precision = self.p

    def __exit__(self, *exc):
        global _precision
        _precision = self.prev


# Pre-BatchNorm storage of the single-MMA "fp16" mode: "f32" (4 bytes, default) or "f16" (2 bytes; the statistics still come
# from the fp32 accumulators). Measured on the B200 (DCGAN-64, batch 1024, D-fake chain in fp16): the 2-byte storage keeps
# the north_star bars (D-fake cosine 0.99968 at batch 64 / 0.99996 at 1024 vs 0.99973 / 0.99996) but does NOT pay: 16.57
# vs 16.40 ms per step — the fp32-input BatchNorm kernels are the tuned ones, and the GEMM epilogue is not the limiter
# there. Kept as an option (GP_FP16_PREBN=f16).
_fp16_prebn = os.environ.get("GP_FP16_PREBN", "f32")
if _fp16_prebn not in ("f16", "f32"):
    raise ValueError("GP_FP16_PREBN must be f16 or f32")


def fp16_prebn():
    return _fp16_prebn


def set_fp16_prebn(v):
    global _fp16_prebn
    if v not in ("f16", "f32"):
        raise ValueError("fp16 pre-BatchNorm storage must be f16 or f32")
    _fp16_prebn = v


# Spectral norm: all hooks of a forward in one batched call (five launches) instead of ~7 launches per hook.
# GP_BATCHED_SN=0 falls back to the per-hook kernels.
_batched_sn = os.environ.get("GP_BATCHED_SN", "1") != "0"


def batched_sn():
    return _batched_sn


def set_batched_sn(on):
    global _batched_sn
    _batched_sn = bool(on)


def resnet_scope():
    """The scope the ResNet (models/sngan_projection.py) and blur (models/dcgan_blur.py) mirrors run their forward in:
    those nodes implement "bf16" and "fp16"; the global default "bf16x3" maps to "fp16" — one MMA on 11-bit operands
    already meets every north_star bar on the projection pair, which has no BatchNorm in D."""
    return precision_scope("fp16" if _precision in ("bf16x3", "fp16") else "bf16")


# BatchNorm batch statistics accumulated in the GEMM epilogue from the fp32 accumulators (one fewer pass over the conv
# output per BN layer). GP_FUSED_STATS=0 falls back to the standalone statistics kernel.
_fused_stats = os.environ.get("GP_FUSED_STATS", "1") != "0"


def fused_stats():
    return _fused_stats


def set_fused_stats(on):
    global _fused_stats
    _fused_stats = bool(on)


# Backward-side epilogue fusion (functional.BwdLink): the data-gradient GEMM of block i+1 applies block i's activation
# derivative (blocks without BatchNorm) or accumulates block i's BatchNorm-backward sums, instead of a separate pass over
# dA. Correct (tests/test_gpu_conv_gemm.py: bit-identical dA, sums within 2e-4) but OFF by default — measured on the B200
# it LOSES: cfg2 15.93 -> 17.24 ms, cfg3 4.11 -> 4.54, cfg5 9.23 -> 9.77 (GP_BWD_FUSION=1 vs 0, same box). The DCGAN
# data-gradient GEMMs have small K per output element (512-2048), so the extra 2-4 bytes per output element read in the
# epilogue (y or a, row-strided, latency-exposed with ~250 registers already live) stretch a 55-80 us GEMM by more than
# the 40-100 us streaming pass it replaces; the standalone passes run at 0.7-0.9 of the HBM peak.
_bwd_fusion = os.environ.get("GP_BWD_FUSION", "0") == "1"


def bwd_fusion():
    return _bwd_fusion


def set_bwd_fusion(on):
    global _bwd_fusion
    _bwd_fusion = bool(on)


# Batched weight staging (functional.WeightCache.get_conv -> ops.stage_conv_weights): every 4x4 conv weight of a network in
# every operand format / orientation from one launch per optimiser step. GP_BATCH_STAGE=0 restores one launch per tensor.
_batch_stage = os.environ.get("GP_BATCH_STAGE", "1") != "0"


def batch_stage():
    return _batch_stage


def set_batch_stage(on):
    global _batch_stage
    _batch_stage = bool(on)


# Weight-gradient GEMMs on a side stream (engine._AdversarialStep): dW of a layer is needed only at the optimiser step, so
# the step drivers fork it off the backward chain (dgrad -> BatchNorm backward -> next dgrad ...) and join before the
# gradient exchange / optimiser. Pays where one launch cannot fill the GPU (the 128-image shards of the 8-GPU run: 64-128
# tiles and 20-60 us per launch); GP_WGRAD_STREAM = 1 / 0 forces it, "auto" (default) enables it at <= 256 images per GPU.
# Only inside a step driver, and only for parameters whose gradient is delivered in place (ops.grad_target): plain
# autograd users read .grad right after backward() with no join.
_wgrad_stream_mode = os.environ.get("GP_WGRAD_STREAM", "auto")
if _wgrad_stream_mode not in ("auto", "0", "1"):
    raise ValueError("GP_WGRAD_STREAM must be auto, 0 or 1")
_wgrad_side = None      # [stream, tensors kept alive until the join, forked?] while a step driver runs its loop body


def set_wgrad_stream_mode(mode):
    global _wgrad_stream_mode
    if mode not in ("auto", "0", "1"):
        raise ValueError("wgrad stream mode must be auto, 0 or 1")
    _wgrad_stream_mode = mode


def wgrad_stream_wanted(per_gpu_batch):
    if _wgrad_stream_mode == "auto":
        return per_gpu_batch <= 256
    return _wgrad_stream_mode == "1"


def wgrad_side():
    return _wgrad_side


class wgrad_side_scope:
    def __init__(self, stream):
        self.stream = stream

    def __enter__(self):
        global _wgrad_side
        self.prev = _wgrad_side
        # [stream, tensors kept alive, forked since the last join]
        _wgrad_side = [self.stream, [], False] if self.stream is not None else None

    def __exit__(self, *exc):
        global _wgrad_side
        join_wgrad_side()
        _wgrad_side = self.prev


def join_wgrad_side():
    """The current stream waits for every weight gradient issued on the side stream; the tensors they read may be freed."""
    if _wgrad_side is None or not _wgrad_side[2]:
        return      # nothing was forked: waiting on the idle stream would tie a graph capture to uncaptured work
    import torch
    stream, hold, _ = _wgrad_side
    torch.cuda.current_stream().wait_stream(stream)
    del hold[:]
    _wgrad_side[2] = False


# Generator forward of the D-fake chain (main_dcgan.py:77, G(z1)) on its own stream NEXT TO the real-image D pass
# (:70-74): neither reads what the other writes, and at shard sizes where one launch cannot fill the GPU the two chains
# interleave. GP_G_AHEAD = 1 / 0 forces it, "auto" (default) enables it at <= 256 images per GPU. Under data parallelism
# the SyncBN exchanges of that forward use the second peer lane (parallel.peer_lane).
_g_ahead_mode = os.environ.get("GP_G_AHEAD", "auto")
if _g_ahead_mode not in ("auto", "0", "1"):
    raise ValueError("GP_G_AHEAD must be auto, 0 or 1")


def set_g_ahead_mode(mode):
    global _g_ahead_mode
    if mode not in ("auto", "0", "1"):
        raise ValueError("g-ahead mode must be auto, 0 or 1")
    _g_ahead_mode = mode


def g_ahead_wanted(per_gpu_batch):
    if _g_ahead_mode == "auto":
        return per_gpu_batch <= 256
    return _g_ahead_mode == "1"
