"""Run-time switches of the hot path.

precision:
  "bf16x3" (default) — forward GEMMs use hi/lo split bf16 operands (x_hi*w_hi + x_lo*w_hi + x_hi*w_lo, fp32 accumulate)
            and pre-BatchNorm outputs are stored in fp32: fp32-class forward activations on bf16 tensor cores, needed for
            the north_star gradient-cosine bar (>= 0.999 on the G step). Backward GEMMs stay single bf16.
  "bf16"   — single bf16 operands everywhere, bf16 activation storage: ~1.6x faster, G-step gradient cosine ~0.97.
Select with gan_playground_b200.config.set_precision(...) or the GP_PRECISION environment variable.
Only the DCGAN-family nodes (dcgan / acgan / dcgan_specnorm) implement bf16x3 so far; the ResNet nodes run "bf16".
"""
import os

_VALID = ("bf16", "bf16x3")
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


# BatchNorm batch statistics accumulated in the GEMM epilogue from the fp32 accumulators (one fewer pass over the conv
# output per BN layer). GP_FUSED_STATS=0 falls back to the standalone statistics kernel.
_fused_stats = os.environ.get("GP_FUSED_STATS", "1") != "0"


def fused_stats():
    return _fused_stats


def set_fused_stats(on):
    global _fused_stats
    _fused_stats = bool(on)
