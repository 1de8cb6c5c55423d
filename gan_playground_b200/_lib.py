"""ctypes binding of libgpb200.so — the C-ABI declared in include/gpb200.h.

The product path has no CPU or library fallback: if the CUDA extension is missing, loading fails loudly.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgpb200.so")

_lib = None

c_void_p = ctypes.c_void_p
c_int = ctypes.c_int32
c_float = ctypes.c_float
c_ll = ctypes.c_longlong


class ConvFwd(ctypes.Structure):
    _fields_ = [
        ("inp", c_void_p), ("w", c_void_p), ("bias", c_void_p), ("out", c_void_p),
        ("col_sum", c_void_p), ("col_sumsq", c_void_p),
        ("NB", c_int), ("Hin", c_int), ("Win", c_int), ("Cin", c_int),
        ("Hout", c_int), ("Wout", c_int), ("Nout", c_int),
        ("kind", c_int), ("act", c_int), ("residual", c_void_p),
        ("in_lo", c_void_p), ("out_lo", c_void_p), ("out_f32", c_void_p),
        ("flags", c_int),
        ("bwd_src", c_void_p), ("bwd_fin", c_void_p), ("bwd_slope", ctypes.c_float), ("bwd_mode", c_int),
    ]


class ConvWgrad(ctypes.Structure):
    _fields_ = [
        ("dense", c_void_p), ("gath", c_void_p), ("dw", c_void_p),
        ("NB", c_int), ("Hs", c_int), ("Ws", c_int), ("Cd", c_int),
        ("Hg", c_int), ("Wg", c_int), ("Cg", c_int),
        ("kind", c_int),
    ]


class GpError(RuntimeError):
    pass


def lib():
    """Return the loaded shared library (building is done by `python -m gan_playground_b200.build`)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise GpError(
                "libgpb200.so not found at %s — run `python -m gan_playground_b200.build` "
                "(there is no CPU / PyTorch fallback for the hot path)" % LIB_PATH)
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.gp_version.restype = ctypes.c_char_p
        _lib.gp_last_error.restype = ctypes.c_char_p
        _lib.gp_launch_count.restype = ctypes.c_uint64
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().gp_last_error().decode("utf-8", "replace")
        raise GpError("%s failed (%d): %s" % (what or "gpb200 call", rc, msg))


def launch_count():
    return int(lib().gp_launch_count())
