"""The C-ABI library loads and exports every symbol include/gpb200.h declares (no compute calls: CPU-safe)."""
import ctypes
import os
import re

from conftest import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "gpb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gp_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_something():
    syms = declared_symbols()
    assert "gp_conv_fwd" in syms and "gp_conv_wgrad" in syms and len(syms) >= 20


def test_library_exports_every_declared_symbol(built_lib):
    lib = ctypes.CDLL(built_lib)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, "declared in gpb200.h but not exported: %s" % missing


def test_python_binding_table_matches_header(built_lib):
    from gan_playground_b200 import ops

    assert sorted(ops.exported_symbols()) == declared_symbols()


def test_version_and_error_strings(built_lib):
    lib = ctypes.CDLL(built_lib)
    lib.gp_version.restype = ctypes.c_char_p
    lib.gp_last_error.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.gp_version()
    assert isinstance(lib.gp_last_error(), bytes)


def test_sass_is_blackwell_native(built_lib):
    """tcgen05.mma / TMA / TMEM loads must be in the SASS (UTCHMMA / UTMALDG / LDTM), and no legacy HMMA."""
    import shutil
    import subprocess

    if shutil.which("cuobjdump") is None:
        import pytest

        pytest.skip("cuobjdump not available")
    sass = subprocess.run(["cuobjdump", "-sass", built_lib], stdout=subprocess.PIPE, text=True).stdout
    assert "UTCHMMA" in sass and "UTMALDG" in sass and "LDTM" in sass
    assert "HMMA." not in sass.replace("UTCHMMA", "")


def test_every_kernel_waits_on_its_grid_dependencies(built_lib):
    """Programmatic dependent launch (csrc/common.h): ordering between launches is only transitive if EVERY kernel of
    the library executes griddepcontrol.wait (SASS: ACQBULK) before it exits, and the overlap only happens if every
    kernel releases its dependents (griddepcontrol.launch_dependents, SASS: PREEXIT)."""
    import shutil
    import subprocess

    if shutil.which("cuobjdump") is None:
        import pytest

        pytest.skip("cuobjdump not available")
    sass = subprocess.run(["cuobjdump", "-sass", built_lib], stdout=subprocess.PIPE, text=True).stdout
    funcs = sass.split("Function : ")[1:]
    assert len(funcs) > 50
    missing = [f.split("\n")[0][:80] for f in funcs if "ACQBULK" not in f or "PREEXIT" not in f]
    assert not missing, missing


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from gan_playground_b200 import _lib
    import pytest

    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.GpError):
        _lib.lib()


_NULL_PROBE = r"""
import ctypes, json, sys
sys.path.insert(0, %r)
from gan_playground_b200 import ops, _lib
lib = _lib.lib()
out = {}
for name, sig in sorted(ops._SIGS.items()):
    f = getattr(lib, name); f.argtypes = sig; f.restype = ctypes.c_int
    args = [None if t is ctypes.c_void_p else (0.0 if t in (ctypes.c_float, ctypes.c_double) else 0) for t in sig]
    rc = f(*args)
    out[name] = [rc, lib.gp_last_error().decode("utf-8", "replace")]
print("PROBE" + json.dumps(out))
"""


def test_every_entry_point_rejects_null_arguments(built_lib):
    """Error convention of the C ABI: bad arguments return a negative code and leave a message for
    gp_last_error() — before any pointer is dereferenced or anything is launched, so this runs without a GPU. Done in a
    child process: an entry point that dereferenced a null pointer would take the interpreter down with it."""
    import json
    import subprocess
    import sys

    r = subprocess.run([sys.executable, "-c", _NULL_PROBE % ROOT], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                       timeout=300)
    assert r.returncode == 0, "probe died (rc %d): %s" % (r.returncode, r.stderr[-400:])
    res = json.loads(r.stdout[r.stdout.index("PROBE") + 5:])
    assert len(res) >= 50
    for name, (rc, msg) in res.items():
        assert rc < 0, "%s accepted null arguments" % name
        assert msg and ("gp_" in msg or "peer" in msg), (name, msg)


def test_conv_fwd_validation_messages(built_lib):
    """gp_conv_fwd names the offending field (nothing is dereferenced when validation fails: the pointers are fakes)."""
    import ctypes

    from gan_playground_b200 import _lib

    lib = _lib.lib()
    lib.gp_conv_fwd.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    lib.gp_conv_fwd.restype = ctypes.c_int

    def call(**over):
        kw = dict(inp=4096, w=4096, bias=None, out=4096, col_sum=None, col_sumsq=None, NB=2, Hin=8, Win=8, Cin=16, Hout=4,
                  Wout=4, Nout=16, kind=0, act=0, residual=None, in_lo=None, out_lo=None, out_f32=None, flags=0)
        kw.update(over)
        p = _lib.ConvFwd(**kw)
        rc = lib.gp_conv_fwd(ctypes.addressof(p), None)
        return rc, lib.gp_last_error().decode()

    rc, msg = call(Cin=12)
    assert rc < 0 and "Cin=12" in msg
    rc, msg = call(Nout=20)
    assert rc < 0 and "Nout=20" in msg
    rc, msg = call(act=3)
    assert rc < 0 and "tanh" in msg
    # past the shape checks the host builds TMA descriptors through the driver: without a GPU driver that step must
    # fail cleanly (negative code + message), never crash or launch
    import torch

    if not torch.cuda.is_available():
        rc, msg = call()
        assert rc < 0 and msg
