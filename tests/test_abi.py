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


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from gan_playground_b200 import _lib
    import pytest

    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.GpError):
        _lib.lib()
