"""The shared library loads on a CPU-only box and exports every entry point include/odis_b200.h declares."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "odis_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(odis_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(built_library):
    lib = ctypes.CDLL(built_library)
    names = declared_symbols()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_binding_covers_header(odis):
    from geodesicodis_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    _lib.load()


def test_version_and_error_text(odis):
    from geodesicodis_b200 import _lib
    lib = _lib.load()
    assert b"sm_100a" in lib.odis_version()
    rc = lib.odis_config_load(b"/nonexistent/run/dir", ctypes.byref(ctypes.c_void_p()))
    assert rc == -2
    assert b"input.in" in lib.odis_last_error()


def test_solver_fails_loudly_without_gpu(odis):
    """No CPU fallback: on a box without a device odis_create must fail with ODIS_ERR_CUDA."""
    import torch
    if torch.cuda.is_available():
        return
    pos, fr, cen = odis.generate_grid(3)
    mesh = odis.Mesh.from_arrays(pos, fr, cen, 1.0e6)
    prm = dict(g=1.0, h=1.0e3, alpha=1e-7, dt=10.0, radius=1.0e6, omega=1e-5, love_reduct=1.0, ecc=0.01, obl=0.0,
               shell_thickness=0.0, semimajor_axis=0.0, potential=5, friction=0, surface=0, init_load=0, reorder=1)
    try:
        odis.Solver(mesh, prm)
    except odis.OdisError as e:
        assert e.code == -5 and "no CPU fallback" in str(e)
    else:
        raise AssertionError("odis_create succeeded without a CUDA device")


def test_header_is_plain_c(tmp_path):
    """The boundary is a C ABI: include/odis_b200.h must compile as C99 (no C++ types leak into the signatures)."""
    import shutil
    import subprocess
    if not shutil.which("gcc"):
        pytest.skip("no gcc")
    src = tmp_path / "chk.c"
    src.write_text('#include "odis_b200.h"\nint main(void) { odis_params p; (void)p; return ODIS_OK; }\n')
    inc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include")
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I" + inc, "-fsyntax-only", str(src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
