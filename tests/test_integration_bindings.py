"""The reference-side bindings in integration/ (what INTEGRATION.md tells a GeodesicODIS maintainer to add) compile against
the reference's own headers and link with libodis_b200.so; without a GPU the resulting programs stop through the
reference's fatal path instead of computing anything on the host. Their results are checked on the GPU box
(tests/test_surface_hybrid_gpu.py)."""
import os
import subprocess

import pytest

from conftest import ROOT, load_case, make_run_dir

REF_SRC = "/root/reference/src"


@pytest.fixture(scope="module")
def hybrids(built_library):
    from oracle.build_oracle import build_reference
    if os.path.isdir(REF_SRC):
        build_reference((3,), hybrid_levels=(3,))
    paths = {k: os.path.join(ROOT, "oracle", "_ref", f"{k}_l3") for k in ("odis_hybrid", "odis_hybridops")}
    if not all(os.path.exists(p) for p in paths.values()):
        pytest.skip("reference tree not present and no prebuilt hybrids")
    return paths


def test_bindings_use_only_the_c_abi():
    """The glue may touch the library through include/odis_b200.h alone (no torch, no CUDA headers, no oracle)."""
    for name in os.listdir(os.path.join(ROOT, "integration")):
        includes = [ln for ln in open(os.path.join(ROOT, "integration", name)) if ln.lstrip().startswith("#include")]
        for banned in ("cuda", "torch", "oracle", "csrc"):
            assert not any(banned in ln for ln in includes), (name, banned)
    assert '#include "odis_b200.h"' in open(os.path.join(ROOT, "integration", "odis_b200_bridge.h")).read()


@pytest.mark.parametrize("kind", ["odis_hybrid", "odis_hybridops"])
def test_hybrid_links_the_library_and_fails_loudly_without_gpu(hybrids, tmp_path, kind):
    exe = hybrids[kind]
    needed = subprocess.run(["readelf", "-d", exe], stdout=subprocess.PIPE, text=True).stdout
    assert "libodis_b200.so" in needed and "$ORIGIN/../../geodesicodis_b200" in needed
    symbols = subprocess.run(["nm", "-C", "--defined-only", exe], stdout=subprocess.PIPE, text=True).stdout
    if kind == "odis_hybrid":
        assert "ab3Explicit(Globals*, Mesh*)" in symbols and "CatchExit" not in symbols      # the reference's loop is not in the program
    else:
        assert "CatchExit" in symbols and "integrateAB3scalar_reference" in symbols            # its loop is; the renamed originals are unused
    import torch
    if torch.cuda.is_available():
        return
    d = make_run_dir(tmp_path, load_case("l3_obliqwest_earth"))
    subprocess.run([exe, "--quiet-restart"], cwd=d, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=300)
    err = open(os.path.join(d, "DATA", "ERROR.txt")).read()
    assert "no CUDA device available" in err and "no CPU fallback" in err and "TERMINATING ODIS" in err
    assert not os.path.exists(os.path.join(d, "DATA", "ref_final.bin"))
