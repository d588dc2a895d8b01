"""TEST INFRASTRUCTURE — compiles oracle/lte_oracle.c into oracle/_build/liblte_oracle.so and, when the
reference tree is present (this container only), the reference binaries oracle/_ref/odis_ref_l<L>."""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "liblte_oracle.so")
REF_DIR = os.path.join(HERE, "_ref")


def build_oracle(force: bool = False) -> str:
    src = os.path.join(HERE, "lte_oracle.c")
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        cmd = ["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-Wall", "-o", LIB, src, "-lm"]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError("oracle build failed:\n" + r.stdout)
    return LIB


def build_reference(levels=(3, 4, 5, 6), reference_root: str = "/root/reference", openmp_levels=(), hybrid_levels=()) -> list[str]:
    """Build oracle/_ref/odis_ref_l<L> from the unmodified reference sources (and, for `openmp_levels`, the
    -fopenmp variant odis_ref_l<L>_omp that bench.py's reference arm runs on all host cores; for `hybrid_levels`, the
    reference program with its hot path bound to libodis_b200.so through integration/*.cpp). No-op (returns
    what is already there) when the reference tree is absent, e.g. on the GPU box."""
    have = [os.path.join(REF_DIR, f"odis_ref_l{L}") for L in levels] + [os.path.join(REF_DIR, f"odis_ref_l{L}_omp") for L in openmp_levels] + \
           [os.path.join(REF_DIR, f"{kind}_l{L}") for L in hybrid_levels for kind in ("odis_hybrid", "odis_hybridops")]
    if not os.path.isdir(os.path.join(reference_root, "src")):
        return [p for p in have if os.path.exists(p)]
    for lv, extra in ((levels, []), (openmp_levels, ["OPENMP=1"])):
        if not lv:
            continue
        cmd = ["make", "-C", os.path.join(HERE, "ref_build"), f"REF={reference_root}", "LEVELS=" + " ".join(str(L) for L in lv), *extra, "-j8"]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError("reference build failed:\n" + r.stdout[-4000:])
    if hybrid_levels:
        # the reference program linked against integration/*.cpp + geodesicodis_b200/libodis_b200.so (which must exist already):
        # the drop-in demonstration tests/test_surface_hybrid_gpu.py runs on the GPU box
        lv = " ".join(str(L) for L in hybrid_levels)
        cmd = ["make", "-C", os.path.join(HERE, "ref_build"), f"REF={reference_root}", "LEVELS=" + lv, "HYBRID_LEVELS=" + lv, "hybrids", "-j8"]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError("hybrid (reference + libodis_b200) build failed:\n" + r.stdout[-4000:])
    return [p for p in have if os.path.exists(p)]


def reference_binary(level: int, openmp: bool = False) -> str | None:
    p = os.path.join(REF_DIR, f"odis_ref_l{level}" + ("_omp" if openmp else ""))
    return p if os.path.exists(p) else None


if __name__ == "__main__":
    print(build_oracle(force=True))
    print(build_reference())
