"""Builds geodesicodis_b200/libodis_b200.so in-tree: C++ host code (g++) + sm_100a CUDA kernels (nvcc).

Flags that matter for parity:
  * host:   -ffp-contract=off  — the mesh-table expressions must evaluate as written (the reference
            oracle is built the same way, oracle/ref_build/Makefile), otherwise GCC's FMA fusion makes
            table values depend on instruction scheduling;
  * device: -fmad=false        — same reason for the time-step kernels (bit-for-bit agreement with the
            CPU solver's operation order). The kernels are HBM-bound; FMA fusion buys nothing.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libodis_b200.so")
# host-only part of the same C ABI (grid generator, mesh tables, input.in, time-step quantisation — no CUDA): what bench.py's
# `--impl reference` arm uses to write the reference's input files, so that the reference process never maps the CUDA library
HOST_LIB = os.path.join(HERE, "libodis_b200_host.so")
HOST_ONLY_EXCLUDE = ["odis_run.cpp"]        # drives the device engine

HOST_SOURCES = ["odis_capi_host.cpp", "odis_config.cpp", "odis_mesh.cpp", "odis_gridgen.cpp", "odis_reorder.cpp", "odis_partition.cpp", "odis_h5lite.cpp",
                "odis_run.cpp", "odis_sh.cpp", "odis_mesh_nl.cpp", "odis_analytic.cpp"]
CUDA_SOURCES = ["odis_kernels.cu", "odis_kernels_pipe.cu", "odis_engine.cu", "odis_ensemble.cu", "odis_sh.cu", "odis_kernels_nl.cu"]

HOST_FLAGS = ["-O2", "-std=c++17", "-fPIC", "-fopenmp", "-ffp-contract=off", "-Wall"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-fmad=false",
              "-Xcompiler", "-fPIC,-fopenmp,-ffp-contract=off", "--expt-relaxed-constexpr"]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA extension cannot be built")


def _digest() -> str:
    h = hashlib.sha256()
    names = [n for n in sorted(os.listdir(CSRC)) if os.path.isfile(os.path.join(CSRC, n))]
    for name in names + ["../build.py", "../../include/odis_b200.h"]:
        with open(os.path.join(CSRC, name), "rb") as f:
            h.update(name.encode())
            h.update(f.read())
    return h.hexdigest()


def build_variant(name: str, defines: list[str]) -> str:
    """Tuning aid: build geodesicodis_b200/_build/libodis_b200_<name>.so with extra -D flags (load it via
    the ODIS_B200_LIB environment variable)."""
    nvcc = _nvcc()
    vdir = os.path.join(BUILD, "variant_" + name)
    os.makedirs(vdir, exist_ok=True)
    objs = []
    for src in HOST_SOURCES:
        obj = os.path.join(vdir, src + ".o")
        subprocess.run(["g++", *HOST_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj], check=True)
        objs.append(obj)
    for src in CUDA_SOURCES:
        obj = os.path.join(vdir, src + ".o")
        subprocess.run([nvcc, *NVCC_FLAGS, *["-D" + d for d in defines], "-c", os.path.join(CSRC, src), "-o", obj], check=True)
        objs.append(obj)
    lib = os.path.join(HERE, f"libodis_b200_{name}.so")
    subprocess.run([nvcc, "-shared", "-o", lib, *objs, "-Xcompiler", "-fopenmp", "-lgomp"], check=True)
    return lib


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile (if sources changed) and return the path of the shared library."""
    os.makedirs(BUILD, exist_ok=True)
    stamp = os.path.join(BUILD, "stamp.txt")
    digest = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(HOST_LIB) and os.path.exists(stamp) and open(stamp).read() == digest:
        return LIB
    nvcc = _nvcc()
    objs = []
    jobs = []
    for src in HOST_SOURCES:
        obj = os.path.join(BUILD, src + ".o")
        jobs.append((["g++", *HOST_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj], obj))
    for src in CUDA_SOURCES:
        obj = os.path.join(BUILD, src + ".o")
        extra = ["-Xptxas", "-v"] if verbose else []
        jobs.append(([nvcc, *NVCC_FLAGS, *extra, "-c", os.path.join(CSRC, src), "-o", obj], obj))
    procs = [(subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True), cmd, obj) for cmd, obj in jobs]
    for p, cmd, obj in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("build failed: %s\n%s" % (" ".join(cmd), out))
        if verbose and out.strip():
            print(out)
        objs.append(obj)
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objs, "-Xcompiler", "-fopenmp", "-lgomp"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed: %s\n%s" % (" ".join(link), r.stdout))
    host_objs = [os.path.join(BUILD, src + ".o") for src in HOST_SOURCES if src not in HOST_ONLY_EXCLUDE]
    r = subprocess.run(["g++", "-shared", "-Wl,-z,defs", "-o", HOST_LIB, *host_objs, "-fopenmp"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("link of the host-only library failed:\n" + r.stdout)
    # the drop-in executable: `ODIS` run from a directory holding input.in (src/main.cpp)
    exe = os.path.join(HERE, "bin", "ODIS")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    cmd = ["g++", "-O2", "-std=c++17", os.path.join(CSRC, "odis_main.cpp"), "-o", exe, "-L" + HERE, "-lodis_b200", "-Wl,-rpath,$ORIGIN/.."]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed: %s\n%s" % (" ".join(cmd), r.stdout))
    with open(stamp, "w") as f:
        f.write(digest)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
