"""bench.py's contract where it can be checked without a GPU: the reference arm (`--impl reference`: the unmodified reference program on
the host cores, the one place besides cpu_baseline / parity where bench.py may execute oracle/) prints one JSON line with the keys the
driver reads, also under a torchrun-style environment (rank 0 alone works, the others exit 0 without output); and the product arm refuses
to run without a CUDA device instead of falling back to anything."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

ARGS = ["--impl", "reference", "--level", "6", "--steps", "1", "--warmup", "0", "--substeps", "4", "--ref-substeps", "2"]


def run_bench(args, env=None):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                          timeout=600, env=dict(os.environ, **(env or {})))


def test_reference_arm_line():
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "odis_ref_l6")):
        pytest.skip("reference binaries not built (needs the reference tree at build time)")
    r = run_bench(ARGS)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "timesteps/s" and d["higher_is_better"] is True and d["n_gpus"] == 1
    assert d["value"] > 0 and d["dtype"] == "f64" and d["vs_baseline"] is None and d["scaling"] == "strong"
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_under_torchrun_environment():
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "odis_ref_l6")):
        pytest.skip("reference binaries not built")
    other = run_bench(ARGS + ["--gpus", "2"], env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert other.returncode == 0 and not [l for l in other.stdout.splitlines() if l.startswith("{")]
    first = run_bench(ARGS + ["--gpus", "2"], env={"RANK": "0", "LOCAL_RANK": "0", "WORLD_SIZE": "2"})
    assert first.returncode == 0, first.stderr[-2000:]
    d = json.loads([l for l in first.stdout.splitlines() if l.startswith("{")][0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["value"] > 0


def test_product_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = run_bench(["--steps", "1"])
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)
