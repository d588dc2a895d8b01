"""Kernel logic and engine sequencing checked WITHOUT a GPU: the `-m gpu` parity tests themselves, run in a subprocess against
tests/_build/libodis_b200_emu.so — the library's own kernel and engine sources compiled for the host against a small emulation
of the CUDA execution model (tests/simt/simt_emu.h: streams as worker threads, CTAs of a launch in sequence, threads as fibers,
barriers, warp shuffles, atomics, captured graphs; tests/simt/build_emu.py rewrites launch syntax and inline PTX). The arithmetic is the kernels' own, so the
bit-for-bit assertions against the reference fixtures and the oracle hold or fail exactly as they would for the device code's
logic — the default bulk-async staged kernels included (mbarrier objects, cp.async.bulk, named barriers are modelled). What this
cannot show: anything about speed or memory ordering between the CTAs of one launch. Partitioned runs are covered: every solver's stream
is a thread of its own, so the "devices" really run concurrently and exchange halos / harmonic sums through each other's memory and flags. It is test infrastructure: the product library has no CPU path and `geodesicodis_b200` never loads this file
unless ODIS_B200_LIB says so (as this test's subprocess does).

Two groups: a control group of tests that have passed on real B200s (the emulation must agree with the hardware's verdict), and
the tests of code written after the last GPU run (operator surface, hybrids, 3-launch self-gravity, 4-launch nonlinear step,
overlapped output, analytical start)."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.timeout(2400)          # whole groups of -m gpu tests run inside one test here

CONTROL = ["tests/test_step_parity_gpu.py", "tests/test_nonlinear_gpu.py", "tests/test_run_gpu.py", "tests/test_self_gravity_gpu.py",
           "tests/test_ensemble_gpu.py::test_members_match_oracle",
           "tests/test_ensemble_gpu.py::test_ensemble_self_gravity_matches_oracle[4-5-2]"]      # FP64 mma.sync fragments modelled
NEW = ["tests/test_surface_planet_gpu.py", "tests/test_variant_blocks_gpu.py", "tests/test_surface_ops_gpu.py", "tests/test_surface_hybrid_gpu.py", "tests/test_surface_analytical_gpu.py",
       "tests/test_surface_sigint_gpu.py", "tests/test_self_gravity_step_gpu.py", "tests/test_variant_nl4_gpu.py", "tests/test_variant_overlap_gpu.py", "tests/test_variant_ids16_gpu.py"]
# left out under emulation: full-size grids and the slowest parameter sets
SKIP = "not large_grid and not high_degree_matrix_free and not 5-12 and not 5-8 and not 6-4 and not 6-2 and not l6_obliqwest and not band_limited"
DESELECT = ["tests/test_step_parity_gpu.py::test_direct_and_pipelined_kernels_agree[6]"]


@pytest.fixture(scope="module")
def emulated_library(built_library):
    sys.path.insert(0, os.path.join(ROOT, "tests", "simt"))
    import build_emu
    lib = build_emu.build()
    libdir = os.path.join(os.path.dirname(lib), "emu_libdir")            # for the hybrid programs, which link libodis_b200.so by name
    os.makedirs(libdir, exist_ok=True)
    link = os.path.join(libdir, "libodis_b200.so")
    if not os.path.islink(link):
        os.symlink(os.path.join("..", os.path.basename(lib)), link)
    return lib, libdir


def run_gpu_tests_on_the_emulation(lib, libdir, files, extra_env=None, select=SKIP, workers=4):
    env = dict(os.environ, ODIS_B200_LIB=lib, LD_LIBRARY_PATH=libdir + os.pathsep + os.environ.get("LD_LIBRARY_PATH", ""))
    env.update(extra_env or {})
    cmd = [sys.executable, "-m", "pytest", *files, "-m", "gpu", "-q", "-x", *(["-k", select] if select else []), "-p", "no:cacheprovider",
           *(["-n", str(workers)] if workers > 1 else []),           # independent tests, one process each (pytest-xdist)
           *[a for d in DESELECT for a in ("--deselect", d)]]
    r = subprocess.run(cmd, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=1500)
    tail = r.stdout[-3000:]
    assert r.returncode == 0, tail
    assert " passed" in tail and "failed" not in tail, tail
    return tail


def test_emulation_agrees_with_hardware_on_validated_kernels(emulated_library):
    tail = run_gpu_tests_on_the_emulation(*emulated_library, CONTROL)
    passed = int(tail.split(" passed")[0].split()[-1])
    assert passed >= 50, tail


def test_code_written_after_the_last_gpu_run(emulated_library):
    tail = run_gpu_tests_on_the_emulation(*emulated_library, NEW)
    passed = int(tail.split(" passed")[0].split()[-1])
    assert passed >= 30, tail


def test_partitioned_runs_on_concurrent_emulated_devices(emulated_library):
    """tests/test_multigpu.py (2, 4 and 8 ranks: halo exchange bit-identical to one device, self-gravity all-reduce through peer memory) with
    every rank's stream running as a thread of its own: the in-kernel flag waits really wait for the neighbour, a missing host-side
    synchronisation or a call that blocks on another rank's progress shows as a time-out or a mismatch. (The same tests run on
    2 and 4 B200s with `-m gpu`.)"""
    tail = run_gpu_tests_on_the_emulation(*emulated_library, ["tests/test_multigpu.py", "tests/test_variant_ids16_gpu.py::test_narrow_ids_on_a_partitioned_grid",
                                                              "tests/test_self_gravity_step_gpu.py::test_three_launch_step_on_a_partitioned_grid",
                                                              "tests/test_surface_hybrid_gpu.py::test_reference_program_with_the_device_time_loop_on_several_gpus"],
                                          extra_env={"ODIS_B200_EMULATED_DEVICES": "8"}, select="", workers=3)
    assert int(tail.split(" passed")[0].split()[-1]) == 25 and "skipped" not in tail, tail


def test_partitioned_runs_with_slow_fences_and_concurrent_ctas(emulated_library):
    """What the plain emulation cannot show — its CTAs run one after another and a fence costs nothing — and the first hardware run of
    round 2's halo warp did: on a grid whose kernels are shorter than a system-scope fence, the CTA that finished last counted the step
    before another CTA's halo warp had published, an epoch one too high. ODIS_EMU_CTA_THREADS runs the CTAs of a launch on several OS
    threads at once, ODIS_EMU_SLOW_FENCE makes __threadfence_system() last that many scheduler rounds (the other threads of the CTA run
    on meanwhile). With that bug put back, test_partitioned_small_grids_stay_in_step[4-2] fails here (checked by hand); the fixed protocol
    (the publisher records the epoch) must pass, together with the LL-line all-reduce and the pipelined host I/O."""
    files = ["tests/test_multigpu.py::test_partitioned_small_grids_stay_in_step", "tests/test_multigpu.py::test_partitioned_pipelined_io_matches_synchronous_calls",
             "tests/test_multigpu.py::test_partitioned_self_gravity_matches_single_gpu[2-False]", "tests/test_multigpu.py::test_partitioned_run_matches_single_gpu[4]"]
    tail = run_gpu_tests_on_the_emulation(*emulated_library, files, select="", workers=4,
                                          extra_env={"ODIS_B200_EMULATED_DEVICES": "4", "ODIS_EMU_CTA_THREADS": "4", "ODIS_EMU_SLOW_FENCE": "40"})
    assert int(tail.split(" passed")[0].split()[-1]) == 8 and "skipped" not in tail, tail


def test_merged_grid_barrier_kernel_with_a_thread_per_cta(emulated_library):
    """The default single-GPU self-gravity step is ONE cell launch with a grid-wide barrier inside (cell update + harmonic analysis |
    barrier | solve + synthesis), which needs every CTA live at once: the emulation runs it when it gives every CTA of a launch an OS
    thread of its own (ODIS_EMU_CTA_THREADS >= the grid; ODIS_EMU_MERGED=1 tells the library so). Same assertions as on the GPU, plus:
    2 launches per step, i.e. the merged kernel really ran. Run once by hand under the address sanitizer (clean) and the thread
    sanitizer (one report: padded lanes of the synthesis loop read eu_out[0] as a dummy while CTA 0 writes it; the value is discarded)."""
    tail = run_gpu_tests_on_the_emulation(*emulated_library, ["tests/test_self_gravity_step_gpu.py::test_three_launch_step_matches_oracle_and_baseline_kernels"],
                                          select="4-2 or 5-2 or 5-3", workers=3,
                                          extra_env={"ODIS_EMU_MERGED": "1", "ODIS_EMU_CTA_THREADS": "128", "ODIS_TEST_EXPECT_MERGED": "1"})
    assert int(tail.split(" passed")[0].split()[-1]) == 3, tail


def test_memcheck_of_the_kernels_under_address_sanitizer(emulated_library):
    """The emulation built with -fsanitize=address: device arrays are host heap blocks, so any out-of-range load or store of a
    kernel (padding rows, the last partial block, per-CTA partial buffers with 256 / 512-thread blocks) aborts the run."""
    import build_emu
    asan_rt = subprocess.run(["gcc", "-print-file-name=libasan.so"], stdout=subprocess.PIPE, text=True).stdout.strip()
    if not os.path.isabs(asan_rt) or not os.path.exists(asan_rt):
        pytest.skip("no AddressSanitizer runtime with this compiler")
    lib = build_emu.build(asan=True)
    files = ["tests/test_variant_blocks_gpu.py", "tests/test_surface_ops_gpu.py", "tests/test_self_gravity_step_gpu.py", "tests/test_variant_nl4_gpu.py",
             "tests/test_step_parity_gpu.py", "tests/test_self_gravity_gpu.py", "tests/test_variant_ids16_gpu.py",
             "tests/test_multigpu.py::test_partitioned_run_matches_single_gpu[2]",                    # halo push / wait between two concurrent "devices"
             "tests/test_multigpu.py::test_partitioned_self_gravity_matches_single_gpu[2-False]"]     # + all-reduce through peer memory
    select = SKIP + " and not l5_ and not l6_ and not 5-2 and not 5-3 and not full_orbit and not random_state and not kernels_agree"
    tail = run_gpu_tests_on_the_emulation(lib, emulated_library[1], files, select=select,
                                          extra_env={"LD_PRELOAD": asan_rt, "ASAN_OPTIONS": "detect_leaks=0:halt_on_error=1", "ODIS_B200_EMULATED_DEVICES": "2"})
    assert int(tail.split(" passed")[0].split()[-1]) >= 45, tail


def test_racecheck_of_partitioned_runs_under_thread_sanitizer(emulated_library):
    """The emulation built with -fsanitize=thread: streams are threads and every rank is a "device" of its own, so an access that is not
    ordered after another rank's (or the host's) write by an epoch flag (system-scope release / acquire = __atomic), an event or a stream
    synchronisation is reported as a data race. Covers the in-kernel halo push / wait (read-after-write on the ghost slots and
    write-after-read on the slots a neighbour may still be reading), the harmonic all-reduce through peer memory, and the engine's host-side
    waits, and the pipelined host I/O of partitioned solvers (second stream, page-locked pack buffer, the harmonic sums as LL lines). (Removing the flag wait from the cell kernel makes this test report the races -- checked by hand.) 2 ranks here; the 4-rank
    cases, the 16-bit-id kernel and the overlapped output pipeline were run the same way once, clean."""
    import build_emu
    tsan_rt = subprocess.run(["gcc", "-print-file-name=libtsan.so"], stdout=subprocess.PIPE, text=True).stdout.strip()
    if not os.path.isabs(tsan_rt) or not os.path.exists(tsan_rt):
        pytest.skip("no ThreadSanitizer runtime with this compiler")
    lib = build_emu.build(tsan=True)
    report = os.path.join(os.path.dirname(lib), "tsan_report")
    for f in os.listdir(os.path.dirname(lib)):
        if f.startswith("tsan_report"):
            os.remove(os.path.join(os.path.dirname(lib), f))
    files = ["tests/test_multigpu.py::test_partitioned_run_matches_single_gpu[2]",
             "tests/test_multigpu.py::test_partitioned_self_gravity_matches_single_gpu[2-False]",
             "tests/test_multigpu.py::test_partitioned_pipelined_io_matches_synchronous_calls"]      # copy stream / page-locked staging vs the steps
    tail = run_gpu_tests_on_the_emulation(lib, emulated_library[1], files, select="", workers=2,
                                          extra_env={"LD_PRELOAD": tsan_rt, "OMP_NUM_THREADS": "1", "ODIS_B200_EMULATED_DEVICES": "2",
                                                     "TSAN_OPTIONS": f"halt_on_error=0:report_signal_unsafe=0:exitcode=0:log_path={report}"})
    assert int(tail.split(" passed")[0].split()[-1]) == 4, tail
    reports = [f for f in os.listdir(os.path.dirname(lib)) if f.startswith("tsan_report")]
    text = "".join(open(os.path.join(os.path.dirname(lib), f)).read() for f in reports)
    assert "WARNING: ThreadSanitizer" not in text, text[:6000]
