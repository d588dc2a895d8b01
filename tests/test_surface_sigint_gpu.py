"""SIGINT semantics of the time loop (src/timeIntegrator.cpp:32-37,120,307-312,316-321): the handler raises a flag, the loop
stops at the next check, the restart files are still written, OUTPUT.txt says so and the program's status is 1 (ab3Explicit's
return value; solveODIS logs "SOLVER RETURNED WITH AN ERROR...", src/solver.cpp:52-55)."""
import os
import re
import signal
import subprocess
import time

import pytest

from conftest import ROOT, load_case, make_run_dir
from h5lite_reader import read_h5

pytestmark = pytest.mark.gpu


def test_sigint_stops_the_run_and_leaves_restart_files(tmp_path):
    case = load_case("l3_ecc_full_orbit")
    d = make_run_dir(tmp_path, case)
    text = open(os.path.join(d, "input.in")).read()
    text, n = re.subn(r"(simulation end time;\s*)[^;]+;", r"\g<1>2000;", text)           # far longer than the test will let it run
    assert n == 1
    open(os.path.join(d, "input.in"), "w").write(text)
    exe = os.path.join(ROOT, "geodesicodis_b200", "bin", "ODIS")
    p = subprocess.Popen([exe, "--dir", d, "--quiet"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    out_txt = os.path.join(d, "DATA", "OUTPUT.txt")
    try:
        t0 = time.time()
        while time.time() - t0 < 120:                        # wait until it is well inside the loop
            if os.path.exists(out_txt) and open(out_txt).read().count("DUMPING DATA AT") >= 3:
                break
            assert p.poll() is None, "the run ended before it could be interrupted"
            time.sleep(0.05)
        else:
            raise AssertionError("no progress lines within 120 s")
        p.send_signal(signal.SIGINT)
        rc = p.wait(timeout=120)
    finally:
        if p.poll() is None:
            p.kill()
    assert rc == 1
    log = open(out_txt).read()
    assert "Terminate signal caught..." in log and "SOLVER RETURNED WITH AN ERROR..." in log
    assert "Calculations appear to have finished!" not in log
    dumps = log.count("DUMPING DATA AT")
    assert 3 <= dumps < 2000 * 10
    for name, rows in (("vel_init.txt", 480), ("pres_init.txt", 162)):
        lines = open(os.path.join(d, "InitialConditions", name)).read().splitlines()
        assert len(lines) == rows and all(len(l.split(",")) == 4 for l in lines)
    h5 = read_h5(os.path.join(d, "DATA", "data.h5"))            # closed properly: readable, rows written so far are non-zero
    assert abs(h5["displacement"][dumps - 1]).max() > 0 and abs(h5["displacement"][-1]).max() == 0
