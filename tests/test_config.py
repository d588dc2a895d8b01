"""input.in reader vs the values the reference's Globals derived from the same text (golden fixtures)."""
import math
import os

import pytest

from conftest import ALL_CASES, load_case, make_run_dir


@pytest.mark.parametrize("name", ALL_CASES)
def test_derived_scalars_match_reference(odis, tmp_path, name):
    case = load_case(name)
    d = make_run_dir(tmp_path, case)
    g = odis.Globals.load(d)
    s = lambda k: float(case["scalar_" + k][0])
    # bit-exact: same expressions, same libm
    assert g["radius"] == s("radius")
    assert g["angular velocity"] == s("angVel")
    assert g["orbital period"] == s("period")
    assert g["surface gravity"] == s("g")
    assert g["ocean thickness"] == s("h")
    assert g["friction coefficient"] == s("alpha")
    assert g["love reduction factor"] == s("loveReduct")
    assert g["eccentricity"] == s("e")
    assert g["obliquity"] == s("theta")
    assert g.tide_type == int(s("tide_type"))
    assert g.fric_type == int(s("fric_type"))
    assert g.surface_type == int(s("surface_type"))
    dt, n = odis.quantise_time_step(g["orbital period"], g["time step"])
    assert dt == s("timeStep") and n == int(s("totalIter"))


def test_period_is_even_integer_seconds(odis):
    g = odis.Globals.defaults(**{"angular velocity": "7.292e-5", "surface type": "FREE", "solver type": "AB3"})
    assert g["orbital period"] == 86166.0                       # 2*round(pi/Omega), src/globals.cpp:209-214
    assert g["angular velocity"] == 2 * math.pi / 86166.0


def test_key_syntax_and_unknown_keys(odis, tmp_path):
    d = str(tmp_path)
    with open(os.path.join(d, "input.in"), "w") as f:
        f.write("RADIUS; 1.5e6; upper-case key is lowered;\n"
                "not a key; 3; silently ignored;\n"
                "surface type; FREE; x;\nsolver type; AB3; x;\npotential; NONE; x;\nfriction type; LINEAR; x;\n"
                "k2; 0.25; x;\nh2; 0.5; x;\nadvection; maybe; keeps the previous bool;\n"
                "obliquity; 90; degrees;\n")
    g = odis.Globals.load(d)
    assert g["radius"] == 1.5e6
    assert g["love reduction factor"] == 1.0 + 0.25 - 0.5        # FREE: src/boundaryConditions.cpp:23
    assert g["advection"] is True                                # valBool starts true, src/globals.cpp:341
    assert g["obliquity"] == 90 * math.pi / 180.
    assert g["ocean thickness"] == 400                           # Titan default, src/globals.cpp:503


@pytest.mark.parametrize("key,val,msg", [("surface type", "SOLID", "SURFACE"), ("solver type", "LEAPFROG", "SOLVER"),
                                         ("potential", "TIDE", "POTENTIAL"), ("friction type", "CUBIC", "DRAG")])
def test_bad_enum_is_rejected(odis, tmp_path, key, val, msg):
    base = {"surface type": "FREE", "solver type": "AB3", "potential": "ECC", "friction type": "LINEAR"}
    base[key] = val
    with open(os.path.join(str(tmp_path), "input.in"), "w") as f:
        f.write("".join(f"{k}; {v}; c;\n" for k, v in base.items()))
    with pytest.raises(odis.OdisError) as e:
        odis.Globals.load(str(tmp_path))
    assert e.value.code == -4 and msg in str(e.value)


def test_missing_input_file(odis, tmp_path):
    with pytest.raises(odis.OdisError) as e:
        odis.Globals.load(str(tmp_path))
    assert e.value.code == -2


def test_lid_love_shrinks_radius(odis):
    g = odis.Globals.defaults(**{"surface type": "LID_LOVE", "solver type": "AB3", "radius": "252.1e3", "shell thickness": "23e3"})
    assert g["radius"] == 252.1e3 - 23e3                         # src/boundaryConditions.cpp:126
