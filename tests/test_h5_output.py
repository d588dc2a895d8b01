"""DATA/data.h5 writer: layout parity with the reference's output (names, shapes, f32, row = dump index).

No HDF5 library exists in this image, so the file format is checked with an independent minimal reader
(tests/h5lite_reader.py) that is itself validated on files written by the real HDF5 library — the two
.h5 files the reference ships — whenever /root/reference is present."""
import os

import numpy as np
import pytest

from h5lite_reader import H5Error, read_h5

REF_H5 = "/root/reference/input_files/grid_l4_10x10_weights.h5"


@pytest.mark.skipif(not os.path.exists(REF_H5), reason="reference tree not present (GPU box)")
def test_reader_parses_files_written_by_the_hdf5_library():
    d = read_h5(REF_H5)
    assert sorted(d) == ["column index", "row index", "weights"]
    assert d["column index"].dtype == np.int32 and d["column index"].shape == (8229,)
    assert d["row index"].shape == (649,) and d["row index"][0] == 0 and d["row index"][-1] == 8229     # CSR row pointers
    assert d["weights"].dtype == np.float64 and d["weights"].shape == (8229,)
    assert np.all(np.diff(d["row index"]) >= 0) and d["column index"].max() < 3 * 642     # 3N columns (x,y,z per cell)
    d6 = read_h5("/root/reference/input_files/grid_l4_6x6_weights.h5")
    assert sorted(d6) == ["column index", "row index", "weights"] and d6["row index"][-1] == d6["weights"].shape[0]


def test_writer_round_trip(odis, tmp_path):
    path = os.path.join(str(tmp_path), "data.h5")
    w = odis.H5Writer(path)
    T, F, N = 4, 480, 162
    names = ["east velocity", "north velocity", "displacement", "dissipated energy"]       # src/outFiles.cpp:250-300
    ids = {n: w.add_dataset(n, (T, F if "velocity" in n or "energy" in n else N)) for n in names}
    ids["dissipation avg output"] = w.add_dataset("dissipation avg output", (T,))
    ids["face longitude"] = w.add_dataset("face longitude", (F,))
    with pytest.raises(odis.OdisError):
        w.add_dataset("displacement", (T, N))               # the reference's pressure/kinetic/dummy2 bug: duplicate name
    rng = np.random.default_rng(0)
    rows = {n: rng.standard_normal((T, F if "velocity" in n or "energy" in n else N)).astype(np.float32) for n in names}
    for t in (0, 2):                                        # rows 1 and 3 stay unwritten -> zeros
        for n in names:
            w.write_rows(ids[n], t, rows[n][t:t + 1])
        w.write_rows(ids["dissipation avg output"], t, np.array([1.5 + t], dtype=np.float32))
    lon = rng.uniform(0, 360, F).astype(np.float32)
    w.write_rows(ids["face longitude"], 0, lon)
    with pytest.raises(odis.OdisError):
        w.write_rows(ids["displacement"], T, rows["displacement"][:1])       # outside the extent
    w.close()
    d = read_h5(path)
    assert sorted(d) == sorted(list(names) + ["dissipation avg output", "face longitude"])
    for n in names:
        assert d[n].dtype == np.float32 and d[n].shape == rows[n].shape
        assert np.array_equal(d[n][0], rows[n][0]) and np.array_equal(d[n][2], rows[n][2])
        assert not d[n][1].any() and not d[n][3].any()
    assert np.array_equal(d["dissipation avg output"], np.array([1.5, 0, 3.5, 0], dtype=np.float32))
    assert np.array_equal(d["face longitude"], lon)
    assert os.path.getsize(path) == int.from_bytes(open(path, "rb").read()[40:48], "little")     # stored EOF == file size


def test_reader_rejects_garbage(tmp_path):
    p = os.path.join(str(tmp_path), "x.h5")
    open(p, "wb").write(b"not an hdf5 file at all" * 10)
    with pytest.raises(H5Error):
        read_h5(p)
