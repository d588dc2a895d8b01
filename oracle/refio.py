"""TEST INFRASTRUCTURE — reader for the record files written by oracle/ref_build/ref_driver.cpp.

Record layout (little endian): u32 name_len, name bytes, u32 dtype (0 = float64, 1 = int32),
u32 ndim, u64 dims[ndim], raw data.  Returns an ordered dict name -> numpy array.
"""
from __future__ import annotations

import struct
from collections import OrderedDict

import numpy as np


def read_records(path: str) -> "OrderedDict[str, np.ndarray]":
    out: "OrderedDict[str, np.ndarray]" = OrderedDict()
    with open(path, "rb") as f:
        buf = f.read()
    off = 0
    while off < len(buf):
        (nl,) = struct.unpack_from("<I", buf, off); off += 4
        name = buf[off:off + nl].decode(); off += nl
        dt, nd = struct.unpack_from("<II", buf, off); off += 8
        dims = struct.unpack_from("<%dQ" % nd, buf, off); off += 8 * nd
        dtype = np.float64 if dt == 0 else np.int32
        n = int(np.prod(dims)) if nd else 1
        arr = np.frombuffer(buf, dtype=dtype, count=n, offset=off).reshape(dims).copy()
        off += n * arr.itemsize
        out[name] = arr
    return out


def read_h5shim(data_dir: str) -> "OrderedDict[str, np.ndarray]":
    """float32 datasets captured by the HDF5 stand-in (oracle/ref_build/shim/H5Cpp.h)."""
    out: "OrderedDict[str, np.ndarray]" = OrderedDict()
    with open(f"{data_dir}/h5shim_index.txt") as f:
        for line in f:
            k, name, dims = line.rstrip("\n").split("|")
            shape = tuple(int(x) for x in dims.split(",")) if dims else ()
            out[name] = np.fromfile(f"{data_dir}/h5shim_{k}.f32", dtype=np.float32).reshape(shape)
    return out
