"""TEST INFRASTRUCTURE — minimal reader for the HDF5 dialect of DATA/data.h5 (superblock v0, symbol-table
root group, v1 object headers, contiguous datasets). Independent of the writer in the product; validated
against files written by the real HDF5 library (the reference's input_files/*.h5) in test_h5_output.py."""
from __future__ import annotations

import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Error(ValueError):
    pass


def read_h5(path: str) -> dict:
    b = open(path, "rb").read()
    if b[:8] != b"\x89HDF\r\n\x1a\n":
        raise H5Error("bad signature")
    ver, _fs, _rg, _r, _sh, so, sl, _r2 = struct.unpack_from("<8B", b, 8)
    if ver != 0 or so != 8 or sl != 8:
        raise H5Error(f"unsupported superblock (version {ver}, offsets {so}, lengths {sl})")
    leaf_k, internal_k, _flags = struct.unpack_from("<HHI", b, 16)
    base, _free, eof, _drv = struct.unpack_from("<4Q", b, 24)
    if eof > len(b):
        raise H5Error(f"file truncated: stored EOF {eof} > size {len(b)}")
    _name_off, root_hdr, cache, _res, btree, heap = struct.unpack_from("<QQIIQQ", b, 56)
    if cache != 1:
        raise H5Error("root entry carries no cached symbol table")
    # local heap
    if b[heap:heap + 4] != b"HEAP":
        raise H5Error("bad heap signature")
    _hsize, _hfree, hdata = struct.unpack_from("<QQQ", b, heap + 8)

    def name_at(off: int) -> str:
        end = b.index(b"\0", hdata + off)
        return b[hdata + off:end].decode()

    # group B-tree -> symbol nodes
    entries = []

    def walk(addr: int):
        if b[addr:addr + 4] != b"TREE":
            raise H5Error("bad B-tree signature")
        ntype, level, used = struct.unpack_from("<BBH", b, addr + 4)
        if ntype != 0:
            raise H5Error("not a group B-tree")
        pos = addr + 24
        for k in range(used):
            child = struct.unpack_from("<Q", b, pos + 8)[0]
            pos += 16
            if level > 0:
                walk(child)
            else:
                if b[child:child + 4] != b"SNOD":
                    raise H5Error("bad symbol node signature")
                nsym = struct.unpack_from("<H", b, child + 6)[0]
                if nsym > 2 * leaf_k:
                    raise H5Error("symbol node over-full for the superblock's leaf K")
                for s in range(nsym):
                    noff, ohdr = struct.unpack_from("<QQ", b, child + 8 + 40 * s)
                    entries.append((name_at(noff), ohdr))

    walk(btree)
    names = [n for n, _ in entries]
    if names != sorted(names):
        raise H5Error("symbol table entries are not sorted by name")
    out = {}
    for name, ohdr in entries:
        out[name] = _read_dataset(b, ohdr)
    return out


def _read_dataset(b: bytes, addr: int) -> np.ndarray:
    ver, _r, nmsg, _ref, hsize = struct.unpack_from("<BBHII", b, addr)
    if ver != 1:
        raise H5Error("object header version != 1")
    pos, end = addr + 16, addr + 16 + hsize
    dims = dtype = layout = None
    seen = 0
    while pos < end and seen < nmsg:
        mtype, msize, _flags = struct.unpack_from("<HHB", b, pos)
        body = pos + 8
        if mtype == 0x0001:
            v, rank, flags = struct.unpack_from("<BBB", b, body)
            if v != 1:
                raise H5Error("dataspace version")
            dims = struct.unpack_from("<%dQ" % rank, b, body + 8)
        elif mtype == 0x0003:
            cv, b0, b1, _b2, size = struct.unpack_from("<BBBBI", b, body)
            cls, v = cv & 0x0F, cv >> 4
            if v != 1 or (b0 & 1):
                raise H5Error("datatype version / byte order")
            if cls == 0:
                dtype = {4: np.int32, 8: np.int64}[size] if (b0 & 8) else {4: np.uint32, 8: np.uint64}[size]
            elif cls == 1:
                _off, prec, epos, esize, mpos, msz, bias = struct.unpack_from("<HHBBBBI", b, body + 8)
                if (size, prec, epos, esize, mpos, msz, bias) == (4, 32, 23, 8, 0, 23, 127):
                    dtype = np.float32
                elif (size, prec, epos, esize, mpos, msz, bias) == (8, 64, 52, 11, 0, 52, 1023):
                    dtype = np.float64
                else:
                    raise H5Error("unknown float layout")
            else:
                raise H5Error("datatype class")
        elif mtype == 0x0008:
            v, cls = struct.unpack_from("<BB", b, body)
            if v != 3 or cls != 1:
                raise H5Error("only version-3 contiguous layout is supported")
            layout = struct.unpack_from("<QQ", b, body + 2)
        pos = body + msize
        seen += 1
    if dims is None or dtype is None or layout is None:
        raise H5Error("dataset header incomplete")
    daddr, dsize = layout
    n = int(np.prod(dims))
    if dsize != n * np.dtype(dtype).itemsize:
        raise H5Error("layout size does not match the dataspace")
    if daddr == UNDEF:
        return np.zeros(dims, dtype=dtype)
    return np.frombuffer(b, dtype=dtype, count=n, offset=daddr).reshape(dims).copy()
