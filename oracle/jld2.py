"""Reader for the reference's JLD2 fixtures (TEST INFRASTRUCTURE — never imported by the product).

The files under ``/root/reference/test/data`` are "HDF5-based Julia Data Format 0.1.1": a 512-byte
user block, an HDF5 v2 superblock, v2 object headers, contiguous or compact v4 layouts, no
compression.  h5py is not available, so the few structures used are decoded
with ``struct``.  Used only by ``tests/golden/make_fixtures.py`` to turn the fixtures into ``.npz``.
"""
from __future__ import annotations

import struct

import numpy as np

BASE = 512


def _u(buf, off, n):
    return int.from_bytes(buf[off:off + n], "little")


def _messages(buf, addr):
    """Yield (type, body) for the messages of the v2 object header at absolute offset addr."""
    assert buf[addr:addr + 4] == b"OHDR", "not a v2 object header"
    flags = buf[addr + 5]
    off = addr + 6
    if flags & 0x20:
        off += 16
    if flags & 0x10:
        off += 4
    nsz = 1 << (flags & 3)
    chunk = _u(buf, off, nsz)
    off += nsz
    yield from _chunk_messages(buf, off, off + chunk, flags)


def _chunk_messages(buf, off, end, flags):
    while off + 4 <= end:
        mtype = buf[off]
        msize = _u(buf, off + 1, 2)
        mflags = buf[off + 3]
        off += 4
        if flags & 0x04:
            off += 2
        body = buf[off:off + msize]
        if mtype == 0x10:                       # continuation
            caddr = _u(body, 0, 8) + BASE
            clen = _u(body, 8, 8)
            assert buf[caddr:caddr + 4] == b"OCHK"
            yield from _chunk_messages(buf, caddr + 4, caddr + clen - 4, flags)
        else:
            # a shared (committed) datatype is reported as pseudo-type 0x103
            yield (0x103 if (mtype == 3 and mflags & 0x02) else mtype), body
        off += msize


def _links(buf, addr):
    out = {}
    for mtype, body in _messages(buf, addr):
        if mtype != 6:
            continue
        flags = body[1]
        off = 2
        if flags & 0x08:
            off += 1
        if flags & 0x04:
            off += 8
        if flags & 0x10:
            off += 1
        nl = 1 << (flags & 3)
        ln = _u(body, off, nl)
        off += nl
        name = body[off:off + ln].decode()
        off += ln
        out[name] = _u(body, off, 8) + BASE
    return out


def _dataset(buf, addr):
    """Return (dims, type class, type size, raw bytes) of the dataset whose header is at addr."""
    dims, cls, size, raw = (), None, None, None
    for mtype, body in _messages(buf, addr):
        if mtype == 1:
            rank = body[1]
            dims = tuple(_u(body, 4 + 8 * i, 8) for i in range(rank))
        elif mtype == 3:
            cls = body[0] & 0x0F
            size = _u(body, 4, 4)
        elif mtype == 0x103:                    # committed Julia struct type
            cls, size = "committed", None
        elif mtype == 8:
            assert body[0] == 4, "layout version 4 expected"
            lclass = body[1]
            if lclass == 1:
                a = _u(body, 2, 8) + BASE
                n = _u(body, 10, 8)
                raw = buf[a:a + n]
            elif lclass == 0:
                n = _u(body, 2, 2)
                raw = body[4:4 + n]
            else:
                raise ValueError("chunked layouts are not supported")
    return dims, cls, size, raw


def _array(buf, addr):
    dims, cls, size, raw = _dataset(buf, addr)
    n = int(np.prod(dims)) if dims else 1
    if cls == 1 and size == 8:
        a = np.frombuffer(raw, dtype="<f8", count=n)
    elif cls == 0 and size == 8:
        a = np.frombuffer(raw, dtype="<i8", count=n)
    else:
        raise ValueError(f"unsupported datatype class {cls} size {size}")
    return a.copy() if dims else a[0]


def read_jld2(path):
    """Return a dict name -> ndarray / scalar / ``("csc", m, n, colptr, rowval, nzval)``."""
    with open(path, "rb") as fh:
        buf = fh.read()
    assert buf[BASE:BASE + 8] == b"\x89HDF\r\n\x1a\n", "HDF5 superblock not found at offset 512"
    root = _u(buf, BASE + 36, 8) + BASE
    out = {}
    for name, addr in _links(buf, root).items():
        if name.startswith("_"):
            continue
        dims, cls, size, raw = _dataset(buf, addr)
        if cls in (0, 1) and size == 8:
            out[name] = _array(buf, addr)
        elif cls == "committed" and raw is not None and len(raw) == 40:
            # SparseMatrixCSC{Float64,Int64}: {m, n, ref colptr, ref rowval, ref nzval}
            m, n, r1, r2, r3 = struct.unpack("<qqQQQ", raw[:40])
            out[name] = ("csc", m, n, _array(buf, r1 + BASE), _array(buf, r2 + BASE),
                         _array(buf, r3 + BASE))
        else:
            raise ValueError(f"{path}:{name}: unsupported dataset (class {cls}, size {size})")
    return out
