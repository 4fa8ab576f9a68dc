"""State I/O: ``save_state`` / ``set_state_from_file!``.  Mirrors reference ``src/IO.jl:1-23``:
``jldsave(ofile; u, p, b, t)`` with the free values in Gridap (un-permuted) order and the scalar
time, and the matching loader.

The reference writes JLD2 ("HDF5-based Julia Data Format 0.1.1": 512-byte user block, HDF5
superblock v2, v2 object headers with Jenkins lookup3 checksums, contiguous / compact v4 layouts,
no chunking or compression).  JLD2.jl and h5py are not available here, so the handful of
structures such a file needs is written (and read back) directly, byte-for-byte in the layout of
the reference's own state files (``test/data/bowl_*.jld2``: Fill-value, Dataspace v2, Datatype
IEEE-754 binary64 LE, Layout v4 messages per dataset; Link-info, Group-info and Link messages in
the root group).  ``tests/test_state_io.py`` checks the checksum routine against the bytes of a
reference file and reads the written files with the independent decoder in ``oracle/jld2.py``.
"""
from __future__ import annotations

import struct

import numpy as np

_BASE = 512
_SIG = b"\x89HDF\r\n\x1a\n"
_UNDEF = b"\xff" * 8
_F64_TYPE = bytes([0x31, 0x20, 0x3F, 0x00]) + struct.pack("<I", 8) + struct.pack(
    "<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)          # class 1 v1, LE IEEE binary64


def _rot(x, k):
    return ((x << k) | (x >> (32 - k))) & 0xFFFFFFFF


def lookup3(data: bytes, initval: int = 0) -> int:
    """Bob Jenkins' lookup3 ``hashlittle`` — the metadata checksum of HDF5 v2 structures."""
    n = len(data)
    a = b = c = (0xDEADBEEF + n + initval) & 0xFFFFFFFF
    off = 0
    M = 0xFFFFFFFF
    while n > 12:
        a = (a + int.from_bytes(data[off:off + 4], "little")) & M
        b = (b + int.from_bytes(data[off + 4:off + 8], "little")) & M
        c = (c + int.from_bytes(data[off + 8:off + 12], "little")) & M
        a = (a - c) & M; a ^= _rot(c, 4); c = (c + b) & M
        b = (b - a) & M; b ^= _rot(a, 6); a = (a + c) & M
        c = (c - b) & M; c ^= _rot(b, 8); b = (b + a) & M
        a = (a - c) & M; a ^= _rot(c, 16); c = (c + b) & M
        b = (b - a) & M; b ^= _rot(a, 19); a = (a + c) & M
        c = (c - b) & M; c ^= _rot(b, 4); b = (b + a) & M
        off += 12
        n -= 12
    if n == 0:
        return c
    tail = data[off:] + b"\x00" * (12 - n)
    a = (a + int.from_bytes(tail[0:4], "little")) & M
    b = (b + int.from_bytes(tail[4:8], "little")) & M
    c = (c + int.from_bytes(tail[8:12], "little")) & M
    c ^= b; c = (c - _rot(b, 14)) & M
    a ^= c; a = (a - _rot(c, 11)) & M
    b ^= a; b = (b - _rot(a, 25)) & M
    c ^= b; c = (c - _rot(b, 16)) & M
    a ^= c; a = (a - _rot(c, 4)) & M
    b ^= a; b = (b - _rot(a, 14)) & M
    c ^= b; c = (c - _rot(b, 24)) & M
    return c


def _msg(mtype: int, body: bytes, flags: int = 0) -> bytes:
    return struct.pack("<BHB", mtype, len(body), flags) + body


def _ohdr(messages: bytes) -> bytes:
    """Version-2 object header, no times / attribute phase change, 1-byte chunk size, with the
    16-byte NIL padding message JLD2 leaves for later growth."""
    messages += _msg(0, b"\x00" * 16)
    if len(messages) > 255:
        raise ValueError("object header too large for a 1-byte chunk size")
    head = b"OHDR" + bytes([2, 0, len(messages)]) + messages
    return head + struct.pack("<I", lookup3(head))


def _dataset_header(dims, layout_body: bytes) -> bytes:
    fill = _msg(5, bytes([3, 0x09]))
    if dims:
        space = _msg(1, bytes([2, len(dims), 0, 1]) + b"".join(struct.pack("<Q", d) for d in dims))
    else:
        space = _msg(1, bytes([2, 0, 0, 0]))
    dtype = _msg(3, _F64_TYPE, flags=1)
    return _ohdr(fill + space + dtype + _msg(8, layout_body))


def write_jld2(path: str, fields: dict, creator: str = "nupgcm_b200") -> None:
    """Write float64 arrays (1-D) and float scalars as a JLD2 file readable by ``jldopen``."""
    user = (b"HDF5-based Julia Data Format, version 0.1.1\x00 (" + creator.encode() + b")\x00")
    buf = bytearray(user.ljust(_BASE, b"\x00"))
    buf += b"\x00" * 48                                     # superblock, filled in at the end
    links = []
    for name, value in fields.items():
        while len(buf) % 8:
            buf += b"\x00"
        addr = len(buf) - _BASE
        if np.ndim(value) == 0:
            body = bytes([4, 0]) + struct.pack("<H", 8) + struct.pack("<d", float(value))
            buf += _dataset_header((), body)
        else:
            arr = np.ascontiguousarray(value, dtype="<f8").ravel()
            # the header's size does not depend on the addresses it stores: lay it out twice
            probe = _dataset_header((arr.size,), bytes([4, 1]) + b"\x00" * 16)
            data_at = (len(buf) + len(probe) + 7) // 8 * 8
            body = bytes([4, 1]) + struct.pack("<QQ", data_at - _BASE, arr.nbytes)
            buf += _dataset_header((arr.size,), body)
            buf += b"\x00" * (data_at - len(buf))
            buf += arr.tobytes()
        links.append((name, addr))
    root = len(buf) - _BASE                                 # JLD2 does not align the root group header
    msgs = _msg(2, bytes([0, 0]) + _UNDEF + _UNDEF) + _msg(0x0A, bytes([0, 0]))
    for name, addr in links:
        nm = name.encode()
        if len(nm) > 255:
            raise ValueError("link name too long")
        msgs += _msg(6, bytes([1, 0x10, 1, len(nm)]) + nm + struct.pack("<Q", addr))
    buf += _ohdr(msgs)
    sb = _SIG + bytes([2, 8, 8, 0]) + struct.pack("<Q", _BASE) + _UNDEF + struct.pack("<QQ", len(buf), root)
    buf[_BASE:_BASE + 48] = sb + struct.pack("<I", lookup3(sb))
    with open(path, "wb") as fh:
        fh.write(bytes(buf))


# ---- reader (files of the shape written above and by the reference's save_state) -------------

def _messages(buf, addr):
    if buf[addr:addr + 4] != b"OHDR" or buf[addr + 4] != 2:
        raise ValueError("not a version-2 object header")
    flags = buf[addr + 5]
    off = addr + 6
    if flags & 0x20:
        off += 16
    if flags & 0x10:
        off += 4
    nsz = 1 << (flags & 3)
    size = int.from_bytes(buf[off:off + nsz], "little")
    off += nsz
    end = off + size
    stored = int.from_bytes(buf[end:end + 4], "little")
    if lookup3(bytes(buf[addr:end])) != stored:
        raise ValueError("object header checksum mismatch")
    if flags & 0x04:
        raise ValueError("object headers with creation-order tracking are not supported")
    while off + 4 <= end:
        mtype, msize, _ = struct.unpack("<BHB", buf[off:off + 4])
        yield mtype, bytes(buf[off + 4:off + 4 + msize])
        off += 4 + msize


def read_jld2(path: str) -> dict:
    """Float64 arrays / scalars of a JLD2 state file (contiguous or compact layouts)."""
    with open(path, "rb") as fh:
        buf = fh.read()
    if buf[_BASE:_BASE + 8] != _SIG or buf[_BASE + 8] != 2:
        raise ValueError(f"{path}: no version-2 HDF5 superblock at offset 512 (not a JLD2 file?)")
    if lookup3(buf[_BASE:_BASE + 44]) != int.from_bytes(buf[_BASE + 44:_BASE + 48], "little"):
        raise ValueError(f"{path}: superblock checksum mismatch")
    base = int.from_bytes(buf[_BASE + 12:_BASE + 20], "little")
    root = int.from_bytes(buf[_BASE + 36:_BASE + 44], "little") + base
    out = {}
    for mtype, body in _messages(buf, root):
        if mtype != 6:
            continue
        flags = body[1]
        off = 2 + (1 if flags & 0x08 else 0) + (8 if flags & 0x04 else 0) + (1 if flags & 0x10 else 0)
        nl = 1 << (flags & 3)
        ln = int.from_bytes(body[off:off + nl], "little")
        name = body[off + nl:off + nl + ln].decode()
        addr = int.from_bytes(body[off + nl + ln:off + nl + ln + 8], "little") + base
        if name.startswith("_"):
            continue
        dims, raw, is_f64, is_i64 = (), None, False, False
        for mt, mb in _messages(buf, addr):
            if mt == 1:
                dims = tuple(int.from_bytes(mb[4 + 8 * i:12 + 8 * i], "little") for i in range(mb[1]))
            elif mt == 3:
                is_f64 = (mb[0] & 0x0F) == 1 and int.from_bytes(mb[4:8], "little") == 8
                # class 0 = fixed point: `t = 0` (an Int64) is what save_state writes for a model
                # without a timestepper (IO.jl:3-4)
                is_i64 = (mb[0] & 0x0F) == 0 and int.from_bytes(mb[4:8], "little") == 8
            elif mt == 8:
                if mb[0] != 4:
                    raise ValueError("layout version 4 expected")
                if mb[1] == 1:
                    a = int.from_bytes(mb[2:10], "little") + base
                    raw = buf[a:a + int.from_bytes(mb[10:18], "little")]
                elif mb[1] == 0:
                    raw = mb[4:4 + int.from_bytes(mb[2:4], "little")]
                else:
                    raise ValueError("chunked layouts are not supported")
        if not (is_f64 or is_i64) or raw is None:
            raise ValueError(f"{path}:{name}: only Float64 / Int64 datasets are supported")
        arr = np.frombuffer(raw, dtype="<f8" if is_f64 else "<i8").astype(np.float64)
        out[name] = arr.reshape(dims) if dims else float(arr[0])
    return out


# ---- the reference's two functions ---------------------------------------------------------------

def save_state(model, ofile: str, history: bool = False) -> None:
    """``save_state(model, ofile)`` (IO.jl:1-10): u, p, b free values in Gridap order and t.

    ``history=True`` additionally stores the previous-step fields ``u_prev``, ``p_prev``, ``b_prev``, the step
    index and Δt, which the reference does not (a resumed BDF2 run of the reference restarts with
    prev = curr, model.jl:120-123): with them ``set_state_from_file_`` + ``run_(..., resume=True)``
    continues a BDF2 run bit for bit.  A file written with ``history=False`` is byte-compatible with the
    reference's."""
    s = model.state
    t = 0.0 if model.timestepper is None else model.timestepper.t
    fields = {"u": s.u, "p": s.p, "b": s.b, "t": t}
    if history and getattr(model, "_has_prev", False):
        d = model.fe_data.dofs
        xp = model._u_prev.download()[d.inv_p_inversion]
        fields.update({"u_prev": xp[:d.nu], "p_prev": xp[d.nu:], "b_prev": model._b_prev.download()[d.inv_p_b],
                       "step_index": float(getattr(model, "_step_index", 1)), "dt": float(model.timestepper.Δt)})
    write_jld2(ofile, fields)


def set_state_from_file_(model, ifile: str):
    """``set_state_from_file!(model, ifile)`` (IO.jl:12-23): upload u, p, b (permuted to solver
    order) and restore t."""
    d = read_jld2(ifile)
    dofs = model.fe_data.dofs
    if d["u"].size != dofs.nu or d["p"].size != dofs.np or d["b"].size != dofs.nb:
        raise ValueError(f"{ifile}: state sizes do not match the model")
    model.inversion.solver.x.upload(np.concatenate([d["u"], d["p"]])[dofs.p_inversion])
    model.xb.upload(d["b"][dofs.p_b])
    if model.timestepper is not None:
        model.timestepper.t = float(d["t"])
    if "b_prev" in d and model.evolution is not None:
        # BDF2 history written by save_state(..., history=True): run_(..., resume=True) continues the run
        model._u_prev.upload(np.concatenate([d["u_prev"], d["p_prev"]])[dofs.p_inversion])
        model._b_prev.upload(d["b_prev"][dofs.p_b])
        model._has_prev = True
        model._step_index = int(d["step_index"])
        model.timestepper.Δt = float(d["dt"])
        if model._step_index >= 2 and model.timestepper.scheme == 2:
            from .evolution import collect_evolution_LHS_
            collect_evolution_LHS_(model.evolution, model.params, model.forcings, model.timestepper)   # model.jl:134-137
    return model
