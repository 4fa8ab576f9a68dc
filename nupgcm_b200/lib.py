"""ctypes binding of ``libnupgcm_b200.so`` (the C ABI declared in ``include/nupgcm_b200.h``).

This is the Python twin of the ``ccall`` stubs a nuPGCM maintainer would put in
``ext/nuPGCMB200Ext.jl``.  There is no fallback: if the shared library is
missing, or no B200 is present when a context is created, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os
from ctypes import POINTER, byref, c_char_p, c_double, c_float, c_int32, c_int64, c_size_t, c_void_p

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libnupgcm_b200.so")

ORTH_MGS, ORTH_CGS2, ORTH_CGS2_FUSED = 0, 1, 2


class NupgcmError(RuntimeError):
    pass


class SolveStats(C.Structure):
    _fields_ = [("niter", c_int64), ("solved", c_int32), ("inconsistent", c_int32),
                ("breakdown", c_int32), ("reserved", c_int32), ("rnorm", c_double),
                ("rnorm0", c_double), ("device_ms", c_float), ("launches", c_int32),
                ("hist_len", c_int64), ("phase_frac", c_float * 4),
                ("sm_mhz", c_float), ("reserved2", c_float)]


_P = c_void_p
_dp = POINTER(c_double)
_ip = POINTER(c_int64)
_i32p = POINTER(c_int32)

# name -> (argtypes); every function returns int32 except the two noted below
SIGNATURES = {
    "nupgcm_create": [c_int32, POINTER(_P)],
    "nupgcm_destroy": [_P],
    "nupgcm_synchronize": [_P],
    "nupgcm_set_grid": [_P, c_int32],
    "nupgcm_host_alloc": [_P, c_int64, POINTER(c_void_p)],
    "nupgcm_host_free": [_P, c_void_p],
    "nupgcm_comm_create": [_P, c_int32, c_int32, c_int64, POINTER(_P)],
    "nupgcm_comm_destroy": [_P],
    "nupgcm_comm_ipc_handle": [_P, c_void_p],
    "nupgcm_comm_connect_ipc": [_P, c_void_p],
    "nupgcm_comm_connect_local": [_P, POINTER(_P)],
    "nupgcm_csr_shard": [_P, _P],
    "nupgcm_csr_shard_info": [_P, c_int32, _ip, _ip, _ip, _ip],
    "nupgcm_mem_status": [_P, POINTER(c_size_t), POINTER(c_size_t)],
    "nupgcm_device_info": [_P, _i32p, _i32p, _i32p, c_char_p],
    "nupgcm_timer_start": [_P],
    "nupgcm_timer_stop": [_P, POINTER(c_float)],
    "nupgcm_launch_count": [_P, _ip],
    "nupgcm_vec_create": [_P, c_int64, POINTER(_P)],
    "nupgcm_vec_destroy": [_P],
    "nupgcm_vec_size": [_P, _ip],
    "nupgcm_vec_upload": [_P, _dp, c_int64],
    "nupgcm_vec_download": [_P, _dp, c_int64],
    "nupgcm_vec_fill": [_P, c_double],
    "nupgcm_vec_copy": [_P, _P],
    "nupgcm_vec_axpby": [_P, c_double, _P, c_double],
    "nupgcm_vec_dot": [_P, _P, _dp],
    "nupgcm_vec_norm2": [_P, _dp],
    "nupgcm_vec_maxabs": [_P, c_int64, _dp, _i32p],
    "nupgcm_diag_apply": [_P, _P, _P],
    "nupgcm_index_create": [_P, _ip, c_int64, c_int32, POINTER(_P)],
    "nupgcm_index_destroy": [_P],
    "nupgcm_vec_gather": [_P, _P, _P],
    "nupgcm_csr_create": [_P, c_int64, c_int64, c_int64, _ip, _ip, _dp, c_int32, c_int32, POINTER(_P)],
    "nupgcm_csr_destroy": [_P],
    "nupgcm_csr_info": [_P, _ip, _ip, _ip, _ip],
    "nupgcm_csr_update_values": [_P, _dp, c_int64],
    "nupgcm_csr_combine": [_P, _P, _P, _P, c_double],
    "nupgcm_csr_inv_diag": [_P, _P],
    "nupgcm_rcm_order": [c_int64, _ip, _ip, c_int32, _ip],
    "nupgcm_shard_plan": [c_int64, _ip, _ip, c_int32, c_int32, c_int32, _ip, _ip, _ip, _ip],
    "nupgcm_diag_stream_spmv_host": [c_int64, _ip, _ip, _dp, _dp, c_int32, c_int32, _dp, _ip, _ip, _ip, _ip],
    "nupgcm_diag_stream_spmv": [_P, _P, _P, c_int32, c_int32, POINTER(c_float), _dp],
    "nupgcm_diag_tma_stream": [_P, c_int64, c_int32, c_int32, c_int32, c_int32, POINTER(c_float)],
    "nupgcm_diag_latency": [_P, _dp],
    "nupgcm_spmv": [_P, _P, _P, c_double, c_double],
    "nupgcm_cg_solve": [_P, _P, c_double, _P, _P, c_double, c_double, c_int64, _dp, c_int64,
                        POINTER(SolveStats)],
    "nupgcm_gmres_solve": [_P, _P, c_double, _P, _P, c_double, c_double, c_int64, c_int32, c_int32,
                           _dp, c_int64, POINTER(SolveStats)],
    "nupgcm_blockprec_create": [_P, _P, _P, c_int64, _P, _P, c_int64, POINTER(_P)],
    "nupgcm_blockprec_destroy": [_P],
    "nupgcm_blockprec_apply": [_P, _P, _P],
    "nupgcm_blockprec_info": [_P, _ip, _ip],
    "nupgcm_gmres_solve_prec": [_P, _P, _P, _P, c_double, c_double, c_int64, c_int32, _dp, c_int64,
                                POINTER(SolveStats)],
    "nupgcm_diag_reduce_latency": [_P, c_int32, c_int32, c_int32, c_int32, POINTER(c_float)],
    "nupgcm_diag_pingpong": [_P, c_int32, c_int32, c_int32, POINTER(c_float)],
    "nupgcm_diag_xping": [_P, c_int32, c_int32, c_int32, c_int32, POINTER(c_float)],
    "nupgcm_diag_xreduce": [_P, c_int32, c_int32, c_int32, POINTER(c_float)],
    "nupgcm_mesh_create": [_P, c_int64, c_int32, _i32p, _i32p, _dp, _dp, c_int32, _dp, _dp, c_int64,
                           _dp, c_int64, c_int64, _dp, c_int64, POINTER(_P)],
    "nupgcm_mesh_create_orders": [_P, c_int64, c_int32, c_int32, _i32p, _i32p, _dp, _dp, c_int32, _dp, _dp, c_int64,
                                  _dp, c_int64, c_int64, _dp, c_int64, POINTER(_P)],
    "nupgcm_mesh_destroy": [_P],
    "nupgcm_mesh_set_cell_sizes": [_P, _dp, c_int64],
    "nupgcm_cfl_dt": [_P, _P, c_double, c_double, _dp],
    "nupgcm_mesh_enable_kv_rebuild": [_P, _P, _dp],
    "nupgcm_rebuild_kv": [_P, c_double, c_double, c_double, c_double, _P, _P, _P, _P],
    "nupgcm_mesh_enable_nu_rebuild": [_P, _P, _dp, _dp],
    "nupgcm_rebuild_friction": [_P, c_double, c_double, c_double, c_double, c_double, c_double, _P, _P],
    "nupgcm_rhs_adv": [_P, c_int32, c_double, c_double, _P, _P, _P, _P, _P],
    "nupgcm_rhs_combine": [_P, _P, c_double, c_double, _P, _P, _P, _P, _P],
}
OTHER_SYMBOLS = ["nupgcm_version", "nupgcm_last_error", "nupgcm_solve_stats_size"]

_lib = None


def load():
    """Load the shared library (once) and declare every prototype.  Fails loudly if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NupgcmError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` or `make -C nupgcm_b200/csrc`.  There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = c_int32
    lib.nupgcm_version.restype = c_int32
    lib.nupgcm_version.argtypes = []
    lib.nupgcm_last_error.restype = c_char_p
    lib.nupgcm_last_error.argtypes = [_P]
    lib.nupgcm_solve_stats_size.restype = c_int64
    lib.nupgcm_solve_stats_size.argtypes = []
    if lib.nupgcm_solve_stats_size() != C.sizeof(SolveStats):
        raise NupgcmError(f"ABI mismatch: nupgcm_solve_stats is {lib.nupgcm_solve_stats_size()} bytes in the "
                          f"library, {C.sizeof(SolveStats)} in this binding")
    _lib = lib
    return lib


def _check(rc, ctx=None):
    if rc != 0:
        msg = load().nupgcm_last_error(ctx)
        raise NupgcmError(f"libnupgcm_b200 error {rc}: {msg.decode() if msg else '?'}")


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _ptr(a, typ=_dp):
    return a.ctypes.data_as(typ)


class Context:
    """One B200 (``nupgcm_create``).  Owns a stream; all objects created from it share it."""

    def __init__(self, device: int = 0):
        self.lib = load()
        h = _P()
        _check(self.lib.nupgcm_create(device, byref(h)))
        self.h = h
        self.device = device
        self._pinned = []

    def close(self):
        if self.h:
            for p in self._pinned:
                self.lib.nupgcm_host_free(self.h, p)
            self._pinned = []
            self.lib.nupgcm_destroy(self.h)
            self.h = None

    def synchronize(self):
        _check(self.lib.nupgcm_synchronize(self.h), self.h)

    def pinned_array(self, n: int) -> np.ndarray:
        """float64 array of ``n`` entries in page-locked host memory (freed with the context)."""
        p = c_void_p()
        _check(self.lib.nupgcm_host_alloc(self.h, max(int(n), 1) * 8, byref(p)), self.h)
        self._pinned.append(p)
        return np.ctypeslib.as_array((c_double * max(int(n), 1)).from_address(p.value))[:int(n)]

    def set_grid(self, grid: int):
        """Restrict the persistent solver kernels to ``grid`` CTAs (several contexts on one device)."""
        _check(self.lib.nupgcm_set_grid(self.h, int(grid)), self.h)
        return self

    def mem_status(self):
        f, t = c_size_t(), c_size_t()
        _check(self.lib.nupgcm_mem_status(self.h, byref(f), byref(t)), self.h)
        return f.value, t.value

    def device_info(self):
        sm, ma, mi = c_int32(), c_int32(), c_int32()
        name = C.create_string_buffer(64)
        _check(self.lib.nupgcm_device_info(self.h, byref(sm), byref(ma), byref(mi), name), self.h)
        return {"sm_count": sm.value, "cc": (ma.value, mi.value), "name": name.value.decode()}

    def timer_start(self):
        _check(self.lib.nupgcm_timer_start(self.h), self.h)

    def timer_stop(self) -> float:
        ms = c_float()
        _check(self.lib.nupgcm_timer_stop(self.h, byref(ms)), self.h)
        return ms.value

    def reduce_latency(self, mode=1, reps=2000, grid=None, threads=512) -> float:
        us = c_float()
        grid = grid or self.device_info()["sm_count"]
        _check(self.lib.nupgcm_diag_reduce_latency(self.h, mode, reps, grid, threads, byref(us)), self.h)
        return us.value

    def pingpong(self, peer=1, variant=0, reps=2000) -> float:
        us = c_float()
        _check(self.lib.nupgcm_diag_pingpong(self.h, peer, variant, reps, byref(us)), self.h)
        return us.value

    def tma_stream(self, total_bytes, piece, slots, warps, reps=3) -> float:
        """GB/s of bulk copies of `piece` bytes, `slots` in flight per warp, `warps` warps per CTA."""
        out = c_float()
        _check(self.lib.nupgcm_diag_tma_stream(self.h, int(total_bytes), int(piece), int(slots), int(warps),
                                               int(reps), C.byref(out)), self.h)
        return float(out.value)

    def latencies(self):
        out = np.zeros(8)
        _check(self.lib.nupgcm_diag_latency(self.h, _ptr(out)), self.h)
        return dict(zip(["lds64", "dfma", "lds16_lds64", "imad", "dfma_ilp8_1warp", "dfma_ilp8_11warps",
                         "block8_1warp", "block8_11warps"], out))

    def launch_count(self) -> int:
        n = c_int64()
        _check(self.lib.nupgcm_launch_count(self.h, byref(n)), self.h)
        return n.value

    # factories ------------------------------------------------------------------------------
    def vector(self, data_or_n):
        return Vector(self, data_or_n)

    def index(self, idx, index_base=0):
        return Index(self, idx, index_base)

    def csr(self, mat, drop_zeros=False):
        return CsrMatrix(self, mat, drop_zeros)


IPC_HANDLE_BYTES = 64


class Comm:
    """One rank of a sharded-solve communicator (``nupgcm_comm_*``): a device arena the other
    ranks store halo rows and reduction words into.  ``connect_ipc`` takes the 64-byte handles of
    all ranks (exchanged by the caller, e.g. ``torch.distributed.all_gather_object``);
    ``Comm.connect_local`` wires up ranks that live in one process."""

    def __init__(self, ctx: Context, rank: int, nranks: int, max_n: int):
        self.ctx = ctx
        self.lib = ctx.lib
        self.rank, self.nranks, self.max_n = int(rank), int(nranks), int(max_n)
        h = _P()
        _check(self.lib.nupgcm_comm_create(ctx.h, self.rank, self.nranks, self.max_n, byref(h)), ctx.h)
        self.h = h

    def close(self):
        if self.h and self.ctx.h:
            self.lib.nupgcm_comm_destroy(self.h)
        self.h = None

    def xping(self, rank_a=0, rank_b=1, variant=0, reps=2000) -> float:
        us = c_float()
        _check(self.lib.nupgcm_diag_xping(self.h, rank_a, rank_b, variant, reps, byref(us)), self.ctx.h)
        return us.value

    def xreduce(self, count=1, publish=False, reps=2000) -> float:
        us = c_float()
        _check(self.lib.nupgcm_diag_xreduce(self.h, count, int(publish), reps, byref(us)), self.ctx.h)
        return us.value

    def ipc_handle(self) -> bytes:
        buf = C.create_string_buffer(IPC_HANDLE_BYTES)
        _check(self.lib.nupgcm_comm_ipc_handle(self.h, buf), self.ctx.h)
        return buf.raw

    def connect_ipc(self, handles):
        blob = b"".join(handles)
        if len(blob) != IPC_HANDLE_BYTES * self.nranks:
            raise ValueError("connect_ipc needs one 64-byte handle per rank, ordered by rank")
        _check(self.lib.nupgcm_comm_connect_ipc(self.h, C.c_char_p(blob)), self.ctx.h)
        return self

    @staticmethod
    def connect_local(comms):
        arr = (_P * len(comms))(*[c.h for c in comms])
        for c in comms:
            _check(c.lib.nupgcm_comm_connect_local(c.h, arr), c.ctx.h)
        return comms

    @staticmethod
    def from_torch_distributed(ctx: Context, max_n: int, group=None):
        """One rank per process: exchange the IPC handles over ``torch.distributed``."""
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        c = Comm(ctx, rank, world, max_n)
        if world > 1:
            handles = [None] * world
            dist.all_gather_object(handles, c.ipc_handle(), group=group)
            c.connect_ipc(handles)
            dist.barrier(group)                 # every arena is mapped before anyone stores into one
        return c


class Vector:
    """Device FP64 vector (the ``CuVector{Float64}`` of the reference's GPU path)."""

    def __init__(self, ctx: Context, data_or_n):
        self.ctx = ctx
        self.lib = ctx.lib
        h = _P()
        if isinstance(data_or_n, (int, np.integer)):
            n, data = int(data_or_n), None
        else:
            data = _f64(data_or_n).ravel()
            n = data.size
        _check(self.lib.nupgcm_vec_create(ctx.h, n, byref(h)), ctx.h)
        self.h = h
        self.n = n
        if data is not None and n:
            self.upload(data)

    def __len__(self):
        return self.n

    def __del__(self):
        try:
            if self.h and self.ctx.h:
                self.lib.nupgcm_vec_destroy(self.h)
        except Exception:
            pass

    def upload(self, data):
        data = _f64(data).ravel()
        _check(self.lib.nupgcm_vec_upload(self.h, _ptr(data), data.size), self.ctx.h)
        return self

    def download(self, out=None):
        out = np.empty(self.n) if out is None else out
        _check(self.lib.nupgcm_vec_download(self.h, _ptr(out), out.size), self.ctx.h)
        return out

    def fill(self, v):
        _check(self.lib.nupgcm_vec_fill(self.h, float(v)), self.ctx.h)
        return self

    def copy_from(self, other: "Vector"):
        _check(self.lib.nupgcm_vec_copy(self.h, other.h), self.ctx.h)
        return self

    def axpby(self, alpha, x: "Vector", beta):
        _check(self.lib.nupgcm_vec_axpby(self.h, float(alpha), x.h, float(beta)), self.ctx.h)
        return self

    def dot(self, other: "Vector") -> float:
        out = c_double()
        _check(self.lib.nupgcm_vec_dot(self.h, other.h, byref(out)), self.ctx.h)
        return out.value

    def norm2(self) -> float:
        out = c_double()
        _check(self.lib.nupgcm_vec_norm2(self.h, byref(out)), self.ctx.h)
        return out.value

    def maxabs(self, count=0):
        m, nan = c_double(), c_int32()
        _check(self.lib.nupgcm_vec_maxabs(self.h, int(count), byref(m), byref(nan)), self.ctx.h)
        return m.value, bool(nan.value)

    def gather_from(self, src: "Vector", idx: "Index"):
        _check(self.lib.nupgcm_vec_gather(self.h, src.h, idx.h), self.ctx.h)
        return self


def diag_apply(z: Vector, d: Vector, r: Vector):
    _check(z.lib.nupgcm_diag_apply(z.h, d.h, r.h), z.ctx.h)
    return z


class Index:
    def __init__(self, ctx: Context, idx, index_base=0):
        self.ctx = ctx
        self.lib = ctx.lib
        idx = np.ascontiguousarray(idx, dtype=np.int64)
        h = _P()
        _check(self.lib.nupgcm_index_create(ctx.h, _ptr(idx, _ip), idx.size, index_base, byref(h)),
               ctx.h)
        self.h = h
        self.n = idx.size

    def __del__(self):
        try:
            if self.h and self.ctx.h:
                self.lib.nupgcm_index_destroy(self.h)
        except Exception:
            pass


class CsrMatrix:
    """Device CSR matrix built from a SciPy sparse matrix (explicit zeros are kept unless
    ``drop_zeros``)."""

    def __init__(self, ctx: Context, mat, drop_zeros=False):
        import scipy.sparse as sp
        self.ctx = ctx
        self.lib = ctx.lib
        m = mat if sp.isspmatrix_csr(mat) else sp.csr_matrix(mat)
        rowptr = np.ascontiguousarray(m.indptr, dtype=np.int64)
        col = np.ascontiguousarray(m.indices, dtype=np.int64)
        val = _f64(m.data)
        h = _P()
        _check(self.lib.nupgcm_csr_create(ctx.h, m.shape[0], m.shape[1], val.size,
                                          _ptr(rowptr, _ip), _ptr(col, _ip), _ptr(val), 0,
                                          1 if drop_zeros else 0, byref(h)), ctx.h)
        self.h = h
        self.shape = m.shape

    def __del__(self):
        try:
            if self.h and self.ctx.h:
                self.lib.nupgcm_csr_destroy(self.h)
        except Exception:
            pass

    def shard(self, comm: "Comm | None"):
        """Make the persistent solvers on this matrix collective over ``comm`` (row-block sharded)."""
        _check(self.lib.nupgcm_csr_shard(self.h, comm.h if comm is not None else None), self.ctx.h)
        self.comm = comm
        return self

    def shard_info(self, rank: int):
        a, b, c, d = c_int64(), c_int64(), c_int64(), c_int64()
        _check(self.lib.nupgcm_csr_shard_info(self.h, int(rank), byref(a), byref(b), byref(c), byref(d)),
               self.ctx.h)
        return {"row_begin": a.value, "row_end": b.value, "nnz_owned": c.value, "halo_rows": d.value}

    def info(self):
        a, b, c, d = c_int64(), c_int64(), c_int64(), c_int64()
        _check(self.lib.nupgcm_csr_info(self.h, byref(a), byref(b), byref(c), byref(d)), self.ctx.h)
        return {"n_rows": a.value, "n_cols": b.value, "nnz_given": c.value, "nnz_stored": d.value}

    def update_values(self, vals):
        vals = _f64(vals)
        _check(self.lib.nupgcm_csr_update_values(self.h, _ptr(vals), vals.size), self.ctx.h)

    def combine(self, M, Kh, Kv, theta):
        _check(self.lib.nupgcm_csr_combine(self.h, M.h, Kh.h, Kv.h, float(theta)), self.ctx.h)

    def inv_diag(self, out: Vector):
        _check(self.lib.nupgcm_csr_inv_diag(self.h, out.h), self.ctx.h)
        return out

    def stream_spmv(self, x: Vector, y: Vector, reps=1, mode=0) -> float:
        """``y = A x`` by the persistent solvers' streaming SpMV engine alone; µs per product."""
        us = c_float()
        cyc = np.zeros((148, 11, 8)) if mode == 3 else None
        _check(self.lib.nupgcm_diag_stream_spmv(self.h, x.h, y.h, int(reps), int(mode), C.byref(us),
                                                _ptr(cyc) if cyc is not None else None), self.ctx.h)
        return (float(us.value), cyc) if mode == 3 else float(us.value)

    def spmv(self, x: Vector, y: Vector, alpha=1.0, beta=0.0):
        _check(self.lib.nupgcm_spmv(self.h, x.h, y.h, float(alpha), float(beta)), self.ctx.h)
        return y


def rcm_order(mat):
    """RCM ordering of a SciPy sparse matrix's symmetrised pattern (host-only utility)."""
    import scipy.sparse as sp
    m = sp.csr_matrix(mat)
    rowptr = np.ascontiguousarray(m.indptr, dtype=np.int64)
    col = np.ascontiguousarray(m.indices, dtype=np.int64)
    out = np.empty(m.shape[0], dtype=np.int64)
    _check(load().nupgcm_rcm_order(m.shape[0], _ptr(rowptr, _ip), _ptr(col, _ip), 0, _ptr(out, _ip)))
    return out


def stream_spmv_host(mat, x, grid: int = 148, arena: int = 6656):
    """Host-only: ``y = mat @ x`` computed by walking the streaming-SpMV tables the persistent
    kernels use (``nupgcm_diag_stream_spmv_host``).  Returns ``(y, info)`` with the number of tiles,
    stream entries, slice positions and shared-memory wavefronts of the vector gathers."""
    import scipy.sparse as sp
    m = sp.csr_matrix(mat)
    m.sort_indices()
    rowptr = np.ascontiguousarray(m.indptr, dtype=np.int64)
    col = np.ascontiguousarray(m.indices, dtype=np.int64)
    vals, xv = _f64(m.data), _f64(x)
    y = np.empty(m.shape[0])
    out = np.zeros(4, dtype=np.int64)
    _check(load().nupgcm_diag_stream_spmv_host(m.shape[0], _ptr(rowptr, _ip), _ptr(col, _ip), _ptr(vals),
                                               _ptr(xv), int(grid), int(arena), _ptr(y),
                                               _ptr(out[0:], _ip), _ptr(out[1:], _ip), _ptr(out[2:], _ip),
                                               _ptr(out[3:], _ip)))
    return y, {"tiles": int(out[0]), "entries": int(out[1]), "gather_wavefronts": int(out[2]),
               "positions": int(out[3])}


def shard_plan(mat, nranks: int, grid_per_rank: int = 148):
    """Host-only: internal ordering, row blocks and halo ranges of a sharded solve
    (``nupgcm_shard_plan``).  Returns ``perm``, ``row_begin[nranks+1]``,
    ``halo_lo/hi[dst, src]`` (rows of ``src`` pushed to ``dst``; internal row ids)."""
    import scipy.sparse as sp
    m = sp.csr_matrix(mat)
    rowptr = np.ascontiguousarray(m.indptr, dtype=np.int64)
    col = np.ascontiguousarray(m.indices, dtype=np.int64)
    n = m.shape[0]
    perm = np.empty(n, dtype=np.int64)
    rb = np.empty(nranks + 1, dtype=np.int64)
    lo = np.empty((nranks, nranks), dtype=np.int64)
    hi = np.empty((nranks, nranks), dtype=np.int64)
    _check(load().nupgcm_shard_plan(n, _ptr(rowptr, _ip), _ptr(col, _ip), 0, int(nranks),
                                    int(grid_per_rank), _ptr(perm, _ip), _ptr(rb, _ip),
                                    _ptr(lo, _ip), _ptr(hi, _ip)))
    return {"perm": perm, "row_begin": rb, "halo_lo": lo, "halo_hi": hi}


def _solve(fn, A, dinv, pscale, y, x, atol, rtol, itmax, extra, history):
    cap = int(history) if history else 0
    hist = np.empty(max(cap, 1))
    st = SolveStats()
    rc = fn(A.h, dinv.h if dinv is not None else None, float(pscale), y.h, x.h, float(atol),
            float(rtol), int(itmax), *extra, _ptr(hist) if cap else None, cap, byref(st))
    _check(rc, A.ctx.h)
    return st, hist[:st.hist_len].copy()


def cg_solve(A: CsrMatrix, y: Vector, x: Vector, dinv: Vector | None = None, pscale=1.0,
             atol=1e-6, rtol=1e-6, itmax=0, history=0):
    return _solve(A.lib.nupgcm_cg_solve, A, dinv, pscale, y, x, atol, rtol, itmax, (), history)


def gmres_solve(A: CsrMatrix, y: Vector, x: Vector, dinv: Vector | None = None, pscale=1.0,
                atol=1e-6, rtol=1e-6, itmax=0, memory=20, orth=ORTH_MGS, history=0):
    return _solve(A.lib.nupgcm_gmres_solve, A, dinv, pscale, y, x, atol, rtol, itmax,
                  (int(memory), int(orth)), history)


class BlockPrec:
    """``BlockDiagonalPreconditioner`` of two ``CgPreconditioner`` blocks (``nupgcm_blockprec_*``)."""

    def __init__(self, ctx: Context, P: CsrMatrix, P_dinv: Vector, P_itmax: int, T: CsrMatrix,
                 T_dinv: Vector, T_itmax: int):
        self.ctx, self.lib = ctx, ctx.lib
        self._keep = (P, P_dinv, T, T_dinv)                 # the library borrows these handles
        h = _P()
        _check(self.lib.nupgcm_blockprec_create(ctx.h, P.h, P_dinv.h, int(P_itmax), T.h, T_dinv.h,
                                                int(T_itmax), byref(h)), ctx.h)
        self.h = h

    def __del__(self):
        try:
            if self.h and self.ctx.h:
                self.lib.nupgcm_blockprec_destroy(self.h)
        except Exception:
            pass

    def apply(self, x: Vector, y: Vector):
        _check(self.lib.nupgcm_blockprec_apply(self.h, x.h, y.h), self.ctx.h)
        return y

    def info(self):
        a, b = c_int64(), c_int64()
        _check(self.lib.nupgcm_blockprec_info(self.h, byref(a), byref(b)), self.ctx.h)
        return {"applies": a.value, "inner_iters": b.value}


def gmres_solve_prec(A: CsrMatrix, M: BlockPrec, y: Vector, x: Vector, atol=1e-6, rtol=1e-6, itmax=0,
                     memory=20, history=0):
    cap = int(history) if history else 0
    hist = np.empty(max(cap, 1))
    st = SolveStats()
    _check(A.lib.nupgcm_gmres_solve_prec(A.h, M.h, y.h, x.h, float(atol), float(rtol), int(itmax),
                                         int(memory), _ptr(hist) if cap else None, cap, byref(st)), A.ctx.h)
    return st, hist[:st.hist_len].copy()


class ElementMesh:
    """Device-side element tables of the advection RHS (``nupgcm_mesh_create_orders``: P2 velocity,
    P2 or P1 buoyancy — told apart by the number of local DOFs in ``cell_b``)."""

    def __init__(self, ctx: Context, tables: dict):
        self.ctx = ctx
        self.lib = ctx.lib
        cb = np.ascontiguousarray(tables["cell_b"], dtype=np.int32)
        cu = np.ascontiguousarray(tables["cell_u"], dtype=np.int32)
        grad = _f64(tables["grad"])
        vol = _f64(tables["vol"])
        bary = _f64(tables["bary"])
        w = _f64(tables["w"])
        bd = _f64(tables["b_dirichlet"])
        ud = _f64(tables["u_dirichlet"])
        h = _P()
        _check(self.lib.nupgcm_mesh_create_orders(
            ctx.h, cb.shape[0], cu.shape[1], cb.shape[1], _ptr(cb, _i32p), _ptr(cu, _i32p), _ptr(grad),
            _ptr(vol), w.size, _ptr(bary), _ptr(w), int(tables["nb"]), _ptr(bd), bd.size,
            int(tables["nu"]), _ptr(ud), ud.size, byref(h)), ctx.h)
        self.h = h
        if "h_cells" in tables:
            self.set_cell_sizes(tables["h_cells"])

    def __del__(self):
        try:
            if self.h and self.ctx.h:
                self.lib.nupgcm_mesh_destroy(self.h)
        except Exception:
            pass

    def set_cell_sizes(self, h_cells):
        h = _f64(h_cells)
        _check(self.lib.nupgcm_mesh_set_cell_sizes(self.h, _ptr(h), h.size), self.ctx.h)
        return self

    def cfl_dt(self, u: Vector, cfl_factor=0.8, u_min=0.01) -> float:
        out = c_double()
        _check(self.lib.nupgcm_cfl_dt(self.h, u.h, float(cfl_factor), float(u_min), byref(out)), self.ctx.h)
        return out.value

    def enable_kv_rebuild(self, pattern: "CsrMatrix", kv_q):
        kv = _f64(kv_q)
        _check(self.lib.nupgcm_mesh_enable_kv_rebuild(self.h, pattern.h, _ptr(kv)), self.ctx.h)
        return self

    def rebuild_kv(self, alpha, N2, kappa_c, N2min, b: Vector, Kv: "CsrMatrix", rhs_v: Vector, rhs_diff: Vector):
        _check(self.lib.nupgcm_rebuild_kv(self.h, float(alpha), float(N2), float(kappa_c), float(N2min),
                                          b.h, Kv.h, rhs_v.h, rhs_diff.h), self.ctx.h)

    def enable_nu_rebuild(self, A: "CsrMatrix", A0_vals, f_q):
        a0, fq = _f64(A0_vals), _f64(f_q)
        _check(self.lib.nupgcm_mesh_enable_nu_rebuild(self.h, A.h, _ptr(a0), _ptr(fq)), self.ctx.h)
        return self

    def rebuild_friction(self, a2e2, alpha, N2, N2min, smoothing, nu_min, b: Vector, A: "CsrMatrix"):
        _check(self.lib.nupgcm_rebuild_friction(self.h, float(a2e2), float(alpha), float(N2), float(N2min),
                                                  float(smoothing), float(nu_min), b.h, A.h), self.ctx.h)

    def rhs_adv(self, scheme, dt, N2, b, b_prev, u, u_prev, out):
        _check(self.lib.nupgcm_rhs_adv(self.h, int(scheme), float(dt), float(N2), b.h, b_prev.h,
                                       u.h, u_prev.h, out.h), self.ctx.h)
        return out


def rhs_combine(out, rhs_adv, theta, dt, rhs_diff, rhs_flux, rhs_m, rhs_h, rhs_v):
    _check(out.lib.nupgcm_rhs_combine(out.h, rhs_adv.h, float(theta), float(dt), rhs_diff.h,
                                      rhs_flux.h, rhs_m.h, rhs_h.h, rhs_v.h), out.ctx.h)
    return out
