"""The CPU/GPU switch.  Mirrors reference ``src/architectures.jl:4-20`` as extended by
``ext/nuPGCMCUDAExt.jl:24-33``; here ``GPU()`` means *this library on a B200*.

``CPU()`` exists so that ``on_architecture(CPU(), a)`` can bring results back to the host, but no
solver runs on it: the reference's CPU path (UMFPACK direct solves + Gridap assembly) is the
reference's own, and this package deliberately has no CPU fallback (toolkits refuse ``CPU()``).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from . import lib


class AbstractArchitecture:
    def __eq__(self, other):
        return type(self) is type(other)

    def __hash__(self):
        return hash(type(self).__name__)

    def __repr__(self):
        return f"{type(self).__name__}()"


class CPU(AbstractArchitecture):
    pass


class GPU(AbstractArchitecture):
    """B200 architecture; carries the library context (created lazily, one per device).

    ``comm`` (a ``lib.Comm``) turns the Krylov solves of toolkits built on this architecture into
    row-block sharded, collective solves over the communicator's ranks (one rank per GPU);
    ``ctx`` supplies an explicit context (several ranks sharing one device)."""

    _contexts: dict = {}

    def __init__(self, device: int = 0, comm=None, ctx=None):
        self.device = device
        self.comm = comm
        self._ctx = ctx if ctx is not None else (comm.ctx if comm is not None else None)

    @property
    def ctx(self) -> lib.Context:
        if self._ctx is not None:
            return self._ctx
        c = GPU._contexts.get(self.device)
        if c is None:
            c = lib.Context(self.device)        # raises if no B200 / no library: no fallback
            GPU._contexts[self.device] = c
        return c


def on_architecture(arch, a, **kw):
    """Move an array / sparse matrix to ``arch`` (architectures.jl:9-10, nuPGCMCUDAExt.jl:24-29)."""
    if isinstance(arch, GPU):
        if isinstance(a, (lib.Vector, lib.CsrMatrix)):
            return a
        if sp.issparse(a):
            return arch.ctx.csr(a, **kw)
        return arch.ctx.vector(np.asarray(a, dtype=np.float64))
    if isinstance(arch, CPU):
        if isinstance(a, lib.Vector):
            return a.download()
        if isinstance(a, lib.CsrMatrix):
            raise NotImplementedError("device CSR matrices are not copied back to the host")
        return a
    raise TypeError(arch)


def architecture(a):
    """Which architecture an array lives on (architectures.jl:13-14, nuPGCMCUDAExt.jl:30-31)."""
    return GPU(a.ctx.device) if isinstance(a, (lib.Vector, lib.CsrMatrix)) else CPU()


def vector_type(arch, T=np.float64):
    """Vector type of an architecture (architectures.jl:17, nuPGCMCUDAExt.jl:32)."""
    return lib.Vector if isinstance(arch, GPU) else np.ndarray


def print_memory_status(arch):
    """architectures.jl:20 / nuPGCMCUDAExt.jl:33."""
    if isinstance(arch, GPU):
        free, total = arch.ctx.mem_status()
        print(f"GPU memory usage: {(total - free) / 2**30:.3f} / {total / 2**30:.3f} GiB")
    else:
        import resource
        print(f"CPU memory usage: {resource.getrusage(resource.RUSAGE_SELF).ru_maxrss / 1e6:.3f} GB")
