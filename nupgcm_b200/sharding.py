"""Helpers for the row-block sharded (multi-GPU) solves.

The reference is single-GPU; the sharded path is this package's answer to BASELINE.json's
"row-block sharded across the 8 B200s of one box".  One rank per GPU:

* ``torch_comm(max_n)`` — one process per rank (torchrun): context on ``LOCAL_RANK``'s device and
  a ``lib.Comm`` whose IPC handles travel over ``torch.distributed``;
* ``local_ranks(nranks, max_n, devices=...)`` — all ranks inside this process, one context each
  (different devices, or — for tests on a single GPU — one device whose SMs are split evenly
  between the ranks);
* ``run_collective(fns)`` — run one callable per rank concurrently (sharded solves are collective:
  every rank must be inside the same solve at the same time).

A sharded solve waits at most 10 s for a peer (then every rank aborts with an error): when host-side
set-up times differ between processes, put a ``torch.distributed.barrier()`` in front of the first
collective solve.
"""
from __future__ import annotations

import os
import threading

from . import lib
from .architectures import GPU


def torch_comm(max_n: int, group=None):
    """(arch, comm) of this process's rank; ``torch.distributed`` must be initialised."""
    local = int(os.environ.get("LOCAL_RANK", "0"))
    ctx = lib.Context(local)
    comm = lib.Comm.from_torch_distributed(ctx, max_n, group=group)
    return GPU(local, comm=comm), comm


def local_ranks(nranks: int, max_n: int, devices=None):
    """``nranks`` connected communicators in this process.  ``devices`` defaults to device 0 for
    every rank, in which case each rank's persistent kernels get ``SM count // nranks`` CTAs so
    that all ranks are co-resident on the one GPU."""
    devices = list(devices) if devices is not None else [0] * nranks
    if len(devices) != nranks:
        raise ValueError("one device per rank")
    ctxs = [lib.Context(d) for d in devices]
    for d in set(devices):
        share = devices.count(d)
        if share > 1:
            for c, dd in zip(ctxs, devices):
                if dd == d:
                    c.set_grid(c.device_info()["sm_count"] // share)
    comms = [lib.Comm(c, r, nranks, max_n) for r, c in enumerate(ctxs)]
    lib.Comm.connect_local(comms)
    return comms


def run_collective(fns):
    """Call ``fns[r]()`` for every rank concurrently (ctypes releases the GIL inside the library);
    returns the results in rank order and re-raises the first exception."""
    import gc
    out = [None] * len(fns)
    err = [None] * len(fns)
    # Ranks that share one device (tests) must not meet a device-wide synchronisation while a peer's
    # persistent kernel is waiting for them: a cyclic-GC pass that finalises old device objects
    # (cudaFree) in one rank's thread would be exactly that.  Collect now, keep the collector off
    # until every rank is done.  (With one process per GPU a cudaFree only waits for its own device.)
    gc.collect()
    gc_was_enabled = gc.isenabled()
    gc.disable()

    def work(r):
        try:
            out[r] = fns[r]()
        except BaseException as e:      # noqa: BLE001 - re-raised below
            err[r] = e

    ts = [threading.Thread(target=work, args=(r,)) for r in range(len(fns))]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    if gc_was_enabled:
        gc.enable()
    for e in err:
        if e is not None:
            raise e
    return out
