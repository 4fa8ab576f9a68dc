"""Throughput of cp.async.bulk (1-D TMA) HBM -> shared memory as a function of the copy size and of the
number of copies in flight: the measurement that sizes the pieces of the streaming SpMV rings."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nupgcm_b200.architectures import GPU            # noqa: E402

ctx = GPU(0).ctx
total = 1 << 30
print("cp.async.bulk HBM -> smem, 1 GiB streamed by 148 CTAs; GB/s   (in flight per SM = warps x slots x piece)")
print(f"{'warps':>5} {'slots':>5} {'piece B':>8} {'inflight KB':>11} {'GB/s':>8}")
for warps, slots, piece in [
        (11, 4, 512), (11, 4, 1024), (11, 4, 2048), (11, 4, 2560), (11, 4, 4096), (11, 2, 8192),
        (11, 2, 2560), (11, 2, 5120), (11, 3, 5120), (11, 8, 2048), (11, 8, 1024), (11, 16, 1024),
        (4, 4, 8192), (4, 3, 16384), (4, 4, 12288), (2, 3, 32768), (1, 3, 65536), (1, 6, 32768), (2, 6, 16384),
        (8, 4, 5120), (8, 3, 8192), (16, 2, 5120), (16, 4, 2560)]:
    if warps * slots * piece > 220 * 1024:
        continue
    gbs = ctx.tma_stream(total, piece, slots, warps)
    print(f"{warps:5d} {slots:5d} {piece:8d} {warps * slots * piece / 1024:11.0f} {gbs:8.0f}", flush=True)
