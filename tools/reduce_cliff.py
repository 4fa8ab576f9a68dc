"""Latency of the grid-wide reduction primitive vs cooperative grid size and slot replication."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nupgcm_b200.architectures import GPU  # noqa: E402

ctx = GPU(0).ctx
os.environ["NUPGCM_POLL_DEPTH"] = "1"
for layout in (0, 1):
    os.environ["NUPGCM_SLOT_LAYOUT"] = str(layout)
    print(f"layout {layout}: grid | replicas 1,2,4: scalar-only / publishing reduction (us)")
    for grid in (2, 16, 64, 96, 128, 148):
        row = []
        for rep in (1, 2, 4):
            os.environ["NUPGCM_REPLICAS"] = str(rep)
            row.append(" ".join("%5.2f" % ctx.reduce_latency(m, 5000, grid, 512) for m in (1, 2)))
        print("%3d | " % grid + " | ".join(row), flush=True)
