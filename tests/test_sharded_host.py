"""CPU tests of the multi-GPU host logic (no GPU needed): the host-only shard planner and a
world-size-2 gloo execution of the planned layout."""
import os
import subprocess
import sys

import numpy as np

from conftest import ROOT, workload
from nupgcm_b200 import lib


def test_shard_plan_covers_every_footprint():
    _, ops = workload("bowl_mixing", dim=2)
    A = ops["A"].tocsr().copy()
    A.eliminate_zeros()
    n = A.shape[0]
    for nranks in (1, 2, 3, 8):
        plan = lib.shard_plan(A, nranks, grid_per_rank=8)
        perm, rb = plan["perm"], plan["row_begin"]
        assert sorted(perm.tolist()) == list(range(n))
        assert rb[0] == 0 and rb[-1] == n and np.all(np.diff(rb) >= 0)
        P = A[perm][:, perm].tocsr()
        nnz = np.diff(P.indptr)
        per_rank = [nnz[rb[r]:rb[r + 1]].sum() for r in range(nranks)]
        assert max(per_rank) <= 1.25 * (P.nnz / nranks) + 4 * n / nranks     # balanced on nnz (+4/row)
        for d in range(nranks):
            cols = np.unique(P[rb[d]:rb[d + 1]].indices)
            for c in cols[(cols < rb[d]) | (cols >= rb[d + 1])]:
                s = int(np.searchsorted(rb, c, side="right") - 1)
                assert plan["halo_lo"][d, s] <= c < plan["halo_hi"][d, s]
            assert plan["halo_lo"][d, d] == 0 and plan["halo_hi"][d, d] == 0
        # pushed ranges lie inside the source rank's block
        for d in range(nranks):
            for s in range(nranks):
                if plan["halo_hi"][d, s] > plan["halo_lo"][d, s]:
                    assert rb[s] <= plan["halo_lo"][d, s] and plan["halo_hi"][d, s] <= rb[s + 1]


def test_shard_plan_rejects_bad_arguments():
    import pytest
    _, ops = workload("bowl_mixing", dim=2)
    with pytest.raises(lib.NupgcmError):
        lib.shard_plan(ops["A"], 9)
    with pytest.raises(lib.NupgcmError):
        lib.shard_plan(ops["A"], 2, grid_per_rank=0)


def test_planned_layout_runs_cg_over_gloo_world2():
    env = dict(os.environ, OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29541",
           os.path.join(ROOT, "tests", "_gloo_sharded_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert "gloo sharded layout ok" in out.stdout
