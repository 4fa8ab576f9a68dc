"""Debug: per-CTA time stamps of a window of grid reductions inside k_gmres (comm-warp pipeline)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nupgcm_b200 import lib, workloads as W          # noqa: E402
from nupgcm_b200.architectures import GPU            # noqa: E402

ctx = GPU(0).ctx
w = W.bowl_example(h=0.08)
ops = W.host_operands(w)
A = ops["A"]
y = ops["B"] @ ops["b_init"] + ops["b0"]
dA = ctx.csr(A, drop_zeros=True)
dy = ctx.vector(y)
os.makedirs("gpurun_out", exist_ok=True)
names = ["post", "seen", "stored", "w0_has_all", "all_have_all", "resume"]
for orth, name in ((lib.ORTH_MGS, "mgs"), (lib.ORTH_CGS2, "cgs2")):
    path = f"gpurun_out/trace_{name}.bin"
    os.environ["NUPGCM_TRACE_FILE"] = path
    x = ctx.vector(y.size)
    st, _ = lib.gmres_solve(dA, dy, x, pscale=ops["pscale"], atol=0, rtol=1e-30, itmax=1500, orth=orth)
    os.environ.pop("NUPGCM_TRACE_FILE")
    raw = np.fromfile(path, dtype=np.uint64)
    grid, win, ns, _ = np.frombuffer(raw[:2].tobytes(), dtype=np.int32)
    t = raw[2:].reshape(win, ns, grid).astype(np.int64)
    ok = (t > 0).all(axis=(1, 2))
    t = t[ok]
    print(f"== {name}: {st.device_ms * 1e3 / st.niter:.2f} us/iter, {ok.sum()} reductions traced, grid {grid}")
    for i in range(1, ns):
        d = (t[:, i, :] - t[:, i - 1, :]) / 1e3
        print(f"  {names[i-1]:>13s} -> {names[i]:<13s}: mean {d.mean():5.2f}  p10 {np.percentile(d,10):5.2f}  p90 {np.percentile(d,90):5.2f} us")
    last_store = t[:, 2, :].max(axis=1, keepdims=True)
    first_post = t[:, 0, :].min(axis=1, keepdims=True)
    print("  post spread (last-first post): %.2f us;  store spread: %.2f us" % (
        (t[:, 0, :].max(1) - t[:, 0, :].min(1)).mean() / 1e3, (t[:, 2, :].max(1) - t[:, 2, :].min(1)).mean() / 1e3))
    print("  last store -> w0_has_all: mean %.2f us;  last store -> resume: mean %.2f us" % (
        ((t[:, 3, :] - last_store).mean()) / 1e3, ((t[:, 5, :] - last_store).mean()) / 1e3))
    print("  cycle (post g -> post g+1, CTA 0): %.2f us; resume -> next post: %.2f us" % (
        np.diff(t[:, 0, 0]).mean() / 1e3, ((t[1:, 0, :] - t[:-1, 5, :]).mean()) / 1e3))
