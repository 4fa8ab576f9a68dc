import sys; import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nupgcm_b200.architectures import GPU
ctx=GPU(0).ctx
print("round trip us; rows: store kind, cols: load kind (0 relaxed.gpu 1 volatile 2 cg 3 atomic 4 acquire)")
for peer in (1, 2, 73, 74, 147):
    print("peer", peer)
    for sk in range(5):
        print("  st%d: "%sk + "  ".join("%6.2f"%ctx.pingpong(peer, 10*sk+lk, 3000) for lk in range(5)), flush=True)
