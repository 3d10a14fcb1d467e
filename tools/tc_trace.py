"""Dump a role-level timeline of CTA 0 of the tcgen05 search kernel (bring-up aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
N, K, D = (int(x) for x in sys.argv[1:4])
dev = torch.device("cuda:0")
trace = torch.zeros(5 * 512 * 2, dtype=torch.int64, device=dev)
os.environ["G2V_TC_TRACE"] = hex(trace.data_ptr())
import gesture2vec_b200 as g
from gesture2vec_b200 import _lib
z = torch.randn(N, D, device=dev); E = torch.randn(K, D, device=dev)
cb = g.prepare_codebook(E)
g.vq_search(z, E, cb, flags=_lib.ALGO_TC | _lib.NO_RECHECK)
torch.cuda.synchronize()
trace.zero_()
g.vq_search(z, E, cb, flags=_lib.ALGO_TC | _lib.NO_RECHECK)
torch.cuda.synchronize()
t = trace.cpu().view(5, 512, 2).numpy()
t0 = min(int(t[r, 0, 1]) for r in range(5) if t[r, 0, 1] > 0)
names = ["Bprod", "MMA", "Aprod", "EPI", "CONV"]
ev = []
for r in range(5):
    for i in range(512):
        if t[r, i, 1] > 0:
            ev.append((int(t[r, i, 1]) - t0, names[r], int(t[r, i, 0])))
ev.sort()
lo, hi = int(sys.argv[4]) if len(sys.argv) > 4 else 0, int(sys.argv[5]) if len(sys.argv) > 5 else 10**9
for e in ev:
    if lo <= e[0] <= hi:
        print(f"{e[0]:9d} ns  {e[1]:5s} {e[2]}")
