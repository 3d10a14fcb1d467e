"""Bring-up aid for the tcgen05 search: compares it with the fp32 SIMT search and the fp64 oracle
on a ladder of shapes and prints where they differ.  Run on the GPU box."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import gesture2vec_b200 as g
from gesture2vec_b200 import _lib
from oracle import vq_oracle as O

dev = torch.device("cuda:0")
shapes = [(128, 256, 64), (128, 256, 128), (128, 256, 80), (256, 512, 64), (1000, 512, 400), (4096, 400, 400),
          (300, 1000, 400), (5000, 2048, 448), (70000, 512, 400), (20000, 4096, 400)]
if len(sys.argv) > 1:
    shapes = [tuple(int(x) for x in a.split(",")) for a in sys.argv[1:]]
for (N, K, D) in shapes:
    E = O.synth_codebook("normal", K, D, seed=3)
    z = O.synth_latents("iid", N, D, seed=11)
    zt, Et = torch.from_numpy(z).to(dev), torch.from_numpy(E).to(dev)
    st = torch.zeros(8, dtype=torch.int64, device=dev)
    t0 = time.time()
    a = g.vq_search(zt, Et, flags=_lib.ALGO_TC, stats=st)
    torch.cuda.synchronize()
    dt = time.time() - t0
    fast = g.vq_search(zt, Et, flags=_lib.ALGO_TC | _lib.NO_RECHECK).cpu().numpy()
    b = g.vq_search(zt, Et, flags=_lib.ALGO_SIMT).cpu().numpy()
    a = a.cpu().numpy()
    ref = O.nearest_code_f64(z[:20000], E)
    bad = np.nonzero(a != b)[0]
    badf = np.nonzero(fast != b)[0]
    print(f"N={N} K={K} D={D}: tc!=simt {bad.size}  tc_fast!=simt {badf.size}  tc!=f64(first 20000) "
          f"{int((a[:20000] != ref).sum())}  stats(pair,full,fallback)={st.cpu().numpy()[1:4].tolist()}  {dt*1e3:.1f} ms", flush=True)
    if badf.size > N // 20:
        d = O.distances_f64(z[:4], E)
        srt = np.argsort(d, 1)[:, :4]
        print("  first rows: fast", fast[:4].tolist(), "simt", b[:4].tolist(), "top4", srt.tolist())
        print("  rank of the fast choice in the true order:",
              [int(np.nonzero(np.argsort(d[i]) == fast[i])[0][0]) if fast[i] < K else -1 for i in range(4)])
        print("  histogram of fast idx (first 16 bins of 32):", np.bincount(fast // max(K // 32, 1), minlength=32)[:16].tolist())
