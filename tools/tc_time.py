"""Time only the tcgen05 search kernel chain for a shape (bring-up aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gesture2vec_b200 as g
from gesture2vec_b200 import _lib
N, K, D = (int(x) for x in sys.argv[1:4])
dev = torch.device("cuda:0")
z = torch.randn(N, D, device=dev); E = torch.randn(K, D, device=dev)
cb = g.prepare_codebook(E)
flags = _lib.ALGO_TC | (_lib.NO_RECHECK if len(sys.argv) > 4 and sys.argv[4] == "fast" else 0)
for _ in range(3): g.vq_search(z, E, cb, flags=flags)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): g.vq_search(z, E, cb, flags=flags)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"N={N} K={K} D={D} dbg={os.environ.get('G2V_TC_DEBUG','0')} cg={os.environ.get('G2V_TC_CG','auto')}: {ms:.3f} ms/search  {2*N*K*D/ms/1e9:.0f} TFLOP/s")
