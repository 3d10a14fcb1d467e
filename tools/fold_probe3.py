"""Kernel-level timeline of the folded search on a degenerate EMA codebook (torch.profiler / CUPTI)."""
import os, sys
import torch
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import gesture2vec_b200 as g
from gesture2vec_b200 import functional as F

dev = torch.device("cuda:0")
K, D, N = 512, 400, 1_000_000
gen = torch.Generator(device=dev).manual_seed(1)
layer = g.VQVAE_VQ_Payam_EMA(K, D, 0.25, 0.85).to(dev).train()
layer.return_encodings = False
x = torch.tanh(0.8 * torch.randn(N, D, device=dev, generator=gen))
for _ in range(3):
    layer(x)
E_fold, cb_fold, Wp, bp = layer._fold(dev)
stats = torch.zeros(8, dtype=torch.int64, device=dev)
F.vq_search_wide(x, E_fold, cb_fold, stats=stats)
torch.cuda.synchronize()
print("folded stats", stats.tolist())
E = layer._embedding.weight.detach()
zs = F.gemm(x, Wp, bias=bp)
st2 = torch.zeros(8, dtype=torch.int64, device=dev)
F.vq_search(zs, E, stats=st2)
print("projected stats", st2.tolist())
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    F.vq_search_wide(x, E_fold, cb_fold)
    torch.cuda.synchronize()
    F.vq_search(zs, E)
    torch.cuda.synchronize()
for ev in prof.events():
    if ev.device_time > 0:
        print(f"{ev.name[:70]:70s} {ev.device_time:10.1f} us")
