"""Folded vs projected search on the evolving EMA codebook of VQVAE_VQ_Payam_EMA: per-step search time, re-rank counters, fold statistics."""
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

def t(fn):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); r = fn(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1), r

for step in range(8):
    E_fold, cb_fold, Wp, bp = layer._fold(dev)
    stats = torch.zeros(8, dtype=torch.int64, device=dev)
    ms, _ = t(lambda: F.vq_search_wide(x, E_fold, cb_fold, stats=stats))
    E = layer._embedding.weight.detach()
    zs = F.gemm(x, Wp, bias=bp)
    st2 = torch.zeros(8, dtype=torch.int64, device=dev)
    ms2, _ = t(lambda: F.vq_search(zs, E, stats=st2))
    print(f"step {step}: folded search {ms:.2f} ms stats {stats.tolist()[:4]} | projected search {ms2:.2f} ms stats {st2.tolist()[:4]} | "
          f"|E| max {float(E.abs().max()):.3g} t max {float(E_fold[:, D].max()):.3g} t median {float(E_fold[:, D].median()):.3g} "
          f"live codes {int((layer._ema_cluster_size > 1e-3).sum())}")
    ms3, _ = t(lambda: layer(x))
    print(f"         layer(x) {ms3:.2f} ms")
