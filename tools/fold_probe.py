"""Where the time of the pre_linear-folded flavour goes (1 M rows, K=512): pad, folded search, step, fold rebuild."""
import os, sys, time
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

def timed(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

E_fold, cb_fold, Wp, bp = layer._fold(dev)
stats = torch.zeros(8, dtype=torch.int64, device=dev)
print("pad ms", timed(lambda: torch.nn.functional.pad(x, (0, 4))))
xp = torch.nn.functional.pad(x, (0, 4))
print("folded search ms", timed(lambda: F.vq_search(xp, E_fold, cb_fold, stats=stats)), "stats", stats.tolist())
print("wide search (no copy) ms", timed(lambda: F.vq_search_wide(x, E_fold, cb_fold)))
print("own pad kernel ms", timed(lambda: F.pad_rows(x, D + 4)))
E = layer._embedding.weight.detach()
zs = F.gemm(x, Wp, bias=bp)
stats.zero_()
print("explicit projection ms", timed(lambda: F.gemm(x, Wp, bias=bp)))
print("projected search ms", timed(lambda: F.vq_search(zs, E, stats=stats)), "stats", stats.tolist())
def rebuild():
    layer._fold_key = None
    layer._fold(dev)
print("fold rebuild ms", timed(rebuild))
xs = x.clone().requires_grad_(True)
gq = torch.randn(N, D, device=dev, generator=gen)
def step():
    xs.grad = None
    loss, q, ppl, _ = layer(xs)
    torch.autograd.backward([loss, q], [torch.ones_like(loss), gq])
ids = layer.tokenize(x)
def step_given():
    xs.grad = None
    loss, q, ppl, _ = layer.forward_with_indices(xs, ids)
    torch.autograd.backward([loss, q], [torch.ones_like(loss), gq])
def fwd_only():
    with torch.no_grad():
        layer(x)
for fold in (True, False):
    layer.fold_projection = fold
    print("step ms fold=%s" % fold, timed(step, 5), "| with indices given", timed(step_given, 5), "| forward only", timed(fwd_only, 5))
import cProfile, pstats
layer.fold_projection = True
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable(); step(); torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
