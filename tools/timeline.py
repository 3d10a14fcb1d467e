"""Per-kernel device timeline of the tokenization step, or with a 4th argument `train` of the fwd+bwd+EMA step of
the drop-in EMA module (torch.profiler / CUPTI: kernel durations and the idle gaps between consecutive kernels).
Usage: python tools/timeline.py [K] [f32|bf16] [flags] [train]"""
import os, sys
import torch
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import gesture2vec_b200 as g
import gpu_synth as S
from torch.profiler import profile, ProfilerActivity

K = int(sys.argv[1]) if len(sys.argv) > 1 else 400
dt = torch.bfloat16 if len(sys.argv) > 2 and sys.argv[2] == "bf16" else torch.float32
flags = int(sys.argv[3]) if len(sys.argv) > 3 else 0
dev = torch.device("cuda:0")
N, D = 1_000_000, 400
E = S.codebook("normal", K, D, dev, seed=3)
train = len(sys.argv) > 4 and sys.argv[4] == "train"
soft = len(sys.argv) > 4 and sys.argv[4] == "soft"
vqvae = len(sys.argv) > 4 and sys.argv[4] == "vqvae"
if vqvae:
    # VQVAE_VQ_Payam_EMA (pre_linear folded): latents whose projection is clustered around the codes
    layer = g.VQVAE_VQ_Payam_EMA(K, D, 0.25, 0.85).to(dev).train()
    with torch.no_grad():
        torch.nn.init.orthogonal_(layer.pre_linear.weight)
        t = S.latents("clustered", N, D, dev, E=E, seed=4)
        z0 = (t - layer.pre_linear.bias) @ layer.pre_linear.weight
        layer._embedding.weight.copy_(E)
        layer._ema_w.copy_(E * (N / K))
        layer._ema_cluster_size.fill_(N / K)
    layer.return_encodings = False
    zf = z0.contiguous().requires_grad_(True)
    gq = torch.randn(N, D, device=dev)

    def step(i):
        zf.grad = None
        loss, q, ppl, _ = layer(zf)
        torch.autograd.backward([loss, q], [torch.ones_like(loss), gq])
elif soft:
    N = 131072
    layer = g.VQVAE_VQ_Payam_GSSoft(K, D, 0.25).to(dev)
    xs = torch.tanh(0.8 * torch.randn(N, D, device=dev)).requires_grad_(True)
    gq = torch.randn(N, D, device=dev)

    def step(i):
        xs.grad = None
        layer.zero_grad(set_to_none=True)
        loss, q, ppl, p = layer(xs)
        torch.autograd.backward([loss, q], [torch.ones_like(loss), gq])
elif train:
    layer = g.DAE_VQ_Payam_EMA(K, D, 0.25, 0.85).to(dev).train()
    with torch.no_grad():
        layer._embedding.weight.copy_(E)
    layer.return_encodings = False
    zf = S.latents("clustered", N, D, dev, E=E, seed=4).requires_grad_(True)
    gq = torch.randn(N, D, device=dev)

    def step(i):
        zf.grad = None
        loss, q, ppl, _ = layer(zf)
        torch.autograd.backward([loss, q], [torch.ones_like(loss), gq])
else:
    zs = [S.latents("iid", N, D, dev, seed=4 + i).to(dt) for i in range(3)]
    cb = g.prepare_codebook(E)

    def step(i):
        g.vq_search(zs[i % 3], E, cb, flags=flags)
for i in range(3):
    step(i)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for i in range(9):
        step(i)
    torch.cuda.synchronize()
evs = sorted([e for e in prof.events() if e.device_time > 0], key=lambda e: e.time_range.start)
agg, gaps, prev_end = {}, {}, None
for e in evs:
    name = e.name.replace("(anonymous namespace)::", "").replace("void ", "").split("<")[0].split("(")[0].split("::")[-1][:40]
    agg.setdefault(name, []).append(e.device_time)
    if prev_end is not None:
        gaps.setdefault(name, []).append(e.time_range.start - prev_end)
    prev_end = e.time_range.end
span = evs[-1].time_range.end - evs[0].time_range.start
print(f"K={K} {dt} flags={flags}: {span / 9:.1f} us per step over 9 steps")
for k, v in agg.items():
    gp = gaps.get(k, [0])
    print(f"  {k:40s} n={len(v):3d} mean {sum(v) / len(v):8.1f} us   per step {sum(v) / 9:8.1f} us   gap before: mean {sum(gp) / len(gp):6.1f} us")
