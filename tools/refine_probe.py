"""Search time with and without the refine pass (whole-row re-ranks on the tensor cores at fp32 accuracy)."""
import os, sys
import torch
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import gesture2vec_b200 as g
from gesture2vec_b200 import _lib as L
import gpu_synth as S

dev = torch.device("cuda:0")

def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

for cbk, lat, K, dt in [("normal", "iid", 400, torch.float32), ("normal", "iid", 512, torch.float32), ("normal", "iid", 512, torch.bfloat16),
                        ("ema_degenerate", "gru", 512, torch.float32), ("normal", "iid", 2048, torch.float32),
                        ("normal", "iid", 16384, torch.float32)]:
    N, D = 1_000_000, 400
    E = S.codebook(cbk, K, D, dev, seed=3)
    z = S.latents(lat, N, D, dev, E=E, seed=4).to(dt)
    cb = g.prepare_codebook(E)
    out = []
    for name, fl in (("refine", 0), ("fp64", L.NO_REFINE)):
        st = torch.zeros(8, dtype=torch.int64, device=dev)
        idx = g.vq_search(z, E, cb, flags=fl, stats=st)
        ms = t(lambda: g.vq_search(z, E, cb, flags=fl))
        out.append((name, ms, st.tolist()[:6], idx))
    same = bool(torch.equal(out[0][3], out[1][3]))
    print(f"{cbk}/{lat} K={K} {str(dt)[6:]}: " + " | ".join(f"{n} {ms:.3f} ms stats {s}" for n, ms, s, _ in out) + f" | identical {same}", flush=True)
    del z
