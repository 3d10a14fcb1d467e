"""The N = 128 training step of config/VQ-VAE.yml a few times (for an ncu launch list of the small-batch path)."""
import os
import sys

import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import gesture2vec_b200 as g  # noqa: E402

dev = torch.device("cuda:0")
layer = g.DAE_VQ_Payam_EMA(512, 400, 0.25, 0.85).to(dev).train()
with torch.no_grad():
    layer._embedding.weight.uniform_(-1, 1)
x = torch.tanh(0.8 * torch.randn(2, 128, 200, device=dev)).requires_grad_(True)
gq = torch.randn(2, 128, 200, device=dev)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    x.grad = None
    loss, q, ppl, enc = layer(x)
    torch.autograd.backward([loss, q], [torch.ones_like(loss), gq])
torch.cuda.synchronize()
print("ok", float(loss))
