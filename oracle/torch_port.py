"""Torch-CPU port of the reference quantizers.  TEST / BASELINE INFRASTRUCTURE ONLY.

The reference is pure PyTorch, so the faithful "reference CPU implementation"
of this path is the same sequence of eager torch ops (two [N,K] fp32
intermediates, three GEMMs) on the host cores.  ``/root/reference`` does not
travel to the GPU box, so ``bench.py``'s ``cpu_baseline`` leg and its
``--impl reference`` arm time THIS port (kind = "port").  It is validated
against the real reference modules in ``tests/golden/make_golden.py`` (run in
the authoring container) and against ``tests/golden/*.npz`` everywhere else.

Never imported by the product package.

Reference statements followed: DAE_model.py:301-348 (hard), :396-482 (EMA);
Autoencoder_VQVAE_model.py:1217-1296 (EMA searching on pre_linear(z)).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _search(flat, E):
    # DAE_model.py:320-327 -- fp32, materialises the [N,K] distance matrix
    d = (flat.pow(2).sum(1, keepdim=True) + E.pow(2).sum(1)) - 2 * torch.matmul(flat, E.t())
    return torch.argmin(d, dim=1).unsqueeze(1)


def _tail(inputs, quantized, enc, coef_commit, coef_codebook):
    # DAE_model.py:340-347 / :474-481
    e_loss = F.mse_loss(quantized.detach(), inputs)
    loss = coef_commit * e_loss
    if coef_codebook:
        loss = coef_codebook * F.mse_loss(quantized, inputs.detach()) + loss
    quantized = inputs + (quantized - inputs).detach()
    p = enc.mean(0)
    perplexity = torch.exp(-(p * torch.log(p + 1e-10)).sum())
    return loss, quantized.contiguous(), perplexity, enc


class PortVQ(torch.nn.Module):
    """Same op sequence as reference VQ_Payam."""

    def __init__(self, K, D, beta):
        super().__init__()
        self.K, self.D, self.beta = K, D, beta
        self._embedding = torch.nn.Embedding(K, D)
        self._embedding.weight.data.uniform_(-1 / K, 1 / K)

    def forward(self, inputs):
        flat = inputs.view(-1, self.D)
        E = self._embedding.weight
        idx = _search(flat, E)
        enc = torch.zeros(idx.shape[0], self.K, device=inputs.device)
        enc.scatter_(1, idx, 1)
        q = torch.matmul(enc, E).reshape(inputs.shape).contiguous()
        return _tail(inputs, q, enc, self.beta, 1.0)


class PortVQEMA(torch.nn.Module):
    """Same op sequence as reference VQ_Payam_EMA (flavour 'dae' or 'vqvae')."""

    def __init__(self, K, D, beta, decay, eps=1e-5, flavour="dae"):
        super().__init__()
        self.K, self.D, self.beta, self.decay, self.eps = K, D, beta, decay, eps
        self.flavour = flavour
        self.pre_linear = torch.nn.Linear(D, D)
        self._embedding = torch.nn.Embedding(K, D)
        if flavour == "dae":
            self._embedding.weight.data.uniform_(-1 / K, 1 / K)
        else:
            self._embedding.weight.data.uniform_(-1, 1)
        self.register_buffer("_ema_cluster_size", torch.zeros(K))
        self._ema_w = torch.nn.Parameter(torch.randn(K, D))

    def forward(self, inputs):
        flat = inputs.view(-1, self.D)
        if self.flavour == "vqvae":
            flat = self.pre_linear(flat)
        E = self._embedding.weight
        idx = _search(flat, E)
        enc = torch.zeros(idx.shape[0], self.K, device=inputs.device)
        enc.scatter_(1, idx, 1)
        q = torch.matmul(enc, E).reshape(inputs.shape).contiguous()
        if self.training:
            cs = self._ema_cluster_size * self.decay + (1 - self.decay) * enc.sum(0)
            n = cs.data.sum()
            self._ema_cluster_size = (cs + self.eps) / (n + self.K * self.eps) * n
            dw = torch.matmul(enc.t(), flat)
            self._ema_w = torch.nn.Parameter(self._ema_w * self.decay + (1 - self.decay) * dw)
            self._embedding.weight = torch.nn.Parameter(
                self._ema_w / self._ema_cluster_size.unsqueeze(1))
        return _tail(inputs, q, enc, self.beta, 0.0)


def tokenize_blocks(z: torch.Tensor, module: torch.nn.Module, block: int = 65536) -> torch.Tensor:
    """The reference's tokenisation consumer (Clustering.py:152-157: forward, then
    argmax of the one-hot), restated over blocks of rows."""
    out = []
    with torch.no_grad():
        for s in range(0, z.shape[0], block):
            _, _, _, enc = module(z[s:s + block])
            out.append(torch.argmax(enc, dim=1))
    return torch.cat(out)
