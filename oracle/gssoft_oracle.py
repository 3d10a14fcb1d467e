"""CPU oracle for the soft quantizer VQ_Payam_GSSoft.  TEST INFRASTRUCTURE ONLY.

numpy restatement of ``Autoencoder_VQVAE_model.py:1304-1433`` -- the layer ``Autoencoder_VQVAE.__init__``
actually instantiates (``:816-820``) -- written from the source, with the closed-form backward a CUDA
implementation needs (the reference relies on autograd).  No product code exists for this row yet; the
oracle is here so the next round starts from a pinned checker.

Forward (``:1374-1433``, ``soft_prob :1349-1372``), all fp32:
    m  = mean_layer(z)                      z = inputs.view(-1, D)
    lv = logvar_layer(m)                    [N, K]
    d  = sum(m^2) + sum(E^2) - 2 m E^T      [N, K]
    s  = 1 / exp(lv)^2
    p~ = exp(-(d / 400) * 0.5 * s) / sqrt(s);   p = p~ / sum_k p~        (the literal 400 is hard-coded)
    q  = p E;  loss = mse(q, x) * (1 + beta) in value;  out = x + (q - x);  perplexity from mean_n p
Backward of  G*loss + <out, g_out>  (closed form; only mse(q, sg x) reaches the parameters):
    gq = 2 G (q - x) / M;   dE = p^T gq;   dp = gq E^T;   g = p * (dp - sum_k p dp)        (d log p~)
    gd = -g s / 800;        glv = g (d s / 400 + 1)
    gm = 2 m * sum_k gd - 2 gd E + glv Wl;        dE += 2 E * (sum_n gd)^T - 2 gd^T m
    dWl = glv^T m; dbl = sum glv;   dWm = gm^T z; dbm = sum gm;   gx = gm Wm + 2 G beta (x - q)/M + g_out

Parity status: PINNED against the real reference (forward values and autograd gradients):
tests/golden/make_gssoft_golden.py -> tests/golden/gssoft_trinity.npz, replayed by tests/test_oracle_golden.py.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32
SCALE = F32(400.0)          # Autoencoder_VQVAE_model.py:1351 `dist = (dist) / 400`


def forward(x, E, Wm, bm, Wl, bl, beta):
    x = np.ascontiguousarray(x, dtype=F32)
    D = E.shape[1]
    z = x.reshape(-1, D)
    m = (z @ Wm.T + bm).astype(F32)
    lv = (m @ Wl.T + bl).astype(F32)
    d = ((m * m).sum(1, keepdims=True, dtype=F32) + (E * E).sum(1, dtype=F32)) - F32(2.0) * (m @ E.T)
    s = (F32(1.0) / np.exp(lv) ** 2).astype(F32)
    pt = (np.exp(-(d / SCALE) * (F32(0.5) * s)) / np.sqrt(s)).astype(F32)
    p = (pt / pt.sum(1, keepdims=True, dtype=F32)).astype(F32)
    q = (p @ E).astype(F32)
    M = F32(z.size)
    mse = F32(((q - z) ** 2).sum(dtype=np.float64) / np.float64(M))
    loss = F32(mse + F32(beta) * mse)
    out = (z + (q - z)).reshape(x.shape)
    avg = p.mean(0, dtype=F32)
    ppl = F32(np.exp(-np.sum(avg * np.log(avg + F32(1e-10)), dtype=F32)))
    return dict(loss=loss, out=out, perplexity=ppl, encodings=p, m=m, lv=lv, d=d, s=s, q=q, z=z)


def backward(fw, E, Wm, Wl, beta, g_loss, g_out):
    """Gradients of g_loss*loss + sum(out*g_out) in fp64 (the reference's fp32 autograd is compared with a tolerance)."""
    f8 = np.float64
    z, m, d, s, p, q = (fw[k].astype(f8) for k in ("z", "m", "d", "s", "encodings", "q"))
    E8, Wm8, Wl8 = E.astype(f8), Wm.astype(f8), Wl.astype(f8)
    M = z.size
    gq = 2.0 * g_loss * (q - z) / M
    dE = p.T @ gq
    dp = gq @ E8.T
    g = p * (dp - (p * dp).sum(1, keepdims=True))
    gd = -g * s / 800.0
    glv = g * (d * s / 400.0 + 1.0)
    gm = 2.0 * m * gd.sum(1, keepdims=True) - 2.0 * gd @ E8 + glv @ Wl8
    dE += 2.0 * E8 * gd.sum(0)[:, None] - 2.0 * gd.T @ m
    gx = gm @ Wm8 + 2.0 * g_loss * beta * (z - q) / M + np.asarray(g_out, dtype=f8).reshape(z.shape)
    return dict(x=gx.reshape(fw["out"].shape), E=dE, Wm=gm.T @ z, bm=gm.sum(0), Wl=glv.T @ m, bl=glv.sum(0))
