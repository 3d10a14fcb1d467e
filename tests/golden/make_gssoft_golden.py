#!/usr/bin/env python
"""Golden vectors for the soft quantizer from the REAL reference module.

Run in the authoring container only:   python tests/golden/make_gssoft_golden.py
Imports Autoencoder_VQVAE_model.VQ_Payam_GSSoft unmodified (configargparse shim as in make_golden.py), sets seeded
parameters, runs forward and autograd on a Trinity-shaped input, and stores outputs + gradients (row subsamples of the
large arrays + fp64 checksums).  Inputs / parameters are regenerated from seeds by the tests (`inputs()`).
"""
from __future__ import annotations

import argparse
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
CFG = dict(K=512, D=400, shape=(2, 64, 200), beta=0.25, seed=31, g_loss=3.0)
ROWS = np.array([0, 1, 7, 31, 63])


def inputs():
    c, rng = CFG, np.random.default_rng(CFG["seed"])
    K, D = c["K"], c["D"]
    x = np.tanh(0.8 * rng.standard_normal(c["shape"])).astype(np.float32)
    E = rng.standard_normal((K, D)).astype(np.float32)
    b = 1.0 / np.sqrt(D)
    Wm = rng.uniform(-b, b, (D, D)).astype(np.float32); bm = rng.uniform(-b, b, D).astype(np.float32)
    Wl = rng.uniform(-b, b, (K, D)).astype(np.float32) * 0.2; bl = rng.uniform(-b, b, K).astype(np.float32)
    g_out = rng.standard_normal(c["shape"]).astype(np.float32)
    return x, E, Wm, bm, Wl, bl, g_out


def main():
    shim = types.ModuleType("configargparse"); shim.argparse = argparse
    sys.modules.setdefault("configargparse", shim)
    sys.path.insert(0, "/root/reference/scripts")
    import model.Autoencoder_VQVAE_model as vqvae
    vqvae.debug = False
    c = CFG
    x, E, Wm, bm, Wl, bl, g_out = inputs()
    layer = vqvae.VQ_Payam_GSSoft(c["K"], c["D"], c["beta"])
    with torch.no_grad():
        layer._embedding.weight.copy_(torch.from_numpy(E))
        layer.mean_layer.weight.copy_(torch.from_numpy(Wm)); layer.mean_layer.bias.copy_(torch.from_numpy(bm))
        layer.logvar_layer.weight.copy_(torch.from_numpy(Wl)); layer.logvar_layer.bias.copy_(torch.from_numpy(bl))
    xt = torch.from_numpy(x).requires_grad_(True)
    loss, out, ppl, enc = layer(xt)
    (c["g_loss"] * loss + (out * torch.from_numpy(g_out)).sum()).backward()
    res = dict(loss=np.float32(loss.item()), perplexity=np.float32(ppl.item()),
               out_rows=out.detach().numpy().reshape(-1, c["D"])[ROWS], enc_rows=enc.detach().numpy()[ROWS],
               enc_sum=np.float64(enc.detach().double().sum()), out_sum=np.float64(out.detach().double().sum()),
               gx_rows=xt.grad.numpy().reshape(-1, c["D"])[ROWS], gx_abssum=np.float64(xt.grad.double().abs().sum()),
               gE_rows=layer._embedding.weight.grad.numpy()[ROWS], gE_abssum=np.float64(layer._embedding.weight.grad.double().abs().sum()),
               gWm_rows=layer.mean_layer.weight.grad.numpy()[ROWS], gbm=layer.mean_layer.bias.grad.numpy(),
               gWl_rows=layer.logvar_layer.weight.grad.numpy()[ROWS], gbl=layer.logvar_layer.bias.grad.numpy(),
               x_sum=np.float64(x.astype(np.float64).sum()), state_keys=np.array(sorted(layer.state_dict().keys())))
    np.savez_compressed(os.path.join(HERE, "gssoft_trinity.npz"), **res)
    print("loss", res["loss"], "ppl", res["perplexity"], "max p", enc.max().item(), list(res["state_keys"]))


if __name__ == "__main__":
    main()
