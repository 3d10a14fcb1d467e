#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ by running the REAL reference.

Run in the authoring container only (needs /root/reference):

    python tests/golden/make_golden.py

It imports the reference quantizer modules unmodified (DAE_model.py imports only
torch; Autoencoder_VQVAE_model.py needs a 3-line `configargparse` shim, SURVEY.md
§8c), feeds them seeded synthetic inputs, and stores what they return: indices,
loss, perplexity, gradients and the EMA state after several training steps.

Inputs, initial codebooks and pre_linear weights are regenerated from numpy PCG64
seeds (tests/golden/cases.py) and only their fp64 checksums are stored; results are
stored as full index vectors, scalars, and row subsamples + fp64 checksums of the
large float arrays, which keeps every fixture under ~100 KiB.
"""
from __future__ import annotations

import argparse
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/scripts"
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
sys.path.insert(0, HERE)


def load_reference():
    shim = types.ModuleType("configargparse")
    shim.argparse = argparse
    sys.modules.setdefault("configargparse", shim)
    sys.path.insert(0, REF)
    import model.DAE_model as dae                      # noqa: E402
    import model.Autoencoder_VQVAE_model as vqvae      # noqa: E402
    vqvae.debug = False
    return dae, vqvae


from cases import CASES, BETA, EPS, G_LOSS, regen, checksums  # noqa: E402


def sub(a, rows):
    """row subsample + fp64 checksum (keeps the fixtures small)."""
    a2 = a.reshape(-1, a.shape[-1])
    return a2[rows].copy(), np.float64(a.astype(np.float64).sum()), np.float64(np.abs(a).astype(np.float64).sum())


def put(out, key, a, rows):
    out[key + "_rows"], out[key + "_sum"], out[key + "_abssum"] = sub(a, rows)


def main():
    dae, vqvae = load_reference()
    torch.manual_seed(0)
    torch.set_num_threads(1)  # fixed reduction order for the fixtures
    for ci, (name, flavour, cls, shape, K, D, lkind, ckind, decay) in enumerate(CASES):
        mod = dae if flavour == "dae" else vqvae
        r = regen(name)
        ema = r["ema"]
        layer = getattr(mod, cls)(K, D, BETA, decay) if ema else getattr(mod, cls)(K, D, BETA)
        with torch.no_grad():
            layer._embedding.weight.copy_(torch.from_numpy(r["E0"]))
            if ema:
                layer._ema_w.copy_(torch.from_numpy(r["ema_w0"]))
            if hasattr(layer, "pre_linear"):
                layer.pre_linear.weight.copy_(torch.from_numpy(r["pre_W"]))
                layer.pre_linear.bias.copy_(torch.from_numpy(r["pre_b"]))
        out = {"chk_" + k: np.float64(v) for k, v in checksums(r).items()}
        n = int(np.prod(shape)) // D
        krows = np.sort(np.random.default_rng(7).choice(K, size=min(K, 24), replace=False))
        nrows = np.sort(np.random.default_rng(8).choice(n, size=min(n, 24), replace=False))
        out["krows"], out["nrows"] = krows, nrows
        layer.train()
        for s in range(r["steps"]):
            xt = torch.from_numpy(r[f"x{s}"]).requires_grad_(True)
            loss, quant, ppl, enc = layer(xt)
            (loss * G_LOSS + (quant * torch.from_numpy(r[f"g{s}"])).sum()).backward()
            assert enc.sum().item() == enc.shape[0] and tuple(enc.shape) == (n, K)
            assert quant.shape == xt.shape and quant.is_contiguous()
            out[f"s{s}_idx"] = torch.argmax(enc, 1).numpy().astype(np.int16)
            out[f"s{s}_loss"] = np.float32(loss.item())
            out[f"s{s}_ppl"] = np.float32(ppl.item())
            put(out, f"s{s}_quant", quant.detach().numpy(), nrows)
            put(out, f"s{s}_gx", xt.grad.numpy(), nrows)
            if not ema:
                put(out, f"s{s}_gE", layer._embedding.weight.grad.numpy(), krows)
                layer._embedding.weight.grad = None
            else:
                assert layer._embedding.weight.grad is None
                assert layer.pre_linear.weight.grad is None     # SURVEY §8 a10 probe
                put(out, f"s{s}_E", layer._embedding.weight.detach().numpy(), krows)
                put(out, f"s{s}_ema_w", layer._ema_w.detach().numpy(), krows)
                out[f"s{s}_cs"] = layer._ema_cluster_size.numpy().copy()
        if ema:  # eval mode must leave the EMA state untouched
            layer.eval()
            before = [layer._embedding.weight.detach().clone(), layer._ema_w.detach().clone(),
                      layer._ema_cluster_size.clone()]
            with torch.no_grad():
                loss, quant, ppl, enc = layer(torch.from_numpy(r[f"x{r['steps']}"]))
            after = [layer._embedding.weight, layer._ema_w, layer._ema_cluster_size]
            assert all(torch.equal(a, b) for a, b in zip(before, after))
            out.update(eval_idx=torch.argmax(enc, 1).numpy().astype(np.int16),
                       eval_loss=np.float32(loss.item()), eval_ppl=np.float32(ppl.item()))
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(f"{name:22s} {flavour:6s} {cls:13s} {os.path.getsize(path)/1024:8.1f} KiB")

    # state_dict key contract of every reference class flavour
    keys = {}
    for flavour, mod in (("dae", dae), ("vqvae", vqvae)):
        keys[f"{flavour}.VQ_Payam"] = sorted(mod.VQ_Payam(8, 4, 0.25).state_dict().keys())
        keys[f"{flavour}.VQ_Payam_EMA"] = sorted(mod.VQ_Payam_EMA(8, 4, 0.25, 0.9).state_dict().keys())
    keys["vqvae.VectorQuantizerEMA"] = sorted(vqvae.VectorQuantizerEMA(8, 4, 0.25, 0.9).state_dict().keys())
    keys["vqvae.VQ_Payam_GSSoft"] = sorted(vqvae.VQ_Payam_GSSoft(8, 4, 0.25).state_dict().keys())
    import json
    with open(os.path.join(HERE, "state_dict_keys.json"), "w") as f:
        json.dump(keys, f, indent=1, sort_keys=True)

    # VectorQuantizerEMA (hstack adapter, Autoencoder_VQVAE_model.py:1745-1812), one train step
    K, D, B = 48, 16, 20
    layer = vqvae.VectorQuantizerEMA(K, D, BETA, 0.9)
    rng = np.random.default_rng(4242)
    x = rng.standard_normal((2, B, D // 2), dtype=np.float32)
    E0 = layer._embedding.weight.detach().numpy().copy()
    w0 = layer._ema_w.detach().numpy().copy()
    xt = torch.from_numpy(x).requires_grad_(True)
    layer.train()
    loss, quant, ppl, enc = layer(xt)
    g_out = rng.standard_normal(tuple(quant.shape), dtype=np.float32)
    (loss + (quant * torch.from_numpy(g_out)).sum()).backward()
    np.savez_compressed(
        os.path.join(HERE, "vqvae_hstack_ema.npz"), x=x, E0=E0, ema_w0=w0, g_out=g_out,
        pre_W=layer.pre_lin.weight.detach().numpy(), pre_b=layer.pre_lin.bias.detach().numpy(),
        idx=torch.argmax(enc, 1).numpy().astype(np.int32), loss=np.float32(loss.item()),
        ppl=np.float32(ppl.item()), quant=quant.detach().numpy(), gx=xt.grad.numpy(),
        gW=layer.pre_lin.weight.grad.numpy(), gb=layer.pre_lin.bias.grad.numpy(),
        E1=layer._embedding.weight.detach().numpy(), ema_w1=layer._ema_w.detach().numpy(),
        cs1=layer._ema_cluster_size.numpy(), beta=np.float32(BETA), decay=np.float64(0.9),
        eps=np.float64(1e-5))
    print("vqvae_hstack_ema")


if __name__ == "__main__":
    main()
