#!/usr/bin/env python
"""Golden vectors for the k-means and batched-tokenisation rows.

Run in the authoring container only:   python tests/golden/make_kmeans_golden.py

 * kmeans_lloyd.npz   -- sklearn.cluster.KMeans (the reference's third-party dependency) fitted from a
                         given init on seeded blobs: centres, labels, inertia, n_iter_, predict() of held-out rows.
 * tokenize_loop.npz  -- the reference's own per-chunk loop (Clustering.py:138-156): the real
                         DAE_model.VQ_Payam imported from /root/reference, called with a batch of ONE per chunk,
                         ids = np.argmax(encodings, axis=1).
Inputs are regenerated from seeds by the tests (oracle.kmeans_oracle.synth_blobs / numpy PCG64).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
sys.path.insert(0, HERE)

from oracle.kmeans_oracle import synth_blobs  # noqa: E402

KM_CASES = {"blobs_small": dict(n=3000, d=8, k=12, seed=7, max_iter=50, tol=1e-4),
            "blobs_latent": dict(n=6000, d=400, k=30, seed=11, max_iter=60, tol=1e-6)}
TOK = dict(n_chunks=257, n_layers=2, hidden=200, K=400, seed=21)


def tok_inputs():
    rng = np.random.default_rng(TOK["seed"])
    hidden = np.tanh(0.8 * rng.standard_normal((TOK["n_layers"], TOK["n_chunks"], TOK["hidden"]))).astype(np.float32)
    E = rng.uniform(-1, 1, size=(TOK["K"], TOK["n_layers"] * TOK["hidden"])).astype(np.float32)
    return hidden, E


def main():
    import sklearn
    from sklearn.cluster import KMeans
    out = {"sklearn_version": np.array(sklearn.__version__)}
    for name, c in KM_CASES.items():
        X, init = synth_blobs(c["n"], c["d"], c["k"], c["seed"])
        Xh, _ = synth_blobs(500, c["d"], c["k"], c["seed"] + 1000)
        km = KMeans(n_clusters=c["k"], init=init, n_init=1, max_iter=c["max_iter"], tol=c["tol"], algorithm="lloyd").fit(X)
        out[f"{name}_centers"] = km.cluster_centers_.astype(np.float32)
        out[f"{name}_labels"] = km.labels_.astype(np.int32)
        out[f"{name}_inertia"] = np.float64(km.inertia_)
        out[f"{name}_n_iter"] = np.int32(km.n_iter_)
        out[f"{name}_predict_heldout"] = km.predict(Xh).astype(np.int32)
        out[f"{name}_x_sum"] = np.float64(X.astype(np.float64).sum())
        print(name, "n_iter", km.n_iter_, "inertia", km.inertia_, "min cluster", np.bincount(km.labels_).min())
    np.savez_compressed(os.path.join(HERE, "kmeans_lloyd.npz"), **out)

    # ---- the reference's per-chunk tokenisation loop ----
    sys.path.insert(0, "/root/reference/scripts")
    import model.DAE_model as dae
    hidden, E = tok_inputs()
    layer = dae.VQ_Payam(TOK["K"], TOK["n_layers"] * TOK["hidden"], 0.25).eval()
    with torch.no_grad():
        layer._embedding.weight.copy_(torch.from_numpy(E))
    ids = []
    with torch.no_grad():
        for b in range(TOK["n_chunks"]):
            decoder_hidden = torch.from_numpy(hidden[:, b:b + 1, :]).contiguous()   # [n_layers, 1, hidden] as the encoder returns it
            _, _, _, encodings = layer(decoder_hidden)
            ids.append(np.argmax(encodings.detach().cpu().numpy(), axis=1))   # shape [1], int64
    ids = np.concatenate(ids)
    np.savez_compressed(os.path.join(HERE, "tokenize_loop.npz"), ids=ids.astype(np.int64),
                        hidden_sum=np.float64(hidden.astype(np.float64).sum()), E_sum=np.float64(E.astype(np.float64).sum()))
    print("tokenize_loop", ids.shape, ids[:8])


if __name__ == "__main__":
    main()
