"""CPU oracle for the k-means rows of the hot path.  TEST INFRASTRUCTURE ONLY.

numpy restatement of the Lloyd loop the reference runs through a third-party dependency,
``sklearn.cluster.KMeans`` (requirements.txt: scikit-learn==1.2.2; call sites Clustering.py:718-720,
train_DAE.py:257-263, lmdb_data_loader.py:1288-1291).  sklearn is not part of /root/reference, so the
published algorithm is restated here -- ``sklearn/cluster/_kmeans.py``: ``_kmeans_single_lloyd`` (loop and
stopping rules), ``_tolerance`` (tol * mean(var(X, axis=0))), ``lloyd_iter_chunked_dense`` (assignment by
argmin ||x - c||^2, centres = per-cluster means) -- for the case the product covers: dense rows, unit sample
weights, one run from a given init, no cluster ever empty (sklearn would relocate an empty cluster; the
product keeps its centre).

Parity status: PINNED against sklearn itself (version 1.9.0 in the authoring container; the Lloyd loop is
unchanged since 1.2.2): tests/golden/make_kmeans_golden.py fits ``KMeans(init=array, n_init=1,
algorithm="lloyd")`` and stores centres, labels, inertia and n_iter_ in tests/golden/kmeans_lloyd.npz;
tests/test_oracle_golden.py replays them against this file.  Only tests/ may import it.
"""
from __future__ import annotations

import numpy as np


def assign(X: np.ndarray, C: np.ndarray) -> np.ndarray:
    """argmin_k ||x - c_k||^2 in fp64, first index on ties."""
    X64, C64 = X.astype(np.float64), C.astype(np.float64)
    d = (X64 * X64).sum(1)[:, None] - 2.0 * (X64 @ C64.T) + (C64 * C64).sum(1)[None, :]
    return np.argmin(d, axis=1).astype(np.int32)


def tolerance(X: np.ndarray, tol: float) -> float:
    """sklearn _tolerance(): mean of the per-feature variances times tol."""
    return float(np.mean(np.var(X.astype(np.float64), axis=0)) * tol)


def lloyd(X: np.ndarray, init: np.ndarray, max_iter: int = 300, tol: float = 1e-4):
    """_kmeans_single_lloyd: returns (centres fp32, labels, inertia, n_iter)."""
    X = np.ascontiguousarray(X, dtype=np.float32)
    C = np.ascontiguousarray(init, dtype=np.float32).copy()
    K = C.shape[0]
    tol_abs = tolerance(X, tol)
    labels_old = np.full(X.shape[0], -1, dtype=np.int32)
    strict = False
    n_iter = 0
    for i in range(max_iter):
        labels = assign(X, C)
        C_new = C.copy()
        for k in range(K):
            m = labels == k
            if m.any():
                C_new[k] = X[m].astype(np.float64).mean(0).astype(np.float32)
        shift2 = float(((C_new.astype(np.float64) - C.astype(np.float64)) ** 2).sum())
        C = C_new
        n_iter = i + 1
        if np.array_equal(labels, labels_old):
            strict = True
            break
        if shift2 <= tol_abs:
            break
        labels_old = labels
    if not strict:
        labels = assign(X, C)
    inertia = float(((X.astype(np.float64) - C.astype(np.float64)[labels]) ** 2).sum())
    return C, labels, inertia, n_iter


def synth_blobs(n: int, d: int, k: int, seed: int, spread: float = 1.0, init_noise: float = 0.8):
    """Overlapping Gaussian blobs and an init of one perturbed true centre per blob: Lloyd needs several
    iterations, but every blob keeps exactly one centre, so the trajectory does not hinge on how near-ties
    round (an init of random rows splits blobs between centres and then does)."""
    rng = np.random.default_rng(seed)
    centres = rng.standard_normal((k, d)).astype(np.float32) * 2.0
    lab = rng.integers(0, k, size=n)
    X = (centres[lab] + spread * rng.standard_normal((n, d))).astype(np.float32)
    init = (centres[rng.permutation(k)] + init_noise * rng.standard_normal((k, d))).astype(np.float32)
    return X, init
