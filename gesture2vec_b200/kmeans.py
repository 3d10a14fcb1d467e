"""k-means on gesture latents with the quantizer kernels.

The reference clusters latents with ``sklearn.cluster.KMeans`` (scikit-learn==1.2.2, a third-party
dependency pinned in its requirements.txt): ``Clustering.py:718-720`` (300 clusters, max_iter=2500),
``train_DAE.py:257-263`` (codebook re-estimate written into ``vq_layer._embedding.weight``) and
``lmdb_data_loader.py:1288-1291`` (``kmeanmodel.predict``).  Lloyd's E-step is the nearest-code search
(g2v_vq_search), its M-step the residual sums + counts of g2v_vq_apply followed by g2v_kmeans_update;
sharded rows only need the one all-reduce of the packed statistics the EMA path already uses.

``KMeans`` mirrors the part of the sklearn estimator the reference touches: constructor
``(n_clusters, max_iter, tol, random_state, init)``, ``fit``, ``predict``, ``fit_predict`` and the
attributes ``cluster_centers_``, ``labels_``, ``inertia_``, ``n_iter_``.  The loop follows sklearn's
``_kmeans_single_lloyd`` (stop when the labels repeat or when the squared centre shift falls below
``tol * mean(var(X))``; labels and inertia come from a final assignment pass with the final centres).  Two stated differences:
a cluster that loses all rows keeps its centre (sklearn relocates it to the farthest rows), and the
default seeding is plain k-means++ drawn from a ``torch.Generator`` (sklearn's greedy variant and its
RNG stream are not reproduced; pass ``init=ndarray`` for a controlled start).  One run (``n_init=1``).
"""
from __future__ import annotations

from typing import Callable, Optional, Union

import numpy as np
import torch

from . import _lib
from .functional import _need_cuda, _on, _ptr, _stream, prepare_codebook, vq_apply, vq_search


def kmeans_update(E_old: torch.Tensor, packed: torch.Tensor, E_new: torch.Tensor, shift2: Optional[torch.Tensor] = None,
                  cb: Optional[torch.Tensor] = None) -> None:
    """Lloyd M-step from a packed statistics buffer (include/g2v_vq.h g2v_kmeans_update)."""
    K, D = E_old.shape
    with _on(E_old.device):
        _lib.check(_lib.load().g2v_kmeans_update(_ptr(E_old), _ptr(packed), K, D, _ptr(E_new), _ptr(shift2), _ptr(cb),
                                                 0 if cb is None else cb.numel(), _stream(E_old.device)),
                   "g2v_kmeans_update")


def kmeans_plusplus(X: torch.Tensor, K: int, generator: torch.Generator) -> torch.Tensor:
    """k-means++ seeding (D^2 sampling), K sequential passes over the rows; torch ops, host-side plumbing."""
    N = X.shape[0]
    first = int(torch.randint(N, (1,), generator=generator, device=X.device).item())
    centres = torch.empty(K, X.shape[1], dtype=torch.float32, device=X.device)
    centres[0] = X[first]
    d2 = (X.float() - centres[0]).square_().sum(1)
    for k in range(1, K):
        tot = d2.sum()
        if float(tot) <= 0.0:                      # fewer distinct rows than clusters: fill with random rows
            pick = int(torch.randint(N, (1,), generator=generator, device=X.device).item())
        else:
            pick = int(torch.multinomial(d2 / tot, 1, generator=generator).item())
        centres[k] = X[pick]
        d2 = torch.minimum(d2, (X.float() - centres[k]).square_().sum(1))
    return centres


class KMeans:
    def __init__(self, n_clusters: int = 8, *, init: Union[str, np.ndarray, torch.Tensor] = "k-means++",
                 max_iter: int = 300, tol: float = 1e-4, random_state: Optional[int] = None,
                 device: Union[str, torch.device, None] = None, stats_reduce: Optional[Callable] = None,
                 count_reduce: Optional[Callable] = None):
        """stats_reduce / count_reduce: data-parallel hooks (rows sharded over ranks): in-place sum
        all-reduce of the packed fp32 statistics, and of small fp64 tensors (changed-label count, column moments)."""
        self.n_clusters, self.init, self.max_iter, self.tol = int(n_clusters), init, int(max_iter), float(tol)
        self.random_state = random_state
        self.device = torch.device(device) if device is not None else None
        self.stats_reduce, self.count_reduce = stats_reduce, count_reduce
        self.cluster_centers_ = None
        self.labels_ = None
        self.inertia_ = None
        self.n_iter_ = 0

    # ---- helpers ----
    def _rows(self, X) -> torch.Tensor:
        if isinstance(X, np.ndarray):
            X = torch.from_numpy(np.ascontiguousarray(X))
        if X.dim() != 2:
            raise ValueError("expected a 2-d array of rows")
        if X.dtype not in (torch.float32, torch.bfloat16, torch.float16):
            X = X.float()
        dev = self.device or (X.device if X.is_cuda else torch.device("cuda", torch.cuda.current_device()))
        X = X.to(dev).contiguous()
        _need_cuda(X, "X")
        return X

    def _init_centres(self, X: torch.Tensor) -> torch.Tensor:
        K = self.n_clusters
        if isinstance(self.init, (np.ndarray, torch.Tensor)):
            c = torch.as_tensor(self.init).to(device=X.device, dtype=torch.float32).contiguous().clone()
            if tuple(c.shape) != (K, X.shape[1]):
                raise ValueError(f"init has shape {tuple(c.shape)}, expected {(K, X.shape[1])}")
            return c
        gen = torch.Generator(device=X.device)
        gen.manual_seed(0 if self.random_state is None else int(self.random_state))
        if self.init == "random":
            return X[torch.randperm(X.shape[0], generator=gen, device=X.device)[:K]].float().contiguous()
        if self.init == "k-means++":
            return kmeans_plusplus(X, K, gen)
        raise ValueError(f"unknown init {self.init!r}")

    def _assign(self, X, E, cb, want_dwr: bool):
        idx = vq_search(X, E, cb)
        xs = X if X.dtype == torch.float32 else X.float()
        _, packed = vq_apply(xs, E, idx, want_out=False, want_stats=True, want_dwr=want_dwr)
        if self.stats_reduce is not None:
            self.stats_reduce(packed)
        return idx, packed

    # ---- estimator surface ----
    def fit(self, X, y=None) -> "KMeans":
        X = self._rows(X)
        N, D = X.shape
        K = self.n_clusters
        if N < K and self.stats_reduce is None:
            raise ValueError(f"n_samples={N} should be >= n_clusters={K}")
        E = self._init_centres(X)
        cb = prepare_codebook(E)
        # tol scaled by the mean per-feature variance, as sklearn's _tolerance()
        colsum = torch.zeros(D, dtype=torch.float64, device=X.device)
        colsq = torch.zeros(D, dtype=torch.float64, device=X.device)
        for r0 in range(0, N, 1 << 20):             # fp64 column moments, a block of rows at a time
            xb = X[r0:r0 + (1 << 20)].double()
            colsum += xb.sum(0)
            colsq += xb.square_().sum(0)
        del xb
        nrows = torch.tensor([float(N)], dtype=torch.float64, device=X.device)
        if self.count_reduce is not None:
            for t in (colsum, colsq, nrows):
                self.count_reduce(t)
        var = colsq / nrows - (colsum / nrows) ** 2
        tol = float(var.mean().item()) * self.tol
        labels_old = None
        shift2 = torch.zeros(1, dtype=torch.float64, device=X.device)
        E_new = torch.empty_like(E)
        self.n_iter_ = 0
        for it in range(self.max_iter):
            idx, packed = self._assign(X, E, cb, want_dwr=True)
            shift2.zero_()
            kmeans_update(E, packed, E_new, shift2, cb)
            E, E_new = E_new, E
            self.n_iter_ = it + 1
            if labels_old is not None:
                changed = (idx != labels_old).sum().double().reshape(1)
                if self.count_reduce is not None:
                    self.count_reduce(changed)
                if float(changed.item()) == 0.0:
                    break
            if float(shift2.item()) <= tol:
                break
            labels_old = idx.clone()
        # labels and inertia belong to the final centres (sklearn re-runs the E-step unless the labels repeated;
        # here the pass is always made: it also yields the inertia)
        idx, packed = self._assign(X, E, cb, want_dwr=False)
        self.cluster_centers_ = E.detach().cpu().numpy()
        self._centres_dev, self._cb = E, cb
        self.labels_ = idx.cpu().numpy().astype(np.int32)
        self.inertia_ = float(packed[K * D + K].item())
        return self

    def predict(self, X) -> np.ndarray:
        if self.cluster_centers_ is None:
            raise RuntimeError("this KMeans instance is not fitted yet")
        X = self._rows(X)
        E = getattr(self, "_centres_dev", None)
        if E is None or E.device != X.device:
            E = torch.from_numpy(np.ascontiguousarray(self.cluster_centers_, dtype=np.float32)).to(X.device)
            self._centres_dev, self._cb = E, prepare_codebook(E)
        return vq_search(X, E, self._cb).cpu().numpy().astype(np.int32)

    def fit_predict(self, X, y=None) -> np.ndarray:
        return self.fit(X).labels_

    @classmethod
    def from_centers(cls, centers: np.ndarray, **kw) -> "KMeans":
        """Wrap centres fitted elsewhere (e.g. an unpickled sklearn model's cluster_centers_) for predict()."""
        km = cls(n_clusters=int(centers.shape[0]), **kw)
        km.cluster_centers_ = np.ascontiguousarray(centers, dtype=np.float32)
        return km
