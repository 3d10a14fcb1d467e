"""k-means on gesture latents with the quantizer kernels.

The reference clusters latents with ``sklearn.cluster.KMeans`` (scikit-learn==1.2.2, a third-party dependency
pinned in its requirements.txt and absent from /root/reference): ``Clustering.py:718-720``
(``KMeans(n_clusters=300, max_iter=2500, random_state=0)``), ``train_DAE.py:257-263`` (codebook re-estimate written
into ``vq_layer._embedding.weight``) and ``lmdb_data_loader.py:1288-1291`` (``kmeanmodel.predict``).  This module
restates that estimator's published algorithm on the quantizer kernels:

  * Lloyd's E-step is the nearest-code search (g2v_vq_search), its M-step the residual sums + counts of
    g2v_vq_apply followed by g2v_vq_step_finalize(G2V_UPDATE_KMEANS); rows sharded over ranks only need the one
    all-reduce of the packed statistics the EMA path already uses;
  * ``fit`` follows ``KMeans.fit`` / ``_kmeans_single_lloyd`` of 1.2.2: the rows are centred on their mean,
    ``n_init`` runs (default 10, that version's default) from greedy k-means++ seeds drawn from ONE
    ``numpy.random.RandomState(random_state)`` stream (``_kmeans_plusplus``: ``2 + int(log K)`` local trials,
    candidates by ``searchsorted`` on the cumulative potential), each run stops when the labels repeat or when
    the squared centre shift falls below ``tol * mean(var(X))``, a cluster that loses all its rows is relocated
    to the row farthest from its centre, labels and inertia come from a final assignment with the final centres,
    and the run with the lowest inertia wins;
  * the random draws are numpy's own (``RandomState`` runs on the host); the distances behind them are computed
    on the device in fp64 and rounded to fp32 as ``euclidean_distances`` does for fp32 rows.  ``seeding``
    selects how the FIRST centre is drawn: ``"sklearn-1.2"`` (``random_state.randint(n)``, the reference's pinned
    version) or ``"sklearn-1.3+"`` (``random_state.choice(n, p=uniform)``; what the scikit-learn installed here
    does, and what tests/test_kmeans_tokenizer.py pins the seeding against).

Stated differences: up to 131 072 rows the sums a k-means++ draw is compared with are numpy's own (host side)
and the seeds reproduce sklearn's row for row; beyond that they are device sums and a candidate can fall on a
neighbouring row (a different, equally distributed draw); which far row goes to which empty cluster follows ``torch.topk`` order, not ``argpartition`` order; a sharded
fit (``stats_reduce``) keeps an empty cluster's centre instead of relocating it.
"""
from __future__ import annotations

from typing import Callable, Optional, Union

import numpy as np
import torch

from . import _lib
from .functional import _need_cuda, _on, _ptr, _stream, packed_numel, prepare_codebook, vq_apply, vq_search


def kmeans_update(E_old: torch.Tensor, packed: torch.Tensor, E_new: torch.Tensor, shift2: Optional[torch.Tensor] = None,
                  cb: Optional[torch.Tensor] = None) -> None:
    """Lloyd M-step from a packed statistics buffer (include/g2v_vq.h g2v_kmeans_update)."""
    K, D = E_old.shape
    with _on(E_old.device):
        _lib.check(_lib.load().g2v_kmeans_update(_ptr(E_old), _ptr(packed), K, D, _ptr(E_new), _ptr(shift2), _ptr(cb),
                                                 0 if cb is None else cb.numel(), _stream(E_old.device)),
                   "g2v_kmeans_update")


def _sq_dists_f32(C: torch.Tensor, X: torch.Tensor, x2: torch.Tensor) -> torch.Tensor:
    """``euclidean_distances(C, X, squared=True)`` for fp32 rows: fp64 arithmetic, clipped at 0, rounded to fp32."""
    d = (C.double() ** 2).sum(1, keepdim=True) + x2.unsqueeze(0) - 2.0 * (C.double() @ X.double().t())
    return d.clamp_min_(0.0).float()


def kmeans_plusplus(X: torch.Tensor, K: int, rs: np.random.RandomState, seeding: str = "sklearn-1.2",
                    n_local_trials: Optional[int] = None, exact_rows: int = 131072):
    """Greedy k-means++ (``sklearn.cluster._kmeans._kmeans_plusplus``): returns (centres [K, D] fp32, row indices).

    Random numbers come from `rs` on the host, distances from torch ops on X's device (fp64, rounded to fp32 like
    ``euclidean_distances`` on fp32 rows).  The reductions the random draws are compared with -- the potential,
    its cumulative sum, the candidates' potentials -- decide WHICH row a draw lands on, so up to `exact_rows` rows
    they are taken on the host with numpy's own fp32 / fp64 summation, exactly as the pinned version does
    (1.2: ``.sum()`` and ``stable_cumsum``; 1.3+: ``@ sample_weight`` and an fp32 ``cumsum``); beyond that they run
    on the device and a draw can land on a neighbouring row."""
    N, D = X.shape
    if n_local_trials is None:
        n_local_trials = 2 + int(np.log(K))
    if seeding not in ("sklearn-1.2", "sklearn-1.3+"):
        raise ValueError(f"unknown seeding {seeding!r}")
    old = seeding == "sklearn-1.2"
    exact = N <= exact_rows
    ones = np.ones(N, dtype=np.float32)
    x2 = (X.double() ** 2).sum(1)
    first = int(rs.randint(N)) if old else int(rs.choice(N, p=np.full(N, 1.0 / N)))
    idx = [first]
    closest = _sq_dists_f32(X[first:first + 1].float(), X, x2)[0]
    if exact:
        cl = closest.cpu().numpy()
        pot = cl.sum() if old else cl @ ones
    else:
        pot = float(closest.double().sum())
    for _ in range(1, K):
        rand_vals = rs.uniform(size=n_local_trials) * pot
        if exact:
            csum = np.cumsum(cl, dtype=np.float64) if old else np.cumsum(ones * cl)
            cand_np = np.searchsorted(csum, rand_vals)
            np.clip(cand_np, None, N - 1, out=cand_np)
            cand = torch.from_numpy(cand_np).to(X.device)
        else:
            cand = torch.searchsorted(torch.cumsum(closest.double(), 0), torch.from_numpy(rand_vals).to(X.device)).clamp_(max=N - 1)
        dc = torch.minimum(closest.unsqueeze(0), _sq_dists_f32(X[cand].float(), X, x2))
        if exact:
            dcn = dc.cpu().numpy()
            pots = dcn.sum(axis=1) if old else (dcn @ ones.reshape(-1, 1)).ravel()
            best = int(np.argmin(pots))
            pot = pots[best]
            cl = dcn[best]
        else:
            pots = dc.double().sum(1)
            best = int(torch.argmin(pots))
            pot = float(pots[best])
        closest = dc[best]
        idx.append(int(cand[best]))
    ids = torch.tensor(idx, device=X.device)
    return X[ids].float().contiguous(), np.asarray(idx)


class KMeans:
    def __init__(self, n_clusters: int = 8, *, init: Union[str, np.ndarray, torch.Tensor] = "k-means++",
                 n_init: Union[int, str] = 10, max_iter: int = 300, tol: float = 1e-4,
                 random_state: Optional[int] = None, device: Union[str, torch.device, None] = None,
                 stats_reduce: Optional[Callable] = None, count_reduce: Optional[Callable] = None,
                 broadcast: Optional[Callable] = None, seeding: str = "sklearn-1.2", relocate_empty: bool = True):
        """stats_reduce / count_reduce / broadcast: data-parallel hooks (rows sharded over ranks): in-place sum
        all-reduce of the packed fp32 statistics, of small fp64 tensors (changed-label count, column moments),
        and broadcast of rank 0's initial centres (default with `stats_reduce`: torch.distributed.broadcast) --
        every rank must start from the SAME centres, whatever its shard would have drawn.
        relocate_empty=False: no per-iteration host read at all (the convergence flags are read one iteration
        late, the superfluous M-step is discarded)."""
        self.n_clusters, self.init, self.max_iter, self.tol = int(n_clusters), init, int(max_iter), float(tol)
        self.n_init, self.random_state, self.seeding = n_init, random_state, seeding
        self.device = torch.device(device) if device is not None else None
        self.stats_reduce, self.count_reduce, self.broadcast = stats_reduce, count_reduce, broadcast
        self.relocate_empty = bool(relocate_empty)
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

    def _sharded(self) -> bool:
        return self.stats_reduce is not None

    def _bcast(self, t: torch.Tensor) -> None:
        if not self._sharded():
            return
        if self.broadcast is not None:
            self.broadcast(t)
            return
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            dist.broadcast(t, src=0)
        else:
            raise RuntimeError("a sharded KMeans needs `broadcast` (or an initialised torch.distributed) so that "
                               "every rank starts from the same centres")

    def _init_centres(self, X: torch.Tensor, rs: np.random.RandomState) -> torch.Tensor:
        K = self.n_clusters
        if isinstance(self.init, (np.ndarray, torch.Tensor)):
            c = torch.as_tensor(self.init).to(device=X.device, dtype=torch.float32).contiguous().clone()
            if tuple(c.shape) != (K, X.shape[1]):
                raise ValueError(f"init has shape {tuple(c.shape)}, expected {(K, X.shape[1])}")
        elif self.init == "random":
            ids = torch.from_numpy(rs.permutation(X.shape[0])[:K].copy()).to(X.device)
            c = X[ids].float().contiguous()
        elif self.init == "k-means++":
            c, _ = kmeans_plusplus(X, K, rs, self.seeding)
        else:
            raise ValueError(f"unknown init {self.init!r}")
        self._bcast(c)            # sharded rows: rank 0's draw is everybody's start
        return c

    def _assign(self, X, E, cb, want_dwr: bool):
        idx = vq_search(X, E, cb)
        xs = X if X.dtype == torch.float32 else X.float()
        _, packed = vq_apply(xs, E, idx, want_out=False, want_stats=True, want_dwr=want_dwr)
        if self.stats_reduce is not None:
            self.stats_reduce(packed)
        return idx, packed

    @staticmethod
    def _relocate(X: torch.Tensor, E: torch.Tensor, idx: torch.Tensor, packed: torch.Tensor, n_empty: int) -> None:
        """sklearn's _relocate_empty_clusters_dense on the packed statistics: the n_empty rows farthest from their
        centres each leave their cluster and become the single member of an empty one."""
        K, D = E.shape
        counts = packed[K * D:K * D + K]
        dwr = packed[:K * D].view(K, D)
        dist = torch.empty(X.shape[0], dtype=torch.float32, device=X.device)
        for r0 in range(0, X.shape[0], 1 << 18):
            xb = X[r0:r0 + (1 << 18)].float()
            dist[r0:r0 + xb.shape[0]] = ((xb - E[idx[r0:r0 + xb.shape[0]].long()]) ** 2).sum(1)
        far = torch.topk(dist, n_empty).indices
        empty = torch.nonzero(counts == 0).flatten()[:n_empty]
        for f, e in zip(far.tolist(), empty.tolist()):
            o = int(idx[f])
            xf = X[f].float()
            dwr[o] -= xf - E[o]
            counts[o] -= 1
            dwr[e] = xf - E[e]
            counts[e] = 1

    def _lloyd(self, X: torch.Tensor, E: torch.Tensor, tol: float):
        """One run of sklearn's _kmeans_single_lloyd from centres E: (labels, inertia, centres, n_iter)."""
        N, D = X.shape
        K = self.n_clusters
        dev = X.device
        cb = prepare_codebook(E)
        bufs = [E, torch.empty_like(E), torch.empty_like(E)]
        cur = 0
        flags_dev = [torch.zeros(3, dtype=torch.float64, device=dev) for _ in range(2)]
        flags_host = [torch.zeros(3, dtype=torch.float64).pin_memory() for _ in range(2)]
        events = [torch.cuda.Event(), torch.cuda.Event()]
        labels = [None, None]
        n_iter, result = 0, None
        sync_each = self.relocate_empty and not self._sharded()

        def stop_reason(fl) -> bool:
            changed, shift2 = float(fl[1]), float(fl[0])
            return changed == 0.0 or shift2 <= tol

        for it in range(self.max_iter):
            Ec, En = bufs[cur], bufs[(cur + 1) % 3]
            idx, packed = self._assign(X, Ec, cb, want_dwr=True)
            labels[it & 1] = idx
            fd = flags_dev[it & 1]
            fd.zero_()
            fd[1] = float("inf") if labels[(it + 1) & 1] is None else (idx != labels[(it + 1) & 1]).sum().double()
            if self.count_reduce is not None and labels[(it + 1) & 1] is not None:
                ch = fd[1:2].clone()
                self.count_reduce(ch)
                fd[1] = ch[0]
            fd[2] = (packed[K * D:K * D + K] == 0).sum().double()
            if sync_each:
                n_empty = int(fd[2].item())                 # the one host read of this iteration (relocation decision)
                if n_empty > 0:
                    self._relocate(X, Ec, idx, packed, min(n_empty, N))
            kmeans_update(Ec, packed, En, fd[0:1], cb)       # centres -> En, ||shift||^2 -> fd[0], cb re-prepared for En
            flags_host[it & 1].copy_(fd, non_blocking=True)
            events[it & 1].record()
            n_iter = it + 1
            if sync_each:
                events[it & 1].synchronize()
                if stop_reason(flags_host[it & 1]):
                    result = (En, it + 1, float(flags_host[it & 1][1]) == 0.0, idx)
                    break
            elif it >= 1:
                # flags of iteration it-1, read while iteration `it` runs; if it-1 converged, iteration `it`'s
                # E-step is exactly the final assignment sklearn makes, and its M-step output is discarded
                events[(it + 1) & 1].synchronize()
                if stop_reason(flags_host[(it + 1) & 1]):
                    result = (Ec, it, True, idx)             # Ec = the centres iteration it-1 produced
                    n_iter = it
                    break
            cur = (cur + 1) % 3
        if result is None:                                   # max_iter reached
            torch.cuda.current_stream(dev).synchronize()
            result = (bufs[cur], n_iter, False, None)
        E_fin, n_iter, have_labels, idx_fin = result
        # labels and inertia belong to the final centres: sklearn re-runs the E-step unless the labels repeated;
        # here the pass is always made (it also yields the inertia), except in the lagged mode where it already ran
        if not (have_labels and not sync_each):
            cb = prepare_codebook(E_fin, cb)
            idx_fin, packed = self._assign(X, E_fin, cb, want_dwr=False)
        else:
            cb = prepare_codebook(E_fin, cb)
        inertia = float(packed[K * D + K].item())
        return idx_fin, inertia, E_fin.clone(), n_iter, cb

    # ---- estimator surface ----
    def fit(self, X, y=None) -> "KMeans":
        X = self._rows(X)
        N, D = X.shape
        K = self.n_clusters
        if N < K and not self._sharded():
            raise ValueError(f"n_samples={N} should be >= n_clusters={K}")
        with _on(X.device):
            # column moments in fp64: the mean the rows are centred on, and sklearn's _tolerance()
            colsum = torch.zeros(D, dtype=torch.float64, device=X.device)
            colsq = torch.zeros(D, dtype=torch.float64, device=X.device)
            for r0 in range(0, N, 1 << 20):
                xb = X[r0:r0 + (1 << 20)].double()
                colsum += xb.sum(0)
                colsq += xb.square_().sum(0)
            nrows = torch.tensor([float(N)], dtype=torch.float64, device=X.device)
            if self.count_reduce is not None:
                for t in (colsum, colsq, nrows):
                    self.count_reduce(t)
            mean = colsum / nrows
            tol = float((colsq / nrows - mean ** 2).mean().item()) * self.tol
            mean32 = mean.float()
            Xc = (X.float() - mean32).contiguous()           # X -= X_mean (KMeans.fit); centres get it back below
            explicit = isinstance(self.init, (np.ndarray, torch.Tensor))
            n_init = 1 if explicit else (10 if self.n_init == "auto" else int(self.n_init))
            rs = np.random.RandomState(self.random_state)
            if explicit:
                self.init = torch.as_tensor(self.init).to(device=X.device, dtype=torch.float32) - mean32
            best = None
            for _ in range(max(1, n_init)):
                E0 = self._init_centres(Xc, rs)
                run = self._lloyd(Xc, E0, tol)
                if best is None or run[1] < best[1]:
                    best = run
            if explicit:
                self.init = self.init + mean32
            idx, inertia, E, n_iter, cb = best
            E = (E + mean32).contiguous()
            self.cluster_centers_ = E.detach().cpu().numpy()
            self._centres_dev, self._cb = E, prepare_codebook(E)
            self.labels_ = idx.cpu().numpy().astype(np.int32)
            self.inertia_ = inertia
            self.n_iter_ = n_iter
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
