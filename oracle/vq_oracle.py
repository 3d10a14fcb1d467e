"""CPU oracle for the Gesture2Vec vector-quantizer hot path.  TEST INFRASTRUCTURE ONLY.

This file is a numpy restatement of the reference quantizers, written from
their source (not copied), so the GPU tests can run on a box where
``/root/reference`` does not exist.  Only ``tests/``, ``__graft_entry__.smoke()``
and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import
it.  The product package ``gesture2vec_b200`` never does.

Parity status: PINNED.  The reference ships no golden vectors or tests
(SURVEY.md §4), so the oracle is pinned against outputs of the reference
modules themselves, imported from ``/root/reference`` in the authoring
container by ``tests/golden/make_golden.py`` and committed under
``tests/golden/*.npz``.  ``tests/test_oracle_golden.py`` replays them.

Reference statements followed (all paths relative to /root/reference/scripts/model):
  hard VQ      DAE_model.py:301-348            Autoencoder_VQVAE_model.py:1115-1174
  EMA VQ       DAE_model.py:396-482            Autoencoder_VQVAE_model.py:1217-1296
  hstack EMA   Autoencoder_VQVAE_model.py:1745-1812
All reference arithmetic is fp32; indices are int64.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32

# Near-tie tolerance used by every parity test: a disagreement between two
# argmin implementations is a "near-tie" iff the fp64 distance gap between the
# two chosen codes is <= EPS_TIE * (|z|^2 + |e|^2).  The reference evaluates
# (sum z^2 + sum e^2) - 2 z.e in fp32, whose rounding noise is a few ulp of
# |z|^2+|e|^2 (2^-23 relative each); 2^-18 leaves ~10x margin over the largest
# fp32-vs-fp64 deviation measured on the synthetic distributions.
EPS_TIE = 2.0 ** -18


# --------------------------------------------------------------------------
# a1: flatten                     DAE_model.py:317,418 / VQVAE_model.py:1127,1229
# --------------------------------------------------------------------------
def flatten_rows(x: np.ndarray, D: int) -> np.ndarray:
    """``inputs.view(-1, D)`` on a contiguous array (pure reinterpretation)."""
    x = np.ascontiguousarray(x, dtype=F32)
    if x.size % D:
        raise ValueError(f"numel {x.size} not divisible by embedding_dim {D}")
    return x.reshape(-1, D)


# --------------------------------------------------------------------------
# a2: distances                   DAE_model.py:320-324,423-427 / VQVAE_model.py:1132-1136
# --------------------------------------------------------------------------
def distances_f32(z: np.ndarray, E: np.ndarray) -> np.ndarray:
    """(sum z^2)[:,None] + (sum E^2)[None,:] - 2 z E^T, in fp32, in that order."""
    z = z.astype(F32, copy=False)
    E = E.astype(F32, copy=False)
    z2 = np.sum(z * z, axis=1, keepdims=True, dtype=F32)
    e2 = np.sum(E * E, axis=1, dtype=F32)
    return (z2 + e2) - F32(2.0) * (z @ E.T)


def distances_f64(z: np.ndarray, E: np.ndarray) -> np.ndarray:
    """Exact-arithmetic stand-in: the same quantity in fp64."""
    z = z.astype(np.float64)
    E = E.astype(np.float64)
    return (np.sum(z * z, 1, keepdims=True) + np.sum(E * E, 1)) - 2.0 * (z @ E.T)


# --------------------------------------------------------------------------
# a3: argmin (first minimal index, like torch.argmin)   DAE_model.py:327,434
# --------------------------------------------------------------------------
def argmin_first(d: np.ndarray) -> np.ndarray:
    return np.argmin(d, axis=1).astype(np.int64)  # numpy also returns the first minimum


def nearest_code_f32(z, E):
    return argmin_first(distances_f32(z, E))


def nearest_code_f64(z, E, block: int = 8192):
    out = np.empty(z.shape[0], np.int64)
    for s in range(0, z.shape[0], block):
        out[s:s + block] = argmin_first(distances_f64(z[s:s + block], E))
    return out


# --------------------------------------------------------------------------
# a4: one-hot                      DAE_model.py:328-331
# --------------------------------------------------------------------------
def one_hot(idx: np.ndarray, K: int) -> np.ndarray:
    enc = np.zeros((idx.shape[0], K), F32)
    enc[np.arange(idx.shape[0]), idx] = 1.0
    return enc


# --------------------------------------------------------------------------
# a8: perplexity                   DAE_model.py:346-347
# --------------------------------------------------------------------------
def perplexity_from_counts(counts: np.ndarray, N: int) -> np.float32:
    p = (counts.astype(F32) / F32(N)).astype(F32)
    return F32(np.exp(-np.sum(p * np.log(p + F32(1e-10)), dtype=F32)))


# --------------------------------------------------------------------------
# a9: EMA codebook update          DAE_model.py:451-471 / VQVAE_model.py:1262-1282
# --------------------------------------------------------------------------
def ema_update(cluster_size, ema_w, counts, dw, decay, eps):
    """Returns (cluster_size', ema_w', E') exactly as the reference orders the ops."""
    K = cluster_size.shape[0]
    g = F32(decay)
    one_m = F32(1.0 - decay)  # python double 1-decay, then cast (torch scalar semantics)
    cs = cluster_size.astype(F32) * g + one_m * counts.astype(F32)
    n = np.sum(cs, dtype=F32)
    cs = ((cs + F32(eps)) / (n + F32(K * eps)) * n).astype(F32)
    w = (ema_w.astype(F32) * g + one_m * dw.astype(F32)).astype(F32)
    E = (w / cs[:, None]).astype(F32)
    return cs, w, E


# --------------------------------------------------------------------------
# forward of the two hard quantizers (Appendix A of SURVEY.md)
# --------------------------------------------------------------------------
def vq_forward(x, E, beta, *, ema=False, search=None, idx=None):
    """Hard VQ forward.

    x      : inputs, any shape, numel % D == 0
    E      : [K, D] codebook used for the search *and* the gather (the old one in EMA mode)
    beta   : commitment cost
    ema    : False -> VQ_Payam loss (1+beta)*mse ; True -> VQ_Payam_EMA loss beta*mse
    search : optional [N, D] rows to run the search on instead of x.view(-1,D)
             (Autoencoder_VQVAE_model.VQ_Payam_EMA searches on pre_linear(z), :1230)
    idx    : optional precomputed indices (skips the search)
    Returns dict(idx, enc, q, out, loss, perplexity, counts, dw).
    """
    E = np.ascontiguousarray(E, F32)
    K, D = E.shape
    z = flatten_rows(x, D)
    zs = z if search is None else np.ascontiguousarray(search, F32)
    N = z.shape[0]
    if idx is None:
        idx = nearest_code_f32(zs, E)
    enc = one_hot(idx, K)
    q = E[idx].reshape(np.shape(x))                       # a5: encodings @ E == E[idx]
    xx = np.asarray(x, F32)
    diff = (q - xx).astype(F32)
    mse = F32(np.mean(diff.astype(np.float64) ** 2))       # a6 (mean over all N*D elements)
    loss = F32(beta) * mse if ema else mse + F32(beta) * mse
    out = (xx + (q - xx)).astype(F32)                      # a7 forward value of the STE
    counts = np.bincount(idx, minlength=K).astype(np.int64)
    dw = np.zeros((K, D), np.float64)
    np.add.at(dw, idx, zs.astype(np.float64))              # encodings^T @ flat_input
    return dict(idx=idx, enc=enc, q=q, out=out, loss=F32(loss),
                perplexity=perplexity_from_counts(counts, N), counts=counts,
                dw=dw.astype(F32), N=N, D=D, K=K)


# --------------------------------------------------------------------------
# a10: backward closed forms (autograd of the reference, verified in make_golden.py)
# --------------------------------------------------------------------------
def vq_backward(x, E, idx, beta, g_loss, g_out, *, ema=False):
    """Returns (grad_x, grad_E or None).

    d loss/dx = 2*beta*(x - q)/M ; STE passes g_out through unchanged.
    VQ_Payam only: d loss/dE[k] = sum_{n: idx=k} 2*(q_n - z_n)/M.
    """
    E = np.ascontiguousarray(E, F32)
    K, D = E.shape
    z = flatten_rows(x, D).astype(np.float64)
    M = z.size
    q = E[idx].astype(np.float64)
    gx = 2.0 * beta * g_loss * (z - q) / M
    if g_out is not None:
        gx = gx + flatten_rows(g_out, D).astype(np.float64)
    gE = None
    if not ema:
        gE = np.zeros((K, D), np.float64)
        np.add.at(gE, idx, 2.0 * g_loss * (q - z) / M)
        gE = gE.astype(F32)
    return gx.astype(F32).reshape(np.shape(x)), gE


# --------------------------------------------------------------------------
# Stateful mirrors used by the tests (one object per reference class flavour)
# --------------------------------------------------------------------------
class HardVQ:
    """VQ_Payam (DAE_model.py:277 / Autoencoder_VQVAE_model.py:1088)."""

    def __init__(self, E, beta):
        self.E = np.array(E, F32)
        self.beta = float(beta)

    def forward(self, x):
        return vq_forward(x, self.E, self.beta, ema=False)


class EmaVQ:
    """VQ_Payam_EMA.  flavour='dae' (DAE_model.py:351) searches on the raw rows;
    flavour='vqvae' (Autoencoder_VQVAE_model.py:1182) searches on pre_linear(rows)."""

    def __init__(self, E, ema_w, beta, decay, eps=1e-5, *, flavour="dae", W=None, b=None,
                 cluster_size=None):
        self.E = np.array(E, F32)
        self.ema_w = np.array(ema_w, F32)
        K = self.E.shape[0]
        self.cluster_size = np.zeros(K, F32) if cluster_size is None else np.array(cluster_size, F32)
        self.beta, self.decay, self.eps = float(beta), float(decay), float(eps)
        self.flavour = flavour
        self.W = None if W is None else np.array(W, F32)
        self.b = None if b is None else np.array(b, F32)
        self.training = True

    def forward(self, x):
        D = self.E.shape[1]
        search = None
        if self.flavour == "vqvae":
            z = flatten_rows(x, D)
            search = (z @ self.W.T + self.b).astype(F32)
        r = vq_forward(x, self.E, self.beta, ema=True, search=search)
        if self.training:
            self.cluster_size, self.ema_w, self.E = ema_update(
                self.cluster_size, self.ema_w, r["counts"], r["dw"], self.decay, self.eps)
        return r


# --------------------------------------------------------------------------
# near-tie audit shared by CPU and GPU parity tests
# --------------------------------------------------------------------------
def audit_indices(z, E, idx_test, idx_ref, eps_tie=EPS_TIE):
    """Compare two index vectors.  Returns dict(mismatch, near_tie, hard, worst_rel_gap).

    A mismatch row is a *near-tie* iff |d64[ref] - d64[test]| <= eps_tie*(|z|^2+|e_ref|^2);
    anything else is a *hard* mismatch, which parity tests require to be zero.
    """
    idx_test = np.asarray(idx_test).astype(np.int64).ravel()
    idx_ref = np.asarray(idx_ref).astype(np.int64).ravel()
    bad = np.nonzero(idx_test != idx_ref)[0]
    res = dict(mismatch=int(bad.size), near_tie=0, hard=0, worst_rel_gap=0.0, rows=bad)
    if bad.size == 0:
        return res
    zz = z[bad].astype(np.float64)
    ea = E[idx_test[bad]].astype(np.float64)
    eb = E[idx_ref[bad]].astype(np.float64)
    da = np.sum((zz - ea) ** 2, 1)
    db = np.sum((zz - eb) ** 2, 1)
    scale = np.sum(zz * zz, 1) + np.sum(eb * eb, 1)
    rel = np.abs(da - db) / np.maximum(scale, 1e-300)
    res["near_tie"] = int(np.sum(rel <= eps_tie))
    res["hard"] = int(np.sum(rel > eps_tie))
    res["worst_rel_gap"] = float(rel.max())
    return res


# --------------------------------------------------------------------------
# synthetic inputs of SURVEY.md §8(d) (value distributions & seeds)
# --------------------------------------------------------------------------
def synth_latents(kind: str, N: int, D: int, E: np.ndarray | None = None, seed: int | None = None):
    """kind in {'iid','gru','clustered'}: seeds 1234 / 1235 / 1236 unless given."""
    if kind == "iid":
        rng = np.random.default_rng(1234 if seed is None else seed)
        return rng.standard_normal((N, D), dtype=F32)
    if kind == "gru":
        rng = np.random.default_rng(1235 if seed is None else seed)
        return np.tanh(F32(0.8) * rng.standard_normal((N, D), dtype=F32)).astype(F32)
    if kind == "clustered":
        rng = np.random.default_rng(1236 if seed is None else seed)
        K = E.shape[0]
        w = 1.0 / np.arange(1, K + 1) ** 1.1
        c = rng.choice(K, size=N, p=w / w.sum())
        return (E[c] + F32(0.1) * rng.standard_normal((N, D), dtype=F32)).astype(F32)
    raise ValueError(kind)


def synth_codebook(kind: str, K: int, D: int, seed: int = 0):
    """kind in {'normal','uniform1','uniform_invK'}: the three class inits (§8 a11)."""
    rng = np.random.default_rng(seed)
    if kind == "normal":
        return rng.standard_normal((K, D), dtype=F32)
    if kind == "uniform1":
        return rng.uniform(-1.0, 1.0, (K, D)).astype(F32)
    if kind == "uniform_invK":
        return rng.uniform(-1.0 / K, 1.0 / K, (K, D)).astype(F32)
    raise ValueError(kind)
