"""Drop-in replacements for the reference's hard quantizer modules.

Same class names, constructor arguments, attribute names (= state_dict keys), forward
contract ``forward(inputs) -> (loss, quantized, perplexity, encodings)`` and train/eval
semantics as

  flavour "vqvae":  scripts/model/Autoencoder_VQVAE_model.py  VQ_Payam :1088, VQ_Payam_EMA :1182,
                    VectorQuantizerEMA :1713
  flavour "dae":    scripts/model/DAE_model.py                VQ_Payam :277,  VQ_Payam_EMA :351

but the arithmetic runs in the sm_100a CUDA library behind include/g2v_vq.h.  Inputs must be
CUDA tensors (CPU tensors are moved to the module's device and the results moved back, which
keeps the DataLoader-worker call site lmdb_data_loader.py:1280 working on a GPU box); there is
no CPU implementation here.

Differences from the reference that are not observable through the interface:
  * `_ema_w` / `_embedding.weight` keep their Parameter identity; their `.data` is re-pointed
    to a fresh tensor each EMA step (the reference re-creates the Parameters, :1276-1282);
  * `encodings` is built by a scatter kernel from the int32 indices (also kept on the module
    as `last_indices`), not by scatter_ into torch.zeros.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.nn as nn

from . import _lib
from . import functional as F


class _HardQuantizerBase(nn.Module):
    #: set by subclasses
    _ema: bool = False
    _projects: bool = False      # the search runs on a projection of the rows (`_search_rows`)

    def _init_common(self, num_embeddings: int, embedding_dim: int, commitment_cost: float):
        self._embedding_dim = int(embedding_dim)
        self._num_embeddings = int(num_embeddings)
        self._commitment_cost = float(commitment_cost)
        # non-persistent runtime state
        self._cb: Optional[torch.Tensor] = None          # prepared codebook aux buffer
        self._cb_key = None                              # (data_ptr, _version, device) it was built for
        self.search_flags: int = _lib.ALGO_AUTO
        self.return_encodings: bool = True               # dense one-hot like the reference
        self.last_indices: Optional[torch.Tensor] = None
        # data-parallel EMA: callable(packed fp32 tensor) doing an in-place sum all-reduce
        self.stats_reduce: Optional[Callable[[torch.Tensor], None]] = None
        self.grad_scale: float = 1.0
        # EMA state updated in place (fixed addresses: a captured CUDA graph of the training step advances the
        # state on every replay).  The backward pass then reads the old codebook from a private copy, so at most
        # one forward may be pending its backward.  Default: fresh tensors per step, like the reference.
        # bit-reproducible statistics (EMA sums, codebook gradient, loss): sorted fixed-order fp64 sums instead of
        # fp32 atomics -- the same bits from run to run, at the cost of a sort and a second pass over the rows
        self.deterministic: bool = False
        self.ema_inplace: bool = False
        self._E_prev: Optional[torch.Tensor] = None

    # -- reference API -------------------------------------------------------------------------
    def embedding_grad(self, what: bool) -> None:
        """Autoencoder_VQVAE_model.py:1110-1113."""
        for param in self._embedding.parameters():
            param.requires_grad = what

    # -- helpers -------------------------------------------------------------------------------
    def _codebook(self, dev: torch.device) -> Tuple[torch.Tensor, torch.Tensor]:
        W = self._embedding.weight
        if W.device != dev:
            raise RuntimeError(f"quantizer codebook is on {W.device}, inputs on {dev}")
        E = W.detach()
        if E.dtype != torch.float32 or not E.is_contiguous():
            raise RuntimeError("codebook must be a contiguous fp32 tensor")
        key = (E.data_ptr(), W._version, str(dev))
        if self._cb is None or self._cb_key != key:
            self._cb = F.prepare_codebook(E, self._cb)
            self._cb_key = key
        return W, self._cb

    def invalidate_codebook_cache(self) -> None:
        """Call after writing the codebook in place through `.data` (`W.data.copy_()`, `dist.broadcast(W.data)`,
        `.data.normal_()` ...): such writes do not bump `W._version`, which is what keys the cached fp16 /
        ||e||^2 aux buffer.  Assigning a new tensor (`W.data = t`, as train_DAE.py:261 does), optimizer steps,
        `load_state_dict` and `copy_` on the Parameter itself are all detected without this call."""
        self._cb_key = None

    def _search_rows(self, flat: torch.Tensor) -> Optional[torch.Tensor]:
        """Rows the search (and the EMA sums) run on, if different from the raw rows."""
        return None

    def _fold(self, dev: torch.device):
        """(E_fold, cb_fold, W_proj, b_proj) when the flavour's projection is folded into the codebook, else None."""
        return None

    def _run(self, inputs: torch.Tensor, indices: Optional[torch.Tensor] = None):
        if inputs.numel() % self._embedding_dim:
            raise RuntimeError(
                f"shape '[-1, {self._embedding_dim}]' is invalid for input of size {inputs.numel()}")
        src_dev = inputs.device
        W0 = self._embedding.weight
        if not inputs.is_cuda:
            if not W0.is_cuda:
                raise RuntimeError(
                    "gesture2vec_b200 quantizers run on a CUDA device only: move the module to a "
                    "B200 (`.cuda()`); there is no CPU implementation of this path")
            inputs = inputs.to(W0.device)
        if inputs.dtype != torch.float32:
            inputs = inputs.float()
        flat = inputs.contiguous().view(-1, self._embedding_dim)          # a1: inputs.view(-1, D)
        with F._on(flat.device):
            W, cb = self._codebook(flat.device)
            fold = self._fold(flat.device)
            zs, dw_transform = None, None
            with torch.no_grad():
                if fold is None:
                    zs = self._search_rows(flat.detach())
                else:
                    # the projection lives in the codebook: search the RAW rows (the TMA zero-fills the extra columns)
                    E_fold, cb_fold, Wp, bp = fold
                    dw_transform = (Wp, bp)
                    if indices is None:
                        indices = F.vq_search_wide(flat.detach(), E_fold, cb_fold, flags=self.search_flags)
            training_ema = self._ema and self.training
            want_dwr = training_ema or (not self._ema and W.requires_grad and torch.is_grad_enabled())
            reduce_fn = self.stats_reduce if training_ema else None
            E_arg = W if (not self._ema) else W.detach()
            ema = None
            if training_ema:
                if not (self._ema_w.data.is_contiguous() and self._ema_cluster_size.is_contiguous()):
                    raise RuntimeError("EMA state must be contiguous")
                E_prev = None
                if self.ema_inplace:
                    if self._E_prev is None or self._E_prev.shape != W.shape or self._E_prev.device != W.device:
                        self._E_prev = torch.empty_like(W.data)
                    E_prev = self._E_prev
                ema = F.EmaState(self._ema_cluster_size, self._ema_w.data, self._decay, self._epsilon, E_prev)
            loss, out, ppl, idx, packed = F.quantize(
                flat, E_arg, zs=zs, cb=cb, beta=self._commitment_cost,
                coef_codebook=0.0 if self._ema else 1.0, want_dwr=want_dwr, reduce_fn=reduce_fn,
                grad_scale=self.grad_scale, flags=self.search_flags, ema=ema, idx=indices,
                deterministic=self.deterministic, dw_transform=dw_transform if training_ema else None)
            self.last_indices = idx
            quantized = out.view(inputs.shape)
            enc = F.one_hot(idx, self._num_embeddings) if self.return_encodings else idx.long().unsqueeze(1)
            if ema is not None:
                if ema.pending is not None:          # statistics exchange + finalise ran on the side stream
                    torch.cuda.current_stream(flat.device).wait_event(ema.pending)
                # the step's single finalise launch wrote the new state into fresh tensors (the reference
                # re-creates its Parameters each step, :1276-1282) and re-prepared `cb` for the new codebook
                if not self.ema_inplace:
                    with torch.no_grad():
                        self._ema_cluster_size = ema.cs_out   # registered buffer: assignment keeps it registered
                        self._ema_w.data = ema.w_out
                        W.data = ema.E_new
                self._cb_key = (ema.E_new.data_ptr(), W._version, str(ema.E_new.device))
        if src_dev != quantized.device:
            loss, quantized, ppl, enc = (t.to(src_dev) for t in (loss, quantized, ppl, enc))
        return loss, quantized, ppl, enc

    def tokenize(self, inputs: torch.Tensor) -> torch.Tensor:
        """Bulk path: int32 code ids for the rows of `inputs` -- the same ids as `argmax(forward(x)[3], 1)` --
        with no one-hot / gather / loss.  16-bit rows are searched as stored unless the flavour projects them
        first (`_search_rows`), in which case they go through the projection in fp32 like the forward pass."""
        flat = inputs.contiguous().view(-1, self._embedding_dim)
        with F._on(flat.device):
            W, cb = self._codebook(flat.device)
            fold = self._fold(flat.device)
            with torch.no_grad():
                if fold is not None:
                    E_fold, cb_fold, _, _ = fold
                    return F.vq_search_wide(flat, E_fold, cb_fold, flags=self.search_flags)
                zs = self._search_rows(flat.float() if self._projects else flat)
            return F.vq_search(flat if zs is None else zs, W.detach(), cb, flags=self.search_flags)

    def forward(self, inputs: torch.Tensor):
        return self._run(inputs)

    def forward_with_indices(self, inputs: torch.Tensor, indices: torch.Tensor):
        """`forward` with the code ids given (one per row of `inputs.view(-1, D)`) instead of searched: rows
        tokenised earlier by `tokenize`, or -- in the parity tests -- the reference's own indices."""
        return self._run(inputs, indices=indices.reshape(-1))


# ================================================================================================
# flavour "dae": scripts/model/DAE_model.py
# ================================================================================================
class DAE_VQ_Payam(_HardQuantizerBase):
    """DAE_model.py:277-348.  state_dict: `_embedding.weight`."""

    def __init__(self, num_embeddings: int, embedding_dim: int, commitment_cost: float):
        super().__init__()
        self._init_common(num_embeddings, embedding_dim, commitment_cost)
        self._embedding = nn.Embedding(self._num_embeddings, self._embedding_dim)
        self._embedding.weight.data.uniform_(-1 / self._num_embeddings, 1 / self._num_embeddings)


class DAE_VQ_Payam_EMA(_HardQuantizerBase):
    """DAE_model.py:351-482.  state_dict: `pre_linear.*`, `_embedding.weight`, `_ema_w`,
    `_ema_cluster_size`.  pre_linear exists but is unused (:419)."""
    _ema = True

    def __init__(self, num_embeddings: int, embedding_dim: int, commitment_cost: float, decay: float,
                 epsilon: float = 1e-5):
        super().__init__()
        self._init_common(num_embeddings, embedding_dim, commitment_cost)
        self.pre_linear = nn.Linear(self._embedding_dim, self._embedding_dim)
        self._embedding = nn.Embedding(self._num_embeddings, self._embedding_dim)
        self._embedding.weight.data.uniform_(-1 / self._num_embeddings, 1 / self._num_embeddings)
        self.register_buffer("_ema_cluster_size", torch.zeros(num_embeddings))
        self._ema_w = nn.Parameter(torch.Tensor(num_embeddings, self._embedding_dim))
        self._ema_w.data.normal_()
        self._decay = decay
        self._epsilon = epsilon


# ================================================================================================
# flavour "vqvae": scripts/model/Autoencoder_VQVAE_model.py
# ================================================================================================
class VQVAE_VQ_Payam(_HardQuantizerBase):
    """Autoencoder_VQVAE_model.py:1088-1174.  Owns an unused pre_linear (:1099); E ~ N(0,1) (:1104)."""

    def __init__(self, num_embeddings: int, embedding_dim: int, commitment_cost: float):
        super().__init__()
        self._init_common(num_embeddings, embedding_dim, commitment_cost)
        self.pre_linear = nn.Linear(self._embedding_dim, self._embedding_dim)
        self._embedding = nn.Embedding(self._num_embeddings, self._embedding_dim)
        self._embedding.weight.data.normal_()


class VQVAE_VQ_Payam_EMA(_HardQuantizerBase):
    """Autoencoder_VQVAE_model.py:1182-1296.  The search and the EMA sums run on pre_linear(z)
    (:1230, :1275); the loss and the straight-through value use the raw inputs (:1285, :1292).
    E ~ U(-1,1) (:1204)."""
    _ema = True

    def __init__(self, num_embeddings: int, embedding_dim: int, commitment_cost: float, decay: float,
                 epsilon: float = 1e-5):
        super().__init__()
        self._init_common(num_embeddings, embedding_dim, commitment_cost)
        self.pre_linear = nn.Linear(self._embedding_dim, self._embedding_dim)
        self._embedding = nn.Embedding(self._num_embeddings, self._embedding_dim)
        self._embedding.weight.data.uniform_(-1, 1)
        self.register_buffer("_ema_cluster_size", torch.zeros(num_embeddings))
        self._ema_w = nn.Parameter(torch.Tensor(num_embeddings, self._embedding_dim))
        self._ema_w.data.normal_()
        self._decay = decay
        self._epsilon = epsilon

    _projects = True
    #: fold pre_linear into the codebook (see _fold); False: project every row with g2v_gemm_f32 first
    fold_projection: bool = True
    _FOLD_PAD = 4           # extra codebook columns: the offset term + zeros up to a 16-byte row pitch

    def _fold(self, dev: torch.device):
        """pre_linear is linear, so  argmin_k |W z + b - e_k|^2 = argmin_k ( c_k - 2 z . (W^T e_k) ),
        c_k = |e_k|^2 - 2 b.e_k  (|W z + b|^2 is constant over k).  With e'_k = W^T e_k and one extra coordinate
        t_k = sqrt(c_k + C - |e'_k|^2)  (C >= 0 makes every radicand non-negative; a constant shift of all
        distances) the folded code  e~_k = [e'_k, t_k, 0, 0, 0]  satisfies  |[z,0] - e~_k|^2 = |z|^2 + c_k + C - 2 z.e'_k,
        i.e. the SAME argmin from the RAW rows against a [K, D+4] codebook: no N x D x D projection, no projected
        copy of the rows.  The fold is K x D x D work (fp64, rounded once to fp32) whenever E, W or b change; the
        exact search then returns the exact argmin for that fp32 codebook, ~1e-4 absolute in the distances away
        from the reference's own fp32 chain -- far inside the near-tie tolerance.  The EMA sums over the projected
        rows follow from the raw sums: sum_n (W x_n + b) = W (sum_n x_n) + count b (functional.quantize)."""
        if not self.fold_projection:
            return None
        Wemb, Wp, bp = self._embedding.weight, self.pre_linear.weight, self.pre_linear.bias
        key = (Wemb.data_ptr(), Wemb._version, Wp.data_ptr(), Wp._version, bp.data_ptr(), bp._version, str(dev))
        if getattr(self, "_fold_key", None) != key:
            with torch.no_grad():
                K, D = Wemb.shape
                E_fold = torch.zeros(K, D + self._FOLD_PAD, dtype=torch.float32, device=dev)
                g = torch.empty(K, dtype=torch.float64, device=dev)
                # rows W^T e_k in fp64, rounded once, and g_k = c_k - |W^T e_k|^2 (g2v_fold_projection: K x D x D work)
                _lib.check(_lib.load().g2v_fold_projection(
                    F._ptr(Wemb.detach()), F._ptr(Wp.detach().contiguous()), F._ptr(bp.detach().contiguous()), K, D,
                    D + self._FOLD_PAD, F._ptr(E_fold), F._ptr(g), F._stream(dev)), "g2v_fold_projection")
                C = torch.clamp(-g.min(), min=0.0)
                E_fold[:, D] = torch.sqrt(g + C).float()
                self._fold_E = E_fold
                self._fold_cb = F.prepare_codebook(E_fold, getattr(self, "_fold_cb", None))
                self._fold_key = key
        return self._fold_E, self._fold_cb, Wp.detach(), bp.detach()

    def _search_rows(self, flat: torch.Tensor) -> Optional[torch.Tensor]:
        # pre_linear on the tensor cores at fp32 accuracy (g2v_gemm_f32: split-fp16 tcgen05 GEMM, bias in the
        # epilogue); no gradient reaches pre_linear
        return F.gemm(flat, self.pre_linear.weight.detach(), bias=self.pre_linear.bias.detach())


class VQVAE_VQ_Payam_GSSoft(nn.Module):
    """Autoencoder_VQVAE_model.py:1304-1433 -- the soft quantizer `Autoencoder_VQVAE.__init__` ends on (:816-820).

    Same constructor, attributes and state_dict keys (`pre_linear.*` exists and is unused, `_embedding.weight` ~
    N(0,1), `mean_layer.*`, `logvar_layer.*`), same forward contract; `encodings` is the dense [N, K] matrix of
    assignment probabilities.  The arithmetic runs in soft.py (tcgen05 GEMMs + the g2v_soft_* row kernels)."""

    def __init__(self, num_embeddings: int, embedding_dim: int, commitment_cost: float):
        super().__init__()
        self._embedding_dim = int(embedding_dim)
        self._num_embeddings = int(num_embeddings)
        self.pre_linear = nn.Linear(self._embedding_dim, self._embedding_dim)
        self._embedding = nn.Embedding(self._num_embeddings, self._embedding_dim)
        self._embedding.weight.data.normal_()
        self._commitment_cost = float(commitment_cost)
        self.mean_layer = nn.Linear(self._embedding_dim, self._embedding_dim)
        self.logvar_layer = nn.Linear(self._embedding_dim, self._num_embeddings)

    def embedding_grad(self, what: bool) -> None:
        for param in self._embedding.parameters():
            param.requires_grad = what

    def forward(self, inputs: torch.Tensor):
        from .soft import soft_quantize
        if inputs.numel() % self._embedding_dim:
            raise RuntimeError(
                f"shape '[-1, {self._embedding_dim}]' is invalid for input of size {inputs.numel()}")
        src_dev = inputs.device
        W0 = self._embedding.weight
        if not inputs.is_cuda:
            if not W0.is_cuda:
                raise RuntimeError(
                    "gesture2vec_b200 quantizers run on a CUDA device only: move the module to a "
                    "B200 (`.cuda()`); there is no CPU implementation of this path")
            inputs = inputs.to(W0.device)
        if inputs.dtype != torch.float32:
            inputs = inputs.float()
        flat = inputs.contiguous().view(-1, self._embedding_dim)
        loss, out, ppl, p = soft_quantize(flat, W0, self.mean_layer.weight, self.mean_layer.bias,
                                          self.logvar_layer.weight, self.logvar_layer.bias, self._commitment_cost)
        quantized = out.view(inputs.shape)
        if src_dev != quantized.device:
            loss, quantized, ppl, p = (t.to(src_dev) for t in (loss, quantized, ppl, p))
        return loss, quantized, ppl, p


class VectorQuantizerEMA(_HardQuantizerBase):
    """Autoencoder_VQVAE_model.py:1713-1812 (never instantiated by the reference; kept for parity).

    inputs [2, B, H] -> hstack(inputs[0], inputs[1]) [B, 2H] -> pre_lin -> EMA VQ on the projected
    rows (loss and STE DO use the projection here) -> literal reshape to (2, B, -1) (:1810)."""
    _ema = True

    def __init__(self, num_embeddings: int, embedding_dim: int, commitment_cost: float, decay: float,
                 epsilon: float = 1e-5):
        super().__init__()
        self._init_common(num_embeddings, embedding_dim, commitment_cost)
        self.pre_lin = nn.Linear(self._embedding_dim, self._embedding_dim)
        self._embedding = nn.Embedding(self._num_embeddings, self._embedding_dim)
        self._embedding.weight.data.normal_()
        self.register_buffer("_ema_cluster_size", torch.zeros(num_embeddings))
        self._ema_w = nn.Parameter(torch.Tensor(num_embeddings, self._embedding_dim))
        self._ema_w.data.normal_()
        self._decay = decay
        self._epsilon = epsilon

    def forward(self, inputs: torch.Tensor):
        rows = torch.hstack((inputs[0], inputs[1]))
        if rows.is_cuda:                                # differentiable: grads reach pre_lin here
            rows = F.linear(rows.float(), self.pre_lin.weight, self.pre_lin.bias)
        else:
            rows = self.pre_lin(rows)                   # CPU input: _run raises or moves it, like the other flavours
        loss, quantized, ppl, enc = self._run(rows)
        quantized = torch.reshape(quantized, (2, quantized.shape[0], -1)).contiguous()
        return loss, quantized, ppl, enc


FLAVOURS = {
    "dae": {"VQ_Payam": DAE_VQ_Payam, "VQ_Payam_EMA": DAE_VQ_Payam_EMA},
    "vqvae": {"VQ_Payam": VQVAE_VQ_Payam, "VQ_Payam_EMA": VQVAE_VQ_Payam_EMA,
              "VectorQuantizerEMA": VectorQuantizerEMA, "VQ_Payam_GSSoft": VQVAE_VQ_Payam_GSSoft},
}
