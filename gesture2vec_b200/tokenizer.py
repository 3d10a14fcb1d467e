"""Batched gesture tokenisation: the reference's per-chunk loops as one search.

The reference turns every gesture chunk into a code id one chunk at a time -- ``Clustering.py:102-166``
(encoder -> ``vq_layer(decoder_hidden)`` with a batch of ONE -> ``np.argmax(encodings, 1)`` -> pickle entry
``quantized_indices``), ``inference_Autoencoder.py:124-182`` and ``lmdb_data_loader.py:1273-1281`` (sentence
latents -> ``argmax(encodings)``), each with a device->host copy per chunk.  Here all chunks of a file (or of
the data set) go through one nearest-code search and the ids come back in the layouts those callers write.

Row layout: a chunk's latent is the last hidden state ``[n_layers, 1, hidden]`` of the GRU encoder; with a batch
of one, ``inputs.view(-1, D)`` (D = n_layers * hidden) is the row ``[layer0 | layer1]``.
The batched form must build that row per chunk -- ``hidden.transpose(0, 1).reshape(B, D)`` -- NOT ``view(-1, D)`` of
the ``[n_layers, B, hidden]`` tensor, which would pair adjacent batch items of one layer (the a1 quirk).
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence, Union

import numpy as np
import torch

from .functional import prepare_codebook, tokenize_host, vq_search

ArrayLike = Union[np.ndarray, torch.Tensor]


def chunk_rows_from_hidden(hidden: ArrayLike) -> ArrayLike:
    """``[n_layers, B, hidden]`` (encoder_hidden[:n_layers]) -> ``[B, n_layers * hidden]``, row b = what
    ``vq_layer(hidden[:, b:b+1])`` flattens to (Clustering.py:138-152 with its batch of one)."""
    if hidden.ndim != 3:
        raise ValueError("expected [n_layers, batch, hidden]")
    L, B, H = hidden.shape
    if isinstance(hidden, torch.Tensor):
        return hidden.transpose(0, 1).reshape(B, L * H).contiguous()
    return np.ascontiguousarray(np.transpose(hidden, (1, 0, 2)).reshape(B, L * H))


class GestureTokenizer:
    """Code ids for many chunks at once, against a quantizer module's codebook (``_embedding.weight``) or a
    ``[K, D]`` array.  The fp16/||e||^2 aux buffer of the codebook is prepared once and reused."""

    def __init__(self, codebook, device: Union[str, torch.device, None] = None, host_chunk_rows: int = 131072):
        # a quantizer module whose search runs on a projection of the rows (Autoencoder_VQVAE_model.VQ_Payam_EMA
        # searches pre_linear(z), :1230; VectorQuantizerEMA its pre_lin, :1755): the ids must be those of
        # module.forward, so the rows go through the module's own projection first
        self._project = None
        self._module = None
        if getattr(codebook, "_projects", False) and hasattr(codebook, "tokenize"):
            self._module = codebook            # its tokenize() searches the raw rows against the folded codebook
            self._project = codebook._search_rows
        elif hasattr(codebook, "pre_lin") and hasattr(codebook, "_embedding"):
            from .functional import gemm
            self._project = lambda rows, _m=codebook: gemm(rows, _m.pre_lin.weight.detach(), bias=_m.pre_lin.bias.detach())
        w = getattr(getattr(codebook, "_embedding", None), "weight", codebook)
        if isinstance(w, np.ndarray):
            w = torch.from_numpy(np.ascontiguousarray(w, dtype=np.float32))
        w = w.detach().to(torch.float32)
        dev = torch.device(device) if device is not None else (
            w.device if w.is_cuda else torch.device("cuda", torch.cuda.current_device()))
        self.E = w.to(dev).contiguous()
        self.K, self.D = self.E.shape
        self.cb = prepare_codebook(self.E)
        self.host_chunk_rows = int(host_chunk_rows)

    def encode_rows(self, rows: ArrayLike) -> np.ndarray:
        """``[N, D]`` latent rows -> int64 ids ``[N]`` (the dtype ``np.argmax`` gives the reference's callers).
        Host arrays stream through the pinned-buffer entry point; CUDA tensors are searched in place."""
        if rows.ndim != 2 or rows.shape[1] != self.D:
            raise ValueError(f"expected rows of {self.D} values, got {tuple(rows.shape)}")
        if rows.shape[0] == 0:
            return np.zeros(0, dtype=np.int64)
        if isinstance(rows, torch.Tensor) and rows.is_cuda:
            r = rows if rows.dtype in (torch.float32, torch.bfloat16, torch.float16) else rows.float()
            if self._module is not None:
                return self._module.tokenize(r).cpu().numpy().astype(np.int64)
            if self._project is not None:
                with torch.no_grad():
                    r = self._project(r.float().contiguous())
            return vq_search(r.contiguous(), self.E, self.cb).cpu().numpy().astype(np.int64)
        t = torch.from_numpy(np.ascontiguousarray(rows, dtype=np.float32)) if isinstance(rows, np.ndarray) else rows
        if t.dtype not in (torch.float32, torch.bfloat16, torch.float16):
            t = t.float()
        t = t.contiguous()
        if self._project is not None:
            # projected flavour: chunks go to the device, through the projection, and are searched there
            out = np.empty(t.shape[0], dtype=np.int64)
            for r0 in range(0, t.shape[0], self.host_chunk_rows):
                blk = t[r0:r0 + self.host_chunk_rows].to(self.E.device, non_blocking=True)
                if self._module is not None:
                    out[r0:r0 + blk.shape[0]] = self._module.tokenize(blk).cpu().numpy()
                    continue
                with torch.no_grad():
                    blk = self._project(blk.float().contiguous())
                out[r0:r0 + blk.shape[0]] = vq_search(blk.contiguous(), self.E, self.cb).cpu().numpy()
            return out
        if not t.is_pinned() and t.numel() >= (1 << 22):
            t = t.pin_memory()
        return tokenize_host(t, self.E, self.cb, chunk_rows=self.host_chunk_rows).numpy().astype(np.int64)

    def encode_hidden(self, hidden: ArrayLike) -> np.ndarray:
        """``[n_layers, B, hidden]`` encoder states of B chunks -> ids ``[B]``."""
        return self.encode_rows(chunk_rows_from_hidden(hidden))

    def cluster_ids(self, sentence_latents: ArrayLike) -> torch.Tensor:
        """``lmdb_data_loader.py:1273-1281``: ``torch.argmax(encodings, dim=1)`` for ``[n, D]`` latents (int64 tensor)."""
        return torch.from_numpy(self.encode_rows(sentence_latents))

    def clustering_entries(self, hidden_per_chunk: Union[ArrayLike, Sequence[ArrayLike]],
                           entries: Optional[List[dict]] = None) -> List[dict]:
        """The ``quantized_indices`` / ``latent_rnn`` fields of the per-chunk dicts ``Clustering.py:155-163`` builds:
        ``quantized_indices`` is an int64 array of shape ``[1]``, ``latent_rnn`` the ``[n_layers, hidden]`` state.
        `hidden_per_chunk`: ``[n_layers, B, hidden]`` or a sequence of ``[n_layers, 1, hidden]`` states.  Existing
        dicts (one per chunk) are filled in place when given."""
        if not isinstance(hidden_per_chunk, (np.ndarray, torch.Tensor)):
            seq = list(hidden_per_chunk)
            if len(seq) == 0:
                return entries if entries is not None else []
            if isinstance(seq[0], torch.Tensor):
                hidden = torch.cat([h.reshape(h.shape[0], 1, -1) for h in seq], dim=1)
            else:
                hidden = np.concatenate([np.asarray(h).reshape(np.asarray(h).shape[0], 1, -1) for h in seq], axis=1)
        else:
            hidden = hidden_per_chunk
        ids = self.encode_hidden(hidden)
        B = ids.shape[0]
        if entries is None:
            entries = [dict() for _ in range(B)]
        if len(entries) != B:
            raise ValueError("one dict per chunk expected")
        lat = hidden.detach().cpu().numpy() if isinstance(hidden, torch.Tensor) else np.asarray(hidden)
        for b, e in enumerate(entries):
            e["quantized_indices"] = ids[b:b + 1].copy()
            e["latent_rnn"] = np.ascontiguousarray(lat[:, b, :])
        return entries
