"""Tensor-level entry points over the C ABI: search, apply, statistics, EMA, backward, and the
autograd.Function the quantizer modules use.  PyTorch owns every buffer; the C side only
enqueues kernels on the current stream of the tensors' device.

A training step is at most five launches of this library's kernels:
  g2v_vq_search         tcgen05 sweep (or the fp32 sweep) + one exact re-rank launch
  g2v_vq_apply          gather, straight-through value, SSE / histogram / EMA residual sums
  g2v_vq_step_finalize  pack, loss, perplexity, EMA update, codebook aux re-preparation (one cooperative launch)
  g2v_vq_backward       (in backward) gradient wrt the inputs
(data-parallel EMA adds one pack launch in front of the all-reduce).
"""
from __future__ import annotations

import contextlib
import ctypes as C
from typing import Optional, Tuple

import torch

from . import _lib

_DT = {torch.float32: _lib.F32, torch.bfloat16: _lib.BF16, torch.float16: _lib.F16}
UPDATE_NONE, UPDATE_EMA, UPDATE_KMEANS = 0, 1, 2


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream(dev) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


_NULL = contextlib.nullcontext()


def _on(dev: torch.device):
    """Make `dev` the CUDA runtime's current device around a library call: kernels, memsets, the arch
    check and the SM count all refer to the current device, while the stream we pass belongs to the
    tensors' device (a module on cuda:1 with current device 0 must work, as the reference's do)."""
    if dev.index is None or dev.index == torch.cuda.current_device():
        return _NULL
    return torch.cuda.device(dev)


def _need_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(
            f"gesture2vec_b200: `{name}` must be a CUDA tensor; the quantizer path has no CPU "
            "implementation (move the tensor to a B200, or use the reference module on CPU)")


class _Scratch:
    """Grow-only per-(device, tag) byte buffers, so steady-state steps allocate nothing."""

    def __init__(self):
        self._bufs = {}

    def get(self, dev: torch.device, tag: str, nbytes: int) -> torch.Tensor:
        key = (dev.index, tag)
        b = self._bufs.get(key)
        if b is None or b.numel() < nbytes:
            # cudaMalloc'd blocks from the caching allocator are >= 512 B aligned
            b = torch.empty(max(int(nbytes), 1024), dtype=torch.uint8, device=dev)
            self._bufs[key] = b
        return b


_scratch = _Scratch()


class _Accum:
    """Accumulators of the row pass (sse, counts, dwr replicas) for one (device, K, D, replicas).

    They are zero whenever `clean` is True: g2v_vq_step_finalize hands them back zeroed after packing, so
    a steady-state step issues no memset for them.  A step that dies between the row pass and the pack
    leaves `clean` False and the next user zeroes them explicitly."""

    def __init__(self, dev: torch.device, K: int, D: int, reps: int):
        self.head = torch.zeros(8 + 4 * K, dtype=torch.uint8, device=dev)      # double sse | int32 counts[K]
        self.sse = self.head[:8].view(torch.float64)
        self.counts = self.head[8:].view(torch.int32)
        self.dwr = torch.zeros(reps * K * D, dtype=torch.float32, device=dev) if reps else None
        self.reps = reps
        self.clean = True

    def take(self) -> "_Accum":
        if not self.clean:
            self.head.zero_()
            if self.dwr is not None:
                self.dwr.zero_()
        self.clean = False
        return self


_accums = {}


def _accum(dev: torch.device, K: int, D: int, reps: int) -> _Accum:
    key = (dev.index, K, D, reps)
    a = _accums.get(key)
    if a is None:
        a = _accums[key] = _Accum(dev, K, D, reps)
    return a.take()


def dwr_replicas(N: int, K: int, D: int) -> int:
    """Private copies of the [K, D] sums that keep hot codes from serialising the fp32 atomics: large
    batches get up to 8 copies (bounded to 64 MB), small ones a single copy."""
    return 1 if N < 32768 else max(1, min(8, (64 << 20) // (K * D * 4)))


# ------------------------------------------------------------------------------------------------
# codebook aux
# ------------------------------------------------------------------------------------------------
def codebook_bytes(K: int, D: int) -> int:
    return int(_lib.load().g2v_codebook_bytes(K, D))


def prepare_codebook(E: torch.Tensor, cb: Optional[torch.Tensor] = None) -> torch.Tensor:
    """||e||^2, norm maxima and the fp16 operand copy of E (re-run whenever E changes)."""
    _need_cuda(E, "codebook")
    assert E.dtype == torch.float32 and E.is_contiguous() and E.dim() == 2
    K, D = E.shape
    lib = _lib.load()
    nbytes = codebook_bytes(K, D)
    if cb is None or cb.numel() < nbytes or cb.device != E.device:
        cb = torch.empty(nbytes, dtype=torch.uint8, device=E.device)
    with _on(E.device):
        _lib.check(lib.g2v_codebook_prepare(_ptr(E), K, D, _ptr(cb), cb.numel(), _stream(E.device)),
                   "g2v_codebook_prepare")
    return cb


# ------------------------------------------------------------------------------------------------
# search
# ------------------------------------------------------------------------------------------------
def vq_search(z: torch.Tensor, E: torch.Tensor, cb: Optional[torch.Tensor] = None, *,
              flags: int = _lib.ALGO_AUTO, stats: Optional[torch.Tensor] = None,
              out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """idx[n] = argmin_k ||z_n - e_k||^2 (int32, first index on exact ties).

    z: [N, D] fp32 / bf16 / fp16 CUDA rows; E: [K, D] fp32.  `stats`: optional int64[8] CUDA
    tensor the kernels accumulate their re-rank counters into."""
    _need_cuda(z, "z")
    _need_cuda(E, "codebook")
    if z.dtype not in _DT:
        raise RuntimeError(f"unsupported latent dtype {z.dtype}")
    assert z.dim() == 2 and z.is_contiguous() and E.is_contiguous() and E.dtype == torch.float32
    if z.device != E.device:
        raise RuntimeError(f"rows are on {z.device}, the codebook on {E.device}")
    N, D = z.shape
    K = E.shape[0]
    assert E.shape[1] == D, "latent dim != codebook dim"
    lib = _lib.load()
    if cb is None:
        cb = prepare_codebook(E)
    idx = out if out is not None else torch.empty(N, dtype=torch.int32, device=z.device)
    if N == 0:
        return idx
    dt = _DT[z.dtype]
    with _on(z.device):
        ws = _scratch.get(z.device, "search", lib.g2v_workspace_bytes(N, K, D, dt, flags))
        _lib.check(lib.g2v_vq_search(_ptr(z), dt, _ptr(E), _ptr(cb), N, K, D, _ptr(idx), _ptr(stats),
                                     _ptr(ws), ws.numel(), flags, _stream(z.device)), "g2v_vq_search")
    return idx


def vq_search_wide(z: torch.Tensor, E: torch.Tensor, cb: torch.Tensor, *, flags: int = _lib.ALGO_AUTO,
                   stats: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Nearest code of rows that are NARROWER than the codebook (z [N, Dz], E [K, D], Dz <= D; the rows count as
    [z | 0]): how a codebook with a projection folded in is searched with the raw rows.  The tensor-core sweep reads
    the rows through TMA and lets it zero-fill the missing columns (g2v_vq_search_wide); shapes it does not cover
    get a widened copy of the rows (g2v_pad_rows) and the ordinary search."""
    _need_cuda(z, "z")
    N, Dz = z.shape
    K, D = E.shape
    if Dz == D:
        return vq_search(z, E, cb, flags=flags, stats=stats)
    lib = _lib.load()
    if z.dtype in _DT and z.is_contiguous() and N > 0:
        dt = _DT[z.dtype]
        with _on(z.device):
            idx = torch.empty(N, dtype=torch.int32, device=z.device)
            ws = _scratch.get(z.device, "search", lib.g2v_workspace_bytes(N, K, D, dt, _lib.ALGO_TC))
            rc = lib.g2v_vq_search_wide(_ptr(z), dt, Dz, _ptr(E), _ptr(cb), N, K, D, _ptr(idx), _ptr(stats), _ptr(ws),
                                        ws.numel(), flags & ~_lib.ALGO_MASK, _stream(z.device))
        if rc == 0:
            return idx
        if rc != _lib.ERR_UNSUPPORTED:
            _lib.check(rc, "g2v_vq_search_wide")
    return vq_search(pad_rows(z.float().contiguous(), D), E, cb, flags=flags, stats=stats)


def vq_search_exact(z: torch.Tensor, E: torch.Tensor) -> torch.Tensor:
    """Verification aid: exact-arithmetic (fp64) nearest code of every row, first index on ties (int32).
    Slow by design (FP64-bound); shares nothing with the fast search paths."""
    _need_cuda(z, "z")
    _need_cuda(E, "codebook")
    assert z.dim() == 2 and z.is_contiguous() and E.is_contiguous() and E.dtype == torch.float32 and z.dtype in _DT
    N, D = z.shape
    K = E.shape[0]
    lib = _lib.load()
    idx = torch.empty(N, dtype=torch.int32, device=z.device)
    if N == 0:
        return idx
    with _on(z.device):
        ws = torch.empty(int(lib.g2v_exact_workspace_bytes(K)), dtype=torch.uint8, device=z.device)
        _lib.check(lib.g2v_vq_search_exact(_ptr(z), _DT[z.dtype], _ptr(E), N, K, D, _ptr(idx), _ptr(ws), ws.numel(),
                                           _stream(z.device)), "g2v_vq_search_exact")
    return idx


def pad_rows(x: torch.Tensor, Dp: int) -> torch.Tensor:
    """[N, D] fp32 rows -> [N, Dp] with zeros in the extra columns (g2v_pad_rows)."""
    N, D = x.shape
    if D % 4 or Dp % 4 or x.dtype != torch.float32 or not x.is_contiguous() or (x.data_ptr() & 15):
        return torch.nn.functional.pad(x.float(), (0, Dp - D)).contiguous()
    with _on(x.device):
        out = torch.empty(N, Dp, dtype=torch.float32, device=x.device)
        _lib.check(_lib.load().g2v_pad_rows(_ptr(x), N, D, Dp, _ptr(out), _stream(x.device)), "g2v_pad_rows")
    return out


def packed_numel(K: int, D: int) -> int:
    return K * D + K + 2


def _apply_rows(x, zs, E, idx, out, acc: Optional[_Accum], want_dwr: bool, deterministic: bool = False) -> None:
    """g2v_vq_apply: out = x + (E[idx]-x) and, if `acc`, SSE / counts (/ residual sums) into the accumulators.
    deterministic: the statistics come from the atomics-free fixed-order pass instead (acc must hold ONE dwr copy)."""
    N, D = x.shape
    K = E.shape[0]
    if deterministic and acc is not None:
        assert acc.reps == 1, "deterministic statistics use a single accumulator copy"
        _deterministic_stats(x, zs, E, idx, acc)
        if out is not None:
            _lib.check(_lib.load().g2v_vq_apply(_ptr(x), None, _ptr(E), _ptr(idx), N, K, D, _ptr(out), None, None, None, 0,
                                                _stream(x.device)), "g2v_vq_apply")
        return
    dwr = acc.dwr if (acc is not None and want_dwr) else None
    lib = _lib.load()
    # bulk passes with residual sums group the rows by code first (scratch: the order array)
    wsb = lib.g2v_apply_workspace_bytes(N, K) if dwr is not None else 0
    ws = _scratch.get(x.device, "apply", wsb) if wsb else None
    _lib.check(lib.g2v_vq_apply_ws(
        _ptr(x), _ptr(zs), _ptr(E), _ptr(idx), N, K, D, _ptr(out),
        _ptr(acc.sse) if acc is not None else None, _ptr(acc.counts) if acc is not None else None,
        _ptr(dwr), acc.reps if dwr is not None else 0, _ptr(ws), wsb, _stream(x.device)), "g2v_vq_apply_ws")


def _deterministic_stats(x, zs, E, idx, acc: _Accum) -> None:
    """Fill the accumulators reproducibly (g2v_vq_stats_deterministic): rows sorted by code with a stable sort
    (plumbing: torch.sort / bincount / cumsum), then fixed-order fp64 chunk sums -- no atomics anywhere."""
    N, D = x.shape
    K = E.shape[0]
    dev = x.device
    idx64 = idx.long()
    order = torch.sort(idx64, stable=True).indices.to(torch.int32)
    counts = torch.bincount(idx64, minlength=K)
    seg = torch.zeros(K + 1, dtype=torch.int64, device=dev)
    seg[1:] = torch.cumsum(counts, 0)
    chunk_off = torch.zeros(K + 1, dtype=torch.int64, device=dev)
    chunk_off[1:] = torch.cumsum((counts + (_lib.DET_CHUNK - 1)) // _lib.DET_CHUNK, 0)
    max_chunks = N // _lib.DET_CHUNK + K
    partial = torch.empty(max_chunks * (D + 1), dtype=torch.float64, device=dev)
    sse_code = torch.empty(K, dtype=torch.float64, device=dev)
    _lib.check(_lib.load().g2v_vq_stats_deterministic(
        _ptr(x), _ptr(zs), _ptr(E), _ptr(order), _ptr(seg), _ptr(chunk_off), max_chunks, K, D, _ptr(partial),
        _ptr(acc.dwr), _ptr(sse_code), _ptr(acc.sse), _stream(dev)), "g2v_vq_stats_deterministic")
    acc.counts.copy_(counts)


def step_finalize(K: int, D: int, packed: torch.Tensor, *, acc: Optional[_Accum] = None, use_dwr: bool = True,
                  rows_local: int = 0, coefs: Optional[Tuple[float, float]] = None, update: int = UPDATE_NONE,
                  cs_in=None, cs_out=None, w_in=None, w_out=None, E_old=None, E_new=None, E_prev=None, decay: float = 0.0,
                  eps: float = 0.0, shift2=None, cb=None, res: Optional[torch.Tensor] = None) -> Optional[torch.Tensor]:
    """g2v_vq_step_finalize.  `acc`: pack these accumulators first (they come back zeroed); otherwise `packed`
    already holds the statistics.  `coefs` = (coef_codebook, coef_commit): also produce [loss, perplexity]
    (returned as a 2-element device tensor)."""
    lib = _lib.load()
    dev = packed.device
    if res is None and coefs is not None:
        res = torch.empty(2, dtype=torch.float32, device=dev)
    cc, cm = coefs if coefs is not None else (0.0, 0.0)
    dwr = acc.dwr if (acc is not None and use_dwr) else None
    rc = lib.g2v_vq_step_finalize(
        _ptr(acc.counts) if acc is not None else None, _ptr(acc.sse) if acc is not None else None, _ptr(dwr),
        acc.reps if dwr is not None else 0, rows_local, _ptr(packed), K, D, cc, cm,
        C.c_void_p(res.data_ptr()) if res is not None else None,
        C.c_void_p(res.data_ptr() + 4) if res is not None else None,
        update, _ptr(cs_in), _ptr(cs_out), _ptr(w_in), _ptr(w_out), _ptr(E_old), _ptr(E_new), _ptr(E_prev), decay, eps,
        _ptr(shift2), _ptr(cb), 0 if cb is None else cb.numel(), _stream(dev))
    _lib.check(rc, "g2v_vq_step_finalize")
    if acc is not None:
        # the pack pass zeroes what it read: sse / counts always, the residual sums when they were packed
        acc.clean = use_dwr or acc.dwr is None
    return res


def vq_apply(x: torch.Tensor, E: torch.Tensor, idx: torch.Tensor, *, zs: Optional[torch.Tensor] = None,
             want_out: bool = True, want_stats: bool = True, want_dwr: bool = False, deterministic: bool = False
             ) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
    """One pass over the rows: out = x + (E[idx]-x) and, if asked, the packed statistics buffer
    [dwr (K*D) | counts (K) | sse | rows] (fp32, ready for an all-reduce).  deterministic=True: bit-reproducible
    statistics (sorted, fixed-order fp64 sums instead of fp32 atomics)."""
    N, D = x.shape
    K = E.shape[0]
    dev = x.device
    with _on(dev):
        out = torch.empty_like(x) if want_out else None
        det = deterministic and want_stats and N > 0
        acc = _accum(dev, K, D, 1 if det else (dwr_replicas(N, K, D) if want_dwr else 0)) if want_stats else None
        _apply_rows(x, zs, E, idx, out, acc, want_dwr or det, det)
        want_dwr = want_dwr or det
        packed = None
        if want_stats:
            packed = torch.empty(packed_numel(K, D), dtype=torch.float32, device=dev)
            step_finalize(K, D, packed, acc=acc, use_dwr=want_dwr, rows_local=N)
    return out, packed


def stats_finalize(packed: torch.Tensor, K: int, D: int, coef_codebook: float, coef_commit: float
                   ) -> Tuple[torch.Tensor, torch.Tensor]:
    with _on(packed.device):
        res = step_finalize(K, D, packed, coefs=(coef_codebook, coef_commit))
    return res[0], res[1]


def ema_update(cs_in: torch.Tensor, cs_out: torch.Tensor, w_in: torch.Tensor, w_out: torch.Tensor,
               E_old: torch.Tensor, E_new: torch.Tensor, packed: torch.Tensor, decay: float, eps: float,
               cb: Optional[torch.Tensor]) -> None:
    """EMA update from a packed statistics buffer (cs_out must not alias cs_in); `cb` is re-prepared for E_new."""
    K, D = E_old.shape
    with _on(E_old.device):
        step_finalize(K, D, packed, update=UPDATE_EMA, cs_in=cs_in, cs_out=cs_out, w_in=w_in, w_out=w_out,
                      E_old=E_old, E_new=E_new, decay=decay, eps=eps, cb=cb)


def one_hot(idx: torch.Tensor, K: int) -> torch.Tensor:
    lib = _lib.load()
    N = idx.numel()
    with _on(idx.device):
        enc = torch.empty(N, K, dtype=torch.float32, device=idx.device)
        _lib.check(lib.g2v_onehot(_ptr(idx), N, K, _ptr(enc), _stream(idx.device)), "g2v_onehot")
    return enc


# ------------------------------------------------------------------------------------------------
# dense projections on the tensor cores (csrc/g2v_gemm.cu)
# ------------------------------------------------------------------------------------------------
def gemm(A: torch.Tensor, B: torch.Tensor, *, bias: Optional[torch.Tensor] = None, transA: bool = False,
         transB: bool = False, out: Optional[torch.Tensor] = None, alpha: float = 1.0, accumulate: bool = False,
         fp16: bool = False) -> torch.Tensor:
    """C = alpha * op(A) @ op(B)^T + bias on the tcgen05 tensor cores at fp32 accuracy (split-fp16 operands).

    A: [M, K] fp32 (or [K, M] with transA), B: [N, K] fp32 -- nn.Linear's weight layout -- (or [K, N] with transB);
    rows may be strided (stride(1) == 1).  `fp16=True`: one fp16 term per operand (for very long reductions)."""
    _need_cuda(A, "A")
    _need_cuda(B, "B")
    assert A.dtype == torch.float32 and B.dtype == torch.float32 and A.dim() == 2 and B.dim() == 2
    if A.stride(1) != 1:
        A = A.contiguous()
    if B.stride(1) != 1:
        B = B.contiguous()
    M, K = (A.shape[1], A.shape[0]) if transA else A.shape
    N, Kb = (B.shape[1], B.shape[0]) if transB else B.shape
    assert K == Kb, "reduction lengths differ"
    dev = A.device
    lib = _lib.load()
    flags = (_lib.GEMM_ACCUMULATE if accumulate else 0) | (_lib.GEMM_FP16 if fp16 else 0)
    with _on(dev):
        if out is None:
            assert not accumulate
            out = torch.empty(M, N, dtype=torch.float32, device=dev)
        assert out.shape == (M, N) and out.stride(1) == 1 and out.dtype == torch.float32
        ws = _scratch.get(dev, "gemm", lib.g2v_gemm_workspace_bytes(M, N, K, flags))
        _lib.check(lib.g2v_gemm_f32(_ptr(A), A.stride(0), int(transA), _ptr(B), B.stride(0), int(transB), M, N, K,
                                    _ptr(bias), _ptr(out), out.stride(0), float(alpha), flags, _ptr(ws), ws.numel(),
                                    _stream(dev)), "g2v_gemm_f32")
    return out


class _LinearFn(torch.autograd.Function):
    """y = x W^T + b on g2v_gemm_f32, with the two backward products on it as well."""

    @staticmethod
    def forward(ctx, x, W, b):
        ctx.save_for_backward(x, W)
        ctx.has_bias = b is not None
        return gemm(x, W.detach(), bias=None if b is None else b.detach())

    @staticmethod
    def backward(ctx, g):
        x, W = ctx.saved_tensors
        g = g.contiguous()
        gx = gemm(g, W.detach(), transB=True) if ctx.needs_input_grad[0] else None
        gW = gemm(g, x, transA=True, transB=True) if ctx.needs_input_grad[1] else None
        gb = g.sum(0) if (ctx.has_bias and ctx.needs_input_grad[2]) else None
        return gx, gW, gb


def linear(x: torch.Tensor, W: torch.Tensor, b: Optional[torch.Tensor] = None) -> torch.Tensor:
    """nn.Linear's arithmetic (fp32 in / out) on the tcgen05 GEMM; differentiable."""
    return _LinearFn.apply(x.contiguous(), W, b)


# ------------------------------------------------------------------------------------------------
# autograd
# ------------------------------------------------------------------------------------------------
class EmaState:
    """EMA inputs of one step and, after `quantize`, its outputs: fresh tensors by default, like the reference's
    re-created Parameters (Autoencoder_VQVAE_model.py:1276-1282).  With `E_prev` (a persistent [K, D] buffer) the
    state is updated IN PLACE instead -- every address stays fixed, so a captured CUDA graph of the step advances
    the state on every replay -- and the old codebook the backward pass needs is kept in `E_prev`."""

    def __init__(self, cluster_size: torch.Tensor, ema_w: torch.Tensor, decay: float, eps: float,
                 E_prev: Optional[torch.Tensor] = None):
        self.cs_in, self.w_in, self.decay, self.eps = cluster_size, ema_w, float(decay), float(eps)
        self.E_prev = E_prev
        self.cs_out = self.w_out = self.E_new = None
        self.pending: Optional[torch.cuda.Event] = None      # side-stream work the caller must wait for


class _QuantizeFn(torch.autograd.Function):
    """(x, E) -> (loss, quantized, perplexity, idx, packed).

    Closed-form backward of the graph the reference builds:
      d/dx = g_quantized + g_loss * 2*beta/M * grad_scale * (x - E[idx])
      d/dE = -g_loss * 2/M * dwr                      (hard VQ only)
    """

    @staticmethod
    def forward(ctx, x2d, E, zs, cb, beta, coef_codebook, want_dwr, reduce_fn, grad_scale, flags, ema, idx_given, det,
                dw_transform):
        dev = x2d.device
        N, D = x2d.shape
        K = E.shape[0]
        with _on(dev):
            idx = idx_given if idx_given is not None else vq_search(x2d if zs is None else zs, E, cb, flags=flags)
            out = torch.empty_like(x2d)
            packed = torch.empty(packed_numel(K, D), dtype=torch.float32, device=dev)
            if N == 0:
                packed.zero_()
                acc = None
            else:
                want_dwr = want_dwr or det
                acc = _accum(dev, K, D, 1 if det else (dwr_replicas(N, K, D) if want_dwr else 0))
                _apply_rows(x2d, zs, E, idx, out, acc, want_dwr, det)
            kw = dict(coefs=(coef_codebook, beta), res=torch.empty(2, dtype=torch.float32, device=dev))
            E_bwd = E
            if ema is not None:
                if ema.E_prev is not None:
                    ema.cs_out, ema.w_out, ema.E_new, E_bwd = ema.cs_in, ema.w_in, E, ema.E_prev
                else:
                    ema.cs_out = torch.empty_like(ema.cs_in)
                    ema.w_out = torch.empty_like(ema.w_in)
                    ema.E_new = torch.empty_like(E)
                kw.update(update=UPDATE_EMA, cs_in=ema.cs_in, cs_out=ema.cs_out, w_in=ema.w_in, w_out=ema.w_out,
                          E_old=E, E_new=ema.E_new, E_prev=ema.E_prev, decay=ema.decay, eps=ema.eps, cb=cb)
            def project_sums():
                # the EMA sums are taken over pre_linear(x): sum_n (W x_n + b) = W (sum_n x_n) + count b, from the
                # raw residual sums of the row pass; K x D x D work on g2v_gemm_f32 instead of N x D x D
                if dw_transform is None:
                    return
                Wp, bp = dw_transform
                KD = K * D
                cnt = packed[KD:KD + K].unsqueeze(1)
                dwr = packed[:KD].view(K, D)
                S = torch.addcmul(dwr, cnt, E)                            # raw sums of the rows of each code
                dwp = gemm(S, Wp).addcmul_(cnt, bp.unsqueeze(0))          # sums of the projected rows
                dwr.copy_(dwp.addcmul_(cnt, E, value=-1.0))               # back to the residual form the finalise takes

            if reduce_fn is None and dw_transform is None:
                res = step_finalize(K, D, packed, acc=acc, use_dwr=want_dwr, rows_local=N, **kw)
            elif reduce_fn is None:
                step_finalize(K, D, packed, acc=acc, use_dwr=want_dwr, rows_local=N)
                project_sums()
                res = step_finalize(K, D, packed, **kw)
            else:
                # data-parallel: pack, ONE sum all-reduce of the packed statistics, then the identical finalise
                # on every rank.  With a side stream (StatsAllReduce(overlap=True)) the exchange and the
                # finalise overlap whatever the caller enqueues next on the main stream (the dense one-hot).
                step_finalize(K, D, packed, acc=acc, use_dwr=want_dwr, rows_local=N)
                side = getattr(reduce_fn, "stream", None)
                if side is None or ema is None:
                    reduce_fn(packed)
                    project_sums()
                    res = step_finalize(K, D, packed, **kw)
                else:
                    main = torch.cuda.current_stream(dev)
                    side.wait_stream(main)
                    with torch.cuda.stream(side):
                        reduce_fn(packed)
                        project_sums()
                        res = step_finalize(K, D, packed, **kw)
                        ema.pending = torch.cuda.Event()
                        ema.pending.record(side)
        loss, ppl = res[0], res[1]
        ctx.set_materialize_grads(False)      # an output nobody differentiates gets None, not a zero tensor of its size
        ctx.save_for_backward(x2d, E_bwd, idx, packed)
        ctx.beta, ctx.coef_codebook, ctx.grad_scale = beta, coef_codebook, grad_scale
        ctx.mark_non_differentiable(ppl, idx, packed)
        return loss, out, ppl, idx, packed

    @staticmethod
    def backward(ctx, g_loss, g_out, _g_ppl, _g_idx, _g_packed):
        x2d, E, idx, packed = ctx.saved_tensors
        lib = _lib.load()
        N, D = x2d.shape
        K = E.shape[0]
        dev = x2d.device
        M = float(N) * float(D)
        gx = gE = None
        with _on(dev):
            st = _stream(dev)
            if g_loss is None:
                g_loss = torch.zeros((), dtype=torch.float32, device=dev)
            g_loss = g_loss.to(torch.float32).contiguous()
            if ctx.needs_input_grad[0]:
                gx = torch.empty_like(x2d)
                if g_out is not None:
                    g_out = g_out.contiguous()
                _lib.check(lib.g2v_vq_backward(_ptr(x2d), _ptr(E), _ptr(idx), _ptr(g_out), _ptr(g_loss),
                                               2.0 * ctx.beta * ctx.grad_scale / M if M else 0.0, N, K, D, _ptr(gx), st),
                           "g2v_vq_backward")
            if ctx.needs_input_grad[1] and ctx.coef_codebook != 0.0:
                gE = torch.empty_like(E)
                _lib.check(lib.g2v_vq_grad_codebook(_ptr(packed), _ptr(g_loss),
                                                    2.0 * ctx.coef_codebook * ctx.grad_scale / M if M else 0.0,
                                                    K, D, _ptr(gE), st), "g2v_vq_grad_codebook")
        return gx, gE, None, None, None, None, None, None, None, None, None, None, None, None


def quantize(x2d, E, *, zs=None, cb=None, beta=0.25, coef_codebook=1.0, want_dwr=False,
             reduce_fn=None, grad_scale=1.0, flags=_lib.ALGO_AUTO, ema: Optional[EmaState] = None,
             idx: Optional[torch.Tensor] = None, deterministic: bool = False, dw_transform=None):
    """The whole layer on [N, D] rows.  deterministic: bit-reproducible statistics (see vq_apply).  `idx`: int32 code ids to use instead of searching (rows tokenised
    earlier; also how the parity tests evaluate the downstream arithmetic at the reference's indices)."""
    if idx is not None:
        idx = idx.to(device=x2d.device, dtype=torch.int32).contiguous()
        if idx.numel() != x2d.shape[0]:
            raise RuntimeError("one code id per row expected")
    return _QuantizeFn.apply(x2d, E, zs, cb, float(beta), float(coef_codebook), bool(want_dwr),
                             reduce_fn, float(grad_scale), int(flags), ema, idx, bool(deterministic), dw_transform)


# ------------------------------------------------------------------------------------------------
# bulk tokenisation
# ------------------------------------------------------------------------------------------------
def tokenize(z: torch.Tensor, E: torch.Tensor, cb: Optional[torch.Tensor] = None, *,
             flags: int = _lib.ALGO_AUTO, stats: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Device-resident bulk tokenisation: [N, D] rows -> int32 code ids (no one-hot, no gather)."""
    return vq_search(z.contiguous(), E, cb, flags=flags, stats=stats)


def pinned_empty(*shape, dtype=torch.float32, device=None) -> torch.Tensor:
    """Page-locked host tensor for the host-buffer entry points, allocated on the NUMA node of `device`.

    cudaHostAlloc backs the pages where the CALLING thread runs; on a multi-socket box a rank whose staging buffer
    sits on the far socket crosses the inter-socket link on every host->device copy.  The calling thread is bound
    to the GPU's ideal CPUs (NVML's nvmlDeviceSetCpuAffinity) for the duration of the allocation and the first
    touch, then its affinity mask is restored.  Falls back to a plain pinned allocation where NVML or the affinity
    call is unavailable (e.g. a cgroup that does not contain those CPUs)."""
    import os
    if len(shape) == 1 and isinstance(shape[0], (tuple, list, torch.Size)):
        shape = tuple(shape[0])
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    old = None
    try:
        old = os.sched_getaffinity(0)
        import pynvml
        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(dev).uuid)
        try:
            h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByUUID("GPU-" + uuid)
        pynvml.nvmlDeviceSetCpuAffinity(h)
    except Exception:
        pass
    try:
        t = torch.empty(*shape, dtype=dtype, pin_memory=True)
        t.zero_()                                   # first touch on the bound CPUs
    finally:
        if old is not None:
            try:
                os.sched_setaffinity(0, old)
            except Exception:
                pass
    return t


def tokenize_host(z_host: torch.Tensor, E: torch.Tensor, cb: Optional[torch.Tensor] = None, *,
                  chunk_rows: int = 131072, out: Optional[torch.Tensor] = None,
                  flags: int = _lib.ALGO_AUTO, return_stats: bool = False):
    """End-to-end tokenisation of HOST rows (pinned for full PCIe rate): chunked H2D copy,
    search and D2H of the ids overlapped inside the C library.  Returns a host int32 tensor."""
    _need_cuda(E, "codebook")
    if z_host.is_cuda or z_host.dtype not in _DT or not z_host.is_contiguous():
        raise RuntimeError("tokenize_host expects a contiguous CPU tensor of fp32/bf16/fp16 rows")
    lib = _lib.load()
    N, D = z_host.shape
    K = E.shape[0]
    if cb is None:
        cb = prepare_codebook(E)
    dt = _DT[z_host.dtype]
    chunk_rows = max(1, min(int(chunk_rows), max(N, 1)))
    if out is None:
        out = pinned_empty(N, dtype=torch.int32, device=E.device)
    stats = torch.zeros(8, dtype=torch.int64)
    torch.cuda.current_stream(E.device).synchronize()   # cb / E were produced on the caller's stream
    with torch.cuda.device(E.device):
        ws = _scratch.get(E.device, "tokhost", lib.g2v_tokenize_host_bytes(chunk_rows, K, D, dt, flags))
        _lib.check(lib.g2v_tokenize_host(C.c_void_p(z_host.data_ptr()), dt, N, _ptr(E), _ptr(cb), K, D,
                                         C.c_void_p(out.data_ptr()), chunk_rows,
                                         C.c_void_p(stats.data_ptr()), _ptr(ws), ws.numel(), flags),
                   "g2v_tokenize_host")
    return (out, stats) if return_stats else out
