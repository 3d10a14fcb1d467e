"""Tensor-level entry points over the C ABI: search, apply, statistics, EMA, backward, and the
autograd.Function the quantizer modules use.  PyTorch owns every buffer; the C side only
enqueues kernels on the current stream.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _lib

_DT = {torch.float32: _lib.F32, torch.bfloat16: _lib.BF16, torch.float16: _lib.F16}


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream(dev) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


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
    ws = _scratch.get(z.device, "search", lib.g2v_workspace_bytes(N, K, D, dt, flags))
    _lib.check(lib.g2v_vq_search(_ptr(z), dt, _ptr(E), _ptr(cb), N, K, D, _ptr(idx), _ptr(stats),
                                 _ptr(ws), ws.numel(), flags, _stream(z.device)), "g2v_vq_search")
    return idx


def packed_numel(K: int, D: int) -> int:
    return K * D + K + 2


def vq_apply(x: torch.Tensor, E: torch.Tensor, idx: torch.Tensor, *, zs: Optional[torch.Tensor] = None,
             want_out: bool = True, want_stats: bool = True, want_dwr: bool = False
             ) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
    """One pass over the rows: out = x + (E[idx]-x) and, if asked, the packed statistics buffer
    [dwr (K*D) | counts (K) | sse | rows] (fp32, ready for an all-reduce)."""
    lib = _lib.load()
    N, D = x.shape
    K = E.shape[0]
    dev = x.device
    st = _stream(dev)
    out = torch.empty_like(x) if want_out else None
    packed = None
    counts = sse = dwr = None
    reps = 0
    if want_stats:
        packed = torch.empty(packed_numel(K, D), dtype=torch.float32, device=dev)
        acc = torch.zeros(K * 4 + 8, dtype=torch.uint8, device=dev)   # int32 counts[K] + double sse
        sse = acc[:8].view(torch.float64)
        counts = acc[8:].view(torch.int32)
        if want_dwr:
            # private copies of the [K, D] sums keep hot codes from serialising the fp32 atomics;
            # large batches get up to 8 copies (bounded to 64 MB), small ones a single copy
            reps = 1 if N < 32768 else max(1, min(8, (64 << 20) // (K * D * 4)))
            dwr = torch.zeros(reps * K * D, dtype=torch.float32, device=dev)
    _lib.check(lib.g2v_vq_apply(_ptr(x), _ptr(zs), _ptr(E), _ptr(idx), N, K, D, _ptr(out), _ptr(sse),
                                _ptr(counts), _ptr(dwr), reps, st), "g2v_vq_apply")
    if want_stats:
        _lib.check(lib.g2v_vq_stats_pack(_ptr(counts), _ptr(sse), _ptr(dwr), reps, N, K, D, _ptr(packed), st),
                   "g2v_vq_stats_pack")
    return out, packed


def stats_finalize(packed: torch.Tensor, K: int, D: int, coef_codebook: float, coef_commit: float
                   ) -> Tuple[torch.Tensor, torch.Tensor]:
    lib = _lib.load()
    res = torch.empty(2, dtype=torch.float32, device=packed.device)
    _lib.check(lib.g2v_vq_stats_finalize(_ptr(packed), K, D, coef_codebook, coef_commit,
                                         C.c_void_p(res.data_ptr()), C.c_void_p(res.data_ptr() + 4),
                                         _stream(packed.device)), "g2v_vq_stats_finalize")
    return res[0], res[1]


def ema_update(cluster_size: torch.Tensor, ema_w: torch.Tensor, E_old: torch.Tensor, E_new: torch.Tensor,
               packed: torch.Tensor, decay: float, eps: float, cb: Optional[torch.Tensor]) -> None:
    lib = _lib.load()
    K, D = E_old.shape
    _lib.check(lib.g2v_vq_ema_update(_ptr(cluster_size), _ptr(ema_w), _ptr(E_old), _ptr(E_new), _ptr(packed),
                                     decay, eps, K, D, _ptr(cb), 0 if cb is None else cb.numel(),
                                     _stream(E_old.device)), "g2v_vq_ema_update")


def one_hot(idx: torch.Tensor, K: int) -> torch.Tensor:
    lib = _lib.load()
    N = idx.numel()
    enc = torch.empty(N, K, dtype=torch.float32, device=idx.device)
    _lib.check(lib.g2v_onehot(_ptr(idx), N, K, _ptr(enc), _stream(idx.device)), "g2v_onehot")
    return enc


# ------------------------------------------------------------------------------------------------
# autograd
# ------------------------------------------------------------------------------------------------
class _QuantizeFn(torch.autograd.Function):
    """(x, E) -> (loss, quantized, perplexity, idx, packed).

    Closed-form backward of the graph the reference builds:
      d/dx = g_quantized + g_loss * 2*beta/M * grad_scale * (x - E[idx])
      d/dE = -g_loss * 2/M * dwr                      (hard VQ only)
    """

    @staticmethod
    def forward(ctx, x2d, E, zs, cb, beta, coef_codebook, want_dwr, reduce_fn, grad_scale, flags):
        search_rows = x2d if zs is None else zs
        idx = vq_search(search_rows, E, cb, flags=flags)
        out, packed = vq_apply(x2d, E, idx, zs=zs, want_out=True, want_stats=True, want_dwr=want_dwr)
        if reduce_fn is not None:
            reduce_fn(packed)                      # data-parallel all-reduce (sum) of the statistics
        K, D = E.shape
        loss, ppl = stats_finalize(packed, K, D, coef_codebook, beta)
        ctx.save_for_backward(x2d, E, idx, packed)
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
        st = _stream(dev)
        M = float(N) * float(D)
        if g_loss is None:
            g_loss = torch.zeros((), dtype=torch.float32, device=dev)
        g_loss = g_loss.to(torch.float32).contiguous()
        gx = gE = None
        if ctx.needs_input_grad[0]:
            gx = torch.empty_like(x2d)
            if g_out is not None:
                g_out = g_out.contiguous()
            _lib.check(lib.g2v_vq_backward(_ptr(x2d), _ptr(E), _ptr(idx), _ptr(g_out), _ptr(g_loss),
                                           2.0 * ctx.beta * ctx.grad_scale / M, N, K, D, _ptr(gx), st),
                       "g2v_vq_backward")
        if ctx.needs_input_grad[1] and ctx.coef_codebook != 0.0:
            gE = torch.empty_like(E)
            _lib.check(lib.g2v_vq_grad_codebook(_ptr(packed), _ptr(g_loss),
                                                2.0 * ctx.coef_codebook * ctx.grad_scale / M, K, D, _ptr(gE), st),
                       "g2v_vq_grad_codebook")
        return gx, gE, None, None, None, None, None, None, None, None


def quantize(x2d, E, *, zs=None, cb=None, beta=0.25, coef_codebook=1.0, want_dwr=False,
             reduce_fn=None, grad_scale=1.0, flags=_lib.ALGO_AUTO):
    return _QuantizeFn.apply(x2d, E, zs, cb, float(beta), float(coef_codebook), bool(want_dwr),
                             reduce_fn, float(grad_scale), int(flags))


# ------------------------------------------------------------------------------------------------
# bulk tokenisation
# ------------------------------------------------------------------------------------------------
def tokenize(z: torch.Tensor, E: torch.Tensor, cb: Optional[torch.Tensor] = None, *,
             flags: int = _lib.ALGO_AUTO, stats: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Device-resident bulk tokenisation: [N, D] rows -> int32 code ids (no one-hot, no gather)."""
    return vq_search(z.contiguous(), E, cb, flags=flags, stats=stats)


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
        out = torch.empty(N, dtype=torch.int32, pin_memory=True)
    ws = _scratch.get(E.device, "tokhost", lib.g2v_tokenize_host_bytes(chunk_rows, K, D, dt, flags))
    stats = torch.zeros(8, dtype=torch.int64)
    torch.cuda.current_stream(E.device).synchronize()   # cb / E were produced on the caller's stream
    with torch.cuda.device(E.device):
        _lib.check(lib.g2v_tokenize_host(C.c_void_p(z_host.data_ptr()), dt, N, _ptr(E), _ptr(cb), K, D,
                                         C.c_void_p(out.data_ptr()), chunk_rows,
                                         C.c_void_p(stats.data_ptr()), _ptr(ws), ws.numel(), flags),
                   "g2v_tokenize_host")
    return (out, stats) if return_stats else out
