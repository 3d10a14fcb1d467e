"""The soft quantizer VQ_Payam_GSSoft (scripts/model/Autoencoder_VQVAE_model.py:1304-1433) on the B200 kernels.

Forward (all fp32 values, like the reference):
    m  = mean_layer(z)            g2v_gemm_f32 (split-fp16 tcgen05 GEMM, bias in the epilogue)
    lv = logvar_layer(m)          g2v_gemm_f32
    m E^T                         g2v_gemm_f32
    d, p                          g2v_soft_assign  (distance assembly + soft_prob + column sums of p, one pass)
    q  = p E                      g2v_gemm_f32
    out, loss, perplexity         g2v_soft_tail
Backward: the closed form derived and pinned in oracle/gssoft_oracle.py (the reference relies on autograd):
    dp = (q - x) E^T ; gd, glv = g2v_soft_backward ; gm = 2 m rowsum(gd) - 2 gd E + glv Wl
    dE = p^T (q - x) - 2 gd^T m + 2 E colsum(gd) ; dWl = glv^T m ; dWm = gm^T z ; gx = gm Wm + ...
every product again on g2v_gemm_f32 (the reductions over the N rows as transposed, split-K GEMMs with one fp16
term per operand), everything scaled by 2 g_loss / M at the end so that no intermediate leaves fp16's range.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from .functional import _need_cuda, _on, _ptr, _stream, gemm


class _SoftQuantizeFn(torch.autograd.Function):
    """(x [N,D], E [K,D], Wm, bm, Wl, bl) -> (loss, out, perplexity, p)."""

    @staticmethod
    def forward(ctx, x, E, Wm, bm, Wl, bl, beta):
        _need_cuda(x, "inputs")
        lib = _lib.load()
        dev = x.device
        N, D = x.shape
        K = E.shape[0]
        with _on(dev):
            st = _stream(dev)
            Ed = E.detach()
            m = gemm(x, Wm.detach(), bias=bm.detach())                       # mean_layer            (:1390)
            lv = gemm(m, Wl.detach(), bias=bl.detach())                      # logvar_layer          (:1391)
            d = gemm(m, Ed)                                                  # m E^T, becomes d      (:1393-1397)
            e2 = Ed.pow(2).sum(1)
            p = torch.empty(N, K, dtype=torch.float32, device=dev)
            colsum = torch.zeros(K, dtype=torch.float32, device=dev)
            _lib.check(lib.g2v_soft_assign(_ptr(m), _ptr(d), _ptr(lv), _ptr(e2), N, K, D, _ptr(p), _ptr(colsum), st),
                       "g2v_soft_assign")
            q = gemm(p, Ed, transB=True)                                     # encodings @ E         (:1407)
            out = torch.empty_like(x)
            sse = torch.zeros(1, dtype=torch.float64, device=dev)
            res = torch.empty(2, dtype=torch.float32, device=dev)
            _lib.check(lib.g2v_soft_tail(_ptr(x), _ptr(q), N, K, D, float(beta), _ptr(out), _ptr(sse), _ptr(colsum),
                                         C.c_void_p(res.data_ptr()), C.c_void_p(res.data_ptr() + 4), st), "g2v_soft_tail")
        ctx.save_for_backward(x, E, Wm, Wl, m, d, lv, p, q)
        ctx.beta = float(beta)
        ctx.set_materialize_grads(False)
        ctx.mark_non_differentiable(res[1], p)
        return res[0], out, res[1], p

    @staticmethod
    def backward(ctx, g_loss, g_out, _g_ppl, _g_p):
        x, E, Wm, Wl, m, d, lv, p, q = ctx.saved_tensors
        lib = _lib.load()
        dev = x.device
        N, D = x.shape
        K = E.shape[0]
        M = float(N) * float(D)
        need = ctx.needs_input_grad
        with _on(dev), torch.no_grad():
            st = _stream(dev)
            if g_loss is None:
                g_loss = torch.zeros((), dtype=torch.float32, device=dev)
            g_loss = g_loss.to(torch.float32).reshape(1)
            c = g_loss * (2.0 / M)                     # device scalar: every gradient below is scaled by it at the end
            a = g_loss * (2.0 * ctx.beta / M)
            Ed = E.detach()
            r = q - x                                  # gq / c
            dp = gemm(r, Ed)                           # [N, K]
            gd = torch.empty_like(dp)
            glv = torch.empty_like(dp)
            rs = torch.empty(N, dtype=torch.float32, device=dev)
            col = torch.zeros(2, K, dtype=torch.float32, device=dev)
            _lib.check(lib.g2v_soft_backward(_ptr(p), _ptr(dp), _ptr(d), _ptr(lv), N, K, _ptr(gd), _ptr(glv), _ptr(rs),
                                             _ptr(col[0]), _ptr(col[1]), st), "g2v_soft_backward")
            del dp
            gm = gemm(gd, Ed, transB=True, alpha=-2.0)                       # -2 gd E
            gemm(glv, Wl.detach(), transB=True, out=gm, accumulate=True)     # + glv Wl
            gm.addcmul_(m, rs.unsqueeze(1), value=2.0)                       # + 2 m rowsum(gd)
            gx = gE = gWm = gbm = gWl = gbl = None
            if need[0]:
                gmw = gemm(gm, Wm.detach(), transB=True)                     # gm Wm
                gx = torch.empty_like(x)
                go = g_out.contiguous() if g_out is not None else None
                _lib.check(lib.g2v_soft_gx(_ptr(gmw), _ptr(x), _ptr(q), _ptr(go), _ptr(c), _ptr(a), x.numel(), _ptr(gx), st),
                           "g2v_soft_gx")
            elif g_out is not None:
                pass
            long_k = N >= 65536                        # reductions over the rows: one fp16 term per operand is ample
            if need[1]:
                gE = gemm(p, r, transA=True, transB=True, fp16=long_k)       # p^T (q - x)
                gemm(gd, m, transA=True, transB=True, out=gE, accumulate=True, alpha=-2.0, fp16=long_k)
                gE.addcmul_(Ed, col[0].unsqueeze(1), value=2.0)              # + 2 E colsum(gd)
                gE.mul_(c)
            if need[2]:
                gWm = gemm(gm, x, transA=True, transB=True, fp16=long_k).mul_(c)
            if need[3]:
                gbm = gm.sum(0).mul_(c)
            if need[4]:
                gWl = gemm(glv, m, transA=True, transB=True, fp16=long_k).mul_(c)
            if need[5]:
                gbl = col[1].clone().mul_(c)
        return gx, gE, gWm, gbm, gWl, gbl, None


def soft_quantize(x2d: torch.Tensor, E: torch.Tensor, Wm: torch.Tensor, bm: torch.Tensor, Wl: torch.Tensor,
                  bl: torch.Tensor, beta: float):
    """The whole soft layer on [N, D] rows: (loss, out, perplexity, p [N, K])."""
    return _SoftQuantizeFn.apply(x2d, E, Wm, bm, Wl, bl, float(beta))
