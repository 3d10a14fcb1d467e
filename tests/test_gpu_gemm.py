"""The split-fp16 tcgen05 GEMM (g2v_gemm_f32) against fp64 matmul: every operand layout the quantizer modules use
(nn.Linear weights, transposed reductions over the rows, split-K with atomics, accumulate, bias, ragged sizes)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def _ref(A, B, tA, tB, bias, alpha):
    a = (A.t() if tA else A).double()
    b = (B.t() if tB else B).double()
    c = alpha * (a @ b.t())
    return c + (bias.double() if bias is not None else 0.0)


def _err(C, R, A, B, tA, tB):
    # error relative to |a_row| |b_row| (what a dot product's rounding scales with)
    a = (A.t() if tA else A).double().norm(dim=1, keepdim=True)
    b = (B.t() if tB else B).double().norm(dim=1, keepdim=True).t()
    return float(((C.double() - R).abs() / (a * b + 1e-30)).max())


@pytest.mark.parametrize("M,N,K,tA,tB,use_bias", [
    (1000, 400, 400, False, False, True),        # pre_linear / mean_layer: rows x nn.Linear weight
    (300, 512, 400, False, False, True),         # logvar_layer
    (129, 400, 512, False, True, False),         # p @ E with E stored [K, D]
    (7, 16, 5, False, False, True),              # tiny, everything ragged
    (512, 400, 20000, True, True, False),        # weight gradient: reduction over the rows, split-K + atomics
    (400, 400, 5000, True, True, False),
    (2048, 250, 333, False, False, False),       # N, K not multiples of anything
])
def test_gemm_matches_fp64(M, N, K, tA, tB, use_bias):
    import gesture2vec_b200 as g
    gen = torch.Generator(device=DEV).manual_seed(M + N + K)
    A = torch.randn((K, M) if tA else (M, K), device=DEV, generator=gen)
    B = torch.randn((K, N) if tB else (N, K), device=DEV, generator=gen) * 0.05
    bias = torch.randn(N, device=DEV, generator=gen) if use_bias else None
    C = g.functional.gemm(A, B, bias=bias, transA=tA, transB=tB, alpha=0.5)
    R = _ref(A, B, tA, tB, bias, 0.5)
    assert _err(C, R, A, B, tA, tB) < 2e-6, _err(C, R, A, B, tA, tB)
    # fp32 SGEMM for scale: the split GEMM must be at least as accurate as cuBLAS fp32 up to a small factor
    torch.backends.cuda.matmul.allow_tf32 = False
    S = 0.5 * ((A.t() if tA else A) @ (B.t() if tB else B).t()) + (bias if bias is not None else 0.0)
    assert _err(C, R, A, B, tA, tB) <= 8 * _err(S, R, A, B, tA, tB) + 1e-7


def test_gemm_accumulate_strided_and_single_term():
    import gesture2vec_b200 as g
    gen = torch.Generator(device=DEV).manual_seed(5)
    big = torch.randn(600, 900, device=DEV, generator=gen)
    A = big[:, 100:500]                               # rows strided by 900
    B = torch.randn(384, 400, device=DEV, generator=gen)
    out = torch.ones(600, 384, device=DEV)
    g.functional.gemm(A, B, out=out, accumulate=True)
    R = 1.0 + A.double() @ B.double().t()
    assert float((out.double() - R).abs().max()) < 1e-3 * float(R.abs().max()) * 1e-2
    # one fp16 term per operand: 2^-11 per product, unbiased
    C1 = g.functional.gemm(A, B, fp16=True)
    rel = float((C1.double() - (R - 1.0)).norm() / (R - 1.0).norm())
    assert rel < 1e-3, rel
    # extreme scales: tiny gradients and huge dead codes survive the power-of-two scaling
    tiny = torch.randn(300, 400, device=DEV, generator=gen) * 1e-9
    huge = torch.randn(64, 400, device=DEV, generator=gen) * 3e5
    C = g.functional.gemm(tiny, huge)
    R = tiny.double() @ huge.double().t()
    assert float((C.double() - R).abs().max()) < 1e-5 * float(R.abs().max())
