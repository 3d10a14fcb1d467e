"""GPU audits of the search paths against exact arithmetic.

  * g2v_vq_search_exact (the fp64 device checker) against the numpy oracle -- pins the checker itself;
  * every search variant at the large codebooks of the K sweep (K = 4096 .. 16384 = the limit of the key's
    9-bit column-group field) against the numpy fp64 oracle, ragged N included;
  * every kernel variant over 131 072 rows per latent / codebook distribution against the device checker
    (the 1 M-row version of the same audit is tools/audit_exact.py; its log is kept under profiles/);
  * bf16 rows at K = 512 over 1 M rows with a 131 072-row audited subset (BASELINE configs[4]).
"""
import numpy as np
import pytest
import torch

import gpu_synth as S
from oracle import vq_oracle as O

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def _g():
    import gesture2vec_b200 as g
    from gesture2vec_b200 import _lib
    return g, _lib


def _variants(L):
    return {"auto": L.ALGO_AUTO, "simt": L.ALGO_SIMT, "tmem": L.ALGO_TC | L.TC_VARIANT_TMEM,
            "fused": L.ALGO_TC | L.TC_VARIANT_FUSED, "prep": L.ALGO_TC | L.TC_VARIANT_PREP}


@pytest.mark.parametrize("N,K,D,dtype", [(777, 400, 400, torch.float32), (130, 512, 400, torch.bfloat16),
                                         (300, 80, 45, torch.float32), (65, 1000, 41, torch.float16),
                                         (1, 3, 7, torch.float32)])
def test_exact_checker_matches_numpy_oracle(N, K, D, dtype):
    g, _ = _g()
    E = O.synth_codebook("normal", K, D, seed=5)
    z = torch.from_numpy(O.synth_latents("iid", N, D, seed=6)).to(dtype)
    got = g.vq_search_exact(z.to(DEV), torch.from_numpy(E).to(DEV)).cpu().numpy()
    zf = z.float().numpy()
    a = O.audit_indices(zf, E, got, O.nearest_code_f64(zf, E), eps_tie=2.0 ** -45)
    assert a["hard"] == 0, a


def test_exact_checker_first_index_on_ties():
    g, _ = _g()
    E = np.tile(O.synth_codebook("normal", 8, 24, seed=1), (5, 1))           # every code appears 5 times
    z = O.synth_latents("gru", 200, 24, seed=2)
    got = g.vq_search_exact(torch.from_numpy(z).to(DEV), torch.from_numpy(E).to(DEV)).cpu().numpy()
    assert np.array_equal(got, O.nearest_code_f64(z, E)) and got.max() < 8


@pytest.mark.parametrize("K", [4096, 8192, 16384])
@pytest.mark.parametrize("N", [2048, 2301])
def test_search_large_codebooks_against_oracle(K, N):
    """The top of the K sweep, where the tensor-bound claim lives: both tcgen05 variants that cover the shape
    and the fp32 path, against the numpy fp64 argmin; indices stay inside [0, K)."""
    g, L = _g()
    D = 400
    E = O.synth_codebook("normal", K, D, seed=K)
    z = O.synth_latents("iid", N, D, seed=N)
    ref = O.nearest_code_f64(z, E)
    zt, Et = torch.from_numpy(z).to(DEV), torch.from_numpy(E).to(DEV)
    cb = g.prepare_codebook(Et)
    for name in ("auto", "prep", "simt"):
        stats = torch.zeros(8, dtype=torch.int64, device=DEV)
        idx = g.vq_search(zt, Et, cb, flags=_variants(L)[name], stats=stats).cpu().numpy()
        a = O.audit_indices(z, E, idx, ref, eps_tie=2.0 ** -40)
        assert a["hard"] == 0, (name, a, stats.cpu().numpy())
        assert idx.min() >= 0 and idx.max() < K


def test_search_large_codebook_bf16_rows():
    g, L = _g()
    K, D, N = 16384, 400, 1500
    E = O.synth_codebook("uniform1", K, D, seed=3)
    z16 = torch.from_numpy(O.synth_latents("gru", N, D, seed=4)).to(torch.bfloat16)
    zf = z16.float().numpy()
    idx = g.vq_search(z16.to(DEV), torch.from_numpy(E).to(DEV)).cpu().numpy()
    assert O.audit_indices(zf, E, idx, O.nearest_code_f64(zf, E), eps_tie=2.0 ** -40)["hard"] == 0


AUDIT_CASES = [  # latents, codebook, K
    ("iid", "normal", 400),
    ("gru", "uniform1", 512),
    ("clustered", "normal", 400),
    ("gru", "ema_degenerate", 512),
]


@pytest.mark.parametrize("lk,ck,K", AUDIT_CASES)
def test_full_row_audit_every_variant(lk, ck, K):
    """131 072 rows, EVERY row compared with the fp64 device checker, for each kernel variant."""
    g, L = _g()
    D, N = 400, 131072
    E = S.codebook(ck, K, D, DEV, seed=7)
    z = S.latents(lk, N, D, DEV, E=E, seed=8)
    exact = g.vq_search_exact(z, E)
    cb = g.prepare_codebook(E)
    for name, flags in _variants(L).items():
        idx = g.vq_search(z, E, cb, flags=flags)
        a = S.audit(z, E, idx, exact, eps_tie=2.0 ** -40)
        assert a["hard"] == 0, (name, a)
    for dtype in (torch.bfloat16, torch.float16):                      # 16-bit rows: their own exact answer
        z16 = z[:32768].to(dtype).contiguous()
        ex16 = g.vq_search_exact(z16, E)
        for name in ("auto", "prep", "simt"):
            idx = g.vq_search(z16, E, cb, flags=_variants(L)[name])
            a = S.audit(z16.float(), E, idx, ex16, eps_tie=2.0 ** -40)
            assert a["hard"] == 0, (name, str(dtype), a)


def test_bf16_million_rows_audited_subset():
    """BASELINE configs[4] shape: bf16 rows, K = 512; 1 M rows tokenised, the first 131 072 audited against the
    fp64 checker on the same bf16 values, and against the fp32 rows they were rounded from (near-tie count of the
    bf16 INPUT rounding, reported, not asserted to be zero: that is the input's precision, not the kernel's)."""
    g, L = _g()
    K, D, N, NA = 512, 400, 1_000_000, 131072
    E = S.codebook("normal", K, D, DEV, seed=0)
    z32 = S.latents("iid", N, D, DEV, seed=1234)
    z16 = z32.to(torch.bfloat16)
    idx = g.tokenize(z16, E)
    assert int(idx.min()) >= 0 and int(idx.max()) < K
    ex16 = g.vq_search_exact(z16[:NA].contiguous(), E)
    a = S.audit(z16[:NA].float(), E, idx[:NA], ex16, eps_tie=2.0 ** -40)
    assert a["hard"] == 0, a
    ex32 = g.vq_search_exact(z32[:NA].contiguous(), E)
    flips = int((ex32 != ex16).sum())
    assert flips < NA // 50, flips                                      # ~0.4 % expected from bf16 inputs


def test_unaligned_and_odd_shapes_hypothesis():
    """Ragged N, K, D (incl. the frame-level D = 40 / 41 / 45) and base pointers that are only 4-byte aligned."""
    from hypothesis import given, settings, strategies as st, HealthCheck
    g, L = _g()

    @settings(max_examples=40, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)
    @given(N=st.integers(1, 700), K=st.integers(1, 700), D=st.sampled_from([7, 40, 41, 45, 64, 100, 200, 400]),
           off=st.integers(0, 3), dt=st.sampled_from(["f32", "bf16"]), seed=st.integers(0, 10 ** 6))
    def run(N, K, D, off, dt, seed):
        E = O.synth_codebook("normal", K, D, seed=seed)
        z = O.synth_latents("iid", N, D, seed=seed + 1)
        tdt = torch.float32 if dt == "f32" else torch.bfloat16
        buf = torch.zeros(N * D + 8, dtype=tdt, device=DEV)
        zt = buf[off:off + N * D].view(N, D)
        zt.copy_(torch.from_numpy(z).to(tdt))
        zf = zt.float().cpu().numpy()
        ref = O.nearest_code_f64(zf, E)
        Et = torch.from_numpy(E).to(DEV)
        for flags in (L.ALGO_AUTO, L.ALGO_SIMT):
            idx = g.vq_search(zt, Et, flags=flags).cpu().numpy()
            a = O.audit_indices(zf, E, idx, ref, eps_tie=2.0 ** -40)
            assert a["hard"] == 0, (N, K, D, off, dt, flags, a)

    run()


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("N,K,D,extra", [(70000, 512, 400, 4), (300, 400, 400, 4), (1000, 1024, 200, 16), (100, 64, 40, 4)])
def test_narrow_rows_against_a_wider_codebook(N, K, D, extra, dtype):
    """g2v_vq_search_wide: rows [N, D] searched against a [K, D + extra] codebook as [z | 0] -- the folded-projection
    search.  Every row against the fp64 checker on explicitly widened rows; shapes the TMA sweep does not cover (tiny N,
    tiny D) take the widened-copy fallback and must agree too."""
    g, L = _g()
    gen = torch.Generator(device=DEV).manual_seed(N + K)
    E = torch.randn(K, D + extra, device=DEV, generator=gen)
    E[:, D + 1:] = 0.0                                     # the fold uses ONE extra column; the rest is alignment padding
    z = torch.randn(N, D, device=DEV, generator=gen).to(dtype)
    cb = g.prepare_codebook(E)
    stats = torch.zeros(8, dtype=torch.int64, device=DEV)
    idx = g.vq_search_wide(z, E, cb, stats=stats)
    zw = torch.nn.functional.pad(z.float(), (0, extra)).contiguous()
    exact = g.vq_search_exact(zw, E)
    a = S.audit(zw, E, idx, exact, eps_tie=2.0 ** -40)
    assert a["hard"] == 0, (a, stats.cpu().tolist())


def test_dead_codes_do_not_push_live_rows_into_the_exact_fallback():
    """A codebook whose largest entries sit 2^22 above the codes that win rows (dead EMA codes, or the norm
    coordinate of a folded codebook): the fp16 copy is scaled to the top of the fp16 range and its residual is
    split into a relative and an absolute part, so live rows still certify in the tensor-core pass."""
    g, L = _g()
    K, D, N = 512, 400, 131072
    E = S.codebook("uniform1", K, D, DEV, seed=8) * 0.2
    E[::23] *= 4.0e6                                            # 23 codes nobody can reach
    E[7] *= 1e-3                                                # and one with a tiny norm
    z = S.latents("gru", N, D, DEV, seed=9) * 0.3
    want = g.vq_search_exact(z, E)
    cb = g.prepare_codebook(E)
    for name in ("auto", "prep", "simt"):
        stats = torch.zeros(8, dtype=torch.int64, device=DEV)
        idx = g.vq_search(z, E, cb, flags=_variants(L)[name], stats=stats)
        a = S.audit(z, E, idx, want, eps_tie=2.0 ** -40)
        assert a["hard"] == 0, (name, a)
        st = stats.cpu().numpy()
        if name != "simt":
            assert st[L.STAT_FALLBACK_ROWS] < N // 10, (name, st)


@pytest.mark.parametrize("cbk,lat,K", [("normal", "iid", 400), ("ema_degenerate", "gru", 512), ("uniform1", "clustered", 1000)])
def test_refine_pass_gives_the_same_indices_as_straight_fp64(cbk, lat, K):
    """Whole-row re-ranks take a second, fp32-accurate tensor-core pass (split-fp16 operands) before fp64 touches
    them; G2V_NO_REFINE sends them straight to fp64.  Both are exact, so the indices are identical."""
    g, L = _g()
    D, N = 400, 131072
    E = S.codebook(cbk, K, D, DEV, seed=21)
    z = S.latents(lat, N, D, DEV, E=E, seed=22)
    cb = g.prepare_codebook(E)
    on, off = (torch.zeros(8, dtype=torch.int64, device=DEV) for _ in range(2))
    a = g.vq_search(z, E, cb, stats=on)
    b = g.vq_search(z, E, cb, flags=L.NO_REFINE, stats=off)
    assert torch.equal(a, b)
    on, off = on.cpu().numpy(), off.cpu().numpy()
    # the pass is taken when rows x codes is worth three launches (else rerank_kernel's own fp64 path covers them)
    assert on[L.STAT_FALLBACK_ROWS] == off[L.STAT_FALLBACK_ROWS] and off[L.STAT_REFINE_ROWS] == 0
    if on[L.STAT_FALLBACK_ROWS] * K >= 131072:
        assert on[L.STAT_REFINE_ROWS] == on[L.STAT_FALLBACK_ROWS] and on[L.STAT_FULL_RECHECK] == 0
        assert on[L.STAT_REFINE_EXACT] >= on[L.STAT_REFINE_ROWS]
    else:
        assert on[L.STAT_REFINE_ROWS] == 0 and on[L.STAT_FULL_RECHECK] == on[L.STAT_FALLBACK_ROWS]
    want = g.vq_search_exact(z, E)
    assert S.audit(z, E, a, want, eps_tie=2.0 ** -40)["hard"] == 0


def test_refine_pass_keeps_the_first_index_on_exact_ties():
    g, L = _g()
    D, N = 64, 65536
    E = S.codebook("normal", 32, D, DEV, seed=1).repeat(8, 1)              # every code 8 times: all rows tie
    z = S.latents("gru", N, D, DEV, seed=2)
    st = torch.zeros(8, dtype=torch.int64, device=DEV)
    idx = g.vq_search(z, E, stats=st)
    assert int(idx.max()) < 32
    assert torch.equal(idx, g.vq_search_exact(z, E))


@pytest.mark.parametrize("N,K0,rep,D,dtype", [(40000, 75, 4, 60, torch.float32), (33001, 40, 4, 404, torch.float32),
                                               (50003, 48, 4, 180, torch.bfloat16), (32768, 80, 4, 496, torch.float16)])
def test_refine_pass_ragged_shapes(N, K0, rep, D, dtype):
    """The refine pass at shapes off its fast paths: D not a multiple of 64 (zero-padded operand panels), K not a
    multiple of the GEMM tile, ragged N, 16-bit rows.  Every code appears four times, spread over the 32 column
    chains so that a row's best distance ties in four chains (or twice in two): such rows are listed whole."""
    g, L = _g()
    E = S.codebook("normal", K0, D, DEV, seed=31).repeat(rep, 1).contiguous()
    z = S.latents("gru", N, D, DEV, seed=32).to(dtype)
    st = torch.zeros(8, dtype=torch.int64, device=DEV)
    idx = g.vq_search(z, E, stats=st)
    st = st.cpu().numpy()
    assert st[L.STAT_REFINE_ROWS] > N // 2, st
    want = g.vq_search_exact(z, E)
    assert torch.equal(idx, want) and int(idx.max()) < K0           # exact ties: the first copy wins


@pytest.mark.parametrize("N,K,D", [(16384 + 17, 70, 52), (20000, 300, 404), (16385, 512, 400), (40001, 33, 8)])
def test_bulk_row_pass_and_backward_ragged_shapes(N, K, D):
    """The bulk kernels of the training step (run-aggregated row pass behind its cp.async ring, N >= 16384; the
    backward as one float4 stream) at ragged N and D off the 128-column grid, against torch fp64 on the same GPU."""
    g, L = _g()
    E = S.codebook("normal", K, D, DEV, seed=41)
    x = S.latents("clustered", N, D, DEV, E=E, seed=42)
    idx = g.vq_search(x, E)
    out, packed = g.vq_apply(x, E, idx, want_out=True, want_stats=True, want_dwr=True)
    lay = g.packed_layout(K, D)
    q = E[idx.long()]
    assert torch.equal(out, x + (q - x))
    counts = packed[lay["counts"][0]:lay["counts"][1]]
    assert torch.equal(counts.long(), torch.bincount(idx.long(), minlength=K))
    sse = ((q.double() - x.double()) ** 2).sum()
    np.testing.assert_allclose(float(packed[lay["sse"]]), float(sse), rtol=2e-5)
    dwr = packed[:K * D].view(K, D).double()
    ref = torch.zeros(K, D, device=DEV, dtype=torch.float64).index_add_(0, idx.long(), x.double() - q.double())
    tol = 2e-5 * ref.abs().amax(dim=1, keepdim=True) + 1e-4
    assert bool(((dwr - ref).abs() <= tol).all()), float((dwr - ref).abs().max())
    # backward: g_x = g_out + c (x - E[idx]) with c = g_loss * coef
    layer = g.DAE_VQ_Payam(K, D, 0.25).to(DEV)
    with torch.no_grad():
        layer._embedding.weight.copy_(E)
    xs = x.clone().requires_grad_(True)
    gq = torch.randn(N, D, device=DEV, generator=torch.Generator(device=DEV).manual_seed(5))
    loss, qq, ppl, enc = layer(xs)
    torch.autograd.backward([loss, qq], [torch.tensor(2.0, device=DEV), gq])
    c = 2.0 * 0.25 * 2.0 / (N * D)                                   # d(beta * mse(x, q.detach())) / dx, times g_loss
    want = gq.double() + c * (x.double() - q.double())
    assert bool(((xs.grad.double() - want).abs() <= 1e-6 * want.abs() + 1e-7).all())


def test_six_million_rows_need_64_bit_offsets():
    """N x D = 2.4e9 elements (9.6 GB of fp32 rows): row * D no longer fits 32 bits for the last 630 000 rows.
    The search, the row pass and the backward are checked on blocks from both ends against the same kernels run on
    the blocks alone (they are exact, so the indices agree bit for bit) and against the fp64 checker."""
    g, L = _g()
    K, D, N = 400, 400, 6_000_000
    E = S.codebook("normal", K, D, DEV, seed=51)
    z = torch.empty(N, D, device=DEV)
    gen = torch.Generator(device=DEV).manual_seed(52)
    for i in range(0, N, 1_000_000):
        z[i:i + 1_000_000].normal_(generator=gen)
    cb = g.prepare_codebook(E)
    idx = g.vq_search(z, E, cb)
    assert int(idx.min()) >= 0 and int(idx.max()) < K
    for lo in (0, N - 70_000):
        blk = z[lo:lo + 70_000]
        assert torch.equal(idx[lo:lo + 70_000], g.vq_search(blk.contiguous(), E, cb))
        a = S.audit(blk, E, idx[lo:lo + 70_000], g.vq_search_exact(blk.contiguous(), E), eps_tie=2.0 ** -40)
        assert a["hard"] == 0, a
    out, packed = g.vq_apply(z, E, idx, want_out=True, want_stats=True, want_dwr=True)
    lay = g.packed_layout(K, D)
    assert float(packed[lay["rows"]]) == N
    tail = slice(N - 50_000, N)
    q = E[idx[tail].long()]
    assert torch.equal(out[tail], z[tail] + (q - z[tail]))
    del out
    gx = torch.empty_like(z)
    one = torch.ones((), device=DEV)
    from gesture2vec_b200 import _lib as LL
    lib = LL.load()
    LL.check(lib.g2v_vq_backward(z.data_ptr(), E.data_ptr(), idx.data_ptr(), None, one.data_ptr(), 0.5, N, K, D,
                                 gx.data_ptr(), torch.cuda.current_stream().cuda_stream), "g2v_vq_backward")
    assert torch.allclose(gx[tail], 0.5 * (z[tail] - q), rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("lk,ck,K", [("iid", "normal", 400), ("gru", "uniform1", 512), ("clustered", "normal", 1000),
                                     ("iid", "ema_degenerate", 512)])
def test_refine_pass_decides_every_row_exactly(lk, ck, K):
    """G2V_LIST_ALL_ROWS hands EVERY row to the refine pass (split-fp16 tensor-core dots + a-priori error bound +
    fp64 on the codes the bound cannot exclude): 131 072 rows per distribution against the fp64 checker.  A bound that
    was too tight would show up as wrong indices here; one that was loose as many fp64 evaluations per row."""
    g, L = _g()
    D, N = 400, 131072
    E = S.codebook(ck, K, D, DEV, seed=61)
    z = S.latents(lk, N, D, DEV, E=E, seed=62)
    st = torch.zeros(8, dtype=torch.int64, device=DEV)
    idx = g.vq_search(z, E, flags=L.ALGO_TC | L.LIST_ALL_ROWS, stats=st)
    st = st.cpu().numpy()
    assert st[L.STAT_REFINE_ROWS] == N, st
    a = S.audit(z, E, idx, g.vq_search_exact(z, E), eps_tie=2.0 ** -40)
    assert a["hard"] == 0, a
    assert st[L.STAT_REFINE_EXACT] < 1.5 * N, st                 # ~1 fp64 evaluation per row: the bound is tight
