"""GPU parity tests: the CUDA path (through the C ABI) against the oracle and the golden vectors
produced by the real reference.  Run on the B200 box: pytest -m gpu."""
import os

import numpy as np
import pytest
import torch

from cases import CASE_NAMES, G_LOSS, load_case
from oracle import vq_oracle as O

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))


def _dev():
    return torch.device("cuda:0")


def _algos():
    import gesture2vec_b200 as g
    from gesture2vec_b200 import _lib
    lib = _lib.load()
    return g, _lib, lib


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(_dev())


def _rows_close(full, g, key, rows, rtol, atol):
    a2 = full.detach().float().cpu().numpy()
    a2 = a2.reshape(-1, a2.shape[-1])
    np.testing.assert_allclose(a2[rows], g[key + "_rows"], rtol=rtol, atol=atol)
    s, a = float(g[key + "_sum"]), float(g[key + "_abssum"])
    assert abs(float(a2.astype(np.float64).sum()) - s) <= rtol * a + atol * a2.size


def _build_layer(g, r):
    cls = g.FLAVOURS[r["flavour"]][r["cls"]]
    layer = cls(r["K"], r["D"], r["beta"], r["decay"], r["eps"]) if r["ema"] else cls(r["K"], r["D"], r["beta"])
    with torch.no_grad():
        layer._embedding.weight.copy_(torch.from_numpy(r["E0"]))
        if r["ema"]:
            layer._ema_w.copy_(torch.from_numpy(r["ema_w0"]))
        if hasattr(layer, "pre_linear"):
            layer.pre_linear.weight.copy_(torch.from_numpy(r["pre_W"]))
            layer.pre_linear.bias.copy_(torch.from_numpy(r["pre_b"]))
    return layer.to(_dev())


# ---------------------------------------------------------------------------------------------
# search: exact-arithmetic argmin
# ---------------------------------------------------------------------------------------------
SEARCH_SHAPES = [  # N, K, D, latents, codebook
    (1000, 512, 400, "iid", "normal"),
    (257, 400, 400, "gru", "uniform1"),
    (4096, 400, 400, "clustered", "normal"),
    (300, 80, 40, "gru", "uniform_invK"),
    (250, 64, 45, "iid", "uniform1"),
    (1, 400, 400, "gru", "uniform1"),
    (129, 1000, 400, "iid", "normal"),
    (640, 2048, 400, "iid", "normal"),
    (77, 3, 7, "iid", "normal"),
]


@pytest.mark.parametrize("algo", ["simt", "auto"])
@pytest.mark.parametrize("N,K,D,lk,ck", SEARCH_SHAPES)
def test_search_is_exact_argmin(N, K, D, lk, ck, algo):
    g, L, _ = _algos()
    E = O.synth_codebook(ck, K, D, seed=3)
    z = O.synth_latents(lk, N, D, E=E, seed=11)
    stats = torch.zeros(8, dtype=torch.int64, device=_dev())
    flags = L.ALGO_SIMT if algo == "simt" else L.ALGO_AUTO
    idx = g.vq_search(_t(z), _t(E), flags=flags, stats=stats).cpu().numpy()
    ref64 = O.nearest_code_f64(z, E)
    ref32 = O.nearest_code_f32(z, E)
    a64 = O.audit_indices(z, E, idx, ref64, eps_tie=2.0 ** -40)   # vs exact arithmetic: essentially equal
    a32 = O.audit_indices(z, E, idx, ref32)                        # vs the reference's fp32 formula
    assert a64["hard"] == 0, (a64, stats.cpu().numpy())
    assert a32["hard"] == 0, a32
    assert idx.min() >= 0 and idx.max() < K


def test_search_exact_ties_first_index():
    g, L, _ = _algos()
    r, gold = load_case("dae_hard_dupcodes")
    z = O.flatten_rows(r["x0"], r["D"])
    for flags in (L.ALGO_SIMT, L.ALGO_AUTO):
        idx = g.vq_search(_t(z), _t(r["E0"]), flags=flags).cpu().numpy()
        assert np.array_equal(idx, gold["s0_idx"].astype(np.int32))
    # all-zero rows against a codebook with equal-norm duplicates
    E = np.tile(O.synth_codebook("normal", 4, 16, seed=5), (8, 1))
    idx = g.vq_search(_t(np.zeros((33, 16), np.float32)), _t(E)).cpu().numpy()
    assert np.array_equal(idx, O.nearest_code_f64(np.zeros((33, 16), np.float32), E))


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_search_16bit_latents(dtype):
    """16-bit rows are searched as the fp32 values they hold (the codebook stays fp32)."""
    g, L, _ = _algos()
    K, D, N = 512, 400, 3000
    E = O.synth_codebook("uniform1", K, D, seed=9)
    z16 = torch.from_numpy(O.synth_latents("gru", N, D, seed=21)).to(dtype)
    zf = z16.float().numpy()
    for flags in (L.ALGO_SIMT, L.ALGO_AUTO):
        idx = g.vq_search(z16.to(_dev()), _t(E), flags=flags).cpu().numpy()
        assert O.audit_indices(zf, E, idx, O.nearest_code_f64(zf, E), eps_tie=2.0 ** -40)["hard"] == 0


def test_degenerate_ema_codebook():
    """After one reference EMA step unused codes blow up to |E| ~ 1e5: still exact."""
    g, _, _ = _algos()
    K, D, N = 512, 400, 2048
    E0 = O.synth_codebook("uniform1", K, D, seed=1)
    z = O.synth_latents("gru", N, D, seed=2)
    layer = O.EmaVQ(E0, np.random.default_rng(3).standard_normal((K, D), dtype=np.float32), 0.25, 0.85)
    layer.forward(z[:128])
    E1 = layer.E
    assert np.abs(E1).max() > 1e3
    from gesture2vec_b200 import _lib
    for flags in (_lib.ALGO_AUTO, _lib.ALGO_SIMT):
        stats = torch.zeros(8, dtype=torch.int64, device=_dev())
        idx = g.vq_search(_t(z), _t(E1), flags=flags, stats=stats).cpu().numpy()
        assert O.audit_indices(z, E1, idx, O.nearest_code_f64(z, E1))["hard"] == 0
        # dead codes with huge norms must not push ordinary rows onto the slow exact paths: the error
        # bounds use the largest code norm that can still win a row, not the global maximum
        st = stats.cpu().numpy()
        assert st[_lib.STAT_FALLBACK_ROWS] + st[_lib.STAT_FULL_RECHECK] <= N // 10, st


# ---------------------------------------------------------------------------------------------
# golden vectors of the real reference, through the drop-in modules
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", CASE_NAMES)
def test_modules_match_reference_golden(name):
    g, _, _ = _algos()
    r, gold = load_case(name)
    layer = _build_layer(g, r)
    layer.train()
    near_ties = 0
    for s in range(r["steps"]):
        x = _t(r[f"x{s}"]).requires_grad_(True)
        gout = _t(r[f"g{s}"])
        E_used = layer._embedding.weight.detach().cpu().numpy().copy()
        zs = O.flatten_rows(r[f"x{s}"], r["D"])
        if r["ema"] and r["flavour"] == "vqvae":
            zs = (zs @ r["pre_W"].T + r["pre_b"]).astype(np.float32)
        ref_idx = gold[f"s{s}_idx"].astype(np.int64)
        # the search alone first (no state change): a near-tie against the reference must not cascade into the
        # float comparisons, so the step itself then runs at the REFERENCE indices (as test_oracle_golden.py does)
        probe = layer.tokenize(x.detach()).cpu().numpy()
        aud = O.audit_indices(zs, E_used, probe, ref_idx)
        assert aud["hard"] == 0, aud
        near_ties += aud["mismatch"]
        if aud["mismatch"]:
            loss, quant, ppl, enc = layer.forward_with_indices(x, _t(ref_idx.astype(np.int32)))
        else:
            loss, quant, ppl, enc = layer(x)
        assert quant.shape == x.shape and quant.is_contiguous()
        assert loss.dim() == 0 and ppl.dim() == 0
        N = x.numel() // r["D"]
        assert tuple(enc.shape) == (N, r["K"]) and enc.dtype == torch.float32
        idx = torch.argmax(enc, 1).cpu().numpy()
        assert float(enc.sum()) == N and np.array_equal(idx, layer.last_indices.cpu().numpy())
        assert np.array_equal(idx, ref_idx if aud["mismatch"] else probe)
        (loss * G_LOSS + (quant * gout).sum()).backward()
        np.testing.assert_allclose(loss.item(), gold[f"s{s}_loss"], rtol=1e-5)
        np.testing.assert_allclose(ppl.item(), gold[f"s{s}_ppl"], rtol=1e-5)
        qtol = (1e-6, 1e-7) if s == 0 else (3e-5, 1e-6)
        _rows_close(quant, gold, f"s{s}_quant", gold["nrows"], *qtol)
        _rows_close(x.grad, gold, f"s{s}_gx", gold["nrows"], 1e-5, 1e-6 * float(np.abs(gold[f's{s}_gx_rows']).max()))
        if not r["ema"]:
            _rows_close(layer._embedding.weight.grad, gold, f"s{s}_gE", gold["krows"], 1e-4, 1e-7)
            layer._embedding.weight.grad = None
        else:
            assert layer._embedding.weight.grad is None
            assert layer.pre_linear.weight.grad is None
            np.testing.assert_allclose(layer._ema_cluster_size.cpu().numpy(), gold[f"s{s}_cs"], rtol=1e-5, atol=1e-9)
            # the vqvae flavour takes its EMA sums over pre_linear(x): the reference adds the fp32-rounded
            # projections row by row, the fold projects the per-code sums (W sum(x) + count b) -- two orders of
            # the same fp32 sum, a few 1e-6 apart in absolute terms on elements that cancel to ~1e-2
            folded = r["flavour"] == "vqvae"
            _rows_close(layer._ema_w, gold, f"s{s}_ema_w", gold["krows"], 1e-5, 4e-6 if folded else 1e-6)
            _rows_close(layer._embedding.weight, gold, f"s{s}_E", gold["krows"], 2e-5, 4e-6 if folded else 1e-6)
    if r["ema"]:
        layer.eval()
        before = [t.detach().clone() for t in (layer._embedding.weight, layer._ema_w, layer._ema_cluster_size)]
        xe = _t(r[f"x{r['steps']}"])
        zs = O.flatten_rows(r[f"x{r['steps']}"], r["D"])
        if r["flavour"] == "vqvae":
            zs = (zs @ r["pre_W"].T + r["pre_b"]).astype(np.float32)
        ref_idx = gold["eval_idx"].astype(np.int64)
        aud = O.audit_indices(zs, before[0].cpu().numpy(), layer.tokenize(xe).cpu().numpy(), ref_idx)
        assert aud["hard"] == 0
        with torch.no_grad():
            loss, quant, ppl, enc = (layer.forward_with_indices(xe, _t(ref_idx.astype(np.int32))) if aud["mismatch"]
                                     else layer(xe))
        after = (layer._embedding.weight, layer._ema_w, layer._ema_cluster_size)
        assert all(torch.equal(a, b) for a, b in zip(before, after))     # eval leaves EMA state bit-unchanged
        np.testing.assert_allclose(loss.item(), gold["eval_loss"], rtol=2e-5)
        np.testing.assert_allclose(ppl.item(), gold["eval_ppl"], rtol=1e-5)


def test_state_dict_contract_and_checkpoint_roundtrip():
    import json
    g, _, _ = _algos()
    keys = json.load(open(os.path.join(HERE, "golden", "state_dict_keys.json")))
    for flavour, classes in g.FLAVOURS.items():
        for cname, cls in classes.items():
            layer = cls(8, 4, 0.25) if cname in ("VQ_Payam", "VQ_Payam_GSSoft") else cls(8, 4, 0.25, 0.9)
            assert sorted(layer.state_dict().keys()) == keys[f"{flavour}.{cname}"]
    a = g.VQVAE_VQ_Payam_EMA(32, 16, 0.25, 0.85).to(_dev())
    a.train()
    a(torch.randn(2, 20, 8, device=_dev()))
    b = g.VQVAE_VQ_Payam_EMA(32, 16, 0.25, 0.85)
    b.load_state_dict({k: v.cpu() for k, v in a.state_dict().items()}, strict=True)
    b = b.to(_dev()).eval()
    a.eval()
    x = torch.randn(2, 10, 8, device=_dev())
    ra, rb = a(x), b(x)
    assert all(torch.equal(u, v) for u, v in zip(ra, rb))


def test_hstack_adapter_matches_reference_golden():
    g, _, _ = _algos()
    gold = np.load(os.path.join(HERE, "golden", "vqvae_hstack_ema.npz"))
    K, D = gold["E0"].shape
    layer = g.VectorQuantizerEMA(K, D, float(gold["beta"]), float(gold["decay"]), float(gold["eps"]))
    with torch.no_grad():
        layer._embedding.weight.copy_(torch.from_numpy(gold["E0"]))
        layer._ema_w.copy_(torch.from_numpy(gold["ema_w0"]))
        layer.pre_lin.weight.copy_(torch.from_numpy(gold["pre_W"]))
        layer.pre_lin.bias.copy_(torch.from_numpy(gold["pre_b"]))
    layer = layer.to(_dev()).train()
    x = _t(gold["x"]).requires_grad_(True)
    loss, quant, ppl, enc = layer(x)
    (loss + (quant * _t(gold["g_out"])).sum()).backward()
    assert np.array_equal(torch.argmax(enc, 1).cpu().numpy(), gold["idx"])
    np.testing.assert_allclose(loss.item(), gold["loss"], rtol=1e-5)
    np.testing.assert_allclose(ppl.item(), gold["ppl"], rtol=1e-5)
    np.testing.assert_allclose(quant.detach().cpu().numpy(), gold["quant"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(x.grad.cpu().numpy(), gold["gx"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(layer.pre_lin.weight.grad.cpu().numpy(), gold["gW"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(layer._embedding.weight.detach().cpu().numpy(), gold["E1"], rtol=2e-5, atol=1e-6)
    np.testing.assert_allclose(layer._ema_cluster_size.cpu().numpy(), gold["cs1"], rtol=1e-5)


# ---------------------------------------------------------------------------------------------
# edge cases, error behaviour, host path
# ---------------------------------------------------------------------------------------------
def test_empty_and_cpu_inputs():
    g, _, _ = _algos()
    layer = g.DAE_VQ_Payam(16, 8, 0.25).to(_dev())
    idx = g.vq_search(torch.empty(0, 8, device=_dev()), layer._embedding.weight.detach())
    assert idx.numel() == 0
    # CPU input on a CUDA module: moved over and back (the DataLoader-worker call site)
    x = torch.randn(6, 8)
    loss, quant, ppl, enc = layer(x)
    assert quant.device.type == "cpu" and enc.shape == (6, 16)
    with pytest.raises(RuntimeError):
        g.DAE_VQ_Payam(16, 8, 0.25)(x)            # CPU module: no CPU implementation, fail loudly
    with pytest.raises(RuntimeError):
        layer(torch.randn(5, 3, device=_dev()))    # numel not divisible by D


def test_c_abi_error_codes():
    g, L, lib = _algos()
    assert lib.g2v_vq_search(None, 0, None, None, 10, 4, 4, None, None, None, 0, 0, None) == -1
    E = torch.randn(4, 4, device=_dev())
    cb = g.prepare_codebook(E)
    z = torch.randn(10, 4, device=_dev())
    idx = torch.empty(10, dtype=torch.int32, device=_dev())
    rc = lib.g2v_vq_search(z.data_ptr(), 7, E.data_ptr(), cb.data_ptr(), 10, 4, 4, idx.data_ptr(), None, None, 0, 0, None)
    assert rc == -3 and b"dtype" in lib.g2v_strerror(rc)
    rc = lib.g2v_vq_search(z.data_ptr(), 0, E.data_ptr(), cb.data_ptr(), 10, 4, 4, idx.data_ptr(), None, None, 0, 0, None)
    assert rc == -4                                  # no workspace
    assert lib.g2v_codebook_prepare(E.data_ptr(), 4, 4, cb.data_ptr(), 8, None) == -4


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_tokenize_host_matches_device_path(dtype):
    g, _, _ = _algos()
    K, D, N = 400, 400, 50000
    E = _t(O.synth_codebook("normal", K, D, seed=4))
    z = torch.from_numpy(O.synth_latents("gru", N, D, seed=5)).to(dtype).pin_memory()
    ids, stats = g.tokenize_host(z, E, chunk_rows=8192, return_stats=True)
    ref = g.tokenize(z.to(_dev()), E).cpu()
    assert torch.equal(ids, ref)


def test_large_n_properties():
    """BASELINE config 2 size (1M rows, K=400, D=400): size-independent properties."""
    g, _, _ = _algos()
    K, D, N = 400, 400, 1_000_000
    E = _t(O.synth_codebook("normal", K, D, seed=0))
    gen = torch.Generator(device=_dev()).manual_seed(1234)
    z = torch.randn(N, D, device=_dev(), generator=gen)
    idx = g.tokenize(z, E)
    assert int(idx.min()) >= 0 and int(idx.max()) < K
    # (1) a row that IS a code maps to that code (idempotence of quantisation)
    q = E[idx.long()[:200000]]
    assert torch.equal(g.tokenize(q.contiguous(), E), idx[:200000])
    # (2) row-permutation equivariance
    perm = torch.randperm(N, device=_dev(), generator=gen)[:300000]
    assert torch.equal(g.tokenize(z[perm].contiguous(), E), idx[perm])
    # (3) audited subset against exact arithmetic on the CPU
    sub = torch.arange(0, N, 997, device=_dev())[:1000]
    zc, Ec = z[sub].cpu().numpy(), E.cpu().numpy()
    assert O.audit_indices(zc, Ec, idx[sub].cpu().numpy(), O.nearest_code_f64(zc, Ec), eps_tie=2.0 ** -40)["hard"] == 0
    # (4) statistics: counts sum to N, loss equals the mean squared distance to the chosen codes
    out, packed = g.vq_apply(z, E, idx, want_out=True, want_stats=True, want_dwr=True)
    lay = g.packed_layout(K, D)
    counts = packed[lay["counts"][0]:lay["counts"][1]]
    assert float(counts.sum()) == N and float(packed[lay["rows"]]) == N
    assert torch.equal(counts.long(), torch.bincount(idx.long(), minlength=K))
    mse = ((q - z[:200000]) ** 2).double().sum()
    loss, ppl = g.stats_finalize(packed, K, D, 0.0, 1.0)
    ref_mse = ((E[idx.long()] - z) ** 2).double().mean()
    np.testing.assert_allclose(loss.item(), ref_mse.item(), rtol=1e-5)
    # (5) residual sums: dwr[k] + counts[k]*E[k] == sum of the rows assigned to k
    dwr = packed[:K * D].view(K, D)
    dw = torch.zeros(K, D, device=_dev(), dtype=torch.float64).index_add_(0, idx.long(), z.double())
    #     (fp32 atomics: with iid latents a few low-norm codes attract >1e5 rows, so the
    #      tolerance is relative to each code's largest sum, not elementwise)
    got = (dwr.double() + counts.double()[:, None] * E.double())
    err = (got - dw).abs().amax(dim=1)
    assert bool((err <= 2e-4 * dw.abs().amax(dim=1) + 1e-3).all()), float((err / dw.abs().amax(dim=1)).max())
    assert mse.item() > 0


def test_reference_patch_swaps_classes():
    import types
    g, _, _ = _algos()
    fake = types.ModuleType("model.Autoencoder_VQVAE_model")
    fake.VQ_Payam = object
    fake.VQ_Payam_EMA = object
    done = g.patch_reference({"model.Autoencoder_VQVAE_model": fake})
    assert fake.VQ_Payam_EMA is g.VQVAE_VQ_Payam_EMA and "VQ_Payam" in done["model.Autoencoder_VQVAE_model"]
    g.unpatch_reference({"model.Autoencoder_VQVAE_model": fake})
    assert fake.VQ_Payam is object
