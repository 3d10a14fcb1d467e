"""Pin the numpy oracle (oracle/vq_oracle.py) and the torch port (oracle/torch_port.py)
against the golden vectors produced by the real reference (tests/golden/make_golden.py)."""
import json
import os

import numpy as np
import pytest
import torch

from cases import CASE_NAMES, G_LOSS, load_case
from oracle import vq_oracle as O
from oracle import torch_port as P

HERE = os.path.dirname(os.path.abspath(__file__))


def _close_rows(full, g, key, rows, rtol, atol):
    a2 = np.asarray(full).reshape(-1, np.asarray(full).shape[-1])
    np.testing.assert_allclose(a2[rows], g[key + "_rows"], rtol=rtol, atol=atol)
    s, a = float(g[key + "_sum"]), float(g[key + "_abssum"])
    assert abs(float(a2.astype(np.float64).sum()) - s) <= rtol * a + atol * a2.size


def _make_oracle(r):
    if not r["ema"]:
        return O.HardVQ(r["E0"], r["beta"])
    return O.EmaVQ(r["E0"], r["ema_w0"], r["beta"], r["decay"], r["eps"], flavour=r["flavour"],
                   W=r.get("pre_W"), b=r.get("pre_b"))


@pytest.mark.parametrize("name", CASE_NAMES)
def test_numpy_oracle_matches_reference(name):
    r, g = load_case(name)
    layer = _make_oracle(r)
    for s in range(r["steps"]):
        x, gout = r[f"x{s}"], r[f"g{s}"]
        E_used = layer.E.copy()
        res = layer.forward(x)
        ref_idx = g[f"s{s}_idx"].astype(np.int64)
        zs = O.flatten_rows(x, r["D"])
        if r["ema"] and r["flavour"] == "vqvae":
            zs = (zs @ r["pre_W"].T + r["pre_b"]).astype(np.float32)
        aud = O.audit_indices(zs, E_used, res["idx"], ref_idx)
        assert aud["hard"] == 0, aud
        # evaluate everything downstream at the REFERENCE indices so one near-tie cannot
        # cascade into the float comparisons
        if aud["mismatch"]:
            res = O.vq_forward(x, E_used, r["beta"], ema=r["ema"], idx=ref_idx,
                               search=None if zs is O.flatten_rows(x, r["D"]) else zs)
        np.testing.assert_allclose(res["loss"], g[f"s{s}_loss"], rtol=1e-5)
        np.testing.assert_allclose(res["perplexity"], g[f"s{s}_ppl"], rtol=1e-5)
        # step 0 gathers from the stored E0 (1 ulp); later EMA steps gather from the oracle's own
        # E, which matches the reference's only to its 2e-5 tolerance
        qtol = (1e-6, 1e-7) if s == 0 else (3e-5, 1e-6)
        _close_rows(res["out"], g, f"s{s}_quant", g["nrows"], *qtol)
        gx, gE = O.vq_backward(x, E_used, ref_idx, r["beta"], G_LOSS, gout, ema=r["ema"])
        _close_rows(gx, g, f"s{s}_gx", g["nrows"], 1e-5, 1e-7)
        if not r["ema"]:
            _close_rows(gE, g, f"s{s}_gE", g["krows"], 1e-4, 1e-7)
        else:
            np.testing.assert_allclose(layer.cluster_size, g[f"s{s}_cs"], rtol=1e-5, atol=1e-9)
            _close_rows(layer.ema_w, g, f"s{s}_ema_w", g["krows"], 1e-5, 1e-6)
            _close_rows(layer.E, g, f"s{s}_E", g["krows"], 2e-5, 1e-6)
    if r["ema"]:
        layer.training = False
        st = (layer.E.copy(), layer.ema_w.copy(), layer.cluster_size.copy())
        res = layer.forward(r[f"x{r['steps']}"])
        assert all(np.array_equal(a, b) for a, b in zip(st, (layer.E, layer.ema_w, layer.cluster_size)))
        zs = O.flatten_rows(r[f"x{r['steps']}"], r["D"])
        if r["flavour"] == "vqvae":
            zs = (zs @ r["pre_W"].T + r["pre_b"]).astype(np.float32)
        aud = O.audit_indices(zs, layer.E, res["idx"], g["eval_idx"].astype(np.int64))
        assert aud["hard"] == 0
        if aud["mismatch"] == 0:
            np.testing.assert_allclose(res["loss"], g["eval_loss"], rtol=1e-5)
            np.testing.assert_allclose(res["perplexity"], g["eval_ppl"], rtol=1e-5)


@pytest.mark.parametrize("name", CASE_NAMES)
def test_torch_port_matches_reference(name):
    r, g = load_case(name)
    torch.set_num_threads(1)
    if r["ema"]:
        m = P.PortVQEMA(r["K"], r["D"], r["beta"], r["decay"], r["eps"], flavour=r["flavour"])
        with torch.no_grad():
            m._ema_w.copy_(torch.from_numpy(r["ema_w0"]))
            m.pre_linear.weight.copy_(torch.from_numpy(r["pre_W"]))
            m.pre_linear.bias.copy_(torch.from_numpy(r["pre_b"]))
    else:
        m = P.PortVQ(r["K"], r["D"], r["beta"])
    with torch.no_grad():
        m._embedding.weight.copy_(torch.from_numpy(r["E0"]))
    m.train()
    for s in range(r["steps"]):
        xt = torch.from_numpy(r[f"x{s}"]).requires_grad_(True)
        loss, quant, ppl, enc = m(xt)
        (loss * G_LOSS + (quant * torch.from_numpy(r[f"g{s}"])).sum()).backward()
        # same ops, same thread count -> the port must reproduce the reference bit for bit
        assert np.array_equal(torch.argmax(enc, 1).numpy(), g[f"s{s}_idx"].astype(np.int64))
        assert np.float32(loss.item()) == g[f"s{s}_loss"]
        assert np.float32(ppl.item()) == g[f"s{s}_ppl"]
        a = xt.grad.numpy().reshape(-1, xt.shape[-1])[g["nrows"]]
        assert np.array_equal(a, g[f"s{s}_gx_rows"])
        if r["ema"]:
            assert np.array_equal(m._ema_cluster_size.numpy(), g[f"s{s}_cs"])
            assert np.array_equal(m._embedding.weight.detach().numpy()[g["krows"]], g[f"s{s}_E_rows"])


def test_flatten_quirk_layer_major():
    """[L=2,B,H].view(-1,2H) pairs adjacent batch items of ONE layer."""
    L, B, H = 2, 4, 3
    x = np.arange(L * B * H, dtype=np.float32).reshape(L, B, H)
    rows = O.flatten_rows(x, 2 * H)
    assert rows.shape == (B, 2 * H)
    np.testing.assert_array_equal(rows[0], np.concatenate([x[0, 0], x[0, 1]]))
    np.testing.assert_array_equal(rows[B // 2], np.concatenate([x[1, 0], x[1, 1]]))


def test_exact_ties_pick_first_index():
    r, g = load_case("dae_hard_dupcodes")
    idx = g["s0_idx"].astype(np.int64)
    assert np.all(idx % 2 == 0)          # duplicates live at odd rows; first index wins
    assert np.array_equal(O.HardVQ(r["E0"], r["beta"]).forward(r["x0"])["idx"], idx)


def test_hstack_adapter_fixture():
    """VectorQuantizerEMA (Autoencoder_VQVAE_model.py:1745-1812) restated literally."""
    g = np.load(os.path.join(HERE, "golden", "vqvae_hstack_ema.npz"))
    x = g["x"]
    rows = np.hstack((x[0], x[1]))                       # [B, 2H]: true per-item layout
    proj = (rows @ g["pre_W"].T + g["pre_b"]).astype(np.float32)
    res = O.vq_forward(proj, g["E0"], float(g["beta"]), ema=True)
    assert O.audit_indices(proj, g["E0"], res["idx"], g["idx"].astype(np.int64))["hard"] == 0
    np.testing.assert_allclose(res["loss"], g["loss"], rtol=1e-5)
    np.testing.assert_allclose(res["perplexity"], g["ppl"], rtol=1e-5)
    out = res["out"].reshape(2, rows.shape[0], -1)        # literal re-split (:1810)
    np.testing.assert_allclose(out, g["quant"], rtol=1e-6, atol=1e-7)
    cs, w, E = O.ema_update(np.zeros(g["E0"].shape[0], np.float32), g["ema_w0"], res["counts"],
                            res["dw"], float(g["decay"]), float(g["eps"]))
    np.testing.assert_allclose(cs, g["cs1"], rtol=1e-5)
    np.testing.assert_allclose(w, g["ema_w1"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(E, g["E1"], rtol=2e-5, atol=1e-6)


def test_state_dict_key_contract_recorded():
    keys = json.load(open(os.path.join(HERE, "golden", "state_dict_keys.json")))
    assert keys["dae.VQ_Payam"] == ["_embedding.weight"]
    assert keys["vqvae.VQ_Payam"] == ["_embedding.weight", "pre_linear.bias", "pre_linear.weight"]
    assert keys["dae.VQ_Payam_EMA"] == keys["vqvae.VQ_Payam_EMA"] == [
        "_ema_cluster_size", "_ema_w", "_embedding.weight", "pre_linear.bias", "pre_linear.weight"]


def test_gssoft_oracle_matches_reference():
    """Soft quantizer: forward values and the closed-form backward of oracle/gssoft_oracle.py
    against the real module's outputs and autograd gradients (tests/golden/make_gssoft_golden.py)."""
    from make_gssoft_golden import CFG, ROWS, inputs
    from oracle import gssoft_oracle as G
    g = np.load(os.path.join(HERE, "golden", "gssoft_trinity.npz"))
    x, E, Wm, bm, Wl, bl, g_out = inputs()
    assert float(x.astype(np.float64).sum()) == float(g["x_sum"]), "RNG stream changed"
    fw = G.forward(x, E, Wm, bm, Wl, bl, CFG["beta"])
    np.testing.assert_allclose(fw["loss"], g["loss"], rtol=2e-6)
    np.testing.assert_allclose(fw["perplexity"], g["perplexity"], rtol=1e-5)
    D = CFG["D"]
    np.testing.assert_allclose(fw["out"].reshape(-1, D)[ROWS], g["out_rows"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(fw["encodings"][ROWS], g["enc_rows"], rtol=2e-5)
    np.testing.assert_allclose(fw["encodings"].astype(np.float64).sum(), float(g["enc_sum"]), rtol=1e-6)
    bw = G.backward(fw, E, Wm, Wl, CFG["beta"], CFG["g_loss"], g_out)
    # the reference's gradients are fp32 autograd through two GEMMs and an un-stabilised normalisation
    for key, ref, scale in (("x", g["gx_rows"], None), ("E", g["gE_rows"], None), ("Wm", g["gWm_rows"], None),
                            ("Wl", g["gWl_rows"], None)):
        got = bw[key].reshape(-1, D)[ROWS]                            # x as rows of D, like the golden
        np.testing.assert_allclose(got, ref, rtol=2e-3, atol=2e-4 * np.abs(ref).max())
    np.testing.assert_allclose(bw["bm"], g["gbm"], rtol=2e-3, atol=2e-4 * np.abs(g["gbm"]).max())
    np.testing.assert_allclose(bw["bl"], g["gbl"], rtol=2e-3, atol=2e-4 * np.abs(g["gbl"]).max())
    np.testing.assert_allclose(np.abs(bw["x"]).sum(), float(g["gx_abssum"]), rtol=1e-4)
    np.testing.assert_allclose(np.abs(bw["E"]).sum(), float(g["gE_abssum"]), rtol=1e-3)
    assert "pre_linear.weight" in set(g["state_keys"].tolist())        # unused Linear still in the state_dict
