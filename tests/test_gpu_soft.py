"""The soft quantizer VQ_Payam_GSSoft on the GPU (tcgen05 GEMMs + g2v_soft_* kernels, through the drop-in module)
against the golden vectors of the REAL reference module (forward values and autograd gradients,
tests/golden/make_gssoft_golden.py) and against the numpy oracle at other shapes."""
import os

import numpy as np
import pytest
import torch

from make_gssoft_golden import CFG, ROWS, inputs
from oracle import gssoft_oracle as G

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")
HERE = os.path.dirname(os.path.abspath(__file__))


def _layer(g, K, D, beta, E, Wm, bm, Wl, bl):
    layer = g.VQVAE_VQ_Payam_GSSoft(K, D, beta)
    with torch.no_grad():
        layer._embedding.weight.copy_(torch.from_numpy(E))
        layer.mean_layer.weight.copy_(torch.from_numpy(Wm)); layer.mean_layer.bias.copy_(torch.from_numpy(bm))
        layer.logvar_layer.weight.copy_(torch.from_numpy(Wl)); layer.logvar_layer.bias.copy_(torch.from_numpy(bl))
    return layer.to(DEV)


def test_gssoft_matches_reference_golden():
    import gesture2vec_b200 as g
    gold = np.load(os.path.join(HERE, "golden", "gssoft_trinity.npz"))
    x, E, Wm, bm, Wl, bl, g_out = inputs()
    assert float(x.astype(np.float64).sum()) == float(gold["x_sum"]), "RNG stream changed"
    K, D = CFG["K"], CFG["D"]
    layer = _layer(g, K, D, CFG["beta"], E, Wm, bm, Wl, bl)
    assert sorted(layer.state_dict().keys()) == sorted(gold["state_keys"].tolist())
    xt = torch.from_numpy(x).to(DEV).requires_grad_(True)
    loss, out, ppl, enc = layer(xt)
    assert out.shape == xt.shape and out.is_contiguous() and tuple(enc.shape) == (x.size // D, K)
    (CFG["g_loss"] * loss + (out * torch.from_numpy(g_out).to(DEV)).sum()).backward()
    # forward: fp32 tolerances (the products run as split-fp16 tensor-core GEMMs at ~2^-21 relative)
    np.testing.assert_allclose(loss.item(), gold["loss"], rtol=5e-6)
    np.testing.assert_allclose(ppl.item(), gold["perplexity"], rtol=1e-5)
    np.testing.assert_allclose(out.detach().cpu().numpy().reshape(-1, D)[ROWS], gold["out_rows"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(enc.detach().cpu().numpy()[ROWS], gold["enc_rows"], rtol=5e-5)
    np.testing.assert_allclose(float(enc.double().sum()), float(gold["enc_sum"]), rtol=1e-6)
    np.testing.assert_allclose(float(out.detach().double().sum()), float(gold["out_sum"]), rtol=1e-5)
    # backward: the same tolerances the pinned oracle meets against the reference's fp32 autograd
    grads = {"x": xt.grad, "E": layer._embedding.weight.grad, "Wm": layer.mean_layer.weight.grad,
             "Wl": layer.logvar_layer.weight.grad}
    for key, ref in (("x", gold["gx_rows"]), ("E", gold["gE_rows"]), ("Wm", gold["gWm_rows"]), ("Wl", gold["gWl_rows"])):
        got = grads[key].cpu().numpy().reshape(-1, D)[ROWS]
        np.testing.assert_allclose(got, ref, rtol=2e-3, atol=2e-4 * np.abs(ref).max(), err_msg=key)
    np.testing.assert_allclose(layer.mean_layer.bias.grad.cpu().numpy(), gold["gbm"], rtol=2e-3, atol=2e-4 * np.abs(gold["gbm"]).max())
    np.testing.assert_allclose(layer.logvar_layer.bias.grad.cpu().numpy(), gold["gbl"], rtol=2e-3, atol=2e-4 * np.abs(gold["gbl"]).max())
    np.testing.assert_allclose(float(xt.grad.double().abs().sum()), float(gold["gx_abssum"]), rtol=1e-4)
    np.testing.assert_allclose(float(layer._embedding.weight.grad.double().abs().sum()), float(gold["gE_abssum"]), rtol=1e-3)
    assert layer.pre_linear.weight.grad is None                      # unused Linear, as in the reference


@pytest.mark.parametrize("N,K,D", [(300, 400, 400), (70000, 512, 400), (64, 80, 40)])
def test_gssoft_matches_oracle_other_shapes(N, K, D):
    """GENEA shapes, a batch long enough for the single-term row reductions, and the frame-level sizes."""
    import gesture2vec_b200 as g
    rng = np.random.default_rng(N + K)
    x = np.tanh(0.8 * rng.standard_normal((N, D))).astype(np.float32)
    E = rng.standard_normal((K, D)).astype(np.float32)
    b = 1.0 / np.sqrt(D)
    Wm = rng.uniform(-b, b, (D, D)).astype(np.float32); bm = rng.uniform(-b, b, D).astype(np.float32)
    Wl = (rng.uniform(-b, b, (K, D)) * 0.2).astype(np.float32); bl = rng.uniform(-b, b, K).astype(np.float32)
    g_out = rng.standard_normal((N, D)).astype(np.float32)
    layer = _layer(g, K, D, 0.25, E, Wm, bm, Wl, bl)
    xt = torch.from_numpy(x).to(DEV).requires_grad_(True)
    loss, out, ppl, enc = layer(xt)
    (2.0 * loss + (out * torch.from_numpy(g_out).to(DEV)).sum()).backward()
    fw = G.forward(x, E, Wm, bm, Wl, bl, 0.25)
    np.testing.assert_allclose(loss.item(), fw["loss"], rtol=1e-5)
    np.testing.assert_allclose(ppl.item(), fw["perplexity"], rtol=2e-5)
    np.testing.assert_allclose(out.detach().cpu().numpy(), fw["out"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(enc.detach().cpu().numpy(), fw["encodings"], rtol=1e-4, atol=1e-9)
    if N <= 1000:                                                     # fp64 oracle backward: small cases only
        bw = G.backward(fw, E, Wm, Wl, 0.25, 2.0, g_out)
        for key, got in (("x", xt.grad), ("E", layer._embedding.weight.grad), ("Wm", layer.mean_layer.weight.grad),
                         ("bm", layer.mean_layer.bias.grad), ("Wl", layer.logvar_layer.weight.grad),
                         ("bl", layer.logvar_layer.bias.grad)):
            ref = bw[key]
            np.testing.assert_allclose(got.cpu().numpy().reshape(ref.shape), ref, rtol=2e-3,
                                       atol=2e-4 * np.abs(ref).max(), err_msg=key)
    else:
        # long batch: gradients against an fp64 torch autograd of the same closed form on the GPU
        xd = torch.from_numpy(x).to(DEV).double().requires_grad_(True)
        P = [torch.from_numpy(a).to(DEV).double().requires_grad_(True) for a in (E, Wm, bm, Wl, bl)]
        m = xd @ P[1].t() + P[2]
        lv = m @ P[3].t() + P[4]
        d = (m * m).sum(1, keepdim=True) + (P[0] * P[0]).sum(1) - 2 * m @ P[0].t()
        s = 1.0 / torch.exp(lv) ** 2
        pt = torch.exp(-(d / 400) * 0.5 * s) / torch.sqrt(s)
        pp = pt / pt.sum(1, keepdim=True)
        q = pp @ P[0]
        l = ((q - xd.detach()) ** 2).mean() + 0.25 * ((q.detach() - xd) ** 2).mean()
        o = xd + (q - xd).detach()
        (2.0 * l + (o * torch.from_numpy(g_out).to(DEV).double()).sum()).backward()
        for name, got, ref in (("x", xt.grad, xd.grad), ("E", layer._embedding.weight.grad, P[0].grad),
                               ("Wm", layer.mean_layer.weight.grad, P[1].grad), ("bm", layer.mean_layer.bias.grad, P[2].grad),
                               ("Wl", layer.logvar_layer.weight.grad, P[3].grad), ("bl", layer.logvar_layer.bias.grad, P[4].grad)):
            err = float((got.double() - ref).abs().max() / ref.abs().max())
            assert err < 2e-3, (name, err)


def test_gssoft_swap_into_a_model_and_eval():
    import gesture2vec_b200 as g
    net = torch.nn.Module()
    net.vq_layer = g.VQVAE_VQ_Payam_EMA(64, 32, 0.25, 0.85).to(DEV)
    new = g.swap_vq_layer(net, kind="VQ_Payam_GSSoft", flavour="vqvae")
    assert isinstance(new, g.VQVAE_VQ_Payam_GSSoft) and net.vq_layer is new
    with torch.no_grad():
        loss, q, ppl, enc = new.eval()(torch.randn(2, 10, 16, device=DEV))
    assert q.shape == (2, 10, 16) and enc.shape == (10, 64) and torch.isfinite(loss)
    np.testing.assert_allclose(enc.sum(1).cpu().numpy(), 1.0, rtol=1e-5)
