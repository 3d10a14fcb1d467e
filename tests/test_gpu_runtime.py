"""Runtime behaviour of the drop-in modules on the GPU: CUDA-graph capture and replay of the training step,
in-place EMA state, a module on a device that is not the current one, and the tokenisers on the flavour that
projects its rows first (Autoencoder_VQVAE_model.VQ_Payam_EMA, :1230)."""
import numpy as np
import pytest
import torch

from oracle import vq_oracle as O

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def _layer(g, dev, K=512, D=400, cls="DAE_VQ_Payam_EMA", seed=0):
    layer = getattr(g, cls)(K, D, 0.25, 0.85)
    with torch.no_grad():
        layer._embedding.weight.copy_(torch.from_numpy(O.synth_codebook("uniform1", K, D, seed=seed)))
        layer._ema_w.copy_(torch.from_numpy(np.random.default_rng(seed + 1).standard_normal((K, D), dtype=np.float32)))
    return layer.to(dev).train()


def _state(layer):
    return [t.detach().clone() for t in (layer._embedding.weight, layer._ema_w, layer._ema_cluster_size)]


def _close(u, v, what=""):
    """Two runs of the same step agree to fp32 atomics order: the per-code residual sums are accumulated with
    red.add.f32 in a scheduling-dependent order, so EMA state (and everything computed from it one step later)
    matches to a few ulp, not bitwise.  Integer outputs and one-hots must be identical."""
    if u.dtype in (torch.int32, torch.int64):
        assert torch.equal(u, v), what
    else:
        # (absolute tolerance relative to the tensor's scale: sums of a few hundred rows cancel to small elements)
        torch.testing.assert_close(u, v, rtol=1e-4, atol=1e-5 * max(1.0, float(v.abs().max())),
                                   msg=lambda m: f"{what}: {m}")


def test_inplace_ema_equals_fresh_tensor_ema():
    import gesture2vec_b200 as g
    a, b = _layer(g, DEV), _layer(g, DEV)
    b.ema_inplace = True
    ptrs = [t.data_ptr() for t in (b._embedding.weight, b._ema_w, b._ema_cluster_size)]
    gq = torch.randn(2, 128, 200, device=DEV)
    for s in range(3):
        x = torch.from_numpy(O.synth_latents("gru", 128, 400, seed=10 + s).reshape(2, 128, 200)).to(DEV)
        outs = []
        for layer in (a, b):
            xi = x.clone().requires_grad_(True)
            loss, q, ppl, enc = layer(xi)
            (loss * 3.0 + (q * gq).sum()).backward()
            outs.append((loss.detach(), q.detach(), ppl, enc, xi.grad))
        for u, v in zip(*outs):
            _close(u, v, f"step {s} outputs")
        for u, v in zip(_state(a), _state(b)):
            _close(u, v, f"step {s} state")
    assert ptrs == [t.data_ptr() for t in (b._embedding.weight, b._ema_w, b._ema_cluster_size)]


def test_cuda_graph_capture_and_replay_of_the_training_step():
    """The N = 128 step of config/VQ-VAE.yml as ONE graph launch: capture forward + backward of the in-place EMA
    layer, replay it three times on new inputs, and compare every output and the state with the eager layer."""
    import gesture2vec_b200 as g
    eager, graphed = _layer(g, DEV), _layer(g, DEV)
    graphed.ema_inplace = True
    xs = [torch.from_numpy(O.synth_latents("gru", 128, 400, seed=20 + s).reshape(2, 128, 200)).to(DEV) for s in range(5)]
    gq = torch.randn(2, 128, 200, device=DEV)
    static_x = xs[0].clone().requires_grad_(True)
    # warm-up on a side stream (allocations, tensor maps, codebook aux), as torch.cuda.graphs asks for
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for s in range(2):
            static_x.grad = None
            with torch.no_grad():
                static_x.copy_(xs[s])
            loss, q, ppl, enc = graphed(static_x)
            (loss * 3.0 + (q * gq).sum()).backward()
    torch.cuda.current_stream().wait_stream(side)
    for s in range(2):
        xi = xs[s].clone().requires_grad_(True)
        loss, q, ppl, enc = eager(xi)
        (loss * 3.0 + (q * gq).sum()).backward()
    graph = torch.cuda.CUDAGraph()
    static_x.grad = None
    with torch.cuda.graph(graph):
        s_loss, s_q, s_ppl, s_enc = graphed(static_x)
        (s_loss * 3.0 + (s_q * gq).sum()).backward()
    s_grad = static_x.grad
    for s in range(2, 5):
        with torch.no_grad():
            static_x.copy_(xs[s])
        graph.replay()
        xi = xs[s].clone().requires_grad_(True)
        loss, q, ppl, enc = eager(xi)
        (loss * 3.0 + (q * gq).sum()).backward()
        torch.cuda.synchronize()
        for u, v in zip((s_loss, s_q, s_ppl, s_enc, s_grad), (loss.detach(), q.detach(), ppl, enc, xi.grad)):
            _close(u, v, f"replay {s} outputs")
        for u, v in zip(_state(eager), _state(graphed)):
            _close(u, v, f"replay {s} state")


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_module_on_a_device_that_is_not_current():
    import gesture2vec_b200 as g
    d1 = torch.device("cuda:1")
    assert torch.cuda.current_device() == 0
    a, b = _layer(g, DEV), _layer(g, d1)
    x = torch.from_numpy(O.synth_latents("gru", 300, 400, seed=3))
    ra = a(x.to(DEV).requires_grad_(True))
    rb = b(x.to(d1).requires_grad_(True))
    assert torch.cuda.current_device() == 0
    for u, v in zip(ra, rb):
        assert v.device == d1 and torch.equal(u.cpu(), v.cpu())
    ids = g.tokenize(x.to(d1), b._embedding.weight.detach())
    assert ids.device == d1 and torch.equal(ids.cpu(), g.tokenize(x.to(DEV), a._embedding.weight.detach()).cpu())
    km = g.KMeans(n_clusters=8, init=x[:8].numpy(), max_iter=3, device=d1).fit(x.numpy())
    km0 = g.KMeans(n_clusters=8, init=x[:8].numpy(), max_iter=3, device=DEV).fit(x.numpy())
    assert np.array_equal(km.labels_, km0.labels_)


def test_tokenizers_apply_the_projection_of_the_vqvae_flavour():
    """GestureTokenizer(module) and module.tokenize() give argmax(forward(x).encodings) for the flavour whose
    search runs on pre_linear(x) -- fp32 and bf16 rows, device and host inputs."""
    import gesture2vec_b200 as g
    K, D, N = 512, 400, 3000
    layer = _layer(g, DEV, K, D, cls="VQVAE_VQ_Payam_EMA", seed=4).eval()
    x = torch.from_numpy(O.synth_latents("gru", N, D, seed=5)).to(DEV)
    with torch.no_grad():
        enc = layer(x)[3]
    want = torch.argmax(enc, 1)
    assert torch.equal(layer.tokenize(x).long(), want)
    tok = g.GestureTokenizer(layer)
    assert np.array_equal(tok.encode_rows(x), want.cpu().numpy())
    assert np.array_equal(tok.encode_rows(x.cpu().numpy()), want.cpu().numpy())
    # 16-bit rows: projected in fp32 from the stored values, like forward() does with a 16-bit input
    x16 = x.to(torch.bfloat16)
    with torch.no_grad():
        want16 = torch.argmax(layer(x16)[3], 1)
    assert torch.equal(layer.tokenize(x16).long(), want16)
    assert np.array_equal(tok.encode_rows(x16), want16.cpu().numpy())
    # and the raw codebook (no projection) differs, which is what the old tokenizer silently returned
    raw = g.tokenize(x, layer._embedding.weight.detach()).long()
    assert not torch.equal(raw, want)


def test_deterministic_statistics_are_bit_reproducible_and_tight():
    """layer.deterministic = True: atomics-free, fixed-order fp64 sums -- two runs give the same bits (the default
    fp32-atomics pass does not promise that), and the sums are within 1e-6 of fp64 (default pass: 1e-4)."""
    import gesture2vec_b200 as g
    K, D, N = 400, 400, 200_000
    E = torch.randn(K, D, device=DEV, generator=torch.Generator(device=DEV).manual_seed(0))
    w = 1.0 / torch.arange(1, K + 1, device=DEV, dtype=torch.float64) ** 1.1
    code = torch.multinomial(w / w.sum(), N, replacement=True, generator=torch.Generator(device=DEV).manual_seed(1))
    z = (E[code] + 0.1 * torch.randn(N, D, device=DEV, generator=torch.Generator(device=DEV).manual_seed(2))).contiguous()
    idx = g.tokenize(z, E)
    runs = [g.vq_apply(z, E, idx, want_out=True, want_stats=True, want_dwr=True, deterministic=True) for _ in range(3)]
    for out, packed in runs[1:]:
        assert torch.equal(packed, runs[0][1]) and torch.equal(out, runs[0][0])
    packed = runs[0][1]
    lay = g.packed_layout(K, D)
    counts = packed[lay["counts"][0]:lay["counts"][1]]
    assert torch.equal(counts.long(), torch.bincount(idx.long(), minlength=K))
    ref = torch.zeros(K, D, device=DEV, dtype=torch.float64).index_add_(0, idx.long(), z.double() - E.double()[idx.long()])
    got = packed[:K * D].view(K, D).double()
    scale = ref.abs().amax(dim=1, keepdim=True).clamp_min(1e-12)
    assert float(((got - ref).abs() / scale).max()) < 1e-6
    sse = ((E[idx.long()] - z).double() ** 2).sum()
    assert abs(float(packed[lay["sse"]]) - float(sse)) <= 1e-6 * float(sse)
    # through the module: two identically initialised layers stay bit-identical over EMA steps
    a, b = _layer(g, DEV, 512, 400), _layer(g, DEV, 512, 400)
    a.deterministic = b.deterministic = True
    for s in range(3):
        x = torch.from_numpy(O.synth_latents("gru", 4096, 400, seed=30 + s)).to(DEV)
        ra, rb = a(x), b(x)
        assert all(torch.equal(u, v) for u, v in zip(ra, rb))
        assert all(torch.equal(u, v) for u, v in zip(_state(a), _state(b)))


def test_pre_linear_fold_matches_the_explicit_projection():
    """Autoencoder_VQVAE_model.VQ_Payam_EMA searches on pre_linear(z) (:1230).  Default here: pre_linear folded into a
    [K, D+4] codebook and the RAW rows searched; `fold_projection = False`: every row projected first (g2v_gemm_f32).
    Same indices (near-ties aside), same loss / perplexity / EMA state, over 65 659 rows and two EMA steps."""
    import gesture2vec_b200 as g
    import gpu_synth as S
    K, D, N = 512, 400, 65536 + 123
    a = _layer(g, DEV, K, D, cls="VQVAE_VQ_Payam_EMA", seed=7)
    b = _layer(g, DEV, K, D, cls="VQVAE_VQ_Payam_EMA", seed=7)
    b.load_state_dict(a.state_dict())
    b.fold_projection = False
    assert a.fold_projection and a._fold(DEV) is not None and b._fold(DEV) is None
    gq = torch.randn(N, D, device=DEV, generator=torch.Generator(device=DEV).manual_seed(1))
    for s in range(2):
        x = torch.from_numpy(O.synth_latents("gru", N, D, seed=40 + s)).to(DEV)
        zs = b._search_rows(x)                                           # the projected rows the reference searches
        E = b._embedding.weight.detach().clone()
        ia, ib = a.tokenize(x), b.tokenize(x)
        aud = S.audit(zs, E, ia, ib)                                     # EPS_TIE: the reference's own fp32 noise floor
        assert aud["hard"] == 0 and aud["mismatch"] <= N // 2000, aud
        outs = []
        for layer, ids in ((a, ib), (b, ib)):                            # same indices: compare the arithmetic downstream
            xi = x.clone().requires_grad_(True)
            loss, q, ppl, enc = layer.forward_with_indices(xi, ids)
            (loss * 2.0 + (q * gq).sum()).backward()
            outs.append((loss.detach(), q.detach(), ppl, xi.grad))
        for u, v in zip(*outs):
            _close(u, v, f"step {s}")
        for u, v in zip(_state(a), _state(b)):
            _close(u, v, f"state {s}")
    tok = g.GestureTokenizer(a)
    x = torch.from_numpy(O.synth_latents("gru", 5000, D, seed=50)).to(DEV)
    assert np.array_equal(tok.encode_rows(x), a.tokenize(x).cpu().numpy())
    assert np.array_equal(tok.encode_rows(x.cpu().numpy()), a.tokenize(x).cpu().numpy())


def test_cuda_graph_capture_of_a_bulk_search_with_the_refine_side_stream():
    """A 65 536-row search forks the refine pass of its whole-row re-ranks onto the library's side stream and joins
    it again (events): the fork / join is capturable, and the replay on new rows gives the eager indices."""
    import gesture2vec_b200 as g
    K, D, N = 1024, 400, 65536
    E = torch.randn(K, D, device=DEV, generator=torch.Generator(device=DEV).manual_seed(3))
    cb = g.prepare_codebook(E)
    zs = [torch.randn(N, D, device=DEV, generator=torch.Generator(device=DEV).manual_seed(10 + s)) for s in range(3)]
    static_z, static_idx = zs[0].clone(), torch.empty(N, dtype=torch.int32, device=DEV)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):                       # warm-up: workspace, tensor maps, the side stream itself
        g.vq_search(static_z, E, cb, out=static_idx)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        g.vq_search(static_z, E, cb, out=static_idx)
    for s in range(1, 3):
        static_z.copy_(zs[s])
        graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(static_idx, g.vq_search(zs[s], E, cb))
