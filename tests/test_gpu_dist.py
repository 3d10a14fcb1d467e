"""Two-GPU NCCL test of data-parallel EMA training through the drop-in module: the codebooks stay
bit-identical across ranks and equal the single-GPU update on the concatenated batch.  Needs two
CUDA devices (skipped on a one-GPU box; run with `gpurun --gpus 2`)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import vq_oracle as O

pytestmark = pytest.mark.gpu
K, D, N, DECAY, BETA = 512, 400, 4096, 0.85, 0.25


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _make_layer(dev):
    import gesture2vec_b200 as g
    layer = g.DAE_VQ_Payam_EMA(K, D, BETA, DECAY)
    with torch.no_grad():
        layer._embedding.weight.copy_(torch.from_numpy(O.synth_codebook("uniform1", K, D, seed=0)))
        layer._ema_w.copy_(torch.from_numpy(np.random.default_rng(1).standard_normal((K, D), dtype=np.float32)))
    return layer.to(dev).train()


def _commitment_grad(layer, x):
    """d loss / d x alone (no gradient through `quantized`): 2*beta*(x-q)/(N_local*D), free of the '+1'."""
    x.grad = None
    loss, _, _, _ = layer(x)
    loss.backward()
    return x.grad.cpu().numpy()


def _worker(rank, world, port, out, overlap):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import gesture2vec_b200 as g
        layer = _make_layer(dev)
        red = g.enable_data_parallel_ema(layer, overlap=overlap)
        assert (red.stream is not None) == overlap
        z = torch.from_numpy(O.synth_latents("gru", N, D, seed=2))
        b, e = g.shard_rows(N, rank, world)
        x = z[b:e].to(dev).requires_grad_(True)
        for _ in range(2):
            x.grad = None
            loss, q, ppl, _ = layer(x)
            (loss + q.sum()).backward()
        torch.cuda.synchronize()
        assert red.calls == 2 and red.bytes == 2 * g.packed_numel(K, D) * 4
        res = dict(E=layer._embedding.weight.detach().cpu().numpy(), w=layer._ema_w.detach().cpu().numpy(),
                   cs=layer._ema_cluster_size.cpu().numpy(), loss=float(loss.detach()), ppl=float(ppl),
                   gx=x.grad.cpu().numpy(), span=(b, e))
        res["gl"] = _commitment_grad(layer, x)
        out[rank] = res
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("overlap", [False, True])
def test_two_gpu_ema_matches_single_gpu(overlap):
    """overlap=True: the all-reduce and the finalise launch run on a side stream under the one-hot kernel."""
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out, overlap), nprocs=world, join=True)
    r0, r1 = out[0], out[1]
    for k in ("E", "w", "cs"):
        assert np.array_equal(r0[k], r1[k]), k             # identical all-reduce result -> identical state
    assert r0["loss"] == r1["loss"] and r0["ppl"] == r1["ppl"]
    # single GPU on the concatenated batch
    dev = torch.device("cuda:0")
    layer = _make_layer(dev)
    x = torch.from_numpy(O.synth_latents("gru", N, D, seed=2)).to(dev).requires_grad_(True)
    for _ in range(2):
        x.grad = None
        loss, q, ppl, _ = layer(x)
        (loss + q.sum()).backward()
    np.testing.assert_allclose(r0["cs"], layer._ema_cluster_size.cpu().numpy(), rtol=1e-6)
    # fp32 atomics accumulate the per-code sums in a different order on 2 x N/2 rows than on N rows
    np.testing.assert_allclose(r0["w"], layer._ema_w.detach().cpu().numpy(), rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(r0["E"], layer._embedding.weight.detach().cpu().numpy(), rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(r0["loss"], float(loss.detach()), rtol=1e-5)
    np.testing.assert_allclose(r0["ppl"], float(ppl), rtol=1e-5)
    # input gradient: local-mean convention, 2*beta*(x-q)/(N_local*D); DDP's mean over ranks turns
    # it into the single-process gradient, so here it is N/N_local times the single-GPU value.
    # With g_quantized = 1 the commitment part (~1e-7) sits below one ulp of the sum, so that comparison can
    # only be made to a few ulp of 1.0; the commitment gradient alone is compared to fp32 accuracy.
    gs = x.grad.cpu().numpy()
    gl = _commitment_grad(layer, x)
    for r in (r0, r1):
        b, e = r["span"]
        np.testing.assert_allclose(r["gx"], 1.0 + (N / (e - b)) * (gs[b:e] - 1.0), rtol=0, atol=3e-7)
        np.testing.assert_allclose(r["gl"], (N / (e - b)) * gl[b:e], rtol=1e-5, atol=1e-12)


def _km_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import gesture2vec_b200 as g
        from oracle import kmeans_oracle as KO
        X, init = KO.synth_blobs(6000, 8, 12, seed=7)
        b, e = g.shard_rows(X.shape[0], rank, world)
        km = g.KMeans(n_clusters=12, init=init, max_iter=50, tol=1e-4, device=dev, stats_reduce=g.StatsAllReduce(),
                      count_reduce=lambda t: dist.all_reduce(t)).fit(X[b:e])
        out[rank] = dict(C=km.cluster_centers_, labels=km.labels_, n_iter=km.n_iter_, inertia=km.inertia_, span=(b, e))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_kmeans_matches_single_gpu():
    """Rows sharded over two ranks, one all-reduce of the packed statistics per Lloyd iteration: the same
    centres on both ranks, and the same fit as one GPU (and the CPU oracle) on all rows."""
    from oracle import kmeans_oracle as KO
    world = 2
    out = mp.Manager().dict()
    mp.spawn(_km_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    r0, r1 = out[0], out[1]
    assert np.array_equal(r0["C"], r1["C"]) and r0["n_iter"] == r1["n_iter"] and r0["inertia"] == r1["inertia"]
    X, init = KO.synth_blobs(6000, 8, 12, seed=7)
    C, labels, inertia, n_iter = KO.lloyd(X, init, 50, 1e-4)
    assert r0["n_iter"] == n_iter
    np.testing.assert_allclose(r0["C"], C, rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(r0["inertia"], inertia, rtol=1e-5)
    for r in (r0, r1):
        b, e = r["span"]
        assert np.array_equal(r["labels"], labels[b:e])
