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


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import gesture2vec_b200 as g
        layer = _make_layer(dev)
        red = g.enable_data_parallel_ema(layer)
        z = torch.from_numpy(O.synth_latents("gru", N, D, seed=2))
        b, e = g.shard_rows(N, rank, world)
        x = z[b:e].to(dev).requires_grad_(True)
        for _ in range(2):
            x.grad = None
            loss, q, ppl, _ = layer(x)
            (loss + q.sum()).backward()
        torch.cuda.synchronize()
        assert red.calls == 2 and red.bytes == 2 * g.packed_numel(K, D) * 4
        out[rank] = dict(E=layer._embedding.weight.detach().cpu().numpy(), w=layer._ema_w.detach().cpu().numpy(),
                         cs=layer._ema_cluster_size.cpu().numpy(), loss=float(loss), ppl=float(ppl),
                         gx=x.grad.cpu().numpy(), span=(b, e))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_ema_matches_single_gpu():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
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
    np.testing.assert_allclose(r0["loss"], float(loss), rtol=1e-5)
    np.testing.assert_allclose(r0["ppl"], float(ppl), rtol=1e-5)
    # input gradient: local-mean convention, 2*beta*(x-q)/(N_local*D); DDP's mean over ranks turns
    # it into the single-process gradient, so here it is N/N_local times the single-GPU value
    gs = x.grad.cpu().numpy()
    for r in (r0, r1):
        b, e = r["span"]
        np.testing.assert_allclose(r["gx"] - 1.0, (N / (e - b)) * (gs[b:e] - 1.0), rtol=1e-4, atol=1e-7)
