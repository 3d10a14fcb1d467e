"""World-size-2 gloo test of the data-parallel EMA exchange: every rank packs its local statistics
into the [dwr | counts | sse | rows] buffer, one sum all-reduce, identical update on every rank,
and the result equals the single-process update on the concatenated batch.  The compute pieces
here come from the oracle (CPU); the exchanged layout, the reducer and the sharding are the
product's host logic."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import vq_oracle as O

K, D, N, DECAY, EPS, BETA = 48, 16, 400, 0.85, 1e-5, 0.25


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _local_packed(z, E):
    """What g2v_vq_apply + g2v_vq_stats_pack produce for a shard (restated with the oracle)."""
    from gesture2vec_b200 import packed_layout
    lay = packed_layout(K, D)
    r = O.vq_forward(z, E, BETA, ema=True)
    packed = np.zeros(lay["numel"], np.float32)
    dwr = r["dw"].astype(np.float64) - r["counts"][:, None] * E.astype(np.float64)
    packed[lay["dwr"][0]:lay["dwr"][1]] = dwr.astype(np.float32).ravel()
    packed[lay["counts"][0]:lay["counts"][1]] = r["counts"]
    packed[lay["sse"]] = np.sum((r["q"].astype(np.float64) - z) ** 2)
    packed[lay["rows"]] = z.shape[0]
    return packed, r


def _update_from_packed(packed, E, cs, w):
    from gesture2vec_b200 import packed_layout
    lay = packed_layout(K, D)
    counts = packed[lay["counts"][0]:lay["counts"][1]]
    dw = packed[lay["dwr"][0]:lay["dwr"][1]].reshape(K, D) + counts[:, None] * E
    cs2, w2, E2 = O.ema_update(cs, w, counts, dw.astype(np.float32), DECAY, EPS)
    loss = np.float32(BETA) * np.float32(packed[lay["sse"]] / (packed[lay["rows"]] * D))
    ppl = O.perplexity_from_counts(counts.astype(np.int64), int(packed[lay["rows"]]))
    return cs2, w2, E2, loss, ppl


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from gesture2vec_b200 import StatsAllReduce, shard_rows
        E = O.synth_codebook("uniform1", K, D, seed=0)
        w0 = np.random.default_rng(1).standard_normal((K, D), dtype=np.float32)
        z = O.synth_latents("gru", N, D, seed=2)
        b, e = shard_rows(N, rank, world)
        packed, _ = _local_packed(z[b:e], E)
        t = torch.from_numpy(packed)
        red = StatsAllReduce()
        red(t)                                               # the one exchange of the step
        assert red.calls == 1 and red.bytes == t.numel() * 4
        cs, w, E2, loss, ppl = _update_from_packed(t.numpy(), E, np.zeros(K, np.float32), w0)
        out[rank] = (cs, w, E2, float(loss), float(ppl))
    finally:
        dist.destroy_process_group()


def test_two_rank_ema_equals_single_process():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    r0, r1 = out[0], out[1]
    for a, b in zip(r0[:3], r1[:3]):                         # bit-identical state on every rank
        assert np.array_equal(a, b)
    assert r0[3:] == r1[3:]
    # single process on the concatenated batch
    E = O.synth_codebook("uniform1", K, D, seed=0)
    w0 = np.random.default_rng(1).standard_normal((K, D), dtype=np.float32)
    z = O.synth_latents("gru", N, D, seed=2)
    ref = O.EmaVQ(E, w0, BETA, DECAY, EPS)
    res = ref.forward(z)
    np.testing.assert_allclose(r0[0], ref.cluster_size, rtol=1e-6)
    np.testing.assert_allclose(r0[1], ref.ema_w, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(r0[2], ref.E, rtol=2e-5, atol=1e-6)
    np.testing.assert_allclose(r0[3], res["loss"], rtol=1e-5)
    np.testing.assert_allclose(r0[4], res["perplexity"], rtol=1e-5)


def test_reducer_requires_process_group():
    from gesture2vec_b200 import StatsAllReduce
    with pytest.raises(RuntimeError):
        StatsAllReduce()(torch.zeros(4))
