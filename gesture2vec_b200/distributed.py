"""Data-parallel plumbing of the quantizer path (one process per GPU, torch.distributed).

Tokenisation shards by rows with no communication.  EMA training exchanges ONE packed fp32
buffer per step, [dwr (K*D) | counts (K) | sse | rows] (include/g2v_vq.h g2v_vq_stats_pack), by a
sum all-reduce (NCCL over NVLink on the GPU box, gloo in the CPU tests); every rank then runs the
identical EMA update, so codebooks stay bit-identical across ranks and equal the single-process
update on the concatenated batch.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_rows(n_rows: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced row range [begin, end) of `rank` (first n%world ranks get one extra)."""
    if world <= 0 or not (0 <= rank < world) or n_rows < 0:
        raise ValueError("bad shard arguments")
    base, extra = divmod(n_rows, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def packed_layout(K: int, D: int) -> dict:
    """Offsets (in floats) of the packed statistics buffer."""
    return {"dwr": (0, K * D), "counts": (K * D, K * D + K), "sse": K * D + K, "rows": K * D + K + 1,
            "numel": K * D + K + 2}


class StatsAllReduce:
    """Callable handed to a quantizer (`layer.stats_reduce = StatsAllReduce(group)`): in-place sum
    all-reduce of the packed statistics.

    overlap=False: stream-ordered on the compute stream between the pack and the finalise launch.
    overlap=True (CUDA only): the module runs the exchange AND the finalise launch (loss, perplexity, EMA
    update, codebook aux) on `self.stream`, a side stream that waits for the pack; the main stream meanwhile
    runs whatever the forward pass still has to enqueue (the dense one-hot the reference's contract returns)
    and waits for the side stream's event before the module returns -- SURVEY.md 8e "Overlap"."""

    def __init__(self, group: Optional[dist.ProcessGroup] = None, overlap: bool = False,
                 device: Optional[torch.device] = None):
        self.group = group
        self.calls = 0
        self.bytes = 0
        self.stream: Optional[torch.cuda.Stream] = None
        if overlap:
            if not torch.cuda.is_available():
                raise RuntimeError("StatsAllReduce(overlap=True) needs a CUDA device")
            self.stream = torch.cuda.Stream(device=device)

    def __call__(self, packed: torch.Tensor) -> None:
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("StatsAllReduce needs an initialised torch.distributed process group")
        dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=self.group)
        self.calls += 1
        self.bytes += packed.numel() * packed.element_size()


def enable_data_parallel_ema(layer: torch.nn.Module, group: Optional[dist.ProcessGroup] = None,
                             ddp_mean_gradients: bool = True, overlap: bool = False) -> StatsAllReduce:
    """Turn on the statistics all-reduce for an EMA quantizer.

    The returned loss is the loss of the concatenated batch (from the all-reduced SSE / row count).
    The input gradient keeps the local-mean convention, 2*beta*(x-q)/(N_local*D): averaged over
    ranks by DDP (`ddp_mean_gradients=True`) it equals the single-process gradient on the
    concatenated batch; if the caller SUMS gradients across ranks instead, it is divided by the
    world size here."""
    dev = next(layer.parameters()).device
    red = StatsAllReduce(group, overlap=overlap and dev.type == "cuda", device=dev if dev.type == "cuda" else None)
    layer.stats_reduce = red
    layer.grad_scale = 1.0 if ddp_mean_gradients else 1.0 / float(dist.get_world_size(group))
    return red


def broadcast_quantizer_state(layer: torch.nn.Module, src: int = 0,
                              group: Optional[dist.ProcessGroup] = None) -> None:
    """Make every rank start from rank `src`'s codebook / EMA state."""
    with torch.no_grad():
        for t in list(layer.parameters()) + list(layer.buffers()):
            dist.broadcast(t.data, src=src, group=group)
    if hasattr(layer, "invalidate_codebook_cache"):
        layer.invalidate_codebook_cache()
