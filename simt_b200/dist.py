"""Batch-sharded (one process per GPU) plumbing for the head and the eval histograms.

The path shards naturally by batch (pixels are independent given T), so there is NO data-path
collective; the only exchange per training step is ONE all-reduce(sum) of the 2 + CK*C float64
`stats` buffer the fused kernel writes ({sum -log q, n_valid, raw dT}, 2.9 KB at CK = C = 19),
and ONE int64 all-reduce of the confusion matrix at the end of an evaluation.  dLogits stay local
(they feed the local backbone replica).  Works with any torch.distributed backend: NCCL over
NVLink on the GPU box, gloo in the CPU tests of this host logic.
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of n items for `rank`; the first n % world ranks get one extra."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(x: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    lo, hi = shard_range(x.size(0), rank, world)
    return x[lo:hi]


def reduce_head_stats(stats: torch.Tensor, group=None) -> torch.Tensor:
    """In-place all-reduce(sum) of a head `stats` buffer (f64 [2 + CK*C]); returns it."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    return stats


def finish_head(stats: torch.Tensor, CK: int, C: int, grad_out: float = 1.0):
    """(loss, dT [CK, C]) from (all-reduced) stats: loss = sum / n_valid, dT = raw dT * grad_out / n_valid.
    The same arithmetic as simt_head_scale, in torch, for host-side checks."""
    n = stats[1]
    loss = stats[0] / n
    dT = (stats[2:2 + CK * C] * (grad_out / n)).reshape(CK, C)
    return loss, dT


def reduce_hist(hist: torch.Tensor, group=None) -> torch.Tensor:
    """In-place all-reduce(sum) of an int64 histogram; exact (integer addition is associative)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(hist, op=dist.ReduceOp.SUM, group=group)
    return hist
