"""Batch-sharded (one process per GPU) plumbing for the head and the eval histograms.

The path shards naturally by batch (pixels are independent given T), so there is NO data-path
collective; the only exchange per training step is ONE all-reduce(sum) of the 2 + CK*C float64
`stats` buffer the fused kernel writes ({sum -log q, n_valid, raw dT}, 2.9 KB at CK = C = 19),
and ONE int64 all-reduce of the confusion matrix at the end of an evaluation.  dLogits stay local
(they feed the local backbone replica).  Works with any torch.distributed backend: NCCL over
NVLink on the GPU box, gloo in the CPU tests of this host logic.

On one node the per-step exchange does not go through a library collective at all: ``PeerMailbox`` sets up CUDA-IPC
mapped mailboxes between the ranks and the step's own kernels exchange the valid counts and the stats over NVLink peer
stores (csrc/xchg.cu, ``simt_head_step_sharded``); torch.distributed only carries the 64-byte handles once.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Tuple

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


class PeerMailbox:
    """CUDA-IPC mailboxes for the fused stats exchange of the sharded head (one per HeadRunner / buffer set).

    Collective constructor: every rank of ``group`` (all on ONE node, <= 8) creates its mailbox, the 64-byte IPC
    handles are all-gathered and every peer's mailbox is opened.  ``ptrs`` is the ctypes array
    ``simt_head_step_sharded`` takes.  Raises ``RuntimeError`` when peer mapping is not possible (other node, no
    P2P); callers fall back to ``reduce_head_stats`` (a library all-reduce).
    """

    def __init__(self, C: int, group=None, device: Optional[torch.device] = None):
        from . import _lib
        self.lib = _lib.load()
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world > 8:
            raise RuntimeError("PeerMailbox: at most 8 ranks (one node)")
        self.dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.own = ctypes.c_void_p()
        self.peers = {}
        handle = ctypes.create_string_buffer(64)
        ok, why = True, ""
        with torch.cuda.device(self.dev):
            rc = self.lib.simt_xchg_create(self.lib.simt_xchg_bytes(int(C)), ctypes.byref(self.own), handle)
            if rc:
                ok, why = False, f"simt_xchg_create: code {rc}"
            handles = [None] * self.world
            dist.all_gather_object(handles, (ok, bytes(handle.raw)), group=group)
            ptrs = (ctypes.c_void_p * self.world)()
            if ok and all(h[0] for h in handles):
                for r, (_, hb) in enumerate(handles):
                    if r == self.rank:
                        ptrs[r] = self.own.value
                        continue
                    p = ctypes.c_void_p()
                    rc = self.lib.simt_xchg_open(hb, ctypes.byref(p))
                    if rc:
                        ok, why = False, f"simt_xchg_open(rank {r}): code {rc}"
                        break
                    self.peers[r] = p
                    ptrs[r] = p.value
            else:
                ok = False
            # agree on the outcome: one rank failing means nobody uses the mailboxes
            flags = [None] * self.world
            dist.all_gather_object(flags, ok, group=group)
        if not all(flags):
            self.close()
            raise RuntimeError("PeerMailbox: peer mapping unavailable" + (f" ({why})" if why else ""))
        self.ptrs = ptrs

    def close(self):
        """Unmap the peers' mailboxes and free this rank's.  Peers that keep stepping afterwards wait for this rank
        until their bound expires (simt_xchg_set_timeout), so close after the last step on every rank."""
        with torch.cuda.device(self.dev):
            torch.cuda.synchronize(self.dev)
            for p in self.peers.values():
                self.lib.simt_xchg_close(p)
            self.peers = {}
            if self.own:
                self.lib.simt_xchg_destroy(self.own)
                self.own = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
