"""T's volume / anchor / convex regularisers (tools/trainV2_simt.py:326-339,375-384,412-424).

Two forms:
  * ``t_regularizers(T, W)``: ONE launch of a single-CTA kernel (csrc/reg.cu) that returns the
    convex and volume terms of one head together with dT and dW -- replaces ~11 tiny eager
    kernels + a cuSOLVER LU + the host sync at ``torch.isinf`` (:420) per head.
  * ``convex_loss`` / ``volume_loss`` / ``w_fit_loss``: the same expressions written with
    differentiable torch ops (a few hundred elements; SURVEY section 8 rows a8/a9 allow either), for
    callers that want autograd through ``sig_NTM`` / ``sig_W`` without the custom Function.
``anchor_loss`` uses the anchor-statistics kernel (per-channel arg-max pixel and the set of
per-pixel arg-max classes of the UPSAMPLED logits, computed from the low-res logits without
materialising the upsampled tensor).
"""
from __future__ import annotations

from typing import Sequence

import torch

from . import _lib
from .head import _stream_ptr


def convex_loss(W_list: Sequence[torch.Tensor], T_list: Sequence[torch.Tensor]) -> torch.Tensor:
    """``0. - sum_heads MSELoss(sum)(W.mm(T), 0)`` (trainV2_simt.py:412-415)."""
    tot = None
    for W, T in zip(W_list, T_list):
        v = W.mm(T).pow(2).sum()
        tot = v if tot is None else tot + v
    return 0.0 - tot


def w_fit_loss(W_list, T_list) -> torch.Tensor:
    """Objective of the 10-step inner W optimisation (trainV2_simt.py:327-339)."""
    return -convex_loss(W_list, T_list)


def volume_loss(T_list: Sequence[torch.Tensor]):
    """``sum_heads log sqrt |det(T^T T)|``; inf/nan -> 0. (trainV2_simt.py:417-421).
    ``0.5 * logabsdet`` is the same number without the det under/overflow of the reference's
    literal ``log(sqrt(abs(det)))``; the inf/nan guard is applied on the device (no host sync)."""
    tot = None
    for T in T_list:
        v = 0.5 * torch.linalg.slogdet(T.transpose(1, 0).mm(T))[1]
        tot = v if tot is None else tot + v
    return torch.where(torch.isfinite(tot), tot, torch.zeros_like(tot))


def t_regularizers(*args, **kwargs):
    raise NotImplementedError("fused regulariser kernel: see csrc/reg.cu")


def anchor_loss(*args, **kwargs):
    raise NotImplementedError("anchor statistics kernel: see csrc/reg.cu")
