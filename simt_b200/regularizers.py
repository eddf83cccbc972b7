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


class _TRegFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, T, W):
        lib = _lib.load()
        if not T.is_cuda:
            raise RuntimeError("simt_b200.t_regularizers runs on CUDA (sm_100a) only; there is no CPU fallback")
        Tc = T.detach().contiguous().float()
        Wc = None if W is None else W.detach().contiguous().float()
        CK, C = Tc.shape
        dev = Tc.device
        out = torch.empty(2, dtype=torch.float32, device=dev)
        dT_c = torch.empty_like(Tc)
        dT_v = torch.empty_like(Tc)
        dW_c = torch.empty(CK, CK, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            rc = lib.simt_t_regularizers(Tc.data_ptr(), None if Wc is None else Wc.data_ptr(), CK, C, out.data_ptr(),
                                         dT_c.data_ptr(), dT_v.data_ptr(), dW_c.data_ptr(), _stream_ptr())
        _lib.check(rc, "simt_t_regularizers")
        ctx.has_W = W is not None
        ctx.save_for_backward(dT_c, dT_v, dW_c)
        return out[0], out[1]

    @staticmethod
    def backward(ctx, g_convex, g_volume):
        dT_c, dT_v, dW_c = ctx.saved_tensors
        dT = g_volume * dT_v
        dW = None
        if ctx.has_W:
            dT = dT + g_convex * dT_c
            dW = g_convex * dW_c
        return dT, dW


def t_regularizers(T: torch.Tensor, W: torch.Tensor = None):
    """(convex, volume) of ONE head in one single-CTA launch, differentiable w.r.t. T and W:
    ``convex = -||W T||_F^2`` (trainV2_simt.py:414-415), ``volume = log sqrt|det(T^T T)|`` with the
    inf/nan -> 0 guard applied on the device (:417-421).  Sum the heads' terms on the caller side."""
    return _TRegFn.apply(T, W)


def anchor_stats(logits_lo: torch.Tensor, out_size):
    """(Anchor_index int64[CK], exist bool[CK]) of trainV2_simt.py:375-377 from LOW-res logits
    [B, CK, h, w]: per channel the arg-max pixel (NHWC-flatten order over B*H*W) of the upsampled
    logit, and which classes are the per-pixel arg-max somewhere.  B > 1 = pixels flattened over the
    batch (the reference's ``.view`` only works for B = 1)."""
    lib = _lib.load()
    if not logits_lo.is_cuda:
        raise RuntimeError("simt_b200.anchor_stats runs on CUDA (sm_100a) only; there is no CPU fallback")
    x = logits_lo.detach().contiguous().float()
    B, CK, h, w = x.shape
    H, W = int(out_size[0]), int(out_size[1])
    dev = x.device
    idx = torch.empty(CK, dtype=torch.int64, device=dev)
    mask = torch.empty(1, dtype=torch.int64, device=dev)
    scratch = torch.empty(CK, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        rc = lib.simt_anchor_stats(x.data_ptr(), B, CK, h, w, H, W, idx.data_ptr(), None, mask.data_ptr(),
                                   scratch.data_ptr(), _stream_ptr())
    _lib.check(rc, "simt_anchor_stats")
    exist = ((mask >> torch.arange(CK, device=dev)) & 1).bool()
    return idx, exist


def bilinear_gather(src_lo: torch.Tensor, pixel_idx: torch.Tensor, out_size) -> torch.Tensor:
    """rows[r] = upsample(src_lo)[b, :, Y, X] at flat pixel ``pixel_idx[r]`` (no upsampled tensor)."""
    lib = _lib.load()
    x = src_lo.detach().contiguous().float()
    B, C, h, w = x.shape
    H, W = int(out_size[0]), int(out_size[1])
    idx = pixel_idx.contiguous().to(torch.int64)
    rows = torch.empty(idx.numel(), C, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        rc = lib.simt_bilinear_gather(x.data_ptr(), B, C, h, w, H, W, idx.data_ptr(), idx.numel(), rows.data_ptr(),
                                      _stream_ptr())
    _lib.check(rc, "simt_bilinear_gather")
    return rows


def anchor_loss(pred_lo_list: Sequence[torch.Tensor], T_list: Sequence[torch.Tensor], fixed_logits_lo: torch.Tensor,
                out_size) -> torch.Tensor:
    """``sum_heads MSELoss(sum)(T[Exist], labelC_flat[Anchor_index][Exist])`` (trainV2_simt.py:354-357,
    375-384), differentiable w.r.t. T.  ``pred_lo_list``: the heads' LOW-res logits; ``fixed_logits_lo``:
    the frozen model's low-res output2 (its softmax is upsampled only at the CK anchor pixels)."""
    probs_lo = torch.softmax(fixed_logits_lo.detach(), dim=1)            # :354 (softmax before the upsample)
    tot = None
    for pred_lo, T in zip(pred_lo_list, T_list):
        a_idx, exist = anchor_stats(pred_lo, out_size)                   # :375-377
        anchor = bilinear_gather(probs_lo, a_idx, out_size)              # :378
        diff = (T - anchor) * exist.to(T.dtype).unsqueeze(1)             # rows in Exist_label only
        v = diff.pow(2).sum()                                            # :379
        tot = v if tot is None else tot + v
    return tot


def pseudo_labels(fixed_logits_lo: torch.Tensor, pred2_lo: torch.Tensor, out_size, num_classes: int,
                  thres_high: float = 0.8, thres_low: float = 0.2) -> torch.Tensor:
    """``Conf_label_target`` of tools/trainV2_simt.py:354-365 + :387-393 as uint8 [B, H, W] on the device, from
    the LOW-res outputs of the frozen model (``output2``, closed-set C channels) and of the student
    (``pred2`` before the upsample, CK channels; may be None).  No high-res tensor, no host round trip."""
    lib = _lib.load()
    if not fixed_logits_lo.is_cuda:
        raise RuntimeError("simt_b200.pseudo_labels runs on CUDA (sm_100a) only; there is no CPU fallback")
    x = fixed_logits_lo.detach().contiguous().float()
    B, C, h, w = x.shape
    if C != num_classes:
        raise ValueError(f"fixed_logits_lo has {C} channels, num_classes = {num_classes}")
    H, W = int(out_size[0]), int(out_size[1])
    p2 = None if pred2_lo is None else pred2_lo.detach().contiguous().float()
    CK = C if p2 is None else p2.shape[1]
    scratch = torch.empty_like(x)
    out = torch.empty(B, H, W, dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        rc = lib.simt_pseudo_labels(x.data_ptr(), None if p2 is None else p2.data_ptr(), B, C, CK, h, w, H, W,
                                    float(thres_high), float(thres_low), scratch.data_ptr(), out.data_ptr(),
                                    _stream_ptr())
    _lib.check(rc, "simt_pseudo_labels")
    return out
