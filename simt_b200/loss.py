"""Drop-in for the reference's utils/loss.py (same class names, ctor and forward signatures)."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib
from .head import _stream_ptr, error_flag, simt_head


class _NLL2dFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, prob, target, ignore):
        lib = _lib.load()
        prob_c = prob.detach().contiguous()
        target = target.contiguous()
        B, C, H, W = prob_c.shape
        dev = prob_c.device
        stats = torch.empty(2, dtype=torch.float64, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        lb = 1 if target.dtype == torch.uint8 else 8
        with torch.cuda.device(dev):
            rc = lib.simt_nll2d_fwd(prob_c.data_ptr(), B, C, H, W, target.data_ptr(), lb, int(ignore),
                                    stats.data_ptr(), loss.data_ptr(), error_flag(dev).data_ptr(), _stream_ptr())
        _lib.check(rc, "simt_nll2d_fwd")
        ctx.ignore = int(ignore)
        ctx.save_for_backward(prob_c, target, stats)
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        lib = _lib.load()
        prob_c, target, stats = ctx.saved_tensors
        B, C, H, W = prob_c.shape
        dev = prob_c.device
        dprob = torch.empty_like(prob_c)
        g = grad_out.detach().to(device=dev, dtype=torch.float32).contiguous()
        lb = 1 if target.dtype == torch.uint8 else 8
        with torch.cuda.device(dev):
            rc = lib.simt_nll2d_bwd(prob_c.data_ptr(), B, C, H, W, target.data_ptr(), lb, ctx.ignore,
                                    stats.data_ptr(), g.data_ptr(), dprob.data_ptr(), _stream_ptr())
        _lib.check(rc, "simt_nll2d_bwd")
        return dprob, None, None


def _weighted_ce2d(predict, target, weight, ignore_label, is_softmax):
    """``CrossEntropy2d.forward(..., weight=w)`` (utils/loss.py:14,36,39): per-class weights, weighted-mean reduction
    (sum_i w[y_i] l_i / sum_i w[y_i], what ``F.cross_entropy`` / ``F.nll_loss(weight=)`` compute).  No call site of the
    reference passes ``weight``, so this rare form is composed from stock torch ops on the tensors' own device (for
    CUDA tensors: torch's CUDA kernels; nothing here runs on the host) instead of a dedicated kernel: the log-probability
    of the label is gathered per pixel, never a [N_valid, C] copy like the reference's boolean-mask gather."""
    valid = (target >= 0) & (target != ignore_label)
    y = torch.where(valid, target, torch.zeros_like(target)).long()
    if is_softmax:
        logp = torch.log_softmax(predict, dim=1)
    else:
        logp = torch.log(predict)
    picked = logp.gather(1, y.unsqueeze(1)).squeeze(1)                  # [n, h, w]
    wy = weight.to(device=predict.device, dtype=predict.dtype)[y] * valid.to(predict.dtype)
    return -(wy * torch.where(valid, picked, torch.zeros_like(picked))).sum() / wy.sum()


class CrossEntropy2d(nn.Module):
    """Masked 2-d cross entropy; mirrors utils/loss.py:6-40 of the reference.

    ``is_softmax=True``: ``predict`` holds logits (-> the fused head kernel with an identity
    resize and T = I).  ``is_softmax=False``: ``predict`` holds probabilities, the loss is
    ``-mean log predict[y]`` over valid pixels (``Tseg_loss`` at tools/trainV2_simt.py:304,408-409).
    Valid = ``target >= 0`` and ``target != ignore_label``; mean over valid pixels; all-ignored
    gives NaN like the reference.  ``weight`` (per-class, weighted mean; unused by the reference's callers) takes a
    torch-op composition on the device instead of the fused kernels.  CUDA only.
    """

    def __init__(self, size_average=True, ignore_label=255, is_softmax=True):
        super().__init__()
        self.size_average = size_average
        self.ignore_label = ignore_label
        self.is_softmax = is_softmax

    def forward(self, predict, target, weight=None):
        assert not target.requires_grad
        assert predict.dim() == 4
        assert target.dim() == 3
        assert predict.size(0) == target.size(0), "{0} vs {1} ".format(predict.size(0), target.size(0))
        assert predict.size(2) == target.size(1), "{0} vs {1} ".format(predict.size(2), target.size(1))
        assert predict.size(3) == target.size(2), "{0} vs {1} ".format(predict.size(3), target.size(2))
        if not predict.is_cuda:
            raise RuntimeError("simt_b200.CrossEntropy2d runs on CUDA (sm_100a) only; there is no CPU fallback")
        if weight is not None:
            return _weighted_ce2d(predict, target, weight, self.ignore_label, self.is_softmax)
        if target.dtype not in (torch.uint8, torch.int64):
            target = target.long()
        if predict.dtype != torch.float32:
            raise TypeError("predict must be float32")
        if self.is_softmax:
            return simt_head(predict, None, target, (predict.size(2), predict.size(3)), self.ignore_label)
        return _NLL2dFn.apply(predict, target, self.ignore_label)


class EntropyLoss(nn.Module):
    """utils/loss.py:42-49.  Instantiated (tools/trainV2_simt.py:306) but never called by the
    reference; kept as a plain torch expression for import compatibility only."""

    def forward(self, x):
        b = torch.softmax(x, dim=1) * torch.log_softmax(x, dim=1)
        return (-1.0 * b.sum(1)).mean()
