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


class CrossEntropy2d(nn.Module):
    """Masked 2-d cross entropy; mirrors utils/loss.py:6-40 of the reference.

    ``is_softmax=True``: ``predict`` holds logits (-> the fused head kernel with an identity
    resize and T = I).  ``is_softmax=False``: ``predict`` holds probabilities, the loss is
    ``-mean log predict[y]`` over valid pixels (``Tseg_loss`` at tools/trainV2_simt.py:304,408-409).
    Valid = ``target >= 0`` and ``target != ignore_label``; mean over valid pixels; all-ignored
    gives NaN like the reference.  CUDA only.
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
        if weight is not None:
            raise NotImplementedError("simt_b200.CrossEntropy2d: per-class `weight` is not on the accelerated path "
                                      "(no call site in the reference passes it)")
        if not predict.is_cuda:
            raise RuntimeError("simt_b200.CrossEntropy2d runs on CUDA (sm_100a) only; there is no CPU fallback")
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
