"""``Placeholder_loss`` of the reference (tools/trainV2_simt.py:202-230, called at :398-399), fused.

The reference builds, on the UPSAMPLED prediction [B, C+K, H, W]: an arg-max, a one-hot of it, ``predict`` (the
arg-max logit replaced by a constant), a soft-max for the confidence threshold, ``predict_open`` and two
``CrossEntropyLoss`` passes -- about ten full-resolution fp32 temporaries forward and as many backward.  Here one
launch of the head kernel (``simt_placeholder_fwdbwd``) takes the LOW-res logits, applies the bilinear upsample of
:371-372 in registers, derives both label maps per pixel, and writes the gradient at low resolution:

    Place_loss = Placeholder_loss(pred1_lo, num_classes, open_classes, thres=args.Threshold_high,
                                  out_size=(H, W), lambda_place=args.lambda_Place)

PyTorch is used for device memory and streams only; there is no CPU path.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib
from .head import _stream_ptr, _workspace, head_scale


class _PlaceholderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, C, out_size, thres, lambda_place, group):
        lib = _lib.load()
        if not (isinstance(logits, torch.Tensor) and logits.is_cuda):
            raise RuntimeError("simt_b200 runs on CUDA (sm_100a) only: logits must be a CUDA tensor; there is no CPU fallback")
        if logits.dtype != torch.float32 or logits.dim() != 4:
            raise TypeError("logits must be float32 [B, C+K, h, w]")
        x = logits.detach().contiguous()
        B, CK, h, w = x.shape
        H, W = out_size
        dev = x.device
        ws = _workspace(dev, lib.simt_head_workspace_bytes(B, CK, C, h, w, H, W))
        stats = torch.empty(2, dtype=torch.float64, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        dl = torch.empty_like(x)
        with torch.cuda.device(dev):
            rc = lib.simt_placeholder_fwdbwd(x.data_ptr(), B, CK, h, w, C, H, W, float(thres), float(lambda_place),
                                             dl.data_ptr(), stats.data_ptr(), loss.data_ptr(), ws.data_ptr(),
                                             ws.numel(), _stream_ptr())
        _lib.check(rc, "simt_placeholder_fwdbwd")
        if group is not None:   # batch-sharded: mean over the GLOBAL valid-pixel count
            import torch.distributed as dist
            dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
            loss = (stats[0] / stats[1]).to(torch.float32)
        ctx.consumed = False
        ctx.save_for_backward(stats, dl)
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        stats, dl = ctx.saved_tensors
        if ctx.consumed:
            raise RuntimeError("Placeholder_loss: backward a second time (the raw gradient buffer is scaled in place)")
        ctx.consumed = True
        dlog, _ = head_scale(dl, stats, 0, 0, grad_out, want_dT=False)
        return dlog, None, None, None, None, None


def Placeholder_loss(pred_lo: torch.Tensor, num_classes: int, open_classes: int, thres: Optional[float] = None, *,
                     out_size: Tuple[int, int], lambda_place: float = 0.1, group=None) -> torch.Tensor:
    """Drop-in for ``Placeholder_loss(pred, num_classes, open_classes, thres)`` (trainV2_simt.py:202-230) with two
    differences in the call: ``pred_lo`` is the model's LOW-res output (the upsample of :371-372 happens inside) with
    ``out_size=(H, W)``, and ``lambda_place`` is passed explicitly instead of being read from the global ``args``
    (:230; default 0.1 as at :63).  Returns ``loss_known + lambda_place * loss_unknown``; the gradient reaches
    ``pred_lo``.  NaN when no pixel qualifies, like the reference's mean over nothing."""
    if pred_lo.dim() != 4 or pred_lo.size(1) != num_classes + open_classes:
        raise ValueError(f"pred_lo must be [B, {num_classes + open_classes}, h, w], got {tuple(pred_lo.shape)}")
    t = -1.0 if thres is None else float(thres)
    if thres is not None and t < 0:
        raise ValueError("thres must be >= 0 (or None)")
    return _PlaceholderFn.apply(pred_lo, int(num_classes), (int(out_size[0]), int(out_size[1])), t,
                                float(lambda_place), group)
