"""Integer eval histograms on the GPU, behind the reference's function names.

    fast_hist(a, b, n)            tools/compute_iou.py:9-11 (dup tools/evaluate_cityscapes.py:81-83)
    fast_hist(a, b, n33, n19)     tools/compute_ConfusionMatrix.py:54-56
    fast_hist(a, n)               tools/compute_ClassDistribution.py:52-54
    per_class_iu(hist)            tools/compute_iou.py:14-15
    label_mapping(input, mapping) tools/compute_iou.py:18-22

numpy in -> numpy out (int64), like the reference; CUDA tensors in -> CUDA int64 tensor out with no
host sync.  ``ConfusionMeter`` is the accumulate-over-a-dataset form (``hist += fast_hist(...)``,
compute_iou.py:35,51) that keeps the table on the device and fuses ``label_mapping`` as a LUT.
All counting is done by libsimt_b200.so; results are bit-exact.  No CPU fallback.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .head import _stream_ptr, check_errors, error_flag


def _dev() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("simt_b200 histograms run on CUDA (sm_100a) only; there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _to_dev(x, dev):
    """uint8 stays uint8 (fast path); every other integer / bool / float-valued-integer input
    goes through int64, the dtype the reference's ``.astype(int)`` produces."""
    if isinstance(x, np.ndarray):
        if x.dtype == np.uint8:
            t = torch.from_numpy(np.ascontiguousarray(x))
        else:
            t = torch.from_numpy(np.ascontiguousarray(x).astype(np.int64, copy=False))
        return t.to(dev, non_blocking=True).reshape(-1)
    if isinstance(x, torch.Tensor):
        if x.dtype not in (torch.uint8, torch.int64):
            x = x.to(torch.int64)
        return x.to(dev).contiguous().reshape(-1)
    raise TypeError(f"expected numpy array or torch tensor, got {type(x)}")


def build_lut(mapping) -> np.ndarray:
    """256-entry uint8 table equivalent to the sequential ``output[input == k] = v`` passes of
    label_mapping on a uint8 image (matches are against the ORIGINAL input, so the passes
    commute; keys outside 0..255, e.g. -1, can never match a uint8 pixel)."""
    lut = np.arange(256, dtype=np.uint8)
    for k, v in np.asarray(mapping).reshape(-1, 2).tolist():
        if 0 <= k <= 255:
            lut[k] = np.uint8(v)      # same wrap as assigning v into the uint8 copy
    return lut


def confusion_into(hist: torch.Tensor, a: torch.Tensor, b: torch.Tensor, n_rows: int, n_cols: int,
                   lut: torch.Tensor = None) -> None:
    """hist[n_cols * a' + b] += 1 on the device (a' = lut[a]); tensors must already be flat CUDA."""
    lib = _lib.load()
    if a.numel() != b.numel():
        raise ValueError(f"a has {a.numel()} elements, b has {b.numel()}")
    dev = a.device
    with torch.cuda.device(dev):
        rc = lib.simt_confusion(a.data_ptr(), a.element_size(), b.data_ptr(), b.element_size(), a.numel(),
                                None if lut is None else lut.data_ptr(), int(n_rows), int(n_cols),
                                hist.data_ptr(), error_flag(dev).data_ptr(), _stream_ptr())
    _lib.check(rc, "simt_confusion")


def class_hist_into(hist: torch.Tensor, a: torch.Tensor, n_bins: int) -> None:
    lib = _lib.load()
    with torch.cuda.device(a.device):
        rc = lib.simt_class_hist(a.data_ptr(), a.element_size(), a.numel(), int(n_bins), hist.data_ptr(),
                                 _stream_ptr())
    _lib.check(rc, "simt_class_hist")


def fast_hist(a, *args):
    """The reference's three ``fast_hist`` arities (see module docstring)."""
    as_numpy = isinstance(a, np.ndarray)
    dev = a.device if isinstance(a, torch.Tensor) and a.is_cuda else _dev()
    if len(args) == 1:                      # fast_hist(a, n)
        n = int(args[0])
        hist = torch.zeros(n, dtype=torch.int64, device=dev)
        class_hist_into(hist, _to_dev(a, dev), n)
    elif len(args) in (2, 3):               # fast_hist(a, b, n) / fast_hist(a, b, n_rows, n_cols)
        b = args[0]
        n_rows = int(args[1])
        n_cols = int(args[2]) if len(args) == 3 else n_rows
        hist = torch.zeros(n_rows, n_cols, dtype=torch.int64, device=dev)
        confusion_into(hist, _to_dev(a, dev), _to_dev(b, dev), n_rows, n_cols)
    else:
        raise TypeError("fast_hist(a, n) | fast_hist(a, b, n) | fast_hist(a, b, n_rows, n_cols)")
    if as_numpy:
        check_errors(dev)                   # numpy raises for an out-of-table index; so do we
        return hist.cpu().numpy()
    return hist


def per_class_iu(hist):
    """diag / (rowsum + colsum - diag) in float64 on the host (19 numbers; 0/0 -> NaN as in numpy)."""
    if isinstance(hist, torch.Tensor):
        hist = hist.detach().cpu().numpy()
    hist = np.asarray(hist)
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.diag(hist) / (hist.sum(1) + hist.sum(0) - np.diag(hist))


def _label_mapping_generic(x: torch.Tensor, mapping) -> torch.Tensor:
    """tools/compute_iou.py:18-22 for any integer dtype: ``out = copy(input); out[input == k] = v`` for every (k, v) in
    order (matches are on the ORIGINAL input; values that are no key stay as they are), int64 result.  Stock torch ops
    on the tensor's own device -- the LUT kernel below covers the uint8 images PIL yields."""
    out = x.to(torch.int64).clone()
    for k, v in mapping:
        out[x == int(k)] = int(v)
    return out


def label_mapping(input, mapping):
    """raw-id image -> int64 train-id image (tools/compute_iou.py:18-22): uint8 images through the 256-entry LUT
    kernel, any other integer dtype through a torch-op composition on the device."""
    lib = _lib.load()
    as_numpy = isinstance(input, np.ndarray)
    dev = input.device if isinstance(input, torch.Tensor) and input.is_cuda else _dev()
    is_u8 = (input.dtype == np.uint8) if as_numpy else (input.dtype == torch.uint8)
    if not is_u8:
        kind = input.dtype.kind if as_numpy else ("i" if not (input.dtype.is_floating_point or input.dtype == torch.bool) else "f")
        if kind not in ("i", "u"):
            raise TypeError("simt_b200.label_mapping takes integer images")
        xt = torch.from_numpy(np.ascontiguousarray(input).astype(np.int64)) if as_numpy else input
        out = _label_mapping_generic(xt.to(dev), [(int(a), int(b)) for a, b in np.asarray(mapping).tolist()])
        return out.cpu().numpy() if as_numpy else out
    shape = tuple(input.shape)
    x = _to_dev(input, dev)
    lut = torch.from_numpy(build_lut(mapping)).to(dev)
    out = torch.empty(x.numel(), dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        rc = lib.simt_label_map(x.data_ptr(), x.numel(), lut.data_ptr(), out.data_ptr(), _stream_ptr())
    _lib.check(rc, "simt_label_map")
    out = out.reshape(shape)
    return out.cpu().numpy() if as_numpy else out


class ConfusionMeter:
    """Device-resident ``hist += fast_hist(label_mapping(gt), pred, n)`` over a dataset
    (tools/compute_iou.py:35-58; rectangular tables for compute_ConfusionMatrix.py:89-98)."""

    def __init__(self, n_rows: int, n_cols: int = None, mapping=None, device=None):
        self.n_rows = int(n_rows)
        self.n_cols = int(n_cols) if n_cols is not None else int(n_rows)
        self.device = torch.device(device) if device is not None else _dev()
        self.hist = torch.zeros(self.n_rows, self.n_cols, dtype=torch.int64, device=self.device)
        self.lut = None if mapping is None else torch.from_numpy(build_lut(mapping)).to(self.device)

    def update(self, gt, pred) -> None:
        a, b = _to_dev(gt, self.device), _to_dev(pred, self.device)
        if self.lut is not None and a.dtype != torch.uint8:
            raise TypeError("a raw-id -> train-id mapping needs uint8 ground truth")
        confusion_into(self.hist, a, b, self.n_rows, self.n_cols, self.lut)

    def value(self) -> np.ndarray:
        check_errors(self.device)
        return self.hist.cpu().numpy()

    def per_class_iu(self) -> np.ndarray:
        return per_class_iu(self.value())

    def miou_percent(self) -> float:
        """``round(np.nanmean(mIoUs) * 100, 2)`` (compute_iou.py:58)."""
        return round(float(np.nanmean(self.per_class_iu())) * 100, 2)


def eval_argmax(logits_a: torch.Tensor, logits_b: torch.Tensor, out_size, num_classes: int) -> torch.Tensor:
    """uint8 prediction map [B, H, W] of tools/evaluate_cityscapes.py:127-138: upsample the first
    ``num_classes`` channels of one (``logits_b=None``) or two low-res head-2 outputs to ``out_size``, add,
    arg-max -- on the device, without the reference's two 159 MB device->host copies per image.  Feed it to
    ``ConfusionMeter.update(gt, pred)``."""
    lib = _lib.load()
    if not logits_a.is_cuda:
        raise RuntimeError("simt_b200.eval_argmax runs on CUDA (sm_100a) only; there is no CPU fallback")
    a = logits_a.detach().contiguous().float()
    b = None if logits_b is None else logits_b.detach().contiguous().float()
    B, CKa, ha, wa = a.shape
    H, W = int(out_size[0]), int(out_size[1])
    out = torch.empty(B, H, W, dtype=torch.uint8, device=a.device)
    with torch.cuda.device(a.device):
        rc = lib.simt_eval_argmax(a.data_ptr(), CKa, ha, wa, None if b is None else b.data_ptr(),
                                  0 if b is None else b.shape[1], 0 if b is None else b.shape[2],
                                  0 if b is None else b.shape[3], B, int(num_classes), H, W, out.data_ptr(), _stream_ptr())
    _lib.check(rc, "simt_eval_argmax")
    return out
