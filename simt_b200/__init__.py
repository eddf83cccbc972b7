"""simt_b200 -- B200-native (sm_100a) implementation of SimT's per-pixel training/eval head.

Host side mirrors the reference's Python call surface for this one path; the arithmetic is in
hand-written CUDA behind a C ABI (include/simt_b200.h, libsimt_b200.so).  See DESIGN.md.
"""
from .head import HeadRunner, HostPrefetcher, SimTHead, check_errors, simt_head
from .hist import ConfusionMeter, build_lut, eval_argmax, fast_hist, label_mapping, per_class_iu
from .loss import CrossEntropy2d, EntropyLoss
from .ntm import sig_NTM, sig_W
from .placeholder import Placeholder_loss
from .regularizers import (anchor_loss, anchor_stats, bilinear_gather, convex_loss, fit_w, pseudo_labels,
                           t_regularizers, volume_loss, w_fit, w_fit_loss)

__all__ = [
    "simt_head", "SimTHead", "HeadRunner", "HostPrefetcher", "check_errors", "CrossEntropy2d", "EntropyLoss", "sig_NTM", "sig_W",
    "fast_hist", "per_class_iu", "label_mapping", "ConfusionMeter", "build_lut", "eval_argmax",
    "convex_loss", "volume_loss", "anchor_loss", "w_fit_loss", "w_fit", "fit_w", "t_regularizers", "anchor_stats", "bilinear_gather", "pseudo_labels",
    "Placeholder_loss",
]
