"""Seeded synthetic workloads: head inputs (logits + label maps) and evaluation pairs (raw-id ground
truth + prediction).  Pure numpy / torch RNG on the CPU; used by bench.py, the profiling scripts and the tests.  There is
no network for datasets, so every measurement and parity test in this repo runs on these generators."""
from __future__ import annotations

import numpy as np
import torch

IGNORE_LABEL = 255

# Cityscapes raw-id -> train-id table ("label2train" of the reference's dataset/cityscapes_list/info.json:3-38; 19 classes)
CITYSCAPES_LABEL2TRAIN = [
    [0, 255], [1, 255], [2, 255], [3, 255], [4, 255], [5, 255], [6, 255], [7, 0], [8, 1], [9, 255],
    [10, 255], [11, 2], [12, 3], [13, 4], [14, 255], [15, 255], [16, 255], [17, 5], [18, 255], [19, 6],
    [20, 7], [21, 8], [22, 9], [23, 10], [24, 11], [25, 12], [26, 13], [27, 14], [28, 15], [29, 255],
    [30, 255], [31, 16], [32, 17], [33, 18], [-1, 255],
]


def synth_head_inputs(B, CK, h, w, H, W, *, C=19, seed=1234, coherent=True, ignore_frac=0.10,
                      class_dist=None, block=32, logit_scale=3.0):
    """logits ~ 3 N(0,1) f32; labels uniform (u) or block-coherent categorical (r); 10 % -> 255."""
    g = torch.Generator().manual_seed(seed)
    logits = logit_scale * torch.randn(B, CK, h, w, generator=g, dtype=torch.float32)
    rng = np.random.default_rng(seed)
    if coherent:
        pdist = np.full(C, 1.0 / C) if class_dist is None else np.asarray(class_dist, dtype=np.float64)
        pdist = pdist / pdist.sum()
        by, bx = (block, block) if np.isscalar(block) else block     # (rows, cols) of a constant block
        bh, bw = -(-H // by) + 1, -(-W // bx) + 1
        coarse = rng.choice(C, size=(B, bh, bw), p=pdist)
        oy, ox = (by // 3, bx // 3) if not np.isscalar(block) else (0, 0)   # tuple blocks are also shifted
        lab = np.repeat(np.repeat(coarse, by, axis=1), bx, axis=2)[:, oy:oy + H, ox:ox + W]
    else:
        lab = rng.integers(0, C, size=(B, H, W))
    lab = lab.astype(np.uint8)
    if ignore_frac > 0:
        lab[rng.random((B, H, W)) < ignore_frac] = IGNORE_LABEL
    return logits, torch.from_numpy(np.ascontiguousarray(lab))


def synth_eval_pair(H, W, *, seed=1234, coherent=True, block=32, n_raw=34, n_pred=19, noise=0.05):
    """(raw-id gt uint8 [H,W] in 0..n_raw-1, pred uint8 [H,W] in 0..n_pred-1).  Coherent maps are constant
    over blocks (gt and pred on DIFFERENT, mutually shifted block grids so their boundaries do not line
    up); ``noise`` = fraction of pred pixels replaced by a random class."""
    rng = np.random.default_rng(seed)
    if coherent:
        by, bx = (block, block) if np.isscalar(block) else block

        def blocks(n, by_, bx_, oy, ox):
            bh, bw = -(-H // by_) + 2, -(-W // bx_) + 2
            m = rng.integers(0, n, size=(bh, bw))
            return np.repeat(np.repeat(m, by_, 0), bx_, 1)[oy:oy + H, ox:ox + W]

        gt = blocks(n_raw, by, bx, 0, 0)
        pr = blocks(n_pred, max(1, (by * 3) // 4), max(1, (bx * 5) // 4), by // 3, bx // 5)
        if noise > 0:
            flip = rng.random((H, W)) < noise
            pr = np.where(flip, rng.integers(0, n_pred, size=(H, W)), pr)
    else:
        gt = rng.integers(0, n_raw, size=(H, W))
        pr = rng.integers(0, n_pred, size=(H, W))
    return np.ascontiguousarray(gt.astype(np.uint8)), np.ascontiguousarray(pr.astype(np.uint8))
