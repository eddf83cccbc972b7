"""CPU oracle for the SimT per-pixel head and the integer eval histograms.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it.  Nothing under ``simt_b200/`` imports it and
the product path raises if the CUDA library is missing.

It restates, op for op, the reference's algorithm for the hot path using the
same third-party primitives the reference calls (PyTorch CPU ops, numpy
bincount).  The arithmetic itself lives in un-vendored third-party code
(PyTorch -- the reference only says "Pytorch 1.3 & 1.7 are ok", README.md:18-21;
here torch 2.11.0, numpy 2.3.5), so what is restated is the reference's
*composition* of those primitives; each function cites the reference lines it
follows (paths relative to the reference root).

Parity pinning: the reference ships NO tests, golden vectors or fixtures for
this path.  The oracle is therefore pinned against outputs of
the reference itself: ``oracle/make_golden.py`` imports the unmodified
reference modules (``utils/loss.py``, ``model/deeplab_multi.py``,
``tools/compute_iou.py``) in the build container, runs them on seeded
synthetic inputs and commits inputs + outputs under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks this file against those vectors
bit-for-bit (fp32 and fp64) on CPU.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

IGNORE_LABEL = 255


# --------------------------------------------------------------------------
# a1: bilinear resize, align_corners=True
# --------------------------------------------------------------------------
def upsample_bilinear_ac(x: torch.Tensor, size) -> torch.Tensor:
    """``nn.Upsample(size=(H, W), mode='bilinear', align_corners=True)``.

    Reference: tools/trainV2_simt.py:301 (``interp_target``), applied at
    :371-372 and again (identity size) at :402,405.  ``nn.Upsample.forward``
    is ``F.interpolate`` with the same arguments.
    """
    return F.interpolate(x, size=tuple(size), mode="bilinear", align_corners=True)


# --------------------------------------------------------------------------
# a4: CrossEntropy2d
# --------------------------------------------------------------------------
def cross_entropy_2d(predict: torch.Tensor, target: torch.Tensor, *, is_softmax: bool,
                     ignore_label: int = IGNORE_LABEL, weight=None) -> torch.Tensor:
    """Masked mean CE / NLL over valid pixels.

    Reference: utils/loss.py:14-40.  Mask ``(t >= 0) & (t != ignore)`` (:29),
    labels gathered by the mask (:30), NCHW -> NHWC contiguous copy (:33),
    rows gathered by the same mask (:34), then ``F.cross_entropy`` on logits
    (:36) or ``log`` + ``F.nll_loss`` on probabilities (:38-39), mean
    reduction.  The early return at :31-32 is unreachable for a 3-d target
    (an empty selection is still 1-d) and is therefore not restated; an
    all-ignored target gives NaN exactly as the reference does.
    """
    assert not target.requires_grad
    assert predict.dim() == 4 and target.dim() == 3
    assert predict.size(0) == target.size(0)
    assert predict.size(2) == target.size(1)
    assert predict.size(3) == target.size(2)
    n, c, h, w = predict.size()
    valid = (target >= 0) * (target != ignore_label)
    picked = target[valid]
    rows = predict.transpose(1, 2).transpose(2, 3).contiguous()
    rows = rows[valid.view(n, h, w, 1).repeat(1, 1, 1, c)].view(-1, c)
    if is_softmax:
        return F.cross_entropy(rows, picked, weight=weight, reduction="mean")
    return F.nll_loss(torch.log(rows), picked, weight=weight, reduction="mean")


# --------------------------------------------------------------------------
# a1-a4 composed: the T-corrected head exactly as the training loop runs it
# --------------------------------------------------------------------------
def simt_head_loss(logits_lo: torch.Tensor, T: torch.Tensor, labels: torch.Tensor, out_size,
                   ignore_label: int = IGNORE_LABEL, ce=None) -> torch.Tensor:
    """loss_y = Tseg_loss(mm(softmax(interp(interp(pred))), T), label).

    Reference: tools/trainV2_simt.py:371-372 (first upsample), :402-403 /
    :405-406 (second, same-size upsample; channel softmax; NHWC flatten;
    ``torch.mm`` with T [CK, C]; reshape back to NCHW), :408-409
    (``CrossEntropy2d(is_softmax=False)``, built at :304).  ``ce``: the reference's own
    ``CrossEntropy2d(is_softmax=False)`` instance (oracle/_ref/loss.py, see oracle/make_ref.py) to
    use for :408 instead of the restatement above.
    """
    B, CK = logits_lo.shape[:2]
    H, W = out_size
    C = T.shape[1]
    up = upsample_bilinear_ac(logits_lo, (H, W))                       # :371-372
    p = torch.softmax(upsample_bilinear_ac(up, (H, W)), dim=1)         # :402
    p = p.permute(0, 2, 3, 1).contiguous().view(-1, CK)
    q = torch.mm(p, T).view(B, H, W, C).permute(0, 3, 1, 2)            # :403
    if ce is not None:
        return ce(q, labels)                                                          # :408, the reference's class
    return cross_entropy_2d(q, labels, is_softmax=False, ignore_label=ignore_label)  # :408


def plain_ce_loss(logits_lo: torch.Tensor, labels: torch.Tensor, out_size,
                  ignore_label: int = IGNORE_LABEL) -> torch.Tensor:
    """``seg_loss(interp(pred), label)`` with ``CrossEntropyLoss(ignore_index=255)``.

    Reference: tools/trainV2_simt.py:303,394-395 and tools/trainV1_warmup.py:203,
    218-224 (the T = identity special case of the head).
    """
    up = upsample_bilinear_ac(logits_lo, out_size)
    return F.cross_entropy(up, labels, ignore_index=ignore_label)


def simt_head_fwd_bwd(logits_lo, T, labels, out_size, dtype=torch.float32, ignore_label=IGNORE_LABEL):
    """Run the head forward + autograd backward on CPU; returns (loss, dLogits, dT)."""
    lg = logits_lo.detach().to(dtype).clone().requires_grad_(True)
    Tt = T.detach().to(dtype).clone().requires_grad_(True)
    loss = simt_head_loss(lg, Tt, labels.long(), out_size, ignore_label)
    loss.backward()                                                    # :428
    return loss.detach(), lg.grad.detach(), Tt.grad.detach()


# --------------------------------------------------------------------------
# a6 / a7: the T layer and the convex-hull weight layer (forward maths only)
# --------------------------------------------------------------------------
def sig_ntm_forward(NTM: torch.Tensor, class_dist: torch.Tensor, num_classes: int,
                    open_classes: int = 0) -> torch.Tensor:
    """T = L1-row-normalise(sigmoid(NTM) * tile(ClassDist) + [I_C; 0]).

    Reference: model/deeplab_multi.py:254-257 (buffers), :259-263 (forward).
    ``class_dist`` is the float64[19] npy cast through ``torch.FloatTensor``.
    """
    ck = num_classes + open_classes
    ident = torch.cat([torch.eye(num_classes, num_classes), torch.zeros(open_classes, num_classes)], 0)
    dist = torch.as_tensor(np.tile(np.asarray(class_dist), (ck, 1))).to(torch.float32)
    T = torch.sigmoid(NTM)
    T = T.mul(dist.to(NTM.dtype)) + ident.to(NTM.dtype)
    return F.normalize(T, p=1, dim=1)


def sig_w_forward(weight: torch.Tensor) -> torch.Tensor:
    """W = softmax(weight with diag forced to -1e4, dim=1) - I.

    Reference: model/deeplab_multi.py:277-286.  Mutates ``weight`` in place
    under no_grad like the reference does (:279-281).
    """
    n = weight.shape[0]
    idx = np.diag_indices(n)
    with torch.no_grad():
        weight[idx[0], idx[1]] = -10000.0 * torch.ones(n, dtype=weight.dtype)
    w = torch.softmax(weight, dim=1)
    return (torch.zeros(n, n, dtype=weight.dtype) - torch.eye(n, dtype=weight.dtype)) + w


# --------------------------------------------------------------------------
# a8 / a9 / a10: T regularisers
# --------------------------------------------------------------------------
def convex_loss(W_list, T_list) -> torch.Tensor:
    """NTM_Convex_loss = 0 - sum_heads MSE_sum(W @ T, 0).  trainV2_simt.py:412-415."""
    mse = torch.nn.MSELoss(reduction="sum")
    tot = 0.0
    for W, T in zip(W_list, T_list):
        tot = tot + mse(W.mm(T), torch.zeros_like(T))
    return 0.0 - tot


def w_fit_loss(W_list, T_list) -> torch.Tensor:
    """Inner W-optimisation objective, sum_heads ||W T||_F^2.  trainV2_simt.py:336."""
    return -convex_loss(W_list, T_list)


def adam_step(p, g, m, v, step, lr, betas=(0.9, 0.999), eps=1e-8):
    """One torch.optim.Adam update (weight_decay 0, amsgrad off), in place on p / m / v; ``step`` is the 1-based
    count AFTER this update.  torch (third-party, torch/optim/adam.py ``_single_tensor_adam``) is the reference's
    optimiser at trainV2_simt.py:277-280; this restates its published algorithm."""
    b1, b2 = betas
    m.lerp_(g, 1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    bc1 = 1 - b1 ** step
    bc2 = 1 - b2 ** step
    denom = (v.sqrt() / (bc2 ** 0.5)).add_(eps)
    p.addcdiv_(m, denom, value=-(lr / bc1))


def w_fit_loop(ntm_params, weights, states, lr, class_dist, num_classes, open_classes, rounds=10):
    """The inner W optimisation, trainV2_simt.py:326-339, for any number of heads.

    ntm_params: list of sig_NTM.NTM tensors (not stepped here); weights: list of sig_W.weight tensors, updated in
    place; states: list of dicts {m, v, step} (Adam state of each weight), updated in place.
    Returns (ntm_grads, losses): the gradient the ``rounds`` backward passes of :337 ACCUMULATE into each NTM
    parameter (zeroed once per outer iteration at :317-318), and the objective of :336 at every round."""
    ntm_grads = [torch.zeros_like(p) for p in ntm_params]
    losses = []
    for _ in range(rounds):
        leaves_n = [p.detach().clone().requires_grad_(True) for p in ntm_params]
        leaves_w = [w.detach().clone().requires_grad_(True) for w in weights]
        Ts = [sig_ntm_forward(p, class_dist, num_classes, open_classes) for p in leaves_n]      # :328-329
        Ws = [sig_w_forward_functional(w) for w in leaves_w]                                     # :330-331
        loss = w_fit_loss(Ws, Ts)                                                                # :336
        loss.backward()                                                                          # :337
        losses.append(loss.detach())
        for i, w in enumerate(weights):
            ntm_grads[i] += leaves_n[i].grad
            st = states[i]
            st["step"] += 1
            with torch.no_grad():
                w.copy_(_fill_diag(w))                    # sig_W.forward's in-place diagonal write (:279-281)
                adam_step(w, leaves_w[i].grad, st["m"], st["v"], st["step"], lr)               # :338-339
    return ntm_grads, torch.stack(losses)


def _fill_diag(w):
    out = w.clone()
    out.fill_diagonal_(-10000.0)
    return out


def sig_w_forward_functional(weight: torch.Tensor) -> torch.Tensor:
    """``sig_w_forward`` without the in-place write: the diagonal is masked to -1e4 (no gradient reaches it,
    like the reference, whose softmax output is exactly 0 there)."""
    n = weight.shape[0]
    eye = torch.eye(n, dtype=torch.bool)
    w = torch.softmax(torch.where(eye, torch.full_like(weight, -10000.0), weight), dim=1)
    return w - torch.eye(n, dtype=weight.dtype)


def volume_loss(T_list):
    """sum_heads log sqrt |det(T^T T)|; inf/nan -> python 0.  trainV2_simt.py:417-421."""
    tot = None
    for T in T_list:
        v = torch.log(torch.sqrt(torch.abs(torch.linalg.det(T.transpose(1, 0).mm(T)))))
        tot = v if tot is None else tot + v
    if torch.isinf(tot) or torch.isnan(tot):
        return 0.0
    return tot


def anchor_stats(pred_up: torch.Tensor):
    """(Anchor_index[CK], Exist_label) from upsampled logits.  trainV2_simt.py:375-377.

    The reference's ``.view`` on the permuted tensor only works for B = 1;
    ``reshape`` keeps the B = 1 result and defines B > 1 as "flatten pixels
    over the batch" (SURVEY section 7).
    """
    ck = pred_up.shape[1]
    flat = pred_up.detach().clone().permute(0, 2, 3, 1).reshape(-1, ck)
    return torch.argmax(flat, dim=0), torch.unique(torch.argmax(flat, dim=1))


def anchor_loss(pred_up_list, T_list, labelC_flat) -> torch.Tensor:
    """sum_heads MSE_sum(T[Exist], labelC_flat[Anchor_index][Exist]).  :375-384."""
    mse = torch.nn.MSELoss(reduction="sum")
    tot = 0.0
    for pred_up, T in zip(pred_up_list, T_list):
        a_idx, exist = anchor_stats(pred_up)
        anchor = labelC_flat[a_idx]
        tot = tot + mse(T[exist], anchor[exist])
    return tot


def label_c_flat(fixed_logits_lo: torch.Tensor, out_size) -> torch.Tensor:
    """labelC_flat = interp(softmax(output2)) as [N, C].  trainV2_simt.py:354,357."""
    c = fixed_logits_lo.shape[1]
    lc = upsample_bilinear_ac(torch.softmax(fixed_logits_lo.clone(), dim=1), out_size)
    return lc.permute(0, 2, 3, 1).reshape(-1, c)


# --------------------------------------------------------------------------
# (f) row 2: pseudo-label generation + class-posterior relabel
# --------------------------------------------------------------------------
def pseudo_labels(fixed_logits_lo: torch.Tensor, pred2_up: torch.Tensor, out_size, num_classes: int,
                  thres_high: float = 0.8, thres_low: float = 0.2) -> torch.Tensor:
    """Conf_label_target of tools/trainV2_simt.py:354-365 (threshold the frozen model's upsampled softmax) and
    :387-393 (low-confidence pixels take the student's arg-max if that is an open-set class, else 255).
    ``pred2_up`` is the student's UPSAMPLED head-2 logits [B, CK, H, W] (:372).  Returns int64 [B, H, W]."""
    labelC = upsample_bilinear_ac(torch.softmax(fixed_logits_lo.clone(), dim=1), out_size)            # :354
    labelC_max = torch.max(labelC, 1)                                                                  # :355
    labelC_argmax = torch.argmax(labelC, dim=1).float()                                                # :356
    lab = torch.where(labelC_max[0] > thres_high, labelC_argmax, 255. * torch.ones_like(labelC_argmax))   # :359
    lab = torch.where(labelC_max[0] < thres_low, num_classes * torch.ones_like(labelC_argmax), lab)       # :361
    conf = lab.long()                                                                                  # :362
    pseudo = torch.argmax(pred2_up.clone(), dim=1).detach()                                            # :387
    ones, zeros = torch.ones_like(conf), torch.zeros_like(conf)
    mask = torch.where(conf == num_classes * ones, ones, zeros)                                        # :390
    pseudo1 = mask * pseudo                                                                            # :391
    pseudo1 = torch.where(pseudo1 >= num_classes * ones, pseudo1, 255 * ones)                          # :392
    return torch.where(conf == num_classes * ones, pseudo1, conf)                                      # :393


def eval_two_scale_argmax(out2_a: torch.Tensor, out2_b, out_size, num_classes: int) -> np.ndarray:
    """Prediction map of tools/evaluate_cityscapes.py:127-138 (``evaluate_simt``): head-2 logits of the 1024x512
    pass and of the 1280x640 pass, closed-set channels only, each upsampled to the label size, added in float32
    on the host, arg-max over classes.  ``out2_b`` may be None (``evaluate_warmup``, :186-196).  uint8 [B, H, W]."""
    outs = []
    for i in range(out2_a.shape[0]):
        output = upsample_bilinear_ac(out2_a[i:i + 1, :num_classes], out_size).cpu().data[0].numpy()      # :128
        if out2_b is not None:
            output += upsample_bilinear_ac(out2_b[i:i + 1, :num_classes], out_size).cpu().data[0].numpy()  # :133
        output = output.transpose(1, 2, 0)                                                                 # :137
        outs.append(np.asarray(np.argmax(output, axis=2)))                                                 # :138
    return np.stack(outs).astype(np.uint8)


def placeholder_loss(pred, num_classes, open_classes, thres=None, lambda_place=0.1):
    """``Placeholder_loss`` of tools/trainV2_simt.py:202-230 on the UPSAMPLED logits ``pred`` [B, C+K, H, W].

    Restated per pixel (a = arg-max channel, first on ties):
      * known-class term (:205,212-217): CE(pred, a) on pixels with a < C (and max softmax prob > thres);
      * ``predict`` (:206-209): pred with the arg-max channel replaced by the CONSTANT -0.0 -- ``ones`` at :208 is
        ``zeros_like``, so ``-1000. * ones`` is zero, not -1000 (the reference's quirk, kept);
      * open-set target (:220-223): arg-max over [0]*C ++ predict[C:], i.e. the first best open-set channel when its
        logit is > 0 and class 0 otherwise; 255 wherever the known-class label is 255;
      * unknown term (:229): CE(predict, target); result = known + lambda_place * unknown (:230).
    """
    ck = num_classes + open_classes
    pseudo = torch.argmax(pred, dim=1)                                                    # :205
    onehot = F.one_hot(pseudo, ck).permute(0, 3, 1, 2).to(pred.dtype)                     # :206
    predict = torch.where(onehot > 0, torch.zeros_like(pred), pred)                       # :207-209 (-1000 * 0)
    ones = torch.ones_like(pseudo)
    pseudo1 = torch.where(pseudo < num_classes, pseudo, 255 * ones)                       # :213
    if thres is not None:
        pred_max = torch.softmax(pred.detach(), dim=1).max(1)[0]                          # :215
        pseudo1 = torch.where(pred_max > thres, pseudo1, 255 * ones)                      # :216
    loss_known = F.cross_entropy(pred, pseudo1, ignore_index=IGNORE_LABEL)                # :217
    predict_open = torch.zeros_like(predict)                                              # :220
    predict_open[:, num_classes:] = predict[:, num_classes:].detach()                     # :221
    y = torch.argmax(predict_open, dim=1)                                                 # :222
    y = torch.where(pseudo1 == 255, 255 * ones, y)                                        # :223
    loss_unknown = F.cross_entropy(predict, y, ignore_index=IGNORE_LABEL)                 # :229
    return loss_known + lambda_place * loss_unknown                                       # :230


def placeholder_fwd_bwd(logits_lo, out_size, num_classes, open_classes, thres=None, lambda_place=0.1,
                        dtype=torch.float32):
    """(loss, dLoss/dlogits_lo) of upsample (:371-372) -> Placeholder_loss (:398-399), via autograd."""
    lg = logits_lo.detach().to(dtype).clone().requires_grad_(True)
    loss = placeholder_loss(upsample_bilinear_ac(lg, out_size), num_classes, open_classes, thres, lambda_place)
    loss.backward()
    return loss.detach(), lg.grad.detach()


def training_step_loss(pred1_lo, pred2_lo, fixed_out2_lo, label_target, T1, T2, W1, W2, out_size, num_classes,
                       lambda_seg=0.1, lambda_convex=0.1, lambda_volume=1.0, lambda_anchor=1.0,
                       thres_high=0.8, thres_low=0.2, lambda_place=None):
    """The head part of one training iteration, tools/trainV2_simt.py:351-424, WITHOUT the backbone: pseudo labels
    (:351-365), upsample (:371-372), anchor (:375-384), class-posterior relabel + seg_loss (:387-395),
    Placeholder_loss (:398-399, only when ``lambda_place`` is given), T-corrected losses (:402-409), convex /
    volume (:412-421), combination (:423-424) with the published weights of sh_simt.sh:16.  B must be 1 (anchor)."""
    up = lambda x: upsample_bilinear_ac(x, out_size)
    labelC_flat = label_c_flat(fixed_out2_lo, out_size)
    pred1, pred2 = up(pred1_lo), up(pred2_lo)                                                  # :371-372
    anchor = anchor_loss([pred1, pred2], [T1, T2], labelC_flat)                                # :375-384
    conf = pseudo_labels(fixed_out2_lo, pred2, out_size, num_classes, thres_high, thres_low)   # :354-365,387-393
    loss_p1 = F.cross_entropy(pred1, conf, ignore_index=IGNORE_LABEL)                          # :394
    loss_p2 = F.cross_entropy(pred2, conf, ignore_index=IGNORE_LABEL)                          # :395
    loss_y1 = simt_head_loss(pred1_lo, T1, label_target, out_size)                             # :402-403,408
    loss_y2 = simt_head_loss(pred2_lo, T2, label_target, out_size)                             # :405-406,409
    convex = convex_loss([W1, W2], [T1, T2])                                                   # :412-415
    volume = volume_loss([T1, T2])                                                             # :417-421
    loss_target = loss_p2 + loss_y2 + lambda_seg * loss_p1 + lambda_seg * loss_y1             # :423
    place = 0.0
    if lambda_place is not None:
        open_classes = pred1_lo.shape[1] - num_classes
        place = lambda_seg * placeholder_loss(pred1, num_classes, open_classes, thres_high, lambda_place)   # :398
        place = place + placeholder_loss(pred2, num_classes, open_classes, thres_high, lambda_place)        # :399
    return place + loss_target + lambda_convex * convex + lambda_volume * volume + lambda_anchor * anchor   # :424


# --------------------------------------------------------------------------
# a12 - a16: integer eval histograms (numpy, single-threaded like the reference)
# --------------------------------------------------------------------------
def fast_hist(a, b, n):
    """tools/compute_iou.py:9-11 (dup tools/evaluate_cityscapes.py:81-83)."""
    k = (a >= 0) & (a < n)
    return np.bincount(n * a[k].astype(int) + b[k], minlength=n ** 2).reshape(n, n)


def fast_hist_rect(a, b, n_rows, n_cols):
    """tools/compute_ConfusionMatrix.py:54-56 (``fast_hist(a, b, n33, n19)``)."""
    ka = (a >= 0) & (a < n_rows)
    return np.bincount(n_cols * a[ka].astype(int) + b[ka], minlength=n_rows * n_cols).reshape(n_rows, n_cols)


def class_hist(a, n):
    """tools/compute_ClassDistribution.py:52-54 (``fast_hist(a, n)``)."""
    ka = (a >= 0) & (a < n)
    return np.bincount(a[ka], minlength=n)


def per_class_iu(hist):
    """tools/compute_iou.py:14-15."""
    return np.diag(hist) / (hist.sum(1) + hist.sum(0) - np.diag(hist))


def label_mapping(inp, mapping):
    """tools/compute_iou.py:18-22: sequential ``out[inp == k] = v`` passes, int64 result."""
    out = np.copy(inp)
    for ind in range(len(mapping)):
        out[inp == mapping[ind][0]] = mapping[ind][1]
    return np.array(out, dtype=np.int64)


def miou_percent(hist) -> float:
    """``round(np.nanmean(per_class_iu(hist)) * 100, 2)``.  tools/compute_iou.py:55-58."""
    with np.errstate(divide="ignore", invalid="ignore"):
        return round(float(np.nanmean(per_class_iu(hist))) * 100, 2)


def class_dist_normalise(counts):
    """``CM / (sum(CM) + 10e-10)``.  tools/compute_ClassDistribution.py:92."""
    counts = np.asarray(counts, dtype=np.float64)
    return counts / (np.sum(counts) + 10e-10)


def confusion_row_normalise(cm):
    """``CM / (rowsum + 10e-6)``.  tools/compute_ConfusionMatrix.py:121."""
    cm = np.asarray(cm, dtype=np.float64)
    return cm / (np.sum(cm, axis=1, keepdims=True) + 10e-6)


# Cityscapes raw-id -> train-id table, dataset/cityscapes_list/info.json:3-38
# ("label2train"); classes = 19 (:2).  Data, restated so the GPU box (which has
# no /root/reference) can build the same 256-entry LUT.
CITYSCAPES_LABEL2TRAIN = [
    [0, 255], [1, 255], [2, 255], [3, 255], [4, 255], [5, 255], [6, 255], [7, 0], [8, 1], [9, 255],
    [10, 255], [11, 2], [12, 3], [13, 4], [14, 255], [15, 255], [16, 255], [17, 5], [18, 255], [19, 6],
    [20, 7], [21, 8], [22, 9], [23, 10], [24, 11], [25, 12], [26, 13], [27, 14], [28, 15], [29, 255],
    [30, 255], [31, 16], [32, 17], [33, 18], [-1, 255],
]


# --------------------------------------------------------------------------
# Seeded synthetic inputs live in simt_b200/synth.py: they are workload generators shared by the
# tests, bench.py and the profiling scripts, not part of the restatement; re-exported for the tests' convenience.
# --------------------------------------------------------------------------
import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
from simt_b200.synth import synth_eval_pair, synth_head_inputs  # noqa: E402,F401
