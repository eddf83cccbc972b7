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


def w_fit(weight: torch.Tensor, exp_avg: torch.Tensor, exp_avg_sq: torch.Tensor, T: torch.Tensor, n_steps: int,
          step0: int, lr: float, betas=(0.9, 0.999), eps: float = 1e-8, dT_accum: torch.Tensor | None = None,
          losses: torch.Tensor | None = None) -> None:
    """``n_steps`` rounds of (sig_W forward, ``||W T||^2``, backward, Adam step) in ONE launch, all in place
    (trainV2_simt.py:326-339 for one head; see ``simt_w_fit`` in include/simt_b200.h)."""
    lib = _lib.load()
    ck, c = T.shape
    for name, t, shape in (("weight", weight, (ck, ck)), ("exp_avg", exp_avg, (ck, ck)),
                           ("exp_avg_sq", exp_avg_sq, (ck, ck)), ("T", T, (ck, c)), ("dT_accum", dT_accum, (ck, c)),
                           ("losses", losses, (n_steps,))):
        if t is None:
            continue
        if not t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous() or tuple(t.shape) != shape:
            raise ValueError(f"w_fit: {name} must be a contiguous CUDA float32 tensor of shape {shape}")
    _lib.check(lib.simt_w_fit(weight.data_ptr(), exp_avg.data_ptr(), exp_avg_sq.data_ptr(), T.data_ptr(), ck, c,
                              int(n_steps), int(step0), float(lr), float(betas[0]), float(betas[1]), float(eps),
                              dT_accum.data_ptr() if dT_accum is not None else None,
                              losses.data_ptr() if losses is not None else None, _stream_ptr()), "simt_w_fit")


def fit_w(ntm: torch.nn.Module, w_module: torch.nn.Module, optimizer: torch.optim.Optimizer, steps: int = 10,
          accumulate_t_grad: bool = True, return_losses: bool = False):
    """The reference's inner loop ``for iter in range(10): ... optimizer_w.step()`` (trainV2_simt.py:326-339)
    for ONE head: ``fit_w(NTM1, NTM_W1, optimizer_w1)``.  ``optimizer`` must be the ``torch.optim.Adam`` that owns
    ``w_module.weight`` (:277); its ``exp_avg`` / ``exp_avg_sq`` / ``step`` state is read and advanced exactly as
    ``steps`` calls of ``optimizer.step()`` would, and ``w_module.weight.grad`` is left zeroed-equivalent (None).
    With ``accumulate_t_grad`` the T-side gradient of the ``steps`` backward passes is accumulated into
    ``ntm.NTM.grad`` as the reference's ``NTM_loss.backward(retain_graph=True)`` (:337) does.
    Returns the per-round objectives (device tensor) when ``return_losses``."""
    p = w_module.weight
    if not isinstance(optimizer, torch.optim.Adam):
        raise TypeError("fit_w: the fused inner loop implements torch.optim.Adam (trainV2_simt.py:277)")
    group = next((g for g in optimizer.param_groups if any(q is p for q in g["params"])), None)
    if group is None:
        raise ValueError("fit_w: optimizer does not own w_module.weight")
    if group.get("amsgrad") or group.get("weight_decay", 0) != 0 or group.get("maximize"):
        raise NotImplementedError("fit_w: amsgrad / weight_decay / maximize are not used by the reference")
    if not p.is_cuda:
        raise RuntimeError("fit_w: CUDA only (no CPU fallback)")
    state = optimizer.state[p]
    if len(state) == 0:
        state["step"] = torch.tensor(0.0, dtype=torch.float32)
        state["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        state["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
    step0 = int(state["step"].item()) if torch.is_tensor(state["step"]) else int(state["step"])
    lr = group["lr"]
    lr = float(lr.item()) if torch.is_tensor(lr) else float(lr)
    T = ntm()
    Td = T.detach().contiguous()
    want_t = accumulate_t_grad and T.requires_grad
    dT = torch.zeros_like(Td) if want_t else None
    losses = torch.empty(steps, dtype=torch.float32, device=p.device) if return_losses else None
    with torch.no_grad():
        w_fit(p.data, state["exp_avg"], state["exp_avg_sq"], Td, steps, step0, lr, group["betas"], group["eps"],
              dT_accum=dT, losses=losses)
    if torch.is_tensor(state["step"]):
        state["step"] += steps
    else:
        state["step"] = step0 + steps
    p.grad = None
    if want_t:
        T.backward(dT)
    return losses


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
