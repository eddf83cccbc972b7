"""Measure the widened rows of SURVEY section 8(f) on the GPU: each fused entry point next to the reference's own lines
written with stock torch ops on the SAME device (the reference has no CUDA kernels of its own; this is what it runs).
Writes one JSON object (also to argv[1] if given).  CUDA events, warm-up, inputs rotate.  GPU only."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch
import torch.nn.functional as F

import simt_b200

dev = torch.device("cuda")
C, K, h, w, H, W, B = 19, 4, 65, 129, 512, 1024, 8
CK = C + K
g = torch.Generator().manual_seed(0)


def rnd(*shape, scale=3.0):
    return (scale * torch.randn(*shape, generator=g)).to(dev)


def timeit(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    torch.cuda.reset_peak_memory_stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, torch.cuda.max_memory_allocated() / 2**20


out = {"device": torch.cuda.get_device_name(0), "shape": f"B={B}, {C}+{K} channels, {h}x{w} -> {H}x{W}", "rows": {}}
up = lambda x, size=(H, W): F.interpolate(x, size=size, mode="bilinear", align_corners=True)

# ---- (f) 1: plain CE on the upsampled logits, trainV2_simt.py:394-395 -------------------------------------------
lo = rnd(B, CK, h, w)
lab = torch.randint(0, CK, (B, H // 32, W // 32), generator=g).repeat_interleave(32, 1).repeat_interleave(32, 2)
lab[torch.rand(B, H, W, generator=g) < 0.1] = 255
lab8, lab64 = lab.to(torch.uint8).to(dev), lab.to(dev)


def eager_ce():
    x = lo.clone().requires_grad_(True)
    F.cross_entropy(up(x), lab64, ignore_index=255).backward()


def fused_ce():
    x = lo.clone().requires_grad_(True)
    simt_b200.simt_head(x, None, lab8, (H, W)).backward()


# ---- (f) 2: pseudo labels, :354-365 + :387-393 ----------------------------------------------------------------------
fixed = rnd(B, C, h, w)


def eager_pseudo():
    labelC = up(torch.softmax(fixed, dim=1))
    mx = torch.max(labelC, 1)
    am = torch.argmax(labelC, dim=1).float()
    lc = torch.where(mx[0] > 0.8, am, 255. * torch.ones_like(am))
    lc = torch.where(mx[0] < 0.2, C * torch.ones_like(am), lc)
    conf = torch.from_numpy(lc.detach().clone().cpu().numpy()).long().to(dev)          # the host round trip of :362
    pseudo = torch.argmax(up(lo), dim=1)
    ones, zeros = torch.ones_like(conf), torch.zeros_like(conf)
    mask = torch.where(conf == C * ones, ones, zeros)
    p1 = mask * pseudo
    p1 = torch.where(p1 >= C * ones, p1, 255 * ones)
    return torch.where(conf == C * ones, p1, conf)


def fused_pseudo():
    return simt_b200.pseudo_labels(fixed, lo, (H, W), C, 0.8, 0.2)


# ---- (f) 3: two-scale eval arg-max + confusion matrix, evaluate_cityscapes.py:127-148 (one 1024x2048 image) ----------
ea, eb = rnd(1, CK, 129, 257), rnd(1, CK, 81, 161)
gt = torch.randint(0, 34, (1, 1024, 2048), generator=g, dtype=torch.uint8)
gt_dev = gt.to(dev)
meter = simt_b200.ConfusionMeter(C, mapping=None, device=dev)


def eager_eval():
    o = up(ea[:, :C], (1024, 2048)).cpu().data[0].numpy()
    o += up(eb[:, :C], (1024, 2048)).cpu().data[0].numpy()
    pred = np.asarray(np.argmax(o.transpose(1, 2, 0), axis=2))
    a, b = gt[0].numpy().flatten(), pred.flatten()
    k = (a >= 0) & (a < C)
    return np.bincount(C * a[k].astype(int) + b[k], minlength=C * C).reshape(C, C)


def fused_eval():
    pred = simt_b200.eval_argmax(ea, eb, (1024, 2048), C)
    meter.update(gt_dev, pred)


# ---- (f) 4b: Placeholder_loss, :202-230 after :371 -------------------------------------------------------------------
def eager_place():
    x = lo.clone().requires_grad_(True)
    pred = up(x)
    pseudo = torch.argmax(pred, dim=1)
    onehot = F.one_hot(pseudo, CK).permute(0, 3, 1, 2).float()
    predict = torch.where(onehot > 0, torch.zeros_like(pred), pred)
    ones = torch.ones_like(pseudo)
    p1 = torch.where(pseudo < C, pseudo, 255 * ones)
    p1 = torch.where(torch.max(torch.softmax(pred.detach(), dim=1), 1)[0] > 0.8, p1, 255 * ones)
    known = F.cross_entropy(pred, p1, ignore_index=255)
    po = torch.zeros_like(predict)
    po[:, C:] = predict[:, C:].detach()
    y = torch.where(p1 == 255, 255 * ones, torch.argmax(po, dim=1))
    (known + 0.1 * F.cross_entropy(predict, y, ignore_index=255)).backward()


def fused_place():
    x = lo.clone().requires_grad_(True)
    simt_b200.Placeholder_loss(x, C, K, 0.8, out_size=(H, W), lambda_place=0.1).backward()


# ---- (f) 4a: the inner W loop, :326-339 (both heads, 10 Adam rounds) ---------------------------------------------------
ntm = [simt_b200.sig_NTM(C, K).to(dev) for _ in range(2)]
wm = [simt_b200.sig_W(C, K).to(dev) for _ in range(2)]
opt_w = [torch.optim.Adam(m.parameters(), lr=2.5e-4, weight_decay=0) for m in wm]
mse = torch.nn.MSELoss(reduction="sum")
zeros = torch.zeros(CK, C, device=dev)


def eager_wfit():
    for _ in range(10):
        T1, T2, W1, W2 = ntm[0](), ntm[1](), wm[0](), wm[1]()
        opt_w[0].zero_grad(); opt_w[1].zero_grad()
        (mse(W1.mm(T1), zeros) + mse(W2.mm(T2), zeros)).backward(retain_graph=True)
        opt_w[0].step(); opt_w[1].step()


def fused_wfit():
    for i in range(2):
        simt_b200.fit_w(ntm[i], wm[i], opt_w[i], steps=10)


# ---- a8-a10: convex + volume + anchor for both heads (B = 1 as the anchor lines require), :375-384,412-421 ------------
p1a, p2a, fx1 = rnd(1, CK, h, w), rnd(1, CK, h, w), rnd(1, C, h, w)


def eager_reg():
    T1, T2, W1, W2 = ntm[0](), ntm[1](), wm[0](), wm[1]()
    labelC_flat = up(torch.softmax(fx1, dim=1)).permute(0, 2, 3, 1).reshape(-1, C)
    anchor = 0.0
    for pr, T in ((p1a, T1), (p2a, T2)):
        flat = up(pr).permute(0, 2, 3, 1).reshape(-1, CK).detach()
        ai = torch.argmax(flat, dim=0)
        ex = torch.unique(torch.argmax(flat, dim=1))
        anchor = anchor + mse(T[ex], labelC_flat[ai][ex])
    convex = 0.0 - (mse(W1.mm(T1), zeros) + mse(W2.mm(T2), zeros))
    vol = torch.log(torch.sqrt(torch.abs(torch.linalg.det(T1.t().mm(T1))))) + \
        torch.log(torch.sqrt(torch.abs(torch.linalg.det(T2.t().mm(T2)))))
    if torch.isinf(vol) or torch.isnan(vol):
        vol = 0.0
    (0.1 * convex + vol + anchor).backward()


def fused_reg():
    T1, T2, W1, W2 = ntm[0](), ntm[1](), wm[0](), wm[1]()
    c1, v1 = simt_b200.t_regularizers(T1, W1)
    c2, v2 = simt_b200.t_regularizers(T2, W2)
    anchor = simt_b200.anchor_loss([p1a, p2a], [T1, T2], fx1, (H, W))
    (0.1 * (c1 + c2) + (v1 + v2) + anchor).backward()


for name, ref_lines, eager, fused in (
        ("f1 plain CE on upsampled logits fwd+bwd", "trainV2_simt.py:394-395", eager_ce, fused_ce),
        ("f2 pseudo labels + class-posterior relabel", "trainV2_simt.py:354-365,387-393", eager_pseudo, fused_pseudo),
        ("f3 two-scale eval arg-max + confusion matrix, one 1024x2048 image", "evaluate_cityscapes.py:127-148", eager_eval, fused_eval),
        ("f4a inner W loop, 2 heads x 10 Adam rounds", "trainV2_simt.py:326-339", eager_wfit, fused_wfit),
        ("f4b Placeholder_loss fwd+bwd", "trainV2_simt.py:202-230,398", eager_place, fused_place),
        ("a8-a10 convex + volume + anchor, 2 heads, B=1, fwd+bwd", "trainV2_simt.py:375-384,412-421", eager_reg, fused_reg)):
    te, me = timeit(eager, n=10)
    tf, mf = timeit(fused, n=50)
    out["rows"][name] = {"reference_lines": ref_lines, "torch_ops_same_gpu_ms": round(te, 4), "torch_peak_mib": round(me),
                         "simt_b200_ms": round(tf, 4), "simt_b200_peak_mib": round(mf), "speedup": round(te / tf, 1)}
    print(f"{name}: torch {te:.3f} ms ({me:.0f} MiB) -> simt_b200 {tf:.3f} ms ({mf:.0f} MiB)  x{te / tf:.1f}", flush=True)
simt_b200.check_errors(dev)
s = json.dumps(out, indent=1)
if len(sys.argv) > 1:
    open(sys.argv[1], "w").write(s + "\n")
