"""Time Placeholder_loss (trainV2_simt.py:202-230 + the upsample of :371): stock torch ops on the GPU vs the fused launch."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch
import torch.nn.functional as F

import simt_b200

dev = torch.device("cuda")
B, C, K, h, w, H, W = 8, 19, 4, 65, 129, 512, 1024
g = torch.Generator().manual_seed(0)
lo = (3.0 * torch.randn(B, C + K, h, w, generator=g)).to(dev)


def eager(x):
    pred = F.interpolate(x, size=(H, W), mode="bilinear", align_corners=True)
    pseudo = torch.argmax(pred, dim=1)
    onehot = F.one_hot(pseudo, C + K).permute(0, 3, 1, 2).float()
    predict = torch.where(onehot > 0, torch.zeros_like(pred), pred)
    ones = torch.ones_like(pseudo)
    pseudo1 = torch.where(pseudo < C, pseudo, 255 * ones)
    pred_max = torch.max(torch.softmax(pred.detach(), dim=1), 1)[0]
    pseudo1 = torch.where(pred_max > 0.8, pseudo1, 255 * ones)
    known = F.cross_entropy(pred, pseudo1, ignore_index=255)
    po = torch.zeros_like(predict)
    po[:, C:] = predict[:, C:].detach()
    y = torch.argmax(po, dim=1)
    y = torch.where(pseudo1 == 255, 255 * ones, y)
    return known + 0.1 * F.cross_entropy(predict, y, ignore_index=255)


def fused(x):
    return simt_b200.Placeholder_loss(x, C, K, 0.8, out_size=(H, W), lambda_place=0.1)


for name, fn in (("torch eager", eager), ("fused", fused)):
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for it in range(13):
        if it == 3:
            torch.cuda.synchronize()
            ev0.record()
        x = lo.clone().requires_grad_(True)
        loss = fn(x)
        loss.backward()
    ev1.record()
    torch.cuda.synchronize()
    print(f"{name}: {ev0.elapsed_time(ev1) / 10:.3f} ms per fwd+bwd (B={B}, {C}+{K} channels, {h}x{w} -> {H}x{W}); loss {float(loss):.6f}; "
          f"peak mem {torch.cuda.max_memory_allocated() / 2**20:.0f} MiB")
    torch.cuda.reset_peak_memory_stats()
