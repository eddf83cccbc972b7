"""Launch every kernel of libsimt_b200.so a few times at the training / eval shapes (for an ncu per-kernel table):
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv ... \\
        python scripts/all_kernels_once.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

import simt_b200

dev = torch.device("cuda")
C, K, h, w, H, W, B = 19, 4, 65, 129, 512, 1024, 8
CK = C + K
g = torch.Generator().manual_seed(0)
rnd = lambda *s: (3.0 * torch.randn(*s, generator=g)).to(dev)
lab = torch.randint(0, C, (B, H // 32, W // 32), generator=g).repeat_interleave(32, 1).repeat_interleave(32, 2)
lab[torch.rand(B, H, W, generator=g) < 0.1] = 255
lab8 = lab.to(torch.uint8).to(dev)
ntm, wm = simt_b200.sig_NTM(C, K).to(dev), simt_b200.sig_W(C, K).to(dev)
opt_w = torch.optim.Adam(wm.parameters(), lr=2.5e-4)
for rep in range(3):
    lo = rnd(B, CK, h, w).requires_grad_(True)
    T, Wm = ntm(), wm()
    loss = simt_b200.simt_head(lo, T, lab8, (H, W))                                   # head_kernel FWDBWD, finalize, scale
    loss = loss + simt_b200.simt_head(lo, None, lab8, (H, W))                         # plain CE
    loss = loss + simt_b200.Placeholder_loss(lo, C, K, 0.8, out_size=(H, W))          # MODE_PLACE
    c, v = simt_b200.t_regularizers(T, Wm)                                            # t_reg_kernel
    loss = loss + 0.1 * c + v + simt_b200.anchor_loss([lo[:1]], [T], rnd(1, C, h, w), (H, W))   # anchor kernels, gather
    loss.backward()
    simt_b200.fit_w(ntm, wm, opt_w, steps=10)                                         # w_fit_kernel
    simt_b200.pseudo_labels(rnd(B, C, h, w), lo.detach(), (H, W), C)                  # softmax_lo, pseudo_label
    r = simt_b200.HeadRunner(B, CK, C, h, w, H, W, device=dev)
    r.step(lo.detach(), T.detach(), lab8)                                             # prep, head_kernel STEP, finalize
    with torch.no_grad():
        simt_b200.simt_head(lo.detach(), T.detach(), lab8, (H, W))                    # head_kernel FWD
    probs = torch.rand(2, C, 64, 128, generator=g).to(dev)
    simt_b200.CrossEntropy2d(is_softmax=False)(probs / probs.sum(1, keepdim=True), lab8[:2, :64, :128].long())   # nll2d
    gt = torch.randint(0, 34, (8, 1024, 2048), generator=g, dtype=torch.uint8).to(dev)
    pred = simt_b200.eval_argmax(rnd(8, CK, 129, 257), rnd(8, CK, 81, 161), (1024, 2048), C)   # eval_argmax_kernel
    m = simt_b200.ConfusionMeter(C, mapping=None, device=dev)
    m.update(gt, pred)                                                                # hist_u8_kernel<true>
    simt_b200.fast_hist(pred, C)                                                      # hist_u8_kernel<false>
    simt_b200.label_mapping(gt[0].cpu().numpy(), np.array([[i, i % 19] for i in range(34)]))   # label_map_kernel
torch.cuda.synchronize()
simt_b200.check_errors(dev)
print("ok")
