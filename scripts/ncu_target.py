"""Small driver for ncu captures (run under gpurun): a few launches of one kernel.
   python scripts/ncu_target.py head [B] [K] [coherent]   |   hist [nimg] [coherent]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import simt_b200  # noqa: E402
from simt_b200 import head  # noqa: E402
from simt_b200 import synth as O  # seeded workload generators

dev = torch.device("cuda")
what = sys.argv[1]
if what == "head":
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    K = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    coh = (sys.argv[4] != "0") if len(sys.argv) > 4 else True
    cd = np.load(os.path.join(ROOT, "simt_b200", "data", "ClassDist_bapa.npy"))
    lg, lab = O.synth_head_inputs(B, 19 + K, 65, 129, 512, 1024, seed=1234, coherent=coh, class_dist=cd,
                                   block=(36, 52))   # bench.py's label pattern
    lg, lab = lg.to(dev), lab.to(dev)
    torch.manual_seed(1234)
    T = simt_b200.sig_NTM(19, K).to(dev)().detach()
    r = simt_b200.HeadRunner(B, 19 + K, 19, 65, 129, 512, 1024, device=dev)
    for _ in range(6):
        r.step(lg, T, lab)
    with torch.no_grad():
        for _ in range(3):
            simt_b200.simt_head(lg, T, lab, (512, 1024))
    torch.cuda.synchronize()
else:
    nimg = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    coh = (sys.argv[3] != "0") if len(sys.argv) > 3 else True
    gts, prs = [], []
    for i in range(8):
        gt, pr = O.synth_eval_pair(1024, 2048, seed=i, coherent=coh, block=(96, 160), noise=0.0)
        gts.append(torch.from_numpy(gt)); prs.append(torch.from_numpy(pr))
    gt = torch.stack(gts).repeat(nimg // 8, 1, 1).to(dev)
    pr = torch.stack(prs).repeat(nimg // 8, 1, 1).to(dev)
    m = simt_b200.ConfusionMeter(19, mapping=O.CITYSCAPES_LABEL2TRAIN)
    for _ in range(4):
        m.update(gt, pr)
    torch.cuda.synchronize()
    print(m.value().sum())
