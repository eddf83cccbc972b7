"""Time the inner W optimisation (trainV2_simt.py:326-339): literal torch loop vs the fused launch. GPU only."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

import simt_b200

dev = torch.device("cuda")
K = 15
ntm = [simt_b200.sig_NTM(19, K).to(dev) for _ in range(2)]
wm = [simt_b200.sig_W(19, K).to(dev) for _ in range(2)]
opt_w = [torch.optim.Adam(m.parameters(), lr=2.5e-4, weight_decay=0) for m in wm]
mse = torch.nn.MSELoss(reduction="sum")
zeros = torch.zeros(19 + K, 19, device=dev)


def eager():
    for _ in range(10):
        T1, T2, W1, W2 = ntm[0](), ntm[1](), wm[0](), wm[1]()
        opt_w[0].zero_grad(); opt_w[1].zero_grad()
        loss = mse(W1.mm(T1), zeros) + mse(W2.mm(T2), zeros)
        loss.backward(retain_graph=True)
        opt_w[0].step(); opt_w[1].step()


def fused():
    for i in range(2):
        simt_b200.fit_w(ntm[i], wm[i], opt_w[i], steps=10)


for name, fn in (("eager", eager), ("fused", fused)):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 20
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    print(f"{name}: {(time.perf_counter() - t0) / n * 1e3:.3f} ms per outer iteration (2 heads x 10 Adam rounds)")
