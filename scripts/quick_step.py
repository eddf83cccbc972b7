"""Quick timing of the MODE_STEP head kernel (what bench.py's roofline reports): python scripts/quick_step.py [B] [K]"""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import simt_b200
from simt_b200 import _lib
from simt_b200 import synth as O  # seeded workload generators
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
K = int(sys.argv[2]) if len(sys.argv) > 2 else 0
lib = _lib.load(); dev = torch.device("cuda")
cd = np.load(os.path.join(ROOT, "simt_b200", "data", "ClassDist_bapa.npy"))
torch.manual_seed(1234); T = simt_b200.sig_NTM(19, K).to(dev)().detach()
res = []
for name, kw in (("bench", dict(coherent=True, block=(36, 52))), ("blk32", dict(coherent=True, block=32)), ("rnd", dict(coherent=False))):
    sets = [tuple(t.to(dev) for t in O.synth_head_inputs(B, 19 + K, 65, 129, 512, 1024, seed=1234 + s, class_dist=cd, **kw)) for s in range(6)]
    rs = [simt_b200.HeadRunner(B, 19 + K, 19, 65, 129, 512, 1024, device=dev) for _ in range(6)]
    for i in range(6): rs[i].step(sets[i][0], T, sets[i][1])
    torch.cuda.synchronize(); lib.simt_b200_profile_enable(1); lib.simt_b200_profile_read(None, None)
    for i in range(24): rs[i % 6].step(sets[i % 6][0], T, sets[i % 6][1])
    torch.cuda.synchronize()
    ms, n = ctypes.c_double(), ctypes.c_longlong(); lib.simt_b200_profile_read(ctypes.byref(ms), ctypes.byref(n)); lib.simt_b200_profile_enable(0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(6): rs[i].graph_step(sets[i][0], T, sets[i][1])
    torch.cuda.synchronize(); e0.record()
    for i in range(60): rs[i % 6].graph_step(sets[i % 6][0], T, sets[i % 6][1])
    e1.record(); torch.cuda.synchronize()
    res.append(f"{name}: kernel {ms.value / n.value * 1e3:.1f}us step {e0.elapsed_time(e1) / 60 * 1e3:.1f}us")
print(os.environ.get("SIMT_B200_LIB", "default").split("/")[-1], f"B={B} K={K}", " | ".join(res), flush=True)
