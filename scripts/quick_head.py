"""Quick A/B timing of the head kernel: python scripts/quick_head.py [B] [K]  (SIMT_B200_LIB selects the build)."""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import simt_b200
from simt_b200 import _lib, head
from simt_b200 import synth as O  # seeded workload generators
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
K = int(sys.argv[2]) if len(sys.argv) > 2 else 0
BLOCK = (36, 52) if (len(sys.argv) <= 3 or sys.argv[3] != 'aligned') else 32
lib = _lib.load(); dev = torch.device("cuda")
RS = int(os.environ.get('SIMT_RS', '0')); FL = int(os.environ.get('SIMT_FLAGS', '0')); LP = int(os.environ.get('SIMT_LPR', '0')); lib.simt_head_set_tuning(0, RS, FL, LP)
cd = np.load(os.path.join(ROOT, "simt_b200", "data", "ClassDist_bapa.npy"))
torch.manual_seed(1234); T = simt_b200.sig_NTM(19, K).to(dev)().detach()
res = []
for coh in (True, False):
    sets = [tuple(t.to(dev) for t in O.synth_head_inputs(B, 19 + K, 65, 129, 512, 1024, seed=1234 + s, coherent=coh, class_dist=cd, block=BLOCK)) for s in range(6)]
    for ng in (True, False):
        for i in range(3): head.head_forward_raw(sets[i][0], T, sets[i][1], (512, 1024), need_grad=ng)
        torch.cuda.synchronize(); lib.simt_b200_profile_enable(1)
        for i in range(18): head.head_forward_raw(sets[i % 6][0], T, sets[i % 6][1], (512, 1024), need_grad=ng)
        torch.cuda.synchronize()
        ms, n = ctypes.c_double(), ctypes.c_longlong(); lib.simt_b200_profile_read(ctypes.byref(ms), ctypes.byref(n)); lib.simt_b200_profile_enable(0)
        res.append(f"{'coh' if coh else 'rnd'}-{'fwdbwd' if ng else 'fwd'}={ms.value / n.value * 1e3:.1f}us")
print(os.environ.get("SIMT_B200_LIB", "default").split("/")[-1], f"small_pct={RS} flags={FL} lpr={LP}", f"B={B} K={K}", " ".join(res), flush=True)
