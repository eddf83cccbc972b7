import os, sys, time, ctypes
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import simt_b200
from simt_b200 import _lib
from simt_b200 import synth as O  # seeded workload generators
lib = _lib.load(); dev = torch.device("cuda")
cd = np.load(os.path.join(ROOT, "simt_b200", "data", "ClassDist_bapa.npy"))
lg, lab = O.synth_head_inputs(8, 19, 65, 129, 512, 1024, seed=1, coherent=True, class_dist=cd, block=(36, 52))
lg, lab = lg.to(dev), lab.to(dev)
torch.manual_seed(1234); T = simt_b200.sig_NTM(19, 0).to(dev)().detach()
r = simt_b200.HeadRunner(8, 19, 19, 65, 129, 512, 1024, device=dev)
for prof in (0, 1):
    lib.simt_b200_profile_enable(prof)
    for _ in range(20): r.step(lg, T, lab)
    torch.cuda.synchronize()
    ms, n = ctypes.c_double(), ctypes.c_longlong(); lib.simt_b200_profile_read(ctypes.byref(ms), ctypes.byref(n))
    t0 = time.perf_counter()
    for _ in range(200): r.step(lg, T, lab)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    lib.simt_b200_profile_read(ctypes.byref(ms), ctypes.byref(n))
    print(f"prof={prof}: cpu issue {1e6*(t1-t0)/200:.1f} us/step, total {1e6*(t2-t0)/200:.1f} us/step, kernel {1e3*ms.value/max(n.value,1):.1f} us")
lib.simt_b200_profile_enable(0)
t0 = time.perf_counter()
for _ in range(200): r.fwdbwd(lg, T, lab)
t1 = time.perf_counter(); torch.cuda.synchronize()
print(f"fwdbwd only: cpu issue {1e6*(t1-t0)/200:.1f} us")
t0 = time.perf_counter()
for _ in range(200): r.scale()
t1 = time.perf_counter(); torch.cuda.synchronize()
print(f"scale only: cpu issue {1e6*(t1-t0)/200:.1f} us")
# CUDA graph
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    r.step(lg, T, lab); torch.cuda.synchronize()
    with torch.cuda.graph(g, stream=s):
        r.step(lg, T, lab)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(10): g.replay()
e0.record()
for _ in range(200): g.replay()
e1.record(); torch.cuda.synchronize()
print(f"graph replay: {1e3*e0.elapsed_time(e1)/200:.1f} us/step; loss {float(r.loss):.6f}")
