import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import simt_b200
from simt_b200 import _lib
from simt_b200 import synth as O  # seeded workload generators
lib = _lib.load(); dev = torch.device("cuda")
cd = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "simt_b200", "data", "ClassDist_bapa.npy"))
torch.manual_seed(1234); T = simt_b200.sig_NTM(19, 0).to(dev)().detach()
sets = [tuple(t.to(dev) for t in O.synth_head_inputs(8, 19, 65, 129, 512, 1024, seed=1234 + s, coherent=True, class_dist=cd, block=(36, 52))) for s in range(12)]
runners = [simt_b200.HeadRunner(8, 19, 19, 65, 129, 512, 1024, device=dev) for _ in range(12)]
def timeit(fn, n=200):
    for i in range(10): fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
def f_step(i): runners[i % 12].step(sets[i % 12][0], T, sets[i % 12][1])
def f_fwdbwd(i): runners[i % 12].fwdbwd(sets[i % 12][0], T, sets[i % 12][1])
def f_scale(i): runners[i % 12].scale()
def f_memset(i): runners[i % 12].dlogits.zero_()
lib.simt_b200_profile_enable(1)
t_step = timeit(f_step)
ms, n = ctypes.c_double(), ctypes.c_longlong(); lib.simt_b200_profile_read(ctypes.byref(ms), ctypes.byref(n)); lib.simt_b200_profile_enable(0)
print("kernel avg us", ms.value / n.value * 1e3)
print("step us", t_step, "(profiler events on)")
print("step us", timeit(f_step), "fwdbwd us", timeit(f_fwdbwd), "scale us", timeit(f_scale), "memset us", timeit(f_memset))
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for i in range(3): f_step(i)
    torch.cuda.synchronize()
    with torch.cuda.graph(g, stream=s):
        for i in range(12): f_step(i)
    torch.cuda.synchronize()
    def f_graph(i): g.replay()
    print("graph of 12 steps: us per step", timeit(f_graph, 30) / 12)
