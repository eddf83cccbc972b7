"""Where the sharded step's time goes (run under torchrun on >= 2 GPUs): CUDA events around prep / kernel / finalize
of eager steps, mean over steps, per rank.  python -m torch.distributed.run --nproc-per-node 2 scripts/xchg_parts.py"""
import os, sys, datetime
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import simt_b200
from simt_b200 import synth as O
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0))); dev = torch.device("cuda", torch.cuda.current_device())
group = None
if world > 1:
    dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=60)); group = dist.group.WORLD
cd = np.load(os.path.join(ROOT, "simt_b200", "data", "ClassDist_bapa.npy"))
torch.manual_seed(1234); T = simt_b200.sig_NTM(19, 0).to(dev)().detach()
sets = [tuple(t.to(dev) for t in O.synth_head_inputs(8, 19, 65, 129, 512, 1024, seed=1234 + (0 if os.environ.get("SAME") else 1000 * rank) + s, coherent=True, class_dist=cd, block=(36, 52))) for s in range(6)]
rs = [simt_b200.HeadRunner(8, 19, 19, 65, 129, 512, 1024, device=dev, group=None if os.environ.get("NOSHARD") else group) for _ in range(6)]
for i in range(12): rs[i % 6].step(sets[i % 6][0], T, sets[i % 6][1])
torch.cuda.synchronize()
outs = [torch.empty(8, 19, 65, 129, device=dev) for _ in range(6)]
one = rs[0]
for mode in ("eager", "graph", "pipelined-eager", "pipelined-graph"):
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fn = (lambda i: rs[i % 6].step(sets[i % 6][0], T, sets[i % 6][1])) if mode == "eager" else (lambda i: rs[i % 6].graph_step(sets[i % 6][0], T, sets[i % 6][1]))
    if mode == "pipelined-eager":
        fn = lambda i: one.step(sets[i % 6][0], T, sets[i % 6][1], next_labels=sets[(i + 1) % 6][1], defer=True, out=outs[i % 6])
    if mode == "pipelined-graph":     # ONE runner, rotating inputs / dLogits buffers; next labels announced, all-reduce deferred
        fn = lambda i: one.graph_step(sets[i % 6][0], T, sets[i % 6][1], next_labels=sets[(i + 1) % 6][1], defer=True, out=outs[i % 6])
    for i in range(6): fn(i)
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    import ctypes
    from simt_b200 import _lib
    lib = _lib.load()
    if mode.endswith("eager"):
        lib.simt_b200_profile_enable(1); lib.simt_b200_profile_read(None, None)
    e0.record()
    for i in range(120): fn(i)
    if mode.startswith("pipelined"): one.finish()
    e1.record(); torch.cuda.synchronize()
    extra = ""
    if mode.endswith("eager"):
        ms, n = ctypes.c_double(), ctypes.c_longlong(); lib.simt_b200_profile_read(ctypes.byref(ms), ctypes.byref(n)); lib.simt_b200_profile_enable(0)
        extra = f" [profiled kernel (SIMT_PROF_WHICH={os.environ.get('SIMT_PROF_WHICH', '0')}): {ms.value / max(n.value, 1) * 1e3:.1f} us]"
    print(f"rank {rank}/{world} {mode}: {e0.elapsed_time(e1) / 120 * 1e3:.1f} us/step{extra}", flush=True)
if world > 1: dist.destroy_process_group()
