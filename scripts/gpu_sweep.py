"""Tuning sweep on a B200 (run under gpurun): times the fused head for tile / lane-split choices and
the histogram kernel modes, with the library's per-launch event profiler.  Writes
gpurun_out/sweep_<tag>.json.  Not part of the product; bench.py is the contract benchmark."""
import ctypes
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import simt_b200  # noqa: E402
from simt_b200 import _lib, head  # noqa: E402
from simt_b200 import synth as O  # seeded workload generators

PEAK = 6559.7
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def prof_read(lib):
    ms, n = ctypes.c_double(), ctypes.c_longlong()
    lib.simt_b200_profile_read(ctypes.byref(ms), ctypes.byref(n))
    return ms.value, n.value


def time_head(lib, sets, T, size, iters=20, need_grad=True):
    for i in range(3):
        lg, lab = sets[i % len(sets)]
        head.head_forward_raw(lg, T, lab, size, need_grad=need_grad)
    torch.cuda.synchronize()
    lib.simt_b200_profile_enable(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        lg, lab = sets[i % len(sets)]
        head.head_forward_raw(lg, T, lab, size, need_grad=need_grad)
    e1.record()
    torch.cuda.synchronize()
    kms, n = prof_read(lib)
    lib.simt_b200_profile_enable(0)
    return e0.elapsed_time(e1) / iters, kms / max(n, 1)


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
    lib = _lib.load()
    dev = torch.device("cuda")
    out = {"peak_gbs": PEAK, "head": [], "hist": [], "gpu": torch.cuda.get_device_name(0)}
    cd = np.load(os.path.join(ROOT, "simt_b200", "data", "ClassDist_bapa.npy"))
    H, W, h, w = 512, 1024, 65, 129
    for (B, K) in ((8, 0), (8, 4), (8, 15), (1, 0)) if 'nohead' not in sys.argv else ():
        CK = 19 + K
        nsets = 12 if B == 8 else 4
        for coherent in (True, False):
            sets = []
            for s in range(nsets):
                lg, lab = O.synth_head_inputs(B, CK, h, w, H, W, seed=1234 + s, coherent=coherent, class_dist=cd)
                sets.append((lg.to(dev), lab.to(dev)))
            torch.manual_seed(1234)
            T = simt_b200.sig_NTM(19, K).to(dev)().detach()
            labeled = float(sum(int((l != 255).sum()) for _, l in sets)) / len(sets)
            cfgs = [(0, 0, 0), (1, 0, 2), (2, 0, 2), (1, 0, 4), (2, 0, 4)]
            for (tcy, tcx, lpr) in cfgs:
                if K > 0 and lpr == 2:
                    continue
                lib.simt_head_set_tuning(tcy, tcx, 0, lpr)
                try:
                    for mode, ng in (("fwdbwd", True), ("fwd", False)):
                        call_ms, k_ms = time_head(lib, sets, T, (H, W), need_grad=ng)
                        alg = B * (4 * CK * h * w * (2 if ng else 1) + H * W)
                        out["head"].append(dict(B=B, K=K, coherent=coherent, tcy=tcy, tcx=tcx, lpr=lpr, mode=mode,
                                                call_ms=call_ms, kernel_ms=k_ms, labeled_px=labeled,
                                                gpx_s=B * H * W / k_ms / 1e6, alg_gbs=alg / k_ms / 1e6,
                                                frac=alg / k_ms / 1e6 / PEAK))
                        print(out["head"][-1], flush=True)
                except RuntimeError as e:
                    print("skip", tcy, tcx, lpr, e, flush=True)
                finally:
                    lib.simt_head_set_tuning(0, 0, 0, 0)
            del sets
    if 'nohist' in sys.argv:
        json.dump(out, open(os.path.join(ROOT, 'gpurun_out', f'sweep_{tag}.json'), 'w'), indent=1)
        return
    # ---- histograms: BIG launches (32 images = 64 Mi pixels per launch), 3 rotating sets > L2 ----------
    lut = torch.from_numpy(simt_b200.build_lut(O.CITYSCAPES_LABEL2TRAIN)).to(dev)
    nimg, nsets = (256, 2) if 'bighist' in sys.argv else (32, 3)
    for coherent in (('clean',) if 'bighist' in sys.argv else ('clean', 'noisy', False)):
        gt_sets, pr_sets = [], []
        for s_ in range(nsets):
            gts, prs = [], []
            for i in range(8):
                gt, pr = O.synth_eval_pair(1024, 2048, seed=8 * s_ + i, coherent=bool(coherent), block=(96, 160) if coherent else 32,
                                            noise=0.05 if coherent == 'noisy' else 0.0)
                gts.append(torch.from_numpy(gt)); prs.append(torch.from_numpy(pr))
            gt_sets.append(torch.stack(gts).repeat(nimg // 8, 1, 1).to(dev).reshape(-1))
            pr_sets.append(torch.stack(prs).repeat(nimg // 8, 1, 1).to(dev).reshape(-1))
        npx = gt_sets[0].numel()
        for (rows, cols, use_lut) in ((19, 19, True), (34, 19, False), (19, 1, False)):
            for mode in (0, 9):
                for warps, unroll in ((16, 1), (16, 2), (16, 4), (8, 4)):
                    lib.simt_hist_set_tuning(mode, warps, unroll)
                    hist = torch.zeros(rows * cols, dtype=torch.int64, device=dev)
                    try:
                        def run(i):
                            if cols == 1:
                                simt_b200.hist.class_hist_into(hist, pr_sets[i % nsets], rows)
                            else:
                                simt_b200.hist.confusion_into(hist, gt_sets[i % nsets], pr_sets[i % nsets], rows, cols,
                                                              lut if use_lut else None)
                        for i in range(2):
                            run(i)
                        torch.cuda.synchronize()
                        lib.simt_b200_profile_enable(1)
                        for i in range(6):
                            run(i)
                        torch.cuda.synchronize()
                        kms, n = prof_read(lib)
                        lib.simt_b200_profile_enable(0)
                        k_ms = kms / n
                        nbytes = (1 if cols == 1 else 2) * npx
                        out["hist"].append(dict(rows=rows, cols=cols, lut=use_lut, coherent=str(coherent), mode=mode,
                                                warps=warps, unroll=unroll, kernel_ms=k_ms, gbs=nbytes / k_ms / 1e6,
                                                frac=nbytes / k_ms / 1e6 / PEAK, npx=npx))
                        print(out["hist"][-1], flush=True)
                    except RuntimeError as e:
                        print("skip hist", rows, cols, mode, warps, unroll, e, flush=True)
                    finally:
                        lib.simt_hist_set_tuning(0, 0, 0)
        del gt_sets, pr_sets
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"sweep_{tag}.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
