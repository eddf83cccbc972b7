"""Memory-safety / undefined-behaviour sweep of the fused head kernel WITHOUT a GPU: the product kernel source compiled
with g++ -fsanitize=address,undefined against the CUDA-on-CPU shim (tests/cpu_simt) and run over random shapes, channel
counts, label dtypes / alignments, modes and emulated grid sizes.  Every global buffer is a malloc'd numpy array (red
zones under ASan), dynamic shared memory a heap block of exactly the size the launch plan asks for.

    python scripts/head_emul_sanitize.py [n_cases] [seed]       (re-executes itself with libasan preloaded)
"""
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "build", "cpu_simt")
SO = os.path.join(OUT, "libhead_emul_asan.so")


def build():
    os.makedirs(OUT, exist_ok=True)
    csrc = os.path.join(ROOT, "simt_b200", "csrc")
    src = os.path.join(ROOT, "tests", "cpu_simt", "head_emul.cpp")
    deps = [src, os.path.join(ROOT, "tests", "cpu_simt", "cuda_shim.h")] + [os.path.join(csrc, f) for f in os.listdir(csrc)]
    if os.path.exists(SO) and os.path.getmtime(SO) >= max(os.path.getmtime(d) for d in deps):
        return
    cmd = ["g++", "-O0", "-g", "-std=c++17", "-pthread", "-shared", "-fPIC", "-ffp-contract=off", "-fsanitize=address,undefined",
           "-fno-sanitize-recover=undefined", "-Wno-unknown-pragmas", "-I", os.path.join(ROOT, "tests", "cpu_simt"), "-I", csrc,
           src, "-o", SO]
    print("building", SO, "(a few minutes)", flush=True)
    subprocess.check_call(cmd)


def sweep(n_cases, seed):
    import numpy as np
    lib = ctypes.CDLL(SO)
    vp, i, f = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
    lib.emul_head.restype = i
    lib.emul_head.argtypes = [i, vp, i, i, i, i, vp, i, vp, i, i, i, i, f, f, f, vp, vp, vp, vp, vp, i, i, i, i, ctypes.c_ulonglong]
    rng = np.random.default_rng(seed)
    done = 0
    for case in range(n_cases):
        K = int(rng.choice([0, 0, 4, 15, 1, 9]))
        C, CK = 19, 19 + K
        B = int(rng.integers(1, 3))
        h, w = int(rng.integers(1, 12)), int(rng.integers(1, 14))
        kind = rng.choice(["up", "up", "same", "down"])
        if kind == "up":
            H, W = h * int(rng.integers(1, 7)) + int(rng.integers(0, 5)), w * int(rng.integers(1, 7)) + int(rng.integers(0, 5))
        elif kind == "same":
            H, W = h, w
        else:
            H, W = max(1, h // 2), max(1, w // 2 + 1)
        mode = int(rng.choice([0, 1, 2, 3, 4, 4, 1]))
        i64 = bool(rng.integers(0, 2))
        ident = mode != 3 and K == 0 and bool(rng.integers(0, 3) == 0)
        logits = (float(rng.choice([0.5, 3.0, 30.0])) * rng.standard_normal((B, CK, h, w))).astype(np.float32)
        T = None if ident else rng.random((CK, C)).astype(np.float32)
        if T is not None:
            T /= T.sum(1, keepdims=True)
        off = int(rng.integers(0, 4)) if not i64 else 0            # misaligned uint8 label pointer
        raw = rng.integers(0, C, size=B * H * W + off)
        raw[rng.random(raw.size) < float(rng.choice([0.0, 0.1, 0.9]))] = 255
        lab_store = raw.astype(np.int64 if i64 else np.uint8)
        lab = lab_store[off:]
        dl = np.empty_like(logits)
        stats = np.empty(2 + CK * C, np.float64); loss = np.empty(1, np.float32); dT = np.empty((CK, C), np.float32)
        err = np.zeros(1, np.int32)
        if os.environ.get('HEAD_EMUL_VERBOSE'): print(case, dict(mode=mode,B=B,K=K,h=h,w=w,H=H,W=W,i64=i64,ident=ident,off=off), flush=True)
        rc = lib.emul_head(mode, logits.ctypes.data, B, CK, h, w, None if T is None else T.ctypes.data, C, lab.ctypes.data,
                           8 if i64 else 1, H, W, 255, 1.0, float(rng.choice([-1.0, 0.5])), 0.1, dl.ctypes.data, stats.ctypes.data,
                           loss.ctypes.data, dT.ctypes.data, err.ctypes.data, int(rng.integers(1, 5)), int(rng.integers(1, 4)),
                           int(rng.choice([0, 0, 1, 2, 8])), int(rng.choice([0, 0, 1, 3])), int(rng.integers(0, 1 << 30)))
        assert rc == 0, (case, rc)
        done += 1
    print(f"{done} cases (modes fwd / fwdbwd / bwd / place / step, K in 0/1/4/9/15, uint8 and int64 labels, misaligned label "
          f"pointers, up / identity / down sampling, 1-wide and 1-high tensors): no AddressSanitizer or UBSan report", flush=True)


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    if os.environ.get("HEAD_EMUL_SANITIZE_CHILD") != "1":
        build()
        gxx = lambda name: subprocess.check_output(["g++", "-print-file-name=" + name], text=True).strip()   # noqa: E731
        env = dict(os.environ, HEAD_EMUL_SANITIZE_CHILD="1", LD_PRELOAD=gxx("libasan.so") + ":" + gxx("libubsan.so"),
                   ASAN_OPTIONS="detect_leaks=0:detect_stack_use_after_return=0")
        sys.exit(subprocess.call([sys.executable, os.path.abspath(__file__), str(n), str(seed)], env=env))
    sweep(n, seed)
