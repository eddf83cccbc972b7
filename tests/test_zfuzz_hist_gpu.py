"""Randomised GPU sweep of the integer histograms against numpy.  pytest (-m gpu) runs a bounded number of cases;
by hand for a longer sweep:  python tests/test_zfuzz_hist_gpu.py [n_cases] [seed]
Random sizes (0 .. 6 M pixels, ragged tails), misaligned views, uint8 / int64 inputs, value ranges that include
out-of-range ground truth (masked like the reference) and the three arities of fast_hist; bit-exact or bust."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import simt_b200  # noqa: E402
from simt_b200 import hist as H  # noqa: E402
from oracle import simt_oracle as O  # noqa: E402
import pytest  # noqa: E402


def run_fuzz(n_cases, seed):
    rng = np.random.default_rng(seed)
    dev = torch.device("cuda")
    fails = []
    for case in range(n_cases):
        n = int(rng.choice([0, 1, 15, 16, 17, 4095, 4096, 65537, int(rng.integers(1, 6_000_000))]))
        off = int(rng.choice([0, 0, 1, 3, 16, 20]))                # element offset of the view: misaligned pointers
        i64 = bool(rng.integers(0, 4) == 0)
        runs = bool(rng.integers(0, 2))                            # coherent runs or salt-and-pepper
        kind = rng.choice(["square", "rect", "class", "lut"])
        hi_a = 34 if kind in ("rect", "lut") else int(rng.choice([19, 25, 256]))   # gt ids beyond n are masked out
        def draw(hi):
            if runs and n > 0:
                seg = rng.integers(1, 700, size=max(1, n // 200 + 1))
                vals = rng.integers(0, hi, size=seg.size)
                return np.repeat(vals, seg)[:n] if np.repeat(vals, seg).size >= n else np.resize(np.repeat(vals, seg), n)
            return rng.integers(0, hi, size=n)
        a = draw(hi_a).astype(np.int64 if i64 else np.uint8)
        b = draw(19).astype(np.int64 if i64 else np.uint8)
        ta = torch.from_numpy(np.concatenate([np.zeros(off, a.dtype), a])).to(dev)[off:]
        tb = torch.from_numpy(np.concatenate([np.zeros(off, b.dtype), b])).to(dev)[off:]
        tag = f"case {case}: {kind} n={n} off={off} int64={i64} runs={runs}"
        try:
            if kind == "square":
                got = H.fast_hist(ta, tb, 19).cpu().numpy()
                ref = O.fast_hist(a.astype(np.int64), b.astype(np.int64), 19)
            elif kind == "rect":
                got = H.fast_hist(ta, tb, 34, 19).cpu().numpy()
                ref = O.fast_hist_rect(a.astype(np.int64), b.astype(np.int64), 34, 19)
            elif kind == "class":
                got = H.fast_hist(tb, 19).cpu().numpy()
                ref = O.class_hist(b.astype(np.int64), 19)
            else:
                if i64:
                    continue
                m = simt_b200.ConfusionMeter(19, mapping=O.CITYSCAPES_LABEL2TRAIN, device=dev)
                m.update(ta, tb)
                got = m.value()
                ref = O.fast_hist(O.label_mapping(a, np.array(O.CITYSCAPES_LABEL2TRAIN)).flatten(), b.astype(np.int64), 19)
            simt_b200.check_errors(dev)
            if not np.array_equal(np.asarray(got).reshape(ref.shape), ref):
                fails.append(f"{tag}: mismatch, |diff| sum {np.abs(np.asarray(got).reshape(ref.shape) - ref).sum()}")
        except Exception as e:   # noqa: BLE001
            fails.append(f"{tag}: EXCEPTION {type(e).__name__}: {e}")
    return fails


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [0])
def test_fuzz_hist(seed):
    fails = run_fuzz(100, seed)
    assert not fails, "\n".join(fails)


if __name__ == "__main__":
    n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 80
    fails = run_fuzz(n_cases, int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    print(f"{n_cases} histogram cases, {len(fails)} failures")
    for f in fails:
        print("FAIL", f)
    sys.exit(1 if fails else 0)
