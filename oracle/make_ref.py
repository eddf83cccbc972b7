"""Copy recipe for oracle/_ref/: the reference's OWN source files of this hot path, taken unmodified from
/root/reference so that they travel to the GPU box (which has no /root/reference) and can be timed / compared there.

    python oracle/make_ref.py          (also run by __graft_entry__.build() when /root/reference is present)

TEST / BENCH INFRASTRUCTURE ONLY (see oracle/simt_oracle.py): nothing under simt_b200/ imports oracle/_ref.
oracle/_ref/ is git-ignored (the reference's sources are never committed here) but NOT gpurun-ignored.

  utils/loss.py            -> _ref/loss.py           CrossEntropy2d, EntropyLoss (utils/loss.py:6-49)
  tools/compute_iou.py     -> _ref/compute_iou.py    fast_hist, per_class_iu, label_mapping (:9-22)
  tools/_init_paths.py     -> _ref/_init_paths.py    (imported by compute_iou.py; a sys.path insert, harmless)
The rest of the path (tools/trainV2_simt.py:371-372,402-409; compute_ConfusionMatrix.py:54-56;
compute_ClassDistribution.py:52-54) lives inside scripts that parse the command line / import absent modules at import
time and is restated in oracle/simt_oracle.py.
"""
import hashlib
import os
import shutil
import sys

REF = os.environ.get("SIMT_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
FILES = {"utils/loss.py": "loss.py", "tools/compute_iou.py": "compute_iou.py", "tools/_init_paths.py": "_init_paths.py"}


def make(verbose=True) -> bool:
    if not os.path.isdir(REF):
        if verbose:
            print(f"{REF} not present: oracle/_ref left as is ({'present' if os.path.isdir(DST) else 'absent'})")
        return os.path.isdir(DST)
    os.makedirs(DST, exist_ok=True)
    lines = []
    for src, dst in FILES.items():
        shutil.copyfile(os.path.join(REF, src), os.path.join(DST, dst))
        sha = hashlib.sha256(open(os.path.join(DST, dst), "rb").read()).hexdigest()[:16]
        lines.append(f"{dst}  <-  {src}  sha256/16 {sha}")
    open(os.path.join(DST, "MANIFEST.txt"), "w").write("\n".join(lines) + "\n")
    if verbose:
        print("oracle/_ref:", "; ".join(lines))
    return True


def load():
    """(CrossEntropy2d class, compute_iou module) from oracle/_ref, or (None, None) when it was never made."""
    if not os.path.exists(os.path.join(DST, "loss.py")):
        return None, None
    import importlib.util
    if DST not in sys.path:
        sys.path.insert(0, DST)          # compute_iou.py does `import _init_paths`
    mods = []
    for name in ("loss", "compute_iou"):
        spec = importlib.util.spec_from_file_location(f"simt_reference_{name}", os.path.join(DST, f"{name}.py"))
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        mods.append(m)
    return mods[0].CrossEntropy2d, mods[1]


if __name__ == "__main__":
    make()
