import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

HEAD_CASES = ["head_small_u", "head_small_r", "head_openset4", "head_openset15", "head_odd",
              "head_identity", "head_down", "head_row1", "head_cfg1_tile"]


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def rel_l2(x, ref):
    x = np.asarray(x, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    den = np.linalg.norm(ref.ravel())
    return float(np.linalg.norm((x - ref).ravel()) / (den if den > 0 else 1.0))


def rel_max(x, ref):
    x = np.asarray(x, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    den = np.abs(ref).max()
    return float(np.abs(x - ref).max() / (den if den > 0 else 1.0))


def class_dist():
    return np.load(os.path.join(GOLDEN, "ClassDist_bapa.npy"))


def run_gpu_head(logits, T, labels, size, int64_labels=False):
    """CUDA path through the public API (-> ctypes -> C ABI); returns numpy (loss, dlogits, dT)."""
    import simt_b200
    dev = torch.device("cuda")
    lg = torch.as_tensor(logits).to(dev).requires_grad_(True)
    Tt = None if T is None else torch.as_tensor(T).to(dev).requires_grad_(True)
    lab = torch.as_tensor(labels)
    lab = (lab.long() if int64_labels else lab.to(torch.uint8)).to(dev)
    loss = simt_b200.simt_head(lg, Tt, lab, tuple(int(s) for s in size))
    loss.backward()
    torch.cuda.synchronize()
    return (loss.detach().cpu().numpy(), lg.grad.cpu().numpy(), None if Tt is None else Tt.grad.cpu().numpy())
