"""CPU: host-side logic and the C-ABI surface (no compute calls)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import simt_b200
from simt_b200 import _lib
from util import GOLDEN, load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(built_lib):
    hdr = open(os.path.join(ROOT, "include", "simt_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(simt_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 14
    lib = ctypes.CDLL(built_lib)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/simt_b200.h but not exported"
    # and the Python binding names only what the header declares
    assert set(_lib.SIGNATURES) <= declared


def test_abi_version_and_strerror(built_lib):
    lib = _lib.load()
    assert lib.simt_b200_abi_version() == 1
    assert b"invalid" in lib.simt_b200_strerror(-1)
    assert b"workspace" in lib.simt_b200_strerror(-3)


def test_product_path_has_no_cpu_fallback():
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        simt_b200.simt_head(torch.zeros(1, 19, 5, 9), None, torch.zeros(1, 32, 64, dtype=torch.uint8))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        simt_b200.CrossEntropy2d()(torch.zeros(1, 19, 4, 4), torch.zeros(1, 4, 4, dtype=torch.long))
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            simt_b200.fast_hist(np.zeros(4, dtype=np.uint8), np.zeros(4, dtype=np.uint8), 19)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "simt_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("the oracle", ""), f"{f} mentions the oracle"


@pytest.mark.parametrize("K", [0, 4, 15])
def test_sig_ntm_matches_reference_init_and_forward(K):
    """Same seed -> same kaiming init, T and dNTM as model/deeplab_multi.py:244-263 (golden from the reference)."""
    g = load_golden(f"ntm_K{K}")
    torch.manual_seed(1234 + K)
    ntm = simt_b200.sig_NTM(19, K)
    assert list(dict(ntm.named_parameters())) == ["NTM"]
    assert np.array_equal(ntm.NTM.detach().numpy(), g["NTM"])
    T = ntm()
    assert np.array_equal(T.detach().numpy(), g["T"])
    T.sum().backward()
    assert np.array_equal(ntm.NTM.grad.numpy(), g["dNTM_of_sumT"])


@pytest.mark.parametrize("K", [0, 4, 15])
def test_sig_w_matches_reference(K):
    g = load_golden(f"ntm_K{K}")
    w = simt_b200.sig_W(19, K)
    assert list(dict(w.named_parameters())) == ["weight"]
    assert torch.allclose(w.weight.detach(), torch.full((19 + K, 19 + K), 1.0 / (19 + K - 1.0)))
    with torch.no_grad():
        w.weight.copy_(torch.from_numpy(g["W_weight_in"]))
    W = w()
    assert np.array_equal(W.detach().numpy(), g["W"])
    assert float(w.weight.detach().diagonal().max()) == -10000.0     # in-place diag like the reference


def test_build_lut_equals_label_mapping():
    from oracle import simt_oracle as O
    lut = simt_b200.build_lut(O.CITYSCAPES_LABEL2TRAIN)
    allv = np.arange(256, dtype=np.uint8)
    assert np.array_equal(lut.astype(np.int64), O.label_mapping(allv, np.array(O.CITYSCAPES_LABEL2TRAIN)))
    # a mapping whose targets collide with later keys still commutes (matches are on the original input)
    m = [[1, 2], [2, 3], [3, 1]]
    assert np.array_equal(simt_b200.build_lut(m).astype(np.int64), O.label_mapping(allv, np.array(m)))


def test_per_class_iu_host():
    from oracle import simt_oracle as O
    g = load_golden("hist")
    iu = simt_b200.per_class_iu(g["hist19"])
    assert np.array_equal(np.nan_to_num(iu, nan=-1), np.nan_to_num(g["iu"], nan=-1))


def test_torch_regulariser_expressions_match_reference():
    for K in (4, 15):
        g = load_golden(f"reg_K{K}")
        T1, T2, W1, W2 = (torch.from_numpy(g[k]).requires_grad_(True) for k in ("T1", "T2", "W1", "W2"))
        convex = simt_b200.convex_loss([W1, W2], [T1, T2])
        volume = simt_b200.volume_loss([T1, T2])
        assert abs(float(convex) - float(g["convex"])) <= 1e-5 * abs(float(g["convex"]))
        assert abs(float(volume) - float(g["volume"])) <= 1e-5 * abs(float(g["volume"]))


def test_workload_generators_are_shared_not_oracle_code():
    """bench.py's GPU leg and the profiling scripts take their seeded inputs from simt_b200.synth, never from oracle/;
    the oracle re-exports the same generators for the tests, and the Cityscapes table is the same data in both."""
    import re
    from oracle import simt_oracle as O
    from simt_b200 import synth
    assert O.synth_head_inputs is synth.synth_head_inputs and O.synth_eval_pair is synth.synth_eval_pair
    assert O.CITYSCAPES_LABEL2TRAIN == synth.CITYSCAPES_LABEL2TRAIN
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for f in os.listdir(os.path.join(root, "scripts")):
        if f.endswith(".py"):
            src = open(os.path.join(root, "scripts", f)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f"scripts/{f} imports oracle/"
    for f in os.listdir(os.path.join(root, "simt_b200")):
        if f.endswith(".py"):
            src = open(os.path.join(root, "simt_b200", f)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f"simt_b200/{f} imports oracle/"
