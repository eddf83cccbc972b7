"""The oracle's restatements against the reference's OWN files in oracle/_ref (made by oracle/make_ref.py from
/root/reference; skipped where that copy does not exist).  CPU only."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import make_ref, simt_oracle as O  # noqa: E402

make_ref.make(verbose=False)   # a no-op where /root/reference does not exist (GPU box): oracle/_ref then travels with the tree
CE, CIOU = make_ref.load()
pytestmark = pytest.mark.skipif(CE is None, reason="oracle/_ref not made (no /root/reference here)")


@pytest.mark.parametrize("is_softmax", [True, False])
def test_cross_entropy_2d_restatement_equals_reference_class(is_softmax):
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 19, 24, 40, generator=g)
    if not is_softmax:
        x = torch.softmax(x, dim=1)
    y = torch.randint(0, 19, (2, 24, 40), generator=g)
    y[torch.rand(2, 24, 40, generator=g) < 0.2] = 255
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    ref = CE(is_softmax=is_softmax)(xa, y)
    got = O.cross_entropy_2d(xb, y, is_softmax=is_softmax)
    ref.backward(); got.backward()
    assert torch.equal(ref, got) and torch.equal(xa.grad, xb.grad)


def test_head_loss_with_reference_class_equals_restatement():
    logits, labels = O.synth_head_inputs(1, 19, 9, 17, 64, 128, seed=5, coherent=True, block=8)
    T = O.sig_ntm_forward(torch.randn(19, 19, generator=torch.Generator().manual_seed(1)), np.full(19, 1 / 19), 19, 0)
    a = O.simt_head_loss(logits, T, labels.long(), (64, 128))
    b = O.simt_head_loss(logits, T, labels.long(), (64, 128), ce=CE(is_softmax=False))
    assert torch.equal(a, b)


def test_histogram_restatements_equal_reference_functions():
    gt, pr = O.synth_eval_pair(128, 256, seed=2)
    m = np.array(O.CITYSCAPES_LABEL2TRAIN)
    assert np.array_equal(O.label_mapping(gt, m), CIOU.label_mapping(gt, m))
    a = O.label_mapping(gt, m).flatten()
    assert np.array_equal(O.fast_hist(a, pr.flatten(), 19), CIOU.fast_hist(a, pr.flatten(), 19))
    h = O.fast_hist(a, pr.flatten(), 19)
    with np.errstate(divide="ignore", invalid="ignore"):
        assert np.array_equal(O.per_class_iu(h), CIOU.per_class_iu(h), equal_nan=True)


def test_weighted_cross_entropy_composition_equals_reference_class():
    """CrossEntropy2d(weight=) (utils/loss.py:14,36,39): the torch-op composition simt_b200 uses for this rare form."""
    from simt_b200.loss import _weighted_ce2d
    g = torch.Generator().manual_seed(9)
    w = torch.rand(19, generator=g) + 0.1
    for is_softmax in (True, False):
        x = torch.randn(2, 19, 12, 20, generator=g)
        if not is_softmax:
            x = torch.softmax(x, dim=1)
        y = torch.randint(0, 19, (2, 12, 20), generator=g)
        y[torch.rand(2, 12, 20, generator=g) < 0.25] = 255
        xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
        ref = CE(is_softmax=is_softmax)(xa, y, weight=w)
        got = _weighted_ce2d(xb, y, w, 255, is_softmax)
        ref.backward(); got.backward()
        assert abs(float(ref) - float(got)) <= 1e-6 * abs(float(ref))
        assert float((xa.grad - xb.grad).abs().max()) <= 1e-6 * float(xa.grad.abs().max())


def test_generic_label_mapping_equals_reference_function():
    """label_mapping on non-uint8 integer images (tools/compute_iou.py:18-22 takes any integer array)."""
    from simt_b200.hist import _label_mapping_generic
    rng = np.random.default_rng(4)
    m = np.array(O.CITYSCAPES_LABEL2TRAIN)
    for dt in (np.int16, np.int32, np.int64):
        x = rng.integers(-1, 40, size=(37, 53)).astype(dt)
        got = _label_mapping_generic(torch.from_numpy(x.astype(np.int64)), [(int(a), int(b)) for a, b in m.tolist()]).numpy()
        assert np.array_equal(got, CIOU.label_mapping(x, m))
