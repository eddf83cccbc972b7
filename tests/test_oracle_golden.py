"""CPU: the oracle restatement reproduces the committed reference outputs bit for bit."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import simt_oracle as O
from util import GOLDEN, HEAD_CASES, load_golden, rel_l2

# float results depend on the torch build's CPU kernels; bit-exactness is asserted when the
# installed torch is the one that generated the vectors, otherwise to 1e-6 relative.
_MAN = json.load(open(os.path.join(GOLDEN, "MANIFEST.json")))
SAME_TORCH = _MAN["torch"] == torch.__version__


def _close(x, ref, tol=1e-6):
    x, ref = np.asarray(x), np.asarray(ref)
    if SAME_TORCH and np.array_equal(x, ref, equal_nan=True):
        return True
    return rel_l2(x, ref) <= tol


@pytest.mark.parametrize("name", HEAD_CASES)
@pytest.mark.parametrize("tag,dtype", [("f32", torch.float32), ("f64", torch.float64)])
def test_head_matches_reference(name, tag, dtype):
    g = load_golden(name)
    loss, dl, dT = O.simt_head_fwd_bwd(torch.from_numpy(g["logits"]), torch.from_numpy(g["T"]),
                                       torch.from_numpy(g["labels"]), tuple(g["size"]), dtype)
    assert _close(loss.numpy(), g[f"loss_{tag}"])
    assert _close(dl.numpy(), g[f"dlogits_{tag}"])
    assert _close(dT.numpy(), g[f"dT_{tag}"])


def test_fp32_reference_own_error_is_below_tolerance():
    """How far the fp32 reference is from its fp64 self: the floor for any fp32 implementation."""
    worst = 0.0
    for name in HEAD_CASES:
        g = load_golden(name)
        worst = max(worst, rel_l2(g["dlogits_f32"], g["dlogits_f64"]), rel_l2(g["dT_f32"], g["dT_f64"]))
    assert worst < 1e-5, worst


@pytest.mark.parametrize("K", [0, 4, 15])
def test_sig_ntm_and_w(K):
    g = load_golden(f"ntm_K{K}")
    T = O.sig_ntm_forward(torch.from_numpy(g["NTM"]), np.load(os.path.join(GOLDEN, "ClassDist_bapa.npy")), 19, K)
    assert _close(T.numpy(), g["T"])
    W = O.sig_w_forward(torch.from_numpy(g["W_weight_in"]).clone())
    assert _close(W.numpy(), g["W"])
    assert np.allclose(T.numpy().sum(1), 1.0, atol=1e-6)


def test_cross_entropy_2d_both_modes():
    g = load_golden("ce2d")
    x, y = torch.from_numpy(g["x"]), torch.from_numpy(g["y"])
    for mode in (1, 0):
        xin = (x if mode else torch.softmax(x, 1)).clone().requires_grad_(True)
        l = O.cross_entropy_2d(xin, y, is_softmax=bool(mode))
        l.backward()
        assert _close(l.detach().numpy(), g[f"loss_softmax{mode}"])
        assert _close(xin.grad.numpy(), g[f"grad_softmax{mode}"])


def test_all_ignored_is_nan():
    x = torch.softmax(torch.randn(1, 19, 4, 4), 1)
    y = torch.full((1, 4, 4), 255, dtype=torch.long)
    assert torch.isnan(O.cross_entropy_2d(x, y, is_softmax=False))


@pytest.mark.parametrize("K", [4, 15])
def test_regularisers(K):
    g = load_golden(f"reg_K{K}")
    T1, T2, W1, W2 = (torch.from_numpy(g[k]).requires_grad_(True) for k in ("T1", "T2", "W1", "W2"))
    size = tuple(g["size"])
    p1, p2, fx = (torch.from_numpy(g[k]) for k in ("p1", "p2", "fixed"))
    convex = O.convex_loss([W1, W2], [T1, T2])
    volume = O.volume_loss([T1, T2])
    anchor = O.anchor_loss([O.upsample_bilinear_ac(p1, size), O.upsample_bilinear_ac(p2, size)], [T1, T2],
                           O.label_c_flat(fx, size))
    assert _close(convex.detach().numpy(), g["convex"])
    assert _close(volume.detach().numpy(), g["volume"])
    assert _close(anchor.detach().numpy(), g["anchor"])
    (0.1 * convex + volume + anchor).backward()
    for k, t in (("dT1", T1), ("dT2", T2), ("dW1", W1), ("dW2", W2)):
        assert _close(t.grad.numpy(), g[k], 1e-5), k
    a1, e1 = O.anchor_stats(O.upsample_bilinear_ac(p1, size))
    assert np.array_equal(a1.numpy(), g["anchor_idx1"]) and np.array_equal(e1.numpy(), g["exist1"])


def test_histograms_bit_exact():
    g = load_golden("hist")
    mapping = g["mapping"]
    assert mapping.tolist() == O.CITYSCAPES_LABEL2TRAIN
    hist = np.zeros((19, 19), dtype=np.int64)
    for gt, pr in zip(g["gt"], g["pred"]):
        lab = O.label_mapping(gt, mapping)
        assert lab.dtype == np.int64
        hist += O.fast_hist(lab.flatten(), pr.flatten(), 19)
    assert np.array_equal(hist, g["hist19"])
    assert O.miou_percent(hist) == float(g["miou"])
    assert np.array_equal(np.nan_to_num(O.per_class_iu(hist), nan=-1), np.nan_to_num(g["iu"], nan=-1))
    assert np.array_equal(O.fast_hist_rect(g["gt"][0].flatten(), g["pred"][0].flatten(), 34, 19), g["rect34x19"])
    assert np.array_equal(O.class_hist(g["pred"][0].flatten(), 19), g["class19"])


def test_fast_hist_aliasing_and_errors():
    """Reference quirk: b is unchecked -- b >= n aliases, an index past n^2 raises."""
    a = np.array([0, 1, 2], dtype=np.int64)
    assert O.fast_hist(a, np.array([0, 20, 1]), 19)[2, 1] == 2          # 19*1+20 == 19*2+1
    with pytest.raises(ValueError):
        O.fast_hist(np.array([18]), np.array([30]), 19)


def test_pseudo_labels_rule():
    g = load_golden("pseudo_K4")
    size = tuple(int(s) for s in g["size"])
    conf = O.pseudo_labels(torch.from_numpy(g["output2"]), O.upsample_bilinear_ac(torch.from_numpy(g["pred2_lo"]), size),
                           size, 19, 0.8, 0.2)
    assert (conf.numpy() != g["conf"]).mean() < 1e-4       # identical up to float near-ties across torch builds
    vals = set(np.unique(g["conf"]).tolist())
    assert 255 in vals and any(v < 19 for v in vals) and any(19 <= v < 255 for v in vals)


@pytest.mark.parametrize("K", [4, 15])
def test_inner_w_loop_matches_reference(K):
    """trainV2_simt.py:326-339 (10 rounds of sig_W forward, ||W T||^2, backward, torch Adam), two outer
    iterations, against weights / Adam state / accumulated NTM grads recorded from the reference's modules."""
    g = load_golden(f"wfit_K{K}")
    CK = 19 + K
    cd = np.load(os.path.join(GOLDEN, "ClassDist_bapa.npy"))
    ntm = [torch.from_numpy(g["ntm1"]).clone(), torch.from_numpy(g["ntm2"]).clone()]
    w = [torch.from_numpy(g["w1_init"]).clone(), torch.from_numpy(g["w2_init"]).clone()]
    st = [dict(m=torch.zeros(CK, CK), v=torch.zeros(CK, CK), step=0) for _ in range(2)]
    off = ~np.eye(CK, dtype=bool)
    for outer in range(2):
        grads, losses = O.w_fit_loop(ntm, w, st, float(g["lr"]), cd, 19, K, rounds=int(g["rounds"]))
        assert _close(losses.numpy(), g[f"losses_outer{outer}"], 1e-5)
        for i in range(2):
            assert st[i]["step"] == 10 * (outer + 1)
            assert np.all(w[i].numpy()[~off] == -10000.0)
            assert _close(w[i].numpy()[off], g[f"w{i + 1}_after{outer}"][off], 1e-6)
            assert _close(st[i]["m"].numpy(), g[f"m{i + 1}_after{outer}"], 1e-5)
            assert _close(st[i]["v"].numpy(), g[f"v{i + 1}_after{outer}"], 1e-5)
            assert _close(grads[i].numpy(), g[f"ntm_grad{i + 1}_outer{outer}"], 1e-5)


@pytest.mark.parametrize("name", ["place_K4", "place_K15", "place_K4_nothres"])
@pytest.mark.parametrize("tag,dtype", [("f32", torch.float32), ("f64", torch.float64)])
def test_placeholder_loss_matches_reference(name, tag, dtype):
    """Placeholder_loss (trainV2_simt.py:202-230) after the upsample of :371, against values recorded from the
    reference's own function (compiled from its source, see oracle/make_golden.py)."""
    g = load_golden(name)
    thres = None if float(g["thres"]) < 0 else float(g["thres"])
    loss, dl = O.placeholder_fwd_bwd(torch.from_numpy(g["logits"]), tuple(int(s) for s in g["size"]), 19, int(g["K"]),
                                     thres, float(g["lambda_place"]), dtype)
    assert _close(loss.numpy(), g[f"loss_{tag}"], 1e-6)
    assert _close(dl.numpy(), g[f"dlogits_{tag}"], 1e-6)


def test_placeholder_quirks():
    """The constant that replaces the arg-max logit is 0 (``-1000. * zeros_like``), and the open-set target falls
    back to class 0 when no open-set logit is positive."""
    z = torch.tensor([2.0, 1.0, -3.0, -0.5, -0.25]).view(1, 5, 1, 1)     # C = 3 known, K = 2 open (both <= 0)
    loss = O.placeholder_loss(z, 3, 2, None, 1.0)
    known = torch.logsumexp(z.flatten(), 0) - 2.0
    zp = z.flatten().clone(); zp[0] = 0.0
    unknown = torch.logsumexp(zp, 0) - zp[0]                               # y = 0 -> the replaced (constant 0) logit
    assert abs(float(loss) - float(known + unknown)) < 1e-6
    z[0, 4] = 0.75                                                         # now an open-set logit is > 0 -> y = 4
    loss = O.placeholder_loss(z, 3, 2, None, 1.0)
    zp = z.flatten().clone(); zp[0] = 0.0
    assert abs(float(loss) - float(known * 0 + torch.logsumexp(z.flatten(), 0) - 2.0 + torch.logsumexp(zp, 0) - 0.75)) < 1e-6
