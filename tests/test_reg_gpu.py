"""GPU parity of T's volume / anchor / convex regularisers against golden vectors produced by the
reference's own lines (tools/trainV2_simt.py:354-357,375-384,412-424; see oracle/make_golden.py)."""
import numpy as np
import pytest
import torch

from util import load_golden, rel_l2, rel_max

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.mark.parametrize("K", [4, 15])
def test_fused_convex_volume_kernel(K):
    import simt_b200
    g = load_golden(f"reg_K{K}")
    dev = torch.device("cuda")
    T1, T2, W1, W2 = (torch.from_numpy(g[k]).to(dev).requires_grad_(True) for k in ("T1", "T2", "W1", "W2"))
    c1, v1 = simt_b200.t_regularizers(T1, W1)
    c2, v2 = simt_b200.t_regularizers(T2, W2)
    convex, volume = c1 + c2, v1 + v2
    assert abs(float(convex) - float(g["convex"])) <= TOL * abs(float(g["convex"]))
    assert abs(float(volume) - float(g["volume"])) <= TOL * abs(float(g["volume"]))
    # anchor through the kernels, then the reference's weighting (sh_simt.sh:16): 0.1, 1, 1
    size = tuple(int(s) for s in g["size"])
    p1, p2, fx = (torch.from_numpy(g[k]).to(dev) for k in ("p1", "p2", "fixed"))
    anchor = simt_b200.anchor_loss([p1, p2], [T1, T2], fx, size)
    assert abs(float(anchor) - float(g["anchor"])) <= TOL * abs(float(g["anchor"]))
    (0.1 * convex + 1.0 * volume + 1.0 * anchor).backward()
    for k, t in (("dT1", T1), ("dT2", T2), ("dW1", W1), ("dW2", W2)):
        assert rel_l2(t.grad.cpu().numpy(), g[k]) <= TOL, k
        assert rel_max(t.grad.cpu().numpy(), g[k]) <= TOL, k


@pytest.mark.parametrize("K", [4, 15])
def test_anchor_stats_match_reference_indices(K):
    import simt_b200
    g = load_golden(f"reg_K{K}")
    dev = torch.device("cuda")
    size = tuple(int(s) for s in g["size"])
    for tag in ("1", "2"):
        idx, exist = simt_b200.anchor_stats(torch.from_numpy(g["p" + tag]).to(dev), size)
        assert np.array_equal(idx.cpu().numpy(), g["anchor_idx" + tag])
        assert np.array_equal(np.nonzero(exist.cpu().numpy())[0], g["exist" + tag])


def test_volume_guard_on_singular_T():
    """Rank-deficient T: det(T^T T) = 0 -> log = -inf -> the reference replaces the term by 0."""
    import simt_b200
    dev = torch.device("cuda")
    T = torch.zeros(23, 19, device=dev, requires_grad=True)
    conv, vol = simt_b200.t_regularizers(T, torch.eye(23, device=dev))
    assert float(vol) == 0.0 and float(conv) == 0.0
    vol.backward()
    assert float(T.grad.abs().max()) == 0.0
    # torch-expression form applies the same guard without a host sync
    assert float(simt_b200.volume_loss([T.detach()])) == 0.0


def test_torch_expression_forms_on_gpu():
    import simt_b200
    g = load_golden("reg_K4")
    dev = torch.device("cuda")
    T1, T2, W1, W2 = (torch.from_numpy(g[k]).to(dev) for k in ("T1", "T2", "W1", "W2"))
    assert abs(float(simt_b200.convex_loss([W1, W2], [T1, T2])) - float(g["convex"])) <= TOL * abs(float(g["convex"]))
    assert abs(float(simt_b200.volume_loss([T1, T2])) - float(g["volume"])) <= TOL * abs(float(g["volume"]))
    assert abs(float(simt_b200.w_fit_loss([W1, W2], [T1, T2])) + float(g["convex"])) <= TOL * abs(float(g["convex"]))
