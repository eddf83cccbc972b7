"""GPU parity of the fused Placeholder_loss (tools/trainV2_simt.py:202-230 after the upsample of :371-372)
against golden vectors recorded from the reference's own function, and against the oracle on seeded inputs."""
import numpy as np
import pytest
import torch

from oracle import simt_oracle as O
from util import load_golden, rel_l2, rel_max

pytestmark = pytest.mark.gpu
TOL = 1e-5   # fp32 loss / gradient, relative (north_star)


def _run(logits, size, C, K, thres, lam):
    import simt_b200
    lg = torch.as_tensor(logits).cuda().requires_grad_(True)
    loss = simt_b200.Placeholder_loss(lg, C, K, thres, out_size=size, lambda_place=lam)
    loss.backward()
    torch.cuda.synchronize()
    return float(loss), lg.grad.cpu().numpy()


@pytest.mark.parametrize("name", ["place_K4", "place_K15", "place_K4_nothres"])
def test_placeholder_vs_reference_golden(name):
    g = load_golden(name)
    thres = None if float(g["thres"]) < 0 else float(g["thres"])
    loss, dl = _run(g["logits"], tuple(int(s) for s in g["size"]), 19, int(g["K"]), thres, float(g["lambda_place"]))
    # the f64 run of the reference is the truth; its own f32 run is within the same tolerance of it
    assert abs(loss - float(g["loss_f64"])) <= TOL * abs(float(g["loss_f64"]))
    assert rel_l2(dl, g["dlogits_f64"]) <= TOL
    assert rel_max(dl, g["dlogits_f64"]) <= 5 * TOL


@pytest.mark.parametrize("B,K,h,w,H,W,thres,scale", [
    (2, 4, 17, 33, 128, 256, 0.8, 3.0),
    (1, 15, 9, 12, 70, 95, 0.5, 2.0),
    (1, 0, 6, 7, 30, 41, None, 2.0),        # no open-set channels: the open-set target is always class 0
    (1, 4, 5, 9, 5, 9, 0.3, 3.0),           # identity-size "upsample"
    (3, 4, 1, 9, 1, 64, None, 2.0),         # a single row
    (1, 4, 9, 17, 64, 128, 0.8, 40.0),      # huge dynamic range: both soft-maxes use their own exact maximum
    (1, 4, 9, 17, 64, 128, None, 1e-3),     # nearly flat logits
])
def test_placeholder_vs_oracle(B, K, h, w, H, W, thres, scale):
    C = 19
    g = torch.Generator().manual_seed(B * 1000 + K * 10 + h)
    lo = scale * torch.randn(B, C + K, h, w, generator=g)
    lo = lo + 5.0 * scale / 3.0 * torch.nn.functional.one_hot(torch.randint(0, C + K, (B, h, w), generator=g), C + K) \
        .permute(0, 3, 1, 2) * (torch.rand(B, 1, h, w, generator=g) < 0.7)
    l_ref, g_ref = O.placeholder_fwd_bwd(lo, (H, W), C, K, thres, 0.1, torch.float64)
    loss, dl = _run(lo.numpy(), (H, W), C, K, thres, 0.1)
    assert np.isfinite(loss)
    assert abs(loss - float(l_ref)) <= TOL * abs(float(l_ref))
    assert rel_l2(dl, g_ref.numpy()) <= TOL


def test_placeholder_negative_logits_everywhere():
    """All logits far below 0: the constant 0 that replaces the arg-max logit dominates the second soft-max."""
    C, K = 19, 4
    g = torch.Generator().manual_seed(5)
    lo = -60.0 + 3.0 * torch.randn(1, C + K, 7, 9, generator=g)
    l_ref, g_ref = O.placeholder_fwd_bwd(lo, (40, 56), C, K, None, 0.1, torch.float64)
    loss, dl = _run(lo.numpy(), (40, 56), C, K, None, 0.1)
    assert abs(loss - float(l_ref)) <= TOL * abs(float(l_ref))
    assert rel_l2(dl, g_ref.numpy()) <= TOL


def test_placeholder_no_valid_pixel_is_nan():
    """Every arg-max is an open-set channel -> both CE means run over nothing -> NaN, like the reference."""
    C, K = 19, 4
    lo = torch.zeros(1, C + K, 4, 5)
    lo[:, C + 1] = 3.0
    assert bool(torch.isnan(O.placeholder_loss(O.upsample_bilinear_ac(lo, (16, 20)), C, K, 0.8, 0.1)))
    loss, dl = _run(lo.numpy(), (16, 20), C, K, 0.8, 0.1)
    assert np.isnan(loss)


def test_placeholder_full_resolution_image():
    """One image at the training resolution (65x129 -> 512x1024, K = 4): oracle in fp64 on the CPU."""
    C, K, h, w, H, W = 19, 4, 65, 129, 512, 1024
    g = torch.Generator().manual_seed(77)
    lo = 3.0 * torch.randn(1, C + K, h, w, generator=g)
    lo = lo + 6.0 * torch.nn.functional.one_hot(torch.randint(0, C + K, (1, h, w), generator=g), C + K) \
        .permute(0, 3, 1, 2) * (torch.rand(1, 1, h, w, generator=g) < 0.7)
    l_ref, g_ref = O.placeholder_fwd_bwd(lo, (H, W), C, K, 0.8, 0.1, torch.float64)
    loss, dl = _run(lo.numpy(), (H, W), C, K, 0.8, 0.1)
    assert abs(loss - float(l_ref)) <= TOL * abs(float(l_ref))
    assert rel_l2(dl, g_ref.numpy()) <= TOL


def test_placeholder_then_head_share_the_workspace():
    """Placeholder and the T-corrected head run back to back on the same stream / workspace (the training order
    of :398-409) and neither disturbs the other's result."""
    import simt_b200
    C, K, h, w, H, W = 19, 4, 9, 17, 64, 128
    lo, labels = O.synth_head_inputs(2, C + K, h, w, H, W, seed=3, coherent=True, ignore_frac=0.1)
    T = O.sig_ntm_forward(torch.randn(C + K, C, generator=torch.Generator().manual_seed(1)),
                          np.full(19, 1.0 / 19), C, K)
    l_head, dl_head, dT_head = O.simt_head_fwd_bwd(lo, T, labels, (H, W), torch.float64)
    l_pl, g_pl = O.placeholder_fwd_bwd(lo, (H, W), C, K, 0.8, 0.1, torch.float64)
    for _ in range(2):
        x = lo.cuda().requires_grad_(True)
        Tt = T.cuda().requires_grad_(True)
        tot = simt_b200.Placeholder_loss(x, C, K, 0.8, out_size=(H, W), lambda_place=0.1) + \
            simt_b200.simt_head(x, Tt, labels.to(torch.uint8).cuda(), (H, W))
        tot.backward()
        assert abs(float(tot) - float(l_head + l_pl)) <= TOL * abs(float(l_head + l_pl))
        assert rel_l2(x.grad.cpu().numpy(), (dl_head + g_pl).numpy()) <= TOL
        assert rel_l2(Tt.grad.cpu().numpy(), dT_head.numpy()) <= TOL
