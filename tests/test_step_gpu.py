"""GPU: one whole training iteration of the head (tools/trainV2_simt.py:351-424 minus the backbone, with
Placeholder_loss) through the simt_b200 API against the CPU oracle's restatement of the same lines:
total loss and every gradient that leaves the head (both heads' low-res logits, both NTM parameters, both
W parameters)."""
import numpy as np
import pytest
import torch

from util import class_dist, rel_l2

pytestmark = pytest.mark.gpu


def _make(K, seed, dev=None):
    import simt_b200
    torch.manual_seed(seed)
    ntm1, ntm2 = simt_b200.sig_NTM(19, K), simt_b200.sig_NTM(19, K)
    w1, w2 = simt_b200.sig_W(19, K), simt_b200.sig_W(19, K)
    with torch.no_grad():
        w1.weight.add_(0.1 * torch.randn_like(w1.weight))
        w2.weight.add_(0.1 * torch.randn_like(w2.weight))
    mods = [ntm1, ntm2, w1, w2]
    if dev is not None:
        mods = [m.to(dev) for m in mods]
    return mods


@pytest.mark.parametrize("K", [4, 15])
def test_training_iteration_head_matches_reference_lines(K):
    import simt_b200
    from oracle import simt_oracle as O
    dev = torch.device("cuda")
    CK, h, w, H, W = 19 + K, 17, 33, 128, 256
    g = torch.Generator().manual_seed(100 + K)
    pred1 = 2.0 * torch.randn(1, CK, h, w, generator=g)
    pred2 = 2.0 * torch.randn(1, CK, h, w, generator=g)
    out2 = 2.0 * torch.randn(1, 19, h, w, generator=g)
    for p in (pred1, pred2):   # confident nodes, so that Placeholder_loss's 0.8 threshold keeps a share of the pixels
        p += 5.0 * torch.nn.functional.one_hot(torch.randint(0, CK, (1, h, w), generator=g), CK).permute(0, 3, 1, 2) * \
            (torch.rand(1, 1, h, w, generator=g) < 0.6)
    _, labels = O.synth_head_inputs(1, CK, h, w, H, W, seed=3, coherent=True, block=(20, 28), class_dist=class_dist())

    # ---- oracle (CPU, fp64 for a clean comparison) ----
    ntm1, ntm2, w1, w2 = [m.double() for m in _make(K, 7)]
    p1o, p2o = pred1.double().requires_grad_(True), pred2.double().requires_grad_(True)
    ref = O.training_step_loss(p1o, p2o, out2.double(), labels.long(), ntm1(), ntm2(), w1(), w2(), (H, W), 19,
                               lambda_place=0.1)
    ref.backward()

    # ---- product (GPU) ----
    m1, m2, v1, v2 = _make(K, 7, dev)
    p1, p2 = pred1.to(dev).requires_grad_(True), pred2.to(dev).requires_grad_(True)
    lab = labels.to(dev)
    T1, T2, W1, W2 = m1(), m2(), v1(), v2()
    conf = simt_b200.pseudo_labels(out2.to(dev), p2, (H, W), 19, 0.8, 0.2)             # :351-365,387-393
    loss_p1 = simt_b200.simt_head(p1, None, conf, (H, W))                                # :394
    loss_p2 = simt_b200.simt_head(p2, None, conf, (H, W))                                # :395
    loss_y1 = simt_b200.simt_head(p1, T1, lab, (H, W))                                   # :371,402-403,408
    loss_y2 = simt_b200.simt_head(p2, T2, lab, (H, W))                                   # :372,405-406,409
    c1, vol1 = simt_b200.t_regularizers(T1, W1)                                          # :412-421
    c2, vol2 = simt_b200.t_regularizers(T2, W2)
    anchor = simt_b200.anchor_loss([p1, p2], [T1, T2], out2.to(dev), (H, W))             # :375-384
    place = 0.1 * simt_b200.Placeholder_loss(p1, 19, K, 0.8, out_size=(H, W), lambda_place=0.1)   # :398
    place = place + simt_b200.Placeholder_loss(p2, 19, K, 0.8, out_size=(H, W), lambda_place=0.1)  # :399
    total = place + (loss_p2 + loss_y2 + 0.1 * loss_p1 + 0.1 * loss_y1) + 0.1 * (c1 + c2) + 1.0 * (vol1 + vol2) + 1.0 * anchor
    total.backward()
    simt_b200.check_errors()

    assert abs(float(total) - float(ref)) <= 2e-5 * abs(float(ref))
    for name, got, want in (("dpred1", p1.grad, p1o.grad), ("dpred2", p2.grad, p2o.grad),
                            ("dNTM1", m1.NTM.grad, ntm1.NTM.grad), ("dNTM2", m2.NTM.grad, ntm2.NTM.grad),
                            ("dW1", v1.weight.grad, w1.weight.grad), ("dW2", v2.weight.grad, w2.weight.grad)):
        assert rel_l2(got.cpu().numpy(), want.numpy()) <= 2e-5, name
