"""GPU parity of the fused head: CUDA path (public API -> ctypes -> C ABI) vs golden vectors
from the reference and vs the CPU oracle on seeded inputs.

Tolerance (BASELINE.json north_star): loss, dLogits, dT within 1e-5 relative of the fp32
reference -- norm-wise (||d||_2 / ||ref||_2) and max-abs-scaled, since element-wise relative
error is ill-posed at zeros.
"""
import numpy as np
import pytest
import torch

from util import HEAD_CASES, class_dist, load_golden, rel_l2, rel_max, run_gpu_head

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _check(got, ref, what):
    assert rel_l2(got, ref) <= TOL, f"{what}: rel l2 {rel_l2(got, ref):.3e}"
    assert rel_max(got, ref) <= TOL, f"{what}: rel max {rel_max(got, ref):.3e}"


@pytest.mark.parametrize("name", HEAD_CASES)
@pytest.mark.parametrize("int64_labels", [False, True])
def test_head_vs_reference_golden(name, int64_labels):
    g = load_golden(name)
    loss, dl, dT = run_gpu_head(g["logits"], g["T"], g["labels"], g["size"], int64_labels)
    assert abs(float(loss) - float(g["loss_f32"])) <= TOL * abs(float(g["loss_f32"]))
    _check(dl, g["dlogits_f32"], "dlogits vs fp32 reference")
    _check(dT, g["dT_f32"], "dT vs fp32 reference")
    # and at least as close to the fp64 truth as 1e-5
    _check(dl, g["dlogits_f64"], "dlogits vs fp64 reference")
    _check(dT, g["dT_f64"], "dT vs fp64 reference")


@pytest.mark.parametrize("lpr", [2, 4])
@pytest.mark.parametrize("tile", [(1, 0), (2, 0), (3, 0), (8, 0), (32, 0)])
def test_head_tilings_and_lane_splits_agree(lpr, tile):
    """Every unit height (cell-rows per warp unit) / lanes-per-cell configuration is the same function."""
    from simt_b200 import _lib
    g = load_golden("head_cfg1_tile")
    lib = _lib.load()
    try:
        lib.simt_head_set_tuning(tile[0], tile[1], 0, lpr)
        loss, dl, dT = run_gpu_head(g["logits"], g["T"], g["labels"], g["size"])
    finally:
        lib.simt_head_set_tuning(0, 0, 0, 0)
    assert abs(float(loss) - float(g["loss_f32"])) <= TOL * abs(float(g["loss_f32"]))
    _check(dl, g["dlogits_f32"], "dlogits")
    _check(dT, g["dT_f32"], "dT")


@pytest.mark.parametrize("B,K,coherent", [(1, 0, True), (1, 0, False), (2, 4, True), (1, 15, True)])
def test_head_config_shapes_vs_oracle(B, K, coherent):
    """BASELINE configs 1-3 at full 65x129 -> 512x1024 resolution against the CPU oracle (seconds)."""
    from oracle import simt_oracle as O
    CK = 19 + K
    logits, labels = O.synth_head_inputs(B, CK, 65, 129, 512, 1024, seed=1234, coherent=coherent,
                                         class_dist=class_dist())
    torch.manual_seed(1234)
    NTM = torch.nn.init.kaiming_normal_(torch.ones(CK, 19), mode="fan_out", nonlinearity="relu")
    T = O.sig_ntm_forward(NTM, class_dist(), 19, K)
    lo, dlo, dTo = O.simt_head_fwd_bwd(logits, T, labels, (512, 1024), torch.float32)
    loss, dl, dT = run_gpu_head(logits.numpy(), T.numpy(), labels.numpy(), (512, 1024))
    assert abs(float(loss) - float(lo)) <= TOL * abs(float(lo))
    _check(dl, dlo.numpy(), "dlogits")
    _check(dT, dTo.numpy(), "dT")


def test_plain_ce_T_none_vs_oracle():
    """T = None is torch's CrossEntropyLoss(ignore_index=255) on the upsampled logits (trainV2_simt.py:394)."""
    from oracle import simt_oracle as O
    logits, labels = O.synth_head_inputs(2, 19, 9, 17, 64, 128, seed=5, coherent=True, block=8)
    lg = logits.clone().requires_grad_(True)
    ref = O.plain_ce_loss(lg, labels.long(), (64, 128))
    ref.backward()
    loss, dl, _ = run_gpu_head(logits.numpy(), None, labels.numpy(), (64, 128))
    assert abs(float(loss) - float(ref)) <= TOL * abs(float(ref))
    _check(dl, lg.grad.numpy(), "dlogits")


def test_forward_only_and_two_pass_backward_agree_with_single_pass():
    import simt_b200
    from simt_b200 import _lib, head
    g = load_golden("head_openset4")
    dev = torch.device("cuda")
    lg = torch.from_numpy(g["logits"]).to(dev)
    T = torch.from_numpy(g["T"]).to(dev)
    lab = torch.from_numpy(g["labels"]).to(dev)
    size = tuple(int(s) for s in g["size"])
    with torch.no_grad():
        l_fwd = simt_b200.simt_head(lg, T, lab, size)
    assert abs(float(l_fwd) - float(g["loss_f32"])) <= TOL * abs(float(g["loss_f32"]))
    # two-pass C entry: scale known on the host
    lib = _lib.load()
    stats, _, _ = head.head_forward_raw(lg, T, lab, size, need_grad=False)
    n_valid = float(stats[1])
    B, CK, h, w = lg.shape
    C = T.shape[1]
    dl = torch.empty_like(lg)
    dT = torch.empty_like(T)
    ws = torch.zeros(lib.simt_head_workspace_bytes(B, CK, C, h, w, *size), dtype=torch.uint8, device=dev)
    rc = lib.simt_head_bwd(lg.data_ptr(), B, CK, h, w, T.data_ptr(), C, lab.data_ptr(), 1, size[0], size[1], 255,
                           1.0 / n_valid, dl.data_ptr(), dT.data_ptr(), head.error_flag(dev).data_ptr(),
                           ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    torch.cuda.synchronize()
    _check(dl.cpu().numpy(), g["dlogits_f32"], "two-pass dlogits")
    _check(dT.cpu().numpy(), g["dT_f32"], "two-pass dT")


def test_grad_output_scaling_and_second_backward():
    import simt_b200
    g = load_golden("head_small_r")
    dev = torch.device("cuda")
    lg = torch.from_numpy(g["logits"]).to(dev).requires_grad_(True)
    T = torch.from_numpy(g["T"]).to(dev).requires_grad_(True)
    lab = torch.from_numpy(g["labels"]).to(dev)
    loss = simt_b200.simt_head(lg, T, lab, tuple(int(s) for s in g["size"]))
    (0.1 * loss).backward(retain_graph=True)        # lambda_seg = 0.1 on head 1 (trainV2_simt.py:423)
    _check(lg.grad.cpu().numpy(), 0.1 * g["dlogits_f32"], "scaled dlogits")
    _check(T.grad.cpu().numpy(), 0.1 * g["dT_f32"], "scaled dT")
    with pytest.raises(RuntimeError, match="second time"):
        loss.backward()


def test_edge_cases():
    import simt_b200
    dev = torch.device("cuda")
    lg = torch.randn(1, 19, 5, 9, device=dev, requires_grad=True)
    T = torch.softmax(torch.randn(19, 19, device=dev), 1)
    # all ignored -> NaN, like the reference's mean over an empty selection
    lab = torch.full((1, 32, 64), 255, dtype=torch.uint8, device=dev)
    assert torch.isnan(simt_b200.simt_head(lg, T, lab, (32, 64)))
    # negative int64 labels are ignored (utils/loss.py:29)
    lab64 = torch.randint(0, 19, (1, 32, 64), device=dev)
    lab_neg = lab64.clone()
    lab_neg[0, :16] = -1
    lab_ign = lab64.clone()
    lab_ign[0, :16] = 255
    assert float(simt_b200.simt_head(lg, T, lab_neg, (32, 64))) == float(simt_b200.simt_head(lg, T, lab_ign, (32, 64)))
    # a label in [C, 254] is a contract violation: loss poisoned with NaN + IndexError on check
    bad = lab64.clone()
    bad[0, 3, 3] = 77
    out = simt_b200.simt_head(lg, T, bad, (32, 64))
    assert torch.isnan(out)
    with pytest.raises(IndexError):
        simt_b200.check_errors()
    simt_b200.check_errors()      # flag cleared
    # huge dynamic range: the run-level softmax bound must fall back to the exact max
    big = torch.zeros(1, 19, 2, 2, device=dev)
    big[0, 0, 0, 0] = 3000.0
    big[0, 1, :, :] = -3000.0
    big[0, 2, 1, 1] = 2500.0
    labs = torch.randint(0, 19, (1, 16, 16), device=dev).to(torch.uint8)
    ref_in = big.cpu().double().requires_grad_(True)
    from oracle import simt_oracle as O
    ref = O.simt_head_loss(ref_in, T.cpu().double(), labs.cpu().long(), (16, 16))
    got = simt_b200.simt_head(big, T, labs, (16, 16))
    assert torch.isfinite(got)
    assert abs(float(got) - float(ref)) <= 1e-5 * abs(float(ref))


def test_deterministic_loss_and_dT():
    g = load_golden("head_cfg1_tile")
    a = run_gpu_head(g["logits"], g["T"], g["labels"], g["size"])
    b = run_gpu_head(g["logits"], g["T"], g["labels"], g["size"])
    assert a[0].tobytes() == b[0].tobytes()


def test_crossentropy2d_dropin_both_modes():
    import simt_b200
    g = load_golden("ce2d")
    dev = torch.device("cuda")
    y = torch.from_numpy(g["y"]).to(dev)
    for mode in (1, 0):
        x = torch.from_numpy(g["x"])
        xin = (x if mode else torch.softmax(x, 1)).to(dev).requires_grad_(True)
        crit = simt_b200.CrossEntropy2d(is_softmax=bool(mode)).cuda()
        loss = crit(xin, y)
        loss.backward()
        ref = g[f"loss_softmax{mode}"]
        assert abs(float(loss) - float(ref)) <= TOL * abs(float(ref)), mode
        _check(xin.grad.cpu().numpy(), g[f"grad_softmax{mode}"], f"grad mode {mode}")


def test_reference_training_lines_with_fused_head():
    """The reference's unfused lines on torch-CUDA vs the fused op, same device, same tensors."""
    import torch.nn as nn
    import simt_b200
    from oracle import simt_oracle as O
    dev = torch.device("cuda")
    logits, labels = O.synth_head_inputs(2, 23, 33, 65, 256, 512, seed=77, coherent=True, class_dist=class_dist())
    torch.manual_seed(3)
    ntm = simt_b200.sig_NTM(19, 4).to(dev)
    lg1 = logits.to(dev).requires_grad_(True)
    T1 = ntm()
    ref = O.simt_head_loss(lg1, T1, labels.to(dev).long(), (256, 512))
    ref.backward()
    g_ref, n_ref = lg1.grad.clone(), ntm.NTM.grad.clone()
    ntm.NTM.grad = None
    lg2 = logits.to(dev).requires_grad_(True)
    out = simt_b200.simt_head(lg2, ntm(), labels.to(dev), (256, 512))
    out.backward()
    assert abs(float(out) - float(ref)) <= TOL * abs(float(ref))
    _check(lg2.grad.cpu().numpy(), g_ref.cpu().numpy(), "dlogits vs torch-CUDA eager")
    _check(ntm.NTM.grad.cpu().numpy(), n_ref.cpu().numpy(), "dNTM vs torch-CUDA eager")


@pytest.mark.parametrize("B,K", [(64, 0), (16, 15)])
def test_full_size_invariants(B, K):
    """BASELINE config 5 / 3 sizes (too big for the CPU oracle in a test): size-independent properties.

    * softmax gradient: sum_k dz_k = (1/N)(sum_k p_k - sum_k p_k T_ky / q) = 0 at every pixel, and U^T is
      linear, so dLogits summed over channels vanishes at every low-res node;
    * <T, dT> = -(1/N) sum_pixels sum_k p_k T_ky / q = -1;
    * permuting the batch permutes dLogits and leaves loss and dT unchanged (to rounding);
    * doubling every image (batch concatenated with itself) leaves loss and dT unchanged and halves dLogits.
    """
    import simt_b200
    from oracle import simt_oracle as O
    dev = torch.device("cuda")
    CK = 19 + K
    logits, labels = O.synth_head_inputs(B, CK, 65, 129, 512, 1024, seed=99, coherent=True, block=(36, 52),
                                         class_dist=class_dist())
    torch.manual_seed(5)
    T = simt_b200.sig_NTM(19, K)().detach().to(dev)

    def run(lg_cpu, lab_cpu):
        lg = lg_cpu.to(dev).requires_grad_(True)
        Tt = T.clone().requires_grad_(True)
        loss = simt_b200.simt_head(lg, Tt, lab_cpu.to(dev), (512, 1024))
        loss.backward()
        return loss.detach(), lg.grad, Tt.grad

    loss, dl, dT = run(logits, labels)
    assert torch.isfinite(loss) and 0.0 < float(loss) < 20.0
    node_sum = dl.sum(dim=1)
    assert float(node_sum.abs().max()) <= 1e-5 * float(dl.abs().max())
    assert abs(float((T * dT).sum()) + 1.0) <= 1e-5
    perm = torch.randperm(B, generator=torch.Generator().manual_seed(1))
    loss_p, dl_p, dT_p = run(logits[perm], labels[perm])
    assert abs(float(loss_p) - float(loss)) <= 1e-6 * abs(float(loss))
    assert float((dT_p - dT).norm() / dT.norm()) <= 1e-5
    assert float((dl_p - dl[perm.to(dev)]).norm() / dl.norm()) <= 1e-5
    if B <= 16:
        loss_2, dl_2, dT_2 = run(torch.cat([logits, logits]), torch.cat([labels, labels]))
        assert abs(float(loss_2) - float(loss)) <= 1e-6 * abs(float(loss))
        assert float((dT_2 - dT).norm() / dT.norm()) <= 1e-5
        assert float((2 * dl_2[:B] - dl).norm() / dl.norm()) <= 1e-5
    simt_b200.check_errors()


def test_headrunner_graph_step_matches_eager_and_sees_new_inputs():
    """HeadRunner.graph_step (CUDA-graph replay of memset + kernel + finalize + scale) gives the eager result, and a
    replay reads the CURRENT contents of the captured buffers."""
    import simt_b200
    from oracle import simt_oracle as O
    dev = torch.device("cuda")
    B, CK, h, w, H, W = 2, 19, 17, 33, 128, 256
    sets = [O.synth_head_inputs(B, CK, h, w, H, W, seed=40 + i, coherent=True, ignore_frac=0.1) for i in range(2)]
    T = O.sig_ntm_forward(torch.randn(CK, 19, generator=torch.Generator().manual_seed(2)), class_dist(), 19, 0).to(dev)
    r = simt_b200.HeadRunner(B, CK, 19, h, w, H, W, device=dev)
    lg, lab = sets[0][0].to(dev), sets[0][1].to(torch.uint8).to(dev)
    for rep in range(3):                     # first call captures, the others replay
        loss, dl, dT = r.graph_step(lg, T, lab)
    ref = O.simt_head_fwd_bwd(sets[0][0], T.cpu(), sets[0][1], (H, W), torch.float64)
    assert abs(float(loss) - float(ref[0])) <= TOL * abs(float(ref[0]))
    assert rel_l2(dl.cpu().numpy(), ref[1].numpy()) <= TOL and rel_l2(dT.cpu().numpy(), ref[2].numpy()) <= TOL
    lg.copy_(sets[1][0]); lab.copy_(sets[1][1].to(torch.uint8))          # refill in place -> same graph
    loss, dl, dT = r.graph_step(lg, T, lab)
    assert len(r._graphs) == 1
    ref = O.simt_head_fwd_bwd(sets[1][0], T.cpu(), sets[1][1], (H, W), torch.float64)
    assert abs(float(loss) - float(ref[0])) <= TOL * abs(float(ref[0]))
    assert rel_l2(dl.cpu().numpy(), ref[1].numpy()) <= TOL and rel_l2(dT.cpu().numpy(), ref[2].numpy()) <= TOL
    simt_b200.check_errors(dev)


def test_host_prefetcher_double_buffering_delivers_every_batch():
    """HostPrefetcher: pinned host batches arrive intact and in order while the previous step is still computing."""
    import simt_b200
    from oracle import simt_oracle as O
    dev = torch.device("cuda")
    B, CK, h, w, H, W = 2, 19, 9, 17, 64, 128
    T = O.sig_ntm_forward(torch.randn(CK, 19, generator=torch.Generator().manual_seed(4)), class_dist(), 19, 0)
    pre = simt_b200.HostPrefetcher(B, CK, h, w, H, W, device=dev)
    host, packed = [], []
    for i in range(5):
        lg, lab = O.synth_head_inputs(B, CK, h, w, H, W, seed=70 + i, coherent=True, ignore_frac=0.1)
        host.append((lg, lab.to(torch.uint8)))
        buf = pre.host_buffer()                  # one pinned buffer per batch: logits bytes, then label bytes
        buf.logits.copy_(lg); buf.labels.copy_(lab.to(torch.uint8))
        packed.append(buf)
    runner = simt_b200.HeadRunner(B, CK, 19, h, w, H, W, device=dev)
    Td = T.to(dev)
    losses = []
    pre.submit(packed[0])
    for i in range(5):
        cur = pre.get()
        if i + 1 < 5:
            pre.submit(packed[i + 1])
        loss, _, _ = runner.step(cur[0], Td, cur[1])
        pre.release(cur)
        losses.append(loss.clone())
    torch.cuda.synchronize()
    for i in range(5):
        ref = O.simt_head_loss(host[i][0], T, host[i][1].long(), (H, W))
        assert abs(float(losses[i]) - float(ref)) <= TOL * abs(float(ref)), i


@pytest.mark.parametrize("int64_labels", [False, True])
@pytest.mark.parametrize("K", [0, 4])
def test_headrunner_step_device_side_scale(K, int64_labels):
    """HeadRunner.step (simt_head_step: label count pass + fused kernel applying grad_out / N itself, no scale pass):
    loss, dLogits, dT equal the oracle's, with and without an upstream gradient, for uint8 and int64 labels."""
    import simt_b200
    from oracle import simt_oracle as O
    dev = torch.device("cuda")
    B, CK, h, w, H, W = 3, 19 + K, 9, 17, 70, 133          # odd sizes: unaligned label rows, tail paths of the count pass
    lg, lab = O.synth_head_inputs(B, CK, h, w, H, W, seed=11 + K, coherent=True, ignore_frac=0.15)
    T = O.sig_ntm_forward(torch.randn(CK, 19, generator=torch.Generator().manual_seed(2)), class_dist(), 19, K)
    ref = O.simt_head_fwd_bwd(lg, T, lab, (H, W), torch.float64)
    r = simt_b200.HeadRunner(B, CK, 19, h, w, H, W, device=dev, label_dtype=torch.int64 if int64_labels else torch.uint8)
    labd = (lab.long() if int64_labels else lab.to(torch.uint8)).to(dev)
    for g in (None, 0.37):
        go = None if g is None else torch.tensor(g, device=dev)
        loss, dl, dT = r.step(lg.to(dev), T.to(dev), labd, grad_out=go)
        s = 1.0 if g is None else g
        assert abs(float(loss) - float(ref[0])) <= TOL * abs(float(ref[0]))
        assert rel_l2(dl.cpu().numpy(), s * ref[1].numpy()) <= TOL
        assert rel_l2(dT.cpu().numpy(), s * ref[2].numpy()) <= TOL
        assert float(r.stats[1]) == float(((lab >= 0) & (lab < 19)).sum())
    # nothing valid: mean over nothing -> NaN, like the reference
    labd.fill_(255)
    loss, dl, dT = r.step(lg.to(dev), T.to(dev), labd)
    assert np.isnan(float(loss))
    simt_b200.check_errors(dev)


@pytest.mark.parametrize("gap", [60.0, 200.0, 1000.0])
def test_plain_ce_huge_margin_is_finite_like_log_softmax(gap):
    """T = None with the label's logit far below the maximum: p_y underflows in fp32 but torch's cross entropy
    (log_softmax) stays finite -- loss = the gap, dlogits = p - onehot (utils/loss.py:35-36 -> F.cross_entropy)."""
    from oracle import simt_oracle as O
    logits, labels = O.synth_head_inputs(1, 19, 5, 9, 32, 64, seed=11, coherent=True, block=8, ignore_frac=0.1)
    logits = logits * 0.5
    logits[:, 3] += gap          # channel 3 dominates everywhere; most labels are other classes
    lg = logits.clone().requires_grad_(True)
    ref = O.plain_ce_loss(lg, labels.long(), (32, 64))
    ref.backward()
    assert np.isfinite(float(ref))
    loss, dl, _ = run_gpu_head(logits.numpy(), None, labels.numpy(), (32, 64))
    assert np.isfinite(float(loss)) and np.isfinite(dl).all()
    assert abs(float(loss) - float(ref)) <= TOL * abs(float(ref))
    _check(dl, lg.grad.numpy(), "dlogits")


@pytest.mark.parametrize("is_softmax", [True, False])
def test_cross_entropy_2d_per_class_weight(is_softmax):
    """CrossEntropy2d(...)(predict, target, weight=w) (utils/loss.py:14,36,39): weighted mean over valid pixels."""
    import simt_b200
    from oracle import simt_oracle as O
    g = torch.Generator().manual_seed(21)
    x = torch.randn(2, 19, 24, 40, generator=g)
    if not is_softmax:
        x = torch.softmax(x, dim=1)
    y = torch.randint(0, 19, (2, 24, 40), generator=g)
    y[torch.rand(2, 24, 40, generator=g) < 0.2] = 255
    w = torch.rand(19, generator=g) + 0.1
    xr = x.clone().requires_grad_(True)
    ref = O.cross_entropy_2d(xr, y, is_softmax=is_softmax, weight=w)
    ref.backward()
    dev = torch.device("cuda")
    xg = x.to(dev).requires_grad_(True)
    loss = simt_b200.CrossEntropy2d(is_softmax=is_softmax)(xg, y.to(dev), weight=w.to(dev))
    loss.backward()
    assert abs(float(loss.detach()) - float(ref.detach())) <= TOL * abs(float(ref.detach()))
    _check(xg.grad.cpu().numpy(), xr.grad.numpy(), "d predict (weighted)")
