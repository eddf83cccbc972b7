"""GPU parity of the integer histograms: bit-exact against the reference golden vectors and the
numpy oracle, including the reference's edge cases."""
import numpy as np
import pytest
import torch

from util import load_golden

pytestmark = pytest.mark.gpu


def test_golden_confusion_label_mapping_miou():
    import simt_b200
    g = load_golden("hist")
    mapping = g["mapping"]
    meter = simt_b200.ConfusionMeter(19, mapping=mapping)
    hist = np.zeros((19, 19), dtype=np.int64)
    for gt, pr in zip(g["gt"], g["pred"]):
        meter.update(gt, pr)                                   # fused LUT + confusion
        lab = simt_b200.label_mapping(gt, mapping)             # the unfused drop-in pair
        assert lab.dtype == np.int64 and lab.shape == gt.shape
        hist += simt_b200.fast_hist(lab.flatten(), pr.flatten(), 19)
    assert np.array_equal(hist, g["hist19"])
    assert np.array_equal(meter.value(), g["hist19"])
    assert meter.miou_percent() == float(g["miou"])
    iu = simt_b200.per_class_iu(hist)
    assert np.array_equal(np.nan_to_num(iu, nan=-1), np.nan_to_num(g["iu"], nan=-1))
    assert np.array_equal(simt_b200.fast_hist(g["gt"][0].flatten(), g["pred"][0].flatten(), 34, 19), g["rect34x19"])
    assert np.array_equal(simt_b200.fast_hist(g["pred"][0].flatten(), 19), g["class19"])


@pytest.mark.parametrize("mode", [0])
@pytest.mark.parametrize("coherent", [True, False])
@pytest.mark.parametrize("n", [0, 1, 15, 16, 17, 4099, 1 << 20, (1 << 21) + 5])
def test_confusion_vs_oracle_sizes(mode, coherent, n):
    import simt_b200
    from simt_b200 import _lib
    from oracle import simt_oracle as O
    side = max(1, int(np.ceil(np.sqrt(max(n, 1)))))
    gt, pr = O.synth_eval_pair(side, side, seed=n % 97 + 3, coherent=coherent, block=16)
    gt, pr = gt.reshape(-1)[:n], pr.reshape(-1)[:n]
    lut = simt_b200.build_lut(O.CITYSCAPES_LABEL2TRAIN)
    ref = O.fast_hist(O.label_mapping(gt, np.array(O.CITYSCAPES_LABEL2TRAIN)), pr, 19)
    lib = _lib.load()
    try:
        lib.simt_hist_set_tuning(mode, 0, 0)
        m = simt_b200.ConfusionMeter(19, mapping=O.CITYSCAPES_LABEL2TRAIN)
        m.update(gt, pr)
        m.update(gt, pr)                       # accumulates like `hist +=`
        got = m.value()
    finally:
        lib.simt_hist_set_tuning(0, 0, 0)
    assert np.array_equal(got, 2 * ref)
    assert lut.dtype == np.uint8


@pytest.mark.parametrize("unroll,warps", [(1, 4), (2, 8), (4, 16)])
def test_confusion_tunings_agree(unroll, warps):
    import simt_b200
    from simt_b200 import _lib
    from oracle import simt_oracle as O
    gt, pr = O.synth_eval_pair(512, 1024, seed=9, coherent=True)
    ref = O.fast_hist_rect(gt.reshape(-1), pr.reshape(-1), 34, 19)
    lib = _lib.load()
    try:
        lib.simt_hist_set_tuning(0, warps, unroll)
        got = simt_b200.fast_hist(gt.reshape(-1), pr.reshape(-1), 34, 19)
    finally:
        lib.simt_hist_set_tuning(0, 0, 0)
    assert np.array_equal(got, ref)


def test_worst_case_contention_and_byte_counter_overflow():
    """One bin gets every pixel (road is 41 % of real labels): private 8-bit counters must fold in time."""
    import simt_b200
    n = 5_000_011
    a = np.zeros(n, dtype=np.uint8)
    b = np.full(n, 3, dtype=np.uint8)
    a[::7919] = 5                              # break the uniform fast path now and then
    got = simt_b200.fast_hist(a, b, 19)
    ref = np.zeros((19, 19), dtype=np.int64)
    ref[5, 3] = len(a[::7919])
    ref[0, 3] = n - ref[5, 3]
    assert np.array_equal(got, ref)
    # alternating pattern defeats every fast path and hammers two bins
    a = (np.arange(n) & 1).astype(np.uint8)
    got = simt_b200.fast_hist(a, np.zeros(n, dtype=np.uint8), 19)
    assert got[0, 0] == (n + 1) // 2 and got[1, 0] == n // 2 and got.sum() == n


def test_int64_inputs_and_masks():
    import simt_b200
    from oracle import simt_oracle as O
    rng = np.random.default_rng(0)
    a = rng.integers(-3, 300, size=100_003).astype(np.int64)       # negatives and >= n are masked
    b = rng.integers(0, 19, size=100_003).astype(np.int64)
    assert np.array_equal(simt_b200.fast_hist(a, b, 19), O.fast_hist(a, b, 19))
    a8 = rng.integers(0, 256, size=100_003).astype(np.uint8)
    assert np.array_equal(simt_b200.fast_hist(a8, 19), O.class_hist(a8, 19))
    assert np.array_equal(simt_b200.fast_hist(a, 19), O.class_hist(a, 19))
    # misaligned views take the generic kernel
    assert np.array_equal(simt_b200.fast_hist(torch.from_numpy(a8).cuda()[3:], torch.from_numpy(b).cuda()[3:].to(torch.uint8), 19).cpu().numpy(),
                          O.fast_hist(a8[3:].astype(np.int64), b[3:], 19))


def test_reference_aliasing_quirk_and_out_of_table_error():
    import simt_b200
    from oracle import simt_oracle as O
    a = np.array([0, 1, 2] * 11, dtype=np.uint8)
    b = np.array([0, 20, 1] * 11, dtype=np.uint8)
    assert np.array_equal(simt_b200.fast_hist(a, b, 19), O.fast_hist(a, b, 19))     # b >= n aliases, like numpy
    with pytest.raises(IndexError):                                                   # numpy: ValueError on reshape
        simt_b200.fast_hist(np.array([18] * 40, dtype=np.uint8), np.array([30] * 40, dtype=np.uint8), 19)


def test_full_resolution_eval_image_bit_exact_miou():
    """BASELINE config 4 shape (2048x1024 per image) on a few images; mIoU string bit-exact."""
    import simt_b200
    from oracle import simt_oracle as O
    mapping = np.array(O.CITYSCAPES_LABEL2TRAIN)
    meter = simt_b200.ConfusionMeter(19, mapping=mapping)
    ref = np.zeros((19, 19), dtype=np.int64)
    for i in range(3):
        gt, pr = O.synth_eval_pair(1024, 2048, seed=40 + i, coherent=(i < 2))
        meter.update(gt, pr)
        ref += O.fast_hist(O.label_mapping(gt, mapping).flatten(), pr.flatten(), 19)
    assert np.array_equal(meter.value(), ref)
    assert meter.miou_percent() == O.miou_percent(ref)
    assert int(meter.value().sum()) == int((O.label_mapping(gt, mapping) < 19).sum() * 0 + ref.sum())


def test_config4_scale_checksums():
    """BASELINE config 4 scale: 200 images 2048x1024 in ONE launch (0.4 G pixels): totals must equal the
    number of gt pixels that map to a train id, row sums the per-class gt counts, and splitting the batch
    in two launches gives the same table bit for bit."""
    import simt_b200
    from oracle import simt_oracle as O
    dev = torch.device("cuda")
    mapping = np.array(O.CITYSCAPES_LABEL2TRAIN)
    base = [O.synth_eval_pair(1024, 2048, seed=70 + i, coherent=(i % 2 == 0), block=(96, 160)) for i in range(4)]
    gt = torch.from_numpy(np.stack([b[0] for b in base])).to(dev).repeat(50, 1, 1)
    pr = torch.from_numpy(np.stack([b[1] for b in base])).to(dev).repeat(50, 1, 1)
    m1 = simt_b200.ConfusionMeter(19, mapping=mapping)
    m1.update(gt, pr)
    h1 = m1.value()
    lut = simt_b200.build_lut(mapping).astype(np.int64)
    gt_train = lut[np.stack([b[0] for b in base])]
    rows = np.bincount(gt_train[gt_train < 19], minlength=19) * 50
    assert np.array_equal(h1.sum(1), rows)
    assert int(h1.sum()) == int((gt_train < 19).sum()) * 50
    ref = np.zeros((19, 19), dtype=np.int64)
    for g, p in base:
        ref += O.fast_hist(O.label_mapping(g, mapping).flatten(), p.flatten(), 19)
    assert np.array_equal(h1, 50 * ref)
    m2 = simt_b200.ConfusionMeter(19, mapping=mapping)
    m2.update(gt[:77], pr[:77]); m2.update(gt[77:], pr[77:])
    assert np.array_equal(m2.value(), h1)


@pytest.mark.parametrize("two_scale", [True, False])
def test_fused_eval_argmax_and_miou(two_scale):
    """evaluate_cityscapes.py:127-148 on synthetic logits: the fused prediction map equals the oracle's except at
    arg-max near-ties (float op order of the upsample), and the resulting confusion matrix differs by at most
    those pixels."""
    import simt_b200
    from oracle import simt_oracle as O
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(5)
    B, H, W = 2, 256, 512
    a = 3.0 * torch.randn(B, 34, 33, 65, generator=g)              # open-set head: only the first 19 channels count
    b = 3.0 * torch.randn(B, 34, 41, 81, generator=g) if two_scale else None
    ref = O.eval_two_scale_argmax(a, b, (H, W), 19)
    got = simt_b200.eval_argmax(a.to(dev), None if b is None else b.to(dev), (H, W), 19).cpu().numpy()
    diff = got != ref
    z = O.upsample_bilinear_ac(a[:, :19].double(), (H, W))
    if b is not None:
        z = z + O.upsample_bilinear_ac(b[:, :19].double(), (H, W))
    top2 = z.topk(2, dim=1).values
    near_tie = ((top2[:, 0] - top2[:, 1]) < 1e-4).numpy()
    assert not (diff & ~near_tie).any()
    assert diff.mean() < 1e-3
    gt, _ = O.synth_eval_pair(B * H, W, seed=1, block=(24, 40))
    gt = gt.reshape(B, H, W)
    m = simt_b200.ConfusionMeter(19, mapping=O.CITYSCAPES_LABEL2TRAIN)
    m.update(gt, torch.from_numpy(got).to(dev))
    lab = O.label_mapping(gt, np.array(O.CITYSCAPES_LABEL2TRAIN))
    ref_hist = O.fast_hist(lab.flatten(), ref.flatten().astype(np.int64), 19)
    assert np.abs(m.value() - ref_hist).sum() <= 2 * int(diff.sum())


@pytest.mark.parametrize("sa,sb,size", [
    ((33, 65), (21, 41), (256, 512)),     # evaluate_cityscapes.py's ratio: the second scale is coarser (three-column path)
    ((9, 12), (14, 23), (70, 95)),        # odd sizes, second scale finer (per-pixel gather path)
    ((17, 33), None, (8, 16)),            # down-sampling, one scale
    ((5, 4), (3, 2), (5, 4)),             # identity size for the first scale
    ((2, 3), (1, 1), (40, 100)),          # 50-pixel runs, a 1x1 second scale
])
def test_eval_argmax_shapes_vs_oracle(sa, sb, size):
    import simt_b200
    from oracle import simt_oracle as O
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(sa[0] * 10 + sa[1])
    a = 3.0 * torch.randn(2, 23, *sa, generator=g)
    b = None if sb is None else 3.0 * torch.randn(2, 23, *sb, generator=g)
    ref = O.eval_two_scale_argmax(a, b, size, 19)
    got = simt_b200.eval_argmax(a.to(dev), None if b is None else b.to(dev), size, 19).cpu().numpy()
    assert got.shape == ref.shape and got.dtype == np.uint8
    z = O.upsample_bilinear_ac(a[:, :19].double(), size)
    if b is not None:
        z = z + O.upsample_bilinear_ac(b[:, :19].double(), size)
    top2 = z.topk(2, dim=1).values
    near_tie = ((top2[:, 0] - top2[:, 1]) < 1e-4).numpy()
    assert not ((got != ref) & ~near_tie).any()


@pytest.mark.parametrize("dtype", [np.int16, np.int32, np.int64])
def test_label_mapping_integer_images(dtype):
    """label_mapping takes any integer array (tools/compute_iou.py:18-22), not only the uint8 PNG images."""
    import simt_b200
    from oracle import simt_oracle as O
    rng = np.random.default_rng(8)
    m = np.array(O.CITYSCAPES_LABEL2TRAIN)
    x = rng.integers(-1, 40, size=(61, 47)).astype(dtype)
    got = simt_b200.label_mapping(x, m)
    ref = O.label_mapping(x, m)
    assert got.dtype == np.int64 and np.array_equal(got, ref)
