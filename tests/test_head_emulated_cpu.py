"""The fused head kernel ITSELF, executed on the CPU against the reference's golden vectors and the oracle (no GPU).

csrc/head_kernel.cuh (all modes of the fused kernel), csrc/step_kernels.cuh (label-count prologue, finalize) and
csrc/head_plan.cuh (launch plan) -- the files the product build compiles with nvcc -- are compiled with g++ against the
CUDA-on-CPU shim of tests/cpu_simt (warps = 32 fibers, shuffles / votes / __syncwarp = barriers among them, random
scheduling of lanes and warps, dynamic shared memory pre-filled with a NaN pattern) and driven through the launch
sequences of csrc/head.cu (tests/cpu_simt/head_emul.cpp).  The inline-PTX helpers have host alternates behind
SIMT_CPU_EMULATION: fma.rn.f32x2 -> two fmaf, ex2 / lg2 / rcp.approx -> libm, red.global.add -> add, %smid -> block index;
everything else -- units, cells, label decoding, soft-max bounds, the transposed lerp, dT tiles, finalize -- is the
product code.  Same bar as on the GPU: loss, dLogits, dT within 1e-5 (norm-wise and max-scaled) of the fp32 AND fp64
reference (tests/golden, recorded from the unmodified reference)."""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest
import torch

from util import HEAD_CASES, class_dist, load_golden, rel_l2, rel_max

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "build", "cpu_simt")
TOL = 1e-5
MODE_FWD, MODE_FWDBWD, MODE_BWD, MODE_PLACE, MODE_STEP = 0, 1, 2, 3, 4

pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not found")


def _build():
    os.makedirs(OUT, exist_ok=True)
    so = os.path.join(OUT, "libhead_emul.so")
    csrc = os.path.join(ROOT, "simt_b200", "csrc")
    deps = [os.path.join(ROOT, "tests", "cpu_simt", f) for f in ("head_emul.cpp", "cuda_shim.h")] + [
        os.path.join(csrc, f) for f in ("common.cuh", "xchg.cuh", "step_xchg.cuh", "step_kernels.cuh", "head_kernel.cuh",
                                        "head_plan.cuh", "stepx_acquire.inc", "stepx_cta0.inc")]
    if os.path.exists(so) and os.path.getmtime(so) >= max(os.path.getmtime(d) for d in deps):
        return so
    cmd = ["g++", "-O0", "-g", "-std=c++17", "-pthread", "-shared", "-fPIC", "-ffp-contract=off", "-Wno-unknown-pragmas",
           "-I", os.path.join(ROOT, "tests", "cpu_simt"), "-I", csrc, os.path.join(ROOT, "tests", "cpu_simt", "head_emul.cpp"),
           "-o", so]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-4000:]
    return so


@pytest.fixture(scope="module")
def emu():
    lib = ctypes.CDLL(_build())
    vp, i, f = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
    lib.emul_head.restype = i
    lib.emul_head.argtypes = [i, vp, i, i, i, i, vp, i, vp, i, i, i, i, f, f, f, vp, vp, vp, vp, vp, i, i, i, i, ctypes.c_ulonglong]
    return lib


def run(lib, mode, logits, T, labels, size, i64=False, scale=1.0, seed=1, sm=3, cps=2, ur=0, rs=0, ignore=255,
        thres=-1.0, lam=0.0, C=None):
    logits = np.ascontiguousarray(logits, dtype=np.float32)
    B, CK, h, w = logits.shape
    if C is None:
        C = CK if T is None else T.shape[1]
    Tc = None if T is None else np.ascontiguousarray(T, dtype=np.float32)
    lab = None if labels is None else np.ascontiguousarray(np.asarray(labels).astype(np.int64 if i64 else np.uint8))
    H, W = int(size[0]), int(size[1])
    dl = np.full_like(logits, 7.0)                      # garbage: the call must zero / overwrite it
    stats = np.full(2 + CK * C, -5.0, np.float64)
    loss = np.full(1, -5.0, np.float32)
    dT = np.full((CK, C), -5.0, np.float32)
    err = np.zeros(1, np.int32)
    rc = lib.emul_head(mode, logits.ctypes.data, B, CK, h, w, None if Tc is None else Tc.ctypes.data, C,
                       None if lab is None else lab.ctypes.data, 8 if i64 else 1, H, W, ignore, scale, thres, lam,
                       dl.ctypes.data, stats.ctypes.data, loss.ctypes.data, dT.ctypes.data, err.ctypes.data, sm, cps, ur, rs, seed)
    assert rc == 0, f"emul_head returned {rc}"
    return float(loss[0]), dl, dT, stats, int(err[0])


def _check(got, ref, what):
    assert rel_l2(got, ref) <= TOL, f"{what}: rel l2 {rel_l2(got, ref):.3e}"
    assert rel_max(got, ref) <= TOL, f"{what}: rel max {rel_max(got, ref):.3e}"


def _scaled(dl_raw, stats, shape_T, g=1.0):
    """simt_head_scale: dlogits = raw * g / N, dT = raw dT * g / N"""
    n = stats[1]
    return dl_raw * (g / n), (stats[2:] * (g / n)).reshape(shape_T)


@pytest.mark.parametrize("name", HEAD_CASES)
@pytest.mark.parametrize("int64_labels", [False, True])
def test_fused_kernel_vs_reference_golden(emu, name, int64_labels):
    g = load_golden(name)
    # simt_head_fwdbwd + simt_head_scale (the autograd entry)
    loss, dl_raw, _, stats, err = run(emu, MODE_FWDBWD, g["logits"], g["T"], g["labels"], g["size"], int64_labels)
    dl, dT = _scaled(dl_raw, stats, g["T"].shape)
    assert err == 0
    assert abs(loss - float(g["loss_f32"])) <= TOL * abs(float(g["loss_f32"]))
    for ref in ("f32", "f64"):
        _check(dl, g["dlogits_" + ref], "dlogits vs " + ref)
        _check(dT, g["dT_" + ref], "dT vs " + ref)
    # simt_head_step (HeadRunner.step): count pass, kernel applying grad_out / N itself, finalize
    loss2, dl2, dT2, _, err = run(emu, MODE_STEP, g["logits"], g["T"], g["labels"], g["size"], int64_labels, scale=0.75, seed=2)
    assert err == 0 and abs(loss2 - float(g["loss_f32"])) <= TOL * abs(float(g["loss_f32"]))
    _check(dl2, 0.75 * g["dlogits_f64"], "step dlogits")
    _check(dT2, 0.75 * g["dT_f64"], "step dT")


@pytest.mark.parametrize("name", ["head_small_r", "head_openset4", "head_openset15", "head_odd"])
def test_forward_only_and_backward_with_host_scale(emu, name):
    g = load_golden(name)
    loss, _, _, stats, _ = run(emu, MODE_FWD, g["logits"], g["T"], g["labels"], g["size"])
    assert abs(loss - float(g["loss_f32"])) <= TOL * abs(float(g["loss_f32"]))
    _, dl, dT, _, _ = run(emu, MODE_BWD, g["logits"], g["T"], g["labels"], g["size"], scale=float(0.5 / stats[1]), seed=3)
    _check(dl, 0.5 * g["dlogits_f64"], "bwd dlogits")
    _check(dT, 0.5 * g["dT_f64"], "bwd dT")


@pytest.mark.parametrize("ur,rs", [(1, 0), (1, 1), (1, 2), (1, 4), (2, 0), (3, 0), (8, 0), (32, 0)])
def test_unit_shapes_are_the_same_function(emu, ur, rs):
    """every unit height (cell-rows per warp unit) and row split gives the same result"""
    g = load_golden("head_cfg1_tile")
    loss, dl_raw, _, stats, _ = run(emu, MODE_FWDBWD, g["logits"], g["T"], g["labels"], g["size"], ur=ur, rs=rs, seed=ur * 10 + rs)
    dl, dT = _scaled(dl_raw, stats, g["T"].shape)
    assert abs(loss - float(g["loss_f32"])) <= TOL * abs(float(g["loss_f32"]))
    _check(dl, g["dlogits_f32"], "dlogits")
    _check(dT, g["dT_f32"], "dT")


@pytest.mark.parametrize("sm,cps,seed", [(1, 1, 1), (2, 3, 2), (5, 2, 3), (7, 1, 4)])
def test_grid_shapes_and_schedules(emu, sm, cps, seed):
    """any number of SMs / resident CTAs and any interleaving of lanes and warps (dynamic unit claims, per-SM dT tiles)"""
    g = load_golden("head_openset4")
    loss, dl, dT, _, _ = run(emu, MODE_STEP, g["logits"], g["T"], g["labels"], g["size"], sm=sm, cps=cps, seed=seed)
    assert abs(loss - float(g["loss_f32"])) <= TOL * abs(float(g["loss_f32"]))
    _check(dl, g["dlogits_f64"], "dlogits")
    _check(dT, g["dT_f64"], "dT")


def test_plain_ce_T_none_and_huge_margins(emu):
    """T = NULL (the IDENT instantiations): F.cross_entropy on the upsampled logits, finite where p_y underflows"""
    from oracle import simt_oracle as O
    logits, labels = O.synth_head_inputs(2, 19, 9, 17, 64, 128, seed=5, coherent=True, block=8)
    for gap in (0.0, 60.0, 200.0, 1000.0):
        x = logits * (0.5 if gap else 1.0)
        if gap:
            x = x.clone()
            x[:, 3] += gap
        lg = x.clone().requires_grad_(True)
        ref = O.plain_ce_loss(lg, labels.long(), (64, 128))
        ref.backward()
        loss, dl, _, _, err = run(emu, MODE_STEP, x.numpy(), None, labels.numpy(), (64, 128), seed=int(gap) + 1)
        assert err == 0 and np.isfinite(loss) and np.isfinite(dl).all(), gap
        assert abs(loss - float(ref.detach())) <= TOL * abs(float(ref.detach())), gap
        _check(dl, lg.grad.numpy(), f"dlogits, gap {gap}")


def test_plain_ce_rows_of_mixed_range_in_one_warp(emu):
    """Regression (found by this emulation): with T = NULL, lanes whose row failed the first range test skipped the
    shuffle of the second one (`range_safe && group_max(...)`: a full-mask collective behind a short-circuit), so a
    warp that held rows of both kinds ran divergent collectives -- undefined behaviour on the GPU, a livelock here.
    Logits of mixed magnitude on a small up-sampling factor give every warp both kinds of rows."""
    from oracle import simt_oracle as O
    rng = np.random.default_rng(0)
    for h, w, H, W, i64 in ((8, 5, 12, 15, True), (6, 9, 17, 30, False), (5, 5, 5, 5, False)):
        x = torch.from_numpy((30.0 * rng.standard_normal((1, 19, h, w))).astype(np.float32))
        x[:, :, : h // 2] *= 0.05                                   # some rows tame, some wild
        labels = torch.from_numpy(rng.integers(0, 19, size=(1, H, W)))
        lg = x.clone().requires_grad_(True)
        ref = O.plain_ce_loss(lg, labels.long(), (H, W))
        ref.backward()
        loss, dl_raw, _, stats, err = run(emu, MODE_FWDBWD, x.numpy(), None, labels.numpy(), (H, W), i64=i64, seed=h)
        assert err == 0 and np.isfinite(loss)
        assert abs(loss - float(ref.detach())) <= TOL * abs(float(ref.detach()))
        assert rel_l2(dl_raw / stats[1], lg.grad.numpy()) <= TOL


def test_edge_cases(emu):
    from oracle import simt_oracle as O
    logits, labels = O.synth_head_inputs(1, 19, 5, 9, 32, 64, seed=9, coherent=True, block=8)
    T = O.sig_ntm_forward(torch.randn(19, 19, generator=torch.Generator().manual_seed(4)), class_dist(), 19, 0).numpy()
    # every pixel ignored: mean over nothing = NaN, like the reference
    loss, _, _, stats, err = run(emu, MODE_FWDBWD, logits.numpy(), T, np.full((1, 32, 64), 255), (32, 64))
    assert np.isnan(loss) and stats[1] == 0 and err == 0
    # a label in [C, 254]: error bit + NaN loss (the reference raises)
    bad = labels.numpy().copy()
    bad[0, 3, 5] = 77
    loss, _, _, _, err = run(emu, MODE_FWDBWD, logits.numpy(), T, bad, (32, 64))
    assert (err & 1) and np.isnan(loss)
    # negative int64 labels are ignored (utils/loss.py:29), equal to marking them 255
    neg = labels.numpy().astype(np.int64)
    mask = neg == 255
    neg[mask] = -3
    a = run(emu, MODE_FWDBWD, logits.numpy(), T, neg, (32, 64), i64=True)
    b = run(emu, MODE_FWDBWD, logits.numpy(), T, labels.numpy(), (32, 64))
    assert a[3][1] == b[3][1] and abs(a[0] - b[0]) <= 1e-6 * abs(b[0])
    # huge dynamic range inside a cell: rows fall back to the exact per-pixel maximum
    big = logits.numpy() * 300.0
    lo, dlo, dTo = O.simt_head_fwd_bwd(torch.from_numpy(big), torch.from_numpy(T), labels, (32, 64), torch.float64)
    _, dl32, dT32 = O.simt_head_fwd_bwd(torch.from_numpy(big), torch.from_numpy(T), labels, (32, 64), torch.float32)
    loss, dl, dT, _, _ = run(emu, MODE_STEP, big, T, labels.numpy(), (32, 64))
    assert abs(loss - float(lo)) <= TOL * abs(float(lo))
    # at |logit| ~ 1000 the fp32 lerp itself is only good to ~1e-4 in the exponent: the bar for the gradients is the
    # fp32 reference's own distance from its fp64 self
    assert np.isfinite(dl).all() and np.isfinite(dT).all()
    assert rel_l2(dl, dlo.numpy()) <= max(TOL, 3 * rel_l2(dl32.numpy(), dlo.numpy()))
    assert rel_l2(dT, dTo.numpy()) <= max(TOL, 3 * rel_l2(dT32.numpy(), dTo.numpy()))


@pytest.mark.parametrize("name", ["place_K4", "place_K15", "place_K4_nothres"])
def test_placeholder_mode_vs_reference_golden(emu, name):
    """MODE_PLACE (Placeholder_loss, tools/trainV2_simt.py:202-230) against the reference's own function"""
    g = load_golden(name)
    K = int(g["K"])
    loss, dl_raw, _, stats, _ = run(emu, MODE_PLACE, g["logits"], None, None, g["size"], thres=float(g["thres"]),
                                    lam=float(g["lambda_place"]), C=19)
    dl = dl_raw / stats[1]
    assert g["logits"].shape[1] == 19 + K
    assert abs(loss - float(g["loss_f64"])) <= TOL * abs(float(g["loss_f64"]))
    assert rel_l2(dl, g["dlogits_f64"]) <= TOL
    assert rel_max(dl, g["dlogits_f64"]) <= 5 * TOL


def test_random_shapes_vs_oracle(emu):
    """a short randomised sweep like tests/test_zfuzz_head_gpu.py: shapes, channel counts, label patterns, dtypes"""
    from oracle import simt_oracle as O
    rng = np.random.default_rng(3)
    done = 0
    while done < 12:
        K = int(rng.choice([0, 4, 15, 1, 9]))
        C, CK = 19, 19 + K
        B = int(rng.integers(1, 3))
        h, w = int(rng.integers(1, 12)), int(rng.integers(1, 14))
        mode = rng.choice(["up", "up", "same", "down"])
        if mode == "up":
            H, W = h * int(rng.integers(1, 7)) + int(rng.integers(0, 5)), w * int(rng.integers(1, 7)) + int(rng.integers(0, 5))
        elif mode == "same":
            H, W = h, w
        else:
            H, W = max(1, h // 2), max(1, w // 2 + 1)
        seed = int(rng.integers(0, 1 << 30))
        lg, lab = O.synth_head_inputs(B, CK, h, w, H, W, seed=seed, coherent=bool(rng.integers(0, 2)),
                                      ignore_frac=float(rng.choice([0.0, 0.1, 0.5])),
                                      block=(int(rng.integers(1, 12)), int(rng.integers(1, 12))), logit_scale=float(rng.choice([0.5, 3.0, 12.0])))
        if not bool(((lab >= 0) & (lab < C)).any()):
            continue
        T = O.sig_ntm_forward(torch.randn(CK, C, generator=torch.Generator().manual_seed(seed)), np.full(19, 1 / 19), C, K)
        lo, dlo, dTo = O.simt_head_fwd_bwd(lg, T, lab, (H, W), torch.float64)
        i64 = bool(rng.integers(0, 2))
        loss, dl, dT, _, err = run(emu, MODE_STEP, lg.numpy(), T.numpy(), lab.numpy(), (H, W), i64=i64, seed=seed & 0xffff,
                                   sm=int(rng.integers(1, 5)), cps=int(rng.integers(1, 4)))
        tag = f"B={B} K={K} {h}x{w}->{H}x{W} int64={i64} seed={seed}"
        assert err == 0, tag
        assert abs(loss - float(lo)) <= TOL * abs(float(lo)), tag
        assert rel_l2(dl, dlo.numpy()) <= TOL and rel_l2(dT, dTo.numpy()) <= TOL, tag
        done += 1


def test_the_gpu_fuzz_sweep_cases_on_the_emulator(emu):
    """tests/test_zfuzz_head_gpu.py draws 120 random cases for the GPU; the same generator, the same cases and the same
    1e-5 bar here, with the entry points replaced by the emulated kernel (simt_head -> fwdbwd + scale, HeadRunner.step ->
    step, Placeholder_loss -> MODE_PLACE + scale)."""
    import types

    src = open(os.path.join(ROOT, "tests", "test_zfuzz_head_gpu.py")).read().replace('torch.device("cuda")', 'torch.device("cpu")')

    class Leaf:
        """stands in for a CUDA tensor that requires grad: .grad is filled by the fake loss' backward()"""
        def __init__(self, t):
            self.t, self.grad = t, None

    class FakeLoss:
        def __init__(self, value, grads):
            self.value, self.grads = value, grads
        def backward(self):
            for leaf, g in self.grads:
                if leaf is not None:
                    leaf.grad = torch.from_numpy(np.ascontiguousarray(g))
        def detach(self):
            return self
        def __float__(self):
            return float(self.value)

    def as_leaf(x):
        return x if isinstance(x, Leaf) else Leaf(x)

    def simt_head(x, T, lab, size):
        x, Tl = as_leaf(x), (None if T is None else as_leaf(T))
        Tn = None if Tl is None else Tl.t.detach().numpy()
        loss, dl_raw, _, stats, err = run(emu, MODE_FWDBWD, x.t.detach().numpy(), Tn, lab.numpy(), size, seed=int(stats_seed[0]))
        stats_seed[0] += 1
        assert err == 0
        dl = dl_raw / stats[1]
        dT = None if Tn is None else (stats[2:] / stats[1]).reshape(Tn.shape).astype(np.float32)
        return FakeLoss(loss, [(x, dl), (Tl, dT)])

    class HeadRunner:
        def __init__(self, B, CK, C, h, w, H, W, device=None, label_dtype=torch.uint8):
            self.size, self.i64 = (H, W), label_dtype == torch.int64
        def step(self, lg, T, lab):
            loss, dl, dT, _, err = run(emu, MODE_STEP, lg.numpy(), None if T is None else T.numpy(), lab.numpy(), self.size,
                                       i64=self.i64, seed=int(stats_seed[0]))
            stats_seed[0] += 1
            assert err == 0
            return loss, torch.from_numpy(dl), torch.from_numpy(dT)

    def Placeholder_loss(x, C, K, thres, out_size=None, lambda_place=0.1):
        x = as_leaf(x)
        loss, dl_raw, _, stats, _ = run(emu, MODE_PLACE, x.t.detach().numpy(), None, None, out_size,
                                        thres=-1.0 if thres is None else float(thres), lam=float(lambda_place), C=C)
        return FakeLoss(loss, [(x, dl_raw / stats[1])])

    stats_seed = [1]
    fake = types.SimpleNamespace(simt_head=simt_head, HeadRunner=HeadRunner, Placeholder_loss=Placeholder_loss,
                                 check_errors=lambda dev=None: None)
    # tensors "moved to the device" that ask for gradients become leaves of the fake autograd
    orig_requires_grad_ = torch.Tensor.requires_grad_
    g = {"__name__": "zfuzz_on_emulator", "__file__": os.path.join(ROOT, "tests", "test_zfuzz_head_gpu.py")}
    try:
        torch.Tensor.requires_grad_ = lambda self, flag=True: Leaf(self) if self.dtype == torch.float32 else orig_requires_grad_(self, flag)
        import sys
        sys.modules["simt_b200_fake_for_fuzz"] = fake
        exec(compile(src.replace("import simt_b200  # noqa: E402", "import simt_b200_fake_for_fuzz as simt_b200"), "zfuzz_on_emulator", "exec"), g)
        worst, fails = g["run_fuzz"](120, 0)
    finally:
        torch.Tensor.requires_grad_ = orig_requires_grad_
        sys.modules.pop("simt_b200_fake_for_fuzz", None)
    assert not fails, "\n".join(fails)
    assert 0.0 < max(worst.values()) <= TOL and all(v > 0.0 for v in worst.values()), worst


@pytest.mark.parametrize("world,pipelined,K", [(2, 0, 0), (2, 1, 0), (4, 0, 4), (4, 1, 4), (8, 1, 0)])
def test_sharded_step_end_to_end_equals_the_global_batch(emu, world, pipelined, K):
    """simt_head_step_sharded on `world` emulated ranks (real prologue, REAL fused kernel in its MODE_STEPX
    instantiation, real finalize / finish; peer stores delayed and reordered) against the reference semantics: the
    single process sees the whole batch (tools/trainV2_simt.py:408-409,428) -- loss and dT of every rank equal the
    global-batch answer, dLogits of a rank equal its slice of it, and the all-reduced stats are bitwise identical on
    all ranks."""
    from oracle import simt_oracle as O
    vp, i, f = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
    emu.emul_head_sharded.restype = i
    emu.emul_head_sharded.argtypes = [i, i, i, vp, i, i, i, i, vp, i, vp, i, i, i, f, vp, vp, vp, vp, vp, i, i, ctypes.c_ulonglong]
    C, CK, h, w, H, W, Bper, g = 19, 19 + K, 5, 9, 32, 64, 1, 0.75
    logits, labels = O.synth_head_inputs(world * Bper, CK, h, w, H, W, seed=300 + world, coherent=True, block=(6, 10), ignore_frac=0.2)
    T = O.sig_ntm_forward(torch.randn(CK, C, generator=torch.Generator().manual_seed(world)), class_dist(), C, K)
    lo, dlo, dTo = O.simt_head_fwd_bwd(logits, T, labels, (H, W), torch.float64)
    lg = np.ascontiguousarray(logits.numpy(), dtype=np.float32)
    lab = np.ascontiguousarray(labels.numpy().astype(np.uint8))
    Tn = np.ascontiguousarray(T.numpy(), dtype=np.float32)
    dl = np.full_like(lg, 7.0)
    loss = np.zeros(world, np.float32)
    dT = np.zeros((world, CK, C), np.float32)
    stats = np.zeros((world, 2 + CK * C), np.float64)
    err = np.zeros(world, np.int32)
    rc = emu.emul_head_sharded(world, 4, pipelined, lg.ctypes.data, Bper, CK, h, w, Tn.ctypes.data, C, lab.ctypes.data, H, W, 255, g,
                               dl.ctypes.data, loss.ctypes.data, dT.ctypes.data, stats.ctypes.data, err.ctypes.data, 2, 2, 11 + world)
    assert rc == 0 and not err.any()
    for r in range(world):
        assert abs(float(loss[r]) - float(lo)) <= TOL * abs(float(lo)), r
        _check(dT[r], g * dTo.numpy(), f"dT of rank {r}")
        _check(dl[r * Bper:(r + 1) * Bper], g * dlo.numpy()[r * Bper:(r + 1) * Bper], f"dlogits of rank {r}")
        assert stats[r].tobytes() == stats[0].tobytes(), r
        assert loss[r].tobytes() == loss[0].tobytes() and dT[r].tobytes() == dT[0].tobytes(), r


def test_ignore_labels_outside_the_byte_range(emu):
    """Regression (found with this emulation): with an ignore label outside 0..255 (e.g. -1), int64 labels that are
    negative or equal to it were converted to the byte code 0xFF, which the kernel then did NOT treat as 'ignored' ->
    a false LABEL_RANGE error and a NaN loss.  The reference ignores `target < 0` and `target == ignore_label`
    whatever the value (utils/loss.py:29-30)."""
    from oracle import simt_oracle as O
    logits, labels = O.synth_head_inputs(2, 19, 5, 9, 32, 64, seed=9, coherent=True, block=8, ignore_frac=0.0)
    T = O.sig_ntm_forward(torch.randn(19, 19, generator=torch.Generator().manual_seed(4)), class_dist(), 19, 0)
    for ign in (-1, 300, -100):
        lab = labels.clone().long()
        lab[0, :5] = ign
        lab[1, 7, 3:40] = -7                      # negative labels are ignored whatever the ignore label is
        lo, dlo, dTo = O.simt_head_fwd_bwd(logits, T, lab, (32, 64), torch.float64, ignore_label=ign)
        loss, dl, dT, _, err = run(emu, MODE_STEP, logits.numpy(), T.numpy(), lab.numpy(), (32, 64), i64=True, ignore=ign, seed=3)
        assert err == 0, ign
        assert abs(loss - float(lo)) <= TOL * abs(float(lo)), ign
        _check(dl, dlo.numpy(), f"dlogits, ignore {ign}")
        _check(dT, dTo.numpy(), f"dT, ignore {ign}")
    # uint8 labels: an ignore label outside the byte range matches nothing, so a raw 255 is a contract violation
    lab8 = labels.clone()
    lab8[0, 0, 0] = 255
    loss, _, _, _, err = run(emu, MODE_STEP, logits.numpy(), T.numpy(), lab8.numpy(), (32, 64), ignore=300, seed=4)
    assert (err & 1) and np.isnan(loss)


@pytest.mark.parametrize("h,w,H,W,K,i64,coherent", [
    (2, 3, 50, 90, 0, False, False), (2, 3, 50, 90, 4, True, False), (2, 2, 33, 129, 15, False, False),
    (1, 2, 7, 300, 0, True, False), (2, 5, 3, 100, 0, False, False)])
def test_runs_longer_than_the_prefetched_labels(emu, h, w, H, W, K, i64, coherent):
    """large up-sampling factors: a cell's pixel run exceeds the 8 labels that travel in one word, so labels are fetched
    one by one (both walking directions of the boustrophedon rows)"""
    from oracle import simt_oracle as O
    CK = 19 + K
    lg, lab = O.synth_head_inputs(1, CK, h, w, H, W, seed=h * 100 + W, coherent=coherent, ignore_frac=0.1, block=(3, 5))
    T = O.sig_ntm_forward(torch.randn(CK, 19, generator=torch.Generator().manual_seed(4)), class_dist(), 19, K)
    lo, dlo, dTo = O.simt_head_fwd_bwd(lg, T, lab, (H, W), torch.float64)
    loss, dl, dT, _, err = run(emu, MODE_STEP, lg.numpy(), T.numpy(), lab.numpy(), (H, W), i64=i64, seed=3)
    assert err == 0 and abs(loss - float(lo)) <= TOL * abs(float(lo))
    _check(dl, dlo.numpy(), "dlogits")
    _check(dT, dTo.numpy(), "dT")


def test_upstream_gradient_values_and_an_all_ignored_image(emu):
    from oracle import simt_oracle as O
    lg, lab = O.synth_head_inputs(3, 19, 5, 9, 32, 64, seed=21, coherent=True, block=8, ignore_frac=0.1)
    lab = lab.clone()
    lab[1] = 255                                   # one image of the batch has no valid pixel at all
    T = O.sig_ntm_forward(torch.randn(19, 19, generator=torch.Generator().manual_seed(4)), class_dist(), 19, 0)
    lo, dlo, dTo = O.simt_head_fwd_bwd(lg, T, lab, (32, 64), torch.float64)
    for g in (1.0, -2.5, 0.0, 1e-8, 3e4):
        loss, dl, dT, _, err = run(emu, MODE_STEP, lg.numpy(), T.numpy(), lab.numpy(), (32, 64), scale=g, seed=5)
        assert err == 0 and abs(loss - float(lo)) <= TOL * abs(float(lo))
        assert not dl[1].any()                      # the all-ignored image receives exactly zero gradient
        if g == 0.0:
            assert not dl.any() and not dT.any()
        else:
            _check(dl, g * dlo.numpy(), f"dlogits, grad_out {g}")
            _check(dT, g * dTo.numpy(), f"dT, grad_out {g}")


@pytest.mark.parametrize("CK,C", [(41, 19), (64, 19), (64, 64)])
def test_widest_instantiation_up_to_64_channels(emu, CK, C):
    """(CPL, LPR) = (16, 4) serves 41..64 channels -- more than any configuration of the reference uses (K <= 15);
    covered here so that the instantiation the library ships is one that has run"""
    from oracle import simt_oracle as O
    K = CK - C
    lg, lab = O.synth_head_inputs(1, CK, 5, 9, 32, 64, seed=CK, coherent=True, ignore_frac=0.1, block=(6, 10))
    if C != 19:
        lab = torch.from_numpy(np.random.default_rng(CK).integers(0, C, size=tuple(lab.shape))).to(lab.dtype)
    Tm = torch.softmax(torch.randn(CK, C, generator=torch.Generator().manual_seed(4)), 1)
    lo, dlo, dTo = O.simt_head_fwd_bwd(lg, Tm, lab, (32, 64), torch.float64)
    loss, dl, dT, _, err = run(emu, MODE_STEP, lg.numpy(), Tm.numpy(), lab.numpy(), (32, 64), seed=3)
    assert err == 0 and abs(loss - float(lo)) <= TOL * abs(float(lo))
    _check(dl, dlo.numpy(), "dlogits")
    _check(dT, dTo.numpy(), "dT")
    rl, rdl = O.placeholder_fwd_bwd(lg, (32, 64), C, K, 0.3, 0.1, torch.float64)
    loss, dl_raw, _, stats, _ = run(emu, MODE_PLACE, lg.numpy(), None, None, (32, 64), thres=0.3, lam=0.1, C=C)
    assert abs(loss - float(rl)) <= TOL * abs(float(rl))
    assert rel_l2(dl_raw / stats[1], rdl.numpy()) <= TOL
