"""The sharded head step's exchange, executed on the CPU: the UNMODIFIED device code of csrc/step_kernels.cuh
(head_prep_kernel, head_finalize_kernel, head_finish_kernel), csrc/stepx_acquire.inc / stepx_cta0.inc (the prologue of
the fused kernel's MODE_STEPX instantiation) and csrc/step_xchg.cuh / xchg.cuh is compiled with g++ against a small
CUDA-on-CPU shim (tests/cpu_simt/cuda_shim.h: fibers for CUDA threads, barriers for __syncthreads / shuffles / votes,
random scheduling) and run for 2..8 emulated ranks, one OS thread each, with the launch sequence of
simt_head_step_sharded / simt_head_finish_sharded and HeadRunner's host logic (tests/cpu_simt/xchg_emul.cpp).
Peer stores (st.relaxed.sys) are delivered late and out of order, some only after the issuing kernel has ended; the
ranks are skewed by random sleeps.  Only the fused kernel's pixel loop is synthesised (dyadic partials, so every sum is
exact): loss, dT, the all-reduced stats, the gradient scale of every CTA and the protocol's bookkeeping words are checked
BIT FOR BIT on every rank after every step.

Complements tests/test_xchg_protocol_cpu.py (a model of the protocol) by running the real code -- indexing of slots,
tags, parities, staging offsets, tickets.  Negative controls show the harness sees what it should: the order that
dead-locked 8 GPUs in round 2 (peers released before the old slots are read), a synchronous step that does not drain
deferred steps first, and a rank that dies (the survivors must poison their outputs and raise
SIMT_ERRBIT_XCHG_TIMEOUT, never continue with a partial sum).  No GPU."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpu_simt", "xchg_emul.cpp")
OUT = os.path.join(ROOT, "build", "cpu_simt")

pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not found")


def _build(name, *defines):
    os.makedirs(OUT, exist_ok=True)
    exe = os.path.join(OUT, name)
    deps = [SRC, os.path.join(ROOT, "tests", "cpu_simt", "cuda_shim.h")] + [
        os.path.join(ROOT, "simt_b200", "csrc", f)
        for f in ("xchg.cuh", "step_xchg.cuh", "step_kernels.cuh", "stepx_acquire.inc", "stepx_cta0.inc")]
    if os.path.exists(exe) and os.path.getmtime(exe) >= max(os.path.getmtime(d) for d in deps):
        return exe
    cmd = ["g++", "-O0", "-g", "-std=c++17", "-pthread", "-Wno-unknown-pragmas", "-I", os.path.join(ROOT, "tests", "cpu_simt"),
           "-I", os.path.join(ROOT, "simt_b200", "csrc"), *defines, SRC, "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-4000:]
    return exe


@pytest.fixture(scope="module")
def emul():
    return _build("xchg_emul")


@pytest.fixture(scope="module")
def emul_push_first():
    return _build("xchg_emul_bug", "-DSIMT_EMU_BUG_PUSH_FIRST")


@pytest.fixture(scope="module")
def emul_bench_geometry():
    return _build("xchg_emul_c19", "-DEMU_C=19", "-DEMU_CK=19")


def _run(exe, *args, env=None, timeout=400):
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run([exe, *map(str, args)], capture_output=True, text=True, timeout=timeout, env=e)
    return r.returncode, (r.stdout + r.stderr)[-3000:]


@pytest.mark.parametrize("world", [2, 3, 8])
@pytest.mark.parametrize("mode", ["sync", "pipelined", "mixed", "announce_sync", "pipelined_noannounce"])
def test_real_exchange_code_is_bit_exact_on_every_rank(emul, mode, world):
    for seed in (1, 2):
        rc, out = _run(emul, world, 8, mode, seed)
        assert rc == 0, out


def test_synchronous_form_long_run_eight_ranks(emul):
    """what bench.py --gpus 8 runs: slot reuse over both parities and all four count rows, many times"""
    rc, out = _run(emul, 8, 40, "sync", 7)
    assert rc == 0, out


@pytest.mark.parametrize("world", [2, 8])
@pytest.mark.parametrize("mode", ["sync", "pipelined", "mixed"])
def test_bench_geometry_19_channels(emul_bench_geometry, mode, world):
    """C = CK = 19 as in bench.py: 13 finalize blocks, 363 stats values in 2 finish blocks, slots of 2 + 19 * 64 entries"""
    rc, out = _run(emul_bench_geometry, world, 6, mode, 5)
    assert rc == 0, out


def test_prep_kernel_label_dtypes_alignments_sizes(emul):
    """head_prep_kernel alone: uint8 / int64 labels (negative and out-of-range values), ignore labels inside and outside
    the uint8 range, misaligned label and dLogits pointers, empty and ragged sizes, several grid / block shapes"""
    for seed in (1, 2, 3):
        rc, out = _run(emul, "prep", seed)
        assert rc == 0, out


def test_single_gpu_step_kernels(emul, emul_bench_geometry):
    """world = 1: head_prep_kernel + head_finalize_kernel as simt_head_step launches them (no exchange): count, zeroing,
    loss, dT scaled by grad_out / N on the device, tiles re-zeroed, scheduler re-armed"""
    for exe in (emul, emul_bench_geometry):
        rc, out = _run(exe, 1, 6, "sync", 3)
        assert rc == 0, out


def test_open_set_geometry_23_channels():
    """K = 4 (BASELINE configs[2]): CK = 23 rows of T, dT tile rows of 24"""
    exe = _build("xchg_emul_k4", "-DEMU_C=19", "-DEMU_CK=23", "-DEMU_CKP=24")
    for world, mode in ((1, "sync"), (2, "sync"), (4, "mixed")):
        rc, out = _run(exe, world, 5, mode, 3)
        assert rc == 0, out


@pytest.mark.parametrize("mode", ["sync", "pipelined", "mixed"])
def test_a_dead_rank_poisons_the_survivors(emul, mode):
    rc, out = _run(emul, 4, 6, mode, 4, 2, 4)      # rank 2 stops before step 4
    assert rc == 0, out
    assert "fault injected" in out


def test_negative_control_sync_step_without_draining_deferred_steps(emul):
    """HeadRunner.step finishes the deferred steps before a synchronous one; without that the real code loses words."""
    bad = 0
    for seed in (1, 2, 3):
        rc, _ = _run(emul, 4, 9, "mixed_nodrain", seed)
        bad += rc != 0
    assert bad == 3


def test_negative_control_peers_released_before_old_slots_are_read(emul_push_first):
    """The order that dead-locked 8 GPUs in round 2 (the peers are released -- stats and next count pushed -- before
    this rank has read the slots they then overwrite), rebuilt with -DSIMT_EMU_BUG_PUSH_FIRST: the emulation loses
    words within one short run."""
    bad = 0
    for seed in (1, 2, 3):
        rc, _ = _run(emul_push_first, 8, 9, "pipelined", seed, env={"XCHG_EMUL_MAX_SPINS": "3000"})
        bad += rc != 0
    assert bad == 3
