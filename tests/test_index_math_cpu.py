"""The kernels' pixel -> cell arithmetic of the align_corners=True bilinear resize (csrc/common.cuh: cell_of, lambda_of,
first_px_of_cell -- host/device functions, so this runs WITHOUT a GPU through the host-only hook
simt_debug_resize_tables) against torch's own interpolation weights (nn.Upsample at
tools/trainV2_simt.py:301,371-372 and evaluate_cityscapes.py:127-138).

torch's weights are read off F.interpolate of one-hot rows: out[j, X] is the weight source j gets at output X.  The
kernels read sources cell and cell + 1 with weights (1 - lambda, lambda); both descriptions must give the same weight
row for EVERY output pixel of every shape pair, bit for bit."""
import ctypes

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from simt_b200 import _lib

PAIRS = [(65, 512), (129, 1024), (65, 1024), (129, 2048), (81, 1024), (161, 2048),     # train / eval shapes of the reference
         (512, 512), (1024, 1024), (7, 7),                                            # identity (second interp_target)
         (512, 65), (100, 33), (33, 100), (2, 9), (9, 2), (2, 2), (3, 1000), (17, 70), (33, 133), (9, 64), (5, 32),
         (1, 5), (5, 1), (1, 1), (640, 1024), (1280, 2048), (97, 1000), (257, 999)]


def tables(n_in, n_out):
    lib = _lib.load()
    cell = np.zeros(n_out, dtype=np.int32)
    lam = np.zeros(n_out, dtype=np.float32)
    ncell = max(n_in - 1, 1)
    first = np.zeros(ncell + 1, dtype=np.int32)
    rc = lib.simt_debug_resize_tables(n_in, n_out, cell.ctypes.data, lam.ctypes.data, first.ctypes.data)
    assert rc == 0
    return cell, lam, first


def torch_weights(n_in, n_out):
    eye = torch.eye(n_in, dtype=torch.float32).reshape(1, n_in, 1, n_in)
    out = F.interpolate(eye, size=(1, n_out), mode="bilinear", align_corners=True)
    return out.reshape(n_in, n_out).numpy()          # [source j, output X]


@pytest.mark.parametrize("n_in,n_out", PAIRS)
def test_cell_and_lambda_reproduce_torch_weights(n_in, n_out):
    cell, lam, first = tables(n_in, n_out)
    ref = torch_weights(n_in, n_out)
    got = np.zeros_like(ref)
    X = np.arange(n_out)
    nxt = np.minimum(cell + 1, n_in - 1)
    np.add.at(got, (cell, X), np.float32(1.0) - lam)
    np.add.at(got, (nxt, X), lam)
    assert cell.min() >= 0 and cell.max() <= max(n_in - 2, 0)
    assert np.array_equal(got, ref), (n_in, n_out, float(np.abs(got - ref).max()))


@pytest.mark.parametrize("n_in,n_out", PAIRS)
def test_first_pixel_of_cell_is_the_partition_of_the_pixels_by_cell(n_in, n_out):
    cell, _, first = tables(n_in, n_out)
    ncell = max(n_in - 1, 1)
    assert first[0] == 0 and first[ncell] == n_out
    assert (np.diff(first) >= 0).all()
    for c in range(ncell):                       # pixels [first[c], first[c + 1]) are exactly the pixels of cell c
        assert (cell[first[c]:first[c + 1]] == c).all()
    assert np.array_equal(np.repeat(np.arange(ncell), np.diff(first)), cell)


def test_random_shape_pairs():
    rng = np.random.default_rng(5)
    for _ in range(150):
        n_in, n_out = int(rng.integers(1, 300)), int(rng.integers(1, 1200))
        cell, lam, first = tables(n_in, n_out)
        ref = torch_weights(n_in, n_out)
        got = np.zeros_like(ref)
        X = np.arange(n_out)
        np.add.at(got, (cell, X), np.float32(1.0) - lam)
        np.add.at(got, (np.minimum(cell + 1, n_in - 1), X), lam)
        assert np.array_equal(got, ref), (n_in, n_out)
        assert np.array_equal(np.repeat(np.arange(max(n_in - 1, 1)), np.diff(first)), cell), (n_in, n_out)
