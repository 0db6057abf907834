"""`resample` (preprocess.py:49-75 = scipy.ndimage.zoom(order=2)): the oracle restatement (CPU) and the CUDA
kernels (-m gpu) against golden vectors made by the unmodified reference function.  Tolerance, stated in
DESIGN.md section 0: 1e-12 absolute on float64 images of 8-bit range (scipy's compiled spline recursion is not
reproducible to the last bit); uint8 images and the thresholded mask of the reference pipeline's recipe
must be identical."""
import os

import numpy as np
import pytest

from tests import cases

TOL = 1e-12


@pytest.fixture(scope="module")
def golden_resample():
    return np.load(os.path.join(cases.GOLDEN_DIR, "resample.npz"), allow_pickle=False)


def _check(fn, g):
    for name in [str(n) for n in g["__names__"]]:
        a, fac, want = g[name + "/in"], float(g[name + "/factor"]), g[name + "/out"]
        got, vs = fn(a, np.array([0.0195, 0.0195, 0.0195]), fac)
        assert got.shape == want.shape and got.dtype == want.dtype, name
        if want.dtype == np.uint8:
            assert np.array_equal(got, want), name
        else:
            assert np.abs(got - want).max() <= TOL, (name, np.abs(got - want).max())
        assert np.array_equal(vs, g[name + "/voxelsize"])


def _check_lhdl(resample, gaussian, g):
    vol = g["lhdl/in"]
    smooth = gaussian(vol)
    got, vs = resample(smooth, np.array([0.0195] * 3), 2)
    want = g["lhdl/out"]
    assert got.shape == want.shape and np.abs(got - want).max() <= TOL
    n = want.size
    assert np.array_equal(got > 63, np.unpackbits(g["lhdl/mask63"])[:n].astype(bool).reshape(want.shape))
    assert np.array_equal(resample(vol, 1.0, 2)[0], g["lhdl/out_u8"])
    assert np.allclose(vs, 0.039)


def test_oracle_resample_matches_reference(golden_resample):
    from oracle import oracle as orc
    _check(orc.resample, golden_resample)
    _check_lhdl(orc.resample, lambda v: orc.gaussian(v, 1), golden_resample)


def test_zoom_tap_tables_match_the_oracle_rule():
    """the product's vectorised tap tables against the scalar statement of NI_ZoomShift in the oracle"""
    import math
    from ciclope_b200.preprocess import _zoom_taps
    for n_in, n_out in [(20, 10), (21, 10), (23, 12), (2, 1), (5, 2), (7, 14), (1, 1), (3, 2), (200, 100), (257, 128)]:
        idx, w, zero = _zoom_taps(n_in, n_out)
        zz = (n_in - 1) / (n_out - 1) if n_out > 1 else 1.0
        for k in range(n_out):
            cc = k * zz
            if cc < 0 or cc > n_in - 1:
                assert zero[k]
                continue
            assert not zero[k]
            start = int(math.floor(cc + 0.5)) - 1
            for tap in range(3):
                i = start + tap
                if n_in <= 1:
                    i = 0
                else:
                    s2 = 2 * n_in - 2
                    if i < 0:
                        i = s2 * int(-i / s2) + i
                        i = i + s2 if i <= 1 - n_in else -i
                    elif i >= n_in:
                        i -= s2 * int(i / s2)
                        if i >= n_in:
                            i = s2 - i
                assert idx[k, tap] == i, (n_in, n_out, k, tap)
            d = cc - math.floor(cc + 0.5)
            assert np.array_equal(w[k], [0.5 * (0.5 - d) ** 2, 0.75 - d * d, 0.5 * (0.5 + d) ** 2])


@pytest.mark.gpu
def test_gpu_resample_matches_reference(golden_resample):
    import torch
    from ciclope_b200 import gaussian, resample
    assert torch.cuda.is_available()
    _check(resample, golden_resample)
    _check_lhdl(resample, lambda v: gaussian(v, sigma=1, preserve_range=True), golden_resample)
    # tensor in -> tensor out, voxel size scaled like the reference does (list * int repeats the list!)
    t = torch.from_numpy(golden_resample["rand0_u8/in"]).cuda()
    out, vs = resample(t, np.array([1.0, 2.0, 3.0]), 2)
    assert out.is_cuda and np.array_equal(out.cpu().numpy(), golden_resample["rand0_u8/out"])
    assert np.array_equal(vs, [2.0, 4.0, 6.0])
