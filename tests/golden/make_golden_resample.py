"""Golden vectors of `resample` (src/ciclope/utils/preprocess.py:49-75) from the UNMODIFIED reference function
(which calls scipy.ndimage.zoom(order=2) of the installed scipy):

    python tests/golden/make_golden_resample.py        (authoring container: needs /root/reference)

Output (committed): tests/golden/resample.npz -- small random float64 / uint8 volumes with several factors,
and the reference pipeline's own recipe (README.md:184: --smooth 1 -r 2 -t 63) on a 64^3 crop of the LHDL
stack: Gaussian (scipy.ndimage.gaussian_filter as skimage.filters.gaussian(preserve_range=True) calls it),
resample by 2, threshold 63."""
import glob
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.ref_shim import load_reference  # noqa: E402
from tests import cases  # noqa: E402

STACK = "/root/reference/test_data/LHDL/3155_D_4_bc/cropped"


def main():
    import scipy
    import scipy.ndimage as ndi
    from PIL import Image
    _, pre = load_reference()
    out = {}
    names = []
    rng = np.random.RandomState(2026)
    for k, (shape, fac) in enumerate([((20, 21, 23), 2), ((9, 30, 17), 2), ((16, 16, 16), 4), ((31, 12, 20), 3),
                                      ((10, 11, 12), 1.5), ((3, 2, 5), 2), ((7, 7, 7), 0.5), ((33, 17, 40), 2)]):
        for kind in ("f64", "u8"):
            a = rng.rand(*shape) * 255
            if kind == "u8":
                a = a.astype(np.uint8)
            vs = np.array([0.0195, 0.0195, 0.0195])
            res, vs2 = pre.resample(a, vs, fac)
            name = "rand%d_%s" % (k, kind)
            names.append(name)
            out[name + "/in"] = a
            out[name + "/factor"] = np.float64(fac)
            out[name + "/out"] = res
            out[name + "/voxelsize"] = vs2
    files = sorted(glob.glob(os.path.join(STACK, "*.tif")))
    vol = np.stack([np.array(Image.open(f)) for f in files[60:124]])[:, 70:134, 60:124].copy()
    smooth = ndi.gaussian_filter(vol.astype(np.float64), 1, mode='nearest', truncate=4.0)
    res, vs2 = pre.resample(smooth, np.array([0.0195] * 3), 2)
    out["lhdl/in"] = vol
    out["lhdl/out"] = res
    out["lhdl/mask63"] = np.packbits((res > 63).ravel())
    out["lhdl/voxelsize"] = vs2
    res8, _ = pre.resample(vol, np.array([0.0195] * 3), 2)
    out["lhdl/out_u8"] = res8
    out["__names__"] = np.asarray(names)
    out["__scipy__"] = np.asarray(scipy.__version__)
    np.savez_compressed(os.path.join(cases.GOLDEN_DIR, "resample.npz"), **out)
    print("resample.npz", os.path.getsize(os.path.join(cases.GOLDEN_DIR, "resample.npz")), "scipy", scipy.__version__,
          "lhdl mask fraction %.3f" % (res > 63).mean())


if __name__ == "__main__":
    main()
