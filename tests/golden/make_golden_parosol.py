"""Golden datasets of vol2h5ParOSol (voxelFE.py:285-472), segment (preprocess.py:20-46, explicit threshold)
and add_cap (preprocess.py:210-231), produced by the UNMODIFIED reference in this container.

    python tests/golden/make_golden_parosol.py

h5py is not installed: oracle/ref_shim.py records the arguments of the reference's create_dataset calls
(np.array(data, dtype=dtype), which is what h5py would store).  Output: tests/golden/parosol.npz.
"""
import hashlib
import os
import sys

os.environ.setdefault("TQDM_DISABLE", "1")

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402
from tests import cases  # noqa: E402


def main():
    vfe, pre = ref_shim.load_reference()
    out = {}
    n_err = 0
    for name, vol, kw in cases.parosol_cases():
        kw = dict(kw)
        top = kw.pop("topDisplacement")
        try:
            vfe.vol2h5ParOSol(vol.copy(), name, top, **kw)
        except BaseException as e:  # noqa: BLE001
            out["po/" + name + "/error"] = np.asarray(type(e).__name__)
            n_err += 1
            continue
        ds = ref_shim.H5PY_FILES[name]
        out["po/" + name + "/names"] = np.asarray(list(ds))
        for k, a in ds.items():
            if k == "Image" or a.nbytes > 4096:       # large arrays: kept as dtype / shape / SHA-256
                out["po/" + name + "/" + k] = np.asarray([a.dtype.str, str(a.shape),
                                                           hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()])
            else:
                out["po/" + name + "/" + k] = a
    for name, img, t in cases.segment_cases():
        bw, T = pre.segment(img, t)
        assert bw.dtype == np.bool_
        out["seg/" + name + "/bw"] = np.packbits(bw.ravel())
        out["seg/" + name + "/T"] = np.asarray(T)
    rng = np.random.default_rng(9)
    for i, (shape, th, val) in enumerate([((4, 5, 6), 2, 1), ((3, 8, 8), 1, 255), ((6, 4, 9), 3, 0)]):
        img = rng.integers(0, 256, shape).astype(np.uint8)
        res = pre.add_cap(img, th, val)
        out["cap/%d/in" % i] = img
        out["cap/%d/args" % i] = np.asarray([th, val])
        out["cap/%d/out" % i] = res
    import warnings
    n_emb_err = 0
    for name, img, kw in cases.embed_cases():
        arg = img.copy()
        try:
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                res, bw = pre.embed(arg, **kw)
        except BaseException as e:  # noqa: BLE001
            out["emb/" + name + "/error"] = np.asarray(type(e).__name__)
            n_emb_err += 1
            continue
        assert bw.dtype == np.bool_ and res.dtype == img.dtype and (kw["makecopy"] or res is arg)
        out["emb/" + name + "/out"] = res
        out["emb/" + name + "/bw"] = np.packbits(bw.ravel())
        out["emb/" + name + "/arg_after"] = np.asarray(hashlib.sha256(arg.tobytes()).hexdigest())
    print("embed:", len(cases.embed_cases()), "cases,", n_emb_err, "error cases")
    # periosteummask: the reference's own loop / fill_holes / remove_unconnected; its two scikit-image calls
    # (disk, binary_closing) go through oracle/ref_shim.py's restatement on top of scipy
    n_per_err = 0
    for name, bw, kw in cases.periosteum_cases():
        try:
            res = pre.periosteummask(bw.copy(), **kw)
        except BaseException as e:  # noqa: BLE001
            out["per/" + name + "/error"] = np.asarray(type(e).__name__)
            n_per_err += 1
            continue
        assert res.dtype == np.bool_ and res.shape == bw.shape
        out["per/" + name + "/out"] = np.packbits(res.ravel())
    print("periosteummask:", len(cases.periosteum_cases()), "cases,", n_per_err, "error cases")
    path = os.path.join(cases.GOLDEN_DIR, "parosol.npz")
    np.savez_compressed(path, **out)
    print("parosol golden:", len(out), "arrays,", os.path.getsize(path), "bytes,", n_err, "error cases")


if __name__ == "__main__":
    main()
