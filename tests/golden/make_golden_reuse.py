"""Golden outputs of the UNMODIFIED reference's remove_largest / fill_voids (preprocess.py:145-208),
the two other users of the labelling kernels.  Authoring container only (needs /root/reference):

    python tests/golden/make_golden_reuse.py

Output (committed): tests/golden/reuse.npz.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.ref_shim import load_reference  # noqa: E402
from tests import cases  # noqa: E402


def main():
    _, pre = load_reference()
    out = {}
    for name, bw in cases.ccl_cases():
        res = pre.remove_largest(bw.copy())
        assert res.dtype == np.bool_
        out["rl/" + name] = np.packbits(res.ravel())
    for name, img, kw in cases.fill_cases():
        arg = img.copy()
        res = pre.fill_voids(arg, **kw)
        assert res.dtype == img.dtype and (kw.get("makecopy") or res is arg)
        out["fv/" + name + "/in"] = img
        out["fv/" + name + "/out"] = res
        if kw.get("makecopy"):
            assert np.array_equal(arg, img)
    errs = {}
    for key, fn in [("remove_largest_all_background", lambda: pre.remove_largest(np.zeros((3, 3, 3), dtype=bool))),
                    ("fill_voids_all_foreground", lambda: pre.fill_voids(np.full((3, 3, 3), 9, dtype=np.uint8)))]:
        try:
            fn()
            errs[key] = ""
        except BaseException as e:  # noqa: BLE001
            errs[key] = type(e).__name__
    for k, v in errs.items():
        out["err/" + k] = np.asarray(v)
    path = os.path.join(cases.GOLDEN_DIR, "reuse.npz")
    np.savez_compressed(path, **out)
    print("reuse golden:", len(out), "arrays,", os.path.getsize(path), "bytes", errs)


if __name__ == "__main__":
    main()
