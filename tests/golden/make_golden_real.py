"""Real-data golden record (BASELINE.json configs[0]): the UNMODIFIED ciclope reference on the LHDL
trabecular bone stack of its own checkout (test_data/LHDL/3155_D_4_bc/cropped, 200 slices of 200 x 200
uint8; threshold 63 as in README.md:184), through remove_unconnected -> vol2ugrid -> mesh2voxelfe.
`trab_sample_mini3`, which configs[0] names, is not in the checkout: this stack is the
data the reference's example 01 and its README command line use.

Run in the authoring container only (needs /root/reference; ~4 minutes):

    python tests/golden/make_golden_real.py

Outputs (committed): tests/golden/real_lhdl.npz (the thresholded volume as packed bits + the grey values of
the kept voxels) and tests/golden/real_lhdl.json (counts, per-section SHA-256 of the binary deck and of the
grey-scale deck with PROPERTY cards).  The decks use the repo's own template files (tests/golden/tmpl_*.inp):
the tail of a deck is a verbatim copy of the template, every other section does not depend on it.
"""
import glob
import hashlib
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from ciclope_b200 import sections  # noqa: E402  (host-only helper: splits a deck into sections)
from oracle.ref_shim import load_reference  # noqa: E402
from tests import cases  # noqa: E402

STACK = "/root/reference/test_data/LHDL/3155_D_4_bc/cropped"
THRESHOLD = 63
VOXELSIZE = [0.0195, 0.0195, 0.0195]


def main():
    import logging

    import scipy.ndimage as ndi
    from PIL import Image
    logging.disable(logging.WARNING)
    vfe, pre = load_reference()
    files = sorted(glob.glob(os.path.join(STACK, "*.tif")))
    vol = np.stack([np.array(Image.open(f)) for f in files])
    assert vol.shape == (200, 200, 200) and vol.dtype == np.uint8
    bw = vol > THRESHOLD                                   # preprocess.segment: image > T
    rec = {"stack": "test_data/LHDL/3155_D_4_bc/cropped", "threshold": THRESHOLD, "shape": list(vol.shape),
           "voxelsize": VOXELSIZE, "foreground_voxels": int(bw.sum()),
           "components": int(ndi.label(bw)[1]), "volume_sha256": hashlib.sha256(vol.tobytes()).hexdigest()}
    t0 = time.time()
    L = pre.remove_unconnected(bw)
    rec["kept_voxels"] = int(L.sum())
    tmp = tempfile.mkdtemp()

    def deck_record(volume, name, **m):
        mesh = vfe.vol2ugrid(volume, VOXELSIZE)
        out = os.path.join(tmp, name + ".inp")
        vfe.mesh2voxelfe(mesh, cases.TEMPLATE, out, **m)
        deck = open(out, "rb").read()
        os.remove(out)
        return {"n_el": int(len(mesh.cells[0].data)), "n_used_nodes": int(len(np.unique(mesh.cells[0].data))),
                "n_points": int(len(mesh.points)), "deck_bytes": len(deck),
                "deck_sha256": hashlib.sha256(cases.mask_timestamp(deck)).hexdigest(),
                "set_sizes": {k: len(v) for k, v in list(mesh.point_sets.items()) + list(mesh.cell_sets.items())},
                "sections": sections.section_hashes(deck)}

    rec["binary"] = deck_record(L, "binary")
    print("binary deck: %.0f s" % (time.time() - t0), flush=True)
    masked = vol.copy()
    masked[~L] = 0                                          # ciclope example 01 / pipeline.py:237-239
    rec["greyscale"] = deck_record(masked, "grey", keywords=['NSET', 'ELSET', 'PROPERTY'],
                                   matprop={"file": [cases.MAT_LINEAR], "range": [[1, 255]]})
    print("grey-scale deck: %.0f s" % (time.time() - t0), flush=True)
    np.savez_compressed(os.path.join(cases.GOLDEN_DIR, "real_lhdl.npz"), bw=np.packbits(bw.ravel()),
                        kept=np.packbits(L.ravel()), kept_gv=vol[L], shape=np.asarray(vol.shape, dtype=np.int64))
    json.dump(rec, open(os.path.join(cases.GOLDEN_DIR, "real_lhdl.json"), "w"), indent=1, sort_keys=True)
    for f in ("real_lhdl.npz", "real_lhdl.json"):
        print(f, os.path.getsize(os.path.join(cases.GOLDEN_DIR, f)))
    print({k: rec[k] for k in ("foreground_voxels", "components", "kept_voxels")})


if __name__ == "__main__":
    main()
