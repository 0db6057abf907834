"""Generate the committed golden vectors by running the UNMODIFIED ciclope reference.

Run in the authoring container only (needs /root/reference):

    python tests/golden/make_golden.py

Outputs (committed): tests/golden/ccl.npz, tests/golden/decks.npz, tests/golden/midsize.json,
tests/golden/errors.json.  Inputs are stored next to the expected outputs so the vectors do
not depend on NumPy's random streams.
"""
import hashlib
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.ref_shim import load_reference  # noqa: E402
from tests import cases  # noqa: E402


def main():
    import logging
    logging.disable(logging.WARNING)
    vfe, pre = load_reference()
    tmp = tempfile.mkdtemp()

    # ---- remove_unconnected
    ccl = {}
    for name, bw in cases.ccl_cases():
        out = pre.remove_unconnected(bw)
        assert out.dtype == np.bool_ and out.shape == bw.shape
        ccl[name + "/shape"] = np.asarray(bw.shape, dtype=np.int64)
        ccl[name + "/in"] = np.packbits(bw.ravel())
        ccl[name + "/out"] = np.packbits(out.ravel())
    np.savez_compressed(os.path.join(cases.GOLDEN_DIR, "ccl.npz"), **ccl)

    # ---- vol2ugrid + mesh2voxelfe
    decks = {}
    names = []
    for case in cases.deck_cases():
        name = case["name"]
        names.append(name)
        u = dict(case["u"])
        m = dict(case["m"])
        if "matprop" in m:
            m["matprop"] = m["matprop"]()
        mesh, refn = vfe.vol2ugrid(case["vol"], refnodes=True, **u)
        fileout = os.path.join(tmp, name + ".inp")
        vfe.mesh2voxelfe(mesh, fileout=fileout, **m)
        deck = cases.mask_timestamp(open(fileout, "rb").read())
        decks[name + "/vol"] = np.asarray(case["vol"])
        decks[name + "/deck"] = np.frombuffer(deck, dtype=np.uint8)
        decks[name + "/points"] = mesh.points
        decks[name + "/cells"] = mesh.cells[0].data.reshape(-1, 8).astype(np.int64)
        decks[name + "/gv"] = np.asarray(mesh.cell_data["GV"][0])
        for k, v in mesh.point_sets.items():
            assert isinstance(v, list)
            decks[name + "/ps/" + k] = np.asarray(v, dtype=np.int64)
        for k, v in mesh.cell_sets.items():
            decks[name + "/cs/" + k] = np.asarray(v, dtype=np.int64)
        for k, v in refn.items():
            decks[name + "/refnode/" + k] = np.asarray(v, dtype=np.float64)
    decks["__names__"] = np.asarray(names)
    np.savez_compressed(os.path.join(cases.GOLDEN_DIR, "decks.npz"), **decks)

    # ---- mid-size decks: hashes only
    mid = {}
    for case in cases.midsize_cases():
        m = dict(case["m"])
        if "matprop" in m:
            m["matprop"] = m["matprop"]()
        mesh = vfe.vol2ugrid(case["vol"], **case["u"])
        fileout = os.path.join(tmp, case["name"] + ".inp")
        vfe.mesh2voxelfe(mesh, fileout=fileout, **m)
        deck = cases.mask_timestamp(open(fileout, "rb").read())
        mid[case["name"]] = dict(sha256=hashlib.sha256(deck).hexdigest(), size=len(deck),
                                 n_el=int(len(mesh.cells[0].data)),
                                 n_used_nodes=int(len(np.unique(mesh.cells[0].data))),
                                 vol_sha256=hashlib.sha256(np.ascontiguousarray(case["vol"]).tobytes()).hexdigest())
    json.dump(mid, open(os.path.join(cases.GOLDEN_DIR, "midsize.json"), "w"), indent=1, sort_keys=True)

    # ---- error behaviour of the reference (exception class names)
    errs = {}

    def record(key, fn):
        try:
            fn()
            errs[key] = None
        except BaseException as e:  # noqa: BLE001 -- we record whatever the reference raises
            errs[key] = type(e).__name__

    b = cases.rand_binary((5, 7, 9), 0.4, 1)
    empty = np.zeros((3, 3, 3), dtype=bool)
    record("remove_unconnected_all_background", lambda: pre.remove_unconnected(empty))
    # a uint8 image with several non-zero values: the reference labels regions of equal value and returns
    # the largest one; ciclope_b200 raises ValueError instead (documented deviation)
    multi = (cases.rand_binary((6, 8, 10), 0.6, 9) * np.arange(1, 11, dtype=np.uint8) % 4).astype(np.uint8)
    record("remove_unconnected_multivalued_uint8", lambda: pre.remove_unconnected(multi))
    record("vol2ugrid_bad_strategy", lambda: vfe.vol2ugrid(b, locking_strategy="both"))
    record("vol2ugrid_all_background", lambda: vfe.vol2ugrid(empty))
    record("mesh2voxelfe_all_background",
           lambda: vfe.mesh2voxelfe(vfe.vol2ugrid(empty), cases.TEMPLATE, os.path.join(tmp, "e.inp")))
    g = cases.rand_grey((4, 4, 6), 0.5, 3)
    record("matprop_two_files_no_range",
           lambda: vfe.mesh2voxelfe(vfe.vol2ugrid(g), cases.TEMPLATE, os.path.join(tmp, "e.inp"),
                                    matprop={"file": [cases.MAT_CONST, cases.MAT_POWER]}))
    record("matprop_range_exceeds",
           lambda: vfe.mesh2voxelfe(vfe.vol2ugrid(g), cases.TEMPLATE, os.path.join(tmp, "e.inp"),
                                    matprop={"file": [cases.MAT_CONST], "range": [[1, 300]]}))
    record("matprop_range_negative",
           lambda: vfe.mesh2voxelfe(vfe.vol2ugrid(g), cases.TEMPLATE, os.path.join(tmp, "e.inp"),
                                    matprop={"file": [cases.MAT_CONST], "range": [[-1, 30]]}))
    record("matprop_overlap",
           lambda: vfe.mesh2voxelfe(vfe.vol2ugrid(g), cases.TEMPLATE, os.path.join(tmp, "e.inp"),
                                    matprop={"file": [cases.MAT_CONST, cases.MAT_POWER],
                                             "range": [[1, 100], [100, 255]]}))
    record("matpropdictionary_bad_len", lambda: vfe.matpropdictionary(["a.inp", "1"]))
    record("matpropdictionary_missing_file", lambda: vfe.matpropdictionary(["/nonexistent/a.inp"]))
    errs["matpropdictionary_none"] = vfe.matpropdictionary(None)
    errs["matpropdictionary_empty"] = vfe.matpropdictionary([])
    errs["matpropdictionary_triples"] = vfe.matpropdictionary(["a.inp", "1", "100", "b.inp", "101", "255"])
    json.dump(errs, open(os.path.join(cases.GOLDEN_DIR, "errors.json"), "w"), indent=1, sort_keys=True)

    # the reference's own known answers must hold for the vectors we just stored
    d = decks
    assert d["trab_gvmin100_exact/cells"].shape == (209, 8)
    assert d["trab_gvmin100_exact/points"].shape == (1331, 3)
    assert d["trab_gvmin100_exact/ps/NODES_X0Y0Z0"].tolist() == [18, 19]
    lines = bytes(d["trab_gvmin100_plane/deck"]).decode().split("\n")
    assert lines[99] == '       359,     0.840000,     1.200000,     0.240000'
    assert lines[860] == '1315,1316,1326,1327,'
    print("golden vectors written:", len(names), "deck cases,", len(cases.ccl_cases()), "ccl cases")
    for f in ("ccl.npz", "decks.npz", "midsize.json", "errors.json"):
        print(f, os.path.getsize(os.path.join(cases.GOLDEN_DIR, f)))


if __name__ == "__main__":
    main()
