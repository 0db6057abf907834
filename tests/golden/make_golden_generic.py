"""Golden decks of the UNMODIFIED reference's mesh2voxelfe for meshes that are not an untouched
vol2ugrid result (generic hexahedral meshio meshes; vol2ugrid meshes edited on the host).

Run in the authoring container only (needs /root/reference):

    python tests/golden/make_golden_generic.py

Output (committed): tests/golden/generic.npz -- the deck bytes per case (the inputs are regenerated
by tests/cases.py, whose generators are seeded and checked by a SHA-256 of every input array).
"""
import hashlib
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.ref_shim import load_reference  # noqa: E402
from tests import cases  # noqa: E402


def input_digest(case):
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(case["points"]).tobytes())
    for _, c in case["blocks"]:
        h.update(np.ascontiguousarray(c).tobytes())
    for v in case["cell_data"].values():
        for a in (v if isinstance(v, list) else [v]):
            h.update(np.ascontiguousarray(a).tobytes())
    for d in (case["point_sets"], case["cell_sets"]):
        for k, v in d.items():
            h.update(k.encode())
            h.update(np.asarray(v, dtype=np.int64).tobytes())
    return h.hexdigest()


def main():
    import logging
    logging.disable(logging.WARNING)
    vfe, _ = load_reference()
    import meshio   # the shim's meshio-5.0.0 container
    tmp = tempfile.mkdtemp()
    out = {}
    names = []
    for case in cases.generic_cases():
        name = case["name"]
        names.append(name)
        m = dict(case["m"])
        if "matprop" in m:
            m["matprop"] = m["matprop"]()
        mesh = meshio.Mesh(case["points"], case["blocks"], cell_data=cases.copy_cell_data(case["cell_data"]),
                           point_sets=dict(case["point_sets"]), cell_sets=dict(case["cell_sets"]))
        if not isinstance(case["cell_data"][list(case["cell_data"])[0]], list):
            mesh.cell_data = dict(case["cell_data"])          # plain-array pattern bypasses the constructor
        fileout = os.path.join(tmp, name + ".inp")
        vfe.mesh2voxelfe(mesh, fileout=fileout, **m)
        out[name + "/deck"] = np.frombuffer(cases.mask_timestamp(open(fileout, "rb").read()), dtype=np.uint8)
        out[name + "/digest"] = np.asarray(input_digest(case))
    for case in cases.edited_cases():
        name = case["name"]
        names.append(name)
        m = dict(case["m"])
        if "matprop" in m:
            m["matprop"] = m["matprop"]()
        mesh = case["edit"](vfe.vol2ugrid(case["vol"], **case["u"]))
        fileout = os.path.join(tmp, name + ".inp")
        vfe.mesh2voxelfe(mesh, fileout=fileout, **m)
        out[name + "/deck"] = np.frombuffer(cases.mask_timestamp(open(fileout, "rb").read()), dtype=np.uint8)
    out["__names__"] = np.asarray(names)
    path = os.path.join(cases.GOLDEN_DIR, "generic.npz")
    np.savez_compressed(path, **out)
    print("generic golden decks:", len(names), os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
