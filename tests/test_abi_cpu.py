"""CPU-only checks of the C-ABI boundary and of the facade's host logic (no compute calls)."""
import ctypes
import json
import os
import re

import numpy as np
import pytest

from tests import cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "ciclope_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cfe_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from ciclope_b200 import _lib, build
    build.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), "libciclope_b200.so does not export %s" % name
    # the ctypes table binds exactly the declared entry points
    assert sorted(_lib.EXPORTED_SYMBOLS) == declared
    lib.cfe_version.restype = ctypes.c_int
    assert lib.cfe_version() == 1


def test_struct_layouts_match_header():
    from ciclope_b200._lib import CfeItemSeg, CfeSetDesc
    assert ctypes.sizeof(CfeSetDesc) == 8 + 14 * 8
    assert ctypes.sizeof(CfeItemSeg) == 8 + 5 * 8


def test_no_cpu_fallback_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from ciclope_b200 import _lib, remove_unconnected, vol2ugrid
    with pytest.raises(_lib.CfeError):
        remove_unconnected(np.ones((2, 2, 2), dtype=bool))
    with pytest.raises(_lib.CfeError):
        vol2ugrid(np.ones((2, 2, 2), dtype=bool))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "ciclope_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("no CPU fallback", ""), "%s mentions the oracle" % f
                assert "/root/reference" not in text


def test_matpropdictionary_matches_reference_behaviour(tmp_path):
    from ciclope_b200 import matpropdictionary
    errs = json.load(open(os.path.join(cases.GOLDEN_DIR, "errors.json")))
    assert matpropdictionary(None) is None
    assert matpropdictionary([]) is None
    assert matpropdictionary(["a.inp", "1", "100", "b.inp", "101", "255"]) == errs["matpropdictionary_triples"]
    assert errs["matpropdictionary_bad_len"] == "OSError"
    with pytest.raises(IOError):
        matpropdictionary(["a.inp", "1"])
    assert errs["matpropdictionary_missing_file"] == "OSError"
    with pytest.raises(IOError):
        matpropdictionary(["/nonexistent/a.inp"])
    assert matpropdictionary([cases.MAT_CONST]) == {"file": [cases.MAT_CONST]}


def test_threshold_and_tables_host_logic(golden):
    from ciclope_b200 import voxelFE as v
    for gvmin, want in [(0, 0), (100, 100), (99.5, 99), (-1, -1), (-0.5, -1), (255, 255), (300, 255), (0.999, 0)]:
        assert v._gv_threshold(gvmin) == want
        x = np.arange(256)
        assert np.array_equal(x > gvmin, x > want)
    # coordinate tables reproduce the reference's points (dtype and values) for every golden case
    for c in cases.deck_cases():
        vol = np.asarray(c["vol"])
        Z, Y, X = vol.shape
        xs, ys, zs = v._coordinate_tables(c["u"]["voxelsize"], Z, Y, X)
        pts = golden[c["name"] + "/points"].reshape(Z + 1, Y + 1, X + 1, 3)
        assert xs.dtype == pts.dtype, c["name"]
        assert np.array_equal(pts[0, 0, :, 0], xs) and np.array_equal(pts[0, :, 0, 1], ys) \
            and np.array_equal(pts[:, 0, 0, 2], zs), c["name"]


def test_boundary_boxes_host_logic():
    from ciclope_b200 import voxelFE as v
    Z, Y, X = 8, 9, 12
    b = v._boundary_boxes("plane", 2, 1, Z, Y, X)
    assert b[0] == ((0, Z), (0, Y), (0, 2)) and b[1] == ((0, Z), (0, Y), (10, 12))
    assert b[4] == ((0, 2), (0, Y), (0, X)) and b[5] == ((6, 8), (0, Y), (0, X))
    b = v._boundary_boxes("plane", 20, 1, Z, Y, X)
    assert b[0] == ((0, Z), (0, Y), (0, X)) and b[1] == ((0, Z), (0, Y), (0, X))
    b = v._boundary_boxes("plane", 0, 1, Z, Y, X)
    assert all(v._box_count(x) == 0 for x in b)
    b = v._boundary_boxes("exact", 1, 1, Z, Y, X)
    assert b[0] == ((0, Z + 1), (0, Y + 1), (0, 1)) and b[1] == ((0, Z + 1), (0, Y + 1), (12, 13))
    b = v._boundary_boxes("exact", 1, 0, Z, Y, X)
    assert all(v._box_count(x) == 0 for x in b)


def test_vol2ugrid_argument_errors_before_any_gpu_work():
    from ciclope_b200 import vol2ugrid
    with pytest.raises(ValueError):
        vol2ugrid(np.ones((2, 2, 2), dtype=bool), locking_strategy="both")
    with pytest.raises(TypeError):
        vol2ugrid(np.ones((2, 2, 2), dtype=bool), voxelsize=(1, 1, 1))


def test_pyset_replay_matches_this_interpreter():
    """cfe_host_pyset_order (host C code, csrc/pyset.cu) must give list(set(...)) of the running CPython:
    that order is what the reference writes into Fixed_Displacement_Coordinates (voxelFE.py:449)."""
    import numpy as np
    from ciclope_b200 import parosol
    rng = np.random.default_rng(2)
    for n, hi in [(1, 2), (7, 2), (40, 3), (700, 9), (6000, 30), (49000, 60), (52000, 200), (260000, 120)]:
        rows = np.stack([rng.integers(0, hi, n), rng.integers(0, hi, n), rng.integers(0, hi, n),
                         rng.integers(0, 3, n)], axis=1)
        want = np.array(list(set(map(tuple, rows.tolist()))), dtype=np.int64).reshape(-1, 4)
        got = parosol._pyset_replay(rows)
        assert got.shape == want.shape and np.array_equal(got, want), n
    # boundary-plane shaped input: element-major corner expansion with many duplicates
    els = np.argwhere(rng.random((90, 110)) < 0.4)
    els = np.hstack((np.zeros((len(els), 1), dtype=np.int64), els))
    rows = parosol._corner_rows(els, (0, 1, 2))
    want = np.array(list(set(map(tuple, rows.tolist()))), dtype=np.int64).reshape(-1, 4)
    assert np.array_equal(parosol._pyset_replay(rows), want)
    assert parosol._replay_matches_interpreter()


def test_host_logic_of_the_preprocessing_facades():
    """Pure host pieces of parosol.py / preprocess.py against the oracle's restatement of the reference."""
    import numpy as np
    from scipy.ndimage import _filters
    from ciclope_b200 import parosol, preprocess
    from oracle import oracle as orc
    rng = np.random.default_rng(8)
    # bounding box with padding (recon_utils.py:264-298) from the three projections
    for shape in [(7, 9, 11), (4, 20, 6), (12, 5, 5)]:
        for pad in (0, 1, 3, 9):
            bw = np.zeros(shape, dtype=bool)
            bw[1:-1, 2:-1, 1:-2] = rng.random(tuple(np.subtract(shape, (2, 3, 3)))) < 0.3
            if not bw.any():
                bw[1, 2, 1] = True
            proj = (np.max(np.max(bw, 0), 1), np.max(np.max(bw, 0), 0), np.max(np.max(bw, 1), 1))
            assert preprocess._bbox_from_projections(*proj, bw.shape, pad) == orc.bbox_pad(bw, pad)
    # list.index(True) on a grey projection looks for entries EQUAL to 1
    assert parosol._first_last_true(np.array([0, 2, 1, 1, 2, 0], dtype=np.uint8)) == (2, 3)
    with pytest.raises(ValueError):
        parosol._first_last_true(np.array([0, 2, 5], dtype=np.uint8))
    # insertion sequence of _collect_from_elements (preprocess.py:449-456)
    els = np.array([[0, 1, 2], [0, 1, 3], [5, 0, 0]])
    seq = []
    for z, y, x in els.tolist():
        for dz in (0, 1):
            for dy in (0, 1):
                for dx in (0, 1):
                    for d in (0, 2):
                        seq.append((z + dz, y + dy, x + dx, d))
    assert parosol._corner_rows(els, (0, 2)).tolist() == [list(t) for t in seq]
    want = set()
    for t in seq:
        want.add(t)                                   # the reference's loop; list(set) is the order it writes
    assert parosol._ordered_unique(parosol._corner_rows(els, (0, 2))).tolist() == [list(t) for t in list(want)]
    # Gaussian weights = scipy's
    for sigma in (0.5, 1.0, 2.7):
        r = int(4.0 * sigma + 0.5)
        assert np.array_equal(preprocess._gaussian_kernel1d(sigma, r), _filters._gaussian_kernel1d(sigma, 0, r))
