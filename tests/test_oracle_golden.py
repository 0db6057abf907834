"""Pins the CPU oracle (oracle/) against golden vectors produced by the unmodified reference
(tests/golden/make_golden.py).  CPU only."""
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import oracle as orc
from tests import cases

DECK_CASES = {c["name"]: c for c in cases.deck_cases()}


def _unpack(g, name):
    shape = tuple(g[name + "/shape"])
    n = int(np.prod(shape))
    return (np.unpackbits(g[name + "/in"])[:n].astype(bool).reshape(shape),
            np.unpackbits(g[name + "/out"])[:n].astype(bool).reshape(shape))


@pytest.mark.parametrize("name", [n for n, _ in cases.ccl_cases()])
def test_oracle_remove_unconnected(golden_ccl, name):
    bw, want = _unpack(golden_ccl, name)
    got = orc.remove_unconnected(bw)
    assert got.dtype == np.bool_ and got.shape == bw.shape
    assert np.array_equal(got, want)


@pytest.mark.parametrize("name", [n for n, _ in cases.ccl_cases()])
def test_oracle_remove_largest(golden_reuse, name):
    bw = dict(cases.ccl_cases())[name]
    want = np.unpackbits(golden_reuse["rl/" + name])[:bw.size].astype(bool).reshape(bw.shape)
    got = orc.remove_largest(bw)
    assert got.dtype == np.bool_ and np.array_equal(got, want)


@pytest.mark.parametrize("name", [c[0] for c in cases.fill_cases()])
def test_oracle_fill_voids(golden_reuse, name):
    _, img, kw = [c for c in cases.fill_cases() if c[0] == name][0]
    assert np.array_equal(golden_reuse["fv/" + name + "/in"], img)
    arg = img.copy()
    got = orc.fill_voids(arg, **kw)
    want = golden_reuse["fv/" + name + "/out"]
    assert got.dtype == want.dtype and np.array_equal(got, want)
    assert (got is arg) != bool(kw.get("makecopy"))


def test_oracle_reuse_errors(golden_reuse):
    assert str(golden_reuse["err/remove_largest_all_background"]) == "ValueError"
    assert str(golden_reuse["err/fill_voids_all_foreground"]) == "ValueError"
    with pytest.raises(ValueError):
        orc.remove_largest(np.zeros((3, 3, 3), dtype=bool))
    with pytest.raises(ValueError):
        orc.fill_voids(np.full((3, 3, 3), 9, dtype=np.uint8))


def test_cases_regenerate_golden_inputs(golden, golden_ccl):
    # the seeded generators in tests/cases.py still produce the stored inputs
    for name, bw in cases.ccl_cases():
        assert np.array_equal(_unpack(golden_ccl, name)[0], bw), name
    for name, c in DECK_CASES.items():
        assert np.array_equal(golden[name + "/vol"], c["vol"]), name
        assert golden[name + "/vol"].dtype == np.asarray(c["vol"]).dtype, name


@pytest.mark.parametrize("name", list(DECK_CASES))
def test_oracle_vol2ugrid_and_deck(golden, name, tmp_path):
    c = DECK_CASES[name]
    mesh, refn = orc.vol2ugrid(c["vol"], refnodes=True, **c["u"])
    g = golden
    assert mesh.points.dtype == g[name + "/points"].dtype
    assert np.array_equal(mesh.points, g[name + "/points"])
    assert np.array_equal(mesh.cells[0].data.reshape(-1, 8), g[name + "/cells"])
    assert mesh.cells[0].type == "hexahedron"
    assert mesh.cell_data["GV"][0].dtype == g[name + "/gv"].dtype
    assert np.array_equal(mesh.cell_data["GV"][0], g[name + "/gv"])
    assert list(mesh.point_sets) == ["NODES_X0", "NODES_X1", "NODES_Y0", "NODES_Y1", "NODES_Z0", "NODES_Z1",
                                     "NODES_X0Y0Z0", "NODES_X0Y0Z1"]
    assert list(mesh.cell_sets) == ["CELLS_X0", "CELLS_X1", "CELLS_Y0", "CELLS_Y1", "CELLS_Z0", "CELLS_Z1"]
    for k, v in mesh.point_sets.items():
        assert isinstance(v, list)
        assert v == g[name + "/ps/" + k].tolist(), k
    for k, v in mesh.cell_sets.items():
        assert v == g[name + "/cs/" + k].tolist(), k
    for k, v in refn.items():
        assert np.array_equal(v, g[name + "/refnode/" + k]), k  # bit-exact sequential sums
    m = dict(c["m"])
    if "matprop" in m:
        m["matprop"] = m["matprop"]()
    out = str(tmp_path / "deck.inp")
    orc.mesh2voxelfe(mesh, fileout=out, **m)
    got = cases.mask_timestamp(open(out, "rb").read())
    assert got == bytes(g[name + "/deck"])


@pytest.mark.parametrize("case", cases.generic_cases(), ids=lambda c: c["name"])
def test_oracle_generic_mesh_deck(golden_generic, case, tmp_path):
    # mesh2voxelfe of meshes that are not an untouched vol2ugrid result (voxelFE.py:612-688)
    from tests.golden.make_golden_generic import input_digest
    assert input_digest(case) == str(golden_generic[case["name"] + "/digest"])
    m = dict(case["m"])
    if "matprop" in m:
        m["matprop"] = m["matprop"]()
    out = str(tmp_path / "deck.inp")
    orc.mesh2voxelfe(cases.PlainMesh(case), fileout=out, **m)
    assert cases.mask_timestamp(open(out, "rb").read()) == bytes(golden_generic[case["name"] + "/deck"])


@pytest.mark.parametrize("case", cases.edited_cases(), ids=lambda c: c["name"])
def test_oracle_edited_mesh_deck(golden_generic, case, tmp_path):
    m = dict(case["m"])
    if "matprop" in m:
        m["matprop"] = m["matprop"]()
    mesh = case["edit"](orc.vol2ugrid(case["vol"], **case["u"]))
    out = str(tmp_path / "deck.inp")
    orc.mesh2voxelfe(mesh, fileout=out, **m)
    assert cases.mask_timestamp(open(out, "rb").read()) == bytes(golden_generic[case["name"] + "/deck"])


@pytest.mark.parametrize("case", cases.midsize_cases(), ids=lambda c: c["name"])
def test_oracle_midsize_hash(case, tmp_path):
    want = json.load(open(os.path.join(cases.GOLDEN_DIR, "midsize.json")))[case["name"]]
    m = dict(case["m"])
    if "matprop" in m:
        m["matprop"] = m["matprop"]()
    mesh = orc.vol2ugrid(case["vol"], **case["u"])
    out = str(tmp_path / "deck.inp")
    orc.mesh2voxelfe(mesh, fileout=out, **m)
    got = cases.mask_timestamp(open(out, "rb").read())
    assert len(mesh.cells[0].data) == want["n_el"]
    assert len(got) == want["size"]
    assert hashlib.sha256(got).hexdigest() == want["sha256"]


def test_reference_known_answers(golden):
    """The reference's own unit-test assertions (test/test_ciclope.py:11-15, 75-87, 89-132)."""
    tr = cases.trab()
    bw = tr > 100
    assert int(bw.sum()) - int(orc.remove_unconnected(bw).sum()) == 4
    mesh = orc.vol2ugrid(tr, [0.12, 0.12, 0.12], GVmin=100, locking_strategy="exact", node_level_lock_num=1)
    assert mesh.cells[0].type == "hexahedron"
    assert mesh.cells[0].data.shape == (209, 8)
    assert mesh.points.shape == (1331, 3)
    assert mesh.point_sets['NODES_X0Y0Z0'] == [18, 19]
    assert (mesh.points[100] == np.array([0.12, 1.08, 0])).all()
    lines = bytes(golden["trab_gvmin100_plane/deck"]).decode().split("\n")
    assert any('*ELEMENT, TYPE=C3D8, ELSET=SET255' in ln for ln in lines)
    assert any('*ELSET, ELSET=CELLS_Z1' in ln for ln in lines)
    assert any('*NSET, NSET=NODES_Z0' in ln for ln in lines)
    assert list(map(float, lines[99].split(","))) == [359.0, 0.84, 1.2, 0.24]
    assert list(map(int, lines[860].strip(',').split(','))) == [1315, 1316, 1326, 1327]


def test_oracle_errors():
    errs = json.load(open(os.path.join(cases.GOLDEN_DIR, "errors.json")))
    assert errs["remove_unconnected_all_background"] == "ValueError"
    with pytest.raises(ValueError):
        orc.remove_unconnected(np.zeros((3, 3, 3), dtype=bool))
    assert errs["vol2ugrid_bad_strategy"] == "ValueError"
    with pytest.raises(ValueError):
        orc.vol2ugrid(np.ones((2, 2, 2), dtype=bool), locking_strategy="both")
    assert errs["mesh2voxelfe_all_background"] == "ValueError"
    with pytest.raises(ValueError):
        orc.mesh2voxelfe(orc.vol2ugrid(np.zeros((3, 3, 3), dtype=bool)), cases.TEMPLATE, os.devnull)
    assert orc.matpropdictionary(None) is None and orc.matpropdictionary([]) is None
    assert orc.matpropdictionary(["a.inp", "1", "100", "b.inp", "101", "255"]) == errs["matpropdictionary_triples"]
    with pytest.raises(IOError):
        orc.matpropdictionary(["a.inp", "1"])


# --------------------------------------------------------------------------------------------
# vol2h5ParOSol / segment / add_cap (tests/golden/make_golden_parosol.py)

PAROSOL_CASES = {c[0]: c for c in cases.parosol_cases()}


@pytest.mark.parametrize("name", list(PAROSOL_CASES))
def test_oracle_parosol(golden_parosol, name):
    _, vol, kw = PAROSOL_CASES[name]
    cases.check_parosol_against_golden(golden_parosol, name, lambda: orc.vol2h5ParOSol_datasets(vol.copy(), **kw))


def test_oracle_segment_and_caps(golden_parosol):
    g = golden_parosol
    for name, img, t in cases.segment_cases():
        bw, T = orc.segment(img, t)
        assert T == int(g["seg/" + name + "/T"]) and bw.dtype == np.bool_
        assert np.array_equal(np.packbits(bw.ravel()), g["seg/" + name + "/bw"]), name
    i = 0
    while "cap/%d/in" % i in g.files:
        th, val = (int(v) for v in g["cap/%d/args" % i])
        res = orc.add_cap(g["cap/%d/in" % i], th, val)
        want = g["cap/%d/out" % i]
        assert res.dtype == want.dtype and np.array_equal(res, want)
        i += 1
    assert i == 3


def test_oracle_otsu_separates_two_populations():
    # no golden vector exists for Otsu (scikit-image is absent): sanity of the restated algorithm only
    rng = np.random.default_rng(3)
    img = np.concatenate([rng.normal(60, 8, 5000), rng.normal(180, 10, 3000)]).clip(0, 255).astype(np.uint8)
    T = orc.threshold_otsu_u8(img.reshape(20, 20, 20))
    assert 90 <= T <= 150
    assert orc.threshold_otsu_u8(np.full((3, 3, 3), 7, dtype=np.uint8)) == 7


@pytest.mark.parametrize("shape,sigma", [((20, 30, 40), 1.0), ((7, 9, 5), 2.5), ((33, 17, 64), 0.5),
                                         ((12, 40, 33), [0.0, 1.3, 2.0]), ((9, 9, 9), 3.7)])
def test_oracle_gaussian_is_scipy_bit_for_bit(shape, sigma):
    from scipy import ndimage
    img = np.random.default_rng(4).integers(0, 256, shape).astype(np.uint8)
    want = ndimage.gaussian_filter(img.astype(float), sigma, mode='nearest', truncate=4.0)
    got = orc.gaussian(img, sigma)
    assert got.dtype == np.float64 and np.array_equal(got, want)


def _check_embed(g, name, img, kw, fn):
    import warnings
    key = "emb/" + name
    arg = img.copy()
    if key + "/error" in g.files:
        with pytest.raises(BaseException) as ei:
            fn(arg, **kw)
        assert type(ei.value).__name__ == str(g[key + "/error"]), name
        return
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        res, bw = fn(arg, **kw)
    res, bw = np.asarray(res), np.asarray(bw)
    want = g[key + "/out"]
    assert res.dtype == want.dtype and np.array_equal(res, want), name
    assert bw.dtype == np.bool_ and np.array_equal(np.packbits(bw.ravel()), g[key + "/bw"]), name
    assert hashlib.sha256(arg.tobytes()).hexdigest() == str(g[key + "/arg_after"]), name   # in place or not


@pytest.mark.parametrize("case", cases.embed_cases(), ids=lambda c: c[0])
def test_oracle_embed(golden_parosol, case):
    name, img, kw = case
    _check_embed(golden_parosol, name, img, kw, orc.embed)


def _check_periosteum(g, name, bw, kw, fn):
    key = "per/" + name
    if key + "/error" in g.files:
        with pytest.raises(BaseException) as ei:
            fn(bw.copy(), **kw)
        assert type(ei.value).__name__ == str(g[key + "/error"]), name
        return
    res = np.asarray(fn(bw.copy(), **kw))
    assert res.dtype == np.bool_ and res.shape == bw.shape
    assert np.array_equal(np.packbits(res.ravel()), g[key + "/out"]), name


@pytest.mark.parametrize("case", cases.periosteum_cases(), ids=lambda c: c[0])
def test_oracle_periosteummask(golden_parosol, case):
    name, bw, kw = case
    _check_periosteum(golden_parosol, name, bw, kw, orc.periosteummask)
