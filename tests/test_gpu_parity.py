"""GPU parity tests (-m gpu): the CUDA path through the C ABI against the golden vectors produced
by the unmodified reference and against the CPU oracle on seeded inputs.  Bit-exact everywhere
(integer / byte work); the *NODE coordinates are compared as bytes, which is stronger than the
"parse to identical float64" bar."""
import hashlib
import json
import os

import numpy as np
import pytest

from tests import cases

pytestmark = pytest.mark.gpu

DECK_CASES = {c["name"]: c for c in cases.deck_cases()}


@pytest.fixture(scope="module")
def cfe():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import ciclope_b200
    from ciclope_b200 import _lib
    _lib.require_cuda()
    return ciclope_b200


@pytest.fixture(scope="module")
def orc():
    from oracle import oracle
    return oracle


def _unpack(g, name):
    shape = tuple(g[name + "/shape"])
    n = int(np.prod(shape))
    return (np.unpackbits(g[name + "/in"])[:n].astype(bool).reshape(shape),
            np.unpackbits(g[name + "/out"])[:n].astype(bool).reshape(shape))


# ------------------------------------------------------------------ remove_unconnected

@pytest.mark.parametrize("name", [n for n, _ in cases.ccl_cases()])
def test_remove_unconnected_golden(cfe, golden_ccl, name):
    bw, want = _unpack(golden_ccl, name)
    got = cfe.remove_unconnected(bw)
    assert isinstance(got, np.ndarray) and got.dtype == np.bool_ and got.shape == bw.shape
    assert np.array_equal(got, want)


def test_remove_unconnected_reference_known_answer(cfe):
    bw = cases.trab() > 100          # test/test_ciclope.py:11-15
    assert int(bw.sum()) - int(cfe.remove_unconnected(bw).sum()) == 4


def test_remove_unconnected_torch_in_torch_out(cfe, orc):
    import torch
    bw = cases.blobs((16, 20, 64), 0.25, 3)
    got = cfe.remove_unconnected(torch.from_numpy(bw).cuda())
    assert got.is_cuda and got.dtype == torch.bool
    assert np.array_equal(got.cpu().numpy(), orc.remove_unconnected(bw))


def test_remove_unconnected_empty_raises(cfe):
    with pytest.raises(ValueError):
        cfe.remove_unconnected(np.zeros((3, 4, 5), dtype=bool))


@pytest.mark.parametrize("value", [1, 7, 255])
def test_remove_unconnected_uint8_single_value(cfe, golden_ccl, value):
    # uint8 images with ONE non-zero value are binary images to measure.label as well
    for name in [n for n, _ in cases.ccl_cases()][:6]:
        bw, want = _unpack(golden_ccl, name)
        got = cfe.remove_unconnected(bw.astype(np.uint8) * np.uint8(value))
        assert got.dtype == np.bool_ and np.array_equal(got, want)


def test_remove_unconnected_multivalued_uint8_raises(cfe):
    """measure.label labels regions of EQUAL value (the reference returns the largest such region,
    tests/golden/errors.json: no error); this build refuses the image instead of reading it as != 0."""
    errs = json.load(open(os.path.join(cases.GOLDEN_DIR, "errors.json")))
    assert errs["remove_unconnected_multivalued_uint8"] is None
    multi = (cases.rand_binary((6, 8, 10), 0.6, 9) * np.arange(1, 11, dtype=np.uint8) % 4).astype(np.uint8)
    with pytest.raises(ValueError):
        cfe.remove_unconnected(multi)
    with pytest.raises(ValueError):
        cfe.remove_largest(multi)
    import torch
    big = torch.zeros((3, 40, 100), dtype=torch.uint8, device="cuda")
    big[1, 5, 7] = 200
    big[2, 39, 99] = 3           # in the scalar tail / another 16-byte chunk
    with pytest.raises(ValueError):
        cfe.remove_unconnected(big)


def test_packed_mask_tag_is_used_and_invalidated(cfe, orc, tmp_path):
    """remove_unconnected of a CUDA tensor leaves the kept voxels bit-packed on the result (vol2ugrid starts
    from them); an in-place edit of the mask must invalidate that."""
    import torch
    from ciclope_b200._util import packed_of
    bw = cases.blobs((12, 20, 64), 0.3, 31)              # X % 32 == 0: the packed path
    L = cfe.remove_unconnected(torch.from_numpy(bw).cuda())
    assert packed_of(L) is not None
    decks = []
    for vol in (L, L.clone()):                           # tagged / untagged: same deck
        out = str(tmp_path / "d.inp")
        cfe.mesh2voxelfe(cfe.vol2ugrid(vol, [0.0195] * 3), cases.TEMPLATE, out)
        decks.append(cases.mask_timestamp(open(out, "rb").read()))
    assert decks[0] == decks[1]
    L[0, 0, :] = True                                    # in-place edit: the tag is stale now
    assert packed_of(L) is None
    host = L.cpu().numpy()
    out = str(tmp_path / "e.inp")
    cfe.mesh2voxelfe(cfe.vol2ugrid(L, [0.0195] * 3), cases.TEMPLATE, out)
    orc.mesh2voxelfe(orc.vol2ugrid(host, [0.0195] * 3), cases.TEMPLATE, str(tmp_path / "o.inp"))
    assert cases.mask_timestamp(open(out, "rb").read()) == cases.mask_timestamp(open(str(tmp_path / "o.inp"), "rb").read())
    # GVmin != 0 must not use the tag either (thr = 0 is what the bits mean)
    L2 = cfe.remove_unconnected(torch.from_numpy(bw).cuda())
    assert cfe.vol2ugrid(L2, [1, 1, 1], GVmin=1).n_el == 0


@pytest.mark.parametrize("shape", [(9, 33, 70), (40, 48, 96), (3, 700, 1030)])
def test_node_emitters_agree(cfe, shape):
    """the list-free *NODE emitter (cfe_emit_nodes_bits: every CTA expands the used-node bitmap itself) writes the
    bytes of the list-driven one (cfe_fill_list + cfe_emit_nodes), including sections whose length is not a
    multiple of four lines and chunks that span empty stretches of the grid"""
    import torch
    from ciclope_b200 import _lib
    bw = cases.blobs(shape, 0.25, 17)
    bw[:, shape[1] // 3: shape[1] // 3 + 6, :] = False        # empty rows: chunks that span many empty words
    m = cfe.vol2ugrid(bw, [0.0195, 0.02, 0.03])
    n = 53 * m.n_nd
    a = torch.zeros(n + 64, dtype=torch.uint8, device="cuda")
    b = torch.zeros(n + 64, dtype=torch.uint8, device="cuda")
    assert a.data_ptr() % 16 == 0 and b.data_ptr() % 16 == 0
    p_xt = m.ftab.data_ptr()
    p_yt = p_xt + 16 * (m.X + 1)
    p_zt = p_yt + 16 * (m.Y + 1)
    s = _lib.stream()
    _lib.call("cfe_emit_nodes", m.node_list.data_ptr(), m.n_nd, p_xt, p_yt, p_zt, m.Y, m.X, 0, a.data_ptr(), s)
    _lib.call("cfe_emit_nodes_bits", m.nbits.data_ptr(), m.npref.data_ptr(), m.chunk_word.data_ptr(),
              m.n_node_words, m.n_nd, p_xt, p_yt, p_zt, m.Y, m.X, 0, b.data_ptr(), s)
    assert torch.equal(a, b) and int(a[n:].sum()) == 0
    assert bytes(a[:53].cpu().numpy()).endswith(b"\n")


def test_mesh_set_before_read_routes_to_generic(cfe, golden):
    """mesh.points = ... before the lazy array was ever read must count as an edit (not AttributeError)."""
    c = DECK_CASES["bool_5x7x9"]
    mesh = cfe.vol2ugrid(c["vol"], **c["u"])
    mesh.points = golden["bool_5x7x9/points"] * 2.0
    assert mesh._host_edited()
    mesh2 = cfe.vol2ugrid(c["vol"], **c["u"])
    mesh2.cells = [("hexahedron", golden["bool_5x7x9/cells"])]
    assert mesh2._host_edited()


@pytest.mark.parametrize("shape,density,seed", [
    ((32, 32, 32), 0.2, 1), ((32, 32, 32), 0.32, 2), ((17, 23, 129), 0.28, 3), ((40, 33, 96), 0.5, 4),
    ((64, 64, 64), 0.31, 5), ((5, 200, 70), 0.3, 6), ((64, 48, 160), 0.26, 7),
    # > 8192 runs per (8 slices x 32 rows) tile: the block-local union-find hands the tile to the global kernel
    ((8, 32, 512), 0.5, 8), ((16, 64, 300), 0.45, 9), ((9, 33, 1024), 0.33, 10)])
def test_remove_unconnected_random_vs_oracle(cfe, orc, shape, density, seed):
    # random noise near the 6-connectivity percolation threshold: thousands of components, many ties
    bw = cases.rand_binary(shape, density, 1000 + seed)
    assert np.array_equal(cfe.remove_unconnected(bw), orc.remove_unconnected(bw))


@pytest.mark.parametrize("shape,density,seed", [((48, 64, 96), 0.3, 11), ((96, 96, 128), 0.12, 12),
                                                ((128, 128, 128), 0.3, 13)])
def test_remove_unconnected_blobs_vs_oracle(cfe, orc, shape, density, seed):
    bw = cases.blobs(shape, density, seed, radius=2, passes=2)
    assert np.array_equal(cfe.remove_unconnected(bw), orc.remove_unconnected(bw))


# ------------------------------------------------------------------ vol2ugrid

@pytest.mark.parametrize("name", list(DECK_CASES))
def test_vol2ugrid_golden(cfe, golden, name):
    c = DECK_CASES[name]
    g = golden
    mesh, refn = cfe.vol2ugrid(c["vol"], refnodes=True, **c["u"])
    assert mesh.cells[0].type == "hexahedron" and mesh.cells[0][0] == "hexahedron"
    assert mesh.cells[0].data.dtype == np.int64
    assert np.array_equal(mesh.cells[0].data.reshape(-1, 8), g[name + "/cells"])
    assert np.array_equal(mesh.cells[0][1].reshape(-1, 8), g[name + "/cells"])
    assert np.array_equal(mesh.cells_dict["hexahedron"].reshape(-1, 8), g[name + "/cells"])
    assert mesh.points.dtype == g[name + "/points"].dtype
    assert mesh.points.shape == g[name + "/points"].shape
    assert np.array_equal(mesh.points, g[name + "/points"])
    assert list(mesh.cell_data) == ["GV"]
    assert mesh.cell_data["GV"][0].dtype == g[name + "/gv"].dtype
    assert np.array_equal(mesh.cell_data["GV"][0], g[name + "/gv"])
    assert list(mesh.point_sets) == ["NODES_X0", "NODES_X1", "NODES_Y0", "NODES_Y1", "NODES_Z0", "NODES_Z1",
                                     "NODES_X0Y0Z0", "NODES_X0Y0Z1"]
    assert list(mesh.cell_sets) == ["CELLS_X0", "CELLS_X1", "CELLS_Y0", "CELLS_Y1", "CELLS_Z0", "CELLS_Z1"]
    for k, v in mesh.point_sets.items():
        assert isinstance(v, list)
        assert v == g[name + "/ps/" + k].tolist(), k
    for k, v in mesh.cell_sets.items():
        assert isinstance(v, list)
        assert v == g[name + "/cs/" + k].tolist(), k
    for k, v in refn.items():
        assert np.array_equal(v, g[name + "/refnode/" + k]), k


@pytest.mark.parametrize("shape", [(5, 6, 64), (4, 7, 45)])
def test_vol2ugrid_gvmin_sweep(cfe, orc, shape):
    # every byte-compare corner of the SWAR threshold test (flat 16-byte path and ragged rows)
    rng = np.random.RandomState(3)
    vol = rng.randint(0, 256, size=shape).astype(np.uint8)
    vol[0, 0, :8] = [0, 1, 127, 128, 129, 254, 255, 200]
    for gvmin in (-1, 0, 1, 126, 127, 128, 199.5, 200, 253, 254, 255, 300):
        got = cfe.vol2ugrid(vol, GVmin=gvmin)
        want = orc.vol2ugrid(vol, GVmin=gvmin)
        assert np.array_equal(got.cells[0].data.reshape(-1, 8), want.cells[0].data.reshape(-1, 8)), gvmin
        assert np.array_equal(got.cell_data["GV"][0], want.cell_data["GV"][0]), gvmin


def test_vol2ugrid_reference_known_answers(cfe):
    # test/test_ciclope.py:75-87
    mesh = cfe.vol2ugrid(cases.trab(), [0.12, 0.12, 0.12], GVmin=100, locking_strategy="exact",
                         node_level_lock_num=1)
    assert mesh.cells[0].type == "hexahedron"
    assert mesh.cells[0].data.shape == (209, 8)
    assert mesh.points.shape == (1331, 3)
    assert mesh.point_sets['NODES_X0Y0Z0'] == [18, 19]
    assert (mesh.points[100] == np.array([0.12, 1.08, 0])).all()


# ------------------------------------------------------------------ mesh2voxelfe

def _deck(cfe, c, tmp_path, **over):
    m = dict(c["m"])
    if "matprop" in m:
        m["matprop"] = m["matprop"]()
    m.update(over)
    mesh = cfe.vol2ugrid(c["vol"], **c["u"])
    out = str(tmp_path / "deck.inp")
    assert cfe.mesh2voxelfe(mesh, fileout=out, **m) is None
    return mesh, cases.mask_timestamp(open(out, "rb").read())


@pytest.mark.parametrize("name", list(DECK_CASES))
def test_mesh2voxelfe_golden(cfe, golden, name, tmp_path):
    _, got = _deck(cfe, DECK_CASES[name], tmp_path)
    want = bytes(golden[name + "/deck"])
    if got != want:
        gl, wl = got.split(b"\n"), want.split(b"\n")
        for i, (a, b) in enumerate(zip(gl, wl)):
            assert a == b, "first differing line %d: got %r want %r" % (i, a, b)
        assert len(gl) == len(wl)
    assert got == want


def test_mesh2voxelfe_reference_known_answers(cfe, tmp_path):
    # test/test_ciclope.py:89-132
    mesh = cfe.vol2ugrid(cases.trab(), [0.12, 0.12, 0.12], GVmin=100)
    out = str(tmp_path / "foo.inp")
    cfe.mesh2voxelfe(mesh, cases.TEMPLATE, out)
    lines = open(out).read().split("\n")
    assert any('*NODE' in ln for ln in lines)
    assert any('*ELEMENT, TYPE=C3D8, ELSET=SET255' in ln for ln in lines)
    assert any('*ELSET, ELSET=CELLS_Z1' in ln for ln in lines)
    assert any('*NSET, NSET=NODES_Z0' in ln for ln in lines)
    assert any('*STEP' in ln for ln in lines)
    assert any('NODES_T, 3, 3, -0.002' in ln for ln in lines)
    assert list(map(float, lines[99].split(","))) == [359.0, 0.84, 1.2, 0.24]
    assert list(map(int, lines[860].strip(',').split(','))) == [1315, 1316, 1326, 1327]


@pytest.mark.parametrize("case", cases.midsize_cases(), ids=lambda c: c["name"])
def test_mesh2voxelfe_midsize_hash(cfe, case, tmp_path):
    want = json.load(open(os.path.join(cases.GOLDEN_DIR, "midsize.json")))[case["name"]]
    mesh, got = _deck(cfe, case, tmp_path)
    assert mesh.n_el == want["n_el"] and mesh.n_nd == want["n_used_nodes"]
    assert len(got) == want["size"]
    assert hashlib.sha256(got).hexdigest() == want["sha256"]


def _oracle_deck(orc, vol, u, m, path):
    m = dict(m)
    if "matprop" in m:
        m["matprop"] = m["matprop"]()
    orc.mesh2voxelfe(orc.vol2ugrid(vol, **u), fileout=path, **m)
    return cases.mask_timestamp(open(path, "rb").read())


@pytest.mark.parametrize("shape,seed,grey", [((33, 47, 65), 31, False), ((64, 64, 64), 32, False),
                                             ((30, 70, 130), 33, True), ((96, 96, 96), 34, True),
                                             ((128, 128, 128), 35, False)])
def test_full_path_vs_oracle(cfe, orc, shape, seed, grey, tmp_path):
    """remove_unconnected -> vol2ugrid -> mesh2voxelfe on seeded trabecular-like volumes."""
    bw = cases.blobs(shape, 0.3, seed)
    L = cfe.remove_unconnected(bw)
    L_ref = orc.remove_unconnected(bw)
    assert np.array_equal(L, L_ref)
    vol = cases.grey_from_mask(L, seed + 100) if grey else L
    u = dict(voxelsize=[0.0195, 0.0195, 0.0195])
    m = dict(templatefile=cases.TEMPLATE)
    if grey:
        m.update(keywords=['NSET', 'ELSET', 'PROPERTY'],
                 matprop=lambda: {"file": [cases.MAT_LINEAR], "range": [[1, 255]]})
    want = _oracle_deck(orc, vol, u, m, str(tmp_path / "cpu.inp"))
    c = dict(vol=vol, u=u, m=m)
    _, got = _deck(cfe, c, tmp_path)
    assert len(got) == len(want)
    assert hashlib.sha256(got).hexdigest() == hashlib.sha256(want).hexdigest()
    assert got == want


def test_device_tensor_input_and_host_sink(cfe, orc, tmp_path):
    import torch
    L = cases.blobs((32, 32, 64), 0.3, 41)
    vol = torch.from_numpy(L).cuda()
    mesh = cfe.vol2ugrid(vol, [0.0195] * 3)
    sink = torch.empty(64 << 20, dtype=torch.uint8).pin_memory()
    n = cfe.mesh2voxelfe(mesh, cases.TEMPLATE, sink)
    got = cases.mask_timestamp(sink[:n].numpy().tobytes())
    want = _oracle_deck(orc, L, dict(voxelsize=[0.0195] * 3), dict(templatefile=cases.TEMPLATE),
                        str(tmp_path / "cpu.inp"))
    assert got == want


def test_notebook_cell_data_astype_pattern(cfe, golden, tmp_path):
    # notebooks do mesh.cell_data['GV'] = mesh.cell_data['GV'][0].astype('uint8') before mesh2voxelfe
    c = DECK_CASES["trab_bool_property"]
    mesh = cfe.vol2ugrid(c["vol"], **c["u"])
    mesh.cell_data['GV'] = mesh.cell_data['GV'][0].astype('uint8')
    out = str(tmp_path / "d.inp")
    cfe.mesh2voxelfe(mesh, cases.TEMPLATE, out, keywords=['NSET', 'ELSET', 'PROPERTY'],
                     matprop={"file": [cases.MAT_LINEAR]})
    text = open(out).read()
    assert "ELSET=SET1, MATERIAL=MATSET1" in text and "SETTrue" not in text


# ------------------------------------------------------------------ remove_largest / fill_voids

@pytest.mark.parametrize("name", [n for n, _ in cases.ccl_cases()])
def test_remove_largest_golden(cfe, golden_reuse, name):
    bw = dict(cases.ccl_cases())[name]
    want = np.unpackbits(golden_reuse["rl/" + name])[:bw.size].astype(bool).reshape(bw.shape)
    got = cfe.remove_largest(bw)
    assert got.dtype == np.bool_ and got.shape == bw.shape
    assert np.array_equal(got, want)


@pytest.mark.parametrize("name", [c[0] for c in cases.fill_cases()])
def test_fill_voids_golden(cfe, golden_reuse, name):
    _, img, kw = [c for c in cases.fill_cases() if c[0] == name][0]
    want = golden_reuse["fv/" + name + "/out"]
    arg = img.copy()
    got = cfe.fill_voids(arg, **kw)
    assert got.dtype == want.dtype and np.array_equal(got, want)
    if kw.get("makecopy"):
        assert got is not arg and np.array_equal(arg, img)
    else:
        assert got is arg                      # the reference fills in place and returns its argument
    # CUDA tensor in, CUDA tensor out (in place unless makecopy)
    import torch
    t = torch.from_numpy(img.copy()).cuda()
    r = cfe.fill_voids(t, **kw)
    assert r.is_cuda and np.array_equal(r.cpu().numpy(), want)
    assert (r.data_ptr() == t.data_ptr()) != bool(kw.get("makecopy"))
    assert np.array_equal(t.cpu().numpy(), img if kw.get("makecopy") else want)


@pytest.mark.parametrize("shape,density,seed", [((40, 48, 70), 0.7, 1), ((64, 64, 64), 0.8, 2), ((33, 65, 130), 0.75, 3)])
def test_reuse_random_vs_oracle(cfe, orc, shape, density, seed):
    g = cases.rand_grey(shape, density, seed)
    assert np.array_equal(cfe.fill_voids(g.copy()), orc.fill_voids(g.copy()))
    assert np.array_equal(cfe.fill_voids(g.copy(), fill_val=3, makecopy=True), orc.fill_voids(g.copy(), 3, True))
    bw = cases.blobs(shape, 0.3, seed)
    assert np.array_equal(cfe.remove_largest(bw), orc.remove_largest(bw))
    # remove_unconnected and remove_largest partition the foreground
    assert np.array_equal(cfe.remove_largest(bw) | cfe.remove_unconnected(bw), bw)


def test_reuse_errors(cfe, golden_reuse):
    assert str(golden_reuse["err/remove_largest_all_background"]) == "ValueError"
    with pytest.raises(ValueError):
        cfe.remove_largest(np.zeros((3, 3, 3), dtype=bool))
    assert str(golden_reuse["err/fill_voids_all_foreground"]) == "ValueError"
    with pytest.raises(ValueError):
        cfe.fill_voids(np.full((3, 3, 3), 9, dtype=np.uint8))
    with pytest.raises(ValueError):
        cfe.fill_voids(np.zeros((3, 3, 3), dtype=np.uint8), fill_val=300)


# ------------------------------------------------------------------ generic / edited meshes

@pytest.mark.parametrize("case", cases.generic_cases(), ids=lambda c: c["name"])
def test_mesh2voxelfe_generic_mesh_golden(cfe, golden_generic, case, tmp_path):
    # any hexahedral meshio mesh (voxelFE.py:612-688): array-driven CUDA emitters, byte-exact
    m = dict(case["m"])
    if "matprop" in m:
        m["matprop"] = m["matprop"]()
    out = str(tmp_path / "deck.inp")
    cfe.mesh2voxelfe(cases.PlainMesh(case), fileout=out, **m)
    assert cases.mask_timestamp(open(out, "rb").read()) == bytes(golden_generic[case["name"] + "/deck"])


@pytest.mark.parametrize("case", cases.edited_cases(), ids=lambda c: c["name"])
def test_mesh2voxelfe_edited_voxelmesh_golden(cfe, golden_generic, case, tmp_path):
    # a vol2ugrid mesh whose host-side sets / grey values were changed is emitted from the host arrays
    m = dict(case["m"])
    if "matprop" in m:
        m["matprop"] = m["matprop"]()
    mesh = case["edit"](cfe.vol2ugrid(case["vol"], **case["u"]))
    assert mesh._host_edited()
    out = str(tmp_path / "deck.inp")
    cfe.mesh2voxelfe(mesh, fileout=out, **m)
    assert cases.mask_timestamp(open(out, "rb").read()) == bytes(golden_generic[case["name"] + "/deck"])


def test_untouched_views_keep_the_fast_path(cfe, golden, tmp_path):
    # reading every host view (as the reference's tests do) must not divert to the generic emitters,
    # in-place edits of a set list must; the big arrays are read-only so they can only be replaced
    c = DECK_CASES["grey_6x7x20"]
    mesh = cfe.vol2ugrid(c["vol"], **c["u"])
    _ = mesh.points[3], mesh.cells[0].data[0], mesh.cells_dict["hexahedron"], mesh.cell_data["GV"][0]
    _ = mesh.point_sets["NODES_Z0"], mesh.cell_sets["CELLS_Z1"]
    assert not mesh._host_edited()
    with pytest.raises(ValueError):
        mesh.points[0, 0] = 1.0
    with pytest.raises(ValueError):
        mesh.cells[0].data[0, 0] = 1
    out = str(tmp_path / "a.inp")
    cfe.mesh2voxelfe(mesh, cases.TEMPLATE, out)
    assert cases.mask_timestamp(open(out, "rb").read()) == bytes(golden["grey_6x7x20/deck"])
    mesh.point_sets["NODES_Z0"].append(5)
    assert mesh._host_edited()
    mesh.point_sets["NODES_Z0"].pop()
    assert not mesh._host_edited()
    mesh.points = mesh.points * 2.0
    assert mesh._host_edited()
    cfe.mesh2voxelfe(mesh, cases.TEMPLATE, out)
    lines = open(out).read().split("\n")
    want = "{0:10d}, {1:12.6f}, {2:12.6f}, {3:12.6f}".format(int(lines[6].split(",")[0]),
                                                             *mesh.points[int(lines[6].split(",")[0])])
    assert lines[6] == want


def test_generic_mesh_midsize_vs_oracle(cfe, orc, tmp_path):
    # 40x48x70 blobs, cells scrambled, random node coordinates, int32 grey values up to 4000
    rng = np.random.RandomState(5)
    vol = cases.grey_from_mask(cases.blobs((40, 48, 70), 0.3, 21), 22)
    pts, cells, gv = cases._grid_mesh(vol)
    perm = rng.permutation(len(cells))
    case = dict(points=rng.uniform(-2e5, 2e5, size=pts.shape), blocks=[("hexahedron", cells[perm])],
                cell_data={"GV": [(gv[perm].astype(np.int32) * 16)]},
                point_sets={"A": rng.randint(0, len(pts), 12345), "B": []},
                cell_sets={"C": rng.randint(1, len(cells) + 1, 30001)})
    a, b = str(tmp_path / "a.inp"), str(tmp_path / "b.inp")
    cfe.mesh2voxelfe(cases.PlainMesh(case), cases.TEMPLATE, a)
    orc.mesh2voxelfe(cases.PlainMesh(case), cases.TEMPLATE, b)
    assert cases.mask_timestamp(open(a, "rb").read()) == cases.mask_timestamp(open(b, "rb").read())


# ------------------------------------------------------------------ error behaviour

def test_error_behaviour_matches_reference(cfe, tmp_path):
    errs = json.load(open(os.path.join(cases.GOLDEN_DIR, "errors.json")))
    empty = np.zeros((3, 3, 3), dtype=bool)
    assert errs["vol2ugrid_all_background"] is None
    mesh = cfe.vol2ugrid(empty)
    assert mesh.n_el == 0 and mesh.cells[0].data.shape[0] == 0 and mesh.points.shape == (64, 3)
    assert errs["mesh2voxelfe_all_background"] == "ValueError"
    with pytest.raises(ValueError):
        cfe.mesh2voxelfe(mesh, cases.TEMPLATE, str(tmp_path / "e.inp"))
    g = cases.rand_grey((4, 4, 6), 0.5, 3)
    mk = lambda: cfe.vol2ugrid(g)  # noqa: E731
    for key, mp in [("matprop_two_files_no_range", {"file": [cases.MAT_CONST, cases.MAT_POWER]}),
                    ("matprop_range_exceeds", {"file": [cases.MAT_CONST], "range": [[1, 300]]}),
                    ("matprop_range_negative", {"file": [cases.MAT_CONST], "range": [[-1, 30]]}),
                    ("matprop_overlap", {"file": [cases.MAT_CONST, cases.MAT_POWER],
                                         "range": [[1, 100], [100, 255]]})]:
        assert errs[key] == "OSError"
        with pytest.raises(IOError):
            cfe.mesh2voxelfe(mk(), cases.TEMPLATE, str(tmp_path / "e.inp"), matprop=mp)
    with pytest.raises(TypeError):
        cfe.vol2ugrid(np.zeros((2, 2, 2), dtype=np.float32))
    with pytest.raises(AttributeError):     # not a mesh: the reference fails on mesh.cell_data too
        cfe.mesh2voxelfe(object(), cases.TEMPLATE, str(tmp_path / "e.inp"))


# ------------------------------------------------------------------ coordinate formatting kernel

def test_format_coords_matches_cpython(cfe):
    """cfe_format_coords must print exactly what '{:12.6f}'.format does (voxelFE.py:630)."""
    import torch
    from ciclope_b200 import _lib
    rng = np.random.default_rng(0)
    vals = np.concatenate([
        rng.uniform(-9999.0, 99999.0, 20000),
        rng.uniform(-1e-5, 1e-5, 5000),
        np.arange(-640, 6400) / 128.0,                      # exact ties of the 6th decimal: half-to-even
        np.arange(1, 4000) / 128.0 + 1000.0,
        np.arange(0, 3000) * 0.0195, np.arange(0, 3000) * 0.0000005, np.arange(0, 3000) * np.float64(np.float32(0.1)),
        np.array([0.0, -0.0, 99999.9999994, 99999.999999, -9999.9999994, 5e-324, -5e-324, 1e-7, 4.9999999e-7,
                  5e-7, 5.0000001e-7, 1.5e-6, 2.5e-6, 0.1, 0.2, 0.3, 1 / 3, 2 / 3, 12345.678901, 12345.6789015]),
    ]).astype(np.float64)
    t = torch.from_numpy(vals).cuda()
    out = torch.zeros(16 * len(vals), dtype=torch.uint8, device="cuda")
    flag = torch.zeros(1, dtype=torch.int32, device="cuda")
    _lib.call("cfe_format_coords", t.data_ptr(), len(vals), out.data_ptr(), flag.data_ptr(), _lib.stream())
    got = out.cpu().numpy().reshape(-1, 16)
    assert int(flag.item()) == 0
    for i, v in enumerate(vals.tolist()):
        want = format(v, "12.6f")
        assert len(want) == 12
        assert bytes(got[i, :12]).decode() == want, (i, v, want, bytes(got[i, :12]))
    for bad in (100000.0, 99999.9999996, -10000.0, float("inf"), float("nan"), 1e300):
        flag.zero_()
        t = torch.tensor([bad], dtype=torch.float64, device="cuda")
        _lib.call("cfe_format_coords", t.data_ptr(), 1, out.data_ptr(), flag.data_ptr(), _lib.stream())
        assert int(flag.item()) == 1, bad
        assert len(format(bad, "12.6f")) != 12 or bad != bad or abs(bad) == float("inf")


def test_coordinates_beyond_1e18_raise(cfe, tmp_path):
    # fields up to 26 characters are formatted (golden cases vs_wide_*); beyond 1e18 the writer refuses
    mesh = cfe.vol2ugrid(np.ones((2, 2, 3), dtype=bool), voxelsize=[1e18, 1.0, 1.0])
    with pytest.raises(NotImplementedError):
        cfe.mesh2voxelfe(mesh, cases.TEMPLATE, str(tmp_path / "w.inp"))


def test_mesh2voxelfe_is_repeatable_on_one_mesh(cfe, golden, tmp_path):
    """the lazy mesh keeps device state between calls: emitting twice (and with another element type)
    must give the reference bytes each time"""
    c = DECK_CASES["blobs_bool"]
    mesh = cfe.vol2ugrid(c["vol"], **c["u"])
    outs = []
    for i, kw in enumerate([{}, {}, {"eltype": "C3D8R"}, {}]):
        p = str(tmp_path / ("d%d.inp" % i))
        cfe.mesh2voxelfe(mesh, cases.TEMPLATE, p, **kw)
        outs.append(cases.mask_timestamp(open(p, "rb").read()))
    want = bytes(golden["blobs_bool/deck"])
    assert outs[0] == want and outs[1] == want and outs[3] == want
    assert outs[2] == want.replace(b"TYPE=C3D8,", b"TYPE=C3D8R,")
    g = DECK_CASES["blobs_grey_property"]
    mesh = cfe.vol2ugrid(g["vol"], **g["u"])
    for i in range(2):
        p = str(tmp_path / ("g%d.inp" % i))
        m = dict(g["m"]); m["matprop"] = m["matprop"]()
        cfe.mesh2voxelfe(mesh, fileout=p, **m)
        assert cases.mask_timestamp(open(p, "rb").read()) == bytes(golden["blobs_grey_property/deck"])


# --------------------------------------------------------------------------------------------
# vol2h5ParOSol (voxelFE.py:285-472), segment (preprocess.py:20-46), add_cap (preprocess.py:210-231)

PAROSOL_CASES = {c[0]: c for c in cases.parosol_cases()}


@pytest.mark.parametrize("name", list(PAROSOL_CASES))
def test_parosol_datasets_vs_reference(cfe, golden_parosol, name):
    from ciclope_b200.parosol import parosol_datasets
    _, vol, kw = PAROSOL_CASES[name]
    cases.check_parosol_against_golden(golden_parosol, name, lambda: parosol_datasets(vol.copy(), **kw))


def test_parosol_file_through_h5py_calls(cfe, orc, tmp_path, monkeypatch):
    """vol2h5ParOSol makes the reference's h5py calls (File / create_group / create_dataset in the same
    order with the same dtypes); h5py itself is not installed, so a recording stand-in takes its place."""
    import sys
    import types
    rec = []

    class _Group:
        def create_dataset(self, name, data=None, dtype=None):
            rec.append((name, np.array(data, dtype=dtype)))

    class File:
        def __init__(self, path, mode):
            rec.append(("File", path, mode))

        def create_group(self, name):
            rec.append(("group", name))
            return _Group()

        def close(self):
            rec.append(("close",))

    monkeypatch.setitem(sys.modules, "h5py", types.SimpleNamespace(File=File))
    vol = cases.blobs((20, 24, 28), 0.3, 3)
    path = str(tmp_path / "x.h5")
    cfe.vol2h5ParOSol(vol, path, -0.04, voxelsize=0.0195)
    want = orc.vol2h5ParOSol_datasets(vol, -0.04, voxelsize=0.0195)
    assert rec[0] == ("File", path, "w") and rec[1] == ("group", "Image_Data") and rec[-1] == ("close",)
    got = rec[2:-1]
    assert [n for n, _ in got] == list(want)
    for (n, a), b in zip(got, want.values()):
        assert a.dtype == b.dtype and a.shape == b.shape and np.array_equal(a, b), n


def test_parosol_without_h5py_raises(cfe, tmp_path):
    try:
        import h5py  # noqa: F401
        pytest.skip("h5py is installed")
    except ImportError:
        pass
    with pytest.raises(ImportError):
        cfe.vol2h5ParOSol(cases.blobs((8, 8, 8), 0.4, 1), str(tmp_path / "x.h5"), 0.1)


@pytest.mark.parametrize("shape,seed", [((40, 64, 96), 1), ((33, 50, 70), 2), ((64, 128, 160), 3)])
def test_parosol_vs_oracle_larger(cfe, orc, shape, seed):
    """bigger volumes with an empty margin, both strategies, device input; sorted order = same rows"""
    import torch
    from ciclope_b200.parosol import parosol_datasets
    inner = cases.blobs(tuple(s - 6 for s in shape), 0.35, seed)
    vol = np.zeros(shape, dtype=bool)
    vol[2:-4, 1:-5, 3:-3] = inner
    for strat, kw in (("plane", dict(plane_lock_num=2)), ("exact", dict(node_level_lock_num=2))):
        want = orc.vol2h5ParOSol_datasets(vol, 0.25, voxelsize=[0.1, 0.2, 0.3], young_modulus=17000.5,
                                          locking_strategy=strat, **kw)
        got = parosol_datasets(torch.from_numpy(vol).cuda(), 0.25, voxelsize=[0.1, 0.2, 0.3],
                               young_modulus=17000.5, locking_strategy=strat, **kw)
        for k in want:
            a, b = np.asarray(got[k]), want[k]
            assert a.dtype == b.dtype and a.shape == b.shape and np.array_equal(a, b), (strat, k)
        srt = parosol_datasets(vol, 0.25, locking_strategy=strat, node_order="sorted", **kw)
        a = srt["Fixed_Displacement_Coordinates"]
        b = want["Fixed_Displacement_Coordinates"]
        assert a.shape == b.shape
        assert sorted(map(tuple, a.tolist())) == sorted(map(tuple, b.tolist()))


def test_segment_and_caps_vs_reference(cfe, golden_parosol):
    import torch
    g = golden_parosol
    for name, img, t in cases.segment_cases():
        bw, T = cfe.segment(img, t)
        assert T == int(g["seg/" + name + "/T"]) and bw.dtype == np.bool_ and bw.shape == img.shape
        assert np.array_equal(np.packbits(bw.ravel()), g["seg/" + name + "/bw"]), name
        bw_t, _ = cfe.segment(torch.from_numpy(img).cuda(), t)
        assert bw_t.dtype == torch.bool and np.array_equal(bw_t.cpu().numpy(), bw)
    i = 0
    while "cap/%d/in" % i in g.files:
        th, val = (int(v) for v in g["cap/%d/args" % i])
        res = cfe.add_cap(g["cap/%d/in" % i], th, val)
        want = g["cap/%d/out" % i]
        assert res.dtype == want.dtype and np.array_equal(res, want)
        i += 1
    with pytest.raises(ValueError):
        cfe.add_cap(g["cap/0/in"], 0, 1)
    with pytest.raises(TypeError):
        cfe.add_cap(g["cap/0/in"], 2, 0.5)


@pytest.mark.parametrize("shape", [(5, 7, 9), (16, 32, 64), (3, 5, 2050), (1, 1, 15), (2, 3, 16)])
def test_mask_image_is_the_reference_lines(cfe, shape):
    """pipeline.py:240-242: `masked = I; masked[~L] = 0` -- in place, like the boolean-index assignment"""
    import torch
    rng = np.random.default_rng(sum(shape))
    img = rng.integers(0, 256, shape).astype(np.uint8)
    L = rng.random(shape) < 0.4
    want = img.copy()
    want[~L] = 0
    a = img.copy()
    r = cfe.preprocess.mask_image(a, L)
    assert r is a and a.dtype == np.uint8 and np.array_equal(a, want)
    b = img.copy()
    r = cfe.preprocess.mask_image(b, L.astype(np.uint8), makecopy=True)
    assert r is not b and np.array_equal(b, img) and np.array_equal(r, want)
    t = torch.from_numpy(img).cuda()
    r = cfe.preprocess.mask_image(t, torch.from_numpy(L).cuda())
    assert r.data_ptr() == t.data_ptr() and np.array_equal(t.cpu().numpy(), want)
    t = torch.from_numpy(img).cuda()
    r = cfe.preprocess.mask_image(t, torch.from_numpy(L).cuda(), makecopy=True)
    assert r.data_ptr() != t.data_ptr() and np.array_equal(t.cpu().numpy(), img) and np.array_equal(r.cpu().numpy(), want)
    bw = img > 100
    r = cfe.preprocess.mask_image(bw.copy(), L)
    assert r.dtype == np.bool_ and np.array_equal(r, bw & L)
    with pytest.raises(IndexError):
        cfe.preprocess.mask_image(img.copy(), L[:, :, :-1] if shape[2] > 1 else np.zeros((1, 1, 2), bool))


def test_segment_otsu_vs_oracle(cfe, orc):
    rng = np.random.default_rng(11)
    for shape in [(20, 30, 40), (7, 33, 65)]:
        img = np.concatenate([rng.normal(70, 12, shape[0] * shape[1] * shape[2] // 2),
                              rng.normal(170, 15, shape[0] * shape[1] * shape[2] - shape[0] * shape[1] * shape[2] // 2)]
                             ).clip(0, 255).astype(np.uint8).reshape(shape)
        bw, T = cfe.segment(img, None)
        bw_o, T_o = orc.segment(img, None)
        assert T == T_o and np.array_equal(bw, bw_o)
    flat = np.full((4, 5, 6), 9, dtype=np.uint8)
    bw, T = cfe.segment(flat, None)
    assert T == 9 and not bw.any()


@pytest.mark.parametrize("shape,sigma", [((20, 30, 40), 1.0), ((7, 9, 5), 2.5), ((33, 17, 64), 0.5),
                                         ((12, 40, 33), [0.0, 1.3, 2.0]), ((64, 96, 130), 1.5), ((9, 9, 9), 3.7),
                                         ((3, 4, 2100), 2.0), ((2, 3, 1024), [0.0, 0.0, 6.0])])
def test_gaussian_bit_exact_vs_scipy_and_oracle(cfe, orc, shape, sigma):
    """pipeline.py:146-148; float64 work, compared BIT FOR BIT (tolerance 0) with the oracle and with the
    installed scipy.ndimage.gaussian_filter that skimage.filters.gaussian wraps"""
    import torch
    from scipy import ndimage
    img = np.random.default_rng(4).integers(0, 256, shape).astype(np.uint8)
    got = cfe.gaussian(img, sigma=sigma, preserve_range=True)
    assert isinstance(got, np.ndarray) and got.dtype == np.float64 and got.shape == img.shape
    assert np.array_equal(got, orc.gaussian(img, sigma))
    assert np.array_equal(got, ndimage.gaussian_filter(img.astype(float), sigma, mode='nearest', truncate=4.0))
    # float64 input, tensor in / tensor out, then the threshold of the smoothed image
    got2 = cfe.gaussian(torch.from_numpy(img.astype(np.float64)).cuda(), sigma=sigma, preserve_range=True)
    assert got2.dtype == torch.float64 and np.array_equal(got2.cpu().numpy(), got)
    bw, T = cfe.segment(got2, 100.9)
    assert T == 100 and bw.dtype == torch.bool and np.array_equal(bw.cpu().numpy(), got > 100)
    bw_np, _ = cfe.segment(got, 100)
    assert np.array_equal(bw_np, got > 100)
    with pytest.raises(NotImplementedError):
        cfe.segment(got, None)
    with pytest.raises(NotImplementedError):
        cfe.gaussian(img, sigma=1.0)


@pytest.mark.parametrize("case", cases.embed_cases(), ids=lambda c: c[0])
def test_embed_vs_reference(cfe, golden_parosol, case):
    """preprocess.py:233-355 against the outputs of the unmodified reference (image, mask, in-place rule)"""
    from tests.test_oracle_golden import _check_embed
    name, img, kw = case
    _check_embed(golden_parosol, name, img, kw, cfe.embed)


def test_embed_tensor_in_place(cfe, orc):
    import torch
    name, img, kw = [c for c in cases.embed_cases() if c[2]["embed_dir"] == "-z" and not c[2]["makecopy"]][0]
    want, bw_want = orc.embed(img.copy(), **kw)
    t = torch.from_numpy(img.copy()).cuda()
    res, bw = cfe.embed(t, **kw)
    assert res is t and np.array_equal(t.cpu().numpy(), want) and np.array_equal(bw.cpu().numpy(), bw_want)
    t2 = torch.from_numpy(img.copy()).cuda()
    res2, _ = cfe.embed(t2, **dict(kw, makecopy=True))
    assert res2 is not t2 and np.array_equal(t2.cpu().numpy(), img) and np.array_equal(res2.cpu().numpy(), want)


@pytest.mark.parametrize("case", cases.periosteum_cases(), ids=lambda c: c[0])
def test_periosteummask_vs_reference(cfe, golden_parosol, case):
    """preprocess.py:357-420: disk closing + hole filling per slice, then remove_unconnected"""
    from tests.test_oracle_golden import _check_periosteum
    name, bw, kw = case
    _check_periosteum(golden_parosol, name, bw, kw, cfe.periosteummask)


def test_periosteummask_larger_vs_oracle(cfe, orc):
    import torch
    bw = cases.blobs((24, 150, 200), 0.25, 9)
    bw[:, 60:90, 80:120] = False                          # a cavity that the closing cannot bridge
    for r in (2, 7, 31):
        want = orc.periosteummask(bw, closepixels=r)
        got = cfe.periosteummask(torch.from_numpy(bw).cuda(), closepixels=r)
        assert got.dtype == torch.bool and np.array_equal(got.cpu().numpy(), want), r
    for cw in (2, 5, 8, 33):                              # final 3D closing with cube(closevoxels)
        want = orc.periosteummask(bw, closepixels=2, closevoxels=cw)
        got = cfe.periosteummask(torch.from_numpy(bw).cuda(), closepixels=2, closevoxels=cw)
        assert np.array_equal(got.cpu().numpy(), want), cw
    with pytest.raises(NotImplementedError):
        cfe.periosteummask(bw, closevoxels=40)
    # preliminary removal of small objects (scikit-image's remove_small_objects on the 3D / 2D image)
    rng = np.random.default_rng(3)
    specks = bw | (rng.random(bw.shape) < 0.01)
    for min_size in (2, 30, 5000):
        want = orc.periosteummask(specks, closepixels=3, remove_objects_smaller_than=min_size)
        got = cfe.periosteummask(specks, closepixels=3, remove_objects_smaller_than=min_size)
        assert got.dtype == np.bool_ and np.array_equal(got, want), min_size
    sl = specks[5]
    assert np.array_equal(cfe.periosteummask(sl, closepixels=2, remove_objects_smaller_than=9, removeunconn=False),
                          orc.periosteummask(sl, closepixels=2, remove_objects_smaller_than=9, removeunconn=False))
    with pytest.raises(TypeError):
        cfe.periosteummask(specks.astype(np.uint8), remove_objects_smaller_than=5)
