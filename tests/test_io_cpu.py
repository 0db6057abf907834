"""Host logic of ciclope_b200.io: the file selection of read_tiff_stack (recon_utils.py:196-230) on a stack
written with Pillow, and the legacy VTK writer / reader pair."""
import os

import numpy as np
import pytest

from ciclope_b200 import io as cio


@pytest.fixture()
def stack(tmp_path):
    from PIL import Image
    rng = np.random.RandomState(4)
    vol = rng.randint(0, 256, (12, 9, 14), dtype=np.uint8)
    for k in range(vol.shape[0]):
        Image.fromarray(vol[k]).save(str(tmp_path / ("slice_%04d.tif" % k)))
    return vol, str(tmp_path / "slice_0000.tif")


def test_read_tiff_stack_host(stack):
    vol, first = stack
    got = cio.read_tiff_stack(first, device=None)
    assert got.dtype == np.uint8 and np.array_equal(got, vol)
    # the reference's range selection: files [index of 0003 : index of 0009)  (recon_utils.py:217-223)
    assert np.array_equal(cio.read_tiff_stack(first, range=[3, 9], device=None), vol[3:9])
    with pytest.warns(UserWarning):
        assert np.array_equal(cio.read_tiff_stack(first, range=[3, 99], device=None), vol)


def test_reference_fixture_stack():
    """the reference's own unit-test fixture (test/trab/slice_00..09.tif) is tests/golden/trab_10x10x10_u8.npy;
    a stack written from it reads back identically"""
    import tempfile
    from PIL import Image
    from tests import cases
    tr = cases.trab()
    with tempfile.TemporaryDirectory() as d:
        for k in range(tr.shape[0]):
            Image.fromarray(tr[k]).save(os.path.join(d, "slice_%02d.tif" % k))
        assert np.array_equal(cio.read_tiff_stack(os.path.join(d, "slice_00.tif"), device=None), tr)


def test_vtk_roundtrip(tmp_path):
    rng = np.random.RandomState(5)
    pts = rng.rand(30, 3)
    cells = rng.randint(0, 30, (7, 8)).astype(np.int64)
    gv = rng.randint(0, 256, 7).astype(np.uint8)
    path = str(tmp_path / "m.vtk")
    cio.write_vtk(path, pts, cells, {"GV": [gv]})
    p2, c2, cd = cio.read_vtk(path)
    assert np.array_equal(p2, pts) and np.array_equal(c2, cells) and np.array_equal(cd["GV"], gv)
    head = open(path, "rb").read(200)
    assert head.startswith(b"# vtk DataFile Version 4.2\n") and b"DATASET UNSTRUCTURED_GRID" in head


@pytest.mark.gpu
def test_read_tiff_stack_device_and_mesh_write(stack, tmp_path, golden):
    import torch
    import ciclope_b200 as cfe
    vol, first = stack
    t = cio.read_tiff_stack(first)
    assert t.is_cuda and t.dtype == torch.uint8 and np.array_equal(t.cpu().numpy(), vol)
    # straight into the hot path: segment -> remove_unconnected -> vol2ugrid -> mesh.write('.vtk')
    bw, _ = cfe.segment(t, 100)
    L = cfe.remove_unconnected(bw)
    mesh = cfe.vol2ugrid(L, [0.5, 0.5, 0.5])
    path = str(tmp_path / "mesh.vtk")
    mesh.write(path)
    pts, cells, cd = cio.read_vtk(path)
    assert np.array_equal(pts, mesh.points) and np.array_equal(cells, mesh.cells[0].data)
    assert np.array_equal(cd["GV"], mesh.cell_data["GV"][0].astype(np.uint8))
    # golden mesh of the reference: same arrays in the file
    from tests import cases
    c = {d["name"]: d for d in cases.deck_cases()}["trab_gvmin100_plane"]
    m2 = cfe.vol2ugrid(c["vol"], **c["u"])
    m2.write(path)
    pts, cells, cd = cio.read_vtk(path)
    assert np.array_equal(pts, golden["trab_gvmin100_plane/points"])
    assert np.array_equal(cells, golden["trab_gvmin100_plane/cells"])
    assert np.array_equal(cd["GV"], golden["trab_gvmin100_plane/gv"])
