"""GPU tests (-m gpu) on real data -- BASELINE.json configs[0]: the LHDL trabecular bone stack of the
reference's own checkout thresholded at 63 (README.md:184), golden record produced by the UNMODIFIED
reference (tests/golden/make_golden_real.py): kept voxels bit-exact, per-section SHA-256 of the binary deck
and of the grey-scale deck with 190-odd material sets; single GPU and 2 / 4 / 8 emulated z-slabs."""
import json
import os

import numpy as np
import pytest

from ciclope_b200 import sections
from tests import cases

pytestmark = pytest.mark.gpu
SECTIONS = ("nodes", "elements", "nset", "elset", "tail")


@pytest.fixture(scope="module")
def real():
    import torch
    assert torch.cuda.is_available()
    rec = json.load(open(os.path.join(cases.GOLDEN_DIR, "real_lhdl.json")))
    bw, kept, gv = cases.real_lhdl()
    return rec, bw, kept, gv


GREY_KW = dict(keywords=['NSET', 'ELSET', 'PROPERTY'])


def _grey_kw():
    return dict(GREY_KW, matprop={"file": [cases.MAT_LINEAR], "range": [[1, 255]]})


def test_real_remove_unconnected(real):
    from ciclope_b200 import remove_unconnected
    rec, bw, kept, _ = real
    L = remove_unconnected(bw)
    assert int(bw.sum()) - int(L.sum()) == rec["foreground_voxels"] - rec["kept_voxels"] == 1691
    assert np.array_equal(L, kept)


@pytest.mark.parametrize("kind", ["binary", "greyscale"])
def test_real_deck_sections(real, kind, tmp_path):
    from ciclope_b200 import mesh2voxelfe, vol2ugrid
    rec, _, kept, gv = real
    want = rec[kind]
    mesh = vol2ugrid(kept if kind == "binary" else gv, rec["voxelsize"])
    assert mesh.n_el == want["n_el"] and mesh.n_nd == want["n_used_nodes"] and mesh.n_points == want["n_points"]
    sizes = {k: len(v) for k, v in list(mesh.point_sets.items()) + list(mesh.cell_sets.items())}
    assert sizes == want["set_sizes"]
    out = str(tmp_path / "real.inp")
    mesh2voxelfe(mesh, cases.TEMPLATE, out, **(_grey_kw() if kind == "greyscale" else {}))
    deck = open(out, "rb").read()
    assert len(deck) == want["deck_bytes"] or abs(len(deck) - want["deck_bytes"]) == 7   # 19 / 26 char time stamp
    got = sections.section_hashes(deck)
    for sec in SECTIONS + ("header",):
        if sec == "header":
            assert got[sec]["sha256"] == want["sections"][sec]["sha256"]
        else:
            assert got[sec] == want["sections"][sec], sec


@pytest.mark.parametrize("n", [2, 4, 8])
@pytest.mark.parametrize("kind", ["binary", "greyscale"])
def test_real_deck_sections_sharded(real, kind, n, tmp_path):
    from ciclope_b200 import sharded
    rec, bw, kept, gv = real
    mask, winner = sharded.emulate_remove_unconnected(bw, n)
    assert np.array_equal(mask.cpu().numpy(), kept) and winner[0] == rec["kept_voxels"]
    out = str(tmp_path / "real_sharded.inp")
    sharded.emulate_ct2inp(kept if kind == "binary" else gv, n, rec["voxelsize"], cases.TEMPLATE, out,
                           m_kwargs=(_grey_kw() if kind == "greyscale" else {}))
    got = sections.section_hashes(open(out, "rb").read())
    for sec in SECTIONS:
        assert got[sec] == rec[kind]["sections"][sec], sec
