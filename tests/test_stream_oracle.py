"""The section-streamed oracle (oracle/stream_oracle.c, oracle.stream_deck) against the golden decks of the
unmodified reference: byte for byte through its file output, and section by section through its SHA-256
sinks.  This is what pins the streamed restatement that the GPU tests use at benchmark size."""
import hashlib
import json
import os

import numpy as np
import pytest

from ciclope_b200 import sections
from oracle import oracle as orc
from tests import cases

DECK_CASES = {c["name"]: c for c in cases.deck_cases()}


def test_sha256_implementation_matches_hashlib():
    import ctypes
    rng = np.random.RandomState(3)
    L = orc.lib()
    for n in [0, 1, 55, 56, 57, 63, 64, 65, 119, 120, 121, 1000, 1 << 20, (1 << 20) + 7]:
        data = rng.randint(0, 256, n, dtype=np.uint8).tobytes()
        out = ctypes.create_string_buffer(32)
        L.orc_sha256(data, n, out)
        assert out.raw == hashlib.sha256(data).digest(), n


def _kwargs(c):
    kw = dict(c["u"])
    m = dict(c["m"])
    if "matprop" in m:
        m["matprop"] = m["matprop"]()
    tmpl = m.pop("templatefile")
    kw.update(m)
    return tmpl, kw


@pytest.mark.parametrize("name", list(DECK_CASES))
def test_streamed_deck_equals_reference_deck(golden, name, tmp_path):
    c = DECK_CASES[name]
    tmpl, kw = _kwargs(c)
    out = str(tmp_path / "stream.inp")
    rec = orc.stream_deck(c["vol"], tmpl, fileout=out, **kw)
    want = bytes(golden[name + "/deck"])
    assert cases.mask_timestamp(open(out, "rb").read()) == want
    hashes = sections.section_hashes(want)
    for sec in ("nodes", "elements", "nset", "elset", "tail"):
        assert (sec in rec) == (sec in hashes), sec
        if sec in rec:
            assert rec[sec] == hashes[sec], sec
    assert rec["n_el"] == len(golden[name + "/cells"])
    assert rec["n_nd"] == len(np.unique(golden[name + "/cells"])) if rec["n_el"] else rec["n_nd"] == 0


def test_streamed_deck_midsize_hashes():
    mid = json.load(open(os.path.join(cases.GOLDEN_DIR, "midsize.json")))
    for c in cases.midsize_cases():
        tmpl, kw = _kwargs(c)
        rec = orc.stream_deck(c["vol"], tmpl, **kw)
        assert rec["n_el"] == mid[c["name"]]["n_el"] and rec["n_nd"] == mid[c["name"]]["n_used_nodes"]
        total = sum(rec[s]["bytes"] for s in ("nodes", "elements", "nset", "elset", "tail"))
        # + header: three comment lines (the stamp is masked in the golden size) + the two *NODE lines
        assert total < mid[c["name"]]["size"] < total + 400


def test_streamed_deck_real_lhdl():
    """BASELINE.json configs[0] stand-in: the reference's own LHDL stack thresholded at 63 (golden record
    made by tests/golden/make_golden_real.py from the unmodified reference)."""
    rec = json.load(open(os.path.join(cases.GOLDEN_DIR, "real_lhdl.json")))
    bw, kept, gv = cases.real_lhdl()
    assert int(bw.sum()) == rec["foreground_voxels"] and int(kept.sum()) == rec["kept_voxels"]
    assert orc.count_components(bw) == rec["components"]
    L = orc.remove_unconnected(bw)
    assert np.array_equal(L, kept)
    got = orc.stream_deck(L, cases.TEMPLATE, voxelsize=rec["voxelsize"])
    want = rec["binary"]
    assert got["n_el"] == want["n_el"] and got["n_nd"] == want["n_used_nodes"]
    for sec in ("nodes", "elements", "nset", "elset", "tail"):
        assert got[sec] == want["sections"][sec], sec
    got = orc.stream_deck(gv, cases.TEMPLATE, voxelsize=rec["voxelsize"], keywords=['NSET', 'ELSET', 'PROPERTY'],
                          matprop={"file": [cases.MAT_LINEAR], "range": [[1, 255]]})
    want = rec["greyscale"]
    for sec in ("nodes", "elements", "nset", "elset", "tail"):
        assert got[sec] == want["sections"][sec], sec
