"""GPU tests (-m gpu) at the benchmark size (512^3, BASELINE.json configs[1]) through size-independent
properties -- the CPU oracle needs minutes and tens of GB there:

* sharding invariance: the deck of the z-slab path (4 emulated ranks) is byte-identical to the
  single-GPU deck (compared on the device);
* CCL: idempotence, mask is a subset of the input, kept voxels == winner size, sharded == single;
* structure: line count, fixed 53-byte *NODE lines, ascending ids, element lines re-derived from the
  voxel coordinates for a sample of elements;
* oracle parity at 256^3 with the materialising C oracle, and at 512^3 (binary and grey-scale with 255 material
  sets) with the section-streamed oracle: per-section SHA-256 of the CUDA deck, streamed off the GPU, against
  oracle.stream_deck -- which tests/test_stream_oracle.py pins on every golden deck of the reference.
"""
import hashlib
import os

import numpy as np
import pytest

from tests import cases

pytestmark = pytest.mark.gpu
VS = [0.0195, 0.0195, 0.0195]


@pytest.fixture(scope="module")
def big():
    import torch
    from ciclope_b200 import synth
    assert torch.cuda.is_available()
    return synth.trabecular((512, 512, 512), seed=0, bvtv=0.30, device="cuda")


def test_fullsize_ccl_properties(big):
    import torch
    from ciclope_b200 import remove_unconnected, sharded
    from ciclope_b200.preprocess import component_stats, largest_component_device
    L = remove_unconnected(big)
    assert L.dtype == torch.bool and L.shape == big.shape
    assert bool((L & ~big).sum() == 0)                          # subset of the input
    mask, best = largest_component_device(big.view(torch.uint8))
    size, _root = component_stats(best)
    assert int(L.sum()) == size                                  # kept voxels == winner size
    assert torch.equal(remove_unconnected(L), L)                 # idempotent
    m4, winner = sharded.emulate_remove_unconnected(big, 4)     # z-slab path, 4 ranks
    assert torch.equal(m4, L) and winner[0] == size


def test_fullsize_deck_sharding_invariance_and_structure(big, tmp_path):
    import torch
    from ciclope_b200 import remove_unconnected, sharded, vol2ugrid
    from ciclope_b200.voxelFE import deck_to_device
    L = remove_unconnected(big)
    mesh = vol2ugrid(L, VS)
    deck = deck_to_device(mesh, cases.TEMPLATE)
    Z, Y, X = L.shape
    RN, SN = X + 1, (X + 1) * (Y + 1)
    # ---- structure
    head_end = 5 * 0 + int((deck[:400] == 10).nonzero()[4].item()) + 1        # after the 5 header lines
    nodes = deck[head_end: head_end + 53 * mesh.n_nd].view(mesh.n_nd, 53)
    assert bool((nodes[:, 52] == 10).all()) and bool((nodes[:, 10] == ord(',')).all())
    ids = torch.zeros(mesh.n_nd, dtype=torch.int64, device="cuda")
    for k in range(10):                                          # parse '{:10d}'
        d = nodes[:, k].to(torch.int64)
        ids = ids * 10 + torch.where(d == 32, torch.zeros_like(d), d - 48)
    assert bool((ids[1:] > ids[:-1]).all())                      # ascending, unique
    used = torch.zeros((Z + 1) * (Y + 1) * (X + 1), dtype=torch.bool, device="cuda")
    used[ids] = True
    idx = L.nonzero()
    base = idx[:, 0] * SN + idx[:, 1] * RN + idx[:, 2]
    for off in (0, 1, RN, RN + 1, SN, SN + 1, SN + RN, SN + RN + 1):
        assert bool(used[base + off].all())                      # every element corner has a *NODE line
    assert int(used.sum()) == mesh.n_nd
    n_lines = int((deck == 10).sum())
    sets = sum(mesh.set_count) + min(2, mesh.set_count[4]) + min(2, mesh.set_count[5])
    assert n_lines >= 5 + mesh.n_nd + 2 + mesh.n_el                # header, nodes, comment + *ELEMENT, elements
    assert n_lines <= 5 + mesh.n_nd + 2 + mesh.n_el + 4 + 2 * 14 + sets // 10 + 14 + 40
    # a sample of element lines, re-derived from the voxel coordinates
    el0 = head_end + 53 * mesh.n_nd
    text = bytes(deck[el0: el0 + 4096].cpu().numpy()).decode().split("\n")
    assert text[0] == "** Elements and Element sets from input model"
    assert text[1] == "*ELEMENT, TYPE=C3D8, ELSET=SET1"
    b = base[:20].cpu().tolist()
    for e in range(20):
        want = ",".join(str(v) for v in [e + 1, b[e], b[e] + 1, b[e] + RN + 1, b[e] + RN, b[e] + SN, b[e] + SN + 1,
                                         b[e] + SN + RN + 1, b[e] + SN + RN])
        assert text[2 + e] == want
    tail = bytes(deck[-200:].cpu().numpy()).decode()
    assert tail.endswith(open(cases.TEMPLATE).read()[-100:])
    # ---- sharding invariance: 4 emulated ranks write the same bytes
    out = str(tmp_path / "sharded.inp")
    del used, ids, nodes
    sharded.emulate_ct2inp(L, 4, VS, cases.TEMPLATE, out)
    assert os.path.getsize(out) == deck.numel()
    h = hashlib.sha256()
    h2 = hashlib.sha256()
    with open(out, "rb") as f:
        pos = 0
        while True:
            chunk = f.read(64 << 20)
            if not chunk:
                break
            mine = deck[pos: pos + len(chunk)].cpu().numpy().tobytes()
            h.update(cases.mask_timestamp(chunk) if pos == 0 else chunk)
            h2.update(cases.mask_timestamp(mine) if pos == 0 else mine)
            pos += len(chunk)
    assert h.hexdigest() == h2.hexdigest()


def test_oracle_parity_256(tmp_path):
    from ciclope_b200 import mesh2voxelfe, remove_unconnected, synth, vol2ugrid
    from oracle import oracle as orc
    bw = synth.trabecular((256, 256, 256), seed=1, bvtv=0.30, device="cuda").cpu().numpy()
    L = remove_unconnected(bw)
    L_ref = orc.remove_unconnected(bw)
    assert np.array_equal(L, L_ref)
    mesh2voxelfe(vol2ugrid(L, VS), cases.TEMPLATE, str(tmp_path / "gpu.inp"))
    orc.mesh2voxelfe(orc.vol2ugrid(L_ref, VS), cases.TEMPLATE, str(tmp_path / "cpu.inp"))
    hs = []
    for name in ("gpu.inp", "cpu.inp"):
        h = hashlib.sha256()
        with open(str(tmp_path / name), "rb") as f:
            first = True
            while True:
                chunk = f.read(64 << 20)
                if not chunk:
                    break
                h.update(cases.mask_timestamp(chunk) if first else chunk)
                first = False
        hs.append((h.hexdigest(), os.path.getsize(str(tmp_path / name))))
    assert hs[0] == hs[1]


def test_fullsize_foam_sharded_equals_single():
    """configs[3] at 512^3: ~4 * 10^3 components; 8 emulated z-slabs keep the same component as one GPU"""
    import torch
    from ciclope_b200 import remove_unconnected, sharded, synth
    from ciclope_b200.preprocess import component_stats, largest_component_device
    vol = synth.foam((512, 512, 512), seed=0, device="cuda")
    L = remove_unconnected(vol)
    _, best = largest_component_device(vol.view(torch.uint8))
    size, _root = component_stats(best)
    assert int(L.sum()) == size and bool((L & ~vol).sum() == 0)
    assert size < 0.5 * int(vol.sum())                           # fragmented: no dominant component
    m8, winner = sharded.emulate_remove_unconnected(vol, 8)
    assert torch.equal(m8, L) and winner[0] == size


# ------------------------------------------------------------------ oracle parity at the benchmark size

SECTIONS = ("nodes", "elements", "nset", "elset", "tail")


@pytest.fixture(scope="module")
def oracle_512(big):
    """Streamed oracle of the 512^3 binary and grey-scale decks (two host threads, ~40 s; the C code
    releases the GIL) + the oracle's remove_unconnected of the binary volume."""
    import threading

    from ciclope_b200 import synth
    from oracle import oracle as orc
    bw = big.cpu().numpy()
    grey = synth.greyscale((512, 512, 512), seed=0, bvtv=0.30, device="cuda")
    grey_host = grey.cpu().numpy()
    out = {"grey": grey}

    def binary():
        out["L"] = orc.remove_unconnected(bw)
        out["bin"] = orc.stream_deck(out["L"], cases.TEMPLATE, voxelsize=VS)

    def greyscale():
        out["gv"] = orc.stream_deck(grey_host, cases.TEMPLATE, voxelsize=VS, keywords=['NSET', 'ELSET', 'PROPERTY'],
                                    matprop={"file": [cases.MAT_LINEAR], "range": [[1, 255]]})

    th = [threading.Thread(target=binary), threading.Thread(target=greyscale)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert "bin" in out and "gv" in out, "oracle thread failed"
    return out


def test_oracle_parity_512_binary(big, oracle_512):
    import torch
    from ciclope_b200 import mesh2voxelfe, remove_unconnected, vol2ugrid
    L = remove_unconnected(big)
    assert torch.equal(L.cpu(), torch.from_numpy(oracle_512["L"]))          # labels kept: bit-exact at 512^3
    mesh = vol2ugrid(L, VS)
    want = oracle_512["bin"]
    assert (mesh.n_el, mesh.n_nd) == (want["n_el"], want["n_nd"])
    sink = cases.SectionStreamHasher(want)
    n = mesh2voxelfe(mesh, cases.TEMPLATE, sink)
    got = sink.result()
    assert n == len(sink.head) + sum(v["bytes"] for v in got.values())
    for sec in SECTIONS:
        assert got[sec] == want[sec], sec


def test_oracle_parity_512_greyscale(oracle_512):
    from ciclope_b200 import mesh2voxelfe, vol2ugrid
    mesh = vol2ugrid(oracle_512["grey"], VS)
    want = oracle_512["gv"]
    assert (mesh.n_el, mesh.n_nd) == (want["n_el"], want["n_nd"])
    sink = cases.SectionStreamHasher(want)
    mesh2voxelfe(mesh, cases.TEMPLATE, sink, keywords=['NSET', 'ELSET', 'PROPERTY'],
                 matprop={"file": [cases.MAT_LINEAR], "range": [[1, 255]]})
    got = sink.result()
    for sec in SECTIONS:
        assert got[sec] == want[sec], sec


def test_oracle_parity_512_sharded_sections(big, oracle_512, tmp_path):
    """the z-slab path (8 emulated ranks, pinned-ring writer, one file) against the same oracle record"""
    import torch
    from ciclope_b200 import sharded
    L = torch.from_numpy(oracle_512["L"]).cuda()
    out = str(tmp_path / "sharded.inp")
    sharded.emulate_ct2inp(L, 8, VS, cases.TEMPLATE, out)
    want = oracle_512["bin"]
    sink = cases.SectionStreamHasher(want)
    with open(out, "rb") as f:
        while True:
            chunk = f.read(64 << 20)
            if not chunk:
                break
            sink.write(chunk)
    os.remove(out)
    got = sink.result()
    for sec in SECTIONS:
        assert got[sec] == want[sec], sec
