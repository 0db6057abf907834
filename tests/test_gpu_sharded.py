"""GPU tests (-m gpu) of the z-slab sharded path.  On a one-GPU box the N ranks are emulated in
lock-step inside one process (sharded.LocalRunner): same kernels, same offsets, same collectives'
results.  With >= 2 GPUs a real 2-rank NCCL run is compared with the single-GPU deck as well."""
import hashlib
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from tests import cases

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def sh():
    import torch
    assert torch.cuda.is_available()
    from ciclope_b200 import sharded
    return sharded


@pytest.fixture(scope="module")
def orc():
    from oracle import oracle
    return oracle


CCL_NAMES = [n for n, v in cases.ccl_cases() if v.shape[0] >= 2]


@pytest.mark.parametrize("n", [2, 3, 4, 8])
@pytest.mark.parametrize("name", CCL_NAMES)
def test_sharded_ccl_golden(sh, golden_ccl, name, n):
    shape = tuple(golden_ccl[name + "/shape"])
    if shape[0] < n:
        pytest.skip("fewer slices than slabs")
    k = int(np.prod(shape))
    bw = np.unpackbits(golden_ccl[name + "/in"])[:k].astype(bool).reshape(shape)
    want = np.unpackbits(golden_ccl[name + "/out"])[:k].astype(bool).reshape(shape)
    mask, _ = sh.emulate_remove_unconnected(bw, n)
    assert np.array_equal(mask.cpu().numpy(), want)


@pytest.mark.parametrize("n", [2, 4, 8])
@pytest.mark.parametrize("shape,density,seed,kind", [
    ((16, 24, 64), 0.3, 1, "rand"), ((24, 32, 96), 0.22, 2, "rand"), ((32, 48, 64), 0.3, 3, "blobs"),
    ((64, 64, 64), 0.12, 4, "blobs"), ((33, 20, 70), 0.4, 5, "rand"), ((48, 48, 128), 0.3, 6, "blobs")])
def test_sharded_ccl_vs_oracle(sh, orc, shape, density, seed, kind, n):
    bw = cases.rand_binary(shape, density, 500 + seed) if kind == "rand" else cases.blobs(shape, density, seed)
    mask, winner = sh.emulate_remove_unconnected(bw, n)
    want = orc.remove_unconnected(bw)
    assert np.array_equal(mask.cpu().numpy(), want)
    assert winner[0] == int(want.sum())


def test_sharded_ccl_u_shape_across_slabs(sh, orc):
    # two prongs that only join in the LAST slab: slab-local components must be merged globally, and a
    # bigger slab-internal blob must not win against the merged component
    bw = np.zeros((16, 12, 40), dtype=bool)
    bw[0:15, 2, 5] = True
    bw[0:15, 9, 30] = True
    bw[15, 2:10, 5] = True
    bw[15, 9, 5:31] = True
    bw[3:5, 5:8, 10:20] = True          # 60-voxel blob inside one slab (smaller than the U: 64)
    for n in (2, 4, 8, 16):
        mask, winner = sh.emulate_remove_unconnected(bw, n)
        assert np.array_equal(mask.cpu().numpy(), orc.remove_unconnected(bw)), n
    bw[3:6, 5:8, 10:20] = True          # now the blob (90) wins
    for n in (2, 4, 8):
        mask, _ = sh.emulate_remove_unconnected(bw, n)
        assert np.array_equal(mask.cpu().numpy(), orc.remove_unconnected(bw)), n


@pytest.mark.parametrize("mode", ["host", "overflow", "message"])
@pytest.mark.parametrize("n", [2, 4, 8])
def test_sharded_ccl_fallback_paths(sh, orc, n, mode, monkeypatch):
    """the host-driven merge, directly, through the capacity-overflow flag of the device path, and through an
    interface plane with more runs than the interface message carries"""
    if mode == "overflow":
        monkeypatch.setattr(sh, "BOUNDARY_VERTS", 4)
        monkeypatch.setattr(sh, "BOUNDARY_EDGES", 8)
    if mode == "message":
        monkeypatch.setattr(sh, "MSG_RUNS_MIN", 8)
        monkeypatch.setattr(sh, "MSG_RUNS_DIV", 10 ** 9)
    for bw in (cases.rand_binary((16, 24, 64), 0.3, 501), cases.blobs((32, 48, 64), 0.3, 3),
               cases.blobs((64, 64, 64), 0.12, 4)):
        mask, winner = sh.emulate_remove_unconnected(bw, n, force_host_merge=(mode == "host"))
        want = orc.remove_unconnected(bw)
        assert np.array_equal(mask.cpu().numpy(), want)
        assert winner[0] == int(want.sum())


def test_sharded_ccl_empty(sh):
    with pytest.raises(ValueError):
        sh.emulate_remove_unconnected(np.zeros((4, 4, 8), dtype=bool), 2)


DECK_NAMES = [c["name"] for c in cases.deck_cases() if np.asarray(c["vol"]).shape[0] >= 2]
DECK_CASES = {c["name"]: c for c in cases.deck_cases()}


@pytest.mark.parametrize("n", [2, 3, 8])
@pytest.mark.parametrize("name", DECK_NAMES)
def test_sharded_deck_golden(sh, golden, name, n, tmp_path):
    c = DECK_CASES[name]
    if np.asarray(c["vol"]).shape[0] < n:
        pytest.skip("fewer slices than slabs")
    u = dict(c["u"])
    vs = u.pop("voxelsize")
    m = dict(c["m"])
    tmpl = m.pop("templatefile")
    out = str(tmp_path / "deck.inp")
    sh.emulate_ct2inp(c["vol"], n, vs, tmpl, out, u_kwargs=u, m_kwargs=m)
    got = cases.mask_timestamp(open(out, "rb").read())
    want = bytes(golden[name + "/deck"])
    if got != want:
        for i, (a, b) in enumerate(zip(got.split(b"\n"), want.split(b"\n"))):
            assert a == b, "first differing line %d: got %r want %r" % (i, a, b)
    assert got == want


@pytest.mark.parametrize("n", [2, 4, 8])
@pytest.mark.parametrize("grey", [False, True])
def test_sharded_full_path_vs_oracle(sh, orc, n, grey, tmp_path):
    bw = cases.blobs((40, 48, 96), 0.3, 77)
    mask, _ = sh.emulate_remove_unconnected(bw, n)
    L = mask.cpu().numpy()
    L_ref = orc.remove_unconnected(bw)
    assert np.array_equal(L, L_ref)
    vol = cases.grey_from_mask(L, 78) if grey else L
    m = {}
    if grey:
        m = dict(keywords=['NSET', 'ELSET', 'PROPERTY'],
                 matprop=lambda: {"file": [cases.MAT_LINEAR], "range": [[1, 255]]})
    out = str(tmp_path / "gpu.inp")
    meshes, decks = sh.emulate_ct2inp(vol, n, [0.0195] * 3, cases.TEMPLATE, out, m_kwargs=m)
    mm = {k: (v() if callable(v) else v) for k, v in m.items()}
    orc.mesh2voxelfe(orc.vol2ugrid(vol, [0.0195] * 3), cases.TEMPLATE, str(tmp_path / "cpu.inp"), **mm)
    got = cases.mask_timestamp(open(out, "rb").read())
    want = cases.mask_timestamp(open(str(tmp_path / "cpu.inp"), "rb").read())
    assert len(got) == len(want)
    assert os.path.getsize(out) == decks[0].total_bytes
    assert hashlib.sha256(got).hexdigest() == hashlib.sha256(want).hexdigest()
    assert sum(mesh.n_el for mesh in meshes) == int(L.sum())


_NCCL_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch, torch.distributed as dist
from ciclope_b200 import sharded
from tests import cases
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
bw = cases.blobs((40, 48, 96), 0.3, 77)
a, b = sharded.slab_ranges(bw.shape[0], world)[rank]
slab = torch.from_numpy(bw[a:b].copy()).cuda().view(torch.uint8)
runner = sharded.DistRunner()
mask, winner = runner.run(sharded.ccl_program(runner.ctx(), slab))
mesh = runner.run(sharded.vol2ugrid_program(runner.ctx(), mask, True, bw.shape[0], a, [0.0195] * 3))
deck = runner.run(sharded.deck_program(runner.ctx(), mesh, cases.TEMPLATE))
runner.run(sharded.write_deck_program(runner.ctx(), deck, {out!r}))
dist.destroy_process_group()
"""


def test_real_nccl_two_ranks(orc, tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    out = str(tmp_path / "nccl.inp")
    script = tmp_path / "worker.py"
    script.write_text(_NCCL_WORKER.format(root=ROOT, out=out))
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                       cwd=ROOT, capture_output=True, timeout=600)
    assert r.returncode == 0, r.stderr.decode()[-3000:]
    bw = cases.blobs((40, 48, 96), 0.3, 77)
    L = orc.remove_unconnected(bw)
    orc.mesh2voxelfe(orc.vol2ugrid(L, [0.0195] * 3), cases.TEMPLATE, str(tmp_path / "cpu.inp"))
    assert cases.mask_timestamp(open(out, "rb").read()) == \
        cases.mask_timestamp(open(str(tmp_path / "cpu.inp"), "rb").read())


_FACADE_WORKER = """
import os, sys, json
sys.path.insert(0, {root!r})
import numpy as np, torch, torch.distributed as dist
from ciclope_b200 import distributed as cfd, sharded
from tests import cases
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
bw = cases.blobs((40, 48, 96), 0.3, 77)
a, b = sharded.slab_ranges(bw.shape[0], world)[rank]
L = cfd.remove_unconnected(bw[a:b])                                  # NumPy slab in -> NumPy slab out
assert isinstance(L, np.ndarray) and L.dtype == np.bool_
Lt = cfd.remove_unconnected(torch.from_numpy(bw[a:b].copy()).cuda())  # tensor in -> tagged tensor out
mesh, refn = cfd.vol2ugrid(Lt, [0.0195] * 3, refnodes=True)           # placement from the gathered slab heights
cfd.mesh2voxelfe(mesh, cases.TEMPLATE_REFNODE, {out!r}, refnode=[float(v) for v in refn["Z1"]])
grey = cases.grey_from_mask(bw, 13)
gm = cfd.vol2ugrid(grey[a:b], [0.0195] * 3, z0=a, Z_global=bw.shape[0])
cfd.mesh2voxelfe(gm, cases.TEMPLATE, {out_grey!r}, keywords=['NSET', 'ELSET', 'PROPERTY'],
                 matprop={{"file": [cases.MAT_LINEAR], "range": [[1, 255]]}})
try:
    cfd.remove_unconnected(np.zeros((4, 8, 32), dtype=bool))
    raised = False
except ValueError:
    raised = True
if rank == 0:
    json.dump({{"refnodes": {{k: [float(x) for x in v] for k, v in refn.items()}}, "empty_raises": raised,
               "mask_sum": int(L.sum())}}, open({meta!r}, "w"))
np.save({mask!r} + str(rank) + ".npy", L)
cfd.shutdown()
dist.destroy_process_group()
"""


def test_distributed_facade_two_ranks(orc, tmp_path):
    """ciclope_b200.distributed (the reference's signatures over real ranks: NCCL + peer-memory exchange) against
    the oracle: kept voxels, reference nodes (float64 sums chained up the slabs), the refnode deck and a
    grey-scale deck with PROPERTY cards, each written by both ranks into one file."""
    import json
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    out, out_grey = str(tmp_path / "facade.inp"), str(tmp_path / "facade_grey.inp")
    meta, mask = str(tmp_path / "meta.json"), str(tmp_path / "mask")
    script = tmp_path / "facade_worker.py"
    script.write_text(_FACADE_WORKER.format(root=ROOT, out=out, out_grey=out_grey, meta=meta, mask=mask))
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                       cwd=ROOT, capture_output=True, timeout=600)
    assert r.returncode == 0, r.stderr.decode()[-3000:]
    bw = cases.blobs((40, 48, 96), 0.3, 77)
    L = orc.remove_unconnected(bw)
    got = np.concatenate([np.load(mask + "%d.npy" % k) for k in range(2)])
    assert np.array_equal(got, L)
    m = json.load(open(meta))
    assert m["empty_raises"] and m["mask_sum"] > 0
    mesh, refn = orc.vol2ugrid(L, [0.0195] * 3, refnodes=True)
    for k, v in refn.items():
        assert np.array_equal(np.asarray(m["refnodes"][k]), v), k          # bit-exact float64 barycentres
    orc.mesh2voxelfe(mesh, cases.TEMPLATE_REFNODE, str(tmp_path / "cpu.inp"), refnode=[float(v) for v in refn["Z1"]])
    assert cases.mask_timestamp(open(out, "rb").read()) == cases.mask_timestamp(open(str(tmp_path / "cpu.inp"), "rb").read())
    grey = cases.grey_from_mask(bw, 13)
    orc.mesh2voxelfe(orc.vol2ugrid(grey, [0.0195] * 3), cases.TEMPLATE, str(tmp_path / "cpu_grey.inp"),
                     keywords=['NSET', 'ELSET', 'PROPERTY'], matprop={"file": [cases.MAT_LINEAR], "range": [[1, 255]]})
    assert cases.mask_timestamp(open(out_grey, "rb").read()) == \
        cases.mask_timestamp(open(str(tmp_path / "cpu_grey.inp"), "rb").read())


def test_refnodes_program_emulated(sh, orc):
    """the chained reference-node sums of the sharded path on emulated ranks (1-GPU boxes run this one)"""
    from ciclope_b200 import distributed as cfd
    bw = cases.blobs((30, 24, 64), 0.3, 5)
    L = orc.remove_unconnected(bw)
    _, want = orc.vol2ugrid(L, [0.0195, 0.02, 0.03], refnodes=True)
    for n in (2, 3, 5):
        slabs, is_bool, Zg = sh._slab_inputs(L, n)
        runner = sh.LocalRunner(n)
        meshes = runner.run([sh.vol2ugrid_program(c, s, is_bool, Zg, z0, [0.0195, 0.02, 0.03])
                             for c, (s, z0) in zip(runner.ctxs(), slabs)])
        refs = runner.run([cfd._refnodes_program(c, m) for c, m in zip(runner.ctxs(), meshes)])
        for got in refs:
            for k, v in want.items():
                assert np.array_equal(got[k], v), (n, k)


@pytest.mark.parametrize("case", ["wide", "digits_8_9", "digits_9_10", "digits_7_8"])
def test_ids_beyond_32_bits(sh, case):
    """Node / element ids of the 2048^3-class configs exceed 2^32: a thin slab placed at a huge global
    slice offset exercises the 64-bit formatting paths; the expected text is re-derived here from the
    voxel coordinates with plain Python integers.  The other cases keep the ids below 2^32 (the 32-bit
    variant of the element emitter) and put the powers of ten where digit counts change inside the slab:
    that is where the benchmark slab (ids up to 1.1e9, element ids up to 3.2e8) lives."""
    import torch
    Zl, Y, X = 6, 48, 64
    RN, SN = X + 1, (X + 1) * (Y + 1)
    if case == "wide":
        z0 = (1 << 32) // SN + 700_000        # node ids around 6.4e9
        eid_base = 5_000_000_000 - 300        # element ids cross 5e9 inside the slab
    elif case == "digits_8_9":
        z0 = 10 ** 9 // SN - 3                # node ids cross 1e9 inside the slab
        eid_base = 10 ** 8 - 300              # element ids cross 1e8
    elif case == "digits_9_10":
        z0 = 4 * 10 ** 9 // SN                # ten-digit node ids below 2^32
        eid_base = 10 ** 9 - 200              # element ids cross 1e9
    else:
        z0 = 10 ** 8 // SN - 2                # node ids cross 1e8
        eid_base = 10 ** 7 - 150              # element ids cross 1e7
    Zg = z0 + Zl                               # this slab is the last one of the (virtual) global volume
    vs = [0.01, 0.02, 0.00003]
    bw = cases.rand_binary((Zl, Y, X), 0.35, 77)
    runner = sh.LocalRunner(1)
    ctx = runner.ctxs()[0]
    slab = torch.from_numpy(bw).cuda().view(torch.uint8)
    mesh = runner.run([sh.vol2ugrid_program(ctx, slab, True, Zg, z0, vs, eid_base=eid_base)])[0]
    deck = runner.run([sh.deck_program(ctx, mesh, cases.TEMPLATE)])[0]
    pieces = {id(buf): bytes(buf[lo:hi].cpu().numpy()) for _, buf, lo, hi in deck.writes}
    nodes_txt = pieces[id(deck.buffers[0])].decode()
    elems_txt = pieces[id(deck.buffers[1])].decode()
    # expected, from first principles (voxelFE.py:145-154, 202, 630, 647)
    used = set()
    elines = []
    eid = eid_base
    for s, r, c in zip(*np.nonzero(bw)):
        s, r, c = int(s), int(r), int(c)
        b = SN * (s + z0) + RN * r + c
        n = [b, b + 1, b + RN + 1, b + RN, b + SN, b + SN + 1, b + SN + RN + 1, b + SN + RN]
        used.update(n)
        eid += 1
        elines.append(",".join(str(v) for v in [eid] + n))
    assert (max(used) > (1 << 32) and eid > 5_000_000_000) if case == "wide" else max(used) < (1 << 32)
    nlines = []
    for nid in sorted(used):
        s, rem = divmod(nid, SN)
        r, c = divmod(rem, RN)
        nlines.append('{0:10d}, {1:12.6f}, {2:12.6f}, {3:12.6f}'.format(nid, vs[0] * c, vs[1] * r, vs[2] * s))
    assert elems_txt == "\n".join(elines) + "\n"
    assert nodes_txt == "\n".join(nlines) + "\n"
    # the set ids are global too: CELLS_Z1 (last slice) holds the last elements
    z1 = [eid_base + 1 + i for i, (s, r, c) in enumerate(zip(*np.nonzero(bw))) if s == Zl - 1]
    got = mesh.set_ids[mesh.set_first[11]: mesh.set_first[11] + mesh.set_count[11]].cpu().tolist()
    assert got == z1


@pytest.mark.parametrize("n", [2, 8])
def test_foam_many_components_across_slabs(sh, orc, n):
    """BASELINE.json configs[3] in small: a fragmented foam (hundreds of components, many of them cut by the
    slab interfaces); single GPU and z-slabs against the CPU oracle."""
    import torch
    from ciclope_b200 import remove_unconnected, synth
    vol = synth.foam((160, 192, 192), seed=3, device="cuda")
    host = vol.cpu().numpy()
    assert orc.count_components(host) > 100
    want = orc.remove_unconnected(host)
    got = remove_unconnected(vol)
    assert np.array_equal(got.cpu().numpy(), want)
    mask, winner = sh.emulate_remove_unconnected(vol, n)
    assert np.array_equal(mask.cpu().numpy(), want) and winner[0] == int(want.sum())
    # components that straddle an interface: foreground voxels facing each other across the cut
    cuts = [a for a, _ in sh.slab_ranges(host.shape[0], n)][1:]
    assert sum(int((host[c - 1] & host[c]).sum()) for c in cuts) > 0
