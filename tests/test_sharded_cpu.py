"""CPU tests of the z-slab sharding host logic: slab geometry, the cross-slab component merge, the
lock-step rank emulator, and the torch.distributed runner over gloo with world_size 2."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch

from ciclope_b200 import sharded

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_slab_ranges():
    assert sharded.slab_ranges(10, 1) == [(0, 10)]
    assert sharded.slab_ranges(10, 3) == [(0, 4), (4, 7), (7, 10)]
    assert sharded.slab_ranges(8, 8) == [(k, k + 1) for k in range(8)]
    for Z in (8, 9, 100, 2048):
        for n in (1, 2, 3, 8):
            r = sharded.slab_ranges(Z, n)
            assert r[0][0] == 0 and r[-1][1] == Z
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1
    with pytest.raises(ValueError):
        sharded.slab_ranges(3, 4)
    assert sharded.exclusive_offsets([3, 0, 5]) == ([0, 3, 3], 8)


def test_clip_box_z():
    box = ((0, 10), (0, 5), (0, 7))
    assert sharded.clip_box_z(box, 4, 8) == ((4, 8), (0, 5), (0, 7))
    assert sharded.clip_box_z(box, 10, 12) == ((0, 0), (0, 0), (0, 0))


def _reference_components(keys, sizes, edges):
    parent = {int(k): int(k) for k in keys}

    def find(x):
        while parent[x] != x:
            parent[x] = parent[parent[x]]
            x = parent[x]
        return x
    for a, b in edges:
        ra, rb = find(int(a)), find(int(b))
        if ra != rb:
            parent[max(ra, rb)] = min(ra, rb)
    roots = [find(int(k)) for k in keys]
    comp = {}
    for r, s in zip(roots, sizes):
        comp[r] = comp.get(r, 0) + int(s)
    return roots, comp


@pytest.mark.parametrize("seed", range(6))
def test_solve_boundary_components_random(seed):
    rng = np.random.default_rng(seed)
    V = int(rng.integers(1, 400))
    keys = np.unique(rng.integers(0, 1 << 36, size=V))
    V = len(keys)
    sizes = rng.integers(1, 1000, size=V)
    E = int(rng.integers(0, 2 * V))
    edges = keys[rng.integers(0, V, size=(E, 2))] if E else np.zeros((0, 2), dtype=np.int64)
    lab, comp = sharded.solve_boundary_components(torch.from_numpy(keys), torch.from_numpy(sizes),
                                                  torch.from_numpy(edges.astype(np.int64)))
    roots, want = _reference_components(keys, sizes, edges)
    got_roots = keys[lab.numpy()]
    assert got_roots.tolist() == roots
    for i, r in enumerate(roots):
        assert int(comp[lab[i]].item()) == want[r]


def test_solve_boundary_components_long_chain():
    # a path graph given in the worst order: label propagation must still converge
    V = 257
    keys = torch.arange(V, dtype=torch.int64) * 7 + 3
    sizes = torch.ones(V, dtype=torch.int64)
    idx = torch.arange(V - 1)
    edges = torch.stack([keys[idx + 1], keys[idx]], dim=1).flip(0)
    lab, comp = sharded.solve_boundary_components(keys, sizes, edges)
    assert torch.all(lab == 0) and int(comp[0]) == V


def test_pick_winner_ties():
    assert sharded.pick_winner([(5, 9), (7, 3), (7, 2), (0, 1)]) == (7, 2)
    assert sharded.pick_winner([(0, 1), (0, 2)]) is None
    assert sharded.pick_winner([]) is None


def _toy_program(ctx, x):
    got = yield from ctx.all_gather(torch.tensor([x + ctx.rank], dtype=torch.int64))
    var = yield from ctx.all_gather_var(torch.arange(ctx.rank + 1, dtype=torch.int64))
    below = yield from ctx.shift_up(torch.tensor([100 + ctx.rank], dtype=torch.int64))
    yield from ctx.barrier()
    return ([int(t.item()) for t in got], [t.tolist() for t in var], None if below is None else int(below.item()))


def test_local_runner_lockstep():
    runner = sharded.LocalRunner(3)
    res = runner.run([_toy_program(c, 10) for c in runner.ctxs()])
    for k, (g, v, b) in enumerate(res):
        assert g == [10, 11, 12]
        assert v == [[0], [0, 1], [0, 1, 2]]
        assert b == (None if k == 0 else 100 + k - 1)


_WORKER = r"""
import os, sys, json
sys.path.insert(0, {root!r})
import torch, torch.distributed as dist
from ciclope_b200 import sharded
from tests.test_sharded_cpu import _toy_program
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
runner = sharded.DistRunner()
res = runner.run(_toy_program(runner.ctx(), 10))
# cross-slab merge on gathered boundary data, identical on both ranks
keys = torch.tensor([3, 5, (1 << 32) | 2, (1 << 32) | 9], dtype=torch.int64)
sizes = torch.tensor([4, 6, 10, 1], dtype=torch.int64)
mine = torch.tensor([[5, (1 << 32) | 2]], dtype=torch.int64) if dist.get_rank() == 1 else torch.zeros((0, 2), dtype=torch.int64)
def prog(ctx):
    e = yield from ctx.all_gather_var(mine.reshape(-1))
    return torch.cat(e).reshape(-1, 2)
edges = runner.run(prog(runner.ctx()))
lab, comp = sharded.solve_boundary_components(keys, sizes, edges)
print(json.dumps({{"rank": dist.get_rank(), "res": res, "lab": lab.tolist(), "comp": comp.tolist()}}))
dist.destroy_process_group()
"""


def test_dist_runner_gloo_world2(tmp_path):
    sock = socket.socket()
    sock.bind(("127.0.0.1", 0))
    port = sock.getsockname()[1]
    sock.close()
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                              cwd=ROOT) for r in range(2)]
    outs = []
    for p in procs:
        out, err = p.communicate(timeout=120)
        assert p.returncode == 0, err.decode()[-2000:]
        outs.append(out.decode().strip().splitlines()[-1])
    import json
    for line in outs:
        d = json.loads(line)
        g, v, b = d["res"]
        assert g == [10, 11] and v == [[0], [0, 1]]
        assert b == (None if d["rank"] == 0 else 100)
        assert d["lab"] == [0, 1, 1, 3]
        assert d["comp"][1] == 16 and d["comp"][0] == 4 and d["comp"][3] == 1
