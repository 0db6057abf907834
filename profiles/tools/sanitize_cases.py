"""Small end-to-end cases for compute-sanitizer (memcheck / racecheck / synccheck):

    compute-sanitizer --tool memcheck python profiles/tools/sanitize_cases.py

smoke-sized binary deck, grey-scale deck with PROPERTY cards (GV-major partition), wide coordinate fields,
a volume whose CCL tiles overflow shared memory, X not a multiple of 32, and the z-slab path on 3 emulated
ranks -- every result is compared with the CPU oracle, so a silent corruption also fails the run."""
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from ciclope_b200 import mesh2voxelfe, remove_unconnected, sharded, vol2ugrid
from oracle import oracle as orc
from tests import cases

VS = [0.0195] * 3
tmp = tempfile.mkdtemp()


def deck(vol, name, u=None, m=None, oracle_vol=None):
    u, m = dict(u or {}), dict(m or {})
    u.setdefault("voxelsize", VS)
    out = os.path.join(tmp, name + ".inp")
    mesh2voxelfe(vol2ugrid(vol, **u), cases.TEMPLATE, out, **{k: (v() if callable(v) else v) for k, v in m.items()})
    ref = os.path.join(tmp, name + "_ref.inp")
    orc.mesh2voxelfe(orc.vol2ugrid(vol if oracle_vol is None else oracle_vol, **u), cases.TEMPLATE, ref,
                     **{k: (v() if callable(v) else v) for k, v in m.items()})
    a, b = (cases.mask_timestamp(open(f, "rb").read()) for f in (out, ref))
    assert a == b, name
    print("ok", name, len(a), flush=True)


bw = cases.blobs((24, 32, 64), 0.3, 5)
L = remove_unconnected(bw)
assert np.array_equal(L, orc.remove_unconnected(bw))
deck(L, "binary")
Lt = remove_unconnected(torch.from_numpy(bw).cuda())            # tagged mask: vol2ugrid starts from the packed bits
deck(Lt, "binary_tagged", oracle_vol=L)
grey = cases.grey_from_mask(cases.blobs((14, 16, 40), 0.3, 12), 13)
deck(grey, "grey_property", m=dict(keywords=['NSET', 'ELSET', 'PROPERTY'],
                                   matprop=lambda: {"file": [cases.MAT_LINEAR], "range": [[1, 255]]}))
deck(cases.rand_binary((5, 7, 9), 0.4, 1), "wide_fields", u=dict(voxelsize=[30000.5, 1e6, 123456.789]))
deck(cases.rand_binary((7, 9, 31), 0.2, 8), "x31_exact", u=dict(locking_strategy="exact", node_level_lock_num=2))
noise = cases.rand_binary((8, 32, 512), 0.5, 1008)               # > capacity runs per CCL tile
assert np.array_equal(remove_unconnected(noise), orc.remove_unconnected(noise))
print("ok ccl overflow tiles", flush=True)
bw3 = cases.blobs((30, 24, 64), 0.3, 77)
mask, winner = sharded.emulate_remove_unconnected(bw3, 3)
want = orc.remove_unconnected(bw3)
assert np.array_equal(mask.cpu().numpy(), want)
out = os.path.join(tmp, "sharded.inp")
sharded.emulate_ct2inp(want, 3, VS, cases.TEMPLATE, out)
orc.mesh2voxelfe(orc.vol2ugrid(want, VS), cases.TEMPLATE, os.path.join(tmp, "sharded_ref.inp"))
assert cases.mask_timestamp(open(out, "rb").read()) == cases.mask_timestamp(open(os.path.join(tmp, "sharded_ref.inp"), "rb").read())
print("ok sharded n=3", flush=True)
torch.cuda.synchronize()
print("ALL OK")
