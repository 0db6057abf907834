"""Timing of BASELINE.json configs[2] (greyscale volume, GV -> material mapping, PROPERTY section) on one GPU.
    python profiles/tools/bench_grey.py [size]
"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
os.environ.setdefault("PYTORCH_CUDA_ALLOC_CONF", "expandable_segments:True")
import torch
from ciclope_b200 import _lib, synth
from ciclope_b200.preprocess import remove_unconnected
from ciclope_b200.voxelFE import deck_to_device, vol2ugrid

S = int(sys.argv[1]) if len(sys.argv) > 1 else 512
VS = [0.0195] * 3
TEMPLATE = os.path.join(ROOT, "tests", "golden", "tmpl_compression.inp")
MAT = os.path.join(ROOT, "tests", "golden", "mat_linear.inp")
grey = synth.greyscale((S, S, S), seed=0, bvtv=0.30, device="cuda")
L = remove_unconnected(grey > 0)
grey = torch.where(L, grey, torch.zeros_like(grey))          # masked[~L] = 0 (ciclope example 01)


def step():
    mesh = vol2ugrid(grey, VS)
    return mesh, deck_to_device(mesh, TEMPLATE, matprop={"file": [MAT], "range": [[1, 255]]},
                                keywords=['NSET', 'ELSET', 'PROPERTY'])

for _ in range(4):
    mesh, deck = step()
n_el, nbytes = mesh.n_el, deck.numel()
del mesh, deck
torch.cuda.synchronize()
K = 10
t0 = time.perf_counter()
for _ in range(K):
    m, d = step(); del m, d
torch.cuda.synchronize()
ms = (time.perf_counter() - t0) / K * 1e3
_lib.PROFILE = []
for _ in range(3):
    m, d = step(); del m, d
torch.cuda.synchronize()
ep = {}
for name, a, b in _lib.PROFILE:
    ep[name] = ep.get(name, 0.0) + a.elapsed_time(b) / 3
print(json.dumps({"workload": "greyscale %d^3, vol2ugrid + mesh2voxelfe with PROPERTY (255 ELSETs)" % S, "ms_per_step": ms,
                  "voxels_per_s": S ** 3 / (ms * 1e-3), "n_el": n_el, "deck_bytes": nbytes,
                  "entry_point_ms": {k: round(v, 4) for k, v in sorted(ep.items(), key=lambda kv: -kv[1])}}))
