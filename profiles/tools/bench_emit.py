"""vol2ugrid + deck of the benchmark slab, per entry point (CUDA events); [grey] = 1024^3 grey-scale deck
python profiles/tools/bench_emit.py [Z,Y,X] [grey]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from ciclope_b200 import _lib, synth, vol2ugrid
from ciclope_b200.voxelFE import deck_to_device
shape = tuple(int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "256,2048,2048").split(","))
grey = len(sys.argv) > 2 and sys.argv[2] == "grey"
vol = synth.greyscale(shape, 0, 0.30, 'cuda') if grey else synth.trabecular(shape, 0, 0.30, 'cuda')
tmpl = os.path.join(ROOT, "tests", "golden", "tmpl_compression.inp")
kw = dict(keywords=['NSET', 'ELSET', 'PROPERTY'], matprop={"file": [os.path.join(ROOT, "tests", "golden", "mat_linear.inp")], "range": [[1, 255]]}) if grey else {}
for _ in range(3):
    d = deck_to_device(vol2ugrid(vol, [0.0195] * 3), tmpl, **kw); del d
torch.cuda.synchronize()
_lib.PROFILE = []
for _ in range(5):
    d = deck_to_device(vol2ugrid(vol, [0.0195] * 3), tmpl, **kw); del d
torch.cuda.synchronize()
ms = {}
for name, a, b in _lib.PROFILE:
    ms[name] = ms.get(name, 0.0) + a.elapsed_time(b) / 5
print(os.environ.get("CICLOPE_B200_LIB", "default"), "grey" if grey else "binary", {k: round(v, 3) for k, v in sorted(ms.items(), key=lambda kv: -kv[1])[:14]})
