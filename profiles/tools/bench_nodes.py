"""vol2ugrid + the *NODE emitter alone (per entry point CUDA-event times).  python profiles/tools/bench_nodes.py [Z,Y,X]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from ciclope_b200 import _lib, synth, vol2ugrid
from ciclope_b200.voxelFE import deck_to_device
shape = tuple(int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "256,2048,2048").split(","))
vol = synth.trabecular(shape, 0, 0.30, 'cuda')
tmpl = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests", "golden", "tmpl_compression.inp")
for _ in range(3):
    d = deck_to_device(vol2ugrid(vol, [0.0195] * 3), tmpl); del d
torch.cuda.synchronize()
_lib.PROFILE = []
for _ in range(5):
    d = deck_to_device(vol2ugrid(vol, [0.0195] * 3), tmpl); del d
torch.cuda.synchronize()
ms = {}
for name, a, b in _lib.PROFILE:
    ms[name] = ms.get(name, 0.0) + a.elapsed_time(b) / 5
print(os.environ.get("CICLOPE_B200_LIB", "default"), {k: round(v, 3) for k, v in sorted(ms.items(), key=lambda kv: -kv[1])[:6]})
