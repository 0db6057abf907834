"""remove_unconnected alone, for CCL tuning.  python profiles/tools/bench_ccl.py [Z,Y,X | edge] [foam]
(alternative builds of the library: CICLOPE_B200_LIB=<.so>, see build_variant.sh)"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from ciclope_b200 import _lib, synth
from ciclope_b200.preprocess import remove_unconnected
shape = tuple(int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "512").split(","))
if len(shape) == 1:
    shape = shape * 3
foam = len(sys.argv) > 2 and sys.argv[2] == "foam"
vol = synth.foam(shape, 0, 0.12, 'cuda') if foam else synth.trabecular(shape, 0, 0.30, 'cuda')
for _ in range(3):
    remove_unconnected(vol)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    remove_unconnected(vol)
e1.record()
torch.cuda.synchronize()
_lib.PROFILE = []
for _ in range(3):
    remove_unconnected(vol)
torch.cuda.synchronize()
ms = {}
for name, a, b in _lib.PROFILE:
    ms[name] = ms.get(name, 0.0) + a.elapsed_time(b) / 3
print("%s %s %s: remove_unconnected %.3f ms  %s" % (os.environ.get("CICLOPE_B200_LIB", "default"), shape,
      "foam" if foam else "trabecular", e0.elapsed_time(e1) / 10, {k: round(v, 3) for k, v in ms.items()}))
