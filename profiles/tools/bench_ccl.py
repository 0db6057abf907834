"""remove_unconnected alone (per-kernel CUDA-event times), for CCL tuning.  python profiles/tools/bench_ccl.py [size]"""
import sys
sys.path.insert(0, '/root/repo')
import torch
from ciclope_b200 import _lib, synth
from ciclope_b200.preprocess import remove_unconnected
S = int(sys.argv[1]) if len(sys.argv) > 1 else 512
vol = synth.trabecular((S, S, S), 0, 0.30, 'cuda')
for _ in range(3):
    remove_unconnected(vol)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    remove_unconnected(vol)
e1.record()
torch.cuda.synchronize()
print("size %d: remove_unconnected %.3f ms" % (S, e0.elapsed_time(e1) / 10))
