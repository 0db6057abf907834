import sys, time, cProfile, pstats, io
sys.path.insert(0, '/root/repo')
import torch
import bench
from ciclope_b200 import synth
vol = synth.trabecular((512,512,512), 0, 0.30, 'cuda')
for _ in range(5): bench.gpu_step(vol)
torch.cuda.synchronize()
# stage-level wall times with sync after each stage
from ciclope_b200.preprocess import remove_unconnected
from ciclope_b200.voxelFE import deck_to_device, vol2ugrid
for rep in range(3):
    torch.cuda.synchronize(); t0=time.perf_counter()
    L = remove_unconnected(vol); torch.cuda.synchronize(); t1=time.perf_counter()
    mesh = vol2ugrid(L, bench.VOXELSIZE); torch.cuda.synchronize(); t2=time.perf_counter()
    deck = deck_to_device(mesh, bench.TEMPLATE); torch.cuda.synchronize(); t3=time.perf_counter()
    print("stage ms: ccl %.3f ugrid %.3f deck %.3f total %.3f"%((t1-t0)*1e3,(t2-t1)*1e3,(t3-t2)*1e3,(t3-t0)*1e3))
    del mesh, deck
pr = cProfile.Profile(); pr.enable()
for _ in range(20):
    m, d = bench.gpu_step(vol); del m, d
torch.cuda.synchronize()
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('tottime').print_stats(28); print(s.getvalue()[:6000])
