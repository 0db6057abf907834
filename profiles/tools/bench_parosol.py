"""Timing of vol2h5ParOSol's dataset generation (voxelFE.py:285-472) and of segment on one GPU.
    python profiles/tools/bench_parosol.py [size]
"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from ciclope_b200 import _lib, synth
from ciclope_b200.parosol import parosol_datasets
from ciclope_b200.preprocess import gaussian, remove_unconnected, segment

S = int(sys.argv[1]) if len(sys.argv) > 1 else 512
grey = synth.greyscale((S, S, S), seed=0, bvtv=0.30, device="cuda")
bw = remove_unconnected(grey > 0)
res = {"workload": "%d^3 synthetic trabecular volume" % S}


def timeit(fn, k=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(k):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / k * 1e3


def profile(fn):
    _lib.PROFILE = []
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ep = {}
    for name, a, b in _lib.PROFILE:
        ep[name] = ep.get(name, 0.0) + a.elapsed_time(b) / 3
    _lib.PROFILE = None
    return {k: round(v, 4) for k, v in sorted(ep.items(), key=lambda kv: -kv[1])}


f = lambda: parosol_datasets(bw, -0.04, voxelsize=0.0195, image_on_device=True, node_order="sorted")
res["parosol_sorted_ms"] = timeit(f)
res["parosol_entry_point_ms"] = profile(f)
f2 = lambda: parosol_datasets(bw, -0.04, voxelsize=0.0195, image_on_device=True)
res["parosol_reference_order_ms"] = timeit(f2, k=2)
g = lambda: segment(grey, 80)
res["segment_ms"] = timeit(g)
res["segment_entry_point_ms"] = profile(g)
res["segment_gbs"] = 2 * S ** 3 / (res["segment_entry_point_ms"]["cfe_threshold_u8"] * 1e-3) / 1e9
if S <= 512:
    gs = lambda: gaussian(grey, sigma=1.0, preserve_range=True)
    res["gaussian_sigma1_ms"] = timeit(gs, k=3)
    res["gaussian_entry_point_ms"] = profile(gs)
ds = f()
img = ds["Image"]
res["image_shape"] = list(img.shape)
res["image_gbs"] = 5 * img.numel() / (res["parosol_entry_point_ms"]["cfe_scale_crop_f32"] * 1e-3) / 1e9
res["bc_rows"] = int(ds["Fixed_Displacement_Coordinates"].shape[0])
if S <= 256:
    sys.path.insert(0, ROOT)
    from oracle import oracle as orc
    h = bw.cpu().numpy()
    t0 = time.perf_counter()
    orc.vol2h5ParOSol_datasets(h, -0.04, voxelsize=0.0195)
    res["cpu_oracle_ms"] = (time.perf_counter() - t0) * 1e3
print(json.dumps(res))
