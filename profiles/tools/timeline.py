"""Kernel timeline of one device-resident step (torch.profiler / CUPTI): start, duration and the
idle gap before every kernel, to see where the GPU waits on the host.  Diagnostic only -- numbers
under a profiler are never bench values.

    python profiles/tools/timeline.py [Z,Y,X] [grey] > gpurun_out/timeline.txt
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import ProfilerActivity, profile

import bench
from ciclope_b200 import synth

shape = tuple(int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "256,2048,2048").split(","))
grey = len(sys.argv) > 2 and sys.argv[2] == "grey"
hp = bench.HotPath(1, 0, shape[0], 0)
if grey:
    g = synth.greyscale(shape, 0, 0.30, 'cuda')
    bw = g > 0
    kw = dict(keywords=['NSET', 'ELSET', 'PROPERTY'], matprop={"file": [bench.MAT_LINEAR], "range": [[1, 255]]})
else:
    g, bw, kw = None, synth.trabecular(shape, 0, 0.30, 'cuda'), {}


def step():
    r = hp.step(bw, g, None, **kw)
    del r


for _ in range(4):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
t_prev = None
rows = []
for e in ev:
    st, en = e.time_range.start, e.time_range.end
    gap = (st - t_prev) if t_prev is not None else 0.0
    rows.append((st, en - st, gap, e.name[:60]))
    t_prev = max(en, t_prev) if t_prev is not None else en
n = len(rows) // 3
last = rows[2 * n:]
t0 = last[0][0]
tot_k = sum(r[1] for r in last)
tot_gap = sum(r[2] for r in last[1:])
print("step: %d device activities, span %.1f us, busy %.1f us, idle %.1f us" % (
    len(last), last[-1][0] + last[-1][1] - t0, tot_k, tot_gap))
print("%10s %9s %9s  %s" % ("start_us", "dur_us", "gap_us", "name"))
for st, du, gap, name in last:
    print("%10.1f %9.1f %9.1f  %s" % (st - t0, du, gap, name))
