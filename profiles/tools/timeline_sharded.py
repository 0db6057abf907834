"""Kernel timeline of one sharded step on rank 0 (torchrun, N ranks): like timeline.py.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 profiles/tools/timeline_sharded.py
Diagnostic only."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import torch.distributed as dist
from torch.profiler import ProfilerActivity, profile

import bench
from ciclope_b200 import sharded, synth

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
Zl, Y, X = (int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "256,2048,2048").split(","))
Zg, z0 = Zl * world, Zl * rank
vol = synth.trabecular_slab((Zg, Y, X), z0, z0 + Zl, seed=0, bvtv=0.30, device=torch.device("cuda", lr))
hp = bench.HotPath(world, rank, Zg, z0)


def step(v):
    r = hp.step(v)
    return r[0], r[1]


for _ in range(6):
    m, d = step(vol)
    del m, d
torch.cuda.synchronize()
dist.barrier()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        m, d = step(vol)
        del m, d
    torch.cuda.synchronize()
if rank == 0:
    ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    ev.sort(key=lambda e: e.time_range.start)
    rows, t_prev = [], None
    for e in ev:
        st, en = e.time_range.start, e.time_range.end
        rows.append((st, en - st, (st - t_prev) if t_prev is not None else 0.0, e.name[:70]))
        t_prev = max(en, t_prev) if t_prev is not None else en
    n = len(rows) // 3
    last = rows[2 * n:]
    t0 = last[0][0]
    print("step: %d device activities, span %.1f us, busy %.1f us, idle %.1f us" % (
        len(last), last[-1][0] + last[-1][1] - t0, sum(r[1] for r in last), sum(r[2] for r in last[1:])))
    for st, du, gap, name in last:
        print("%10.1f %9.1f %9.1f  %s" % (st - t0, du, gap, name))
dist.destroy_process_group()
