"""CCL stage alone on the fragmented foam of BASELINE.json configs[3] (remove_unconnected; single GPU, or
z-slabs under torchrun):   python profiles/tools/bench_foam.py [size]
                           torchrun --nproc-per-node N profiles/tools/bench_foam.py [size per GPU]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from ciclope_b200 import _lib, sharded, synth
from ciclope_b200.preprocess import remove_unconnected

S = int(sys.argv[1]) if len(sys.argv) > 1 else 512
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    vol = synth.trabecular_slab((S * world, S, S), S * rank, S * (rank + 1), seed=0, bvtv=0.12, device="cuda")
    runner = sharded.DistRunner()
    step = lambda: runner.run(sharded.ccl_program(runner.ctx(), vol.view(torch.uint8)))
else:
    vol = synth.foam((S, S, S), seed=0, device="cuda")
    step = lambda: remove_unconnected(vol)
for _ in range(3):
    out = step()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 10
e0.record()
for _ in range(K):
    out = step()
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / K], device="cuda")
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
_lib.PROFILE = []
for _ in range(3):
    step()
torch.cuda.synchronize()
ep = {}
for name, a, b in _lib.PROFILE:
    ep[name] = ep.get(name, 0.0) + a.elapsed_time(b) / 3
if rank == 0:
    kept = out[0] if isinstance(out, tuple) else out
    n_vox = S ** 3 * world
    print(json.dumps({"workload": "fragmented foam %d x %d^2 (BV/TV 0.12), remove_unconnected" % (S * world, S),
                      "n_gpus": world, "ms": float(ms.item()), "voxels_per_s": n_vox / (float(ms.item()) * 1e-3),
                      "algorithmic_GBps_per_gpu": 2 * S ** 3 / (float(ms.item()) * 1e-3) / 1e9,
                      "kept_voxels_rank0": int(kept.sum().item()),
                      "entry_point_ms": {k: round(v, 4) for k, v in sorted(ep.items(), key=lambda kv: -kv[1])}}))
if world > 1:
    dist.destroy_process_group()
