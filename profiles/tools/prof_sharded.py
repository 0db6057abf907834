"""torchrun --nproc-per-node N profiles/tools/prof_sharded.py : host-side profile of the sharded step."""
import cProfile, io, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, torch.distributed as dist
import bench
from ciclope_b200 import sharded, synth
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
S = 512
Zg, z0 = S * world, S * rank
vol = synth.trabecular_slab((Zg, S, S), z0, z0 + S, seed=0, bvtv=0.30, device="cuda")
runner = sharded.DistRunner()
step, _ = bench.make_sharded_steps(runner, Zg, z0)
for _ in range(5):
    step(vol)
torch.cuda.synchronize(); dist.barrier()
u8 = vol.view(torch.uint8)
for rep in range(3):
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    mask, w = runner.run(sharded.ccl_program(runner.ctx(), u8)); torch.cuda.synchronize(); t1 = time.perf_counter()
    mesh = runner.run(sharded.vol2ugrid_program(runner.ctx(), mask, True, Zg, z0, bench.VOXELSIZE)); torch.cuda.synchronize(); t2 = time.perf_counter()
    deck = runner.run(sharded.deck_program(runner.ctx(), mesh, bench.TEMPLATE)); torch.cuda.synchronize(); t3 = time.perf_counter()
    if rank == 0:
        print("stage ms: ccl %.3f ugrid %.3f deck %.3f total %.3f" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t3 - t0) * 1e3))
pr = cProfile.Profile(); pr.enable()
for _ in range(20):
    step(vol)
torch.cuda.synchronize()
pr.disable()
if rank == 0:
    s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('tottime').print_stats(30); print(s.getvalue()[:7000])
dist.destroy_process_group()
