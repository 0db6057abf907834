"""torchrun --nproc-per-node N profiles/tools/peer_check.py : the peer-memory exchange kernel against NCCL
(same results over hundreds of back-to-back exchanges of mixed sizes) and its latency next to all_gather."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
from ciclope_b200 import sharded

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
runner = sharded.DistRunner()
if rank == 0:
    print("peer exchange:", "on" if runner.peer is not None else "OFF (%s)" % runner.peer_error, flush=True)
assert runner.peer is not None, runner.peer_error
g = torch.Generator(device="cuda"); g.manual_seed(1234 + rank)
sizes = [1, 14, 3, 272, 4096, 1, 32768, 100000, 7, 131072, 5, 17, 1100000, 196610]
bad = 0
for it in range(300):
    n = sizes[it % len(sizes)]
    t = torch.randint(0, 2 ** 31, (n,), dtype=torch.int64, device="cuda", generator=g)
    want = torch.empty(world * n, dtype=torch.int64, device="cuda")
    dist.all_gather_into_tensor(want, t)
    # two shifts out of three exchanges, and a burst of consecutive shifts (neighbour-only synchronisation:
    # the slot-reuse rule of PeerExchange is what keeps them safe)
    if it % 3 != 0 or 100 <= it < 120:
        got = runner._fulfil("shift_up", t)
        if rank > 0:
            bad += int(not torch.equal(got, want.view(world, n)[rank - 1]))
        else:
            bad += int(got is not None)
    else:
        got = runner._gather(t)
        bad += int(not torch.equal(got.reshape(-1), want))
torch.cuda.synchronize()
tot = torch.tensor([bad], device="cuda"); dist.all_reduce(tot)
# latency: 200 back-to-back exchanges of 112 bytes (the count vector of the sharded path)
t = torch.arange(14, dtype=torch.int64, device="cuda")
res = {}
for name, fn in (("peer", lambda: runner._gather(t)),
                 ("nccl", lambda: dist.all_gather_into_tensor(torch.empty(world * 14, dtype=torch.int64, device="cuda"), t))):
    for _ in range(20):
        fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter(); e0.record()
    for _ in range(200):
        fn()
    e1.record(); host = (time.perf_counter() - w0) / 200 * 1e6
    torch.cuda.synchronize()
    res[name] = (e0.elapsed_time(e1) / 200 * 1e3, host)
for name, n, fn in (("shift of a 2048^2 interface (8.5 MB)", 1114112, lambda x: runner._fulfil("shift_up", x)),
                    ("all-gather of the boundary message (1.5 MB)", 196610, lambda x: runner._gather(x))):
    x = torch.zeros(n, dtype=torch.int64, device="cuda")
    for _ in range(5):
        fn(x)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        fn(x)
    e1.record()
    torch.cuda.synchronize()
    res[name] = (e0.elapsed_time(e1) / 20 * 1e3, 0.0)
if rank == 0:
    print("mismatches over 300 exchanges (all ranks):", int(tot.item()))
    for k, (dev_us, host_us) in res.items():
        print("%s: %.1f us per call on the device, %.1f us of host time to enqueue" % (k if " " in k else k + " all-gather of 112 B", dev_us, host_us))
dist.destroy_process_group()
sys.exit(1 if int(tot.item()) else 0)
