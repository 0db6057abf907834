#!/usr/bin/env python
"""Benchmark of the CT -> .INP hot path (remove_unconnected + vol2ugrid + mesh2voxelfe).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--size S]

One "step" = one pass of the whole path over one synthetic trabecular-like volume (BASELINE.json
configs[1]: S^3 binary volume, 30 % BV/TV; S = 512 per GPU by default).  `value` is voxels/s with
the volume resident in HBM and the deck left in HBM; `e2e` is the same path fed from a pinned HOST
volume with the deck copied back into a pinned HOST buffer inside the timed region.

Prints ONE JSON line on rank 0.  `--impl reference` times the CPU restatement of the reference
(oracle/, single-threaded C port) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
# the step allocates and frees multi-GB buffers (deck, lists): expandable segments keep torch's
# caching allocator from fragmenting into occasional cudaMalloc / cudaFree stalls
os.environ.setdefault("PYTORCH_CUDA_ALLOC_CONF", "expandable_segments:True")

METRIC = "voxels/s CT->.INP (CCL+vol2ugrid+deck)"
VOXELSIZE = [0.0195, 0.0195, 0.0195]
TEMPLATE = os.path.join(ROOT, "tests", "golden", "tmpl_compression.inp")
CPU_SAMPLE = 192          # edge of the cubic sample every CPU port process is timed on


# ------------------------------------------------------------------------------------------ CPU arm

def cpu_port_pass(sample):
    """One pass of the CPU port (oracle/) over a bool numpy volume; returns seconds per stage."""
    import tempfile

    from oracle import oracle as orc
    t0 = time.perf_counter()
    L = orc.remove_unconnected(sample)
    t1 = time.perf_counter()
    mesh = orc.vol2ugrid(L, VOXELSIZE)
    t2 = time.perf_counter()
    with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as d:
        orc.mesh2voxelfe(mesh, TEMPLATE, os.path.join(d, "cpu.inp"))
        size = os.path.getsize(os.path.join(d, "cpu.inp"))
    t3 = time.perf_counter()
    return dict(ccl=t1 - t0, vol2ugrid=t2 - t1, mesh2voxelfe=t3 - t2, total=t3 - t0, deck_bytes=size)


def cpu_sample(edge, seed=0):
    import torch

    from ciclope_b200 import synth
    return synth.trabecular((edge, edge, edge), seed, 0.30, "cpu").numpy()


def cpu_workers():
    """Host threads for the CPU arm: one single-threaded port process per core (at most 16), each on its own
    copy of the sample -- the path is embarrassingly parallel over independent volumes -- bounded by memory
    (a 192^3 pass peaks at ~1.5 GB of arrays + 0.3 GB of deck in /dev/shm)."""
    n = min(os.cpu_count() or 1, 16)
    try:
        with open("/proc/meminfo") as f:
            avail_kb = [int(ln.split()[1]) for ln in f if ln.startswith("MemAvailable:")][0]
        n = max(1, min(n, int(avail_kb / 1e6 / 4)))
    except (OSError, IndexError, ValueError):
        n = min(n, 4)
    return n


def _cpu_worker(i, sample, barrier, passes, warmup, q):
    try:
        for _ in range(warmup):
            cpu_port_pass(sample)
        barrier.wait()
        t0 = time.time()
        last = None
        for _ in range(passes):
            last = cpu_port_pass(sample)
        q.put((i, t0, time.time(), last))
    except BaseException as e:  # noqa: BLE001
        q.put((i, None, None, repr(e)))


def cpu_pool_run(edge, passes=1, warmup=0):
    """`workers` processes run `passes` passes of the CPU port each, started together; returns
    (voxels/s over all workers, seconds per pass, workers, stage times of one pass)."""
    import multiprocessing as mp
    sample = cpu_sample(edge)
    workers = cpu_workers()
    ctx = mp.get_context("fork")          # the children run NumPy + the C oracle only (no CUDA, no torch)
    barrier, q = ctx.Barrier(workers), ctx.Queue()
    procs = [ctx.Process(target=_cpu_worker, args=(i, sample, barrier, passes, warmup, q)) for i in range(workers)]
    for p in procs:
        p.start()
    res = [q.get() for _ in procs]
    for p in procs:
        p.join()
    bad = [r for r in res if r[1] is None]
    if bad:
        raise RuntimeError("CPU port worker failed: %s" % bad[0][3])
    wall = max(r[2] for r in res) - min(r[1] for r in res)
    return workers * passes * edge ** 3 / wall, wall / passes, workers, res[0][3]


def cpu_baseline_obj(edge, value, workers, times):
    return {"value": value, "unit": "voxels/s", "cores": workers, "kind": "port",
            "sample": "%d^3 synthetic trabecular sample (same generator as the workload) per process, %d "
                      "single-threaded port processes side by side; one pass of one process: ccl %.3fs "
                      "vol2ugrid %.3fs mesh2voxelfe %.3fs" % (edge, workers, times["ccl"], times["vol2ugrid"],
                                                              times["mesh2voxelfe"]),
            "host_cpus": os.cpu_count()}


def run_reference(args, rank):
    if rank != 0:
        return
    edge = CPU_SAMPLE
    value, dt, workers, last = cpu_pool_run(edge, passes=max(args.steps, 1), warmup=min(args.warmup, 1))
    base = cpu_baseline_obj(edge, value, workers, last)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "voxels/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": workload_config(args.size, args.gpus,
                                      note="CPU port: each step = one pass over a %d^3 sample in each of %d "
                                           "processes" % (edge, workers)),
            "cpu_baseline": base,
            "e2e": {"value": value, "unit": "voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit_line(line)


# ------------------------------------------------------------------------------------------ GPU arm

def workload_config(size, n_gpus, note=None):
    cfg = {"workload": "synthetic %d^3 trabecular-like binary volume (30%% BV/TV) per GPU: remove_unconnected + "
                       "vol2ugrid + mesh2voxelfe (BASELINE.json configs[1])" % size,
           "volume_per_gpu": [size, size, size], "voxelsize": VOXELSIZE[0], "bvtv": 0.30, "seed": 0,
           "template": "tests/golden/tmpl_compression.inp", "keywords": ["NSET", "ELSET"],
           "parallelism": "1 GPU" if n_gpus == 1 else
           "z-slabs: global volume [%d,%d,%d] cut into %d slabs of %d slices, one process per GPU, NCCL "
           "plane shifts + count/byte all-gathers" % (size * n_gpus, size, size, n_gpus, size),
           "l2": "inputs+outputs per step (0.13 GB volume, ~6 GB deck at 512^3) exceed the 126 MB L2; no flush"}
    if note:
        cfg["note"] = note
    return cfg


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index, period=0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:  # noqa: BLE001
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:  # noqa: BLE001
                pass
            self._stop_evt.wait(self.period)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def gpu_step(vol_dev):
    """One pass of the hot path, volume and deck resident in HBM."""
    from ciclope_b200.preprocess import remove_unconnected
    from ciclope_b200.voxelFE import deck_to_device, vol2ugrid
    L = remove_unconnected(vol_dev)
    mesh = vol2ugrid(L, VOXELSIZE)
    deck = deck_to_device(mesh, TEMPLATE)
    return mesh, deck


def gpu_step_e2e(vol_host, vol_dev, sink):
    """Same path through the public API with HOST buffers: pinned volume in, pinned deck out."""
    from ciclope_b200.preprocess import remove_unconnected
    from ciclope_b200.voxelFE import mesh2voxelfe, vol2ugrid
    vol_dev.copy_(vol_host, non_blocking=True)
    L = remove_unconnected(vol_dev)
    mesh = vol2ugrid(L, VOXELSIZE)
    n = mesh2voxelfe(mesh, TEMPLATE, sink)
    return mesh, n


class _Sized:
    """what the bookkeeping below needs from a step's result"""

    def __init__(self, n_el, n_nd, nbytes):
        self.n_el, self.n_nd, self.nbytes = n_el, n_nd, nbytes

    def numel(self):
        return self.nbytes


def make_sharded_steps(runner, Zg, z0):
    """N > 1: the same path as rank programs over torch.distributed (ciclope_b200.sharded)."""
    from ciclope_b200 import sharded

    def step(vol_dev):
        u8 = vol_dev.view(__import__("torch").uint8)
        mask, winner = runner.run(sharded.ccl_program(runner.ctx(), u8))
        if winner is None:
            raise ValueError("attempt to get argmax of an empty sequence")
        mesh = runner.run(sharded.vol2ugrid_program(runner.ctx(), mask, True, Zg, z0, VOXELSIZE))
        deck = runner.run(sharded.deck_program(runner.ctx(), mesh, TEMPLATE))
        return mesh, deck

    def step_e2e(vol_host, vol_dev, sink):
        vol_dev.copy_(vol_host, non_blocking=True)
        mesh, deck = step(vol_dev)
        pos = 0
        for _, buf, lo, hi in deck.writes:       # this rank's pieces of the global deck
            sink[pos:pos + hi - lo].copy_(buf[lo:hi], non_blocking=True)
            pos += hi - lo
        __import__("torch").cuda.current_stream().synchronize()
        return mesh, pos

    return step, step_e2e


def bind_to_gpu_numa_node(local_rank):
    """Run this rank on the CPUs next to its GPU (sysfs local_cpulist of the GPU's PCI function), so that
    the pinned host buffers of the end-to-end leg are allocated on that NUMA node: with 8 ranks copying
    6 GB decks at once, cross-socket pinned memory is what limits the aggregate D2H rate.  Best effort."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
        dom = torch.cuda.get_device_properties(local_rank).pci_domain_id
        dv = torch.cuda.get_device_properties(local_rank).pci_device_id
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/local_cpulist" % (dom, bus, dv)
        cpus = set()
        for part in open(path).read().strip().split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return "%s (%d cpus)" % (path.split("/")[-2], len(cpus))
    except Exception as e:  # noqa: BLE001 -- affinity is an optimisation, never a failure
        return "unbound (%s)" % type(e).__name__
    return "unbound"


def run_gpu(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    from ciclope_b200 import _lib, synth

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else "not bound (1 GPU)"
    if rank == 0:
        sys.stderr.write("[bench] host affinity: %s\n" % numa)
    lib = _lib.require_cuda()
    S = args.size
    global gpu_step, gpu_step_e2e
    if world == 1:
        vol = synth.trabecular((S, S, S), seed=0, bvtv=0.30, device=dev)
    else:
        from ciclope_b200 import sharded
        Zg, z0 = S * world, S * rank
        vol = synth.trabecular_slab((Zg, S, S), z0, z0 + S, seed=0, bvtv=0.30, device=dev)
        sh_step, sh_step_e2e = make_sharded_steps(sharded.DistRunner(), Zg, z0)

        def gpu_step(v):
            mesh, deck = sh_step(v)
            return mesh, _Sized(mesh.n_el, mesh.n_nd, deck.local_bytes)
        gpu_step_e2e = sh_step_e2e
    n_vox = vol.numel()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (also sizes the caching allocator and, for N > 1, NCCL's channels)
    n_warm = max(args.warmup, 3 if world == 1 else 5)
    for _ in range(n_warm):
        mesh, deck = gpu_step(vol)
    deck_bytes = deck.numel()
    n_el, n_nd = mesh.n_el, mesh.n_nd
    del mesh, deck

    # ---- timed region: K steps, CUDA events, max over ranks
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    l0 = lib.cfe_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    step_wall = []
    for _ in range(args.steps):
        t_s = time.perf_counter()
        mesh, deck = gpu_step(vol)          # every step ends with the host read of its byte totals
        del mesh, deck
        step_wall.append((time.perf_counter() - t_s) * 1e3)
    e1.record()
    barrier()
    launches = lib.cfe_launch_count() - l0
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = t.item()
    ms_per_step = ms_total / args.steps
    value = n_vox * world / (ms_per_step * 1e-3)

    # ---- per-entry-point device times (separate profiled steps, CUDA events on the launch stream)
    _lib.PROFILE = []
    prof_steps = 3
    for _ in range(prof_steps):
        mesh, deck = gpu_step(vol)
        del mesh, deck
    torch.cuda.synchronize()
    stage_ms = {}
    for name, a, b in _lib.PROFILE:
        stage_ms[name] = stage_ms.get(name, 0.0) + a.elapsed_time(b) / prof_steps
    _lib.PROFILE = None

    # ---- end to end through the public API with host buffers
    e2e = None
    if not args.no_e2e:
        vol_host = torch.empty((S, S, S), dtype=torch.bool).pin_memory()
        vol_host.copy_(vol)
        sink = torch.empty(int(deck_bytes * 1.02) + (1 << 20), dtype=torch.uint8).pin_memory()
        vol_dev = torch.empty_like(vol)
        for _ in range(2):
            gpu_step_e2e(vol_host, vol_dev, sink)
        barrier()
        k_e2e = max(3, min(args.steps, 5))
        t0 = time.perf_counter()
        for _ in range(k_e2e):
            _, n_out = gpu_step_e2e(vol_host, vol_dev, sink)
        barrier()
        dt = torch.tensor([(time.perf_counter() - t0) / k_e2e], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": n_vox * world / dt.item(), "unit": "voxels/s", "h2d_bytes_per_step": int(n_vox),
               "d2h_bytes_per_step": int(n_out) + 16 + 128 + 16, "ms_per_step": dt.item() * 1e3, "steps": k_e2e}
        del sink, vol_host

    if rank != 0:
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s"
    # dominant kernel: k_emit_elements2 (entry point cfe_emit_elements = that one launch).  Algorithmic
    # bytes per launch = element-section bytes written + 6 B/element read (voxel index, line length,
    # digit flags) -- DESIGN.md section 4.
    elem_bytes = deck_bytes - 53 * n_nd   # everything but the *NODE lines is element lines + O(N^2) text
    dom = max(stage_ms, key=stage_ms.get)
    alg = {"cfe_emit_elements": elem_bytes + 6 * n_el, "cfe_emit_nodes": 57 * n_nd}
    dom_for_roofline = "cfe_emit_elements"
    achieved = alg[dom_for_roofline] / (stage_ms[dom_for_roofline] * 1e-3) / 1e9
    traffic = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        if tr.get("size") == S:
            traffic = tr.get("k_emit_elements2")
    except Exception:  # noqa: BLE001
        pass
    roofline = {"bound": "hbm", "kernel": "k_emit_elements2 (entry point cfe_emit_elements, one launch per deck)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": alg[dom_for_roofline],
                "ms_per_launch": stage_ms[dom_for_roofline], "dominant_entry_point": dom,
                "other_kernels": {"k_emit_nodes4": {"achieved": alg["cfe_emit_nodes"] / (stage_ms["cfe_emit_nodes"] * 1e-3) / 1e9,
                                                    "frac": alg["cfe_emit_nodes"] / (stage_ms["cfe_emit_nodes"] * 1e-3) / 1e9 / peak,
                                                    "algorithmic_bytes_per_launch": alg["cfe_emit_nodes"],
                                                    "ms_per_launch": stage_ms["cfe_emit_nodes"]}}}
    total_alg = 2 * n_vox + deck_bytes
    cpu = None
    if world == 1 and not args.no_cpu:
        _v, _dt, _w, _t = cpu_pool_run(CPU_SAMPLE)
        cpu = cpu_baseline_obj(CPU_SAMPLE, _v, _w, _t)
    line = {"metric": METRIC, "value": value, "unit": "voxels/s", "n_gpus": world, "steps": args.steps,
            "warmup": n_warm, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": workload_config(S, world),
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
            "cpu_baseline": cpu,
            "job": {"voxels": n_vox * world, "elements_rank0": n_el, "used_nodes_rank0": n_nd,
                    "deck_bytes_rank0": deck_bytes, "algorithmic_bytes_rank0": total_alg,
                    "whole_path_frac_of_hbm_peak": total_alg / (ms_per_step * 1e-3) / 1e9 / peak},
            "step_ms": {"min": round(min(step_wall), 3), "median": round(statistics.median(step_wall), 3),
                        "max": round(max(step_wall), 3)},
            "entry_point_ms": {k: round(v, 4) for k, v in sorted(stage_ms.items(), key=lambda kv: -kv[1])}}
    emit_line(line)


_JSON_FD = None


def _claim_stdout():
    """Keep stdout for the ONE JSON line: everything else that writes to fd 1 (NCCL prints its version
    banner there) goes to stderr."""
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)


def emit_line(line):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ciclope_b200", choices=["ciclope_b200", "reference"])
    ap.add_argument("--size", type=int, default=512, help="edge of the cubic volume per GPU")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        if args.steps == 20:
            args.steps = 3
        run_reference(args, rank)
        return
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_gpu(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
