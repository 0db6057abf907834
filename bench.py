#!/usr/bin/env python
"""Benchmark of the CT -> .INP hot path (remove_unconnected + vol2ugrid + mesh2voxelfe).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--shape Z,Y,X]

One "step" = one pass of the whole path over one synthetic trabecular-like volume.  The default workload
is 2^30 voxels per GPU as a 256 x 2048 x 2048 z-slab of a binary volume at 30 % BV/TV: at N = 1 that is a
1024^3-class job on one GPU, at N = 8 the global volume is exactly the 2048^3 cube of BASELINE.json
configs[4] (weak scaling, one process per GPU, z-slabs).  `value` is voxels/s with the volume resident in
HBM and the deck left in HBM; `e2e` is the same path fed from a pinned HOST volume with the deck streamed
back into pinned HOST memory inside the timed region.

The same JSON line carries `configs`: BASELINE.json configs[2] (1024^3 uint8 grey-scale volume, 255
ELSETs + PROPERTY cards; cut into N z-slabs) and configs[3] (fragmented foam of the workload's shape: CCL
stage alone and end to end, component count, components cut by slab interfaces), and `parity_check`: small
volumes run through the same code path as the timed one (for N > 1: real ranks, peer-memory exchange,
NCCL) and compared section by section (SHA-256) with the single-GPU deck; a mismatch fails the run.

Prints ONE JSON line on rank 0.  `--impl reference` times the CPU restatement of the reference
(oracle/, C port, one process per host core) on a bounded sample of the same workload.
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
MAT_LINEAR = os.path.join(ROOT, "tests", "golden", "mat_linear.inp")
DEFAULT_SHAPE = (256, 2048, 2048)     # per GPU; N = 8 slabs = the 2048^3 cube
GREY_GLOBAL = (1024, 1024, 1024)      # BASELINE.json configs[2], cut into N z-slabs
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
                      "single-threaded port processes side by side (one step = one pass of every process); one "
                      "pass of one process: ccl %.3fs vol2ugrid %.3fs mesh2voxelfe %.3fs"
                      % (edge, workers, times["ccl"], times["vol2ugrid"], times["mesh2voxelfe"]),
            "host_cpus": os.cpu_count()}


def run_reference(args, rank):
    if rank != 0:
        return
    edge = CPU_SAMPLE
    value, dt, workers, last = cpu_pool_run(edge, passes=max(args.steps, 1), warmup=max(args.warmup, 0))
    base = cpu_baseline_obj(edge, value, workers, last)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "voxels/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": workload_config(args.shape, args.gpus),
            "cpu_baseline": base,
            "e2e": {"value": value, "unit": "voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit_line(line)


# ------------------------------------------------------------------------------------------ GPU arm

def workload_config(shape, n_gpus):
    Z, Y, X = shape
    return {"workload": "synthetic %dx%dx%d trabecular-like binary z-slab (30%% BV/TV) per GPU = 2^%.2f voxels: "
                        "remove_unconnected + vol2ugrid + mesh2voxelfe; N GPUs = N slabs of one global volume "
                        "(N = 8: the 2048^3 cube of BASELINE.json configs[4]; N = 1: a 1024^3-class job)"
                        % (Z, Y, X, __import__("math").log2(Z * Y * X)),
            "volume_per_gpu": [Z, Y, X], "global_volume": [Z * n_gpus, Y, X], "voxelsize": VOXELSIZE[0],
            "bvtv": 0.30, "seed": 0, "template": "tests/golden/tmpl_compression.inp", "keywords": ["NSET", "ELSET"],
            "parallelism": "1 GPU" if n_gpus == 1 else
            "z-slabs: %d slabs of %d slices, one process per GPU; peer-memory / NCCL plane shifts + count and "
            "byte all-gathers, no data-path collective on the deck" % (n_gpus, Z),
            "l2": "inputs+outputs per step (1 GB volume, ~49 GB deck per GPU at the default shape) exceed the "
                  "126 MB L2; no flush"}


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


class CountingSink:
    """Host consumer of the streamed deck in the end-to-end leg: the bytes arrive in pinned host memory
    (ring of 3 x 64 MiB) and are counted; a real caller writes them to a file or a socket.  The
    deck of the default workload (~49 GB per GPU) does not fit pinned host memory N times over."""

    def __init__(self):
        self.n = 0
        self.check = 0

    def write(self, mv):
        self.n += len(mv)
        self.check ^= mv[0] ^ mv[-1]      # touch both ends of the chunk on the host

    def pwrite(self, mv, off):
        self.n += len(mv)
        if len(mv):
            self.check ^= mv[0] ^ mv[-1]


class HotPath:
    """The three reference calls on one GPU (the public single-GPU API) or, for N > 1, on this rank's z-slab
    through ciclope_b200.distributed (same signatures + the slab's placement)."""

    def __init__(self, world, rank, Zg, z0):
        self.world, self.rank, self.Zg, self.z0 = world, rank, Zg, z0
        import torch
        self.torch = torch
        from ciclope_b200.preprocess import mask_image
        self.mask_image = mask_image
        if world == 1:
            from ciclope_b200 import preprocess, voxelFE
            self.pre, self.vfe = preprocess, voxelFE
        else:
            from ciclope_b200 import distributed
            self.dist = distributed

    def ccl(self, bw):
        return self.pre.remove_unconnected(bw) if self.world == 1 else self.dist.remove_unconnected(bw)

    def grid(self, vol):
        if self.world == 1:
            return self.vfe.vol2ugrid(vol, VOXELSIZE)
        return self.dist.vol2ugrid(vol, VOXELSIZE, z0=self.z0, Z_global=self.Zg)

    def deck(self, mesh, **kw):
        """-> (deck object kept alive by the caller, bytes this rank formatted, bytes of the whole deck)"""
        if self.world == 1:
            d = self.vfe.deck_to_device(mesh, TEMPLATE, **kw)
            return d, d.numel(), d.numel()
        d = self.dist.deck_to_device(mesh, TEMPLATE, **kw)
        return d, d.local_bytes, d.total_bytes

    def out(self, mesh, sink, **kw):
        if self.world == 1:
            return self.vfe.mesh2voxelfe(mesh, TEMPLATE, sink, **kw)
        return self.dist.mesh2voxelfe(mesh, TEMPLATE, sink, **kw)

    def step(self, bw, grey=None, marks=None, **kw):
        """One pass: remove_unconnected -> (grey-scale: masked[~L] = 0, ciclope example 01) -> vol2ugrid ->
        deck in HBM.  `marks` collects CUDA events at the stage boundaries."""
        torch = self.torch

        def mark():
            if marks is not None:
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                marks.append(e)
        mark()
        L = self.ccl(bw)
        mark()
        vol = L if grey is None else self.mask_image(grey, L, makecopy=True)     # masked = I; masked[~L] = 0
        mesh = self.grid(vol)
        mark()
        deck, local_b, total_b = self.deck(mesh, **kw)
        mark()
        return mesh, deck, local_b, total_b

    def step_e2e(self, vol_host, vol_dev, sink):
        """Same path through the public API with HOST buffers: pinned volume in, deck streamed to the host."""
        vol_dev.copy_(vol_host, non_blocking=True)
        L = self.ccl(vol_dev)
        mesh = self.grid(L)
        self.out(mesh, sink)
        return mesh


def bind_to_gpu_numa_node(local_rank):
    """Run this rank on the CPUs next to its GPU (sysfs local_cpulist of the GPU's PCI function), so that
    the pinned host buffers of the end-to-end leg are allocated on that NUMA node.  Best effort."""
    try:
        import torch
        prop = torch.cuda.get_device_properties(local_rank)
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/local_cpulist" % (prop.pci_domain_id, prop.pci_bus_id,
                                                                         prop.pci_device_id)
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


def _sync_all(world):
    import torch
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()


def _max_over_ranks(x, world, dev):
    import torch
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def _sum_over_ranks(x, world, dev):
    import torch
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.item()


def time_steps(hp, world, dev, steps, warmup, bw, grey=None, **kw):
    """W warm-up steps, K timed steps (CUDA events, barrier + synchronize on both sides, max over ranks) with
    stage boundaries marked.  -> dict(ms_per_step, stage_ms, sizes ...)."""
    import torch
    for _ in range(warmup):
        mesh, deck, local_b, total_b = hp.step(bw, grey, None, **kw)
    n_el, n_nd = mesh.n_el, mesh.n_nd
    del mesh, deck
    _sync_all(world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    marks, wall = [], []
    e0.record()
    for _ in range(steps):
        t_s = time.perf_counter()
        mesh, deck, local_b, total_b = hp.step(bw, grey, marks, **kw)
        del mesh, deck
        wall.append((time.perf_counter() - t_s) * 1e3)
    e1.record()
    _sync_all(world)
    ms = _max_over_ranks(e0.elapsed_time(e1), world, dev) / steps
    stage = {"ccl": 0.0, "grid": 0.0, "emit": 0.0}
    for k in range(steps):
        m = marks[4 * k: 4 * k + 4]
        for name, a, b in (("ccl", 0, 1), ("grid", 1, 2), ("emit", 2, 3)):
            stage[name] += m[a].elapsed_time(m[b]) / steps
    stage = {k: _max_over_ranks(v, world, dev) for k, v in stage.items()}
    return dict(ms_per_step=ms, stage_ms=stage, n_el=n_el, n_nd=n_nd, local_bytes=int(local_b),
                total_bytes=int(total_b), wall=wall)


def stage_fractions(stage_ms, n_vox, deck_bytes, peak, grey_elems=0):
    """Fraction of the HBM roofline per stage on the ALGORITHMIC bytes of BASELINE.md section 4 (per GPU):
    A_ccl = 2 N_vox, A_grid = N_vox (+ A_bin = N_vox + 8 n_el for grey-scale volumes), A_emit = N_vox + |INP|."""
    alg = {"ccl": 2 * n_vox, "grid": n_vox + ((n_vox + 8 * grey_elems) if grey_elems else 0),
           "emit": n_vox + deck_bytes}
    return {k: {"ms": round(stage_ms[k], 4), "algorithmic_bytes": int(alg[k]),
                "frac": round(alg[k] / (stage_ms[k] * 1e-3) / 1e9 / peak, 4) if stage_ms[k] > 0 else None}
            for k in alg}


# ---- parity inside the run -------------------------------------------------------------------------

def parity_check(world, rank, dev):
    """Small volumes through the code path the timed region uses, compared with the single-GPU deck of the
    same global volume section by section (SHA-256).  N > 1: real ranks (ciclope_b200.distributed:
    peer-memory exchange + NCCL, the pinned-ring writer, one shared file); N = 1: four emulated slabs."""
    import tempfile

    import torch

    from ciclope_b200 import preprocess, sections, synth, voxelFE
    cases = []
    shm = "/dev/shm" if os.path.isdir("/dev/shm") else None
    n_slabs = world if world > 1 else 4
    for (Z, Y, X), grey in (((8 * n_slabs, 48, 96), False), ((64 * n_slabs, 192, 192), False),
                            ((16 * n_slabs, 64, 96), True)):
        kw = dict(keywords=['NSET', 'ELSET', 'PROPERTY'],
                  matprop={"file": [MAT_LINEAR], "range": [[1, 255]]}) if grey else {}
        g = synth.greyscale((Z, Y, X), 5, 0.30, dev) if grey else None
        bw = (g > 0) if grey else synth.trabecular((Z, Y, X), 5, 0.30, dev)
        # single GPU (rank 0 compares; every rank computes its own copy, the volumes are small)
        L = preprocess.remove_unconnected(bw)
        vol = torch.where(L, g, torch.zeros((), dtype=torch.uint8, device=dev)) if grey else L
        single = bytes(voxelFE.deck_to_device(voxelFE.vol2ugrid(vol, VOXELSIZE), TEMPLATE,
                                              **{k: (dict(v) if isinstance(v, dict) else v)
                                                 for k, v in kw.items()}).cpu().numpy())
        path = None
        if world > 1:
            import torch.distributed as dist

            from ciclope_b200 import distributed, sharded
            a, b = sharded.slab_ranges(Z, world)[rank]
            names = [os.path.join(tempfile.mkdtemp(dir=shm), "parity.inp") if rank == 0 else None]
            dist.broadcast_object_list(names, src=0)
            path = names[0]
            Ls = distributed.remove_unconnected(bw[a:b].contiguous())
            vs = torch.where(Ls, g[a:b], torch.zeros((), dtype=torch.uint8, device=dev)) if grey else Ls
            mesh = distributed.vol2ugrid(vs, VOXELSIZE)            # placement from the all-gathered heights
            distributed.mesh2voxelfe(mesh, TEMPLATE, path, **kw)
            mask_equal = bool(torch.equal(Ls, L[a:b]))
            sharded_deck = open(path, "rb").read() if rank == 0 else None
        else:
            from ciclope_b200 import sharded
            mask, _ = sharded.emulate_remove_unconnected(bw, n_slabs)
            mask_equal = bool(torch.equal(mask, L))
            with tempfile.TemporaryDirectory(dir=shm) as d:
                path = os.path.join(d, "parity.inp")
                sharded.emulate_ct2inp(vol, n_slabs, VOXELSIZE, TEMPLATE, path, m_kwargs=kw)
                sharded_deck = open(path, "rb").read()
            path = None
        rec = {"global_volume": [Z, Y, X], "greyscale": grey, "slabs": n_slabs}
        if rank == 0:
            ok, per = sections.compare_decks(single, sharded_deck)
            rec.update(deck_bytes=len(single), sections=per, sections_equal=bool(ok),
                       sha256=sections.section_hashes(single)["elements"]["sha256"][:16])
            if path:
                os.remove(path)
                os.rmdir(os.path.dirname(path))
        rec["mask_equal_on_all_ranks"] = _sum_over_ranks(0.0 if mask_equal else 1.0, world, dev) == 0.0
        cases.append(rec)
    ok = all(c.get("sections_equal", True) and c["mask_equal_on_all_ranks"] for c in cases)
    bad = _sum_over_ranks(0.0 if ok else 1.0, world, dev)
    return {"ranks": world, "path": "ciclope_b200.distributed over real ranks vs single GPU" if world > 1 else
            "4 emulated z-slabs (sharded.LocalRunner) vs single GPU",
            "sections_equal": bad == 0.0, "cases": cases}


# ---- the other BASELINE configs ---------------------------------------------------------------------

def foam_components(hp, world, rank, dev, foam):
    """Component count of the (global) foam and the number of components cut by slab interfaces, from the
    union-find arrays of one extra labelling (diagnostics, not timed)."""
    import torch
    if world == 1:
        from ciclope_b200 import preprocess
        info = {}
        preprocess.largest_component_device(foam.view(torch.uint8), info=info)
        n = info["n_runs"]
        roots = int((info["parent"][:n].to(torch.int64) == torch.arange(n, device=dev)).sum().item())
        return {"components": roots, "cut_by_interfaces": 0}
    from ciclope_b200 import distributed, sharded
    info = {}
    r = distributed.runner()
    r.run(sharded.ccl_program(r.ctx(), foam.view(torch.uint8), info=info))
    n = int(info["stats"][0].item())
    local_roots = int((info["parent"][:n].to(torch.int64) == torch.arange(n, device=dev)).sum().item())
    total_local = _sum_over_ranks(float(local_roots), world, dev)
    if info["host_merge"]:
        return {"components": None, "note": "boundary graph above the device capacity: host merge path",
                "local_components_sum": int(total_local)}
    P, Cv = info["packs"], info["Cv"]
    nv = P[:, 0].to(torch.int64)
    lab = info["label"].to(torch.int64).reshape(world, Cv)
    valid = torch.arange(Cv, device=dev).unsqueeze(0) < nv.unsqueeze(1)
    labs = lab[valid]
    ranks = torch.arange(world, device=dev).unsqueeze(1).expand(world, Cv)[valid]
    uniq, inv = torch.unique(labs, return_inverse=True)
    lo = torch.full((uniq.numel(),), world, device=dev).scatter_reduce(0, inv, ranks, reduce="amin")
    hi = torch.full((uniq.numel(),), -1, device=dev).scatter_reduce(0, inv, ranks, reduce="amax")
    return {"components": int(total_local) - int(nv.sum().item()) + int(uniq.numel()),
            "cut_by_interfaces": int((hi > lo).sum().item()),
            "boundary_vertices_per_rank": [int(v) for v in nv.tolist()],
            "boundary_edges_per_rank": [int(v) for v in P[:, 1].tolist()],
            "device_merge_capacity": [int(Cv), int(info["Ce"])]}


def run_other_configs(hp, world, rank, dev, shape, peak, steps):
    import torch

    from ciclope_b200 import sharded, synth
    out = {}
    Z, Y, X = shape
    # ---- configs[3]: fragmented foam of the workload's shape
    Zg = Z * world
    foam = synth.foam_slab((Zg, Y, X), rank * Z, rank * Z + Z, seed=0, bvtv=0.12, device=dev)
    n_vox = foam.numel()
    comp = foam_components(hp, world, rank, dev, foam)
    r = time_steps(hp, world, dev, steps, 2, foam)
    fg = _sum_over_ranks(float(foam.sum().item()), world, dev)
    out["configs[3] foam"] = {
        "workload": "fragmented foam (BV/TV 0.12, below percolation), global volume %dx%dx%d in %d z-slab(s)"
                    % (Zg, Y, X, world),
        "foreground_voxels": int(fg), "kept_voxels": int(_sum_over_ranks(float(r["n_el"]), world, dev)),
        **comp,
        "ccl_stage_ms": round(r["stage_ms"]["ccl"], 4),
        "ccl_voxels_per_s": n_vox * world / (r["stage_ms"]["ccl"] * 1e-3),
        "ms_per_step": round(r["ms_per_step"], 4), "voxels_per_s": n_vox * world / (r["ms_per_step"] * 1e-3),
        "steps": steps, "deck_bytes": r["total_bytes"],
        "stages": stage_fractions(r["stage_ms"], n_vox, r["local_bytes"], peak)}
    del foam
    torch.cuda.empty_cache()
    # ---- configs[2]: 1024^3 uint8 grey-scale volume, GV -> material mapping (255 ELSETs + PROPERTY cards)
    Zg, Y, X = GREY_GLOBAL
    if Zg >= world:
        a, b = sharded.slab_ranges(Zg, world)[rank]
        grey = synth.greyscale_slab(GREY_GLOBAL, a, b, seed=0, bvtv=0.30, device=dev)
        bw = grey > 0                                  # segment(I, 0): BW = I > T
        ghp = HotPath(world, rank, Zg, a)
        kw = dict(keywords=['NSET', 'ELSET', 'PROPERTY'], matprop={"file": [MAT_LINEAR], "range": [[1, 255]]})
        r = time_steps(ghp, world, dev, steps, 2, bw, grey, **kw)
        n_vox_g = Zg * Y * X
        out["configs[2] greyscale 1024^3"] = {
            "workload": "1024^3 uint8 grey-scale volume (GV 1..255 inside the bone mask), %d z-slab(s): "
                        "remove_unconnected(I > 0) + masked[~L] = 0 + vol2ugrid + mesh2voxelfe(matprop, "
                        "keywords NSET ELSET PROPERTY)" % world,
            "scaling": "strong", "material_card": "tests/golden/mat_linear.inp (the law of bone_Elinear.inp)",
            "elements": int(_sum_over_ranks(float(r["n_el"]), world, dev)), "deck_bytes": r["total_bytes"],
            "ms_per_step": round(r["ms_per_step"], 4), "voxels_per_s": n_vox_g / (r["ms_per_step"] * 1e-3),
            "steps": steps,
            "stages": stage_fractions(r["stage_ms"], grey.numel(), r["local_bytes"], peak, grey_elems=r["n_el"])}
        del grey, bw
        torch.cuda.empty_cache()
    return out


def run_gpu(args, rank, world, local_rank):
    import torch

    from ciclope_b200 import _lib, synth

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else "not bound (1 GPU)"
    if rank == 0:
        sys.stderr.write("[bench] host affinity: %s\n" % numa)
    lib = _lib.require_cuda()
    Z, Y, X = args.shape
    Zg, z0 = Z * world, Z * rank
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s"

    # ---- parity first: a wrong deck must not be timed
    parity = None
    if not args.no_parity:
        parity = parity_check(world, rank, dev)
        if rank == 0:
            sys.stderr.write("[bench] parity_check: %s\n" % json.dumps(parity))
        if not parity["sections_equal"]:
            if rank == 0:
                emit_line({"metric": METRIC, "value": None, "error": "parity check failed", "parity_check": parity})
            raise SystemExit(3)

    hp = HotPath(world, rank, Zg, z0)
    vol = synth.trabecular_slab((Zg, Y, X), z0, z0 + Z, seed=0, bvtv=0.30, device=dev)
    torch.cuda.empty_cache()
    n_vox = vol.numel()

    # ---- warm-up + timed region: K steps, CUDA events, max over ranks
    n_warm = max(args.warmup, 3)
    sampler = ClockSampler(local_rank)
    for _ in range(n_warm):                 # (also sizes the caching allocator and NCCL's channels)
        hp.step(vol)
    _sync_all(world)
    sampler.start()
    l0 = lib.cfe_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    marks, step_wall = [], []
    e0.record()
    for _ in range(args.steps):
        t_s = time.perf_counter()
        mesh, deck, local_b, total_b = hp.step(vol, None, marks)   # every step ends with the host read of its totals
        n_el, n_nd = mesh.n_el, mesh.n_nd
        del mesh, deck
        step_wall.append((time.perf_counter() - t_s) * 1e3)
    e1.record()
    _sync_all(world)
    launches = lib.cfe_launch_count() - l0
    clocks = sampler.stop()
    ms_per_step = _max_over_ranks(e0.elapsed_time(e1), world, dev) / args.steps
    value = n_vox * world / (ms_per_step * 1e-3)
    stage_ms = {"ccl": 0.0, "grid": 0.0, "emit": 0.0}
    for k in range(args.steps):
        m = marks[4 * k: 4 * k + 4]
        for name, a, b in (("ccl", 0, 1), ("grid", 1, 2), ("emit", 2, 3)):
            stage_ms[name] += m[a].elapsed_time(m[b]) / args.steps
    stage_ms = {k: _max_over_ranks(v, world, dev) for k, v in stage_ms.items()}
    deck_bytes = int(local_b)

    # ---- per-entry-point device times (separate profiled steps, CUDA events on the launch stream)
    _lib.PROFILE = []
    prof_steps = 3
    for _ in range(prof_steps):
        hp.step(vol)
    torch.cuda.synchronize()
    entry_ms = {}
    for name, a, b in _lib.PROFILE:
        entry_ms[name] = entry_ms.get(name, 0.0) + a.elapsed_time(b) / prof_steps
    _lib.PROFILE = None

    # ---- end to end through the public API with host buffers (deck streamed through the pinned ring)
    e2e = None
    if not args.no_e2e:
        vol_host = torch.empty((Z, Y, X), dtype=torch.bool).pin_memory()
        vol_host.copy_(vol)
        vol_dev = torch.empty_like(vol)
        sink = CountingSink()
        hp.step_e2e(vol_host, vol_dev, sink)
        _sync_all(world)
        k_e2e = args.e2e_steps
        sink.n = 0
        t0 = time.perf_counter()
        for _ in range(k_e2e):
            hp.step_e2e(vol_host, vol_dev, sink)
        _sync_all(world)
        dt = _max_over_ranks((time.perf_counter() - t0) / k_e2e, world, dev)
        d2h = sink.n // k_e2e
        e2e = {"value": n_vox * world / dt, "unit": "voxels/s", "h2d_bytes_per_step": int(n_vox),
               "d2h_bytes_per_step": int(d2h) + 16 + 136 + 16, "ms_per_step": dt * 1e3, "steps": k_e2e,
               "sink": "pinned ring (3 x 64 MiB), bytes counted on the host; per rank for N > 1",
               "d2h_gbs_per_gpu": d2h / dt / 1e9}
        del vol_host, vol_dev

    # ---- the other BASELINE configs (grey-scale 1024^3, foam)
    others = None
    if not args.no_configs:
        del vol
        torch.cuda.empty_cache()
        others = run_other_configs(hp, world, rank, dev, args.shape, peak, args.config_steps)

    if rank != 0:
        return
    # dominant kernel: k_emit_elements2 (entry point cfe_emit_elements = that one launch).  Algorithmic
    # bytes per launch = element-section bytes written + 6 B/element read (voxel index, line length,
    # digit flags) -- DESIGN.md section 4.
    elem_bytes = deck_bytes - 53 * n_nd   # everything but the *NODE lines is element lines + O(N^2) text
    dom = max(entry_ms, key=entry_ms.get)
    # *NODE emitter: 53 B per line written + the used-node bitmap and its prefix (8 B per 32 grid nodes) read
    nodes_key = "cfe_emit_nodes_bits" if "cfe_emit_nodes_bits" in entry_ms else "cfe_emit_nodes"
    alg = {"cfe_emit_elements": elem_bytes + 6 * n_el,
           "cfe_emit_nodes": 53 * n_nd + ((Z + 1) * (Y + 1) * (X + 1)) // 4 if nodes_key.endswith("bits") else 57 * n_nd}
    achieved = alg["cfe_emit_elements"] / (entry_ms["cfe_emit_elements"] * 1e-3) / 1e9
    traffic, traffic_src = None, None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        ent = tr.get("%dx%dx%d" % (Z, Y, X))
        if ent:
            traffic, traffic_src = ent.get("k_emit_elements2"), ent.get("source")
    except Exception:  # noqa: BLE001
        pass
    nodes_ach = alg["cfe_emit_nodes"] / (entry_ms[nodes_key] * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "k_emit_elements2 (entry point cfe_emit_elements, one launch per deck)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "traffic_source": traffic_src, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg["cfe_emit_elements"],
                "ms_per_launch": entry_ms["cfe_emit_elements"], "dominant_entry_point": dom,
                "other_kernels": {"k_emit_nodes4b" if nodes_key.endswith("bits") else "k_emit_nodes4": {
                    "achieved": nodes_ach, "frac": nodes_ach / peak,
                    "algorithmic_bytes_per_launch": alg["cfe_emit_nodes"], "ms_per_launch": entry_ms[nodes_key]}}}
    total_alg = 2 * n_vox + deck_bytes
    cpu = None
    if world == 1 and not args.no_cpu:
        _v, _dt, _w, _t = cpu_pool_run(CPU_SAMPLE)
        cpu = cpu_baseline_obj(CPU_SAMPLE, _v, _w, _t)
    line = {"metric": METRIC, "value": value, "unit": "voxels/s", "n_gpus": world, "steps": args.steps,
            "warmup": n_warm, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": workload_config(args.shape, world),
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
            "cpu_baseline": cpu, "parity_check": parity,
            "job": {"voxels": n_vox * world, "elements_rank0": n_el, "used_nodes_rank0": n_nd,
                    "deck_bytes_rank0": deck_bytes, "deck_bytes_total": int(total_b),
                    "algorithmic_bytes_rank0": total_alg,
                    "whole_path_frac_of_hbm_peak": total_alg / (ms_per_step * 1e-3) / 1e9 / peak},
            "stages": stage_fractions(stage_ms, n_vox, deck_bytes, peak),
            "configs": others,
            "step_ms": {"min": round(min(step_wall), 3), "median": round(statistics.median(step_wall), 3),
                        "max": round(max(step_wall), 3)},
            "entry_point_ms": {k: round(v, 4) for k, v in sorted(entry_ms.items(), key=lambda kv: -kv[1])}}
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


def _shape(text):
    v = tuple(int(p) for p in text.replace("x", ",").split(","))
    if len(v) == 1:
        v = v * 3
    if len(v) != 3 or min(v) < 1:
        raise argparse.ArgumentTypeError("shape must be Z,Y,X")
    return v


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="timed steps (default 20; reference arm 3)")
    ap.add_argument("--warmup", type=int, default=None, help="warm-up steps (default 3; reference arm 1)")
    ap.add_argument("--impl", default="ciclope_b200", choices=["ciclope_b200", "reference"])
    ap.add_argument("--shape", type=_shape, default=DEFAULT_SHAPE, help="volume per GPU as Z,Y,X (or one edge)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--config-steps", type=int, default=5)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-configs", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        args.steps = 3 if args.steps is None else args.steps
        args.warmup = 1 if args.warmup is None else args.warmup
        run_reference(args, rank)
        return
    args.steps = 20 if args.steps is None else args.steps
    args.warmup = 3 if args.warmup is None else args.warmup
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

            from ciclope_b200 import distributed
            distributed.shutdown()
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
