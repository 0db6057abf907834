"""Z-slab sharding of the CT -> .INP hot path over the GPUs of one box.

The global volume [Z,Y,X] is cut into contiguous z-slabs, one per rank (one process per GPU under
torchrun).  Every voxel / node / text loop is the same CUDA kernel the single-GPU path uses; the
kernels take the slab's global slice offset ``z0``, the global element-id offset and the global
position of every set entry, so their output bytes are the bytes of the single-GPU deck.

Exchange steps (NCCL over NVLink; all latency-bound, O(plane) or O(1) sized):

* ``shift_up``     one voxel plane of packed bits (Y*W words) to the rank above: node flags /
                   plane-set membership on a slab's first node layer, and the CCL interface;
* ``all_gather``   per-slab element / node / set counts and byte totals -> global numbering and
                   file offsets; CCL boundary vertices, sizes and cross-slab edges;
* the CCL winner is the max over (merged size, min global run id), computed identically on every
  rank from the gathered boundary graph.

Each rank's logic is written ONCE as a generator that yields its collective requests; the same
program runs under ``torch.distributed`` (``DistRunner``) or as N emulated ranks inside one
process on one GPU (``LocalRunner``), which is how the single-GPU test box exercises N = 2..8.
"""
import os

import numpy as np
import torch

from . import _lib
from . import voxelFE as vfe
from ._lib import CfeItemSeg, CfeSetDesc
from ._util import bytes_to_device_async, empty_u32, read_back, scan_workspace, tag_packed, to_device_async


# --------------------------------------------------------------------------------------------
# slab geometry (pure host logic)

def slab_ranges(Z, n):
    """Contiguous z ranges [z0, z1) of n slabs; the first Z % n slabs get one extra slice."""
    if n < 1:
        raise ValueError("need at least one slab")
    if Z < n:
        raise ValueError("cannot cut %d slices into %d non-empty slabs" % (Z, n))
    base, extra = divmod(Z, n)
    out, z = [], 0
    for k in range(n):
        h = base + (1 if k < extra else 0)
        out.append((z, z + h))
        z += h
    return out


def exclusive_offsets(counts):
    """[sum(counts[:k]) for k] and the total."""
    out, acc = [], 0
    for c in counts:
        out.append(acc)
        acc += int(c)
    return out, acc


def clip_box_z(box, lo, hi):
    (s0, s1), r, c = box
    s0, s1 = max(s0, lo), min(s1, hi)
    if s1 <= s0:
        return ((0, 0), (0, 0), (0, 0))
    return ((s0, s1), r, c)


# --------------------------------------------------------------------------------------------
# collectives as generator requests

class Ctx:
    """Per-rank context handed to the rank programs."""

    def __init__(self, rank, size, device):
        self.rank, self.size, self.device = rank, size, device

    def all_gather(self, t):
        """same-shape tensor from every rank -> list of tensors (index = rank)"""
        return (yield ("all_gather", t))

    def all_gather_var(self, t):
        """1-D tensors of different lengths -> list of tensors"""
        return (yield ("all_gather_var", t))

    def shift_up(self, t):
        """send `t` to rank+1; returns the tensor of rank-1 (None on rank 0).  Same shape on all ranks."""
        return (yield ("shift_up", t))

    def barrier(self):
        return (yield ("barrier", None))


class LocalRunner:
    """Runs N rank programs in lock-step inside one process (all on the current device)."""

    def __init__(self, size, device=None):
        self.size = size
        self.device = device

    def ctxs(self):
        return [Ctx(k, self.size, self.device) for k in range(self.size)]

    def run(self, programs):
        gens = list(programs)
        results = [None] * len(gens)
        reqs = []
        for k, g in enumerate(gens):
            try:
                reqs.append(next(g))
            except StopIteration as e:
                results[k] = e.value
                reqs.append(None)
        while any(r is not None for r in reqs):
            if any(r is None for r in reqs):
                raise RuntimeError("rank programs diverged (some finished while others wait on a collective)")
            kinds = {r[0] for r in reqs}
            if len(kinds) != 1:
                raise RuntimeError("rank programs issued different collectives: %s" % kinds)
            kind = kinds.pop()
            vals = [r[1] for r in reqs]
            if kind in ("all_gather", "all_gather_var"):
                answers = [[v.clone() for v in vals] for _ in gens]
            elif kind == "shift_up":
                answers = [None] + [vals[k - 1].clone() for k in range(1, len(gens))]
            elif kind == "barrier":
                answers = [None] * len(gens)
            else:
                raise RuntimeError("unknown collective %s" % kind)
            new = []
            for k, g in enumerate(gens):
                try:
                    new.append(g.send(answers[k]))
                except StopIteration as e:
                    results[k] = e.value
                    new.append(None)
            reqs = new
        return results


class PeerExchange:
    """The small collectives of the z-slab path (counts, one packed bit plane, boundary vertices) as ONE
    kernel over NVLink peer memory (``cfe_peer_exchange``, csrc/peer.cu) instead of an NCCL call each.
    torch.distributed._symmetric_memory allocates the buffer and exchanges the peer pointers once; the
    data path is the library's kernel: remote stores into every peer's slot, a release / acquire signal per
    peer, then a local gather.  Payloads above ``CHUNK`` bytes stay with NCCL."""

    SLOTS = 4
    CHUNK = 16 << 20         # bytes per source rank and slot (a 2048^2 interface: 0.5 MB of bits + 8 MB of run roots)
    SIG_OFFSET = 256         # first u32 of the signal pad used here (the front is symmetric memory's own)

    def __init__(self, dist, group, device):
        import torch.distributed._symmetric_memory as symm
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        if self.world > 32:
            raise RuntimeError("peer exchange handles up to 32 ranks")
        pg = group if group is not None else dist.group.WORLD
        with torch.cuda.device(device):
            self.buf = symm.empty(self.SLOTS * self.world * self.CHUNK, dtype=torch.uint8, device=device)
            self.hdl = symm.rendezvous(self.buf, pg)
            if self.hdl.signal_pad_size < 4 * (self.SIG_OFFSET + self.SLOTS * self.world):
                raise RuntimeError("signal pad too small")
            self.counter = torch.zeros(2, dtype=torch.int32, device=device)
            self.hdl.barrier()                  # every rank's pad and counter are ready before the first kernel
        self.device = device
        self.epoch = 0
        self.shifts_in_a_row = 0

    def exchange(self, t, shift=False):
        """-> tensor [world, *t.shape].  All-gather: row p = rank p's ``t``.  ``shift``: ``t`` goes to the rank
        above only and only row rank - 1 is defined; neighbours alone synchronise, except that every
        (SLOTS - 2)-th shift in a row is run as a full synchronisation (see csrc/peer.cu)."""
        t = t.contiguous()
        nbytes = t.numel() * t.element_size()
        self.epoch += 1
        slot = self.epoch % self.SLOTS
        out = torch.empty((self.world,) + tuple(t.shape), dtype=t.dtype, device=t.device)
        full = (1 << self.world) - 1
        if shift and self.shifts_in_a_row < self.SLOTS - 2:
            self.shifts_in_a_row += 1
            send = (1 << (self.rank + 1)) if self.rank + 1 < self.world else 0
            wait = (1 << (self.rank - 1)) if self.rank > 0 else 0
        else:
            self.shifts_in_a_row = 0
            send = wait = full
        _lib.call("cfe_peer_exchange", t.data_ptr(), nbytes, self.hdl.buffer_ptrs_dev, self.hdl.signal_pad_ptrs_dev,
                  self.rank, self.world, slot * self.world * self.CHUNK, self.CHUNK, self.epoch, send, wait,
                  self.SIG_OFFSET + slot * self.world, self.counter.data_ptr(), out.data_ptr(), _lib.stream())
        return out


class DistRunner:
    """Runs one rank program per process over torch.distributed (NCCL on GPUs, gloo on CPU).  On GPUs the
    small exchanges go through :class:`PeerExchange` (one kernel over NVLink peer memory); NCCL carries the
    large planes, and everything when symmetric memory cannot be set up (``CICLOPE_B200_PEER=0`` forces
    that)."""

    def __init__(self, group=None, device=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.size = dist.get_world_size(group)
        self.device = device
        self.peer = None
        self.peer_error = None
        if device is None and dist.get_backend(group) == "nccl" and torch.cuda.is_available():
            device = torch.device("cuda", torch.cuda.current_device())      # one process per GPU
        if (device is not None and torch.device(device).type == "cuda" and self.size > 1
                and dist.get_backend(group) == "nccl" and os.environ.get("CICLOPE_B200_PEER", "1") != "0"):
            try:
                self.peer = PeerExchange(dist, group, torch.device(device))
            except Exception as e:  # noqa: BLE001 -- no P2P / symmetric memory on this box: NCCL does it all
                self.peer_error = "%s: %s" % (type(e).__name__, e)
            # all ranks take the same path
            ok = torch.tensor([1 if self.peer is not None else 0], dtype=torch.int32, device=device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
            if int(ok.item()) == 0:
                self.peer = None

    def ctx(self):
        return Ctx(self.rank, self.size, self.device)

    def _gather(self, t):
        """all_gather_into_tensor: one flat output, far less Python / launch overhead than the list API"""
        t = t.contiguous()
        if self.peer is not None and t.is_cuda and 0 < t.numel() * t.element_size() <= PeerExchange.CHUNK:
            return self.peer.exchange(t)
        flat = t.reshape(-1)
        out = torch.empty(self.size * flat.numel(), dtype=t.dtype, device=t.device)
        self.dist.all_gather_into_tensor(out, flat, group=self.group)
        return out.reshape((self.size,) + tuple(t.shape))

    def _fulfil(self, kind, t):
        dist = self.dist
        if kind == "all_gather":
            return list(self._gather(t).unbind(0))
        if kind == "all_gather_var":
            n = torch.tensor([t.numel()], dtype=torch.int64, device=t.device)
            ns = [int(v) for v in self._gather(n).reshape(-1).tolist()]
            m = max(max(ns), 1)
            pad = torch.zeros(m, dtype=t.dtype, device=t.device)
            pad[: t.numel()] = t.reshape(-1)
            out = self._gather(pad)
            return [out[k, :c] for k, c in enumerate(ns)]
        if kind == "shift_up":
            if self.peer is not None and t.is_cuda and 0 < t.numel() * t.element_size() <= PeerExchange.CHUNK:
                # the plane goes to the rank above only, and only neighbours wait for each other
                out = self.peer.exchange(t, shift=True)
                return out[self.rank - 1] if self.rank > 0 else None
            if t.numel() * t.element_size() <= self.SHIFT_VIA_GATHER_BYTES:
                # small planes: one all-gather is cheaper to issue than a send/recv pair
                out = self._gather(t)
                return out[self.rank - 1] if self.rank > 0 else None
            ops, recv = [], None
            if self.rank + 1 < self.size:
                ops.append(dist.P2POp(dist.isend, t.contiguous(), self.rank + 1, group=self.group))
            if self.rank > 0:
                recv = torch.empty_like(t)
                ops.append(dist.P2POp(dist.irecv, recv, self.rank - 1, group=self.group))
            if ops:
                for w in dist.batch_isend_irecv(ops):
                    w.wait()
            return recv
        if kind == "barrier":
            dist.barrier(group=self.group)
            return None
        raise RuntimeError("unknown collective %s" % kind)

    SHIFT_VIA_GATHER_BYTES = 4 << 20

    def run(self, program):
        try:
            req = next(program)
            while True:
                req = program.send(self._fulfil(*req))
        except StopIteration as e:
            return e.value


# --------------------------------------------------------------------------------------------
# cross-slab component merge (device-agnostic torch ops on O(boundary) data)

def solve_boundary_components(keys, sizes, edges):
    """Connected components of the boundary graph.

    keys  : int64 [V] sorted unique vertex keys (rank << 32 | slab-local root run id)
    sizes : int64 [V] voxels of each slab-local component
    edges : int64 [E,2] vertex keys of overlapping runs across slab interfaces
    Returns (label [V] = index of the smallest-key vertex of the component, comp_size [V] (valid at
    label positions)).  On CUDA tensors the components come from the library's union-find kernels
    (cfe_uf_solve); on CPU tensors (unit tests, gloo) from label propagation with pointer jumping."""
    V = keys.numel()
    lab = torch.arange(V, dtype=torch.int64, device=keys.device)
    if V == 0:
        return lab, sizes.clone()
    if edges.numel():
        a = torch.searchsorted(keys, edges[:, 0].contiguous())
        b = torch.searchsorted(keys, edges[:, 1].contiguous())
        if keys.is_cuda:
            lab32 = torch.empty(V, dtype=torch.int32, device=keys.device)
            with torch.cuda.device(keys.device):
                _lib.call("cfe_uf_solve", lab32.data_ptr(), V, a.data_ptr(), b.data_ptr(), a.numel(), _lib.stream())
            lab = lab32.to(torch.int64)
        else:
            while True:
                m = torch.minimum(lab[a], lab[b])
                new = lab.clone()
                new.scatter_reduce_(0, a, m, reduce="amin")
                new.scatter_reduce_(0, b, m, reduce="amin")
                for _ in range(3):
                    new = new[new]
                if torch.equal(new, lab):
                    break
                lab = new
            while True:           # full compression
                nxt = lab[lab]
                if torch.equal(nxt, lab):
                    break
                lab = nxt
    comp = torch.zeros(V, dtype=torch.int64, device=keys.device)
    comp.scatter_add_(0, lab, sizes)
    return lab, comp


def pick_winner(cands):
    """cands: iterable of (size, key); largest size, ties -> smallest key.  None if empty/zero."""
    best = None
    for size, key in cands:
        if size <= 0:
            continue
        if best is None or size > best[0] or (size == best[0] and key < best[1]):
            best = (size, key)
    return best


# --------------------------------------------------------------------------------------------
# rank programs

def _u32_to_i64(t):
    return t.to(torch.int64) & 0xFFFFFFFF


BOUNDARY_VERTS = 1 << 16      # per-rank capacity of the device-side cross-slab merge (vertices = slab-local
                              # components touching an interface); larger cases take the host-driven path
BOUNDARY_EDGES = 2 * BOUNDARY_VERTS   # spanning-forest edges of one interface: never more than 2 Cv - 1
MSG_RUNS_DIV = 16             # runs of an interface plane carried by the device-side merge: Y * X / 16 ...
MSG_RUNS_MIN = 4096           # ... but at least this many
_MARK = -1                    # 0xffffffff in the u32 size array = "keep this root"


def ccl_program(ctx, vol, force_host_merge=False, info=None):
    """remove_unconnected of the global volume; `vol` = this rank's uint8 slab [Zl,Y,X] on the GPU.
    Returns (uint8 mask slab, winner (size, key) or None).

    The cross-slab merge runs entirely on the device with fixed-capacity buffers (one plane shift,
    one all-gather, ONE host read at the end); when a rank has more boundary components / edges
    than the capacity, the flag in that read sends every rank through `_ccl_host_merge`."""
    lib = _lib.require_cuda()
    dev = vol.device
    Zl, Y, X = (int(v) for v in vol.shape)
    W = (X + 31) // 32
    rows = Zl * Y
    PW = Y * W
    N, rank = ctx.size, ctx.rank
    with torch.cuda.device(dev):
        s = _lib.stream()
        bits = empty_u32(rows * W, dev)
        pref = empty_u32(rows * W, dev)
        stats = torch.zeros(4, dtype=torch.int64, device=dev)   # n_runs, best
        ws = scan_workspace(max(rows * W, PW), dev)
        _lib.call("cfe_pack_bits", vol.data_ptr(), rows, X, 0, bits.data_ptr(), s)
        _lib.call("cfe_scan_run_starts", bits.data_ptr(), rows, X, pref.data_ptr(), stats.data_ptr(),
                  ws.data_ptr(), s)
        max_runs = lib.cfe_ccl_max_runs(rows, X)
        parent = empty_u32(max_runs, dev)
        size = empty_u32(max_runs, dev)
        _lib.call("cfe_ccl_label", bits.data_ptr(), pref.data_ptr(), stats.data_ptr(), Zl, Y, X,
                  parent.data_ptr(), size.data_ptr(), s)
        out = torch.empty((Zl, Y, X), dtype=torch.uint8, device=dev)
        if N == 1:
            _lib.call("cfe_ccl_select", parent.data_ptr(), size.data_ptr(), stats.data_ptr(), stats.data_ptr() + 8, s)
            _lib.call("cfe_ccl_write_mask", bits.data_ptr(), pref.data_ptr(), Zl, Y, X, parent.data_ptr(),
                      stats.data_ptr() + 8, None, out.data_ptr(), s)
            best = int(stats[1].item()) & 0xFFFFFFFFFFFFFFFF
            return (out, None) if best == 0 else (out, (best >> 32, 0xFFFFFFFF - (best & 0xFFFFFFFF)))

        # ---- roots of the runs of my bottom (0) and top (Zl-1) voxel planes
        cap = Y * ((X + 1) // 2)
        roots_bot = empty_u32(cap, dev)
        roots_top = empty_u32(cap, dev)
        cnt = torch.zeros(4, dtype=torch.int32, device=dev)
        _lib.call("cfe_ccl_plane_roots", bits.data_ptr(), pref.data_ptr(), 0, Y, X, parent.data_ptr(),
                  roots_bot.data_ptr(), cnt.data_ptr(), s)
        _lib.call("cfe_ccl_plane_roots", bits.data_ptr(), pref.data_ptr(), Zl - 1, Y, X, parent.data_ptr(),
                  roots_top.data_ptr(), cnt.data_ptr() + 4, s)
        top_bits = bits[(Zl - 1) * PW: Zl * PW]
        state = dict(bits=bits, pref=pref, parent=parent, size=size, stats=stats, cnt=cnt, out=out, ws=ws,
                     roots_bot=roots_bot, roots_top=roots_top, top_bits=top_bits, cap=cap, max_runs=max_runs,
                     dims=(Zl, Y, X))
    if force_host_merge:
        return (yield from _ccl_host_merge(ctx, state))

    Cv, Ce = BOUNDARY_VERTS, BOUNDARY_EDGES
    PS = 4 + 2 * Cv + 2 * Ce
    with torch.cuda.device(dev):
        s = _lib.stream()
        pack = torch.empty(PS, dtype=torch.int32, device=dev)
        pack[:4].zero_()
        vfirst = empty_u32(max_runs, dev)
        didx = empty_u32(2 * cap, dev)
        vbot = empty_u32(cap, dev)
        vtop = torch.zeros(cap, dtype=torch.int32, device=dev)
        # dense numbering of my boundary components, their roots / sizes, the per-run vertex indices
        # and the best slab-internal component -- three launches, nothing read back
        _lib.call("cfe_ccl_boundary_vertices", roots_bot.data_ptr(), roots_top.data_ptr(), cnt.data_ptr(),
                  1 if rank > 0 else 0, 1 if rank + 1 < N else 0, stats.data_ptr(), vfirst.data_ptr(),
                  didx.data_ptr(), parent.data_ptr(), size.data_ptr(), Cv, pack.data_ptr(), vbot.data_ptr(),
                  vtop.data_ptr(), s)
    # the interface message: the top bit plane + the vertex of each of its runs.  A plane can hold Y * X / 2 runs,
    # a real one holds a few per cent of that: the message carries the first MSG_RUNS of them and a plane
    # with more runs sends the step through the host-driven merge (flagged by the receiver's edge pass)
    msg_runs = min(cap, max(MSG_RUNS_MIN, (Y * X) // MSG_RUNS_DIV))
    recv = yield from ctx.shift_up(torch.cat([top_bits, vtop[:msg_runs]]))
    with torch.cuda.device(dev):
        s = _lib.stream()
        if rank > 0:
            bitsA, vA = recv[:PW], recv[PW:]
            prefA = empty_u32(PW, dev)
            scratch = torch.zeros(1, dtype=torch.int64, device=dev)
            _lib.call("cfe_scan_run_starts", bitsA.data_ptr(), Y, X, prefA.data_ptr(), scratch.data_ptr(),
                      ws.data_ptr(), s)
            # plane B is this slab's first plane: its run ids are 0..n_bot-1, so `vbot` maps them to
            # vertices; the kernel emits a spanning forest of (vertex of A's run, vertex of B's run)
            # straight into the message
            forest = empty_u32(2 * Cv, dev)
            _lib.call("cfe_ccl_interface_forest", bitsA.data_ptr(), prefA.data_ptr(), vA.data_ptr(), msg_runs,
                      bits.data_ptr(), pref.data_ptr(), vbot.data_ptr(), Y, X, Cv, forest.data_ptr(),
                      pack.data_ptr() + 4 * (4 + 2 * Cv), pack.data_ptr() + 4, Ce, s)
    packs = yield from ctx.all_gather(pack)
    with torch.cuda.device(dev):
        s = _lib.stream()
        P = torch.stack(packs).contiguous()                     # [N, PS]
        label = empty_u32(N * Cv, dev)
        comp = torch.empty(N * Cv, dtype=torch.int64, device=dev)
        ckey = torch.empty(N * Cv, dtype=torch.int64, device=dev)
        result = torch.zeros(4, dtype=torch.int64, device=dev)
        _lib.call("cfe_ccl_boundary_solve", P.data_ptr(), N, Cv, Ce, rank, label.data_ptr(), comp.data_ptr(),
                  ckey.data_ptr(), size.data_ptr(), result.data_ptr(), s)
        if X % 32 == 0 and out.data_ptr() % 16 == 0:
            sel = empty_u32(rows * W, dev)          # the kept voxels bit-packed, for vol2ugrid_program
            _lib.call("cfe_ccl_write_mask_bits", bits.data_ptr(), pref.data_ptr(), Zl, Y, X, parent.data_ptr(),
                      stats.data_ptr() + 8, size.data_ptr(), out.data_ptr(), sel.data_ptr(), s)
            tag_packed(out, sel)
        else:
            _lib.call("cfe_ccl_write_mask", bits.data_ptr(), pref.data_ptr(), Zl, Y, X, parent.data_ptr(),
                      stats.data_ptr() + 8, size.data_ptr(), out.data_ptr(), s)
        res = read_back(result)                                 # the one host read of the stage
    if info is not None:      # diagnostics: the slab's forest and the gathered boundary graph
        info.update(parent=parent, stats=stats, packs=P, label=label, Cv=Cv, Ce=Ce, host_merge=bool(res[2]))
    if res[2]:
        # some rank has more boundary components / edges than the fixed capacity: every rank sees the
        # same flag (nothing was marked) and redoes the merge host-driven at exact sizes
        if hasattr(out, "_cfe_bits"):
            del out._cfe_bits
        return (yield from _ccl_host_merge(ctx, state))
    if res[0] <= 0:
        return out, None
    return out, (int(res[0]), int(res[1]))


def _ccl_host_merge(ctx, st):
    """Host-driven cross-slab merge (exact sizes, a few host reads): the fallback of ccl_program."""
    dev = st["bits"].device
    bits, pref, parent, size, stats, cnt, out, ws = (st[k] for k in
                                                      ("bits", "pref", "parent", "size", "stats", "cnt", "out", "ws"))
    Zl, Y, X = st["dims"]
    cap = st["cap"]
    PW = Y * ((X + 31) // 32)
    if st.get("needs_relabel"):
        # the device path overwrote a few sizes with markers: recompute them
        with torch.cuda.device(dev):
            _lib.call("cfe_ccl_label", bits.data_ptr(), pref.data_ptr(), stats.data_ptr(), Zl, Y, X,
                      parent.data_ptr(), size.data_ptr(), _lib.stream())
    roots_bot, roots_top, top_bits = st["roots_bot"], st["roots_top"], st["top_bits"]
    recv = yield from ctx.shift_up(torch.cat([top_bits, roots_top]))
    with torch.cuda.device(dev):
        s = _lib.stream()
        cnt[2] = 0
        edges_dev = None
        if ctx.rank > 0:
            bitsA, rootsA = recv[:PW], recv[PW:]
            prefA = empty_u32(PW, dev)
            scratch = torch.zeros(1, dtype=torch.int64, device=dev)
            _lib.call("cfe_scan_run_starts", bitsA.data_ptr(), Y, X, prefA.data_ptr(), scratch.data_ptr(),
                      ws.data_ptr(), s)
            edges_dev = torch.empty((max(cap, 1), 2), dtype=torch.int32, device=dev)
            _lib.call("cfe_ccl_interface_edges", bitsA.data_ptr(), prefA.data_ptr(), rootsA.data_ptr(),
                      bits.data_ptr(), pref.data_ptr(), parent.data_ptr(), Y, X, edges_dev.data_ptr(),
                      cnt.data_ptr() + 8, cap, s)
        h = torch.cat([cnt[:3].to(torch.int64), stats[:1]]).cpu().tolist()   # sync: plane runs, edges, runs
        n_bot, n_top, n_edges, n_runs = h
        if n_edges > cap:
            raise _lib.CfeError("interface edge buffer overflow")
        parts = []
        if ctx.rank > 0:
            parts.append(roots_bot[:n_bot])
        if ctx.rank + 1 < ctx.size:
            parts.append(roots_top[:n_top])
        my_roots = torch.unique(_u32_to_i64(torch.cat(parts))) if parts else torch.empty(0, dtype=torch.int64,
                                                                                         device=dev)
        my_sizes = _u32_to_i64(size[my_roots]) if my_roots.numel() else my_roots.clone()
        if ctx.rank > 0 and n_edges:
            e = _u32_to_i64(edges_dev[:n_edges])
            my_edges = torch.stack([e[:, 0] | ((ctx.rank - 1) << 32), e[:, 1] | (ctx.rank << 32)], dim=1)
        else:
            my_edges = torch.empty((0, 2), dtype=torch.int64, device=dev)
        nr, ne = my_roots.numel(), my_edges.shape[0]
        if nr:
            saved = size[my_roots].clone()
            size[my_roots] = 0
        _lib.call("cfe_ccl_select", parent.data_ptr(), size.data_ptr(), stats.data_ptr(), stats.data_ptr() + 8, s)
        if nr:
            size[my_roots] = saved
        head = torch.cat([to_device_async([nr, ne], torch.int64, dev), stats[1:2]])
    heads = yield from ctx.all_gather(head)
    hh = torch.stack(heads).cpu().tolist()               # sync: everybody's sizes + best internal
    maxlen = max(max(2 * a + 2 * b for a, b, _ in hh), 1)
    with torch.cuda.device(dev):
        pack = torch.zeros(maxlen, dtype=torch.int64, device=dev)
        pack[:nr] = my_roots
        pack[nr:2 * nr] = my_sizes
        pack[2 * nr:2 * nr + 2 * ne] = my_edges.reshape(-1)
    packs = yield from ctx.all_gather(pack)
    with torch.cuda.device(dev):
        s = _lib.stream()
        keys = torch.cat([packs[k][:a] | (k << 32) for k, (a, b, _) in enumerate(hh)])   # sorted by rank, root
        sizes = torch.cat([packs[k][a:2 * a] for k, (a, b, _) in enumerate(hh)])
        edges = torch.cat([packs[k][2 * a:2 * a + 2 * b] for k, (a, b, _) in enumerate(hh)]).reshape(-1, 2)
        lab, comp = solve_boundary_components(keys, sizes, edges)
        cands = []
        for k, (_, _, best) in enumerate(hh):
            v = best & 0xFFFFFFFFFFFFFFFF
            cands.append((v >> 32, (k << 32) | (0xFFFFFFFF - (v & 0xFFFFFFFF))))
        wlab = -1
        t = None
        if keys.numel():
            iota = torch.arange(keys.numel(), device=dev)
            rep_size = torch.where(lab == iota, comp, torch.zeros_like(comp))
            top = rep_size.max()
            top_idx = torch.where(rep_size == top, iota, torch.full_like(lab, keys.numel())).min()
            t = torch.stack([top, keys[top_idx.clamp(max=keys.numel() - 1)], top_idx]).cpu().tolist()   # sync
            cands.append((t[0], t[1]))
        winner = pick_winner(cands)
        if winner is None:
            out.zero_()
            return out, None
        wsize, wkey = winner
        if t is not None and wkey == t[1] and wsize == t[0]:
            wlab = t[2]
        off = sum(a for a, _, _ in hh[: ctx.rank])
        if wlab >= 0:                        # the winner is a merged (boundary) component
            if nr:
                size[my_roots] = torch.where(lab[off: off + nr] == wlab, torch.full_like(saved, _MARK), saved)
        elif (wkey >> 32) == ctx.rank:       # a slab-internal component of this rank
            size[wkey & 0xFFFFFFFF] = _MARK
        _lib.call("cfe_ccl_write_mask", bits.data_ptr(), pref.data_ptr(), Zl, Y, X, parent.data_ptr(),
                  stats.data_ptr() + 8, size.data_ptr(), out.data_ptr(), s)
    return out, winner


class SlabMesh:
    """One rank's part of the global voxel mesh (device-resident)."""

    @property
    def node_list(self):
        """Local grid index of every used node of the slab (only the variable-width *NODE path reads it)."""
        if getattr(self, "_node_list", None) is None:
            dev = self.vol.device
            with torch.cuda.device(dev):
                self._node_list = empty_u32(self.n_nd, dev)
                _lib.call("cfe_fill_list", self.nbits.data_ptr(), self.npref.data_ptr(), self.n_node_words,
                          (self.X + 32) // 32, self.X + 1, self._node_list.data_ptr(), _lib.stream())
        return self._node_list


def vol2ugrid_program(ctx, vol, is_bool, Zg, z0, voxelsize, GVmin=0, plane_lock_num=1, node_level_lock_num=1,
                      locking_strategy="plane", eid_base=0, packed=None):
    """vol2ugrid of the global volume [Zg,Y,X]; `vol` = this rank's uint8 slab starting at slice z0.
    `packed`: the slab's ``vol > 0`` already bit-packed (``_util.packed_of`` of a remove_unconnected mask)."""
    _lib.require_cuda()
    if locking_strategy not in ("plane", "exact"):
        raise ValueError("locking_strategy must be 'plane' or 'exact'")
    p = vfe._int_lock(plane_lock_num, "plane_lock_num")
    kk = vfe._int_lock(node_level_lock_num, "node_level_lock_num")
    dev = vol.device
    Zl, Y, X = (int(v) for v in vol.shape)
    last = ctx.rank == ctx.size - 1
    n_layers = Zl + (1 if last else 0)
    thr = vfe._gv_threshold(GVmin)
    W, WN = (X + 31) // 32, (X + 32) // 32
    rows, nrows = Zl * Y, n_layers * (Y + 1)
    PW = Y * W
    m = SlabMesh()
    m.rank, m.size = ctx.rank, ctx.size
    m.vol, m.is_bool, m.thr = vol, is_bool, thr
    m.Zg, m.Zl, m.Y, m.X, m.z0, m.n_layers = Zg, Zl, Y, X, z0, n_layers
    with torch.cuda.device(dev):
        s = _lib.stream()
        if packed is not None and thr == 0 and packed.numel() >= rows * W:
            bits = packed
        else:
            bits = empty_u32(rows * W, dev)
            # the first kernel goes out before anything else is allocated or computed on the host
            _lib.call("cfe_pack_bits", vol.data_ptr(), rows, X, thr, bits.data_ptr(), s)
        epref = empty_u32(rows * W, dev)
        nbits = empty_u32(nrows * WN, dev)
        npref = empty_u32(nrows * WN, dev)
        ws = scan_workspace(max(rows * W, nrows * WN), dev)
        stats = torch.zeros(4, dtype=torch.int64, device=dev)
        top_bits = bits[(Zl - 1) * PW: Zl * PW]
    halo = yield from ctx.shift_up(top_bits)
    with torch.cuda.device(dev):
        s = _lib.stream()
        _lib.call("cfe_scan_elements", bits.data_ptr(), rows * W, epref.data_ptr(), stats.data_ptr(),
                  ws.data_ptr(), s)
        chunk_word = empty_u32(n_layers * (Y + 1) * (X + 1) // 512 + 2, dev)   # rank index of the *NODE emitter
        _lib.call("cfe_scan_nodes_index", bits.data_ptr(), _lib.ptr(halo), Zl, Y, X, n_layers, nbits.data_ptr(),
                  npref.data_ptr(), stats.data_ptr() + 8, chunk_word.data_ptr(), ws.data_ptr(), s)
        # boundary sets: global boxes clipped to the layers this slab owns
        boxes = vfe._boundary_boxes(locking_strategy, p, kk, Zg, Y, X)
        descs = (CfeSetDesc * 12)()
        tile_base = 0
        for q in range(12):
            d = descs[q]
            b = boxes[q % 6]
            if locking_strategy == "plane":
                if q < 6:
                    d.kind = 0
                    cand = tuple((lo, hi + 1) if hi > lo else (0, 0) for lo, hi in b)
                    vbox = b
                    cand = clip_box_z(cand, z0, z0 + n_layers) if vfe._box_count(cand) else cand
                else:
                    d.kind = 2
                    cand = clip_box_z(b, z0, z0 + Zl) if vfe._box_count(b) else b
                    vbox = b
            else:
                if q < 6:
                    d.kind = 1
                    cand = clip_box_z(b, z0, z0 + n_layers) if vfe._box_count(b) else b
                    vbox = ((0, Zg), (0, Y), (0, X))
                else:
                    d.kind = 2
                    cand = ((0, 0), (0, 0), (0, 0))
                    vbox = cand
            (d.s0, d.s1), (d.r0, d.r1), (d.c0, d.c1) = cand
            (d.vs0, d.vs1), (d.vr0, d.vr1), (d.vc0, d.vc1) = vbox
            d.n_cand = vfe._box_count(cand)
            d.tile_base = tile_base
            tile_base += (d.n_cand + 255) // 256
        n_tiles = tile_base
        descs_dev = bytes_to_device_async(bytes(descs), dev)
        tile_count = empty_u32(n_tiles, dev)
        tile_off = torch.empty(max(n_tiles, 1) + 1, dtype=torch.int64, device=dev)
        # element ids need the global offset -> counts first, ids after the all-gather
        _lib.call("cfe_sets_count", descs_dev.data_ptr(), 12, n_tiles, Zl, Y, X, z0, 0, bits.data_ptr(),
                  epref.data_ptr(), _lib.ptr(halo), tile_count.data_ptr(), s)
        _lib.call("cfe_scan_u32", tile_count.data_ptr(), n_tiles, tile_off.data_ptr(), stats.data_ptr() + 16,
                  ws.data_ptr(), s)
        # per-axis coordinate values and their text fields (memoised: a pure function of voxel size and shape)
        tkey = vfe._tables_key(voxelsize, Zg, Y, X, dev, s)
        tables = vfe._TABLES.get(tkey)
        if tables is None:
            m.xs, m.ys, m.zs = vfe._coordinate_tables(voxelsize, Zg, Y, X)
            m.ftab, m.coord_vals = vfe._format_tables_device(m.xs, m.ys, m.zs, dev, stats.data_ptr() + 24, s)
        else:
            m.xs, m.ys, m.zs, m.ftab, m.coord_vals, cached_overflow = tables
            if cached_overflow:
                stats[3] = 1
        mine = torch.empty(15, dtype=torch.int64, device=dev)   # n_el, n_nd, 12 set counts, field overflow
        _lib.call("cfe_set_counts", descs_dev.data_ptr(), 12, n_tiles, tile_off.data_ptr(), stats.data_ptr(),
                  mine.data_ptr(), s)
    gathered = yield from ctx.all_gather(mine)
    with torch.cuda.device(dev):
        s = _lib.stream()
        table = read_back(torch.stack(gathered)) if len(gathered) > 1 else read_back(gathered[0])
        table = [table[15 * k: 15 * k + 15] for k in range(len(gathered))]    # the one host sync: [rank][15]
        if tables is None:
            vfe._TABLES[tkey] = (m.xs, m.ys, m.zs, m.ftab, m.coord_vals, bool(table[ctx.rank][14]))
            while len(vfe._TABLES) > 8:
                vfe._TABLES.popitem(last=False)
        n_el_r = [t[0] for t in table]
        n_nd_r = [t[1] for t in table]
        m.n_el, m.n_nd = n_el_r[ctx.rank], n_nd_r[ctx.rank]
        m.eid_offset = eid_base + sum(n_el_r[: ctx.rank])   # eid_base: elements of slabs outside this group
        m.node_rank_offset = sum(n_nd_r[: ctx.rank])
        m.n_el_global, m.n_nd_global = sum(n_el_r), sum(n_nd_r)
        m.n_el_by_rank = [int(v) for v in n_el_r]
        m.max_eid = eid_base + m.n_el_global                 # largest element id that can be printed
        m.set_count = [int(c) for c in table[ctx.rank][2:14]]
        m.field_overflow = any(t[14] for t in table)
        m.set_first, n_ids = exclusive_offsets(m.set_count)
        m.set_rank0 = [sum(t[2 + q] for t in table[: ctx.rank]) for q in range(12)]
        m.set_count_global = [sum(t[2 + q] for t in table) for q in range(12)]
        m.elem_vox = empty_u32(m.n_el, dev)
        m.set_ids = torch.empty(max(n_ids, 1), dtype=torch.int64, device=dev)
        lib = _lib.load()
        m.ws_el = torch.empty(lib.cfe_emit_elements_workspace_bytes(m.n_el) + 16, dtype=torch.uint8, device=dev)
        # grey-scale slabs: cell_GV compacted in element order (the GV histogram and partition read it)
        m.gv8 = None if (is_bool and thr >= 0) else torch.empty(max(m.n_el, 1), dtype=torch.uint8, device=dev)
        _lib.call("cfe_fill_elements_gv", bits.data_ptr(), epref.data_ptr(), rows * W, Y, X, z0, m.eid_offset,
                  m.n_el, m.elem_vox.data_ptr(), m.ws_el.data_ptr(), vol.data_ptr(), _lib.ptr(m.gv8), s)
        m.nbits, m.npref, m.chunk_word, m.n_node_words = nbits, npref, chunk_word, nrows * WN
        if vfe.NODE_EMITTER != "bits":
            m.node_list                       # k_fill_list now
        _lib.call("cfe_sets_fill", descs_dev.data_ptr(), 12, n_tiles, Zl, Y, X, z0, m.eid_offset, bits.data_ptr(),
                  epref.data_ptr(), _lib.ptr(halo), tile_off.data_ptr(), m.set_ids.data_ptr(), s)
    m.bits, m.epref = bits, epref
    return m


class SlabDeck:
    """One rank's deck pieces: device buffers + (file offset, buffer, lo, hi) write list."""


def deck_program(ctx, m, templatefile, matprop=None, keywords=['NSET', 'ELSET'], eltype='C3D8', matpropbits=8,
                 refnode=None):
    """mesh2voxelfe of the global mesh.  Every rank formats its own elements / nodes / set entries
    into HBM; the all-gathered byte totals give each piece its offset in the single .INP file."""
    lib = _lib.require_cuda()
    dev = m.vol.device
    Zl, Y, X = m.Zl, m.Y, m.X
    rank = ctx.rank
    wide = m.field_overflow      # variable-length *NODE lines (some '{:12.6f}' field needs > 12 characters)
    nodes_buf = None
    if not wide:
        # fixed 53-byte *NODE lines need nothing but the mesh: start them before any host work so that the
        # GPU formats nodes while the host prepares the element / set sections
        with torch.cuda.device(dev):
            nodes_buf = torch.empty(max(53 * m.n_nd, 1) + 64, dtype=torch.uint8, device=dev)
            p_xt0 = m.ftab.data_ptr()
            p_yt0 = p_xt0 + 16 * (X + 1)
            vfe._emit_node_lines(m, p_xt0, p_yt0, p_yt0 + 16 * (Y + 1) + 16 * m.z0, Y, X, m.z0,
                                 nodes_buf.data_ptr(), _lib.stream())
    if m.gv8 is None:
        # a boolean volume has one grey value (True -> SET1): the histogram is the element count,
        # which every rank already knows from vol2ugrid's all-gather
        hist_r = np.zeros((ctx.size, 256), dtype=np.int64)
        hist_r[:, 1] = m.n_el_by_rank
    else:
        with torch.cuda.device(dev):
            s = _lib.stream()
            hist_dev = torch.zeros(256, dtype=torch.int64, device=dev)
            # over the compact grey values of the slab's elements (every one counts: threshold -1)
            _lib.call("cfe_gv_histogram", m.gv8.data_ptr(), m.n_el, -1, hist_dev.data_ptr(), s)
        hists = yield from ctx.all_gather(hist_dev)
        hist_r = torch.stack(hists).cpu().numpy()        # [rank][256]
    hist = hist_r.sum(axis=0)
    if m.n_el_global == 0:
        raise ValueError("min() iterable argument is empty")
    gv_dtype = np.bool_ if m.is_bool else np.uint8
    gvs = np.flatnonzero(hist)
    existing_GV = gvs.astype(gv_dtype)
    GVmax = 2 ** matpropbits - 1
    binary_model = bool(gvs.min() == 0) and bool(gvs.max() == 1) and issubclass(gv_dtype, np.integer)
    if max(existing_GV) > GVmax:
        GVmax = max(existing_GV)
    prop_GV = {0: (1, 1)} if binary_model else vfe._resolve_matprop(matprop, GVmax)

    from datetime import datetime
    # every rank derives the file layout from the header's length, and str(datetime.now()) is 26 or 19
    # characters long (microsecond == 0): rank 0's stamp travels with the byte totals below and is the
    # one every rank uses
    stamp = str(datetime.now()).encode()[:32].ljust(32, b"\0")
    elem_comment = b'** Elements and Element sets from input model\n'
    tail = ''
    if 'PROPERTY' in keywords:
        tail += vfe._property_cards(matprop, prop_GV, existing_GV)
    tail += vfe._template_text(templatefile, refnode)
    tail = tail.encode()
    hdrs = {int(g): '*ELEMENT, TYPE={0}, ELSET={1}\n'.format(eltype, 'SET' + str(int(g))).encode() for g in gvs}

    my_hist = hist_r[rank]
    first = np.zeros(256, dtype=np.int64)
    first[1:] = np.cumsum(my_hist)[:-1]   # emission index of the first element with grey value >= g
    blob = vfe._Blob()
    o_first = blob.add(first.tobytes())
    o_hdr = blob.add(bytes(256 * vfe._ELEM_HDR_SLOT))
    o_hlen = blob.add(bytes(256))                      # no headers inside the slab buffers
    slot = 32 if wide else 16
    if wide:
        with torch.cuda.device(dev):
            s = _lib.stream()
            nv = m.coord_vals.numel()
            wtab = torch.empty(32 * nv, dtype=torch.uint8, device=dev)
            nstat = torch.zeros(2, dtype=torch.int64, device=dev)          # *NODE bytes of this slab, bad flag
            _lib.call("cfe_format_coords_wide", m.coord_vals.data_ptr(), nv, wtab.data_ptr(), nstat.data_ptr() + 8, s)
        p_xt = wtab.data_ptr()
    else:
        p_xt = m.ftab.data_ptr()
    p_yt = p_xt + slot * (X + 1)
    p_zt = p_yt + slot * (Y + 1) + slot * m.z0      # this slab's node layers

    # id-list segments of this rank, one per non-empty set, in file order
    set_names = []
    if 'NSET' in keywords:
        set_names += [("N", q) for q in range(8)]
    if 'ELSET' in keywords:
        set_names += [("E", q) for q in range(6)]
    segs, seg_sets, n_items = [], [], 0
    for kind, q in set_names:
        if kind == "N" and q >= 6:       # corner sets: first two entries of the GLOBAL NODES_Z0 / Z1
            src = 4 if q == 6 else 5
            before = m.set_rank0[src]
            cnt = max(0, min(2 - before, m.set_count[src]))
            f, r0 = m.set_first[src], before
        else:
            src = q if kind == "N" else 6 + q
            cnt, f, r0 = m.set_count[src], m.set_first[src], m.set_rank0[src]
        if cnt:
            segs.append((1, n_items, 0, 0, f, r0))
            seg_sets.append((kind, q))
            n_items += cnt
    seg_arr = (CfeItemSeg * max(len(segs), 1))()
    for i, sg in enumerate(segs):
        a = seg_arr[i]
        a.kind, a.item0, a.lit_off, a.lit_len, a.id0, a.rank0 = sg
    o_segs = blob.add(bytes(seg_arr))

    max_node_id = (m.Zg + 1) * (Y + 1) * (X + 1) - 1
    if max_node_id >= 10 ** 10:
        raise NotImplementedError("node ids wider than the 10-character field are not supported")
    if m.max_eid >= 10 ** 10:
        raise NotImplementedError("element ids wider than 10 digits are not supported")
    max_line = len(str(m.max_eid)) + 8 * len(str(max_node_id)) + 9
    with torch.cuda.device(dev):
        s = _lib.stream()
        dblob = blob.to_device(dev)
        base = dblob.data_ptr()
        if wide:
            nodes_buf = torch.empty(max(96 * m.n_nd, 1) + 64, dtype=torch.uint8, device=dev)
        elem_buf = torch.empty(max(m.n_el * max_line, 1) + 64, dtype=torch.uint8, device=dev)
        items_buf = torch.empty(n_items * (len(str(max(max_node_id, m.max_eid))) + 1) + 64, dtype=torch.uint8,
                                device=dev)
        totals = torch.zeros(2, dtype=torch.int64, device=dev)
        blk_off = torch.zeros(256, dtype=torch.int64, device=dev)
        seg_off = torch.zeros(max(len(segs), 1), dtype=torch.int64, device=dev)
        ws = scan_workspace(max(n_items, 1), dev)
        ws_el = m.ws_el
        if wide:
            n_ntiles = (m.n_nd + 255) // 256
            nlen8 = torch.empty(max(m.n_nd, 1), dtype=torch.uint8, device=dev)
            ntile_tot = empty_u32(n_ntiles, dev)
            ntile_off = torch.empty(max(n_ntiles, 1), dtype=torch.int64, device=dev)
            _lib.call("cfe_node_lengths_wide", m.node_list.data_ptr(), m.n_nd, p_xt, p_yt, p_zt, Y, X, m.z0,
                      nlen8.data_ptr(), ntile_tot.data_ptr(), s)
            _lib.call("cfe_scan_u32", ntile_tot.data_ptr(), n_ntiles, ntile_off.data_ptr(), nstat.data_ptr(),
                      scan_workspace(max(n_ntiles, 1), dev).data_ptr(), s)
            _lib.call("cfe_emit_nodes_wide", m.node_list.data_ptr(), m.n_nd, p_xt, p_yt, p_zt, Y, X, m.z0,
                      nlen8.data_ptr(), ntile_off.data_ptr(), nodes_buf.data_ptr(), s)
        perm = None
        single_gv = int(gvs[0])
        if len(gvs) > 1:
            perm = torch.empty((max(m.n_el, 1), 2), dtype=torch.int32, device=dev)
            ws_el = torch.empty(m.ws_el.numel(), dtype=torch.uint8, device=dev)
            ws_p = torch.empty(lib.cfe_gv_partition_workspace_bytes(m.n_el), dtype=torch.uint8, device=dev)
            _lib.call("cfe_gv_partition", m.vol.data_ptr(), m.elem_vox.data_ptr(), m.n_el, _lib.ptr(m.gv8),
                      m.ws_el.data_ptr(), ws_el.data_ptr(), perm.data_ptr(), ws_p.data_ptr(), s)
            single_gv = -1
        el_args = (m.elem_vox.data_ptr(), _lib.ptr(perm), m.n_el, m.vol.data_ptr(), base + o_first, base + o_hdr,
                   base + o_hlen, single_gv if single_gv >= 0 else 0, Zl, Y, X, m.z0, m.eid_offset)
        _lib.call("cfe_element_lengths", *el_args, 1, totals.data_ptr(), ws_el.data_ptr(), s)
        _lib.call("cfe_emit_elements", *el_args, elem_buf.data_ptr(), blk_off.data_ptr(), ws_el.data_ptr(), s)
        _lib.call("cfe_emit_items", base + o_segs, len(segs), n_items, m.set_ids.data_ptr(), base, items_buf.data_ptr(),
                  None, totals.data_ptr() + 8, seg_off.data_ptr(), ws.data_ptr(), s)
        # byte sizes of my (gv) element blocks and of my set segments, computed on the device
        my_gvs = [int(g) for g in gvs if my_hist[g]]
        gv_bytes = torch.zeros(256, dtype=torch.int64, device=dev)
        if my_gvs:
            gi = to_device_async(my_gvs, torch.int64, dev)
            starts = blk_off[gi]
            gv_bytes[gi] = torch.cat([starts[1:], totals[0:1]]) - starts
        set_bytes = torch.zeros(14, dtype=torch.int64, device=dev)
        cols = [(q if kind == "N" else 8 + q) for kind, q in seg_sets]
        if cols:
            ci = to_device_async(cols, torch.int64, dev)
            st = seg_off[: len(cols)]
            set_bytes[ci] = torch.cat([st[1:], totals[1:2]]) - st
        node_b = nstat[0:1] if wide else to_device_async([53 * m.n_nd], torch.int64, dev)
        bad_b = nstat[1:2] if wide else torch.zeros(1, dtype=torch.int64, device=dev)
        stamp_t = to_device_async(np.frombuffer(stamp, dtype=np.int64).tolist(), torch.int64, dev)
        mine = torch.cat([gv_bytes, set_bytes, node_b, bad_b, stamp_t])
    allb = yield from ctx.all_gather(mine)
    host = torch.cat([torch.stack(allb).reshape(-1), totals, blk_off, seg_off]).cpu().numpy()   # the one sync
    nb = ctx.size * 276
    allb = host[:nb].reshape(ctx.size, 276)   # [rank][256 grey values + 14 sets + nodes + bad flag + 4 stamp words]
    head = ('** ---------------------------------------------------------\n'
            '** Abaqus .INP file written by Voxel_FEM script on {}\n'
            '** ---------------------------------------------------------\n'
            '** Node coordinates from input model\n'
            '*NODE\n').format(allb[0, 272:276].tobytes().rstrip(b"\0").decode()).encode()
    if allb[:, 271].any():
        raise NotImplementedError("node coordinates that are not finite or not below 1e18 cannot be formatted")
    blk = host[nb + 2: nb + 258].tolist()
    sgo = host[nb + 258: nb + 258 + len(segs)].tolist()
    it_total = int(host[nb + 1])
    seg_local = {}
    for i, col in enumerate(cols):
        seg_local[col] = (sgo[i], sgo[i + 1] if i + 1 < len(cols) else it_total)

    # ---- global file layout (identical on every rank)
    writes = []        # (file offset, device tensor, lo, hi)
    lits = []          # (file offset, bytes) -- written by rank 0
    pos = 0
    lits.append((pos, head)); pos += len(head)
    for k in range(ctx.size):
        nb = int(allb[k, 270])
        if k == rank and nb:
            writes.append((pos, nodes_buf, 0, nb))
        pos += nb
    lits.append((pos, elem_comment)); pos += len(elem_comment)
    for g in gvs.tolist():
        lits.append((pos, hdrs[g])); pos += len(hdrs[g])
        for k in range(ctx.size):
            nb = int(allb[k, g])
            if k == rank and nb:
                writes.append((pos, elem_buf, blk[g], blk[g] + nb))
            pos += nb

    def set_block(names, prologue, header_fmt, col0):
        nonlocal pos
        lits.append((pos, prologue)); pos += len(prologue)
        for q, name in enumerate(names):
            h = header_fmt.format(name).encode()
            lits.append((pos, h)); pos += len(h)
            for k in range(ctx.size):
                nb = int(allb[k, 256 + col0 + q])
                if k == rank and nb:
                    lo, hi = seg_local[col0 + q]
                    writes.append((pos, items_buf, lo, hi))
                pos += nb
            lits.append((pos, b'\n')); pos += 1

    if 'NSET' in keywords:
        set_block(vfe.NODE_SET_NAMES + ["NODES_X0Y0Z0", "NODES_X0Y0Z1"],
                  ('** Additional nset from voxel model. New generated nsets are:\n'
                   '** {}, \n'.format(vfe.NODE_SET_NAMES[0])).encode(), '*NSET, NSET={0}\n', 0)
    if 'ELSET' in keywords:
        set_block(vfe.CELL_SET_NAMES,
                  ('** Additional elset from voxel model. New generated elsets are:\n'
                   '** {}, \n'.format(vfe.CELL_SET_NAMES[0])).encode(), '*ELSET, ELSET={}\n', 8)
    lits.append((pos, tail)); pos += len(tail)

    d = SlabDeck()
    d.rank, d.size = rank, ctx.size
    d.total_bytes = pos
    d.writes = writes
    d.literals = lits if rank == 0 else []
    d.buffers = (nodes_buf, elem_buf, items_buf, dblob)
    d.local_bytes = sum(hi - lo for _, _, lo, hi in writes)
    return d


def write_deck_program(ctx, deck, fileout):
    """All ranks put their pieces of the global deck at the offsets of the global layout: into ONE file
    (``fileout`` = path; ``os.pwrite``) or into any object with ``pwrite(buffer, offset)``.  The bytes
    leave the GPU through the pinned staging ring of ``voxelFE.stream_device_bytes`` (the copy of the
    next chunk overlaps the write of this one)."""
    to_path = isinstance(fileout, (str, bytes, os.PathLike))
    if to_path:
        if ctx.rank == 0:
            with open(fileout, "wb") as f:
                f.truncate(deck.total_bytes)
        yield from ctx.barrier()
        fd = os.open(fileout, os.O_WRONLY)

        def put(mv, off):
            os.pwrite(fd, mv, off)
    else:
        put = fileout.pwrite
    try:
        vfe.stream_device_bytes([(off, buf[lo:hi]) for off, buf, lo, hi in deck.writes], put)
        for off, data in deck.literals:
            put(data, off)
    finally:
        if to_path:
            os.close(fd)
    yield from ctx.barrier()
    return deck.total_bytes


# --------------------------------------------------------------------------------------------
# convenience drivers

def _slab_inputs(vol, n):
    """Cut a global (host or device) volume into n device slabs: [(uint8 slab tensor, z0)], is_bool."""
    from ._util import as_device_volume
    t, is_bool, _ = as_device_volume(vol)
    return [(t[a:b].contiguous(), a) for a, b in slab_ranges(int(t.shape[0]), n)], is_bool, int(t.shape[0])


def emulate_remove_unconnected(vol, n, force_host_merge=False):
    """N emulated ranks on one GPU: returns the global bool mask (numpy) and the winner."""
    slabs, _, _ = _slab_inputs(vol, n)
    runner = LocalRunner(n)
    res = runner.run([ccl_program(c, s, force_host_merge) for c, (s, _) in zip(runner.ctxs(), slabs)])
    if res[0][1] is None:
        raise ValueError("attempt to get argmax of an empty sequence")
    mask = torch.cat([r[0] for r in res]).view(torch.bool)
    return mask, res[0][1]


def emulate_ct2inp(vol, n, voxelsize, templatefile, fileout, u_kwargs=None, m_kwargs=None):
    """vol2ugrid + mesh2voxelfe on N emulated ranks of one GPU, written to `fileout`."""
    slabs, is_bool, Zg = _slab_inputs(vol, n)
    runner = LocalRunner(n)
    meshes = runner.run([vol2ugrid_program(c, s, is_bool, Zg, z0, voxelsize, **(u_kwargs or {}))
                         for c, (s, z0) in zip(runner.ctxs(), slabs)])
    mk = dict(m_kwargs or {})
    decks = runner.run([deck_program(c, m, templatefile, **{k: (v() if callable(v) else v) for k, v in mk.items()})
                        for c, m in zip(runner.ctxs(), meshes)])
    runner.run([write_deck_program(c, d, fileout) for c, d in zip(runner.ctxs(), decks)])
    return meshes, decks
