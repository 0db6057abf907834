"""The hot path with the reference's signatures for a volume sharded as z-slabs over the GPUs of one box
(one process per GPU under ``torchrun``):

* ``remove_unconnected(bwslab)``                 reference: src/ciclope/utils/preprocess.py:119-143
* ``vol2ugrid(volslab, voxelsize, ...)``         reference: src/ciclope/core/voxelFE.py:15-283
* ``mesh2voxelfe(mesh, templatefile, fileout)``  reference: src/ciclope/core/voxelFE.py:474-772

Every rank passes ITS slab (contiguous slices ``[z0, z0 + Zl)`` of the global volume, slabs in rank order)
and the same remaining arguments; the results are the global ones: the kept component is the largest of
the whole volume, node / element ids and set positions are global, and ``mesh2voxelfe`` writes ONE file
whose bytes are those of the single-GPU (and of the reference's) deck.  The additive keyword arguments
are ``group`` (a ``torch.distributed`` process group, default: the world), ``z0`` and ``Z_global`` (default:
derived from the all-gathered slab heights).  Thin wrappers over the rank programs of ``sharded.py``.
"""
import logging

import numpy as np
import torch

from . import sharded
from . import voxelFE as vfe
from ._util import as_device_volume, packed_of, read_back, tag_packed, to_device_async
from . import _lib

_RUNNERS = {}


def runner(group=None):
    """The DistRunner of a process group (peer-memory exchange buffers are set up once per group)."""
    key = id(group) if group is not None else None
    r = _RUNNERS.get(key)
    if r is None:
        r = _RUNNERS[key] = sharded.DistRunner(group)
    return r


def shutdown():
    """Drop the cached runners (before ``destroy_process_group``)."""
    _RUNNERS.clear()


def _binary_check_program(ctx, vol, who):
    """uint8 slabs: the GLOBAL image must hold one non-zero value (see preprocess.remove_unconnected)."""
    span = torch.empty(2, dtype=torch.int32, device=vol.device)
    with torch.cuda.device(vol.device):
        _lib.call("cfe_value_span_u8", vol.data_ptr(), vol.numel(), span.data_ptr(), _lib.stream())
    spans = yield from ctx.all_gather(span)
    host = read_back(torch.stack(spans).reshape(-1))
    hi = max(int(v) & 0xFFFFFFFF for v in host[0::2])
    lo = min(int(v) & 0xFFFFFFFF for v in host[1::2])
    if hi and lo != hi:
        raise ValueError("%s: uint8 image holds several non-zero values (%d..%d); pass a bool image or one "
                         "with a single non-zero value" % (who, lo, hi))


def remove_unconnected(bwimage, group=None):
    """Largest 6-connected component of the GLOBAL binary image; ``bwimage`` is this rank's z-slab.
    Returns this rank's slab of the result (bool; NumPy in -> NumPy out, CUDA tensor in -> CUDA tensor out).
    Raises ``ValueError`` on every rank when the global image is all background (preprocess.py:141)."""
    vol, is_bool, from_numpy = as_device_volume(bwimage)
    r = runner(group)
    if not is_bool:
        r.run(_binary_check_program(r.ctx(), vol, "remove_unconnected"))
    mask, winner = r.run(sharded.ccl_program(r.ctx(), vol))
    if winner is None:
        raise ValueError("attempt to get argmax of an empty sequence")
    out = mask.view(torch.bool)
    if from_numpy:
        return out.cpu().numpy()
    kept = packed_of(mask)
    if kept is not None:
        tag_packed(out, kept)
    return out


def _placement_program(ctx, Zl, dev):
    h = yield from ctx.all_gather(to_device_async([Zl], torch.int64, dev))
    hs = [int(v) for v in read_back(torch.stack(h).reshape(-1))]
    return sum(hs[: ctx.rank]), sum(hs)


def _refnodes_program(ctx, m):
    """Barycentres of the six GLOBAL boundary node sets (voxelFE.py:254-268).  float64 sums are order
    dependent: the running sums travel up the slabs (N - 1 tiny shifts), every rank continuing the
    accumulation in ascending node order where the rank below stopped."""
    dev = m.vol.device
    with torch.cuda.device(dev):
        tabs = [torch.from_numpy(np.asarray(a, dtype=np.float64)).to(dev) for a in (m.xs, m.ys, m.zs)]
        first = to_device_async(m.set_first[:6], torch.int64, dev)
        count = to_device_async(m.set_count[:6], torch.int64, dev)
        acc = torch.zeros((6, 3), dtype=torch.float64, device=dev)

        def accumulate():
            _lib.call("cfe_refnode_sums", m.set_ids.data_ptr(), first.data_ptr(), count.data_ptr(), 6, m.Y, m.X,
                      tabs[0].data_ptr(), tabs[1].data_ptr(), tabs[2].data_ptr(), acc.data_ptr(), _lib.stream())

        if ctx.rank == 0:
            accumulate()
    for hop in range(1, ctx.size):
        below = yield from ctx.shift_up(acc)
        if ctx.rank == hop:
            with torch.cuda.device(dev):
                acc.copy_(below)
                accumulate()
    accs = yield from ctx.all_gather(acc)
    sums = accs[ctx.size - 1].cpu().numpy()
    out = {}
    for q, key in enumerate(["X0", "X1", "Y0", "Y1", "Z0", "Z1"]):
        v = sums[q].copy()
        v /= m.set_count_global[q] if m.set_count_global[q] else 1
        out[key] = v
    return out


def vol2ugrid(voldata, voxelsize=[1, 1, 1], GVmin=0, plane_lock_num=1, node_level_lock_num=1,
              locking_strategy="plane", refnodes=False, verbose=False, *, group=None, z0=None, Z_global=None):
    """``vol2ugrid`` of the GLOBAL volume; ``voldata`` is this rank's z-slab (bool / uint8, NumPy or CUDA
    tensor).  Returns this rank's :class:`sharded.SlabMesh` (global numbering, device resident) and, with
    ``refnodes``, the dictionary of the six global reference nodes."""
    if verbose:
        logging.basicConfig(level=logging.INFO)
    if locking_strategy not in ("plane", "exact"):
        raise ValueError("locking_strategy must be 'plane' or 'exact'")
    if isinstance(voxelsize, tuple):
        raise TypeError("voxelsize must be a list, a numpy array or a scalar (the reference mis-handles tuples)")
    vol, is_bool, _ = as_device_volume(voldata)
    packed = packed_of(voldata)
    r = runner(group)
    if z0 is None or Z_global is None:
        z0, Z_global = r.run(_placement_program(r.ctx(), int(vol.shape[0]), vol.device))
    logging.info('Calculating cell array')
    mesh = r.run(sharded.vol2ugrid_program(r.ctx(), vol, is_bool, int(Z_global), int(z0), voxelsize, GVmin=GVmin,
                                           plane_lock_num=plane_lock_num,
                                           node_level_lock_num=node_level_lock_num,
                                           locking_strategy=locking_strategy, packed=packed))
    logging.info('Generated the following mesh with {0} nodes and {1} elements:'.format(
        (mesh.Zg + 1) * (mesh.Y + 1) * (mesh.X + 1), mesh.n_el_global))
    if refnodes:
        return mesh, r.run(_refnodes_program(r.ctx(), mesh))
    return mesh


def deck_to_device(mesh, templatefile, matprop=None, keywords=['NSET', 'ELSET'], eltype='C3D8', matpropbits=8,
                   refnode=None, group=None):
    """This rank's pieces of the global deck, formatted into HBM (:class:`sharded.SlabDeck`: device buffers +
    the file offset of every piece; rank 0 also holds the literal lines)."""
    r = runner(group)
    return r.run(sharded.deck_program(r.ctx(), mesh, templatefile, matprop=matprop, keywords=keywords,
                                      eltype=eltype, matpropbits=matpropbits, refnode=refnode))


def mesh2voxelfe(mesh, templatefile, fileout, matprop=None, keywords=['NSET', 'ELSET'], eltype='C3D8',
                 matpropbits=8, refnode=None, verbose=False, *, group=None):
    """``mesh2voxelfe`` of the global mesh: all ranks write their byte ranges of ONE ``.INP`` file at
    ``fileout`` (a path every rank can open; or, additively, any object with ``pwrite(buffer, offset)``).
    Returns the size of the whole deck in bytes."""
    if verbose:
        logging.basicConfig(level=logging.INFO)
    if not isinstance(mesh, sharded.SlabMesh):
        raise TypeError("distributed.mesh2voxelfe needs the SlabMesh of distributed.vol2ugrid")
    logging.info('Start writing INP file')
    r = runner(group)
    deck = deck_to_device(mesh, templatefile, matprop=matprop, keywords=keywords, eltype=eltype,
                          matpropbits=matpropbits, refnode=refnode, group=group)
    n = r.run(sharded.write_deck_program(r.ctx(), deck, fileout))
    logging.info('Model with {0} nodes and {1} elements written to file {fname}'.format(
        (mesh.Zg + 1) * (mesh.Y + 1) * (mesh.X + 1), mesh.n_el_global, fname=fileout))
    return n
