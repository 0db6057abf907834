"""Array-driven deck emission for ANY hexahedral meshio-like mesh.

``ciclope.core.voxelFE.mesh2voxelfe`` (reference: src/ciclope/core/voxelFE.py:474-772) only reads
``mesh.points``, ``mesh.cells[0][1]``, ``mesh.cells_dict["hexahedron"]``, the first ``cell_data`` entry,
``mesh.point_sets`` and ``mesh.cell_sets`` -- it works for a mesh read back from VTK as well as for one
made by ``vol2ugrid``.  The volume-driven fast path of ``ciclope_b200.voxelFE`` needs the device state of
its own lazy mesh; everything else (foreign meshes, lazy meshes whose host-side sets or grey values
were edited) comes here: the arrays are uploaded once and the same text is produced by plain CUDA
kernels (lengths -> tile totals -> scan -> staged emission).  No CPU fallback.
"""
import logging
import warnings
from datetime import datetime

import numpy as np
import torch

from . import _lib
from ._lib import CfeItemSeg
from ._util import bytes_to_device_async, empty_u32, scan_workspace, to_device_async


def _h2d(a, dev):
    """Host array -> device tensor (read-only arrays, e.g. VoxelMesh's materialised views, included)."""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", UserWarning)     # "The given NumPy array is not writable": only read
        return torch.from_numpy(a).to(dev)


def _as_torch(a, dev):
    a = np.ascontiguousarray(a)
    if a.dtype == np.bool_:
        return _h2d(a.view(np.uint8), dev)
    if a.dtype.kind == "u" and a.dtype.itemsize > 1:
        a = a.astype(np.int64)
    if a.dtype.kind == "f" and a.dtype.itemsize == 2:
        a = a.astype(np.float32)
    return _h2d(a, dev)


def deck_generic(mesh, templatefile, matprop=None, keywords=['NSET', 'ELSET'], eltype='C3D8', matpropbits=8,
                 refnode=None):
    """The bytes ``mesh2voxelfe`` writes for `mesh`, built on the GPU.  Returns a uint8 CUDA tensor."""
    from . import voxelFE as vfe
    _lib.require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device())

    # ---- prologue (voxelFE.py:540-610)
    keys = list(mesh.cell_data)
    if not keys:
        raise ValueError("mesh has no cell data: the element sets of the deck are made from its first entry")
    cell_GV = mesh.cell_data.get(keys[0])
    if type(cell_GV) is list:
        cell_GV = cell_GV[0]
    cell_GV = np.asarray(cell_GV)
    if cell_GV.size == 0:
        raise ValueError("min() iterable argument is empty")            # voxelFE.py:564
    if cell_GV.dtype.kind not in "biuf":
        raise TypeError("cell data of dtype %s cannot be mapped to element sets" % cell_GV.dtype)
    cells0 = np.ascontiguousarray(np.asarray(mesh.cells[0][1]), dtype=np.int64)      # voxelFE.py:647
    if cells0.ndim != 2 or cells0.shape[1] != 8:
        raise ValueError("mesh2voxelfe needs 8-node hexahedra, got cells of shape %s" % (cells0.shape,))
    if len(cell_GV) != len(cells0):
        raise IndexError("cell data and cells have different lengths")
    allhex = np.ascontiguousarray(np.asarray(mesh.cells_dict["hexahedron"]), dtype=np.int64)   # voxelFE.py:628
    points = np.asarray(mesh.points)
    if points.ndim != 2 or points.shape[1] < 3:
        raise ValueError("mesh.points must be (n, 3)")
    if points.dtype.kind in "iu" and points.size and int(np.abs(points).max()) >= 2 ** 53:
        raise NotImplementedError("integer node coordinates beyond 2^53 are not supported")
    if allhex.size and (allhex.min() < 0 or allhex.max() >= len(points)):
        raise IndexError("cell node index out of range of mesh.points")

    with torch.cuda.device(dev):
        s = _lib.stream()
        gv_dev = _as_torch(cell_GV, dev)
        uniq, inv = torch.unique(gv_dev, sorted=True, return_inverse=True)      # np.unique(cell_GV), :570
        counts = torch.bincount(inv, minlength=uniq.numel())
        order = torch.sort(inv, stable=True).indices                            # np.where(cell_GV == GV)[0], :645
        group = inv[order].to(torch.int32)
        gfirst = torch.cumsum(counts, 0) - counts
        uniq_host = uniq.cpu().numpy()
    existing_GV = uniq_host.astype(np.bool_) if cell_GV.dtype == np.bool_ else uniq_host.astype(cell_GV.dtype)
    GVmax = 2 ** matpropbits - 1
    binary_model = bool(existing_GV.min() == 0) and bool(existing_GV.max() == 1) and \
        issubclass(cell_GV.dtype.type, np.integer)
    if binary_model:
        logging.warning('cell_data is binary. Material property mapping disabled.')
    if max(existing_GV) > GVmax:
        GVmax = max(existing_GV)
    prop_GV = {0: (1, 1)} if binary_model else vfe._resolve_matprop(matprop, GVmax)

    head = ('** ---------------------------------------------------------\n'
            '** Abaqus .INP file written by Voxel_FEM script on {}\n'
            '** ---------------------------------------------------------\n'
            '** Node coordinates from input model\n'
            '*NODE\n').format(datetime.now()).encode()
    elem_comment = b'** Elements and Element sets from input model\n'
    hdr_lit = bytearray()
    hdr_off, hdr_len = [], []
    for GV in existing_GV:
        h = '*ELEMENT, TYPE={0}, ELSET={1}\n'.format(eltype, 'SET' + str(int(GV))).encode()
        hdr_off.append(len(hdr_lit))
        hdr_len.append(len(h))
        hdr_lit.extend(h)

    # NSET / ELSET item stream (voxelFE.py:650-688)
    lits = bytearray()
    segs = []
    id_chunks = []
    n_items = 0
    n_ids = 0

    def literal(text):
        nonlocal n_items
        t = text.encode()
        segs.append((0, n_items, len(lits), len(t), 0, 0))
        lits.extend(t)
        n_items += 1

    def idlist(values):
        nonlocal n_items, n_ids
        a = np.asarray(values, dtype=np.int64).reshape(-1)
        if a.size:
            if a.min() < 0:
                raise ValueError("negative ids in a node / element set")
            segs.append((1, n_items, 0, 0, n_ids, 0))
            id_chunks.append(a)
            n_items += a.size
            n_ids += a.size

    max_id = 0
    if 'NSET' in keywords:
        literal('** Additional nset from voxel model. New generated nsets are:\n')
        literal('** {}, \n'.format(*list(mesh.point_sets)))
        for name in list(mesh.point_sets):
            literal('*NSET, NSET={0}\n'.format(name))
            idlist(mesh.point_sets[name])
            literal('\n')
    if 'ELSET' in keywords:
        literal('** Additional elset from voxel model. New generated elsets are:\n')
        literal('** {}, \n'.format(*list(mesh.cell_sets)))
        for name in list(mesh.cell_sets):
            literal('*ELSET, ELSET={}\n'.format(name))
            idlist(mesh.cell_sets[name])
            literal('\n')
    ids_host = np.concatenate(id_chunks) if id_chunks else np.zeros(1, dtype=np.int64)
    if id_chunks:
        max_id = int(ids_host.max())
    seg_arr = (CfeItemSeg * max(len(segs), 1))()
    for i, sg in enumerate(segs):
        a = seg_arr[i]
        a.kind, a.item0, a.lit_off, a.lit_len, a.id0, a.rank0 = sg
    tail_bound = vfe._tail_bound(templatefile, refnode, matprop if 'PROPERTY' in keywords else None, len(existing_GV))

    with torch.cuda.device(dev):
        s = _lib.stream()
        n_el = len(cells0)
        cells_dev = _h2d(cells0, dev)
        used = torch.unique(_h2d(allhex.reshape(-1), dev))       # ascending
        n_nd = used.numel()
        pts = _h2d(np.ascontiguousarray(points[:, :3], dtype=np.float64), dev)
        vals = pts[used].contiguous()                                            # (n_nd, 3)
        slots = torch.empty(96 * max(n_nd, 1), dtype=torch.uint8, device=dev)
        stat = torch.zeros(3, dtype=torch.int64, device=dev)                     # node bytes, element bytes, bad
        _lib.call("cfe_format_coords_wide", vals.data_ptr(), 3 * n_nd, slots.data_ptr(), stat.data_ptr() + 16, s)
        # *NODE lines
        nt_n = (n_nd + 255) // 256
        nlen8 = torch.empty(max(n_nd, 1), dtype=torch.uint8, device=dev)
        ntot = empty_u32(nt_n, dev)
        noff = torch.empty(max(nt_n, 1), dtype=torch.int64, device=dev)
        _lib.call("cfe_gen_node_lengths", used.data_ptr(), n_nd, slots.data_ptr(), nlen8.data_ptr(), ntot.data_ptr(), s)
        _lib.call("cfe_scan_u32", ntot.data_ptr(), nt_n, noff.data_ptr(), stat.data_ptr(),
                  scan_workspace(max(nt_n, 1), dev).data_ptr(), s)
        # *ELEMENT lines
        d_hoff = to_device_async(hdr_off, torch.int32, dev)
        d_hlen = to_device_async(hdr_len, torch.int32, dev)
        d_hlit = bytes_to_device_async(bytes(hdr_lit), dev)
        nt_e = (n_el + 255) // 256
        elen = torch.empty(max(n_el, 1), dtype=torch.int16, device=dev)
        etot = empty_u32(nt_e, dev)
        eoff = torch.empty(max(nt_e, 1), dtype=torch.int64, device=dev)
        _lib.call("cfe_gen_elem_lengths", cells_dev.data_ptr(), order.data_ptr(), group.data_ptr(), gfirst.data_ptr(),
                  d_hoff.data_ptr(), d_hlen.data_ptr(), n_el, elen.data_ptr(), etot.data_ptr(), s)
        _lib.call("cfe_scan_u32", etot.data_ptr(), nt_e, eoff.data_ptr(), stat.data_ptr() + 8,
                  scan_workspace(max(nt_e, 1), dev).data_ptr(), s)
        node_bytes, elem_bytes, bad = stat.cpu().tolist()
        if bad:
            raise NotImplementedError("node coordinates that are not finite or not below 1e18 cannot be formatted")
        off_nodes = len(head)
        off_cmt = off_nodes + node_bytes
        off_elem = off_cmt + len(elem_comment)
        off_items = off_elem + elem_bytes
        cap = off_items + n_ids * (len(str(max_id)) + 1) + len(lits) + tail_bound + 64
        out = torch.empty(cap, dtype=torch.uint8, device=dev)
        d_head = bytes_to_device_async(head + elem_comment, dev)
        _lib.call("cfe_put_literal", d_head.data_ptr(), len(head), out.data_ptr(), None, None, s)
        _lib.call("cfe_put_literal", d_head.data_ptr() + len(head), len(elem_comment), out.data_ptr() + off_cmt,
                  None, None, s)
        _lib.call("cfe_gen_emit_nodes", used.data_ptr(), n_nd, slots.data_ptr(), nlen8.data_ptr(), noff.data_ptr(),
                  out.data_ptr() + off_nodes, s)
        _lib.call("cfe_gen_emit_elements", cells_dev.data_ptr(), order.data_ptr(), group.data_ptr(),
                  gfirst.data_ptr(), d_hoff.data_ptr(), d_hlen.data_ptr(), d_hlit.data_ptr(), n_el, elen.data_ptr(),
                  eoff.data_ptr(), out.data_ptr() + off_elem, s)
        d_ids = _h2d(ids_host, dev)
        d_lits = bytes_to_device_async(bytes(lits), dev)
        d_segs = bytes_to_device_async(bytes(seg_arr), dev)
        itot = torch.zeros(1, dtype=torch.int64, device=dev)
        _lib.call("cfe_emit_items", d_segs.data_ptr(), len(segs), n_items, d_ids.data_ptr(), d_lits.data_ptr(),
                  out.data_ptr() + off_items, None, itot.data_ptr(), None,
                  scan_workspace(max(n_items, 1), dev).data_ptr(), s)
        # host text while the GPU formats (voxelFE.py:691-768)
        tail = ''
        if 'PROPERTY' in keywords:
            tail += vfe._property_cards(matprop, prop_GV, existing_GV)
        tail += vfe._template_text(templatefile, refnode)
        tail = tail.encode()
        if len(tail) > tail_bound:
            raise _lib.CfeError("internal error: PROPERTY/template text exceeds its reserved bound")
        d_tail = bytes_to_device_async(tail, dev)
        _lib.call("cfe_put_literal", d_tail.data_ptr(), len(tail), out.data_ptr() + off_items, itot.data_ptr(), None, s)
        items_bytes = int(itot.cpu().item())
    length = off_items + items_bytes + len(tail)
    if length > cap:
        raise _lib.CfeError("internal error: deck length %d exceeds the reserved capacity %d" % (length, cap))
    return out[:length]
