"""B200-native drop-in for the voxel-FE hot path of ``ciclope.core.voxelFE``:

* ``vol2ugrid``          (reference: src/ciclope/core/voxelFE.py:15-283)
* ``mesh2voxelfe``       (reference: src/ciclope/core/voxelFE.py:474-772)
* ``matpropdictionary``  (reference: src/ciclope/core/voxelFE.py:774-819)

Signatures, defaults, argument meaning and error behaviour are the reference's.  All voxel,
node and text loops run as hand-written sm_100a CUDA kernels behind the C ABI of
``include/ciclope_b200.h``; the mesh object is *lazy* (device-resident numbering, host arrays
only when an attribute is read) because ``points``/``cells`` of a 1024^3 volume do not fit a
host.  Host Python keeps only what is O(#grey values): the deck header, the PROPERTY cards
(the reference evaluates them with ``eval`` on NumPy scalars, voxelFE.py:729-736, so the same
interpreter arithmetic must produce the digits) and the template copy.

There is no CPU fallback: without the CUDA library or a CUDA device every function raises.
"""
import collections
import logging
import math
import os
from datetime import datetime

import numpy as np
import torch

from . import _lib
from ._lib import CfeItemSeg, CfeSetDesc
from ._util import (as_device_volume, bytes_to_device_async, check_slab_size, empty_u32, packed_of, read_back,
                    scan_workspace, to_device_async)

CellBlock = collections.namedtuple("CellBlock", ["type", "data"])  # meshio 5.0.0 CellBlock

NODE_SET_NAMES = ["NODES_X0", "NODES_X1", "NODES_Y0", "NODES_Y1", "NODES_Z0", "NODES_Z1"]
CELL_SET_NAMES = ["CELLS_X0", "CELLS_X1", "CELLS_Y0", "CELLS_Y1", "CELLS_Z0", "CELLS_Z1"]
_ELEM_HDR_SLOT = 64


# --------------------------------------------------------------------------------------------
# host helpers

def _gv_threshold(GVmin):
    """Integer t such that (GV > GVmin) <=> (GV > t) for 8-bit GV (voxelFE.py:140-141)."""
    if GVmin < 0:
        return -1
    return int(min(math.floor(GVmin), 255))


def _coordinate_tables(voxelsize, Z, Y, X):
    """Per-axis coordinate values computed by the reference's own expressions
    (voxelFE.py:70-73, 202) and converted to the dtype ``np.asarray(nodes)`` gives."""
    if not isinstance(voxelsize, (list, np.ndarray)):
        voxelsize = [voxelsize, voxelsize, voxelsize]
    if len(voxelsize) == 1:
        voxelsize = [voxelsize[0]] * 3
    xs = [voxelsize[0] * c for c in range(X + 1)]
    ys = [voxelsize[1] * r for r in range(Y + 1)]
    zs = [voxelsize[2] * s for s in range(Z + 1)]
    dtype = np.asarray([[xs[0], ys[0], zs[0]], [xs[-1], ys[-1], zs[-1]]]).dtype
    if dtype.kind not in "fiu" or dtype.itemsize not in (4, 8):
        raise TypeError("unsupported voxelsize type (points dtype %s)" % dtype)
    return tuple(np.asarray(v).astype(dtype) for v in (xs, ys, zs))


def _format_tables_device(xs, ys, zs, dev, overflow_ptr, stream):
    """'{:12.6f}' fields of every per-axis coordinate, formatted exactly on the GPU (16-byte slots,
    x then y then z).  The values are float64 copies of what np.asarray(nodes) holds, which is what
    format() sees (np.float32 / np.int64 scalars are widened exactly by their __format__)."""
    vals = []
    for a in (xs, ys, zs):
        if a.dtype.kind in "iu" and a.size and int(np.abs(a).max()) >= 2 ** 53:
            raise NotImplementedError("integer node coordinates beyond 2^53 are not supported")
        vals.append(np.asarray(a, dtype=np.float64))
    tab = torch.from_numpy(np.concatenate(vals)).pin_memory().to(dev, non_blocking=True)
    out = torch.empty(16 * tab.numel(), dtype=torch.uint8, device=dev)
    _lib.call("cfe_format_coords", tab.data_ptr(), tab.numel(), out.data_ptr(), overflow_ptr, stream)
    return out, tab


_TABLES = collections.OrderedDict()


def _tables_key(voxelsize, Z, Y, X, dev, stream):
    """Cache key of the per-axis coordinate tables: they are a pure function of the voxel size (value
    AND type: np.float32 / int voxel sizes give float32 / int64 points) and of the grid shape."""
    if isinstance(voxelsize, np.ndarray):
        vs = ("nd", str(voxelsize.dtype), voxelsize.shape, voxelsize.tobytes())
    elif isinstance(voxelsize, list):
        vs = tuple((type(v).__name__, repr(v)) for v in voxelsize)
    else:
        vs = (type(voxelsize).__name__, repr(voxelsize))
    return (vs, Z, Y, X, dev.index, stream)


def _int_lock(v, name):
    if isinstance(v, (bool, np.bool_)) or int(v) != v:
        raise TypeError("%s must be an integer" % name)
    return int(v)


def _clip_box(lo, hi, n):
    lo, hi = max(lo, 0), min(hi, n)
    return (lo, hi) if hi > lo else (0, 0)


def _boundary_boxes(strategy, p, k, Z, Y, X):
    """Voxel boxes of the six locked strips ("plane", voxelFE.py:163-187) or node boxes of the
    six locked levels ("exact", voxelFE.py:228-245), order X0, X1, Y0, Y1, Z0, Z1.
    Each box is ((s0, s1), (r0, r1), (c0, c1)), half-open."""
    full = ((0, Z), (0, Y), (0, X))
    boxes = []
    if strategy == "plane":
        for axis, n in ((2, X), (1, Y), (0, Z)):
            for lo, hi in ((0, p), (n - p, n)):          # idx < p ; idx >= n - p
                b = list(full)
                b[axis] = _clip_box(lo, hi, n)
                boxes.append(tuple(b))
    else:
        fulln = ((0, Z + 1), (0, Y + 1), (0, X + 1))
        for axis, n in ((2, X), (1, Y), (0, Z)):
            for lo, hi in ((0, k), (n - k + 1, n + 1)):  # idx < k ; idx > n - k
                b = list(fulln)
                b[axis] = _clip_box(lo, hi, n + 1)
                boxes.append(tuple(b))
    return boxes


def _box_count(b):
    return max(b[0][1] - b[0][0], 0) * max(b[1][1] - b[1][0], 0) * max(b[2][1] - b[2][0], 0)


# --------------------------------------------------------------------------------------------
# lazy mesh

class VoxelMesh:
    """Device-resident hexahedral voxel mesh with the meshio-5.0.0 surface the reference and its
    tests touch (voxelFE.py:271-274): ``points``, ``cells[0].type/.data/[0]/[1]``, ``cells_dict``,
    ``cell_data["GV"][0]``, ``point_sets``, ``cell_sets``.  Host arrays are materialised by CUDA
    kernels the first time an attribute is read."""

    def __init__(self):
        self._points = None
        self._cells = None
        self._cell_data = None
        self._point_sets = None
        self._cell_sets = None
        # what the lazy getters handed out (a set-before-read then compares as edited)
        self._points_made = None
        self._cells_made = None
        self._cells_block = None

    # -- sizes ------------------------------------------------------------------------------
    @property
    def n_points(self):
        return (self.Z + 1) * (self.Y + 1) * (self.X + 1)

    @property
    def n_cells(self):
        return self.n_el

    @property
    def node_list(self):
        """Local grid index of every used node, ascending (u32).  The *NODE emitter expands the used-node
        bitmap itself; only the variable-width coordinate path still reads this list, built on first use."""
        if getattr(self, "_node_list", None) is None:
            dev = self.vol.device
            with torch.cuda.device(dev):
                self._node_list = empty_u32(self.n_nd, dev)
                _lib.call("cfe_fill_list", self.nbits.data_ptr(), self.npref.data_ptr(), self.n_node_words,
                          (self.X + 32) // 32, self.X + 1, self._node_list.data_ptr(), _lib.stream())
        return self._node_list

    # -- materialisers ----------------------------------------------------------------------
    @property
    def points(self):
        if self._points is None:
            dev = self.vol.device
            dt = self.xs.dtype
            tdt = {4: torch.int32, 8: torch.int64}[dt.itemsize]
            with torch.cuda.device(dev):
                tabs = [torch.from_numpy(np.ascontiguousarray(a).view({4: np.int32, 8: np.int64}[dt.itemsize])).to(dev)
                        for a in (self.xs, self.ys, self.zs)]
                out = torch.empty((self.n_points, 3), dtype=tdt, device=dev)
                _lib.call("cfe_fill_points", tabs[0].data_ptr(), tabs[1].data_ptr(), tabs[2].data_ptr(),
                          dt.itemsize, self.Z + 1, self.Y, self.X, out.data_ptr(), _lib.stream())
                self._points = out.cpu().numpy().view(dt)
                self._points.flags.writeable = False      # replace, do not edit in place (_host_edited)
                self._points_made = self._points
        return self._points

    @points.setter
    def points(self, value):
        self._points = value

    @property
    def cells(self):
        if self._cells is None:
            dev = self.vol.device
            with torch.cuda.device(dev):
                out = torch.empty((self.n_el, 8), dtype=torch.int64, device=dev)
                _lib.call("cfe_fill_cells", self.elem_vox.data_ptr(), self.n_el, self.Y, self.X, 0,
                          out.data_ptr(), _lib.stream())
                data = out.cpu().numpy()
                data.flags.writeable = False
                self._cells = self._cells_made = [CellBlock("hexahedron", data)]
                self._cells_block = self._cells[0]
        return self._cells

    @cells.setter
    def cells(self, value):
        self._cells = value

    @property
    def cells_dict(self):
        d = {}
        for blk in self.cells:
            d.setdefault(blk[0], []).append(np.asarray(blk[1]))
        return {k: (v[0] if len(v) == 1 else np.concatenate(v)) for k, v in d.items()}

    def _gv_array(self):
        dev = self.vol.device
        with torch.cuda.device(dev):
            if getattr(self, "gv8", None) is not None:
                out = self.gv8
            else:
                out = torch.empty(max(self.n_el, 1), dtype=torch.uint8, device=dev)
                _lib.call("cfe_gather_gv", self.vol.data_ptr(), self.elem_vox.data_ptr(), self.n_el,
                          out.data_ptr(), _lib.stream())
            a = out[: self.n_el].cpu().numpy()
        return a.view(np.bool_) if self.is_bool else a

    @property
    def cell_data(self):
        if self._cell_data is None:
            self._gv_host = self._gv_array()
            self._cell_data = {"GV": [self._gv_host]}
        return self._cell_data

    @cell_data.setter
    def cell_data(self, value):
        self._cell_data = value

    def _set_lists(self, first, count):
        ids = self.set_ids
        out = []
        for f, c in zip(first, count):
            out.append(ids[f:f + c].cpu().numpy().tolist() if c else [])
        return out

    @property
    def point_sets(self):
        if self._point_sets is None:
            lists = self._set_lists(self.set_first[:6], self.set_count[:6])
            d = dict(zip(NODE_SET_NAMES, lists))
            d["NODES_X0Y0Z0"] = d["NODES_Z0"][0:2]   # voxelFE.py:250-251
            d["NODES_X0Y0Z1"] = d["NODES_Z1"][0:2]
            self._point_sets = d
        return self._point_sets

    @property
    def cell_sets(self):
        if self._cell_sets is None:
            lists = self._set_lists(self.set_first[6:], self.set_count[6:])
            self._cell_sets = dict(zip(CELL_SET_NAMES, lists))
        return self._cell_sets

    def _host_edited(self):
        """True when a host-side view (sets, grey values, points, cells) no longer shows what
        vol2ugrid computed; mesh2voxelfe then emits from the host arrays (generic.py) instead of
        the device state.  The big materialised arrays are handed out read-only, so they can only
        be replaced (seen by identity); sets and grey values are compared by content."""
        if self._points is not None and self._points is not self._points_made:
            return True
        if self._cells is not None and (self._cells is not self._cells_made or len(self._cells) != 1 or
                                        self._cells[0] is not self._cells_block):
            return True
        ids = None
        for d, names, first, count in ((self._point_sets, NODE_SET_NAMES + ["NODES_X0Y0Z0", "NODES_X0Y0Z1"],
                                        self.set_first[:6], self.set_count[:6]),
                                       (self._cell_sets, CELL_SET_NAMES, self.set_first[6:], self.set_count[6:])):
            if d is None:
                continue
            if list(d) != names:
                return True
            if ids is None:
                ids = self.set_ids.cpu().numpy()
            for q, name in enumerate(names):
                if q < 6:
                    want = ids[first[q]:first[q] + count[q]] if count[q] else ids[:0]
                else:   # corner sets: first two entries of Z0 / Z1
                    want = ids[first[q - 2]:first[q - 2] + min(2, count[q - 2])] if count[q - 2] else ids[:0]
                got = np.asarray(d[name])
                if got.shape != want.shape or got.dtype.kind not in "iu" or not np.array_equal(got, want):
                    return True
        if self._cell_data is not None:
            keys = list(self._cell_data)
            if not keys:
                return True
            cell_GV = self._cell_data.get(keys[0])
            if type(cell_GV) is list:
                cell_GV = cell_GV[0] if len(cell_GV) else None
            if cell_GV is None:
                return True
            cell_GV = np.asarray(cell_GV)
            # the notebooks' cell_data['GV'] = cell_data['GV'][0].astype('uint8') keeps the values
            if cell_GV.dtype not in (np.bool_, np.uint8) or cell_GV.shape != (self.n_el,) or \
                    not np.array_equal(cell_GV, self._gv_array()):
                return True
        return False

    # -- export ------------------------------------------------------------------------------
    def to_meshio(self):
        """The mesh as a ``meshio.Mesh`` (points, one hexahedron block, GV cell data, point / cell sets):
        what the reference's ``vol2ugrid`` returns (voxelFE.py:271-274).  Needs meshio and host memory for the
        materialised arrays (24 B per grid node + 64 B per element)."""
        import meshio
        return meshio.Mesh(self.points, [(blk[0], np.asarray(blk[1])) for blk in self.cells],
                           cell_data={k: list(v) if isinstance(v, list) else [np.asarray(v)]
                                      for k, v in self.cell_data.items()},
                           point_sets=self.point_sets, cell_sets=self.cell_sets)

    def write(self, path, file_format=None, **kwargs):
        """``mesh.write(path)`` of the reference's notebooks and pipeline (pipeline.py:222-233): through meshio
        when it is installed (every format meshio knows); otherwise ``.vtk`` files are written by the built-in
        legacy-VTK writer (points, hexahedra, cell data) and other formats raise ImportError."""
        try:
            import meshio  # noqa: F401
        except ImportError:
            if file_format not in (None, "vtk") or not str(path).lower().endswith(".vtk"):
                raise ImportError("meshio is not installed: only legacy .vtk files can be written")
            from .io import write_vtk
            cells = np.concatenate([np.asarray(blk[1]).reshape(-1, 8) for blk in self.cells]) if self.cells else \
                np.empty((0, 8), dtype=np.int64)
            return write_vtk(path, self.points, cells, self.cell_data)
        return self.to_meshio().write(path, file_format=file_format, **kwargs)

    def __repr__(self):
        return "<ciclope_b200 voxel mesh>\n  Number of points: {}\n  Number of cells:\n    hexahedron: {}\n" \
               "  Point sets: {}\n  Cell sets: {}\n  Cell data: GV".format(
                   self.n_points, self.n_el, ", ".join(NODE_SET_NAMES + ["NODES_X0Y0Z0", "NODES_X0Y0Z1"]),
                   ", ".join(CELL_SET_NAMES))


# --------------------------------------------------------------------------------------------
# vol2ugrid

def vol2ugrid(voldata, voxelsize=[1, 1, 1], GVmin=0, plane_lock_num=1, node_level_lock_num=1,
              locking_strategy="plane", refnodes=False, verbose=False):
    """Generate unstructured grid mesh from 3D volume data (drop-in for voxelFE.py:15-283).

    Parameters are the reference's.  ``voldata`` may be a NumPy array or a (CUDA) torch tensor
    of dtype bool or uint8.  Returns a :class:`VoxelMesh` (and the reference-node dictionary
    when ``refnodes`` is True).
    """
    if verbose:
        logging.basicConfig(level=logging.INFO)
    if locking_strategy not in ("plane", "exact"):
        raise ValueError("locking_strategy must be 'plane' or 'exact'")
    if isinstance(voxelsize, tuple):
        raise TypeError("voxelsize must be a list, a numpy array or a scalar (the reference mis-handles tuples)")
    p = _int_lock(plane_lock_num, "plane_lock_num")
    k = _int_lock(node_level_lock_num, "node_level_lock_num")
    vol, is_bool, _ = as_device_volume(voldata)
    Z, Y, X = (int(v) for v in vol.shape)
    check_slab_size(Z, Y, X)
    thr = _gv_threshold(GVmin)
    dev = vol.device
    # the mask of remove_unconnected arrives with its voxels already bit-packed (GV > 0  <=>  mask)
    packed = packed_of(voldata) if thr == 0 and vol.dim() == 3 else None

    m = VoxelMesh()
    m.vol, m.is_bool, m.thr = vol, is_bool, thr
    m.Z, m.Y, m.X = Z, Y, X
    m.strategy, m.plane_lock_num, m.node_level_lock_num = locking_strategy, p, k

    logging.info('Calculating cell array')
    with torch.cuda.device(dev):
        s = _lib.stream()
        W, WN = (X + 31) // 32, (X + 32) // 32
        rows, nrows = Z * Y, (Z + 1) * (Y + 1)
        # the first kernel goes out before anything else is allocated: the GPU is idle until then
        if packed is not None and packed.numel() >= rows * W:
            bits = packed
        else:
            bits = empty_u32(rows * W, dev)
            _lib.call("cfe_pack_bits", vol.data_ptr(), rows, X, thr, bits.data_ptr(), s)
        epref = empty_u32(rows * W, dev)
        nbits = empty_u32(nrows * WN, dev)
        npref = empty_u32(nrows * WN, dev)
        ws = scan_workspace(max(rows * W, nrows * WN), dev)
        # n_el, n_nd, -, field overflow, then first[0..12] (position of every set's first id; [12] = n_ids)
        stats = torch.zeros(17, dtype=torch.int64, device=dev)
        _lib.call("cfe_scan_elements", bits.data_ptr(), rows * W, epref.data_ptr(), stats.data_ptr(),
                  ws.data_ptr(), s)
        chunk_word = empty_u32((Z + 1) * (Y + 1) * (X + 1) // 512 + 2, dev)   # rank index of the *NODE emitter
        _lib.call("cfe_scan_nodes_index", bits.data_ptr(), None, Z, Y, X, Z + 1, nbits.data_ptr(), npref.data_ptr(),
                  stats.data_ptr() + 8, chunk_word.data_ptr(), ws.data_ptr(), s)
        # per-axis coordinate values and their text fields (memoised: a pure function of voxel size and shape)
        tkey = _tables_key(voxelsize, Z, Y, X, dev, s)
        tables = _TABLES.get(tkey)
        if tables is None:
            m.xs, m.ys, m.zs = _coordinate_tables(voxelsize, Z, Y, X)
            m.ftab, m.coord_vals = _format_tables_device(m.xs, m.ys, m.zs, dev, stats.data_ptr() + 24, s)
        else:
            m.xs, m.ys, m.zs, m.ftab, m.coord_vals, cached_overflow = tables
        # (host work from here to the read-back overlaps the kernels above)
        # boundary-set descriptors: node sets X0..Z1 then cell sets X0..Z1
        boxes = _boundary_boxes(locking_strategy, p, k, Z, Y, X)
        descs = (CfeSetDesc * 12)()
        set_tile = _lib.load().cfe_sets_onepass_tile()
        tile_base = 0
        n_cand = 0
        for q in range(12):
            d = descs[q]
            b = boxes[q % 6]
            if locking_strategy == "plane":
                if q < 6:      # nodes of the locked voxel strip: both node layers of every voxel
                    d.kind = 0
                    cand = tuple((lo, hi + 1) if hi > lo else (0, 0) for lo, hi in b)
                    vbox = b
                else:
                    d.kind = 2
                    cand = b
                    vbox = b
            else:
                if q < 6:      # existing nodes on the locked node levels
                    d.kind = 1
                    cand = b
                    vbox = ((0, Z), (0, Y), (0, X))
                else:          # voxelFE.py: the cell sets stay empty under "exact"
                    d.kind = 2
                    cand = ((0, 0), (0, 0), (0, 0))
                    vbox = cand
            (d.s0, d.s1), (d.r0, d.r1), (d.c0, d.c1) = cand
            (d.vs0, d.vs1), (d.vr0, d.vr1), (d.vc0, d.vc1) = vbox
            d.n_cand = _box_count(cand)
            d.tile_base = tile_base
            tile_base += (d.n_cand + set_tile - 1) // set_tile
            n_cand += d.n_cand
        n_tiles = tile_base
        descs_dev = bytes_to_device_async(bytes(descs), dev)
        set_ids = torch.empty(max(n_cand, 1), dtype=torch.int64, device=dev)
        ws_sets = scan_workspace(max(n_tiles, 1) * 256, dev)

        logging.info('Detecting node coordinates and boundary nodes')
        # one pass: predicates, look-back positions and the ids themselves (voxelFE.py:163-192, 205-247)
        _lib.call("cfe_sets_onepass", descs_dev.data_ptr(), 12, n_tiles, Z, Y, X, 0, 0, bits.data_ptr(),
                  epref.data_ptr(), None, ws_sets.data_ptr(), set_ids.data_ptr(), stats.data_ptr() + 32, s)
        # the one device->host read: counts + first id of every set
        host = read_back(stats)
        n_el, n_nd = host[0], host[1]
        if tables is None:
            m.field_overflow = bool(host[3])
            _TABLES[tkey] = (m.xs, m.ys, m.zs, m.ftab, m.coord_vals, m.field_overflow)
            while len(_TABLES) > 8:
                _TABLES.popitem(last=False)
        else:
            m.field_overflow = cached_overflow
        first = host[4:17]
        n_ids = first[12]
        for q in range(11, -1, -1):
            if first[q] < 0:             # a set without candidates starts where the next one does
                first[q] = first[q + 1]
        count = [first[q + 1] - first[q] for q in range(12)]
        first = first[:12]

        # element list fused with the *ELEMENT emitter's line-length pass
        lib = _lib.load()
        elem_vox = empty_u32(n_el, dev)
        ws_el = torch.empty(lib.cfe_emit_elements_workspace_bytes(n_el) + 16, dtype=torch.uint8, device=dev)
        # grey-scale volumes: the grey value of every element, compact and in element order (cell_GV)
        # (a bool volume has the one grey value True -- unless GVmin < 0 makes the False voxels elements too)
        m.gv8 = None if (is_bool and thr >= 0) else torch.empty(max(n_el, 1), dtype=torch.uint8, device=dev)
        _lib.call("cfe_fill_elements_gv", bits.data_ptr(), epref.data_ptr(), rows * W, Y, X, 0, 0, n_el,
                  elem_vox.data_ptr(), ws_el.data_ptr(), vol.data_ptr(), _lib.ptr(m.gv8), s)
        m.ws_el = ws_el
        m.bits, m.epref = bits, epref
        m.nbits, m.npref, m.chunk_word, m.n_node_words = nbits, npref, chunk_word, nrows * WN
        m.n_el, m.n_nd = int(n_el), int(n_nd)
        m.elem_vox = elem_vox
        if NODE_EMITTER != "bits":
            m.node_list                       # k_fill_list now, behind the element list on the stream

    m.set_ids, m.set_first, m.set_count = set_ids, [int(v) for v in first], [int(v) for v in count]

    logging.info('Generated the following mesh with {0} nodes and {1} elements:'.format(m.n_points, m.n_el))
    logging.info(m)
    if refnodes:
        return m, _refnodes(m)
    return m


def _refnodes(m):
    """Barycentres of the six boundary node sets (voxelFE.py:254-268), float64 sums accumulated on
    the GPU in the reference's (ascending node) order."""
    dev = m.vol.device
    with torch.cuda.device(dev):
        tabs = [torch.from_numpy(np.asarray(a, dtype=np.float64)).to(dev) for a in (m.xs, m.ys, m.zs)]
        first = to_device_async(m.set_first[:6], torch.int64, dev)
        count = to_device_async(m.set_count[:6], torch.int64, dev)
        acc = torch.zeros((6, 3), dtype=torch.float64, device=dev)
        _lib.call("cfe_refnode_sums", m.set_ids.data_ptr(), first.data_ptr(), count.data_ptr(), 6, m.Y, m.X,
                  tabs[0].data_ptr(), tabs[1].data_ptr(), tabs[2].data_ptr(), acc.data_ptr(), _lib.stream())
        sums = acc.cpu().numpy()
    out = {}
    for q, key in enumerate(["X0", "X1", "Y0", "Y1", "Z0", "Z1"]):
        v = sums[q].copy()
        v /= m.set_count[q] if m.set_count[q] else 1
        out[key] = v
    return out


# --------------------------------------------------------------------------------------------
# mesh2voxelfe

def _resolve_matprop(matprop, GVmax):
    """Range checks of voxelFE.py:579-610 (mutates ``matprop`` like the reference does)."""
    prop_GV = {}
    if matprop is None:
        logging.warning('matprop not given: material property mapping disabled.')
        return prop_GV
    if "range" not in matprop and len(matprop["file"]) > 1:
        raise IOError('Incorrect material property definition: GV ranges required when using multiple '
                      'property files.')
    elif "range" not in matprop and len(matprop["file"]) == 1:
        matprop["range"] = [[1, GVmax]]
    spans = []
    for n, GVrange in enumerate(matprop["range"]):
        lo, hi = int(GVrange[0]), int(GVrange[1])
        if lo < 0:
            raise IOError('property GV range below min representable integer')
        if hi > GVmax:
            raise IOError('property GV range exceeds max representable integer')
        prop_GV[n] = (lo, hi)
        spans.append((lo, hi))
    for i in range(len(spans) - 1):
        if max(spans[i][0], spans[i + 1][0]) <= min(spans[i][1], spans[i + 1][1]):
            raise IOError('GV ranges assigned to different material properties cannot overlap')
    return prop_GV


_TEXT_CACHE = {}


def _read_lines(path):
    """readlines() of a small text file (template, property cards), cached on (mtime, size): the deck
    writer looks at these files twice per call (size bound, then the text)."""
    st = os.stat(path)
    key = (os.path.abspath(path), st.st_mtime_ns, st.st_size)
    hit = _TEXT_CACHE.get(key)
    if hit is None:
        with open(path, 'r') as f:
            hit = tuple(f.readlines())
        if len(_TEXT_CACHE) > 64:
            _TEXT_CACHE.clear()
        _TEXT_CACHE[key] = hit
    return list(hit)


_CARD_CACHE = {}


def _file_key(fname):
    st = os.stat(fname)
    return (os.path.abspath(fname), st.st_mtime_ns, st.st_size)


def _card_for_gv(fname, fkey, GV):  # noqa: N803 -- the cards' expressions name this variable
    """The property card of one grey value (voxelFE.py:716-746).  The text is a function of the card file
    and of the NumPy scalar GV (value and dtype) only, so it is kept per (file, dtype, value)."""
    key = (fkey, type(GV), repr(GV))
    hit = _CARD_CACHE.get(key)
    if hit is not None:
        return hit
    parts = []
    setname = 'SET' + str(GV)
    scope = {"GV": GV, "np": np}
    for line in _read_lines(fname):
        line = line.replace('\n', '')
        if line.startswith('**'):
            parts.append(line + '\n')
            continue
        line = line.replace('SetName', setname).replace('MatName', 'MAT' + setname)
        if 'GV' in line:
            fields = ['{}'.format(eval(f, scope)) if 'GV' in f else f for f in line.split(',')]   # voxelFE.py:734
            parts.append(','.join(fields) + '\n')
        else:
            parts.append(line + '\n')
    hit = ''.join(parts)
    if len(_CARD_CACHE) > 4096:
        _CARD_CACHE.clear()
    _CARD_CACHE[key] = hit
    return hit


def _property_cards(matprop, prop_GV, existing_GV):
    """PROPERTY section (voxelFE.py:691-746): O(#GV) lines of host string work (eval of the card
    expressions on NumPy scalars, as the reference does), memoised per card and per whole section: a deck
    with 255 material sets costs one dictionary look-up when the same cards were printed before."""
    fkeys = [_file_key(f) for f in matprop["file"]]
    ex = np.asarray(existing_GV)
    skey = (tuple(fkeys), tuple(sorted(prop_GV.items())), ex.dtype.str, ex.tobytes())
    hit = _CARD_CACHE.get(skey)
    if hit is not None:
        return hit
    parts = ['** User material property definition:\n',
             '** internal variables are: "SetName", "MatName", "GV"\n']
    for n, fname in enumerate(matprop["file"]):
        lo, hi = prop_GV[n]
        for GV in existing_GV:  # noqa: N806
            if lo <= GV <= hi:
                parts.append(_card_for_gv(fname, fkeys[n], GV))
    hit = ''.join(parts)
    _CARD_CACHE[skey] = hit
    return hit


def _template_text(templatefile, refnode):
    """Analysis template copy with refnodeX/Y/Z substitution (voxelFE.py:748-768)."""
    lines = _read_lines(templatefile)
    if refnode is not None:
        subs = [('refnodeX', str(round(refnode[0], 4))), ('refnodeY', str(round(refnode[1], 4))),
                ('refnodeZ', str(round(refnode[2], 4)))]
        out = []
        for line in lines:
            for key, val in subs:
                line = line.replace(key, val)
            out.append(line)
        lines = out
    return ''.join(lines)


def _tail_bound(templatefile, refnode, matprop, n_gv):
    """Upper bound (bytes) of the PROPERTY section + template copy, from file sizes only."""
    t = ''.join(_read_lines(templatefile))
    bound = len(t.encode()) + 64
    if refnode is not None:
        bound += 32 * (t.count('refnodeX') + t.count('refnodeY') + t.count('refnodeZ'))
    if matprop is not None:
        bound += 256
        for fname in matprop["file"]:
            lines = _read_lines(fname)
            # per output line: substitutions at most double it; every eval'ed field prints < 32 chars
            per_gv = sum(2 * len(ln.encode()) + 32 * (ln.count(',') + 1) + 16 for ln in lines)
            bound += n_gv * per_gv
    return bound


class _Blob:
    """Host byte blob with 16-byte aligned pieces, shipped to the device in one copy."""

    def __init__(self):
        self.parts = []
        self.size = 0

    def add(self, data):
        off = self.size
        self.parts.append(data)
        padn = (-len(data)) % 16
        if padn:
            self.parts.append(b"\0" * padn)
        self.size += len(data) + padn
        return off

    def to_device(self, dev):
        return bytes_to_device_async(b"".join(self.parts), dev)


# "list": k_fill_list + k_emit_nodes4 (4 B / node of scratch, written and read once); "bits": k_emit_nodes4b
# expands the used-node bitmap inside the emitter (no list).  Measured on the 256 x 2048 x 2048 slab: 0.51 +
# 3.43 ms against 4.11 ms -- with eight 27 KB staging areas per SM the emitter only stays at the HBM rate while
# its CTAs do nothing but format, so the list is the default.
NODE_EMITTER = os.environ.get("CICLOPE_B200_NODE_EMITTER", "list")


def _emit_node_lines(m, p_xt, p_yt, p_zt, Y, X, z0, out_ptr, stream):
    if NODE_EMITTER == "bits" and out_ptr % 16 == 0:
        _lib.call("cfe_emit_nodes_bits", m.nbits.data_ptr(), m.npref.data_ptr(), m.chunk_word.data_ptr(),
                  m.n_node_words, m.n_nd, p_xt, p_yt, p_zt, Y, X, z0, out_ptr, stream)
    else:
        _lib.call("cfe_emit_nodes", m.node_list.data_ptr(), m.n_nd, p_xt, p_yt, p_zt, Y, X, z0, out_ptr, stream)


def deck_to_device(mesh, templatefile, matprop=None, keywords=['NSET', 'ELSET'], eltype='C3D8', matpropbits=8,
                   refnode=None):
    """Build the whole .INP deck in HBM.  Returns a uint8 CUDA tensor holding exactly the bytes
    the reference's mesh2voxelfe would write (voxelFE.py:612-768)."""
    if not isinstance(mesh, VoxelMesh) or mesh._host_edited():
        # foreign meshio meshes and vol2ugrid meshes edited on the host: array-driven CUDA emitters
        from .generic import deck_generic
        return deck_generic(mesh, templatefile, matprop=matprop, keywords=keywords, eltype=eltype,
                            matpropbits=matpropbits, refnode=refnode)
    m = mesh
    lib = _lib.require_cuda()
    dev = m.vol.device
    Z, Y, X = m.Z, m.Y, m.X

    # ---- prologue (voxelFE.py:540-610)
    gv_dtype = np.bool_ if m.is_bool else np.uint8
    if m._cell_data is not None:   # vetted by _host_edited: same values, dtype bool or uint8
        cell_GV = m._cell_data.get(list(m._cell_data)[0])
        if type(cell_GV) is list:
            cell_GV = cell_GV[0]
        gv_dtype = np.asarray(cell_GV).dtype.type
    if m.n_el == 0:
        raise ValueError("min() iterable argument is empty")   # voxelFE.py:564
    with torch.cuda.device(dev):
        s = _lib.stream()
        if m.gv8 is None:
            hist = np.zeros(256, dtype=np.int64)
            hist[1] = m.n_el
        else:
            hist_dev = torch.empty(256, dtype=torch.int64, device=dev)
            # over the n_el compact grey values (every one of them is an element: threshold -1 = all)
            _lib.call("cfe_gv_histogram", m.gv8.data_ptr(), m.n_el, -1, hist_dev.data_ptr(), s)
            hist = np.asarray(read_back(hist_dev), dtype=np.int64)
    gvs = np.flatnonzero(hist)
    existing_GV = gvs.astype(gv_dtype)                        # np.unique(cell_GV), voxelFE.py:570
    GVmax = 2 ** matpropbits - 1
    binary_model = bool(gvs.min() == 0) and bool(gvs.max() == 1) and issubclass(gv_dtype, np.integer)
    if binary_model:
        logging.warning('cell_data is binary. Material property mapping disabled.')
    if max(existing_GV) > GVmax:
        GVmax = max(existing_GV)
    prop_GV = {0: (1, 1)} if binary_model else _resolve_matprop(matprop, GVmax)

    # ---- sizes needed to reserve the deck buffer; the *NODE emitter is launched before any other host
    # work so that the GPU formats nodes while the host prepares the element / set sections
    head = ('** ---------------------------------------------------------\n'
            '** Abaqus .INP file written by Voxel_FEM script on {}\n'
            '** ---------------------------------------------------------\n'
            '** Node coordinates from input model\n'
            '*NODE\n').format(datetime.now()).encode()
    elem_comment = b'** Elements and Element sets from input model\n'
    # The PROPERTY cards + template (host string work, O(#GV) lines with eval) are generated AFTER
    # the emit kernels have been enqueued so that they overlap the GPU; only a bound on their size
    # is needed to reserve the deck buffer.
    tail_bound = _tail_bound(templatefile, refnode, matprop if 'PROPERTY' in keywords else None, len(gvs))
    max_node_id = (Z + 1) * (Y + 1) * (X + 1) - 1
    if max_node_id >= 10 ** 10:
        raise NotImplementedError("node ids wider than the 10-character field are not supported")
    if len('*ELEMENT, TYPE={0}, ELSET=SET255\n'.format(eltype)) > _ELEM_HDR_SLOT:
        raise ValueError("eltype too long for the element block header slot")
    max_line = len(str(m.n_el)) + 8 * len(str(max_node_id)) + 9
    n_set_ids = sum(m.set_count) + 4
    lits_bound = 1024                      # comments + 14 '*NSET, NSET=<name>' headers + separators
    wide = m.field_overflow     # some '{:12.6f}' field needs more than 12 characters: variable-length *NODE lines
    slot = 32 if wide else 16
    node_bytes = 53 * m.n_nd
    if wide:
        with torch.cuda.device(dev):
            s = _lib.stream()
            nv = m.coord_vals.numel()
            wtab = torch.empty(32 * nv, dtype=torch.uint8, device=dev)
            nstat = torch.zeros(2, dtype=torch.int64, device=dev)          # *NODE bytes, bad-value flag
            _lib.call("cfe_format_coords_wide", m.coord_vals.data_ptr(), nv, wtab.data_ptr(), nstat.data_ptr() + 8, s)
            n_ntiles = (m.n_nd + 255) // 256
            nlen8 = torch.empty(max(m.n_nd, 1), dtype=torch.uint8, device=dev)
            ntile_tot = empty_u32(n_ntiles, dev)
            ntile_off = torch.empty(max(n_ntiles, 1), dtype=torch.int64, device=dev)
            p_xt = wtab.data_ptr()
            _lib.call("cfe_node_lengths_wide", m.node_list.data_ptr(), m.n_nd, p_xt, p_xt + 32 * (X + 1),
                      p_xt + 32 * (X + 1) + 32 * (Y + 1), Y, X, 0, nlen8.data_ptr(), ntile_tot.data_ptr(), s)
            _lib.call("cfe_scan_u32", ntile_tot.data_ptr(), n_ntiles, ntile_off.data_ptr(), nstat.data_ptr(),
                      scan_workspace(max(n_ntiles, 1), dev).data_ptr(), s)
            node_bytes, bad = read_back(nstat)
        if bad:
            raise NotImplementedError("node coordinates that are not finite or not below 1e18 cannot be formatted")
    else:
        p_xt = m.ftab.data_ptr()
    p_yt = p_xt + slot * (X + 1)
    p_zt = p_yt + slot * (Y + 1)
    off_nodes = len(head)
    off_cmt = off_nodes + node_bytes
    off_elem = off_cmt + len(elem_comment)
    cap = (off_elem + m.n_el * max_line + len(gvs) * _ELEM_HDR_SLOT
           + n_set_ids * (len(str(max(max_node_id, m.n_el))) + 1) + lits_bound + tail_bound + 64)

    with torch.cuda.device(dev):
        s = _lib.stream()
        # the *NODE section starts 16-byte aligned in memory (word stores in k_emit_nodes4)
        out_full = torch.empty(cap + 16, dtype=torch.uint8, device=dev)
        out = out_full[(-(out_full.data_ptr() + off_nodes)) % 16:]
        dhead = bytes_to_device_async(head + elem_comment, dev)
        _lib.call("cfe_put_literal", dhead.data_ptr(), len(head), out.data_ptr(), None, None, s)
        if wide:
            _lib.call("cfe_emit_nodes_wide", m.node_list.data_ptr(), m.n_nd, p_xt, p_yt, p_zt, Y, X, 0,
                      nlen8.data_ptr(), ntile_off.data_ptr(), out.data_ptr() + off_nodes, s)
        else:
            _emit_node_lines(m, p_xt, p_yt, p_zt, Y, X, 0, out.data_ptr() + off_nodes, s)
        _lib.call("cfe_put_literal", dhead.data_ptr() + len(head), len(elem_comment), out.data_ptr() + off_cmt,
                  None, None, s)

    # ---- element block headers, NSET / ELSET item stream (host work overlapping the node emitter)
    hdr_text = bytearray(256 * _ELEM_HDR_SLOT)
    hdr_len = np.zeros(256, dtype=np.uint8)
    for g in gvs.tolist():
        h = '*ELEMENT, TYPE={0}, ELSET={1}\n'.format(eltype, 'SET' + str(int(g))).encode()
        hdr_text[g * _ELEM_HDR_SLOT: g * _ELEM_HDR_SLOT + len(h)] = h
        hdr_len[g] = len(h)
    first = np.zeros(256, dtype=np.int64)
    first[1:] = np.cumsum(hist)[:-1]   # emission index of the first element with grey value >= g

    blob = _Blob()
    o_hdr = blob.add(bytes(hdr_text))
    o_hlen = blob.add(hdr_len.tobytes())
    o_first = blob.add(first.tobytes())

    lits = bytearray()
    segs = []
    n_items = 0

    def literal(text):
        nonlocal n_items
        t = text.encode()
        segs.append((0, n_items, len(lits), len(t), 0, 0))
        lits.extend(t)
        n_items += 1

    def idlist(first_id, cnt):
        nonlocal n_items
        if cnt:
            segs.append((1, n_items, 0, 0, first_id, 0))
            n_items += cnt

    if 'NSET' in keywords:
        literal('** Additional nset from voxel model. New generated nsets are:\n')
        literal('** {}, \n'.format(NODE_SET_NAMES[0]))
        corner = {"NODES_X0Y0Z0": (m.set_first[4], min(2, m.set_count[4])),
                  "NODES_X0Y0Z1": (m.set_first[5], min(2, m.set_count[5]))}
        for q, name in enumerate(NODE_SET_NAMES + ["NODES_X0Y0Z0", "NODES_X0Y0Z1"]):
            literal('*NSET, NSET={0}\n'.format(name))
            f, c = (m.set_first[q], m.set_count[q]) if q < 6 else corner[name]
            idlist(f, c)
            literal('\n')
    if 'ELSET' in keywords:
        literal('** Additional elset from voxel model. New generated elsets are:\n')
        literal('** {}, \n'.format(CELL_SET_NAMES[0]))
        for q, name in enumerate(CELL_SET_NAMES):
            literal('*ELSET, ELSET={}\n'.format(name))
            idlist(m.set_first[6 + q], m.set_count[6 + q])
            literal('\n')
    if len(lits) > lits_bound:
        raise _lib.CfeError("internal error: set headers exceed their reserved bound")
    seg_arr = (CfeItemSeg * max(len(segs), 1))()
    for i, (kind, item0, lo, ll, id0, rank0) in enumerate(segs):
        sg = seg_arr[i]
        sg.kind, sg.item0, sg.lit_off, sg.lit_len, sg.id0, sg.rank0 = kind, item0, lo, ll, id0, rank0
    o_lits = blob.add(bytes(lits))
    o_segs = blob.add(bytes(seg_arr))

    with torch.cuda.device(dev):
        s = _lib.stream()
        dblob = blob.to_device(dev)
        base = dblob.data_ptr()
        totals = torch.zeros(2, dtype=torch.int64, device=dev)   # element bytes, set bytes
        ws = scan_workspace(max(n_items, 1), dev)
        ws_el = m.ws_el
        perm = None
        single_gv = int(gvs[0])
        if len(gvs) > 1:
            # stable partition by grey value; the raster-order line lengths of vol2ugrid travel along
            perm = torch.empty((m.n_el, 2), dtype=torch.int32, device=dev)
            ws_el = torch.empty(m.ws_el.numel(), dtype=torch.uint8, device=dev)
            ws_p = torch.empty(lib.cfe_gv_partition_workspace_bytes(m.n_el), dtype=torch.uint8, device=dev)
            _lib.call("cfe_gv_partition", m.vol.data_ptr(), m.elem_vox.data_ptr(), m.n_el, m.gv8.data_ptr(),
                      m.ws_el.data_ptr(), ws_el.data_ptr(), perm.data_ptr(), ws_p.data_ptr(), s)
            single_gv = -1
        el_args = (m.elem_vox.data_ptr(), _lib.ptr(perm), m.n_el, m.vol.data_ptr(), base + o_first, base + o_hdr,
                   base + o_hlen, single_gv, Z, Y, X, 0, 0)
        _lib.call("cfe_element_lengths", *el_args, 1, totals.data_ptr(), ws_el.data_ptr(), s)
        _lib.call("cfe_emit_elements", *el_args, out.data_ptr() + off_elem, None, ws_el.data_ptr(), s)
        _lib.call("cfe_emit_items", base + o_segs, len(segs), n_items, m.set_ids.data_ptr(), base + o_lits,
                  out.data_ptr() + off_elem, totals.data_ptr(), totals.data_ptr() + 8, None, ws.data_ptr(), s)
        # host text while the GPU formats: PROPERTY cards (voxelFE.py:691-746) + template (:748-768)
        tail = ''
        if 'PROPERTY' in keywords:
            tail += _property_cards(matprop, prop_GV, existing_GV)
        tail += _template_text(templatefile, refnode)
        tail = tail.encode()
        if len(tail) > tail_bound:
            raise _lib.CfeError("internal error: PROPERTY/template text (%d bytes) exceeds its reserved bound (%d)"
                                % (len(tail), tail_bound))
        dtail = bytes_to_device_async(tail, dev)
        _lib.call("cfe_put_literal", dtail.data_ptr(), len(tail), out.data_ptr() + off_elem, totals.data_ptr(),
                  totals.data_ptr() + 8, s)
        t = read_back(totals)
    length = off_elem + t[0] + t[1] + len(tail)
    if length > cap:
        raise _lib.CfeError("internal error: deck length %d exceeds the reserved capacity %d" % (length, cap))
    mesh._last_keepalive = (dblob, dtail, dhead)
    return out[:length]


_PINNED_CHUNKS = {}
STREAM_CHUNK = 64 << 20


RING = 3        # staging buffers: two copies are always queued behind the one being consumed


def _pinned_ring(chunk):
    bufs = _PINNED_CHUNKS.get(chunk)
    if bufs is None:
        bufs = _PINNED_CHUNKS[chunk] = [torch.empty(chunk, dtype=torch.uint8).pin_memory() for _ in range(RING)]
    return bufs


def stream_device_bytes(pieces, sink, chunk=None):
    """Device bytes -> host consumer through a ring of pinned staging buffers: the copies of chunks k + 1 and
    k + 2 are queued on the copy engine while ``sink(memoryview, offset)`` consumes chunk k, so the engine
    never waits for the host to wake up between chunks (no pageable copies, no intermediate bytes objects,
    host memory bounded by RING * chunk whatever the deck size).  ``pieces`` = [(offset, uint8 CUDA tensor)];
    returns the number of bytes handed to the sink."""
    chunk = chunk or STREAM_CHUNK
    pieces = [(off, t) for off, t in pieces if t.numel()]
    if not pieces:
        return 0
    dev = pieces[0][1].device
    bufs = _pinned_ring(chunk)
    spans = [(off + lo, t, lo, min(lo + chunk, t.numel())) for off, t in pieces for lo in range(0, t.numel(), chunk)]
    total = 0
    with torch.cuda.device(dev):
        copy_stream = torch.cuda.Stream(device=dev)
        copy_stream.wait_stream(torch.cuda.current_stream(dev))
        events = [None] * RING

        def start(k):
            _, t, lo, hi = spans[k]
            with torch.cuda.stream(copy_stream):
                bufs[k % RING][:hi - lo].copy_(t[lo:hi], non_blocking=True)
                events[k % RING] = torch.cuda.Event()
                events[k % RING].record(copy_stream)

        for k in range(min(RING - 1, len(spans))):
            start(k)
        for k, (off, _, lo, hi) in enumerate(spans):
            events[k % RING].synchronize()
            if k + RING - 1 < len(spans):
                start(k + RING - 1)          # its buffer held chunk k - 1, consumed in the previous iteration
            sink(memoryview(bufs[k % RING].numpy())[:hi - lo], off)
            total += hi - lo
        torch.cuda.current_stream(dev).wait_stream(copy_stream)
    return total


def _write_device_bytes(deck, path, chunk=None):
    """Device bytes -> file (sequential writes of the pinned chunks)."""
    with open(path, 'wb') as INP:
        stream_device_bytes([(0, deck)], lambda mv, off: INP.write(mv), chunk)


def mesh2voxelfe(mesh, templatefile, fileout, matprop=None, keywords=['NSET', 'ELSET'], eltype='C3D8',
                 matpropbits=8, refnode=None, verbose=False):
    """Generate ABAQUS voxel Finite Element (FE) input file from 3D Unstructured Grid mesh data
    (drop-in for voxelFE.py:474-772).  Parameters are the reference's.

    ``fileout`` is the path of the .INP file.  Additively it may be a pinned/CPU uint8
    ``torch.Tensor`` that receives the deck bytes, or an object with a ``write(buffer)`` method (an open
    binary file, a socket file, a hasher ...) that is fed the deck in order through a three-buffer pinned
    ring -- decks larger than the host's memory stream through it.  Both return the number of bytes.
    """
    if verbose:
        logging.basicConfig(level=logging.INFO)
    logging.info('Start writing INP file')
    deck = deck_to_device(mesh, templatefile, matprop=matprop, keywords=keywords, eltype=eltype,
                          matpropbits=matpropbits, refnode=refnode)
    n = deck.numel()
    if isinstance(fileout, torch.Tensor):
        if fileout.dtype != torch.uint8 or fileout.numel() < n:
            raise ValueError("fileout tensor must be uint8 with at least %d elements" % n)
        fileout[:n].copy_(deck, non_blocking=True)
        torch.cuda.current_stream(deck.device).synchronize()
        return n
    if hasattr(fileout, "write"):
        write = fileout.write
        return stream_device_bytes([(0, deck)], lambda mv, off: write(mv))
    _write_device_bytes(deck, fileout)
    if isinstance(mesh, VoxelMesh):
        n_points, n_cells = mesh.n_points, mesh.n_el
    else:
        n_points, n_cells = len(mesh.points), len(mesh.cells[0][1])
    logging.info('Model with {0} nodes and {1} elements written to file {fname}'.format(
        n_points, n_cells, fname=fileout))
    return


def matpropdictionary(proplist):
    """Compose dictionary of material properties and property mapping GV ranges
    (drop-in for voxelFE.py:774-819)."""
    if proplist is None:
        return None
    if len(proplist) == 0:
        return None
    elif len(proplist) == 1:
        if not os.path.isfile(proplist[0]):
            raise IOError('Property file {0} does not exist.'.format(proplist[0]))
        return {"file": [proplist[0]]}
    if not len(proplist) % 3 == 0:
        raise IOError('Mapping dictionary error (probably due to incorrect use of --mapping flag): each material '
                      'property file must be followed by MIN and MAX of the corresponding GV range.')
    matprop = {'file': [], 'range': []}
    for i in range(len(proplist) // 3):
        matprop['file'].append(proplist[3 * i])
        matprop['range'].append([int(proplist[3 * i + 1]), int(proplist[3 * i + 2])])
    return matprop


# vol2h5ParOSol (voxelFE.py:285-472) lives in parosol.py; the reference exposes it from this module
from .parosol import parosol_datasets, vol2h5ParOSol  # noqa: E402,F401
