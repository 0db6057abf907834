"""TEST INFRASTRUCTURE ONLY -- Python front-end of the CPU parity oracle.

Restates the reference's hot path (ciclope 2.0.3) on the CPU:

* ``remove_unconnected``  <- src/ciclope/utils/preprocess.py:119-143
* ``vol2ugrid``           <- src/ciclope/core/voxelFE.py:15-283
* ``mesh2voxelfe``        <- src/ciclope/core/voxelFE.py:474-772
* ``matpropdictionary``   <- src/ciclope/core/voxelFE.py:774-819

The voxel / node / text loops are in ``ciclope_oracle.c`` (sequential C, built by
``oracle/build_oracle.py``); the header, PROPERTY cards and template copy are Python string
work here.  Parity status: PINNED against golden vectors produced by the unmodified
reference (tests/golden/, tests/test_oracle_golden.py).

Nothing under ``ciclope_b200/`` may import this module: it is the checker, not the product.
"""
import collections
import ctypes
import math
import os
from datetime import datetime

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None

_i64 = ctypes.c_int64
_p_u8 = ctypes.POINTER(ctypes.c_uint8)
_p_i64 = ctypes.POINTER(ctypes.c_int64)
_p_f64 = ctypes.POINTER(ctypes.c_double)


class _Sets6(ctypes.Structure):
    _fields_ = [("ids", _p_i64 * 6), ("n", _i64 * 6)]


class _Section(ctypes.Structure):
    _fields_ = [("bytes", _i64), ("digest", ctypes.c_uint8 * 32)]


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            from . import build_oracle
            build_oracle.build()
        L = ctypes.CDLL(_LIB_PATH)
        L.orc_remove_unconnected.restype = _i64
        L.orc_remove_unconnected.argtypes = [_p_u8, _i64, _i64, _i64, _p_u8]
        L.orc_count_components.restype = _i64
        L.orc_count_components.argtypes = [_p_u8, _i64, _i64, _i64]
        L.orc_count_cells.restype = _i64
        L.orc_count_cells.argtypes = [_p_u8, _i64, _i64, _i64, ctypes.c_int]
        L.orc_fill_cells.restype = None
        L.orc_fill_cells.argtypes = [_p_u8, _i64, _i64, _i64, ctypes.c_int, _p_i64, _p_u8]
        L.orc_boundary_sets.restype = ctypes.c_int
        L.orc_boundary_sets.argtypes = [_p_u8, _i64, _i64, _i64, ctypes.c_int, ctypes.c_int, _i64, _i64,
                                        ctypes.POINTER(_Sets6), ctypes.POINTER(_Sets6)]
        L.orc_free_sets.restype = None
        L.orc_free_sets.argtypes = [ctypes.POINTER(_Sets6)]
        L.orc_refnode.restype = None
        L.orc_refnode.argtypes = [_p_i64, _i64, _i64, _i64, _p_f64, _p_f64, _p_f64, _p_f64]
        L.orc_drop_largest.restype = _i64
        L.orc_drop_largest.argtypes = [_p_u8, _i64, _i64, _i64, _p_u8]
        L.orc_deck_body.restype = ctypes.c_int
        L.orc_deck_body.argtypes = [_p_f64, _i64, _p_i64, _i64, _p_i64, _i64,
                                    ctypes.POINTER(ctypes.c_int32), ctypes.c_int, _p_i64, ctypes.c_char_p,
                                    ctypes.c_int, ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(_p_i64), _p_i64,
                                    ctypes.c_int, ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(_p_i64), _p_i64,
                                    ctypes.POINTER(ctypes.c_void_p), _p_i64]
        L.orc_free.restype = None
        L.orc_free.argtypes = [ctypes.c_void_p]
        L.orc_stream_deck.restype = ctypes.c_int
        L.orc_stream_deck.argtypes = [_p_u8, _i64, _i64, _i64, ctypes.c_int, ctypes.c_char_p, ctypes.c_char_p,
                                      ctypes.c_char_p, ctypes.c_int, ctypes.c_char_p, ctypes.c_int, _i64, _i64,
                                      ctypes.c_int, ctypes.c_int, ctypes.c_char_p, ctypes.POINTER(_Section * 4),
                                      _p_i64, _p_i64]
        L.orc_sha256.restype = None
        L.orc_sha256.argtypes = [ctypes.c_char_p, _i64, ctypes.c_char_p]
        _lib = L
    return _lib


def _u8(a):
    a = np.ascontiguousarray(a)
    if a.dtype == np.bool_:
        a = a.view(np.uint8)
    if a.dtype != np.uint8:
        raise TypeError("oracle handles bool / uint8 volumes, got %s" % a.dtype)
    return a


def _ptr(a, t):
    return a.ctypes.data_as(t)


# ----------------------------------------------------------------------------- CCL

def remove_unconnected(bwimage):
    """preprocess.py:119-143 -- largest 6-connected component, ties -> first in raster order."""
    bw = np.asarray(bwimage)
    if bw.ndim != 3:
        raise ValueError("oracle remove_unconnected expects a 3-D volume")
    u = _u8(bw)
    out = np.empty(u.shape, dtype=np.uint8)
    Z, Y, X = u.shape
    kept = lib().orc_remove_unconnected(_ptr(u, _p_u8), Z, Y, X, _ptr(out, _p_u8))
    if kept < 0:
        # occurrences[1:].argmax() on an empty array (preprocess.py:141)
        raise ValueError("attempt to get argmax of an empty sequence")
    return out.view(np.bool_)


def _drop_largest(u):
    out = np.empty(u.shape, dtype=np.uint8)
    Z, Y, X = u.shape
    if lib().orc_drop_largest(_ptr(u, _p_u8), Z, Y, X, _ptr(out, _p_u8)) < 0:
        raise ValueError("attempt to get argmax of an empty sequence")
    return out.view(np.bool_)


def remove_largest(bwimage):
    """preprocess.py:187-208 -- everything but the largest 6-connected component."""
    return _drop_largest(_u8(np.asarray(bwimage)))


def fill_voids(I, fill_val=None, makecopy=False):  # noqa: E741 -- the reference's parameter name
    """preprocess.py:145-185 -- background regions other than the largest one are set to fill_val."""
    if fill_val is None:
        fill_val = I.max()
    voids = _drop_largest(np.ascontiguousarray(~(I > 0)).view(np.uint8))
    if makecopy:
        I_filled = I.copy()
        I_filled[voids] = fill_val
        return I_filled
    I[voids] = fill_val
    return I


def count_components(bwimage):
    u = _u8(np.asarray(bwimage))
    Z, Y, X = u.shape
    return int(lib().orc_count_components(_ptr(u, _p_u8), Z, Y, X))


# ----------------------------------------------------------------------------- vol2ugrid

CellBlock = collections.namedtuple("CellBlock", ["type", "data"])

NODE_SET_NAMES = ["NODES_X0", "NODES_X1", "NODES_Y0", "NODES_Y1", "NODES_Z0", "NODES_Z1"]
CELL_SET_NAMES = ["CELLS_X0", "CELLS_X1", "CELLS_Y0", "CELLS_Y1", "CELLS_Z0", "CELLS_Z1"]


class OracleMesh:
    """The meshio-5.0.0 surface the reference and its tests touch (voxelFE.py:271-274)."""

    def __init__(self, points, cells, cell_data, point_sets, cell_sets):
        self.points = points
        self.cells = [CellBlock("hexahedron", cells)]
        self.cell_data = cell_data
        self.point_sets = point_sets
        self.cell_sets = cell_sets

    @property
    def cells_dict(self):
        return {"hexahedron": self.cells[0].data}


def gv_threshold(GVmin):
    """Integer t with (GV > GVmin) <=> (GV > t) for GV in 0..255 (voxelFE.py:141)."""
    import math
    if GVmin < 0:
        return -1
    return int(min(math.floor(GVmin), 255))


def coordinate_tables(voxelsize, shape):
    """Per-axis coordinate values exactly as voxelFE.py:70-73, 202 computes them, converted to
    the dtype np.asarray(nodes) would pick (meshio.Mesh -> np.asarray(points))."""
    Z, Y, X = shape
    if not isinstance(voxelsize, (list, np.ndarray)):
        voxelsize = [voxelsize, voxelsize, voxelsize]
    if len(voxelsize) == 1:
        voxelsize = [voxelsize[0]] * 3
    xs = [voxelsize[0] * c for c in range(X + 1)]
    ys = [voxelsize[1] * r for r in range(Y + 1)]
    zs = [voxelsize[2] * s for s in range(Z + 1)]
    dtype = np.asarray([[xs[0], ys[0], zs[0]], [xs[-1], ys[-1], zs[-1]]]).dtype
    return (np.asarray(xs).astype(dtype), np.asarray(ys).astype(dtype), np.asarray(zs).astype(dtype))


def vol2ugrid(voldata, voxelsize=[1, 1, 1], GVmin=0, plane_lock_num=1, node_level_lock_num=1,
              locking_strategy="plane", refnodes=False, verbose=False):
    voldata = np.asarray(voldata)
    if locking_strategy not in ("plane", "exact"):
        raise ValueError("locking_strategy must be 'plane' or 'exact'")
    u = _u8(voldata)
    Z, Y, X = u.shape
    thr = gv_threshold(GVmin)
    L = lib()
    n_el = int(L.orc_count_cells(_ptr(u, _p_u8), Z, Y, X, thr))
    cells = np.empty((n_el, 8), dtype=np.int64)
    gv = np.empty(n_el, dtype=np.uint8)
    L.orc_fill_cells(_ptr(u, _p_u8), Z, Y, X, thr, _ptr(cells, _p_i64), _ptr(gv, _p_u8))
    if voldata.dtype == np.bool_:
        gv = gv.view(np.bool_)

    xs, ys, zs = coordinate_tables(voxelsize, (Z, Y, X))
    points = np.empty((Z + 1, Y + 1, X + 1, 3), dtype=xs.dtype)
    points[..., 0] = xs[None, None, :]
    points[..., 1] = ys[None, :, None]
    points[..., 2] = zs[:, None, None]
    points = points.reshape(-1, 3)

    ns, cs = _Sets6(), _Sets6()
    L.orc_boundary_sets(_ptr(u, _p_u8), Z, Y, X, thr, 0 if locking_strategy == "plane" else 1,
                        int(plane_lock_num), int(node_level_lock_num), ctypes.byref(ns), ctypes.byref(cs))
    point_sets, cell_sets = {}, {}
    node_arrays = []
    for q in range(6):
        a = np.ctypeslib.as_array(ns.ids[q], shape=(ns.n[q],)).copy() if ns.n[q] else np.empty(0, np.int64)
        node_arrays.append(a)
        point_sets[NODE_SET_NAMES[q]] = a.tolist()
        b = np.ctypeslib.as_array(cs.ids[q], shape=(cs.n[q],)).copy() if cs.n[q] else np.empty(0, np.int64)
        cell_sets[CELL_SET_NAMES[q]] = b.tolist()
    L.orc_free_sets(ctypes.byref(ns))
    L.orc_free_sets(ctypes.byref(cs))
    point_sets["NODES_X0Y0Z0"] = point_sets["NODES_Z0"][0:2]
    point_sets["NODES_X0Y0Z1"] = point_sets["NODES_Z1"][0:2]

    mesh = OracleMesh(points, cells, {"GV": [gv]}, point_sets, cell_sets)
    if not refnodes:
        return mesh
    # voxelFE.py:202-223, 254-268: sequential float64 sums of the Python coordinate values
    x64 = np.asarray(xs, dtype=np.float64)
    y64 = np.asarray(ys, dtype=np.float64)
    z64 = np.asarray(zs, dtype=np.float64)
    rn = {}
    for q, key in enumerate(["X0", "X1", "Y0", "Y1", "Z0", "Z1"]):
        out = np.zeros(3)
        a = node_arrays[q]
        L.orc_refnode(_ptr(a, _p_i64), len(a), Y, X, _ptr(x64, _p_f64), _ptr(y64, _p_f64), _ptr(z64, _p_f64),
                      _ptr(out, _p_f64))
        rn[key] = out
    return mesh, rn


# ----------------------------------------------------------------------------- mesh2voxelfe

def _resolve_matprop(matprop, existing_GV, GVmax):
    """voxelFE.py:579-610 -- returns {file index: (lo, hi)}; mutates matprop like the reference."""
    prop_GV = {}
    if matprop is None:
        return prop_GV
    if "range" not in matprop and len(matprop["file"]) > 1:
        raise IOError('Incorrect material property definition: GV ranges required when using multiple property files.')
    elif "range" not in matprop and len(matprop["file"]) == 1:
        matprop["range"] = [[1, GVmax]]
    ranges = []
    for n, GVrange in enumerate(matprop["range"]):
        lo, hi = int(GVrange[0]), int(GVrange[1])
        if lo < 0:
            raise IOError('property GV range below min representable integer')
        if hi > GVmax:
            raise IOError('property GV range exceeds max representable integer')
        prop_GV[n] = (lo, hi)
        ranges.append(set(range(lo, hi + 1)))
    for i in range(len(ranges) - 1):
        if not ranges[i].isdisjoint(ranges[i + 1]):
            raise IOError('GV ranges assigned to different material properties cannot overlap')
    return prop_GV


def _property_section(matprop, prop_GV, existing_GV):
    """voxelFE.py:691-746."""
    out = ['** User material property definition:\n',
           '** internal variables are: "SetName", "MatName", "GV"\n']
    for n, fname in enumerate(matprop["file"]):
        with open(fname, 'r') as f:
            lines = f.readlines()
        for GV in existing_GV:  # noqa: N806 -- 'GV' is the name the cards' expressions use
            if not (prop_GV[n][0] <= GV <= prop_GV[n][1]):
                continue
            setname = 'SET' + str(GV)
            for line in lines:
                line = line.replace('\n', '')
                if line.startswith('**'):
                    out.append(line + '\n')
                    continue
                if 'SetName' in line:
                    line = line.replace('SetName', setname)
                if 'MatName' in line:
                    line = line.replace('MatName', 'MAT' + setname)
                if 'GV' in line:
                    pieces = line.split(',')
                    env = {"GV": GV, "np": np}
                    out.append(','.join('{}'.format(eval(p, env)) if 'GV' in p else p for p in pieces) + '\n')
                else:
                    out.append(line + '\n')
    return ''.join(out)


def mesh2voxelfe(mesh, templatefile, fileout, matprop=None, keywords=['NSET', 'ELSET'], eltype='C3D8',
                 matpropbits=8, refnode=None, verbose=False):
    keys = list(mesh.cell_data)
    cell_GV = mesh.cell_data.get(keys[0])
    if type(cell_GV) is list:
        cell_GV = cell_GV[0]
    cell_GV = np.asarray(cell_GV)
    if cell_GV.size == 0:
        raise ValueError("min() iterable argument is empty")  # voxelFE.py:564
    GVmax = 2 ** matpropbits - 1
    binary_model = bool(cell_GV.min() == 0) and bool(cell_GV.max() == 1) and \
        issubclass(cell_GV.dtype.type, np.integer)
    existing_GV = np.unique(cell_GV)
    if max(existing_GV) > GVmax:
        GVmax = max(existing_GV)
    if binary_model:
        prop_GV = {0: (1, 1)}
    else:
        prop_GV = _resolve_matprop(matprop, existing_GV, GVmax)

    L = lib()
    points = np.ascontiguousarray(mesh.points, dtype=np.float64)
    cells = np.ascontiguousarray(mesh.cells[0][1], dtype=np.int64)                   # voxelFE.py:647
    allhex = np.ascontiguousarray(mesh.cells_dict["hexahedron"], dtype=np.int64)     # voxelFE.py:628
    _, inv = np.unique(cell_GV, return_inverse=True)
    group = np.ascontiguousarray(inv.reshape(-1), dtype=np.int32)
    labels = np.asarray([int(GV) for GV in existing_GV], dtype=np.int64)             # 'SET' + str(int(GV))

    def pack(sets):
        names = list(sets)
        arrs = [np.asarray(sets[k], dtype=np.int64) for k in names]
        c_names = (ctypes.c_char_p * max(len(names), 1))(*[k.encode() for k in names])
        c_ids = (_p_i64 * max(len(names), 1))(*[_ptr(a, _p_i64) for a in arrs])
        c_n = np.asarray([len(a) for a in arrs] + [0], dtype=np.int64)
        return len(names), c_names, c_ids, c_n, arrs

    nn, nnames, nids, ncnt, _keep1 = pack(mesh.point_sets)
    en, enames, eids, ecnt, _keep2 = pack(mesh.cell_sets)
    if 'NSET' not in keywords:
        nn = -1
    if 'ELSET' not in keywords:
        en = -1
    out_p = ctypes.c_void_p()
    out_n = _i64()
    if 'NSET' in keywords:
        '** {}, \n'.format(*list(mesh.point_sets))      # voxelFE.py:652 raises IndexError on an empty dict
    if 'ELSET' in keywords:
        '** {}, \n'.format(*list(mesh.cell_sets))
    L.orc_deck_body(_ptr(points, _p_f64), len(points), _ptr(cells, _p_i64), len(cells),
                    _ptr(allhex, _p_i64), len(allhex), _ptr(group, ctypes.POINTER(ctypes.c_int32)), len(labels),
                    _ptr(labels, _p_i64), eltype.encode(), nn, nnames, nids, _ptr(ncnt, _p_i64), en, enames, eids, _ptr(ecnt, _p_i64),
                    ctypes.byref(out_p), ctypes.byref(out_n))
    body = ctypes.string_at(out_p.value, out_n.value)
    L.orc_free(out_p)

    with open(fileout, 'wb') as INP:
        INP.write(b'** ---------------------------------------------------------\n')
        INP.write('** Abaqus .INP file written by Voxel_FEM script on {}\n'.format(datetime.now()).encode())
        INP.write(b'** ---------------------------------------------------------\n')
        INP.write(body)
        if 'PROPERTY' in keywords:
            INP.write(_property_section(matprop, prop_GV, existing_GV).encode())
        with open(templatefile, 'r') as template:
            for line in template.readlines():
                if refnode is not None:
                    line = line.replace('refnodeX', str(round(refnode[0], 4)))
                    line = line.replace('refnodeY', str(round(refnode[1], 4)))
                    line = line.replace('refnodeZ', str(round(refnode[2], 4)))
                INP.write(line.encode())


def stream_deck(voldata, templatefile, voxelsize=[1, 1, 1], GVmin=0, plane_lock_num=1, node_level_lock_num=1,
                locking_strategy="plane", matprop=None, keywords=['NSET', 'ELSET'], eltype='C3D8', matpropbits=8,
                refnode=None, fileout=None):
    """mesh2voxelfe(vol2ugrid(voldata, ...), templatefile, ...) WITHOUT materialising the mesh: the deck is
    streamed section by section from the volume into SHA-256 sinks (oracle/stream_oracle.c), which is how the
    oracle reaches the benchmark sizes (512^3: ~0.5 GB of host memory instead of ~60 GB).  Returns an ordered
    dict section -> {"bytes", "sha256"} with the sections of ciclope_b200.sections.split_deck ("header" is
    left out: it holds the time stamp; "tail" is the PROPERTY cards + template text) plus "n_el" / "n_nd".
    With `fileout` the same bytes are also written to that file (small cases: compared byte for byte with
    the golden decks of the reference, which pins the streamed restatement)."""
    import collections
    import hashlib
    voldata = np.asarray(voldata)
    if locking_strategy not in ("plane", "exact"):
        raise ValueError("locking_strategy must be 'plane' or 'exact'")
    u = _u8(voldata)
    Z, Y, X = u.shape
    thr = gv_threshold(GVmin)
    xs, ys, zs = coordinate_tables(voxelsize, (Z, Y, X))
    # '{n[0]:12.6f}'.format(n=point): the field of every distinct coordinate value, by the interpreter
    fields = [['{0:12.6f}'.format(v).encode() for v in a] for a in (xs, ys, zs)]
    stride = max(len(f) for fl in fields for f in fl) + 1
    tabs = [b"".join(f.ljust(stride, b"\0") for f in fl) for fl in fields]
    # prologue checks of mesh2voxelfe on the grey values present (voxelFE.py:564-610)
    present = np.flatnonzero(np.bincount(u.reshape(-1), minlength=256) * (np.arange(256) > thr))
    if present.size == 0:
        raise ValueError("min() iterable argument is empty")
    gv_dtype = np.bool_ if voldata.dtype == np.bool_ else np.uint8
    existing_GV = present.astype(gv_dtype)
    GVmax = 2 ** matpropbits - 1
    binary_model = bool(present.min() == 0) and bool(present.max() == 1) and issubclass(gv_dtype, np.integer)
    if max(existing_GV) > GVmax:
        GVmax = max(existing_GV)
    prop_GV = {0: (1, 1)} if binary_model else _resolve_matprop(matprop, existing_GV, GVmax)
    if fileout is not None:
        with open(fileout, 'wb') as INP:
            INP.write(b'** ---------------------------------------------------------\n')
            INP.write('** Abaqus .INP file written by Voxel_FEM script on {}\n'.format(datetime.now()).encode())
            INP.write(b'** ---------------------------------------------------------\n')
    secs = (_Section * 4)()
    n_el, n_nd = _i64(), _i64()
    rc = lib().orc_stream_deck(_ptr(u, _p_u8), Z, Y, X, thr, tabs[0], tabs[1], tabs[2], stride, eltype.encode(),
                               0 if locking_strategy == "plane" else 1, int(plane_lock_num),
                               int(node_level_lock_num), int('NSET' in keywords), int('ELSET' in keywords),
                               None if fileout is None else os.fsencode(fileout), ctypes.byref(secs),
                               ctypes.byref(n_el), ctypes.byref(n_nd))
    if rc:
        raise RuntimeError("orc_stream_deck failed (%d)" % rc)
    tail = _property_section(matprop, prop_GV, existing_GV) if 'PROPERTY' in keywords else ''
    with open(templatefile, 'r') as template:
        for line in template.readlines():
            if refnode is not None:
                line = line.replace('refnodeX', str(round(refnode[0], 4)))
                line = line.replace('refnodeY', str(round(refnode[1], 4)))
                line = line.replace('refnodeZ', str(round(refnode[2], 4)))
            tail += line
    tail = tail.encode()
    if fileout is not None:
        with open(fileout, 'ab') as INP:
            INP.write(tail)
    out = collections.OrderedDict()
    for name, sec, present_kw in (("nodes", secs[0], True), ("elements", secs[1], True),
                                  ("nset", secs[2], 'NSET' in keywords), ("elset", secs[3], 'ELSET' in keywords)):
        if present_kw:
            out[name] = {"bytes": int(sec.bytes), "sha256": bytes(sec.digest).hex()}
    out["tail"] = {"bytes": len(tail), "sha256": hashlib.sha256(tail).hexdigest()}
    out["n_el"], out["n_nd"] = int(n_el.value), int(n_nd.value)
    return out


def matpropdictionary(proplist):
    """voxelFE.py:774-819."""
    if proplist is None or len(proplist) == 0:
        return None
    if len(proplist) == 1:
        if not os.path.isfile(proplist[0]):
            raise IOError('Property file {0} doesnt exist.', format(proplist[0]))
        return {"file": [proplist[0]]}
    if len(proplist) % 3 != 0:
        raise IOError('Mapping dictionary error (probably due to incorrect use of --mapping flag): each material '
                      'property file must be followed by MIN and MAX of the corresponding GV range.')
    matprop = {'file': [], 'range': []}
    for i in range(len(proplist) // 3):
        matprop['file'].append(proplist[3 * i])
        matprop['range'].append([int(proplist[3 * i + 1]), int(proplist[3 * i + 2])])
    return matprop


# --------------------------------------------------------------------------------------------
# vol2h5ParOSol <- src/ciclope/core/voxelFE.py:285-472 (+ recon_utils.bbox, recon_utils.py:232-318, and
# preprocess._collect_from_elements, preprocess.py:422-456).  The HDF5 container is h5py's; this restates
# the five datasets the reference hands to create_dataset, in the reference's order.

def bbox(bw):
    """recon_utils.py:258-318 with pad = 0: ([row0, col0, slice0], [rowd, cold, sliced])."""
    maxROW = np.max(np.max(bw, 0), 1)
    maxCOL = np.max(np.max(bw, 0), 0)
    maxSLICE = np.max(np.max(bw, 1), 1)
    row0 = list(maxROW).index(True)
    row1 = len(maxROW) - list(maxROW[::-1]).index(True) - 1
    col0 = list(maxCOL).index(True)
    col1 = len(maxCOL) - list(maxCOL[::-1]).index(True) - 1
    slice0 = list(maxSLICE).index(True)
    slice1 = len(maxSLICE) - list(maxSLICE[::-1]).index(True) - 1
    return [row0, col0, slice0], [row1 - row0, col1 - col0, slice1 - slice0]


def vol2h5ParOSol_datasets(voldata, topDisplacement, voxelsize=1, poisson_ratio=0.3, young_modulus=18e3,
                           topHorizontaFixedlDisplacement=True, locking_strategy="plane", plane_lock_num=1,
                           node_level_lock_num=1):
    out = collections.OrderedDict()
    origin, dims = bbox(voldata)                                       # voxelFE.py:336
    origin = [origin[2], origin[0], origin[1]]
    dims = [dims[2], dims[0], dims[1]]
    end = [origin[i] + dims[i] for i in range(3)]
    sub_block = voldata[origin[0]:end[0] + 1, origin[1]:end[1] + 1, origin[2]:end[2] + 1]
    out["Image"] = np.asarray((sub_block * young_modulus).astype(np.float32), dtype='<f4')      # :354-357
    out["Voxelsize"] = np.asarray(voxelsize, dtype='<f8')
    out["Poisons_ratio"] = np.asarray(poisson_ratio, dtype='<f8')
    lb, tb = set(), set()

    def collect(els, nodes, dirs):                                     # preprocess.py:449-456
        for el in els:
            z, y, x = el
            for dz in (0, 1):
                for dy in (0, 1):
                    for dx in (0, 1):
                        for d in dirs:
                            nodes.add((z + dz, y + dy, x + dx, d))

    if locking_strategy == "plane":                                    # :371-398
        low, top = [], []
        for i in range(plane_lock_num):
            low2d = np.argwhere(sub_block[i] != 0)
            low.append(np.hstack((np.full((low2d.shape[0], 1), i), low2d)))
            idx = sub_block.shape[0] - 1 - i
            top2d = np.argwhere(sub_block[idx] != 0)
            top.append(np.hstack((np.full((top2d.shape[0], 1), idx), top2d)))
        low, top = np.vstack(low), np.vstack(top)
        collect(low, lb, (0, 1, 2))
        collect(top, tb, (0, 1, 2) if topHorizontaFixedlDisplacement else (2,))
    elif locking_strategy == "exact":                                  # :400-443
        Nz, Ny, Nx = sub_block.shape
        slice_nodes, row_nodes = (Ny + 1) * (Nx + 1), Nx + 1
        existing = np.zeros((Nz + 1) * slice_nodes, dtype=bool)
        for z, y, x in np.argwhere(sub_block != 0):
            base = z * slice_nodes + y * row_nodes + x
            for off in (0, 1, row_nodes, row_nodes + 1, slice_nodes, slice_nodes + 1,
                        slice_nodes + row_nodes, slice_nodes + row_nodes + 1):
                existing[base + off] = True
        for idx in np.flatnonzero(existing).tolist():
            z = idx // slice_nodes
            y = (idx % slice_nodes) // row_nodes
            x = (idx % slice_nodes) % row_nodes
            if z < node_level_lock_num:
                for d in (0, 1, 2):
                    lb.add((z, y, x, d))
            if z >= (Nz + 1) - node_level_lock_num:
                if topHorizontaFixedlDisplacement:
                    for d in (0, 1, 2):
                        tb.add((z, y, x, d))
                else:
                    tb.add((z, y, x, 2))
    else:
        raise ValueError("locking_strategy must be 'plane' or 'exact'")
    all_nodes = list(lb) + list(tb)                                    # :449 -- CPython set order
    out["Fixed_Displacement_Coordinates"] = np.array(all_nodes, dtype=np.uint16)
    values = np.zeros(len(all_nodes), dtype=np.float32)
    for i, (_, _, _, d) in enumerate(all_nodes[len(lb):], start=len(lb)):
        if d == 2:
            values[i] = topDisplacement
    out["Fixed_Displacement_Values"] = values
    return out


# --------------------------------------------------------------------------------------------
# segment <- src/ciclope/utils/preprocess.py:20-46;  add_cap <- preprocess.py:210-231

def threshold_otsu_u8(image):
    """skimage.filters.threshold_otsu for an integer image (scikit-image is not in the checkout nor
    installed: PARITY UNPINNED for this function).  Restated from scikit-image 0.19's published algorithm:
    histogram with one bin per integer from image.min() to image.max(), class weights / means by cumulative
    sums in float64, threshold = bin centre maximising weight1 * weight2 * (mean1 - mean2)^2."""
    image = np.asarray(image)
    lo, hi = int(image.min()), int(image.max())
    if lo == hi:
        return image.dtype.type(lo)
    counts = np.bincount(image.ravel().astype(np.int64) - lo, minlength=hi - lo + 1).astype(np.float64)
    centers = np.arange(lo, hi + 1)
    weight1 = np.cumsum(counts)
    weight2 = np.cumsum(counts[::-1])[::-1]
    mean1 = np.cumsum(counts * centers) / weight1
    mean2 = (np.cumsum((counts * centers)[::-1]) / weight2[::-1])[::-1]
    variance12 = weight1[:-1] * weight2[1:] * (mean1[:-1] - mean2[1:]) ** 2
    return centers[np.argmax(variance12)]


def segment(image, threshold_value):
    if threshold_value is None:
        T = threshold_otsu_u8(image)
    else:
        T = int(threshold_value)
    return image > T, T


def add_cap(I, cap_thickness, cap_val):  # noqa: E741
    I_cap = np.ones([I.shape[0] + 2 * cap_thickness, I.shape[1], I.shape[2]], I.dtype) * cap_val
    I_cap[cap_thickness:-cap_thickness, :, :] = I
    return I_cap


# --------------------------------------------------------------------------------------------
# gaussian <- pipeline.py:146-148: skimage.filters.gaussian(I, sigma, preserve_range=True), i.e.
# scipy.ndimage.gaussian_filter(I.astype(float64), sigma, mode='nearest', truncate=4.0) (scikit-image's
# defaults).  Restated in NumPy with the order of operations of scipy's NI_Correlate1D symmetric loop;
# tests pin it bit-for-bit against the installed scipy (the dependency the reference itself imports).

def gaussian(image, sigma, truncate=4.0):
    a = np.asarray(image).astype(np.float64)
    sig = [float(sigma)] * a.ndim if np.ndim(sigma) == 0 else [float(v) for v in sigma]
    for axis in range(a.ndim):
        sd = sig[axis]
        if not sd > 1e-15:
            continue
        lw = int(truncate * sd + 0.5)
        x = np.arange(-lw, lw + 1)
        w = np.exp(-0.5 / (sd * sd) * x ** 2)
        w = (w / w.sum())[::-1]
        n = a.shape[axis]
        idx = np.arange(n)
        acc = np.take(a, idx, axis=axis) * w[lw]
        for jj in range(-lw, 0):
            acc = acc + (np.take(a, np.clip(idx + jj, 0, n - 1), axis=axis) +
                         np.take(a, np.clip(idx - jj, 0, n - 1), axis=axis)) * w[lw + jj]
        a = acc
    return a


# --------------------------------------------------------------------------------------------
# embed <- src/ciclope/utils/preprocess.py:233-355 (with recon_utils.bbox incl. its padding rules,
# recon_utils.py:258-298)

_SPLINE_POLE2 = math.sqrt(8.0) - 3.0


def _spline_filter_axis(a, axis):
    """scipy ni_splines.c apply_filter for order 2 (one pole), mirror boundary -- what spline_filter uses for
    mode='constant': gain, causal recursion from the mirror sum, anti-causal recursion."""
    a = np.moveaxis(a, axis, 0).copy()
    n = a.shape[0]
    if n < 2:
        return np.moveaxis(a, 0, axis)
    z = _SPLINE_POLE2
    a *= (1.0 - z) * (1.0 - 1.0 / z)
    z_n_1 = z ** (n - 1)
    c0 = a[0] + z_n_1 * a[n - 1]
    z_i = z
    for i in range(1, n - 1):
        c0 = c0 + z_i * (a[i] + z_n_1 * a[n - 1 - i])
        z_i *= z
    a[0] = c0 / (1 - z_n_1 * z_n_1)
    for i in range(1, n):
        a[i] += z * a[i - 1]
    a[n - 1] = (z * a[n - 2] + a[n - 1]) * z / (z * z - 1)
    for i in range(n - 2, -1, -1):
        a[i] = z * (a[i + 1] - a[i])
    return np.moveaxis(a, 0, axis)


def resample(image, voxelsize, resampling_factor):
    """preprocess.py:49-75: ndimage.zoom(image, 1 / resampling_factor, output=None, order=2) restated (scipy
    ni_interpolation.c NI_ZoomShift, mode='constant', cval 0, grid_mode False) + voxelsize * factor.
    Pinned at tolerance 1e-12 (float64) / exactly (uint8) by tests/golden/resample.npz, made with scipy."""
    inp = np.asarray(image)
    zoom = 1 / resampling_factor
    out_shape = tuple(int(round(s * zoom)) for s in inp.shape)
    f = inp.astype(np.float64)
    for ax in range(inp.ndim):
        f = _spline_filter_axis(f, ax)
    out = f
    for ax in range(inp.ndim):
        n_in, n_out = inp.shape[ax], out_shape[ax]
        zz = (n_in - 1) / (n_out - 1) if n_out > 1 else 1.0
        idx = np.zeros((n_out, 3), dtype=np.int64)
        w = np.zeros((n_out, 3))
        zero = np.zeros(n_out, dtype=bool)
        for k in range(n_out):
            cc = k * zz
            if cc < 0 or cc > n_in - 1:           # map_coordinate, NI_EXTEND_CONSTANT
                zero[k] = True
                continue
            start = int(math.floor(cc + 0.5)) - 1
            d = cc - math.floor(cc + 0.5)
            w[k] = [0.5 * (0.5 - d) ** 2, 0.75 - d * d, 0.5 * (0.5 + d) ** 2]
            for tap in range(3):
                i = start + tap
                if n_in <= 1:
                    i = 0
                else:                             # spline mode: mirror
                    s2 = 2 * n_in - 2
                    if i < 0:
                        i = s2 * int(-i / s2) + i
                        i = i + s2 if i <= 1 - n_in else -i
                    elif i >= n_in:
                        i -= s2 * int(i / s2)
                        if i >= n_in:
                            i = s2 - i
                idx[k, tap] = i
        g = np.take(out, idx.reshape(-1), axis=ax)
        shp = list(out.shape)
        shp[ax:ax + 1] = [n_out, 3]
        wshape = [1] * (out.ndim + 1)
        wshape[ax], wshape[ax + 1] = n_out, 3
        g = (g.reshape(shp) * w.reshape(wshape)).sum(axis=ax + 1)
        zshape = [1] * g.ndim
        zshape[ax] = n_out
        out = np.where(zero.reshape(zshape), 0.0, g)
    if inp.dtype == np.uint8:
        t = np.where(out > 0, out + 0.5, 0.0)
        out = np.minimum(t, 255.0).astype(np.uint8)
    return out, voxelsize * resampling_factor


def bbox_pad(bw, pad=0):
    maxROW = np.max(np.max(bw, 0), 1)
    maxCOL = np.max(np.max(bw, 0), 0)
    maxSLICE = np.max(np.max(bw, 1), 1)
    row0 = list(maxROW).index(True)
    row1 = len(maxROW) - list(maxROW[::-1]).index(True) - 1
    col0 = list(maxCOL).index(True)
    col1 = len(maxCOL) - list(maxCOL[::-1]).index(True) - 1
    slice0 = list(maxSLICE).index(True)
    slice1 = len(maxSLICE) - list(maxSLICE[::-1]).index(True) - 1
    row0 = row0 - pad
    rowd = row1 - row0 + pad
    col0 = col0 - pad
    cold = col1 - col0 + pad
    slice0 = slice0 - pad
    sliced = slice1 - slice0 + pad
    if pad > 0:
        row0, col0, slice0 = max(row0, 0), max(col0, 0), max(slice0, 0)
        if slice0 + sliced > bw.shape[0]:
            sliced = bw.shape[0] - slice0
        if row0 + rowd > bw.shape[1]:
            rowd = bw.shape[1] - row0
        if col0 + cold > bw.shape[2]:
            cold = bw.shape[2] - col0
    return [row0, col0, slice0], [rowd, cold, sliced]


def embed_box(BW_I, embed_depth, embed_dir, pad=0):
    """The three slices of the embedding box (preprocess.py:268-325), before the emboss."""
    if embed_dir == "-z":
        start = np.where(np.max(BW_I.max(1), 1) == True)[-1][-1]  # noqa: E712
        o, s = bbox_pad(BW_I[start - embed_depth:, :, :], pad)
        return (slice(start - embed_depth, None), slice(o[0], o[0] + s[0]), slice(o[1], o[1] + s[1]))
    if embed_dir == "+z":
        start = np.where(np.max(BW_I.max(1), 1) == True)[0][0]  # noqa: E712
        o, s = bbox_pad(BW_I[:start + embed_depth, :, :], pad)
        return (slice(None, start + embed_depth), slice(o[0], o[0] + s[0]), slice(o[1], o[1] + s[1]))
    if embed_dir == "-x":
        start = np.where(np.max(BW_I.max(0), 0) == True)[-1][-1]  # noqa: E712
        o, s = bbox_pad(BW_I[:, :, start - embed_depth:], pad)
        return (slice(o[2], o[2] + s[2]), slice(o[0], o[0] + s[0]), slice(start - embed_depth, None))
    if embed_dir == "+x":
        start = np.where(np.max(BW_I.max(0), 0) == True)[0][0]  # noqa: E712
        o, s = bbox_pad(BW_I[:, :, :start + embed_depth], pad)
        return (slice(o[2], o[2] + s[2]), slice(o[0], o[0] + s[0]), slice(None, start + embed_depth))
    if embed_dir == "+y":
        start = np.where(np.max(BW_I.max(0), 1) == True)[-1][-1]  # noqa: E712
        o, s = bbox_pad(BW_I[:, start - embed_depth:, :], pad)
        return (slice(o[2], o[2] + s[2]), slice(start - embed_depth, None), slice(o[1], o[1] + s[1]))
    if embed_dir == "-y":
        start = np.where(np.max(BW_I.max(0), 1) == True)[0][0]  # noqa: E712
        o, s = bbox_pad(BW_I[:, :start + embed_depth, :])          # the reference passes no pad here
        return (slice(o[2], o[2] + s[2]), slice(None, start + embed_depth), slice(o[1], o[1] + s[1]))
    raise IOError("EMBED_DIR parameter unknown. Valid entries are -x, +x, -y, +y, -z, and +z.")


def embed(I, embed_depth, embed_dir, embed_val=None, pad=0, makecopy=False):  # noqa: E741
    if embed_val is None:
        embed_val = I.max() + 1
    BW_I = np.zeros(I.shape, dtype='bool')
    BW_I[I > 0] = True
    BW_embedding = np.zeros(BW_I.shape, dtype='bool')
    BW_embedding[embed_box(BW_I, embed_depth, embed_dir, pad)] = True
    BW_embedding = remove_unconnected(BW_embedding & ~BW_I)
    if makecopy:
        I_output = I.copy()
        I_output[BW_embedding] = embed_val
        return I_output, BW_embedding
    I[BW_embedding] = embed_val
    return I, BW_embedding


# --------------------------------------------------------------------------------------------
# periosteummask <- src/ciclope/utils/preprocess.py:357-420.  The scikit-image pieces (not in the checkout, not
# installed) are restated from their definitions on top of the installed scipy, which is what they wrap:
#   morphology.disk(r)              = {x^2 + y^2 <= r^2} on a (2r+1)^2 grid
#   morphology.binary_closing(a, f) = ndi.binary_erosion(ndi.binary_dilation(a, f), f, border_value=True)
# PARITY of these two wrappers is unpinned; binary_fill_holes and the loop are the reference's own calls.

def disk(radius):
    L = np.arange(-radius, radius + 1)
    X, Y = np.meshgrid(L, L)
    return (X ** 2 + Y ** 2) <= radius ** 2


def binary_closing(image, footprint):
    from scipy import ndimage
    dilated = ndimage.binary_dilation(image, structure=footprint)
    return ndimage.binary_erosion(dilated, structure=footprint, border_value=True)


def remove_small_objects(ar, min_size):
    """skimage.morphology.remove_small_objects for a boolean image (connectivity=1), restated on scipy."""
    from scipy import ndimage
    out = ar.copy()
    if min_size == 0:
        return out
    ccs = np.zeros_like(ar, dtype=np.int32)
    ndimage.label(ar, ndimage.generate_binary_structure(ar.ndim, 1), output=ccs)
    component_sizes = np.bincount(ccs.ravel())
    too_small = component_sizes < min_size
    out[too_small[ccs]] = 0
    return out


def periosteummask(bwimage, closepixels=10, removeunconn=True, remove_objects_smaller_than=None, closevoxels=0):
    from scipy import ndimage
    if remove_objects_smaller_than:
        bwimage = remove_small_objects(bwimage, remove_objects_smaller_than)
    if bwimage.ndim == 3:
        perimask = np.zeros(np.shape(bwimage), dtype=bool)
        for sl in range(bwimage.shape[0]):
            perimask[sl, :, :] = ndimage.binary_fill_holes(binary_closing(bwimage[sl, :, :], disk(closepixels)))
        if removeunconn:
            perimask = remove_unconnected(perimask)
        if closevoxels > 0:      # preprocess.py:405-408; morphology.cube(w) = ones((w, w, w))
            perimask = binary_closing(perimask, np.ones((closevoxels,) * 3, dtype=bool))
        return perimask
    perimask = ndimage.binary_fill_holes(binary_closing(bwimage, disk(closepixels)))
    if removeunconn:
        perimask = remove_unconnected(perimask[None])[0]
    return perimask
