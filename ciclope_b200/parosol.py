"""ParOSol HDF5 deck from a 3D volume: drop-in for ``ciclope.core.voxelFE.vol2h5ParOSol``
(reference: src/ciclope/core/voxelFE.py:285-472).

What the reference does, and where it runs here:

* bounding box (``recon_utils.bbox``, recon_utils.py:258-272): the three max-projections are reduced on the
  GPU from the packed bits (``cfe_pack_bits`` + ``cfe_proj_bits``); the reference's ``list.index(True)``
  look-up then runs on those small arrays on the host, literally;
* ``Image`` = ``(sub_block * young_modulus).astype(float32)`` (voxelFE.py:350-357), 4 bytes per voxel of the
  cropped block -- GPU (``cfe_scale_crop_f32``): every voxel byte goes through a 256-entry table that this
  module fills with NumPy's own products, so each float has the reference's bits;
* the boundary-condition node list (voxelFE.py:366-456) is O(N^2): it only looks at the first and last
  ``plane_lock_num`` voxel planes (or ``node_level_lock_num`` node levels) of the cropped block.  Those planes
  are copied off the device and the list is assembled on the host.  The reference builds it as
  ``list(set_of_tuples)``, i.e. in CPython's hash-table order; the same insertion sequence into the same
  interpreter's ``set`` reproduces that order exactly (``node_order="reference"``, the default).
  ``node_order="sorted"`` (additive) skips the Python set and returns the rows sorted by (z, y, x, dof).

``h5py`` writes the container, as in the reference; when it is not installed ``vol2h5ParOSol`` raises
``ImportError`` (the datasets themselves are available from :func:`parosol_datasets`).
"""
import logging

import numpy as np
import torch

from . import _lib
from ._util import as_device_volume, empty_u32

# HDF5 data types of the reference (voxelFE.py:318-320)
H5T_IEEE_F64LE = '<f8'
H5T_IEEE_F32LE = '<f4'
H5T_STD_U16LE = '<u2'


def _projections(vol, is_bool, s):
    """maxROW, maxCOL, maxSLICE of recon_utils.bbox (recon_utils.py:260-262) as small host arrays.

    The reference then looks for ``True`` in them with ``list.index``, i.e. for entries EQUAL TO 1: on a
    grey-scale volume a projection whose maximum is above 1 does not count.  The projections are therefore
    reduced from two bit volumes, ``v > 0`` and (uint8 only) ``v > 1``: max = 0, 1, or 2 standing for "> 1".
    """
    Z, Y, X = (int(v) for v in vol.shape)
    dev = vol.device
    W = (X + 31) // 32
    bits = empty_u32(Z * Y * W, dev)
    levels = []
    for thr in ((0,) if is_bool else (0, 1)):
        _lib.call("cfe_pack_bits", vol.data_ptr(), Z * Y, X, thr, bits.data_ptr(), s)
        rowany = torch.empty(Z * Y, dtype=torch.uint8, device=dev)
        colbits = empty_u32(W, dev)
        _lib.call("cfe_proj_bits", bits.data_ptr(), Z * Y, X, rowany.data_ptr(), colbits.data_ptr(), s)
        levels.append((rowany, colbits))
    out = []
    for rowany, colbits in levels:
        ra = rowany.cpu().numpy().reshape(Z, Y).astype(bool)
        cb = colbits.cpu().numpy().view(np.uint8)[: W * 4]
        col = np.unpackbits(cb, bitorder='little')[:X].astype(bool)
        out.append((ra.any(axis=0), col, ra.any(axis=1)))          # rows (y), cols (x), slices (z)
    if is_bool:
        return out[0]
    return tuple(a.astype(np.uint8) + b.astype(np.uint8) for a, b in zip(out[0], out[1]))


def _planes_to_host(vol, z0, planes, y0, y1, x0, x1):
    """{plane index in the cropped block: bool array (sub_block[i] != 0)} -- O(N^2) bytes off the device."""
    out = {}
    for i in planes:
        out[i] = vol[z0 + i, y0:y1 + 1, x0:x1 + 1].cpu().numpy() != 0
    return out


def _first_last_true(m):
    """recon_utils.py:265-272, literally: first and last entry equal to True (ValueError when none)."""
    lst = list(m)
    first = lst.index(True)
    last = len(m) - list(m[::-1]).index(True) - 1
    return first, last


def _corner_rows(els, dirs):
    """Insertion sequence of _collect_from_elements (preprocess.py:449-456): for every element its eight
    corners (dz, dy, dx nested) times `dirs`, as an int array [n, 4]."""
    n = len(els)
    nd = len(dirs)
    corners = np.array([(dz, dy, dx) for dz in (0, 1) for dy in (0, 1) for dx in (0, 1)], dtype=np.int64)
    out = np.empty((n, 8, nd, 4), dtype=np.int64)
    out[..., :3] = (els[:, None, :] + corners[None, :, :])[:, :, None, :]
    out[..., 3] = np.asarray(dirs, dtype=np.int64)[None, None, :]
    return out.reshape(-1, 4)


_REPLAY_OK = None


def _pyset_replay(rows):
    """list(set(map(tuple, rows))) through the C replay of CPython's hash table (csrc/pyset.cu)."""
    import ctypes
    rows = np.ascontiguousarray(rows, dtype=np.int64)
    out = np.empty_like(rows)
    n = ctypes.c_int64(0)
    rc = _lib.load().cfe_host_pyset_order(rows.ctypes.data, len(rows), out.ctypes.data, ctypes.addressof(n))
    if rc != 0:
        raise _lib.CfeError(_lib.load().cfe_last_error().decode())
    return out[:n.value]


def _replay_matches_interpreter():
    """The replay is only used when it reproduces THIS interpreter's set order on a sample that crosses
    every growth rule of the table (checked once per process)."""
    global _REPLAY_OK
    if _REPLAY_OK is None:
        rng = np.random.default_rng(12345)
        ok = True
        for n, hi in ((300, 6), (20000, 40), (150000, 70)):
            rows = np.stack([rng.integers(0, hi, n), rng.integers(0, hi, n), rng.integers(0, hi, n),
                             rng.integers(0, 3, n)], axis=1)
            want = np.array(list(set(map(tuple, rows.tolist()))), dtype=np.int64).reshape(-1, 4)
            try:
                got = _pyset_replay(rows)
            except Exception:  # noqa: BLE001
                ok = False
                break
            if got.shape != want.shape or not np.array_equal(got, want):
                ok = False
                break
        _REPLAY_OK = ok
    return _REPLAY_OK


def _ordered_unique(rows):
    """The reference's ``list(set)`` of the row tuples (CPython set order); an int64 array [n, 4]."""
    if len(rows) == 0:
        return np.zeros((0, 4), dtype=np.int64)
    if rows.min() >= 0 and _replay_matches_interpreter():
        return _pyset_replay(rows)
    s = set()
    s.update(map(tuple, rows.tolist()))       # one add per row, in order, like the reference's loop
    return np.array(list(s), dtype=np.int64).reshape(-1, 4)


def parosol_datasets(voldata, topDisplacement, voxelsize=1, poisson_ratio=0.3, young_modulus=18e3,
                     topHorizontaFixedlDisplacement=True, locking_strategy="plane", plane_lock_num=1,
                     node_level_lock_num=1, node_order="reference", image_on_device=False):
    """The five datasets vol2h5ParOSol writes into the ``Image_Data`` group, in the reference's order:
    ``Image``, ``Voxelsize``, ``Poisons_ratio``, ``Fixed_Displacement_Coordinates``,
    ``Fixed_Displacement_Values`` (name -> array).  ``Image`` is a CUDA tensor when ``image_on_device``."""
    if node_order not in ("reference", "sorted"):
        raise ValueError("node_order must be 'reference' or 'sorted'")
    vol, is_bool, _ = as_device_volume(voldata)
    Z, Y, X = (int(v) for v in vol.shape)
    if Z * Y >= 2 ** 32 or max(Z, Y, X) >= 2 ** 31:
        raise ValueError("volume too large for one GPU")
    dev = vol.device
    # value table of  sub_block * young_modulus  cast to float32, by NumPy itself
    if is_bool:
        lut = np.zeros(256, dtype=np.float32)
        lut[:2] = (np.array([False, True]) * young_modulus).astype(np.float32)
    else:
        lut = (np.arange(256, dtype=np.uint8) * young_modulus).astype(np.float32)
    if lut.shape != (256,):
        raise TypeError("young_modulus must be a scalar")
    if vol.numel() == 0:
        raise ValueError("zero-size array to reduction operation maximum which has no identity")
    with torch.cuda.device(dev):
        s = _lib.stream()
        max_row, max_col, max_slice = _projections(vol, is_bool, s)
        y0, y1 = _first_last_true(max_row)           # ValueError('True is not in list') on an empty volume
        x0, x1 = _first_last_true(max_col)
        z0, z1 = _first_last_true(max_slice)
        # the reference slices [origin, origin + dims + 1) with dims = last - first (voxelFE.py:336-352)
        nz, ny, nx = z1 - z0 + 1, y1 - y0 + 1, x1 - x0 + 1
        image = torch.empty((nz, ny, nx), dtype=torch.float32, device=dev)
        dlut = torch.from_numpy(lut).pin_memory().to(dev, non_blocking=True)
        _lib.call("cfe_scale_crop_f32", vol.data_ptr(), Y, X, z0, y0, x0, nz, ny, nx, dlut.data_ptr(),
                  image.data_ptr(), s)

        if locking_strategy == "plane":
            p = plane_lock_num
            low_planes = [i for i in range(p)]
            top_planes = [nz - 1 - i for i in range(p)]
            for i in low_planes + top_planes:
                if not -nz <= i < nz:                           # sub_block[i] in the reference
                    raise IndexError("index {} is out of bounds for axis 0 with size {}".format(i, nz))
            low_planes = [i % nz for i in low_planes]
            top_planes = [i % nz for i in top_planes]
            need = sorted(set(low_planes + top_planes))
            plane = _planes_to_host(vol, z0, need, y0, y1, x0, x1)

            def elements(planes, labels):
                parts = []
                for i, lab in zip(planes, labels):
                    yx = np.argwhere(plane[i])
                    parts.append(np.hstack((np.full((yx.shape[0], 1), lab), yx)))
                return np.vstack(parts) if parts else np.zeros((0, 3), dtype=np.int64)

            top_dirs = (0, 1, 2) if topHorizontaFixedlDisplacement else (2,)
            if node_order == "sorted":
                # no insertion sequence needed: the nodes of the selected planes as boolean node layers
                def plane_nodes(planes, dirs):
                    layers = {}
                    for i in planes:
                        q = plane[i]
                        for dz in (0, 1):
                            ex = layers.setdefault(i + dz, np.zeros((ny + 1, nx + 1), dtype=bool))
                            ex[:-1, :-1] |= q
                            ex[:-1, 1:] |= q
                            ex[1:, :-1] |= q
                            ex[1:, 1:] |= q
                    rows = []
                    for z in sorted(layers):
                        yx = np.argwhere(layers[z])
                        rows.append(np.hstack((np.full((yx.shape[0], 1), z), yx)))
                    nodes = np.vstack(rows) if rows else np.zeros((0, 3), dtype=np.int64)
                    out = np.empty((len(nodes), len(dirs), 4), dtype=np.int64)
                    out[..., :3] = nodes[:, None, :]
                    out[..., 3] = np.asarray(dirs, dtype=np.int64)[None, :]
                    return out.reshape(-1, 4)

                lb_rows = plane_nodes(low_planes, (0, 1, 2))
                tb_rows = plane_nodes(top_planes, top_dirs)
            else:
                # the reference labels the rows with i / idx as computed, before any wrap-around
                low = elements(low_planes, [i for i in range(p)])
                top = elements(top_planes, [nz - 1 - i for i in range(p)])
                lb_rows = _corner_rows(low, (0, 1, 2))
                tb_rows = _corner_rows(top, top_dirs)
        elif locking_strategy == "exact":
            k = node_level_lock_num
            lo_levels = [z for z in range(nz + 1) if z < k]
            hi_levels = [z for z in range(nz + 1) if z >= (nz + 1) - k]
            need = sorted({v for z in lo_levels + hi_levels for v in (z - 1, z) if 0 <= v < nz})
            plane = _planes_to_host(vol, z0, need, y0, y1, x0, x1)

            def level_nodes(levels):
                rows = []
                for z in levels:
                    ex = np.zeros((ny + 1, nx + 1), dtype=bool)
                    for v in (z - 1, z):
                        if 0 <= v < nz:
                            q = plane[v]
                            ex[:-1, :-1] |= q
                            ex[:-1, 1:] |= q
                            ex[1:, :-1] |= q
                            ex[1:, 1:] |= q
                    yx = np.argwhere(ex)
                    rows.append(np.hstack((np.full((yx.shape[0], 1), z), yx)))
                return np.vstack(rows) if rows else np.zeros((0, 3), dtype=np.int64)

            def with_dirs(nodes, dirs):
                out = np.empty((len(nodes), len(dirs), 4), dtype=np.int64)
                out[..., :3] = nodes[:, None, :]
                out[..., 3] = np.asarray(dirs, dtype=np.int64)[None, :]
                return out.reshape(-1, 4)

            lb_rows = with_dirs(level_nodes(lo_levels), (0, 1, 2))
            tb_rows = with_dirs(level_nodes(hi_levels), (0, 1, 2) if topHorizontaFixedlDisplacement else (2,))
        else:
            raise ValueError("locking_strategy must be 'plane' or 'exact'")

    if node_order == "sorted":
        lb, tb = lb_rows, tb_rows               # distinct and in (z, y, x, dof) order by construction
    else:
        lb = _ordered_unique(lb_rows)
        tb = _ordered_unique(tb_rows)
    all_nodes = np.concatenate([lb, tb])                             # voxelFE.py:449
    n_all = len(all_nodes)
    # np.array(list_of_tuples, dtype=uint16): (n, 4), or (0,) when there is no node at all
    if n_all and (all_nodes.min() < 0 or all_nodes.max() > 65535):
        raise OverflowError("Python integer {} out of bounds for uint16".format(
            int(all_nodes.max() if all_nodes.max() > 65535 else all_nodes.min())))
    fix_coord = all_nodes.astype(np.uint16) if n_all else np.array([], dtype=np.uint16)
    values = np.zeros(n_all, dtype=np.float32)
    if len(tb):
        values[len(lb):][tb[:, 3] == 2] = topDisplacement            # voxelFE.py:453-456
    out = {}
    out["Image"] = image if image_on_device else image.cpu().numpy()
    out["Voxelsize"] = np.asarray(voxelsize, dtype=H5T_IEEE_F64LE)
    out["Poisons_ratio"] = np.asarray(poisson_ratio, dtype=H5T_IEEE_F64LE)
    out["Fixed_Displacement_Coordinates"] = fix_coord
    out["Fixed_Displacement_Values"] = values
    return out


def vol2h5ParOSol(voldata, fileout, topDisplacement, voxelsize=1, poisson_ratio=0.3, young_modulus=18e3,
                  topHorizontaFixedlDisplacement=True, locking_strategy="plane", plane_lock_num=1,
                  node_level_lock_num=1, verbose=False, node_order="reference"):
    """Generate ParOSol HDF5 (.h5) input file from 3D volume data (drop-in for voxelFE.py:285-472;
    parameters are the reference's, ``node_order`` is additive).  ``voldata``: bool / uint8 NumPy array or
    CUDA tensor."""
    if verbose:
        logging.basicConfig(level=logging.INFO)
    try:
        import h5py
    except ImportError as e:
        raise ImportError("vol2h5ParOSol writes its file with h5py, which is not installed; "
                          "ciclope_b200.parosol.parosol_datasets() returns the datasets") from e
    logging.info('Opening Output file')
    file = h5py.File(fileout, 'w')
    logging.info('creating h5 Image_Data Group')
    img_data = file.create_group("Image_Data")
    ds = parosol_datasets(voldata, topDisplacement, voxelsize=voxelsize, poisson_ratio=poisson_ratio,
                          young_modulus=young_modulus,
                          topHorizontaFixedlDisplacement=topHorizontaFixedlDisplacement,
                          locking_strategy=locking_strategy, plane_lock_num=plane_lock_num,
                          node_level_lock_num=node_level_lock_num, node_order=node_order)
    img_data.create_dataset("Image", data=ds["Image"], dtype=H5T_IEEE_F32LE)
    logging.info('Setting voxel size')
    img_data.create_dataset("Voxelsize", data=voxelsize, dtype=H5T_IEEE_F64LE)
    logging.info('Setting Poisson ratio')
    img_data.create_dataset("Poisons_ratio", data=poisson_ratio, dtype=H5T_IEEE_F64LE)
    img_data.create_dataset("Fixed_Displacement_Coordinates", data=ds["Fixed_Displacement_Coordinates"],
                            dtype=H5T_STD_U16LE)
    img_data.create_dataset("Fixed_Displacement_Values", data=ds["Fixed_Displacement_Values"],
                            dtype=H5T_IEEE_F32LE)
    file.close()
    logging.info('h5 export done!')
