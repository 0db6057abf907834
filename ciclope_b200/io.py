"""Volume input / mesh output around the hot path.

* ``read_tiff_stack``  drop-in for ``ciclope.utils.recon_utils.read_tiff_stack`` (recon_utils.py:196-230): same
  file selection (every file of the folder of ``filename``, sorted; optional slice range matched on the
  zero-filled number in the file name); the slices are decoded into ONE pinned host buffer and reach HBM with
  a single asynchronous copy, so that ``segment`` / ``remove_unconnected`` / ``vol2ugrid`` start from a device
  tensor.  The reference decodes with tifffile; this build uses Pillow (same pixel values for the 8 / 16-bit
  grey-scale stacks of the reference's data sets).
* ``write_vtk``        legacy VTK (binary, unstructured grid) writer used by ``VoxelMesh.write`` when meshio is
  not installed: points, hexahedra, the GV cell data -- what ``mesh.write('....vtk')`` of the reference's
  notebooks and pipeline (pipeline.py:222-233) leaves for ParaView."""
import os
import re
import warnings

import numpy as np


def stack_files(filename, range=None, zfill=4):  # noqa: A002 -- the reference's parameter name
    """The file list of recon_utils.read_tiff_stack (recon_utils.py:213-227)."""
    folder = os.path.dirname(filename)
    files = [os.path.join(folder, f) for f in os.listdir(folder) if os.path.isfile(os.path.join(folder, f))]
    files.sort()
    if range is not None:
        slice_in = [i for i, item in enumerate(files) if re.search(str(range[0]).zfill(4) + ".", item)]
        slice_end = [i for i, item in enumerate(files) if re.search(str(range[1]).zfill(4) + ".", item)]
        if len(slice_in) == 1 and len(slice_end) == 1:
            files = files[slice_in[0]:slice_end[0]]
        else:
            warnings.warn('Given slice range is ambiguous or non existing.. loading whole stack.')
    return files


def read_tiff_stack(filename, range=None, zfill=4, device="cuda"):  # noqa: A002
    """Read stack of tiff files (drop-in for recon_utils.py:196-230).  Returns a CUDA tensor [slices, rows,
    cols] (``device=None``: a NumPy array, as the reference)."""
    from PIL import Image
    files = stack_files(filename, range, zfill)
    if not files:
        raise ValueError("no files next to %s" % filename)
    first = np.asarray(Image.open(files[0]))
    if device is None:
        out = np.empty((len(files),) + first.shape, dtype=first.dtype)
        host = out
    else:
        import torch
        tdt = {np.dtype(np.uint8): torch.uint8, np.dtype(np.uint16): torch.uint16, np.dtype(np.int16): torch.int16,
               np.dtype(np.int32): torch.int32, np.dtype(np.float32): torch.float32,
               np.dtype(np.bool_): torch.bool}.get(first.dtype)
        if tdt is None:
            raise TypeError("unsupported TIFF sample type %s" % first.dtype)
        pinned = torch.empty((len(files),) + first.shape, dtype=tdt).pin_memory()
        host = pinned.numpy()
    host[0] = first
    for k, f in enumerate(files[1:], start=1):
        a = np.asarray(Image.open(f))
        if a.shape != first.shape or a.dtype != first.dtype:
            raise ValueError("%s: slice shape / type differs from the first slice" % f)
        host[k] = a
    if device is None:
        return out
    t = pinned.to(device, non_blocking=True)
    torch.cuda.current_stream(t.device).synchronize()
    return t


_VTK_TYPES = {np.dtype(np.uint8): "unsigned_char", np.dtype(np.bool_): "unsigned_char", np.dtype(np.int8): "char",
              np.dtype(np.uint16): "unsigned_short", np.dtype(np.int16): "short", np.dtype(np.int32): "int",
              np.dtype(np.uint32): "unsigned_int", np.dtype(np.int64): "vtktypeint64", np.dtype(np.float32): "float",
              np.dtype(np.float64): "double"}


def write_vtk(path, points, cells, cell_data=None, title="ciclope_b200 voxel mesh"):
    """Legacy VTK 4.2 file, BINARY, DATASET UNSTRUCTURED_GRID: ``points`` (n, 3), ``cells`` (n_el, 8) hexahedra
    (VTK cell type 12, the node order of vol2ugrid is VTK's), ``cell_data`` name -> (n_el,) array."""
    points = np.asarray(points)
    cells = np.asarray(cells).reshape(-1, 8)
    if points.shape[0] >= 2 ** 31:
        raise ValueError("legacy VTK files index points with 32-bit integers")
    pts = points.astype(np.float64 if points.dtype.kind != "f" else points.dtype, copy=False)
    with open(path, "wb") as f:
        f.write(b"# vtk DataFile Version 4.2\n" + title.encode()[:200] + b"\nBINARY\nDATASET UNSTRUCTURED_GRID\n")
        f.write(("POINTS %d %s\n" % (len(pts), _VTK_TYPES[pts.dtype])).encode())
        f.write(pts.astype(pts.dtype.newbyteorder(">")).tobytes())
        f.write(("\nCELLS %d %d\n" % (len(cells), 9 * len(cells))).encode())
        rec = np.empty((len(cells), 9), dtype=">i4")
        rec[:, 0] = 8
        rec[:, 1:] = cells
        f.write(rec.tobytes())
        f.write(("\nCELL_TYPES %d\n" % len(cells)).encode())
        f.write(np.full(len(cells), 12, dtype=">i4").tobytes())
        if cell_data:
            f.write(("\nCELL_DATA %d\n" % len(cells)).encode())
            for name, arr in cell_data.items():
                a = np.asarray(arr[0] if isinstance(arr, list) else arr).reshape(-1)
                if a.dtype == np.bool_:
                    a = a.astype(np.uint8)
                f.write(("SCALARS %s %s 1\nLOOKUP_TABLE default\n" % (name, _VTK_TYPES[a.dtype])).encode())
                f.write(a.astype(a.dtype.newbyteorder(">")).tobytes())
                f.write(b"\n")


def read_vtk(path):
    """Reader for the files :func:`write_vtk` writes (tests; not a general VTK parser).
    -> (points, cells (n_el, 8) int64, {name: cell data})"""
    inv = {v: k for k, v in _VTK_TYPES.items() if k != np.dtype(np.bool_)}
    data = open(path, "rb").read()
    pos = 0

    def line():
        nonlocal pos
        while data[pos:pos + 1] == b"\n":
            pos += 1
        end = data.index(b"\n", pos)
        ln = data[pos:end].decode()
        pos = end + 1
        return ln

    assert line().startswith("# vtk DataFile Version")
    line()
    assert line() == "BINARY" and line() == "DATASET UNSTRUCTURED_GRID"
    _, n, typ = line().split()
    dt = inv[typ].newbyteorder(">")
    points = np.frombuffer(data, dtype=dt, count=3 * int(n), offset=pos).reshape(-1, 3).astype(inv[typ])
    pos += 3 * int(n) * dt.itemsize
    _, n_el, size = line().split()
    rec = np.frombuffer(data, dtype=">i4", count=int(size), offset=pos).reshape(int(n_el), 9)
    pos += int(size) * 4
    assert (rec[:, 0] == 8).all()
    _, n_t = line().split()
    types = np.frombuffer(data, dtype=">i4", count=int(n_t), offset=pos)
    pos += int(n_t) * 4
    assert (types == 12).all()
    cd = {}
    if pos < len(data) and data[pos:].strip():
        assert line().split()[0] == "CELL_DATA"
        while pos < len(data) and data[pos:].strip():
            _, name, typ, _ = line().split()
            line()
            dt = inv[typ].newbyteorder(">")
            cd[name] = np.frombuffer(data, dtype=dt, count=int(n_el), offset=pos).astype(inv[typ])
            pos += int(n_el) * dt.itemsize
    return points, rec[:, 1:].astype(np.int64), cd
