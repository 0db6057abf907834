"""TEST INFRASTRUCTURE ONLY -- import shim that runs the UNMODIFIED ciclope reference.

Only usable where /root/reference exists (the authoring container). It is used by
``tests/golden/make_golden.py`` to generate the committed golden vectors and by
``oracle/check_oracle_vs_reference.py`` to pin the C oracle; nothing in the product
(``ciclope_b200/``), the ``-m gpu`` tests, ``smoke()`` or ``bench.py`` imports it.

The reference (``src/ciclope/core/voxelFE.py``, ``src/ciclope/utils/preprocess.py``)
imports h5py / skimage / imageio / png / tifffile / matplotlib / meshio at module
level; none is installed here and there is no network.  The hot path only *uses*
``skimage.measure.label`` (preprocess.py:135) and ``meshio.Mesh`` as a passive
container (voxelFE.py:271-274), so we pre-populate ``sys.modules`` with:

* empty modules for the unused imports,
* ``skimage.measure.label(img, background, return_num, connectivity)`` ->
  ``scipy.ndimage.label`` (6-neighbourhood == connectivity 1; both number the
  components by first raster appearance),
* a ``meshio`` with the meshio-5.0.0 ``CellBlock`` namedtuple and a ``Mesh``
  constructor doing what 5.0.0 does to these arguments.
"""
import collections
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("CICLOPE_REFERENCE_ROOT", "/root/reference")
H5PY_FILES = {}      # datasets "written" through the h5py stand-in, by file path


def _install_stubs():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    for name in ("imageio", "png", "tifffile"):
        if name not in sys.modules:
            mod(name)
    if "h5py" not in sys.modules:
        # recording stand-in for the three h5py calls vol2h5ParOSol makes (voxelFE.py:328-330, 357-471):
        # H5PY_FILES[path][dataset name] = np.array(data, dtype=dtype), which is what h5py stores
        class _Group:
            def __init__(self, store):
                self._store = store

            def create_dataset(self, name, data=None, dtype=None):
                self._store[name] = np.array(data, dtype=dtype)

        class File:
            def __init__(self, path, mode):
                assert mode == 'w'
                self._store = H5PY_FILES[path] = collections.OrderedDict()

            def create_group(self, name):
                assert name == "Image_Data"
                return _Group(self._store)

            def close(self):
                pass

        mod("h5py", File=File)
    if "matplotlib" not in sys.modules:
        mpl = mod("matplotlib")
        mpl.pyplot = mod("matplotlib.pyplot")

    import scipy.ndimage as ndi

    def label(img, background=None, return_num=False, connectivity=None):
        assert connectivity == 1, "shim only restates connectivity=1 (6-neighbourhood)"
        a = np.asarray(img)
        if a.dtype == np.bool_:
            lab, n = ndi.label(a)
            return (lab, n) if return_num else lab
        # scikit-image labels regions of EQUAL value (background=None: 0 is the background), numbered by
        # first appearance in raster order
        bg = 0 if background is None else background
        lab = np.zeros(a.shape, dtype=np.int64)
        n = 0
        for v in np.unique(a):
            if v == bg:
                continue
            part, k = ndi.label(a == v)
            lab[part > 0] = part[part > 0] + n
            n += k
        if n:
            flat = lab.ravel()
            idx = np.flatnonzero(flat)
            uniq, first = np.unique(flat[idx], return_index=True)
            remap = np.zeros(n + 1, dtype=np.int64)
            remap[uniq[np.argsort(first, kind="stable")]] = np.arange(1, n + 1)
            lab = remap[lab]
        return (lab, n) if return_num else lab

    measure = mod("skimage.measure", label=label)
    # periosteummask (preprocess.py:385-400) calls these scikit-image functions; restated from their
    # definitions on top of scipy (oracle.disk / oracle.binary_closing carry the same statements)
    def _disk(radius):
        L = np.arange(-radius, radius + 1)
        X, Y = np.meshgrid(L, L)
        return (X ** 2 + Y ** 2) <= radius ** 2

    def _binary_closing(image, footprint=None):
        dilated = ndi.binary_dilation(image, structure=footprint)
        return ndi.binary_erosion(dilated, structure=footprint, border_value=True)

    def _remove_small_objects(ar, min_size=64, connectivity=1):
        out = ar.copy()
        if min_size == 0:
            return out
        ccs = np.zeros_like(ar, dtype=np.int32)
        ndi.label(ar, ndi.generate_binary_structure(ar.ndim, connectivity), output=ccs)
        too_small = np.bincount(ccs.ravel()) < min_size
        out[too_small[ccs]] = 0
        return out

    def _cube(width):
        return np.ones((width, width, width), dtype=np.uint8)

    morphology = mod("skimage.morphology", disk=_disk, cube=_cube, binary_closing=_binary_closing,
                     remove_small_objects=_remove_small_objects)
    filters = mod("skimage.filters", threshold_otsu=None, gaussian=None)
    mod("skimage", measure=measure, morphology=morphology, filters=filters)

    CellBlock = collections.namedtuple("CellBlock", ["type", "data"])

    class Mesh:  # meshio 5.0.0 constructor semantics for the arguments vol2ugrid passes
        def __init__(self, points, cells, point_data=None, cell_data=None, field_data=None,
                     point_sets=None, cell_sets=None, gmsh_periodic=None, info=None):
            self.points = np.asarray(points)
            if isinstance(cells, dict):
                cells = list(cells.items())
            self.cells = [CellBlock(t, np.asarray(d)) for t, d in cells]
            self.point_data = {} if point_data is None else point_data
            self.cell_data = {} if cell_data is None else cell_data
            self.field_data = {} if field_data is None else field_data
            self.point_sets = {} if point_sets is None else point_sets
            self.cell_sets = {} if cell_sets is None else cell_sets
            for key, data in self.cell_data.items():
                self.cell_data[key] = [np.asarray(d) for d in data]

        @property
        def cells_dict(self):
            d = {}
            for t, data in self.cells:
                d.setdefault(t, []).append(data)
            return {k: np.concatenate(v) for k, v in d.items()}

        def __repr__(self):
            return "<meshio mesh object>\n  Number of points: {}\n  Number of cells:\n{}".format(
                len(self.points), "".join("    {}: {}\n".format(t, len(d)) for t, d in self.cells))

    mod("meshio", Mesh=Mesh, CellBlock=CellBlock)


def load_reference():
    """Return (voxelFE module, preprocess module) of the unmodified reference."""
    if not os.path.isdir(REFERENCE_ROOT):
        raise RuntimeError("reference checkout not present at %s" % REFERENCE_ROOT)
    _install_stubs()
    src = os.path.join(REFERENCE_ROOT, "src")
    if src not in sys.path:
        sys.path.insert(0, src)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import ciclope.core.voxelFE as voxelFE
        import ciclope.utils.preprocess as preprocess
    return voxelFE, preprocess


def read_tiff_stack(one_file):
    """PIL stand-in for recon_utils.read_tiff_stack (recon_utils.py:196-230)."""
    from PIL import Image
    d = os.path.dirname(one_file)
    files = sorted(os.path.join(d, f) for f in os.listdir(d) if os.path.isfile(os.path.join(d, f)))
    return np.stack([np.array(Image.open(f)) for f in files])
