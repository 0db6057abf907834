"""B200-native drop-ins for the connected-component functions of ``ciclope.utils.preprocess``:
``remove_unconnected`` (reference: src/ciclope/utils/preprocess.py:119-143), ``fill_voids``
(:145-185) and ``remove_largest`` (:187-208).  All three are one labelling pass (``cfe_ccl_*``)
followed by a different selection in the mask-writing kernel."""
import numpy as np
import torch

from . import _lib
from ._util import as_device_volume, check_slab_size, empty_u32, packed_of, read_back, scan_workspace, tag_packed


def remove_unconnected(bwimage):
    """Remove all unconnected voxels. Returns a binary of the largest connected cluster.

    Same contract as the reference (preprocess.py:119-143): 6-connectivity
    (``measure.label(bwimage, None, True, 1)``), the component with the most voxels is kept,
    ties go to the component met first in C-order raster scan, an all-background image raises
    ``ValueError``.  The result is a boolean array of the input's shape: a ``numpy.ndarray``
    for NumPy input, a CUDA ``torch.Tensor`` for tensor input (additive, zero-copy).

    Parameters
    ----------
    bwimage
        Binary image: bool, or uint8 with ONE non-zero value (0/1, 0/255 ...).  ``measure.label``
        labels regions of *equal value*, so a uint8 image with several non-zero values is not a binary
        image to the reference; it raises ``ValueError`` here instead of being
        read as ``!= 0``.

    Returns
    -------
    bwcluster
        Binary image of the largest connected cluster of voxels.
    """
    vol, is_bool, from_numpy = as_device_volume(bwimage, allow_2d=True)
    squeeze = (bwimage.ndim == 2)
    span = None if is_bool else _value_span(vol)
    mask, best = largest_component_device(vol)
    _check_binary(span, "remove_unconnected")
    if best == 0:
        # reference: occurrences[1:].argmax() on an empty array (preprocess.py:141)
        raise ValueError("attempt to get argmax of an empty sequence")
    out = mask.view(torch.bool)
    if squeeze:
        out = out[0]
    if from_numpy:
        return out.cpu().numpy()
    kept = packed_of(mask)
    if kept is not None and not squeeze:
        tag_packed(out, kept)
    return out


def remove_largest(bwimage):
    """Remove largest cluster of voxels in binary image (drop-in for preprocess.py:187-208).

    Parameters
    ----------
    bwimage
        Binary image (bool / uint8; NumPy array or CUDA tensor).

    Returns
    -------
    bwcluster
        Binary image in which the largest cluster of voxels is removed.
    """
    vol, is_bool, from_numpy = as_device_volume(bwimage, allow_2d=True)
    squeeze = (bwimage.ndim == 2)
    span = None if is_bool else _value_span(vol)
    mask, best = largest_component_device(vol, drop=True)
    _check_binary(span, "remove_largest")
    if best == 0:
        raise ValueError("attempt to get argmax of an empty sequence")   # preprocess.py:204
    out = mask.view(torch.bool)
    if squeeze:
        out = out[0]
    return out.cpu().numpy() if from_numpy else out


def fill_voids(I, fill_val=None, makecopy=False):  # noqa: E741 -- the reference's parameter name
    """Fill voids within color image with given value (drop-in for preprocess.py:145-185).

    The background ``~(I>0)`` is labelled with 6-connectivity; every background region except
    the largest one (the outside) is set to ``fill_val`` (default ``I.max()``).  As in the
    reference the input is modified in place and returned unless ``makecopy`` is true.

    Parameters
    ----------
    I
        Input image (bool / uint8; NumPy array or CUDA tensor).
    fill_val
        Filling value.
    makecopy : bool
        Make copy of input image.
    """
    vol, is_bool, from_numpy = as_device_volume(I, allow_2d=True)
    if fill_val is not None:
        if is_bool:
            fill = int(bool(fill_val))
        else:
            fill = int(fill_val)
            if fill != fill_val or not 0 <= fill <= 255:
                raise ValueError("fill_val %r cannot be stored in a uint8 image" % (fill_val,))
    else:
        fill = None
    in_place = isinstance(I, torch.Tensor) and I.is_cuda and I.is_contiguous() and not makecopy
    filled, best = largest_component_device(vol, drop=True, invert=True, src=vol, fill=fill,
                                            out=vol if in_place else None)
    if best == 0:
        raise ValueError("attempt to get argmax of an empty sequence")   # preprocess.py:175
    if is_bool:
        filled = filled.view(torch.bool)
    filled = filled.reshape(I.shape)
    if from_numpy:
        res = filled.cpu().numpy()
        if makecopy:
            return res
        I[...] = res
        return I
    if makecopy or in_place:
        return I if in_place else filled.to(I.device)
    I.copy_(filled)
    return I


def _value_span(vol):
    """Enqueue the (max, min non-zero) reduction of a uint8 volume; read with _check_binary after the
    labelling's own host read, so that the stream is synchronised once."""
    span = torch.empty(2, dtype=torch.int32, device=vol.device)
    with torch.cuda.device(vol.device):
        _lib.call("cfe_value_span_u8", vol.data_ptr(), vol.numel(), span.data_ptr(), _lib.stream())
    return span


def _check_binary(span, who):
    if span is None:
        return
    hi, lo = (int(v) & 0xFFFFFFFF for v in read_back(span))
    if hi and lo != hi:
        raise ValueError("%s: uint8 image holds several non-zero values (%d..%d); the reference labels regions "
                         "of equal value (measure.label), which this build does not: pass a bool image or one "
                         "with a single non-zero value" % (who, lo, hi))


def largest_component_device(vol, drop=False, invert=False, src=None, fill=1, out=None, info=None):
    """vol: uint8 CUDA tensor [Z,Y,X].  Returns (uint8 tensor, best) where
    best = (size << 32) | (0xffffffff - root run id), 0 when nothing was labelled.

    The tensor is the mask of the largest 6-connected component of ``vol != 0`` -- or, with
    ``drop``, of every labelled voxel outside it; ``invert`` labels ``vol == 0`` instead; with
    ``src`` the unselected voxels keep their ``src`` value and the selected ones get ``fill``
    (None: the maximum of ``src``, computed on the device).  One device->host read (16 bytes).
    ``info`` (a dict) receives the union-find arrays of the labelling (diagnostics)."""
    lib = _lib.require_cuda()
    Z, Y, X = (int(v) for v in vol.shape)
    check_slab_size(Z, Y, X)
    dev = vol.device
    if out is None:
        out = torch.empty((Z, Y, X), dtype=torch.uint8, device=dev)
    if Z * Y * X == 0:
        return out, 0
    with torch.cuda.device(dev):
        s = _lib.stream()
        W = (X + 31) // 32
        rows = Z * Y
        bits = empty_u32(rows * W, dev)
        # the first kernel goes out before anything else is allocated: the GPU is idle until then
        _lib.call("cfe_pack_bits_le" if invert else "cfe_pack_bits", vol.data_ptr(), rows, X, 0, bits.data_ptr(), s)
        pref = empty_u32(rows * W, dev)
        stats = torch.empty(2, dtype=torch.int64, device=dev)  # [n_runs, best]: both written by the kernels
        ws = scan_workspace(rows * W, dev)
        _lib.call("cfe_scan_run_starts", bits.data_ptr(), rows, X, pref.data_ptr(), stats.data_ptr(),
                  ws.data_ptr(), s)
        max_runs = lib.cfe_ccl_max_runs(rows, X)
        parent = empty_u32(max_runs, dev)
        size = empty_u32(max_runs, dev)
        if drop or src is not None:
            if fill is None:
                # I.max() on the device: the label/select kernels run first, the fill value is then read
                # back with the winner (one synchronisation) and the selection pass launched
                mx = torch.zeros(1, dtype=torch.int32, device=dev)
                _lib.call("cfe_max_u8", src.data_ptr(), Z * Y * X, mx.data_ptr(), s)
                _lib.call("cfe_ccl_label", bits.data_ptr(), pref.data_ptr(), stats.data_ptr(), Z, Y, X,
                          parent.data_ptr(), size.data_ptr(), s)
                _lib.call("cfe_ccl_select", parent.data_ptr(), size.data_ptr(), stats.data_ptr(),
                          stats.data_ptr() + 8, s)
                fill = int(read_back(mx)[0])
                _lib.call("cfe_ccl_write_select", bits.data_ptr(), pref.data_ptr(), Z, Y, X, parent.data_ptr(),
                          stats.data_ptr() + 8, None, int(drop), _lib.ptr(src), fill, out.data_ptr(), s)
            else:
                _lib.call("cfe_ccl_largest_select", bits.data_ptr(), pref.data_ptr(), stats.data_ptr(), Z, Y, X,
                          parent.data_ptr(), size.data_ptr(), stats.data_ptr() + 8, int(drop), _lib.ptr(src),
                          int(fill), out.data_ptr(), s)
        elif X % 32 == 0 and out.data_ptr() % 16 == 0:
            # the mask writer also leaves the kept voxels bit-packed: vol2ugrid(out) starts from them
            sel = empty_u32(rows * W, dev)
            _lib.call("cfe_ccl_largest_bits", bits.data_ptr(), pref.data_ptr(), stats.data_ptr(), Z, Y, X,
                      parent.data_ptr(), size.data_ptr(), stats.data_ptr() + 8, out.data_ptr(), sel.data_ptr(), s)
            tag_packed(out, sel)
        else:
            _lib.call("cfe_ccl_largest", bits.data_ptr(), pref.data_ptr(), stats.data_ptr(), Z, Y, X,
                      parent.data_ptr(), size.data_ptr(), stats.data_ptr() + 8, out.data_ptr(), s)
        host = read_back(stats)
    best = int(host[1]) & 0xFFFFFFFFFFFFFFFF
    if info is not None:      # diagnostics (benchmarks count the components): parent[i] == i <=> run i is a root
        info.update(parent=parent, size=size, n_runs=int(host[0]))
    return out, best


def component_stats(best):
    """(size, root run id) of the kept component from the packed `best` word."""
    return best >> 32, 0xFFFFFFFF - (best & 0xFFFFFFFF)


# --------------------------------------------------------------------------------------------
# The two steps right before the hot path in ciclope's pipeline (pipeline.py:163-169)

# dtype of the histogram counts in skimage.filters.threshold_otsu: float64 in scikit-image 0.18 (the version of the
# reference's environment file, envs/thkdebian.yml: 0.18.2, and of its notebooks: 0.18.1), float32 from 0.19 on
# (`_validate_image_histogram` returns `counts.astype('float32')`).  pyproject.toml leaves the version open; on
# volumes above 2^24 voxels the float32 cumulative sums round and the argmax can move by a bin.  Set this to
# numpy.float32 to follow a >= 0.19 installation.
OTSU_COUNTS_DTYPE = np.float64


def threshold_otsu(image):
    """Otsu threshold of a uint8 image, as ``skimage.filters.threshold_otsu`` computes it for integer images
    (one bin per integer from min to max, cumulative sums in ``OTSU_COUNTS_DTYPE``, argmax of the between-class
    variance).  The 256-bin histogram is the GPU's (``cfe_gv_histogram``); the 256-term arithmetic is NumPy's, in
    the library's order of operations.  scikit-image is neither in the reference checkout nor installed here, so
    this function is restated from its published algorithm and its parity is NOT pinned by a golden vector
    (see ``OTSU_COUNTS_DTYPE`` for the one known version dependence)."""
    vol, is_bool, _ = as_device_volume(image, allow_2d=True)
    if vol.numel() == 0:
        raise ValueError("zero-size array to reduction operation minimum which has no identity")
    dev = vol.device
    with torch.cuda.device(dev):
        hist_dev = torch.empty(256, dtype=torch.int64, device=dev)
        _lib.call("cfe_gv_histogram", vol.data_ptr(), vol.numel(), -1, hist_dev.data_ptr(), _lib.stream())
        hist = np.asarray(read_back(hist_dev), dtype=np.int64)
    present = np.flatnonzero(hist)
    lo, hi = int(present[0]), int(present[-1])
    if lo == hi:
        return np.bool_(lo) if is_bool else np.uint8(lo)
    counts = hist[lo:hi + 1].astype(OTSU_COUNTS_DTYPE)
    centers = np.arange(lo, hi + 1)
    weight1 = np.cumsum(counts)
    weight2 = np.cumsum(counts[::-1])[::-1]
    mean1 = np.cumsum(counts * centers) / weight1
    mean2 = (np.cumsum((counts * centers)[::-1]) / weight2[::-1])[::-1]
    variance12 = weight1[:-1] * weight2[1:] * (mean1[:-1] - mean2[1:]) ** 2
    return centers[np.argmax(variance12)]


def segment(image, threshold_value):
    """Threshold image (drop-in for preprocess.py:20-46).

    Parameters
    ----------
    image
        Image data (uint8; NumPy array or CUDA tensor).
    threshold_value (optional)
        Threshold value. If empty an Otsu threshold is calculated.

    Returns
    -------
    BWimage
        Binary image after thresholding (``image > T``; a CUDA bool tensor for tensor input).
    T
        Threshold value.
    """
    if _is_float64(image):
        # the smoothed image of pipeline.py:146-148 (gaussian() below)
        if threshold_value is None:
            raise NotImplementedError("Otsu threshold of a float image is not available on the GPU path: "
                                      "pass threshold_value")
        T = int(threshold_value)
        from_numpy = isinstance(image, np.ndarray)
        t = torch.from_numpy(np.ascontiguousarray(image)) if from_numpy else image.contiguous()
        _lib.require_cuda()
        if not t.is_cuda:
            t = t.cuda()
        with torch.cuda.device(t.device):
            out = torch.empty(t.shape, dtype=torch.uint8, device=t.device)
            _lib.call("cfe_threshold_f64", t.data_ptr(), t.numel(), float(T), out.data_ptr(), _lib.stream())
        out = out.view(torch.bool)
        return (out.cpu().numpy() if from_numpy else out), T
    if threshold_value is None:
        T = threshold_otsu(image)
    else:
        T = int(threshold_value)
    vol, _is_bool, from_numpy = as_device_volume(image, allow_2d=True)
    squeeze = (image.ndim == 2)
    dev = vol.device
    with torch.cuda.device(dev):
        out = torch.empty(vol.shape, dtype=torch.uint8, device=dev)
        _lib.call("cfe_threshold_u8", vol.data_ptr(), vol.numel(), max(-1, min(255, int(T))), out.data_ptr(),
                  _lib.stream())
    out = out.view(torch.bool)
    if squeeze:
        out = out[0]
    return (out.cpu().numpy() if from_numpy else out), T


def mask_image(I, L, makecopy=False):  # noqa: E741 -- the reference's variable names
    """``masked = I; masked[~L] = 0``: the grey values outside the mask ``L`` cleared -- the two lines between
    ``remove_unconnected`` and the grey-scale ``mesh2voxelfe`` in the reference's pipeline (pipeline.py:240-242; the
    notebooks mesh ``masked``).  Not a function of the reference: there it is inline NumPy, here one pass of
    ``cfe_masked_keep_u8`` (3 bytes per voxel) instead of a boolean-index assignment.

    ``I``: uint8 image, ``L``: bool / 0-1 uint8 mask of the same shape.  Like the reference's lines the image is
    modified in place and returned (a NumPy array through a device round trip, a contiguous CUDA tensor without
    any copy); ``makecopy=True`` leaves ``I`` alone and returns a new image."""
    vol, is_bool, from_numpy = as_device_volume(I)
    mask, _, _ = as_device_volume(L)
    if mask.shape != vol.shape:
        raise IndexError("boolean index did not match indexed array: image %s, mask %s"
                         % (tuple(vol.shape), tuple(mask.shape)))
    in_place = isinstance(I, torch.Tensor) and I.is_cuda and I.is_contiguous() and not makecopy
    with torch.cuda.device(vol.device):
        out = vol if in_place else torch.empty_like(vol)
        _lib.call("cfe_masked_keep_u8", vol.data_ptr(), mask.data_ptr(), vol.numel(), out.data_ptr(), _lib.stream())
    res = out.view(torch.bool) if is_bool else out
    if from_numpy:
        res_h = res.cpu().numpy()
        if makecopy:
            return res_h
        I[...] = res_h
        return I
    if isinstance(I, torch.Tensor) and not makecopy and not in_place:
        I.copy_(res)          # CPU or non-contiguous tensor: same in-place rule
        return I
    return res


def add_cap(I, cap_thickness, cap_val):  # noqa: E741 -- the reference's parameter name
    """Add caps to 3D image (drop-in for preprocess.py:210-231): ``cap_thickness`` slices of value
    ``cap_val`` before the first and after the last slice.  Pure data movement (two fills and a copy in
    device memory).  uint8 images with a cap value that NumPy keeps in uint8."""
    vol, is_bool, from_numpy = as_device_volume(I)
    probe = np.ones(1, np.bool_ if is_bool else np.uint8) * cap_val      # dtype / range rules of the reference
    if probe.dtype != np.uint8:
        raise TypeError("add_cap on the GPU keeps uint8 images; cap_val %r gives %s" % (cap_val, probe.dtype))
    c = int(cap_thickness)
    Z, Y, X = (int(v) for v in vol.shape)
    if c <= 0 or Z + 2 * c <= 0:
        # I_cap[c:-c] is empty for c = 0 (and the shapes are negative below that): the reference fails here
        raise ValueError("could not broadcast input array from shape ({},{},{}) into shape (0,{},{})"
                         .format(Z, Y, X, Y, X))
    out = torch.empty((Z + 2 * c, Y, X), dtype=torch.uint8, device=vol.device)
    out[:c].fill_(int(probe[0]))
    out[Z + c:].fill_(int(probe[0]))
    out[c:Z + c].copy_(vol)
    return out.cpu().numpy() if from_numpy else out


def _is_float64(image):
    return (isinstance(image, np.ndarray) and image.dtype == np.float64) or \
           (isinstance(image, torch.Tensor) and image.dtype == torch.float64)


def _gaussian_kernel1d(sigma, radius):
    """scipy.ndimage._filters._gaussian_kernel1d(sigma, 0, radius): the same NumPy expressions, so the
    weights have scipy's bits."""
    sigma2 = sigma * sigma
    x = np.arange(-radius, radius + 1)
    phi_x = np.exp(-0.5 / sigma2 * x ** 2)
    return phi_x / phi_x.sum()


def gaussian(image, sigma=1, mode='nearest', cval=0, preserve_range=False, truncate=4.0):
    """Gaussian smoothing of a 3D image: the call ciclope's pipeline makes before segmenting
    (pipeline.py:146-148, ``skimage.filters.gaussian(I, sigma=..., preserve_range=True)``), which is
    ``scipy.ndimage.gaussian_filter(I.astype(float64), sigma, mode='nearest', truncate=4.0)``: one
    float64 pass per axis (0, 1, 2) in scipy's order of operations, so the result is bit-identical to
    scipy's.  ``image``: uint8 or float64, NumPy array or CUDA tensor; returns float64 of the same kind.
    Only ``preserve_range=True`` and ``mode='nearest'`` (what the pipeline uses) are available."""
    if not preserve_range:
        raise NotImplementedError("gaussian: only preserve_range=True (the pipeline's call) is available")
    if mode != 'nearest':
        raise NotImplementedError("gaussian: only mode='nearest' (skimage's default) is available")
    from_numpy = isinstance(image, np.ndarray)
    if _is_float64(image):
        t = torch.from_numpy(np.ascontiguousarray(image)) if from_numpy else image.contiguous()
        is_u8 = False
    else:
        t, is_bool, from_numpy = as_device_volume(image)
        if is_bool:
            raise TypeError("gaussian: uint8 or float64 images")
        is_u8 = True
    if t.dim() != 3:
        raise ValueError("image must be 3-D [slices, rows, cols]")
    sig = np.asarray(sigma, dtype=np.float64)
    if np.any(sig < 0.0):
        raise ValueError("Sigma values less than zero are not valid")
    sigmas = [float(sig)] * 3 if sig.ndim == 0 else [float(v) for v in sig]
    if len(sigmas) != 3:
        raise RuntimeError("sequence argument must have length equal to input rank")
    _lib.require_cuda()
    if not t.is_cuda:
        t = t.cuda()
    dev = t.device
    shape = [int(v) for v in t.shape]
    n_total = t.numel()
    with torch.cuda.device(dev):
        s = _lib.stream()
        cur, cur_u8 = t, is_u8
        for axis in range(3):
            sd = sigmas[axis]
            if not sd > 1e-15:                                  # scipy skips these axes
                continue
            r = int(truncate * sd + 0.5)
            w = _gaussian_kernel1d(sd, r)[::-1]
            dw = torch.from_numpy(np.ascontiguousarray(w)).pin_memory().to(dev, non_blocking=True)
            out = torch.empty(shape, dtype=torch.float64, device=dev)
            stride = int(np.prod(shape[axis + 1:], dtype=np.int64))
            if n_total:
                _lib.call("cfe_gauss1d", cur.data_ptr(), 1 if cur_u8 else 0, out.data_ptr(), n_total, shape[axis],
                          stride, dw.data_ptr(), r, s)
            cur, cur_u8 = out, False
        if cur_u8:                                              # every sigma was 0: the float64 copy
            cur = cur.to(torch.float64)
    return cur.cpu().numpy() if from_numpy else cur


def _zoom_taps(n_in, n_out):
    """Per-axis tables of scipy's NI_ZoomShift for order 2, mode='constant', grid_mode=False: input
    coordinate of output k = k * (n_in - 1) / (n_out - 1); outside [0, n_in - 1] -> the constant (flag);
    three taps around round-half-up(coordinate) with quadratic B-spline weights, taps beyond the edge
    mirrored.  Returns (idx int32 [n_out, 3], w float64 [n_out, 3], zero uint8 [n_out])."""
    k = np.arange(n_out, dtype=np.float64)
    zoom = (n_in - 1) / (n_out - 1) if n_out > 1 else 1.0
    cc = k * zoom
    zero = (cc < 0) | (cc > n_in - 1)
    mid = np.floor(cc + 0.5)
    d = cc - mid
    w = np.stack([0.5 * (0.5 - d) ** 2, 0.75 - d * d, 0.5 * (0.5 + d) ** 2], axis=1)
    idx = mid.astype(np.int64)[:, None] - 1 + np.arange(3)[None, :]
    if n_in <= 1:
        idx[:] = 0
    else:
        s2 = 2 * n_in - 2
        neg = idx < 0
        m = np.where(neg, -idx, idx) % s2            # mirror: period 2 n - 2
        idx = np.where(m >= n_in, s2 - m, m)
    idx = np.where(zero[:, None], 0, idx)
    return idx.astype(np.int32), np.ascontiguousarray(w), zero.astype(np.uint8)


def resample(image, voxelsize, resampling_factor):
    """Resize image (drop-in for preprocess.py:49-75): quadratic spline interpolation,
    ``scipy.ndimage.zoom(image, 1 / resampling_factor, output=None, order=2)``, and the voxel size scaled
    by the factor.

    Parameters
    ----------
    image
        Image data: uint8 or float64 (the Gaussian-smoothed image of the pipeline), 3-D; NumPy array or
        CUDA tensor.
    voxelsize
        Voxel size.
    resampling_factor
        Scaling factor.

    Returns
    -------
    image
        Resized image (dtype of the input, as scipy; a CUDA tensor for tensor input).
    voxelsize
        Voxel size after rescaling.

    Parity: within 1e-12 (absolute, float64 images of 8-bit range) of scipy, identical for uint8 images and
    for the thresholded masks of the test volumes -- not bit-exact on float64 (csrc/filters.cu)."""
    from_numpy = isinstance(image, np.ndarray)
    if _is_float64(image):
        t = torch.from_numpy(np.ascontiguousarray(image)) if from_numpy else image.contiguous()
        is_u8 = False
    else:
        t, is_bool, from_numpy = as_device_volume(image)
        if is_bool:
            raise TypeError("resample: uint8 or float64 images")
        is_u8 = True
    if t.dim() != 3:
        raise ValueError("image must be 3-D [slices, rows, cols]")
    _lib.require_cuda()
    if not t.is_cuda:
        t = t.cuda()
    dev = t.device
    zoom = 1 / resampling_factor
    shape_in = [int(v) for v in t.shape]
    shape_out = [int(round(n * zoom)) for n in shape_in]        # scipy: int(round(ii * jj))
    if min(shape_out) < 1:
        raise RuntimeError("resample: the output would be empty")
    with torch.cuda.device(dev):
        s = _lib.stream()
        coef = t.to(torch.float64) if is_u8 else t.clone()
        for axis in range(3):                                   # spline_filter: one pass per axis
            outer = int(np.prod(shape_in[:axis], dtype=np.int64))
            inner = int(np.prod(shape_in[axis + 1:], dtype=np.int64))
            _lib.call("cfe_spline_prefilter", coef.data_ptr(), outer, shape_in[axis], inner, s)
        tabs = []
        for axis in range(3):
            idx, w, zero = _zoom_taps(shape_in[axis], shape_out[axis])
            tabs.append((torch.from_numpy(idx).to(dev), torch.from_numpy(w).to(dev), torch.from_numpy(zero).to(dev)))
        out = torch.empty(shape_out, dtype=torch.uint8 if is_u8 else torch.float64, device=dev)
        _lib.call("cfe_spline_zoom", coef.data_ptr(), shape_in[1], shape_in[2], shape_out[0], shape_out[1],
                  shape_out[2], tabs[0][0].data_ptr(), tabs[1][0].data_ptr(), tabs[2][0].data_ptr(),
                  tabs[0][1].data_ptr(), tabs[1][1].data_ptr(), tabs[2][1].data_ptr(), tabs[0][2].data_ptr(),
                  tabs[1][2].data_ptr(), tabs[2][2].data_ptr(), 1 if is_u8 else 0, out.data_ptr(), s)
    voxelsize = voxelsize * resampling_factor                   # preprocess.py:73
    return (out.cpu().numpy() if from_numpy else out), voxelsize


# --------------------------------------------------------------------------------------------
# embed (preprocess.py:233-355): another client of the labelling kernels

def _bool_projections(sub):
    """(maxROW, maxCOL, maxSLICE) of ``sub > 0`` as host bool arrays (recon_utils.py:260-262)."""
    from .parosol import _projections
    if sub.numel() == 0:
        raise ValueError("zero-size array to reduction operation maximum which has no identity")
    return _projections(sub, True, _lib.stream())


def _bbox_from_projections(max_row, max_col, max_slice, shape, pad):
    """recon_utils.bbox(bw, pad) (recon_utils.py:264-298) from the three max-projections of ``bw`` (pure host
    arithmetic): ([row0, col0, slice0], [rowd, cold, sliced])."""
    from .parosol import _first_last_true
    row0, row1 = _first_last_true(max_row)
    col0, col1 = _first_last_true(max_col)
    slice0, slice1 = _first_last_true(max_slice)
    row0 = row0 - pad
    rowd = row1 - row0 + pad
    col0 = col0 - pad
    cold = col1 - col0 + pad
    slice0 = slice0 - pad
    sliced = slice1 - slice0 + pad
    if pad > 0:
        row0, col0, slice0 = max(row0, 0), max(col0, 0), max(slice0, 0)
        nz, ny, nx = (int(v) for v in shape)
        if slice0 + sliced > nz:
            sliced = nz - slice0
        if row0 + rowd > ny:
            rowd = ny - row0
        if col0 + cold > nx:
            cold = nx - col0
    return [row0, col0, slice0], [rowd, cold, sliced]


def _bbox_pad(sub, pad):
    """recon_utils.bbox(bw, pad) (recon_utils.py:258-298) of the device volume ``sub > 0``."""
    max_row, max_col, max_slice = _bool_projections(sub)
    return _bbox_from_projections(max_row, max_col, max_slice, sub.shape, pad)


def embed(I, embed_depth, embed_dir, embed_val=None, pad=0, makecopy=False):  # noqa: E741
    """Add embedding to 3D image (drop-in for preprocess.py:233-355).

    Direction and depth of the embedded region should be given; zeroes in the input image are background.
    The box of the embedding comes from the reference's projections and bounding box (with its padding
    rules), the embedding itself is the largest 6-connected part of ``box & ~(I > 0)``
    (``remove_unconnected``, i.e. the labelling kernels), which gets ``embed_val`` (default ``I.max() + 1``).
    As in the reference the input is modified in place and returned unless ``makecopy`` is true.

    Returns ``(I, BW_embedding)``; NumPy in, NumPy out; CUDA tensors stay on the device.
    """
    vol, is_bool, from_numpy = as_device_volume(I)
    Z, Y, X = (int(v) for v in vol.shape)
    dev = vol.device
    if embed_dir not in ("-z", "+z", "-x", "+x", "-y", "+y"):
        # the reference binarises the image first and fails on the direction afterwards: same exception
        raise IOError("EMBED_DIR parameter unknown. Valid entries are -x, +x, -y, +y, -z, and +z.")
    with torch.cuda.device(dev):
        s = _lib.stream()
        if embed_val is None:
            mx = torch.zeros(1, dtype=torch.int32, device=dev)
            _lib.call("cfe_max_u8", vol.data_ptr(), vol.numel(), mx.data_ptr(), s)
            top = int(read_back(mx)[0])
            with np.errstate(over="ignore"):
                embed_val = (np.bool_(top) if is_bool else np.uint8(top)) + 1      # I.max() + 1, NumPy's rules
        probe = np.zeros(1, dtype=np.bool_ if is_bool else np.uint8)
        probe[0] = embed_val                                   # the cast (or OverflowError) of I[mask] = embed_val
        val = int(probe[0])
        max_row, max_col, max_slice = _bool_projections(vol)

        def first_last(m, which):
            idx = np.where(m == True)[-1]  # noqa: E712 -- the reference's expression
            return idx[which]                                  # IndexError on an all-background image

        d = int(embed_depth)
        if embed_dir == "-z":
            start = int(first_last(max_slice, -1))
            sl = (slice(start - d, None), slice(None), slice(None))
        elif embed_dir == "+z":
            start = int(first_last(max_slice, 0))
            sl = (slice(None, start + d), slice(None), slice(None))
        elif embed_dir == "-x":
            start = int(first_last(max_col, -1))
            sl = (slice(None), slice(None), slice(start - d, None))
        elif embed_dir == "+x":
            start = int(first_last(max_col, 0))
            sl = (slice(None), slice(None), slice(None, start + d))
        elif embed_dir == "+y":
            start = int(first_last(max_row, -1))
            sl = (slice(None), slice(start - d, None), slice(None))
        else:
            start = int(first_last(max_row, 0))
            sl = (slice(None), slice(None, start + d), slice(None))
        sub = vol[sl].contiguous()
        o, size = _bbox_pad(sub, pad if embed_dir != "-y" else 0)     # the reference passes no pad for "-y"
        rows, cols, slices = (slice(o[0], o[0] + size[0]), slice(o[1], o[1] + size[1]),
                              slice(o[2], o[2] + size[2]))
        if embed_dir in ("-z", "+z"):
            box = (sl[0], rows, cols)
        elif embed_dir in ("-x", "+x"):
            box = (slices, rows, sl[2])
        else:
            box = (slices, sl[1], cols)
        (z0, z1, _), (y0, y1, _), (x0, x1, _) = (b.indices(n) for b, n in zip(box, (Z, Y, X)))
        cand = torch.empty((Z, Y, X), dtype=torch.uint8, device=dev)
        _lib.call("cfe_box_andnot_u8", vol.data_ptr(), Z, Y, X, z0, max(z1, z0), y0, max(y1, y0), x0, max(x1, x0),
                  cand.data_ptr(), s)
    mask, best = largest_component_device(cand)                # remove_unconnected(BW_embedding & ~BW_I)
    if best == 0:
        raise ValueError("attempt to get argmax of an empty sequence")
    in_place = isinstance(I, torch.Tensor) and I.is_cuda and I.is_contiguous() and not makecopy
    with torch.cuda.device(dev):
        out = vol if in_place else torch.empty_like(vol)
        _lib.call("cfe_masked_set_u8", vol.data_ptr(), mask.data_ptr(), vol.numel(), val, out.data_ptr(), _lib.stream())
    bw = mask.view(torch.bool)
    res = out.view(torch.bool) if is_bool else out
    if from_numpy:
        res_h = res.cpu().numpy()
        if makecopy:
            return res_h, bw.cpu().numpy()
        I[...] = res_h
        return I, bw.cpu().numpy()
    if makecopy or in_place:
        return (I if in_place else res), bw
    I.copy_(res)
    return I, bw


# --------------------------------------------------------------------------------------------
# periosteummask (preprocess.py:357-420)

def _remove_small_objects(vol, min_size):
    """skimage.morphology.remove_small_objects(bool image, min_size) (preprocess.py:385): the components
    (6-connectivity in 3D, 4 in 2D -- scikit-image's default connectivity=1) with fewer than ``min_size`` voxels
    are cleared.  The labelling kernels give every root its size; a per-root flag drives the mask writer.
    vol: uint8 0/1 CUDA tensor [Z,Y,X]; returns a new one."""
    lib = _lib.require_cuda()
    Z, Y, X = (int(v) for v in vol.shape)
    check_slab_size(Z, Y, X)
    dev = vol.device
    out = torch.zeros((Z, Y, X), dtype=torch.uint8, device=dev)
    if Z * Y * X == 0:
        return out
    with torch.cuda.device(dev):
        s = _lib.stream()
        W = (X + 31) // 32
        rows = Z * Y
        bits = empty_u32(rows * W, dev)
        _lib.call("cfe_pack_bits", vol.data_ptr(), rows, X, 0, bits.data_ptr(), s)
        pref = empty_u32(rows * W, dev)
        stats = torch.empty(2, dtype=torch.int64, device=dev)
        ws = scan_workspace(rows * W, dev)
        _lib.call("cfe_scan_run_starts", bits.data_ptr(), rows, X, pref.data_ptr(), stats.data_ptr(), ws.data_ptr(), s)
        max_runs = lib.cfe_ccl_max_runs(rows, X)
        parent = empty_u32(max_runs, dev)
        size = empty_u32(max_runs, dev)
        _lib.call("cfe_ccl_label", bits.data_ptr(), pref.data_ptr(), stats.data_ptr(), Z, Y, X,
                  parent.data_ptr(), size.data_ptr(), s)
        # keep[root] = 0xffffffff when the component is NOT too small (sizes are u32 in an int32 tensor)
        big = ~((size.to(torch.int64) & 0xFFFFFFFF) < min_size)
        keep = torch.where(big, torch.full_like(size, -1), torch.zeros_like(size))
        _lib.call("cfe_ccl_write_select", bits.data_ptr(), pref.data_ptr(), Z, Y, X, parent.data_ptr(), None,
                  keep.data_ptr(), 0, None, 1, out.data_ptr(), s)
    return out


def periosteummask(bwimage, closepixels=10, closevoxels=0, remove_objects_smaller_than=None, removeunconn=True,
                   verbose=False):
    """Binary mask of periosteum (whole bone) -- drop-in for preprocess.py:357-420.

    Every slice is closed with a disk of radius ``closepixels`` (``skimage.morphology.binary_closing``:
    scipy's binary dilation with border value 0, then erosion with border value 1) and its holes are filled
    (``scipy.ndimage.binary_fill_holes``: background not 4-connected to the slice border); isolated clusters
    are then removed in 3D (``remove_unconnected``).  All of it runs on packed bits on the GPU: the disk
    dilation is word-parallel, the hole filling is one labelling of the background of the whole stack
    (slices only meet through their border rows, which are outside by definition) plus a border-flag pass.

    ``remove_objects_smaller_than`` (scikit-image's ``remove_small_objects``) clears the small components first:
    one more labelling, sizes per root, a flag-driven mask writer.  ``closevoxels > 0``: the final 3D closing with
    ``cube(closevoxels)`` as six separable line passes on packed bits; ``closepixels <= 31``.
    """
    import logging
    if verbose:
        logging.basicConfig(level=logging.INFO)
    cw = int(closevoxels)
    if cw != closevoxels or cw > 33:
        raise NotImplementedError("periosteummask: closevoxels must be an integer <= 33")
    r = int(closepixels)
    if r != closepixels or not 0 <= r <= 31:
        raise NotImplementedError("periosteummask: closepixels must be an integer in 0..31")
    vol, is_bool, from_numpy = as_device_volume(bwimage, allow_2d=True)
    squeeze = (bwimage.ndim == 2)
    Z, Y, X = (int(v) for v in vol.shape)
    dev = vol.device
    lib = _lib.require_cuda()
    if remove_objects_smaller_than:
        if not is_bool:
            # scikit-image reads an integer image as a label image here: not the binary case this path handles
            raise TypeError("periosteummask: remove_objects_smaller_than needs a boolean image")
        logging.info('Preliminary removal of objects smaller than {} pixels.'.format(remove_objects_smaller_than))
        vol = _remove_small_objects(vol, remove_objects_smaller_than)
    if Z * Y * X == 0:
        out = torch.zeros((Z, Y, X), dtype=torch.bool, device=dev)
    else:
        with torch.cuda.device(dev):
            s = _lib.stream()
            W = (X + 31) // 32
            rows = Z * Y
            bits = empty_u32(rows * W, dev)
            _lib.call("cfe_pack_bits", vol.data_ptr(), rows, X, 0, bits.data_ptr(), s)
            dil = empty_u32(rows * W, dev)
            _lib.call("cfe_disk_dilate_bits", bits.data_ptr(), Z, Y, X, r, 0, dil.data_ptr(), s)
            # background of the closed slices = dilation of the complement of the dilation
            bg = bits
            _lib.call("cfe_disk_dilate_bits", dil.data_ptr(), Z, Y, X, r, 1, bg.data_ptr(), s)
            # fill_holes: label the background of the stack as one tall slice (4-connectivity inside the
            # slices; consecutive slices only touch through their border rows) and flag what meets a border
            pref = empty_u32(rows * W, dev)
            stats = torch.empty(2, dtype=torch.int64, device=dev)
            ws = scan_workspace(rows * W, dev)
            _lib.call("cfe_scan_run_starts", bg.data_ptr(), rows, X, pref.data_ptr(), stats.data_ptr(), ws.data_ptr(), s)
            max_runs = lib.cfe_ccl_max_runs(rows, X)
            parent = empty_u32(max_runs, dev)
            size = empty_u32(max_runs, dev)
            _lib.call("cfe_ccl_label", bg.data_ptr(), pref.data_ptr(), stats.data_ptr(), 1, rows, X,
                      parent.data_ptr(), size.data_ptr(), s)
            flag = torch.full((max_runs,), -1, dtype=torch.int32, device=dev)        # 0xffffffff = not at a border
            _lib.call("cfe_border_flags", bg.data_ptr(), pref.data_ptr(), Z, Y, X, parent.data_ptr(), flag.data_ptr(), s)
            # mask = 1 everywhere except the background runs that reach a border (flag == 0)
            mask = torch.ones((Z, Y, X), dtype=torch.uint8, device=dev)
            _lib.call("cfe_ccl_write_select", bg.data_ptr(), pref.data_ptr(), 1, rows, X, parent.data_ptr(), None,
                      flag.data_ptr(), 1, mask.data_ptr(), 0, mask.data_ptr(), s)
        if removeunconn:
            logging.info("Removing isolated clusters of voxels.")
            kept, best = largest_component_device(mask)
            if best == 0:
                raise ValueError("attempt to get argmax of an empty sequence")      # remove_unconnected, :141
            mask = kept
        if cw > 0 and not squeeze:
            # final 3D closing with cube(closevoxels) (preprocess.py:405-408; the 2D branch has none)
            logging.info('Final 3D image closing.\n Structuring element CUBE of radius: {}'.format(closepixels))
            mask = _cube_closing(mask, cw)
        out = mask.view(torch.bool)
    if squeeze:
        out = out[0]
    return out.cpu().numpy() if from_numpy else out


def _cube_closing(mask, w):
    """``skimage.morphology.binary_closing(mask, cube(w))`` of a uint8 0/1 CUDA volume: scipy's dilation
    (border 0) then erosion (border 1) by ones((w, w, w)), each as three line passes on packed bits; an
    even-sized structure is centred at w // 2, so the offsets of the two are mirrored (csrc/morph.cu)."""
    Z, Y, X = (int(v) for v in mask.shape)
    dev = mask.device
    L, R = w // 2, w - 1 - w // 2
    with torch.cuda.device(dev):
        s = _lib.stream()
        W = (X + 31) // 32
        a = empty_u32(Z * Y * W, dev)
        b = empty_u32(Z * Y * W, dev)
        _lib.call("cfe_pack_bits", mask.data_ptr(), Z * Y, X, 0, a.data_ptr(), s)
        for axis in range(3):                                   # dilation: sources q - R .. q + L
            _lib.call("cfe_line_or_bits", a.data_ptr(), Z, Y, X, axis, R, L, 0, 0, b.data_ptr(), s)
            a, b = b, a
        for axis in range(3):                                   # erosion = ~ dilation'(~x): sources q - L .. q + R
            _lib.call("cfe_line_or_bits", a.data_ptr(), Z, Y, X, axis, L, R, 1 if axis == 0 else 0,
                      1 if axis == 2 else 0, b.data_ptr(), s)
            a, b = b, a
        out = torch.empty((Z, Y, X), dtype=torch.uint8, device=dev)
        _lib.call("cfe_unpack_bits", a.data_ptr(), Z * Y, X, out.data_ptr(), s)
    return out
