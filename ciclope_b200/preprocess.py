"""B200-native drop-in for ``ciclope.utils.preprocess.remove_unconnected``
(reference: src/ciclope/utils/preprocess.py:119-143)."""
import numpy as np
import torch

from . import _lib
from ._util import as_device_volume, check_slab_size, empty_u32, scan_workspace


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
        Binary image (bool; uint8 is accepted and read as ``!= 0``).

    Returns
    -------
    bwcluster
        Binary image of the largest connected cluster of voxels.
    """
    vol, _is_bool, from_numpy = as_device_volume(bwimage, allow_2d=True)
    squeeze = (bwimage.ndim == 2)
    mask, best = largest_component_device(vol)
    if best == 0:
        # reference: occurrences[1:].argmax() on an empty array (preprocess.py:141)
        raise ValueError("attempt to get argmax of an empty sequence")
    out = mask.view(torch.bool)
    if squeeze:
        out = out[0]
    if from_numpy:
        return out.cpu().numpy()
    return out


def largest_component_device(vol):
    """vol: uint8 CUDA tensor [Z,Y,X].  Returns (uint8 mask tensor, best) where
    best = (size << 32) | (0xffffffff - root run id), 0 when there is no foreground.
    One device->host read (16 bytes) at the end."""
    lib = _lib.require_cuda()
    Z, Y, X = (int(v) for v in vol.shape)
    check_slab_size(Z, Y, X)
    dev = vol.device
    out = torch.empty((Z, Y, X), dtype=torch.uint8, device=dev)
    if Z * Y * X == 0:
        return out, 0
    with torch.cuda.device(dev):
        s = _lib.stream()
        W = (X + 31) // 32
        rows = Z * Y
        bits = empty_u32(rows * W, dev)
        pref = empty_u32(rows * W, dev)
        stats = torch.zeros(2, dtype=torch.int64, device=dev)  # [n_runs, best]
        ws = scan_workspace(rows * W, dev)
        _lib.call("cfe_pack_bits", vol.data_ptr(), rows, X, 0, bits.data_ptr(), s)
        _lib.call("cfe_scan_run_starts", bits.data_ptr(), rows, X, pref.data_ptr(), stats.data_ptr(),
                  ws.data_ptr(), s)
        max_runs = lib.cfe_ccl_max_runs(rows, X)
        parent = empty_u32(max_runs, dev)
        size = empty_u32(max_runs, dev)
        _lib.call("cfe_ccl_largest", bits.data_ptr(), pref.data_ptr(), stats.data_ptr(), Z, Y, X,
                  parent.data_ptr(), size.data_ptr(), stats.data_ptr() + 8, out.data_ptr(), s)
        host = stats.cpu()
    best = int(host[1].item()) & 0xFFFFFFFFFFFFFFFF
    return out, best


def component_stats(best):
    """(size, root run id) of the kept component from the packed `best` word."""
    return best >> 32, 0xFFFFFFFF - (best & 0xFFFFFFFF)
