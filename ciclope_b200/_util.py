"""Host-side helpers shared by the facade modules (device volumes, scratch buffers)."""
import numpy as np
import torch

from . import _lib


def as_device_volume(vol, allow_2d=False):
    """Return (uint8 CUDA tensor [Z,Y,X] contiguous, is_bool, came_from_numpy).

    Accepts numpy arrays and torch tensors (CPU or CUDA) of dtype bool or uint8.  Other dtypes
    raise TypeError: the CUDA path computes on 8-bit grey values and there is no CPU fallback."""
    _lib.require_cuda()
    from_numpy = False
    if isinstance(vol, np.ndarray):
        from_numpy = True
        if vol.dtype not in (np.bool_, np.uint8):
            raise TypeError("ciclope_b200 handles bool / uint8 volumes, got %s" % vol.dtype)
        a = np.ascontiguousarray(vol)
        is_bool = a.dtype == np.bool_
        t = torch.from_numpy(a.view(np.uint8))
    elif isinstance(vol, torch.Tensor):
        if vol.dtype not in (torch.bool, torch.uint8):
            raise TypeError("ciclope_b200 handles bool / uint8 volumes, got %s" % vol.dtype)
        is_bool = vol.dtype == torch.bool
        t = vol.contiguous()
        if is_bool:
            t = t.view(torch.uint8)
    else:
        raise TypeError("volume must be a numpy.ndarray or torch.Tensor, got %s" % type(vol).__name__)
    if t.dim() == 2 and allow_2d:
        t = t.unsqueeze(0)
    if t.dim() != 3:
        raise ValueError("volume must be 3-D [slices, rows, cols], got %d-D" % t.dim())
    if not t.is_cuda:
        t = t.cuda(non_blocking=True)
    return t, is_bool, from_numpy


def empty_u32(n, device):
    return torch.empty(max(int(n), 1), dtype=torch.int32, device=device)


def scan_workspace(n_items, device):
    nbytes = _lib.load().cfe_scan_workspace_bytes(int(n_items))
    return torch.empty((nbytes + 7) // 8, dtype=torch.int64, device=device)


def check_slab_size(Z, Y, X):
    if (Z + 1) * (Y + 1) * (X + 1) >= 2 ** 32:
        raise ValueError("a single-GPU slab must hold fewer than 2^32 grid nodes; shard the volume as "
                         "z-slabs over more GPUs")


def to_device_async(data, dtype, device):
    """Small host data -> device tensor without a blocking copy: staged in pinned memory (torch's
    caching host allocator) and copied with non_blocking=True on the current stream.
    torch.tensor(..., device=cuda) would block the host until everything queued before it is done."""
    if isinstance(data, torch.Tensor):
        host = data.to(dtype)
    else:
        host = torch.tensor(data, dtype=dtype)
    return host.pin_memory().to(device, non_blocking=True)


def bytes_to_device_async(raw, device):
    host = torch.frombuffer(bytearray(raw) if len(raw) else bytearray(b"\0"), dtype=torch.uint8)
    return host.pin_memory().to(device, non_blocking=True)


_READBACK = {}


def read_back(t):
    """Small device tensor -> Python list through a persistent pinned buffer: one asynchronous copy
    and a stream synchronisation (``.cpu()`` goes through a pageable staging copy)."""
    key = (t.device.index, t.dtype)
    host = _READBACK.get(key)
    if host is None or host.numel() < t.numel():
        host = torch.empty(max(64, t.numel()), dtype=t.dtype).pin_memory()
        _READBACK[key] = host
    view = host[: t.numel()]
    view.copy_(t.reshape(-1), non_blocking=True)
    torch.cuda.current_stream(t.device).synchronize()
    return view.tolist()


def tag_packed(t, bits):
    """Remember that `bits` (u32 words, the layout of cfe_pack_bits) holds ``t != 0`` bit-packed: the mask
    writer of remove_unconnected produces them for free, and vol2ugrid of that mask starts from them.  The
    tag lives on the tensor OBJECT and carries its version counter, so an in-place edit of the mask (or of
    any view of it) invalidates it."""
    t._cfe_bits = (bits, t._version, tuple(t.shape))
    return t


def packed_of(t):
    """The bit-packed ``t != 0`` left by :func:`tag_packed`, or None when absent / stale."""
    tag = getattr(t, "_cfe_bits", None) if isinstance(t, torch.Tensor) else None
    if tag is None:
        return None
    bits, version, shape = tag
    if t._version != version or tuple(t.shape) != shape or not t.is_contiguous() or bits.device != t.device:
        return None
    return bits
