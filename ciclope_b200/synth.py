"""Synthetic trabecular-like microCT volumes, bit-reproducible on CPU and GPU (integer-only
torch ops): splitmix64 noise of the voxel index -> three passes of separable box blur (radius r,
exact integer arithmetic) -> threshold at the level giving the requested bone volume fraction."""
import torch

_M64 = (1 << 64) - 1


def _i64(v):
    v &= _M64
    return v - (1 << 64) if v >= (1 << 63) else v


def _lsr(x, k):
    return (x >> k) & ((1 << (64 - k)) - 1)


def noise16(shape, seed, device, z0=0, full_shape=None):
    """uint16-range white noise as int32; value depends only on (seed, global voxel index)."""
    Z, Y, X = shape
    fz, fy, fx = full_shape or shape
    z = torch.arange(z0, z0 + Z, device=device, dtype=torch.int64).view(Z, 1, 1)
    y = torch.arange(Y, device=device, dtype=torch.int64).view(1, Y, 1)
    x = torch.arange(X, device=device, dtype=torch.int64).view(1, 1, X)
    idx = (z * fy + y) * fx + x
    v = idx + _i64(seed * 0x9E3779B97F4A7C15 + 0x632BE59BD9B4E019)
    v = (v ^ _lsr(v, 30)) * _i64(0xBF58476D1CE4E5B9)
    v = (v ^ _lsr(v, 27)) * _i64(0x94D049BB133111EB)
    v = v ^ _lsr(v, 31)
    return (_lsr(v, 48)).to(torch.int32)


def _box(f, axis, r):
    """periodic box filter of width 2r+1 along `axis`, floor-divided to stay in 16 bits"""
    n = f.shape[axis]
    idx = torch.arange(-r - 1, n + r, device=f.device) % n
    c = torch.cumsum(f.index_select(axis, idx).to(torch.int64), dim=axis)
    hi = c.narrow(axis, 2 * r + 1, n)
    lo = c.narrow(axis, 0, n)
    return ((hi - lo) // (2 * r + 1)).to(torch.int32)


def field(shape, seed, device, radius=4, passes=3):
    f = noise16(shape, seed, device)
    for _ in range(passes):
        for axis in range(3):
            f = _box(f, axis, radius)
    return f


def level_for_fraction(f, fraction):
    """smallest integer t with mean(f > t) <= fraction"""
    h = torch.bincount(f.reshape(-1).to(torch.int64), minlength=65536)
    above = h.flip(0).cumsum(0).flip(0)          # above[t] = #(f >= t)
    n = f.numel()
    ok = (above[1:] <= int(fraction * n)).nonzero()   # #(f > t) = above[t+1]
    return int(ok[0].item())


def trabecular(shape, seed=0, bvtv=0.30, device="cuda", radius=4, passes=3):
    """bool volume [Z,Y,X] with bone volume fraction ~bvtv (smooth, mostly one connected component)."""
    f = field(shape, seed, device, radius, passes)
    t = level_for_fraction(f, bvtv)
    return f > t


def foam(shape, seed=0, bvtv=0.12, device="cuda", radius=4, passes=3):
    """Fragmented foam (BASELINE.json configs[3]): the same field thresholded BELOW the percolation level,
    so that the foreground falls apart into many components of similar size (232 at 192^3 with the defaults,
    i.e. ~3 * 10^5 at 2048^3; the largest holds ~20 % of the foreground) -- the case that stresses the
    cross-slab merge of the sharded labelling."""
    return trabecular(shape, seed, bvtv, device, radius, passes)


def greyscale(shape, seed=0, bvtv=0.30, device="cuda", radius=4, passes=3):
    """uint8 volume: GV = 1 + blurred field quantised to 0..254 inside the bone mask, 0 outside
    (the `masked[~L] = 0` pattern of ciclope example 01)."""
    f = field(shape, seed, device, radius, passes)
    t = level_for_fraction(f, bvtv)
    fmax = int(f.max().item())
    g = 1 + ((f - t).clamp(min=0).to(torch.int64) * 254 // max(fmax - t, 1)).clamp(max=254)
    return torch.where(f > t, g.to(torch.uint8), torch.zeros((), dtype=torch.uint8, device=f.device))


def field_slab(global_shape, z0, z1, seed=0, device="cuda", radius=4, passes=3):
    """Slab [z0, z1) of the GLOBAL blurred field of shape `global_shape`, generated without ever holding
    the global field: the noise is a pure function of the global voxel index and the z-blur only needs a
    halo of radius*passes slices (periodic in the global z)."""
    Zg, Y, X = global_shape
    h = radius * passes
    if Zg < z1 - z0 + 2 * h:
        return field(global_shape, seed, device, radius, passes)[z0:z1].contiguous()
    zs = (torch.arange(z0 - h, z1 + h, device=device, dtype=torch.int64) % Zg).view(-1, 1, 1)
    y = torch.arange(Y, device=device, dtype=torch.int64).view(1, Y, 1)
    x = torch.arange(X, device=device, dtype=torch.int64).view(1, 1, X)
    idx = (zs * Y + y) * X + x
    v = idx + _i64(seed * 0x9E3779B97F4A7C15 + 0x632BE59BD9B4E019)
    v = (v ^ _lsr(v, 30)) * _i64(0xBF58476D1CE4E5B9)
    v = (v ^ _lsr(v, 27)) * _i64(0x94D049BB133111EB)
    v = v ^ _lsr(v, 31)
    f = _lsr(v, 48).to(torch.int32)
    del v, idx
    for _ in range(passes):
        for axis in range(3):
            f = _box(f, axis, radius)
    return f[h: h + (z1 - z0)].contiguous()


def _global_level(f, bvtv, global_shape, z0, z1, group):
    """(threshold level, global maximum of the field) from the all-reduced histogram of the slabs"""
    import torch.distributed as dist
    Zg, Y, X = global_shape
    hist = torch.bincount(f.reshape(-1).to(torch.int64), minlength=65536)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(hist, group=group)
    elif (z0, z1) != (0, Zg):
        raise RuntimeError("a partial volume needs an initialised process group to find the global threshold")
    n = Zg * Y * X
    above = hist.flip(0).cumsum(0).flip(0)
    ok = (above[1:] <= int(bvtv * n)).nonzero()
    return int(ok[0].item()), int(hist.nonzero().max().item())


def trabecular_slab(global_shape, z0, z1, seed=0, bvtv=0.30, device="cuda", radius=4, passes=3, group=None):
    """Slab [z0, z1) of ``trabecular(global_shape, ...)``: every rank holds exactly the slices the global
    generator would give (the threshold comes from the all-reduced histogram)."""
    f = field_slab(global_shape, z0, z1, seed, device, radius, passes)
    t, _ = _global_level(f, bvtv, global_shape, z0, z1, group)
    return f > t


def foam_slab(global_shape, z0, z1, seed=0, bvtv=0.12, device="cuda", radius=4, passes=3, group=None):
    """Slab [z0, z1) of ``foam(global_shape, ...)`` (BASELINE.json configs[3])."""
    return trabecular_slab(global_shape, z0, z1, seed, bvtv, device, radius, passes, group)


def greyscale_slab(global_shape, z0, z1, seed=0, bvtv=0.30, device="cuda", radius=4, passes=3, group=None):
    """Slab [z0, z1) of ``greyscale(global_shape, ...)`` (BASELINE.json configs[2])."""
    f = field_slab(global_shape, z0, z1, seed, device, radius, passes)
    t, fmax = _global_level(f, bvtv, global_shape, z0, z1, group)
    g = 1 + ((f - t).clamp(min=0).to(torch.int64) * 254 // max(fmax - t, 1)).clamp(max=254)
    return torch.where(f > t, g.to(torch.uint8), torch.zeros((), dtype=torch.uint8, device=f.device))
