"""Seeded input volumes and parameter sets shared by the golden generator, the oracle tests and
the GPU parity tests.  Pure NumPy, no reference / oracle / product imports."""
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TEMPLATE = os.path.join(GOLDEN_DIR, "tmpl_compression.inp")
TEMPLATE_REFNODE = os.path.join(GOLDEN_DIR, "tmpl_refnode.inp")
MAT_LINEAR = os.path.join(GOLDEN_DIR, "mat_linear.inp")
MAT_CONST = os.path.join(GOLDEN_DIR, "mat_const.inp")
MAT_POWER = os.path.join(GOLDEN_DIR, "mat_power.inp")


def trab():
    """The reference's unit-test fixture test/trab/slice_00..09.tif (10x10x10 uint8)."""
    return np.load(os.path.join(GOLDEN_DIR, "trab_10x10x10_u8.npy"))


def rand_binary(shape, density, seed, dtype=np.bool_, one=1):
    rng = np.random.default_rng(seed)
    m = rng.random(shape) < density
    if dtype == np.bool_:
        return m
    return (m.astype(np.uint8) * np.uint8(one)).astype(dtype)


def rand_grey(shape, density, seed, lo=1, hi=255):
    rng = np.random.default_rng(seed)
    m = rng.random(shape) < density
    g = rng.integers(lo, hi + 1, size=shape, dtype=np.int64).astype(np.uint8)
    return np.where(m, g, np.uint8(0)).astype(np.uint8)


def blobs(shape, density, seed, radius=2, passes=2):
    """Smooth trabecular-like binary structure: box-blurred integer noise thresholded at the
    requested volume fraction (integer arithmetic only)."""
    rng = np.random.default_rng(seed)
    f = rng.integers(0, 1 << 16, size=shape, dtype=np.int64)
    for _ in range(passes):
        for ax in range(3):
            n = f.shape[ax]
            pad = [(0, 0)] * 3
            pad[ax] = (radius + 1, radius)
            c = np.cumsum(np.pad(f, pad, mode="wrap"), axis=ax)
            hi = [slice(None)] * 3
            lo = [slice(None)] * 3
            hi[ax] = slice(2 * radius + 1, 2 * radius + 1 + n)
            lo[ax] = slice(0, n)
            f = c[tuple(hi)] - c[tuple(lo)]
    flat = np.sort(f.ravel())
    thr = flat[int((1.0 - density) * (flat.size - 1))]
    return f > thr


def grey_from_mask(mask, seed):
    """Greyscale uint8 volume, 0 outside the mask (the `masked[~L] = 0` pattern of ex01)."""
    rng = np.random.default_rng(seed)
    g = rng.integers(1, 256, size=mask.shape, dtype=np.int64)
    # smooth a little so neighbouring voxels share grey values
    g = (g + np.roll(g, 1, 0) + np.roll(g, 1, 1) + np.roll(g, 1, 2)) // 4
    g = np.clip(g, 1, 255).astype(np.uint8)
    return np.where(mask, g, np.uint8(0)).astype(np.uint8)


# ------------------------------------------------------------------------------------ CCL cases

def ccl_cases():
    """(name, bool volume) -- adversarial shapes for remove_unconnected."""
    out = []
    out.append(("trab_gt100", trab() > 100))
    out.append(("single_voxel", _single((3, 4, 5), (1, 2, 3))))
    out.append(("full", np.ones((4, 5, 6), dtype=bool)))
    # two equal-size components: tie -> the one met first in raster order
    t = np.zeros((3, 5, 9), dtype=bool)
    t[0, 0, 6:9] = True
    t[2, 4, 0:3] = True
    out.append(("tie_two_equal", t))
    t2 = np.zeros((2, 6, 40), dtype=bool)
    t2[1, 5, 0:7] = True      # later in raster order
    t2[0, 2, 30:37] = True    # earlier
    t2[1, 0, 10:16] = True    # smaller
    out.append(("tie_far_apart", t2))
    # U shape: prongs are met before the junction
    u = np.zeros((2, 8, 40), dtype=bool)
    u[0, 0:7, 3] = True
    u[0, 0:7, 36] = True
    u[0, 6, 3:37] = True
    u[1, 1, 10:14] = True
    out.append(("u_shape", u))
    # diagonal neighbours must NOT connect (6-connectivity)
    d = np.zeros((4, 4, 4), dtype=bool)
    for i in range(4):
        d[i, i, i] = True
    d[0, 3, 0:2] = True
    out.append(("diagonal_not_connected", d))
    # spiral / snake that forces long union chains
    s = np.zeros((1, 21, 45), dtype=bool)
    for r in range(0, 21, 2):
        s[0, r, :] = True
        if (r // 2) % 2 == 0 and r + 1 < 21:
            s[0, r + 1, 44] = True
        elif r + 1 < 21:
            s[0, r + 1, 0] = True
    out.append(("snake", s))
    # z-direction connectivity only
    zc = np.zeros((9, 3, 35), dtype=bool)
    zc[:, 1, 33] = True
    zc[0, 0, 0:5] = True
    out.append(("z_column", zc))
    # runs crossing 32-bit word boundaries, X not a multiple of 32
    w = np.zeros((3, 3, 70), dtype=bool)
    w[1, 1, 20:69] = True
    w[0, 1, 31:33] = True
    w[2, 1, 63:65] = True
    w[0, 0, 0:3] = True
    out.append(("word_crossing", w))
    for i, (shape, dens) in enumerate([((1, 1, 1), 1.0), ((1, 1, 37), 0.6), ((1, 9, 1), 0.7), ((7, 1, 1), 0.8),
                                       ((5, 7, 9), 0.35), ((6, 6, 33), 0.5), ((4, 5, 64), 0.45),
                                       ((3, 4, 100), 0.55), ((9, 8, 31), 0.3), ((8, 16, 32), 0.4),
                                       ((12, 13, 47), 0.25), ((16, 16, 96), 0.32)]):
        out.append(("rand_%dx%dx%d" % shape, rand_binary(shape, dens, 100 + i)))
    out.append(("blobs_20x24x70", blobs((20, 24, 70), 0.3, 7)))
    out.append(("blobs_sparse_24x24x64", blobs((24, 24, 64), 0.12, 8, radius=1, passes=1)))
    return out


def _single(shape, pos):
    v = np.zeros(shape, dtype=bool)
    v[pos] = True
    return v


# ------------------------------------------------------------------------------------ deck cases

def deck_cases():
    """List of dicts: name, vol, u (vol2ugrid kwargs), m (mesh2voxelfe kwargs).
    `m["matprop"]` is a factory (fresh dict per call: the reference mutates it, voxelFE.py:585)."""
    C = []

    def add(name, vol, u=None, m=None):
        u = dict(u or {})
        u.setdefault("voxelsize", [0.0195, 0.0195, 0.0195])
        m = dict(m or {})
        m.setdefault("templatefile", TEMPLATE)
        C.append(dict(name=name, vol=vol, u=u, m=m))

    tr = trab()
    add("trab_gvmin100_plane", tr, dict(voxelsize=[0.12, 0.12, 0.12], GVmin=100))
    add("trab_gvmin100_exact", tr, dict(voxelsize=[0.12, 0.12, 0.12], GVmin=100, locking_strategy="exact",
                                        node_level_lock_num=1))
    add("trab_grey_property", tr, dict(voxelsize=[0.12, 0.12, 0.12]),
        dict(keywords=['NSET', 'ELSET', 'PROPERTY'], matprop=lambda: {"file": [MAT_LINEAR], "range": [[1, 255]]}))
    add("trab_grey_property_norange", tr, dict(voxelsize=[0.12, 0.12, 0.12], GVmin=30),
        dict(keywords=['NSET', 'ELSET', 'PROPERTY'], matprop=lambda: {"file": [MAT_LINEAR]}))
    add("trab_two_matfiles", tr, dict(voxelsize=0.12, GVmin=10),
        dict(keywords=['NSET', 'PROPERTY'],
             matprop=lambda: {"file": [MAT_CONST, MAT_POWER], "range": [[1, 120], [121, 255]]}))
    add("trab_bool", tr > 100, dict(voxelsize=[0.12]))
    add("trab_bool_property", tr > 100, dict(voxelsize=[0.12]),
        dict(keywords=['NSET', 'ELSET', 'PROPERTY'], matprop=lambda: {"file": [MAT_LINEAR]}))

    b = rand_binary((5, 7, 9), 0.4, 1)
    add("bool_5x7x9", b)
    add("u8_01_5x7x9", b.astype(np.uint8))
    add("u8_0_255_5x7x9", b.astype(np.uint8) * np.uint8(255))
    add("single_voxel", _single((3, 4, 5), (1, 2, 3)))
    add("full_4x5x6", np.ones((4, 5, 6), dtype=bool))
    add("z1_1x6x35", rand_binary((1, 6, 35), 0.5, 2))
    add("y1_6x1x35", rand_binary((6, 1, 35), 0.5, 3))
    add("x1_6x7x1", rand_binary((6, 7, 1), 0.5, 4))
    add("one_1x1x1", np.ones((1, 1, 1), dtype=bool))
    add("x33_4x5x33", rand_binary((4, 5, 33), 0.45, 5))
    add("x64_3x4x64", rand_binary((3, 4, 64), 0.5, 6))
    add("x100_3x3x100", rand_binary((3, 3, 100), 0.3, 7))
    add("x31_7x9x31", rand_binary((7, 9, 31), 0.2, 8))

    # plane locking variants (Z0/Z1 overlap, p > extent, p = 0)
    v = rand_binary((8, 9, 12), 0.45, 9)
    add("plane_p2", v, dict(plane_lock_num=2))
    add("plane_p4_overlap", v, dict(plane_lock_num=4))
    add("plane_p5_overlap", v, dict(plane_lock_num=5))
    add("plane_p20_gt_extent", v, dict(plane_lock_num=20))
    add("plane_p0", v, dict(plane_lock_num=0))
    add("exact_k1", v, dict(locking_strategy="exact", node_level_lock_num=1))
    add("exact_k2", v, dict(locking_strategy="exact", node_level_lock_num=2))
    add("exact_k0", v, dict(locking_strategy="exact", node_level_lock_num=0))
    add("exact_k30", v, dict(locking_strategy="exact", node_level_lock_num=30))

    # set sizes that hit the 10-per-line wrap rule: z-plane 0 holds exactly k voxels in a row
    for k in (9, 10, 11, 20):
        w = np.zeros((3, 4, 24), dtype=bool)
        w[0, 0, 0:k] = True
        w[1, 2, 5] = True
        w[2, 3, 1:4] = True
        add("wrap_%d" % k, w)
    e = np.zeros((3, 3, 3), dtype=bool)
    e[1, 1, 1] = True
    add("empty_sets_centre_voxel", e)

    # keyword subsets / element type / refnode
    add("kw_none", b, m=dict(keywords=[]))
    add("kw_nset", b, m=dict(keywords=['NSET']))
    add("kw_elset", b, m=dict(keywords=['ELSET']))
    add("eltype_c3d8r", b, m=dict(eltype='C3D8R'))
    add("refnode", b, m=dict(templatefile=TEMPLATE_REFNODE, refnode=[0.08776, 0.06, 0.123456]))

    # voxelsize flavours (voxelFE.py:70-73, 202; dtype of np.asarray(nodes))
    add("vs_scalar", b, dict(voxelsize=0.5))
    add("vs_len1", b, dict(voxelsize=[0.25]))
    add("vs_ndarray", b, dict(voxelsize=np.array([0.01, 0.02, 0.03])))
    add("vs_int", b, dict(voxelsize=[1, 2, 3]))
    add("vs_mixed_int_float", b, dict(voxelsize=[1, 0.5, 2]))
    add("vs_float32", b, dict(voxelsize=np.array([0.1, 0.2, 0.3], dtype=np.float32)))
    add("vs_negative", b, dict(voxelsize=[-0.1, 0.1, -0.25]))
    add("vs_rounding", b, dict(voxelsize=[0.0000005, 0.0000015, 0.0000025]))
    add("vs_big_but_12chars", b, dict(voxelsize=[11111.11, 1234.5, 9999.999]))
    # '{:12.6f}' overflows its 12 characters from 1e5 on: variable-length *NODE lines
    add("vs_wide_fields", b, dict(voxelsize=[30000.5, 1e6, 123456.789]))
    add("vs_wide_negative_mixed", b, dict(voxelsize=[-25000.25, 0.0195, 5e9]))
    add("vs_wide_int", b, dict(voxelsize=[100000, 7, 12345678]))

    # greyscale with several ELSETs, GVmin, float GVmin
    g = rand_grey((6, 7, 20), 0.5, 10)
    add("grey_6x7x20", g)
    add("grey_gvmin100", g, dict(GVmin=100))
    add("grey_gvmin_float", g, dict(GVmin=99.5))
    add("grey_gvmin_neg", g[:3, :3, :5], dict(GVmin=-1))
    add("grey_few_values", (rand_grey((5, 6, 40), 0.6, 11, 1, 3) * np.uint8(60)).astype(np.uint8),
        m=dict(keywords=['NSET', 'ELSET', 'PROPERTY'], matprop=lambda: {"file": [MAT_POWER], "range": [[1, 255]]}))
    m3 = blobs((14, 16, 40), 0.3, 12)
    add("blobs_grey_property", grey_from_mask(m3, 13),
        m=dict(keywords=['NSET', 'ELSET', 'PROPERTY'], matprop=lambda: {"file": [MAT_LINEAR], "range": [[1, 255]]}))
    add("blobs_bool", m3)
    return C


def midsize_cases():
    """Larger cases whose golden record is a SHA-256 only (the reference needs ~10 s for them)."""
    C = []
    m = blobs((40, 48, 70), 0.3, 21)
    C.append(dict(name="mid_blobs_bool_40x48x70", vol=m, u=dict(voxelsize=[0.0195] * 3),
                  m=dict(templatefile=TEMPLATE)))
    C.append(dict(name="mid_blobs_grey_40x48x70", vol=grey_from_mask(m, 22), u=dict(voxelsize=[0.0195] * 3),
                  m=dict(templatefile=TEMPLATE, keywords=['NSET', 'ELSET', 'PROPERTY'],
                         matprop=lambda: {"file": [MAT_LINEAR], "range": [[1, 255]]})))
    return C


def mask_timestamp(deck_bytes):
    """Replace the non-deterministic header line 2 (voxelFE.py:619) so decks can be compared."""
    lines = deck_bytes.split(b"\n", 3)
    if len(lines) >= 3 and lines[1].startswith(b"** Abaqus .INP file written by Voxel_FEM script on "):
        lines[1] = b"** Abaqus .INP file written by Voxel_FEM script on <timestamp>"
    return b"\n".join(lines)
