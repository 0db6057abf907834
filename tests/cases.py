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
    # GVmin < 0 on a BOOL volume: the False voxels are elements too (two blocks, SET0 and SET1)
    add("bool_gvmin_neg", b[:3, :4, :6], dict(GVmin=-1))
    add("bool_gvmin_neg_property", b[:3, :4, :6], dict(GVmin=-0.5),
        m=dict(keywords=['NSET', 'ELSET', 'PROPERTY'], matprop=lambda: {"file": [MAT_CONST], "range": [[0, 1]]}))
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


def real_lhdl():
    """The LHDL trabecular bone stack of the reference checkout thresholded at 63 (tests/golden/
    make_golden_real.py): (bw = image > 63, kept = remove_unconnected(bw) of the reference, masked grey
    volume = image with everything outside `kept` set to 0), all [200,200,200]."""
    g = np.load(os.path.join(GOLDEN_DIR, "real_lhdl.npz"))
    shape = tuple(int(v) for v in g["shape"])
    n = int(np.prod(shape))
    bw = np.unpackbits(g["bw"])[:n].astype(bool).reshape(shape)
    kept = np.unpackbits(g["kept"])[:n].astype(bool).reshape(shape)
    gv = np.zeros(shape, dtype=np.uint8)
    gv[kept] = g["kept_gv"]
    return bw, kept, gv


class SectionStreamHasher:
    """write()-able sink for ``mesh2voxelfe(mesh, template, sink)``: cuts the streamed deck into the sections
    of ciclope_b200.sections.split_deck from their expected byte counts (the header ends with the first
    ``*NODE`` line; `lengths` = bytes of nodes, elements, nset, elset in file order, absent sections left
    out; the tail is the rest) and hashes each one, so that a multi-GB GPU deck is compared with the
    streamed oracle without ever being held on the host."""

    def __init__(self, lengths):
        import hashlib
        self.names = [k for k in ("nodes", "elements", "nset", "elset") if k in lengths] + ["tail"]
        self.left = [int(lengths[k]["bytes"] if isinstance(lengths[k], dict) else lengths[k])
                     for k in self.names[:-1]] + [None]
        self.h = [hashlib.sha256() for _ in self.names]
        self.count = [0] * len(self.names)
        self.cur = -1            # -1: still inside the header
        self.head = b""

    def write(self, mv):
        data = bytes(mv)
        if self.cur < 0:
            self.head += data
            i = self.head.find(b"*NODE\n")
            if i < 0:
                return
            data, self.head, self.cur = self.head[i + 6:], self.head[:i + 6], 0
        pos = 0
        while pos < len(data):
            left = self.left[self.cur]
            take = len(data) - pos if left is None else min(left, len(data) - pos)
            self.h[self.cur].update(data[pos:pos + take])
            self.count[self.cur] += take
            pos += take
            if left is not None:
                self.left[self.cur] -= take
                if self.left[self.cur] == 0:
                    self.cur += 1

    def result(self):
        return {k: {"bytes": self.count[i], "sha256": self.h[i].hexdigest()} for i, k in enumerate(self.names)}


def mask_timestamp(deck_bytes):
    """Replace the non-deterministic header line 2 (voxelFE.py:619) so decks can be compared."""
    lines = deck_bytes.split(b"\n", 3)
    if len(lines) >= 3 and lines[1].startswith(b"** Abaqus .INP file written by Voxel_FEM script on "):
        lines[1] = b"** Abaqus .INP file written by Voxel_FEM script on <timestamp>"
    return b"\n".join(lines)


# ------------------------------------------------------------------------------------ generic meshes

def _grid_mesh(vol, voxelsize=(0.5, 0.25, 2.0), thr=0):
    """points / cells / gv of vol2ugrid's mesh, by the vectorised formulas of SURVEY appendix A.4
    (only an input generator for the generic-mesh cases: every array is then scrambled)."""
    vol = np.asarray(vol)
    Z, Y, X = vol.shape
    RN, SN = X + 1, (X + 1) * (Y + 1)
    lin = np.flatnonzero(vol.ravel() > thr)
    s, r, c = np.unravel_index(lin, vol.shape)
    b = SN * s + RN * r + c
    cells = b[:, None] + np.asarray([0, 1, RN + 1, RN, SN, SN + 1, SN + RN + 1, SN + RN])
    zz, yy, xx = np.meshgrid(np.arange(Z + 1), np.arange(Y + 1), np.arange(X + 1), indexing="ij")
    points = np.stack([xx.ravel() * voxelsize[0], yy.ravel() * voxelsize[1], zz.ravel() * voxelsize[2]], axis=1)
    return points, cells.astype(np.int64), vol.ravel()[lin]


def generic_cases():
    """Meshes that did NOT come out of vol2ugrid unchanged (mesh2voxelfe accepts any hexahedral meshio
    mesh, voxelFE.py:612-688).  Each case: name, points, blocks [(type, cells)], cell_data (dict),
    point_sets, cell_sets, m (mesh2voxelfe kwargs)."""
    C = []
    rng = np.random.RandomState(77)

    def add(name, points, blocks, cell_data, point_sets, cell_sets, m=None):
        m = dict(m or {})
        m.setdefault("templatefile", TEMPLATE)
        C.append(dict(name=name, points=points, blocks=blocks, cell_data=cell_data, point_sets=point_sets,
                      cell_sets=cell_sets, m=m))

    g = rand_grey((5, 6, 13), 0.5, 31, 1, 6)
    pts, cells, gv = _grid_mesh(g)
    n_el, n_pt = len(cells), len(pts)
    perm = rng.permutation(n_el)
    std_ps = {"NODES_A": sorted(rng.choice(n_pt, 23, replace=False).tolist()), "NODES_B": [5, 3, 3, 1]}
    std_cs = {"CELLS_A": list(range(1, 11)), "CELLS_B": []}

    # cells in a scrambled order, nodes anywhere in space
    rpts = rng.uniform(-50, 50, size=(n_pt, 3))
    rpts[0] = [-0.0, -1e-9, 0.0000005]
    rpts[1] = [123456789.123456789, -99999.9999996, 99999.9999994]
    rpts[2] = [1e15 + 0.5, 2.5e-7, -2.5e-7]
    rpts[3] = [0.1234565, 0.1234575, 1e17]
    add("scrambled_cells_random_points", rpts, [("hexahedron", cells[perm])], {"GV": [gv[perm]]}, std_ps, std_cs)
    # float grey values: several GVs share int(GV) -> repeated SETn headers (voxelFE.py:641)
    fgv = np.asarray([1.5, 2.25, 2.75, 7.0])[rng.randint(0, 4, n_el)]
    add("float_gv", pts, [("hexahedron", cells)], {"GV": [fgv]}, std_ps, std_cs)
    add("float32_gv_property", pts, [("hexahedron", cells)], {"density": [fgv.astype(np.float32)]}, std_ps, std_cs,
        dict(keywords=['NSET', 'ELSET', 'PROPERTY'], matprop=lambda: {"file": [MAT_LINEAR], "range": [[1, 255]]}))
    igv = np.asarray([3, 1000, 70000, 2 ** 18], dtype=np.int64)[rng.randint(0, 4, n_el)]
    add("int64_gv_beyond_8bit", pts, [("hexahedron", cells)], {"GV": [igv]}, std_ps, std_cs,
        dict(keywords=['NSET', 'ELSET', 'PROPERTY'], matprop=lambda: {"file": [MAT_POWER]}))
    add("uint16_gv", pts, [("hexahedron", cells)], {"GV": [(gv.astype(np.uint16) * 9000)]}, std_ps, std_cs)
    add("int8_gv_negative", pts, [("hexahedron", cells)], {"GV": [gv.astype(np.int8) - 3]}, std_ps, std_cs)
    add("bool_gv", pts, [("hexahedron", cells[perm])], {"GV": [np.ones(n_el, dtype=bool)]}, std_ps, std_cs)
    add("binary_model_01", pts, [("hexahedron", cells)], {"GV": [(gv > 3).astype(np.uint8)]}, std_ps, std_cs,
        dict(keywords=['NSET', 'ELSET', 'PROPERTY'], matprop=lambda: {"file": [MAT_CONST]}))
    # the notebooks' pattern: cell_data value is an array, not a list (voxelFE.py:560-562)
    add("cell_data_plain_array", pts, [("hexahedron", cells)], {"GV": gv.astype(np.uint8)}, std_ps, std_cs)
    # set flavours: numpy arrays, unsorted with duplicates, wrap at 10, empty, many sets
    ps = {"top": np.arange(10), "bottom nodes": np.asarray([7, 7, 2]), "none": [], "all": list(range(n_pt)),
          "big_ids": [n_pt - 1, 0]}
    cs = {"E10": list(range(1, 21)), "E_last": np.asarray([n_el]), "E_empty": np.zeros(0, dtype=np.int64)}
    add("custom_sets", pts, [("hexahedron", cells)], {"GV": [gv]}, ps, cs)
    add("custom_sets_nset_only", pts, [("hexahedron", cells)], {"GV": [gv]}, ps, {}, dict(keywords=['NSET']))
    # two hexahedron blocks: lines come from block 0, used nodes from both (voxelFE.py:628 vs :647)
    h = n_el // 2
    add("two_hex_blocks", pts, [("hexahedron", cells[:h]), ("hexahedron", cells[h:])],
        {"GV": [gv[:h], gv[h:]]}, std_ps, std_cs)
    # point dtypes
    ipts = np.stack([np.arange(n_pt) % 14, (np.arange(n_pt) // 14) % 7 * 3, np.arange(n_pt) // 98 * 100000], axis=1)
    add("points_int64", ipts.astype(np.int64), [("hexahedron", cells)], {"GV": [gv]}, std_ps, std_cs)
    add("points_float32", pts.astype(np.float32) * np.float32(0.1), [("hexahedron", cells)], {"GV": [gv]},
        std_ps, std_cs)
    add("points_wide", pts * np.asarray([30000.5, -1e6, 123456.789]), [("hexahedron", cells)], {"GV": [gv]},
        std_ps, std_cs)
    # a mesh using only a few of its points, ids far apart
    few = np.asarray([[0, 1, 15, 14, 98, 99, 113, 112], [n_pt - 1, n_pt - 2, 5, 4, 3, 2, 1, 0]], dtype=np.int64)
    add("two_cells_sparse_nodes", rpts, [("hexahedron", few)], {"GV": [np.asarray([9, 4], dtype=np.uint8)]},
        {"N": [0]}, {"C": [2, 1]}, dict(eltype="C3D8R", keywords=['ELSET', 'NSET']))
    add("single_cell", rpts[:8], [("hexahedron", np.arange(8, dtype=np.int64)[None, ::-1])],
        {"GV": [np.asarray([200], dtype=np.uint8)]}, {"N": []}, {"C": []},
        dict(templatefile=TEMPLATE_REFNODE, refnode=[1.23456, 2.0, -0.00004]))
    return C


def edit_mesh_sets(mesh):
    """The host-side edit of a vol2ugrid mesh used by the `edited_*` cases (applied alike to the
    reference's mesh and to ciclope_b200's VoxelMesh)."""
    ps = mesh.point_sets
    ps["NODES_EXTRA"] = ps["NODES_Z1"][::3]
    ps["NODES_X0"] = ps["NODES_X0"][:7]
    del ps["NODES_X0Y0Z1"]
    mesh.cell_sets["CELLS_Z0"] = [int(v) for v in mesh.cell_sets["CELLS_Z0"]][::-1]
    return mesh


def edit_mesh_gv(mesh):
    """Overwrite the grey values on the host (cell_data assignment, as the notebooks do)."""
    gv = np.asarray(mesh.cell_data["GV"][0]).astype(np.int64)
    mesh.cell_data["GV"] = [(gv % 5 + 1) * 300]
    return mesh


def edited_cases():
    g = rand_grey((6, 7, 20), 0.5, 10)
    return [dict(name="edited_sets", vol=g, u=dict(voxelsize=[0.0195] * 3), edit=edit_mesh_sets,
                 m=dict(templatefile=TEMPLATE)),
            dict(name="edited_gv", vol=g, u=dict(voxelsize=[0.0195] * 3, GVmin=50), edit=edit_mesh_gv,
                 m=dict(templatefile=TEMPLATE, keywords=['NSET', 'ELSET', 'PROPERTY'],
                        matprop=lambda: {"file": [MAT_POWER]}))]


def copy_cell_data(cd):
    return {k: ([np.array(a) for a in v] if isinstance(v, list) else np.array(v)) for k, v in cd.items()}


class PlainMesh:
    """Minimal meshio-5.0.0-like container (namedtuple cell blocks, cells_dict concatenating the
    blocks of one type) used to hand generic meshes to the oracle and to ciclope_b200."""

    def __init__(self, case):
        import collections
        CellBlock = collections.namedtuple("CellBlock", ["type", "data"])
        self.points = np.asarray(case["points"])
        self.cells = [CellBlock(t, np.asarray(d)) for t, d in case["blocks"]]
        self.cell_data = copy_cell_data(case["cell_data"])
        self.point_sets = {k: (v.copy() if isinstance(v, np.ndarray) else list(v)) for k, v in case["point_sets"].items()}
        self.cell_sets = {k: (v.copy() if isinstance(v, np.ndarray) else list(v)) for k, v in case["cell_sets"].items()}

    @property
    def cells_dict(self):
        d = {}
        for t, data in self.cells:
            d.setdefault(t, []).append(data)
        return {k: np.concatenate(v) for k, v in d.items()}


# ------------------------------------------------------------------------------------ CCL reuse

def fill_cases():
    """(name, image, kwargs) for fill_voids (preprocess.py:145-185): images with enclosed background."""
    C = []
    box = np.zeros((7, 8, 40), dtype=np.uint8)
    box[1:6, 1:7, 2:38] = 120
    box[2:5, 2:6, 3:37] = 0          # one big cavity
    box[3, 3, 10] = 200              # a speck inside the cavity
    C.append(("hollow_box_u8", box, {}))
    C.append(("hollow_box_fill77_copy", box, dict(fill_val=77, makecopy=True)))
    C.append(("hollow_box_fill0", box, dict(fill_val=0)))
    two = np.zeros((5, 9, 70), dtype=np.uint8)
    two[1:4, 1:8, 1:30] = 50
    two[2, 2:7, 2:29] = 0            # cavity A (5 x 27)
    two[1:4, 1:8, 33:69] = 90
    two[2, 2:7, 34:68] = 0           # cavity B larger than A, still smaller than the outside
    C.append(("two_cavities", two, {}))
    # the "outside" is NOT the largest background region: the reference then fills the outside instead
    inv = np.zeros((3, 12, 34), dtype=np.uint8)
    inv[:, 1:11, 1:33] = 255
    inv[:, 2:10, 2:32] = 0
    inv[:, 0, :] = 255
    inv[:, 11, :] = 255
    inv[:, :, 0] = 255
    inv[:, :, 33] = 255
    inv[0, 5, 0] = 0
    C.append(("cavity_larger_than_outside", inv, {}))
    C.append(("bool_image", box > 0, {}))
    C.append(("bool_image_copy", box > 0, dict(makecopy=True)))
    for i, (shape, dens) in enumerate([((6, 7, 33), 0.75), ((4, 9, 64), 0.8), ((9, 10, 31), 0.7),
                                       ((12, 13, 47), 0.85), ((1, 20, 45), 0.7), ((16, 16, 96), 0.78)]):
        C.append(("rand_dense_%dx%dx%d" % shape, rand_grey(shape, dens, 200 + i), {}))
    m = blobs((20, 24, 70), 0.6, 9)
    C.append(("blobs_dense_grey", grey_from_mask(m, 10), dict(fill_val=255, makecopy=True)))
    return C


def parosol_cases():
    """(name, volume, kwargs of vol2h5ParOSol) -- small volumes whose non-zero voxels do not fill the array
    (so that the bounding-box crop does something), bool and uint8, both locking strategies."""
    rng = np.random.default_rng(77)
    out = []
    shapes = [(5, 6, 7), (9, 8, 10), (3, 12, 5), (1, 4, 4), (2, 3, 3), (12, 33, 40), (7, 64, 65)]
    for si, shape in enumerate(shapes):
        for dens in (0.25, 0.6):
            for dt in ("bool", "u01", "grey"):
                if dt == "grey" and (si > 1 or dens > 0.5):
                    continue
                v = rng.random(shape) < dens
                if shape[0] > 2:
                    v[0] = False
                v[:, 0] = False
                v[:, :, -1] = False
                if not v.any():
                    v[-1, -1, 0] = True
                if dt == "u01":
                    v = v.astype(np.uint8)
                elif dt == "grey":
                    # recon_utils.bbox looks for projections EQUAL to 1: grey values above 1 hide a row /
                    # column / slice from it (often a ValueError, sometimes a smaller box)
                    g = (v * rng.integers(1, 256, shape)).astype(np.uint8)
                    if si == 1:
                        # ones, with one grey voxel in the far corner of their box: its slice, row and column
                        # have maximum 7, so the reference's box stops one short of it on every axis
                        g = v.astype(np.uint8)
                        g[-1, -1, -2] = 7
                    v = g
                for strat, kw in (("plane", dict(plane_lock_num=1)), ("plane", dict(plane_lock_num=2)),
                                  ("exact", dict(node_level_lock_num=1)), ("exact", dict(node_level_lock_num=3))):
                    for thf in (True, False):
                        name = "s%d_d%d_%s_%s%d_%d" % (si, int(dens * 100), dt, strat,
                                                       list(kw.values())[0], int(thf))
                        k = dict(topDisplacement=-0.04, voxelsize=0.0195, poisson_ratio=0.3, young_modulus=18e3,
                                 topHorizontaFixedlDisplacement=thf, locking_strategy=strat)
                        k.update(kw)
                        out.append((name, v, k))
    # a blob volume, anisotropic voxel size given as a list, another modulus
    b = blobs((20, 24, 28), 0.3, 3)
    out.append(("blobs_plane", b, dict(topDisplacement=0.1, voxelsize=[0.01, 0.02, 0.03], poisson_ratio=0.25,
                                       young_modulus=1234.5, topHorizontaFixedlDisplacement=True,
                                       locking_strategy="plane", plane_lock_num=1)))
    g = grey_from_mask(b, 5)
    out.append(("grey_exact", g, dict(topDisplacement=-1.0, voxelsize=0.5, poisson_ratio=0.3, young_modulus=70.0,
                                      topHorizontaFixedlDisplacement=False, locking_strategy="exact",
                                      node_level_lock_num=2)))
    return out


def segment_cases():
    """(name, uint8 image, threshold)"""
    rng = np.random.default_rng(5)
    out = []
    for si, shape in enumerate([(4, 5, 6), (3, 17, 33), (8, 16, 64), (5, 7, 11)]):
        img = rng.integers(0, 256, shape).astype(np.uint8)
        for t in (0, 1, 63, 127, 128, 200, 254, 255, 300, -5, 63.7):
            out.append(("s%d_t%s" % (si, t), img, t))
    return out


def check_parosol_against_golden(g, name, run):
    """run() -> {dataset name: array} (or raises); compared with what the unmodified reference handed to
    h5py for the same case: same error class, same dataset order, same dtype / shape / bytes."""
    import hashlib
    key = "po/" + name
    if key + "/error" in g.files:
        try:
            run()
        except BaseException as e:  # noqa: BLE001
            assert type(e).__name__ == str(g[key + "/error"]), (name, type(e).__name__)
            return
        raise AssertionError("%s: the reference raises %s" % (name, g[key + "/error"]))
    ds = run()
    names = [str(n) for n in g[key + "/names"]]
    assert list(ds) == names
    for k in names:
        want = g[key + "/" + k]
        got = np.asarray(ds[k])
        if want.dtype.kind == "U":          # dtype / shape / SHA-256
            assert got.dtype.str == str(want[0]) and str(got.shape) == str(want[1]), (name, k)
            assert hashlib.sha256(np.ascontiguousarray(got).tobytes()).hexdigest() == str(want[2]), (name, k)
        else:
            assert got.dtype == want.dtype and got.shape == want.shape, (name, k, got.dtype, got.shape)
            assert np.array_equal(got, want), (name, k)


def embed_cases():
    """(name, uint8 image, kwargs of embed) -- small images with an empty margin, all six directions,
    depths that run past the margin (negative slice starts), padding, explicit / default value."""
    rng = np.random.default_rng(21)
    out = []
    for si, shape in enumerate([(12, 14, 16), (9, 20, 11), (16, 33, 40)]):
        for dens in (0.15, 0.5):
            img = np.zeros(shape, np.uint8)
            inner = rng.random(tuple(s - 6 for s in shape)) < dens
            img[3:-3, 3:-3, 3:-3] = inner * rng.integers(1, 200, inner.shape)
            for d in ["-z", "+z", "-x", "+x", "-y", "+y"]:
                for depth, pad, val, mc in [(1, 0, None, True), (3, 2, 77, False), (6, 0, 250, True),
                                            (2, 5, None, False)]:
                    out.append(("s%d_d%d_%s_%d_%d_%s_%d" % (si, int(dens * 100), d, depth, pad, val, int(mc)), img,
                                dict(embed_depth=depth, embed_dir=d, embed_val=val, pad=pad, makecopy=mc)))
    full = np.full((6, 6, 6), 255, np.uint8)
    out.append(("full_255", full, dict(embed_depth=2, embed_dir="-z", embed_val=None, pad=0, makecopy=True)))
    out.append(("empty", np.zeros((5, 5, 5), np.uint8), dict(embed_depth=2, embed_dir="+x", embed_val=None, pad=0,
                                                             makecopy=True)))
    out.append(("bad_dir", img, dict(embed_depth=2, embed_dir="up", embed_val=1, pad=0, makecopy=True)))
    return out


def periosteum_cases():
    """(name, bool image 2D / 3D, kwargs of periosteummask): hollow shapes (rings that closing seals, rings it
    does not), specks, shapes touching the border, widths that are not a multiple of 32."""
    rng = np.random.default_rng(31)
    out = []
    for si, shape in enumerate([(6, 40, 50), (3, 33, 65), (40, 50), (5, 64, 96), (2, 70, 31)]):
        yy, xx = np.ogrid[:shape[-2], :shape[-1]]
        cy, cx = shape[-2] / 2.0, shape[-1] / 2.0
        rad = min(shape[-2:]) / 3.0
        d = np.hypot(yy - cy, xx - cx)
        ring = (d < rad) & (d > rad - 2)
        gap = ring & ~((abs(yy - cy) < 2) & (xx > cx))          # ring with a 3-pixel opening
        for ri, r in enumerate((0, 1, 3, 10)):
            for dens in (0.02, 0.1):
                bw = rng.random(shape) < dens
                bw = bw | (gap if (ri + si) % 2 else ring)
                if si == 3:
                    bw[..., :3, :] = True                        # a slab along the border
                for ru in (True, False):
                    out.append(("s%d_r%d_d%d_%d" % (si, r, int(dens * 100), int(ru)), bw,
                                dict(closepixels=r, removeunconn=ru)))
    out.append(("empty", np.zeros((3, 20, 20), dtype=bool), dict(closepixels=2, removeunconn=True)))
    out.append(("full", np.ones((2, 20, 20), dtype=bool), dict(closepixels=2, removeunconn=True)))
    # final 3D closing with cube(closevoxels) (preprocess.py:405-408): even and odd widths, shapes touching
    # the borders, widths that are not a multiple of 32
    for ci, (shape, cw) in enumerate([((9, 40, 50), 2), ((9, 40, 50), 3), ((12, 33, 65), 4), ((12, 33, 65), 5),
                                      ((7, 20, 31), 7), ((10, 24, 96), 6)]):
        bw = rng.random(shape) < 0.08
        bw[2:-2, 8:-8, 10:-10] |= rng.random((shape[0] - 4, shape[1] - 16, shape[2] - 20)) < 0.5
        bw[:, :2, :] |= rng.random((shape[0], 2, shape[2])) < 0.5
        for ru in (True, False):
            out.append(("cube%d_w%d_%d" % (ci, cw, int(ru)), bw, dict(closepixels=1, closevoxels=cw, removeunconn=ru)))
    return out
