/*
 * ciclope_b200 -- C ABI of the B200-native CT -> CalculiX .INP hot path.
 *
 * This is the drop-in boundary: everything the Python facade (ciclope_b200/voxelFE.py,
 * ciclope_b200/preprocess.py) needs from the GPU goes through these entry points, and a
 * ciclope maintainer can bind the same symbols with ctypes.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host;
 *   - sizes and indices are int64_t; volumes are uint8 [Z][Y][X] (C order, x fastest);
 *   - the last argument is the CUDA stream the work is enqueued on; nothing synchronises;
 *   - no hidden allocation: scratch/workspace buffers are passed in by the caller;
 *   - return 0 on success, < 0 on error (cfe_last_error() holds the message).
 *
 * Each entry point names the reference lines (ciclope 2.0.3, paths relative to the ciclope
 * checkout) whose work it replaces.
 */
#ifndef CICLOPE_B200_H
#define CICLOPE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* cfe_stream_t; /* == cudaStream_t */

#define CFE_ABI_VERSION 1

int cfe_version(void);
const char* cfe_last_error(void);
/* kernels launched by this library since it was loaded (benchmark bookkeeping) */
uint64_t cfe_launch_count(void);

/* ------------------------------------------------------------------ bit packing + scans */

/* bits[row][w] bit b = (vol[row][32w+b] > thr); W = ceil(X/32) words per row, pad bits 0.
 * `GV > GVmin` of src/ciclope/core/voxelFE.py:140-141 (thr = floor(GVmin)) and the
 * foreground test of measure.label, src/ciclope/utils/preprocess.py:135 (thr = 0). */
int cfe_pack_bits(const uint8_t* vol, int64_t rows, int64_t X, int thr, uint32_t* bits, cfe_stream_t s);
/* bit = (vol <= thr): the background `~(I>0)` labelled by fill_voids, src/ciclope/utils/preprocess.py:167 */
int cfe_pack_bits_le(const uint8_t* vol, int64_t rows, int64_t X, int thr, uint32_t* bits, cfe_stream_t s);
/* *out = max(*out, max(vol[0..n))): `I.max()`, the default fill value of fill_voids (preprocess.py:163) */
int cfe_max_u8(const uint8_t* vol, int64_t n, uint32_t* out, cfe_stream_t s);
/* out[0] = max v, out[1] = min of the non-zero v (0xffffffff if none): `measure.label(bwimage, None, True, 1)`
 * (src/ciclope/utils/preprocess.py:135) labels regions of EQUAL value, so a uint8 image is only read as a
 * binary one when both agree; the facade raises otherwise. */
int cfe_value_span_u8(const uint8_t* vol, int64_t n, uint32_t* out, cfe_stream_t s);

/* bytes of look-back workspace for a scan over n_items items */
size_t cfe_scan_workspace_bytes(int64_t n_items);

/* pref[i] = number of elements (set bits) before word i; *total = n_el.
 * The `cell_i += 1` counter of voxelFE.py:142. */
int cfe_scan_elements(const uint32_t* bits, int64_t n_words, uint32_t* pref, uint64_t* total, void* ws,
                      cfe_stream_t s);

/* pref[i] = number of x-run starts before word i; *total = number of runs (CCL vertices). */
int cfe_scan_run_starts(const uint32_t* bits, int64_t rows, int64_t X, uint32_t* pref, uint64_t* total,
                        void* ws, cfe_stream_t s);

/* Used-node bitmap + numbering: nbits[(layer,row)][w] bit b = grid node (layer,row,32w+b) is
 * a corner of >= 1 element; npref = exclusive count of used nodes; *total = used nodes.
 * np.unique(mesh.cells_dict["hexahedron"]) of voxelFE.py:628 without building the cells.
 * halo_below: element bits [Y][W] of the voxel plane under local slice 0 (z-slab sharding)
 * or NULL; n_layers = node layers owned by this slab (Zl, or Zl+1 for the top slab). */
int cfe_scan_nodes(const uint32_t* bits, const uint32_t* halo_below, int64_t Zl, int64_t Y, int64_t X,
                   int64_t n_layers, uint32_t* nbits, uint32_t* npref, uint64_t* total, void* ws,
                   cfe_stream_t s);
/* Same, and chunk_word[m] = index of the node-grid word holding the used node of rank 512 m (u32
 * [grid nodes / 512 + 2]): the entry point of the list-free *NODE emitter cfe_emit_nodes_bits. */
int cfe_scan_nodes_index(const uint32_t* bits, const uint32_t* halo_below, int64_t Zl, int64_t Y, int64_t X,
                         int64_t n_layers, uint32_t* nbits, uint32_t* npref, uint64_t* total,
                         uint32_t* chunk_word, void* ws, cfe_stream_t s);

/* list[pref[i] + k] = row*row_len + 32*w + (k-th set bit of word i): ascending index lists of
 * the elements (row_len = X) or used nodes (row_len = X+1). */
int cfe_fill_list(const uint32_t* bits, const uint32_t* pref, int64_t n_words, int64_t words_per_row,
                  int64_t row_len, uint32_t* list, cfe_stream_t s);

/* cfe_fill_list for the elements, fused with the *ELEMENT emitter's line-length pass: also writes
 * len8 / nd8 into the element workspace `ws` (cfe_emit_elements_workspace_bytes(n_el) bytes) so that
 * cfe_element_lengths(have_lengths = 1) only has to sum tile totals (raster order, same z0/eid_offset). */
int cfe_fill_elements(const uint32_t* bits, const uint32_t* pref, int64_t n_words, int64_t Y, int64_t X,
                      int64_t z0, int64_t eid_offset, int64_t n_el, uint32_t* list, void* ws, cfe_stream_t s);
/* same, and gv8[j] = vol[list[j]]: `cell_GV.append(GV)` of voxelFE.py:160, the grey value of every element
 * in element order (vol = the slab the bits were packed from) */
int cfe_fill_elements_gv(const uint32_t* bits, const uint32_t* pref, int64_t n_words, int64_t Y, int64_t X,
                         int64_t z0, int64_t eid_offset, int64_t n_el, uint32_t* list, void* ws,
                         const uint8_t* vol, uint8_t* gv8, cfe_stream_t s);

/* generic exclusive scan of u32 counts into u64 offsets (tile counts, histograms) */
int cfe_scan_u32(const uint32_t* in, int64_t n, uint64_t* out, uint64_t* total, void* ws, cfe_stream_t s);

/* ------------------------------------------------------------------ CCL */

/* capacity (entries) of the parent / size scratch arrays for a slab of `rows` rows */
size_t cfe_ccl_max_runs(int64_t rows, int64_t X);

/* Largest 6-connected component of the slab: measure.label(bw, None, True, 1) + np.bincount +
 * argmax + (labels == id), src/ciclope/utils/preprocess.py:135-143.  Ties go to the component
 * whose first voxel comes first in C order.
 *   *best = (size << 32) | (0xffffffff - root run id), 0 when there is no foreground. */
int cfe_ccl_largest(const uint32_t* bits, const uint32_t* run_pref, const uint64_t* n_runs, int64_t Z,
                    int64_t Y, int64_t X, uint32_t* parent, uint32_t* size, uint64_t* best,
                    uint8_t* out_mask, cfe_stream_t s);

/* The same work in three phases, for z-slab sharding (the cross-slab merge sits between 1 and 2):
 *   cfe_ccl_label       parent[run] = slab-local root, size[root] = voxels inside the slab
 *   cfe_ccl_select      *best = max over roots of (size << 32) | (0xffffffff - root)
 *   cfe_ccl_write_mask  mask = runs whose root is the winner (keep == NULL) or has keep[root] == 0xffffffff */
int cfe_ccl_label(const uint32_t* bits, const uint32_t* run_pref, const uint64_t* n_runs, int64_t Z, int64_t Y,
                  int64_t X, uint32_t* parent, uint32_t* size, cfe_stream_t s);
int cfe_ccl_select(const uint32_t* parent, const uint32_t* size, const uint64_t* n_runs, uint64_t* best,
                   cfe_stream_t s);
int cfe_ccl_write_mask(const uint32_t* bits, const uint32_t* run_pref, int64_t Z, int64_t Y, int64_t X,
                       const uint32_t* parent, const uint64_t* best, const uint32_t* keep, uint8_t* out_mask,
                       cfe_stream_t s);
/* General phase 3 and the matching one-call form.  Selected voxels = the winner's (drop = 0) or every
 * labelled voxel except the winner's (drop = 1): `labels[labels == largest] = 0; labels != 0` of
 * remove_largest (preprocess.py:204-206) and of fill_voids (:176-177).  out = fill on selected voxels,
 * src[voxel] elsewhere (0 when src is NULL); out == src is allowed (`I[labels != 0] = fill_val`, :184). */
int cfe_ccl_write_select(const uint32_t* bits, const uint32_t* run_pref, int64_t Z, int64_t Y, int64_t X,
                         const uint32_t* parent, const uint64_t* best, const uint32_t* keep, int drop,
                         const uint8_t* src, int fill, uint8_t* out, cfe_stream_t s);
/* cfe_ccl_largest / cfe_ccl_write_mask that also leave the kept voxels bit-packed in sel_bits (u32
 * [Z*Y*ceil(X/32)], the layout of cfe_pack_bits), so that `vol2ugrid(remove_unconnected(bw))`
 * (src/ciclope/pipeline.py:177,222) starts from the bits instead of re-reading the byte mask.  Needs
 * X % 32 == 0 and a 16-byte aligned out_mask (error otherwise). */
int cfe_ccl_largest_bits(const uint32_t* bits, const uint32_t* run_pref, const uint64_t* n_runs, int64_t Z,
                         int64_t Y, int64_t X, uint32_t* parent, uint32_t* size, uint64_t* best,
                         uint8_t* out_mask, uint32_t* sel_bits, cfe_stream_t s);
int cfe_ccl_write_mask_bits(const uint32_t* bits, const uint32_t* run_pref, int64_t Z, int64_t Y, int64_t X,
                            const uint32_t* parent, const uint64_t* best, const uint32_t* keep,
                            uint8_t* out_mask, uint32_t* sel_bits, cfe_stream_t s);
int cfe_ccl_largest_select(const uint32_t* bits, const uint32_t* run_pref, const uint64_t* n_runs, int64_t Z,
                           int64_t Y, int64_t X, uint32_t* parent, uint32_t* size, uint64_t* best, int drop,
                           const uint8_t* src, int fill, uint8_t* out, cfe_stream_t s);
/* roots[k] = slab-local root of the k-th run of voxel plane `plane`; *count = runs in the plane */
int cfe_ccl_plane_roots(const uint32_t* bits, const uint32_t* run_pref, int64_t plane, int64_t Y, int64_t X,
                        const uint32_t* parent, uint32_t* roots, uint32_t* count, cfe_stream_t s);
/* edges (rootA, rootB) between the neighbour slab's top plane A (bits, plane-local run prefixes,
 * run roots: received over NCCL) and this slab's bottom plane B; *count may exceed capacity */
int cfe_ccl_interface_edges(const uint32_t* bitsA, const uint32_t* prefA, const uint32_t* rootsA,
                            const uint32_t* bitsB, const uint32_t* prefB, const uint32_t* parentB, int64_t Y,
                            int64_t X, uint32_t* edges, uint32_t* count, int64_t capacity, cfe_stream_t s);

/* The same interface walk for the device-side merge: runs carry dense boundary-vertex indices (vA: the
 * neighbour's numbering, of its first vA_len plane runs; vB: this slab's; both < Cv); the pairs are united in the scratch forest
 * (u32[2 Cv]) and only the unions that hooked two trees are appended, i.e. a spanning forest of the
 * interface graph: <= 2 Cv - 1 edges however many run pairs overlap. */
int cfe_ccl_interface_forest(const uint32_t* bitsA, const uint32_t* prefA, const uint32_t* vA, int64_t vA_len,
                             const uint32_t* bitsB, const uint32_t* prefB, const uint32_t* vB, int64_t Y,
                             int64_t X, int64_t Cv, uint32_t* forest, uint32_t* edges, uint32_t* count,
                             int64_t capacity, cfe_stream_t s);

/* Device-side cross-slab merge.  cfe_ccl_boundary_vertices numbers this slab's boundary components
 * (roots owning a run of voxel plane 0 / Zl-1), fills the per-rank message
 *   pack = [n_vertices, n_edges, best internal (u64), roots[Cv], sizes[Cv], edges[Ce][2]]  (u32)
 * and the per-run vertex indices vbot / vtop; cfe_ccl_interface_forest appends the edges; after the all-gather cfe_ccl_boundary_solve merges the
 * graph, picks the winner (largest, ties -> first in raster order) and writes the keep markers.
 * result = {size, key = rank << 32 | root, capacity-overflow flag, winner vertex or -1}. */
int cfe_ccl_boundary_vertices(const uint32_t* roots_bot, const uint32_t* roots_top, const uint32_t* cnt,
                              int use_bot, int use_top, const uint64_t* n_runs, uint32_t* vfirst,
                              uint32_t* didx, const uint32_t* parent, const uint32_t* size, int64_t Cv,
                              uint32_t* pack, uint32_t* vbot, uint32_t* vtop, cfe_stream_t s);
int cfe_ccl_boundary_solve(const uint32_t* packs, int N, int64_t Cv, int64_t Ce, int rank, uint32_t* label,
                           uint64_t* comp, uint64_t* ckey, uint32_t* size, int64_t* result, cfe_stream_t s);

/* connected components of the (small) cross-slab boundary graph: label[v] = smallest vertex index
 * of v's component; ea/eb = dense int64 endpoints of the edges */
int cfe_uf_solve(uint32_t* label, int64_t n_vertices, const int64_t* ea, const int64_t* eb, int64_t n_edges,
                 cfe_stream_t s);

/* ------------------------------------------------------------------ boundary sets */

/* One boundary set = the candidates of a box of the node grid (kind 0/1) or voxel grid
 * (kind 2), in raster order, kept when the predicate holds (voxelFE.py:163-192, 205-247):
 *   kind 0/1  node (s,r,c) is kept when one of its <= 8 adjacent voxels inside the voxel box
 *             [vs0,vs1) x [vr0,vr1) x [vc0,vc1) is an element ("plane": the locked voxel
 *             planes; "exact": the whole volume);  value = 0-based grid node id;
 *   kind 2    voxel (s,r,c) is kept when it is an element;  value = 1-based element id.
 * All coordinates are GLOBAL (z-slab sharding passes z0); boxes are half-open. */
typedef struct {
    int32_t kind;
    int32_t pad_;
    int64_t s0, s1, r0, r1, c0, c1;
    int64_t vs0, vs1, vr0, vr1, vc0, vc1;
    int64_t n_cand;    /* (s1-s0)*(r1-r0)*(c1-c0) */
    int64_t tile_base; /* first 256-candidate tile of this set */
} CfeSetDesc;

int cfe_sets_count(const CfeSetDesc* descs, int n_sets, int64_t n_tiles, int64_t Zl, int64_t Y, int64_t X,
                   int64_t z0, int64_t eid_offset, const uint32_t* bits, const uint32_t* pref,
                   const uint32_t* halo_below, uint32_t* tile_count, cfe_stream_t s);
int cfe_sets_fill(const CfeSetDesc* descs, int n_sets, int64_t n_tiles, int64_t Zl, int64_t Y, int64_t X,
                  int64_t z0, int64_t eid_offset, const uint32_t* bits, const uint32_t* pref,
                  const uint32_t* halo_below, const uint64_t* tile_off, int64_t* ids, cfe_stream_t s);

/* Between cfe_sets_count (+ cfe_scan_u32 of the tile counts, total in stats[2]) and cfe_sets_fill:
 * out = [stats[0] (n_el), stats[1] (n_nd), ids per set (n_sets values), stats[3] (field overflow)] --
 * the vector a z-slab rank all-gathers (global ids / positions need every rank's counts). */
int cfe_set_counts(const CfeSetDesc* descs, int n_sets, int64_t n_tiles, const uint64_t* tile_off,
                   const uint64_t* stats, int64_t* out, cfe_stream_t s);

/* Single-pass form of the two calls above (one predicate evaluation per candidate; positions by
 * decoupled look-back over the tiles).  Tiles hold cfe_sets_onepass_tile() candidates: tile_base of the
 * descriptors is counted in those units.  ids must hold sum(n_cand) entries, ws = look-back workspace for
 * n_tiles items (cfe_scan_workspace_bytes), first[q] = position of set q's first id (-1 for a set without
 * candidates), first[n_sets] = number of ids written. */
int64_t cfe_sets_onepass_tile(void);
int cfe_sets_onepass(const CfeSetDesc* descs, int n_sets, int64_t n_tiles, int64_t Zl, int64_t Y, int64_t X,
                     int64_t z0, int64_t eid_offset, const uint32_t* bits, const uint32_t* pref,
                     const uint32_t* halo_below, void* ws, int64_t* ids, int64_t* first, cfe_stream_t s);

/* Reference-node barycentre sums (voxelFE.py:208-223, 254-268): acc[q][0..2] += coordinates of
 * the ids of set q, float64, in ascending order (one thread per set: the reference's
 * summation order).  xs/ys/zs index GLOBAL column / row / layer. */
int cfe_refnode_sums(const int64_t* ids, const int64_t* first, const int64_t* count, int n_sets, int64_t Y,
                     int64_t X, const double* xs, const double* ys, const double* zs, double* acc,
                     cfe_stream_t s);

/* ------------------------------------------------------------------ lazy-mesh materialisers */

/* cells (n_el,8) int64 in C3D8 order, voxelFE.py:145-157 */
int cfe_fill_cells(const uint32_t* elem_vox, int64_t n_el, int64_t Y, int64_t X, int64_t z0, int64_t* cells,
                   cfe_stream_t s);
/* cell_GV, voxelFE.py:160 */
int cfe_gather_gv(const uint8_t* vol, const uint32_t* elem_vox, int64_t n_el, uint8_t* gv, cfe_stream_t s);
/* points (n_layers*(Y+1)*(X+1), 3) from per-axis value tables, voxelFE.py:202-203 */
int cfe_fill_points(const void* xs, const void* ys, const void* zs, int elem_bytes, int64_t n_layers,
                    int64_t Y, int64_t X, void* out, cfe_stream_t s);

/* ------------------------------------------------------------------ deck emission */

/* '{:12.6f}'.format(v) for every per-axis node coordinate (voxelFE.py:630), exact (round-half-even
 * of the binary value in 128-bit integers, as CPython does): 16-byte slots, 12 characters each.
 * *overflow (device int, caller-zeroed) is set when a field needs more than 12 characters. */
int cfe_format_coords(const double* vals, int64_t n, char* out16, int* overflow, cfe_stream_t s);

/* General path for coordinates whose '{:12.6f}' field is wider than 12 characters (|v| >= 1e5): exact
 * formatting into 32-byte slots (slot[0] = length, |v| < 1e18 else *bad = 1), per-node line lengths and
 * 256-node tile totals, then (after cfe_scan_u32 of the totals) the variable-length *NODE lines. */
int cfe_format_coords_wide(const double* vals, int64_t n, char* out32, int* bad, cfe_stream_t s);
int cfe_node_lengths_wide(const uint32_t* node_list, int64_t n_nodes, const char* xtab32, const char* ytab32,
                          const char* ztab32, int64_t Y, int64_t X, int64_t z0, uint8_t* len8,
                          uint32_t* tile_tot, cfe_stream_t s);
int cfe_emit_nodes_wide(const uint32_t* node_list, int64_t n_nodes, const char* xtab32, const char* ytab32,
                        const char* ztab32, int64_t Y, int64_t X, int64_t z0, const uint8_t* len8,
                        const uint64_t* tile_off, char* out, cfe_stream_t s);

/* *NODE lines (voxelFE.py:629-630).  xtab/ytab/ztab: 16-byte slots holding the 12-character
 * '{:12.6f}' field of every column / row / layer coordinate (formatted on the host by the
 * same interpreter arithmetic as the reference).  Writes 53*n_nodes bytes at out. */
int cfe_emit_nodes(const uint32_t* node_list, int64_t n_nodes, const char* xtab, const char* ytab,
                   const char* ztab, int64_t Y, int64_t X, int64_t z0, char* out, cfe_stream_t s);
/* The same lines straight from the used-node bitmap (nbits / npref / chunk_word of cfe_scan_nodes_index,
 * n_words = words of the bitmap): every CTA expands its 512 nodes itself, no node list is materialised.
 * `out` must be 16-byte aligned. */
int cfe_emit_nodes_bits(const uint32_t* nbits, const uint32_t* npref, const uint32_t* chunk_word, int64_t n_words,
                        int64_t n_nodes, const char* xtab, const char* ytab, const char* ztab, int64_t Y,
                        int64_t X, int64_t z0, char* out, cfe_stream_t s);

/* *ELEMENT lines (voxelFE.py:639-647).  perm == NULL: raster order, one grey value
 * (single_gv); otherwise perm holds {voxel, raster rank} pairs in GV-major order.
 * first[g] (256 entries, non-decreasing) = emission index of the first element whose grey value is >= g
 * (n_el when there is none): block gv starts at first[gv] and the grey value of emission index j is the
 * largest g with first[g] <= j, so the volume is not read again;
 * hdr_text/hdr_len = '*ELEMENT, TYPE=.., ELSET=SET<gv>\n' per gv in 64-byte slots.
 * *total = bytes written; blk_off[gv] = offset where block gv (header included) starts (optional).
 * Two entry points (three launches): cfe_element_lengths (line lengths + exclusive scan of the
 * 32-line tile totals) then cfe_emit_elements (formatting); ws = cfe_emit_elements_workspace_bytes(n_el)
 * bytes of scratch shared by both (it also holds the counter from which the emitter's warps take their
 * chunks of tiles: one workspace per deck in flight, not to be shared by concurrent calls).
 * n_el < 2^32 (elem_vox holds 32-bit voxel indices of one slab). */
size_t cfe_emit_elements_workspace_bytes(int64_t n_el);
/* pass 1: line lengths + tile offsets; *total = bytes of the section */
int cfe_element_lengths(const uint32_t* elem_vox, const void* perm, int64_t n_el, const uint8_t* vol,
                        const uint64_t* first, const char* hdr_text, const uint8_t* hdr_len, int single_gv,
                        int64_t Zl, int64_t Y, int64_t X, int64_t z0, int64_t eid_offset, int have_lengths,
                        uint64_t* total, void* ws, cfe_stream_t s);
/* pass 2: the lines themselves (same arguments, same workspace) */
int cfe_emit_elements(const uint32_t* elem_vox, const void* perm, int64_t n_el, const uint8_t* vol,
                      const uint64_t* first, const char* hdr_text, const uint8_t* hdr_len, int single_gv,
                      int64_t Zl, int64_t Y, int64_t X, int64_t z0, int64_t eid_offset, char* out,
                      uint64_t* blk_off, void* ws, cfe_stream_t s);

/* *NSET / *ELSET sections (voxelFE.py:650-688) as a stream of items. */
typedef struct {
    int32_t kind;     /* 0 = literal text, 1 = id list */
    int32_t pad_;
    int64_t item0;    /* index of the segment's first item */
    int64_t lit_off;  /* kind 0: offset / length into `literals` */
    int64_t lit_len;
    int64_t id0;      /* kind 1: first entry in `ids` */
    int64_t rank0;    /* kind 1: 0-based position of that entry inside its (global) set */
} CfeItemSeg;

/* seg_off (optional, [n_segs]): byte offset of every segment inside the emitted stream */
int cfe_emit_items(const CfeItemSeg* segs, int n_segs, int64_t n_items, const int64_t* ids,
                   const char* literals, char* out_base, const uint64_t* dyn_off, uint64_t* total,
                   uint64_t* seg_off, void* ws, cfe_stream_t s);

/* literal text to out + *dyn0 + *dyn1 (either may be NULL) */
int cfe_put_literal(const char* src, int64_t n, char* out, const uint64_t* dyn0, const uint64_t* dyn1,
                    cfe_stream_t s);

/* ------------------------------------------------------------------ generic (array-driven) meshes */

/* mesh2voxelfe of ANY hexahedral meshio mesh (voxelFE.py:612-688 reads only points, cells, the first
 * cell_data entry and the sets).  ids = np.unique(cells); slots = cfe_format_coords_wide(points[ids]);
 * order / group = element indices grouped by ascending grey value; tile_off = cfe_scan_u32(tile_tot). */
int cfe_gen_node_lengths(const int64_t* ids, int64_t n, const char* slots, uint8_t* len8, uint32_t* tile_tot,
                         cfe_stream_t s);
int cfe_gen_emit_nodes(const int64_t* ids, int64_t n, const char* slots, const uint8_t* len8,
                       const uint64_t* tile_off, char* out, cfe_stream_t s);
int cfe_gen_elem_lengths(const int64_t* cells, const int64_t* order, const int* group, const int64_t* gfirst,
                         const int* hdr_off, const int* hdr_len, int64_t n, uint16_t* len16, uint32_t* tile_tot,
                         cfe_stream_t s);
int cfe_gen_emit_elements(const int64_t* cells, const int64_t* order, const int* group, const int64_t* gfirst,
                          const int* hdr_off, const int* hdr_len, const char* lit, int64_t n, const uint16_t* len16,
                          const uint64_t* tile_off, char* out, cfe_stream_t s);

/* ------------------------------------------------------------------ grey values */

/* hist256[g] = number of elements (v > thr) with grey value g: np.unique(cell_GV), voxelFE.py:570 */
int cfe_gv_histogram(const uint8_t* vol, int64_t n_vox, int thr, uint64_t* hist256, cfe_stream_t s);

/* stable partition of the element list by grey value: np.where(cell_GV == GV), voxelFE.py:645.
 * perm[k] = {local voxel index, raster rank} (uint2) of the k-th element in GV-major order.
 * gv8 (optional) = grey values in element order (cfe_fill_elements_gv); without it the kernels gather
 * vol[elem_vox[j]].  ws_src / ws_dst (optional, both or neither) = element workspaces
 * (cfe_emit_elements_workspace_bytes): the line lengths cfe_fill_elements left in ws_src in raster
 * order are written to ws_dst in GV-major order, so that cfe_element_lengths(perm, have_lengths = 1)
 * and cfe_emit_elements(perm) can run on ws_dst.  ws_part: cfe_gv_partition_workspace_bytes(n_el). */
size_t cfe_gv_partition_workspace_bytes(int64_t n_el);
int cfe_gv_partition(const uint8_t* vol, const uint32_t* elem_vox, int64_t n_el, const uint8_t* gv8,
                     const void* ws_src, void* ws_dst, void* perm, void* ws_part, cfe_stream_t s);

/* ------------------------------------------------------------------ ParOSol deck, segmentation */

/* vol2h5ParOSol (voxelFE.py:285-472), the O(N^3) parts.  cfe_proj_bits: the max-projections behind the
 * bounding box (recon_utils.py:258-272, bbox) from the packed bits of `rows` voxel rows of X voxels:
 * rowany[row] = 1 when the row has a set bit, colbits (ceil(X/32) words) = OR of all rows.
 * cfe_scale_crop_f32: Image dataset (voxelFE.py:350-357): out[z][y][x] = lut256[vol[z0+z][y0+y][x0+x]],
 * lut256 = (arange(256) * young_modulus).astype(float32) computed by the caller with NumPy. */
int cfe_proj_bits(const uint32_t* bits, int64_t rows, int64_t X, uint8_t* rowany, uint32_t* colbits, cfe_stream_t s);
int cfe_scale_crop_f32(const uint8_t* vol, int64_t Y, int64_t X, int64_t z0, int64_t y0, int64_t x0,
                       int64_t nz, int64_t ny, int64_t nx, const float* lut256, float* out, cfe_stream_t s);

/* HOST function (no device work): the distinct rows of `rows` (n x 4 int64, 0 <= v < 2^61 - 1, insertion
 * order) in the order CPython's list(set(map(tuple, rows))) returns them -- the order in which the reference
 * writes Fixed_Displacement_Coordinates (voxelFE.py:449).  out: capacity n x 4. */
int cfe_host_pyset_order(const int64_t* rows, int64_t n, int64_t* out, int64_t* n_out);

/* segment (preprocess.py:20-46): out[i] = img[i] > t, bytes 0 / 1; buffers 16-byte aligned */
int cfe_threshold_u8(const uint8_t* img, int64_t n, int t, uint8_t* out, cfe_stream_t s);

/* embed (preprocess.py:233-355).  cfe_box_andnot_u8: out = 1 where (z, y, x) is inside the box
 * [z0,z1) x [y0,y1) x [x0,x1) and img == 0 (BW_embedding & ~BW_I, :327), else 0.  cfe_masked_set_u8:
 * out[i] = mask[i] ? val : img[i] (I[BW_embedding] = embed_val, :331-337); out may alias img. */
int cfe_box_andnot_u8(const uint8_t* img, int64_t Z, int64_t Y, int64_t X, int64_t z0, int64_t z1, int64_t y0,
                      int64_t y1, int64_t x0, int64_t x1, uint8_t* out, cfe_stream_t s);
int cfe_masked_set_u8(const uint8_t* img, const uint8_t* mask, int64_t n, int val, uint8_t* out, cfe_stream_t s);

/* grey-scale branch of the pipeline (pipeline.py:240-242: `masked = I; masked[~L] = 0`; example 01 meshes `masked`):
 * out[i] = mask[i] ? img[i] : 0, mask bytes 0 / 1; buffers 16-byte aligned; out may alias img. */
int cfe_masked_keep_u8(const uint8_t* img, const uint8_t* mask, int64_t n, uint8_t* out, cfe_stream_t s);

/* periosteummask (preprocess.py:357-420), slice-wise morphology on packed bits ([Z*Y rows][ceil(X/32)]).
 * cfe_disk_dilate_bits: every (Y, X) slice dilated by skimage's disk(r) (0 outside the slice); invert_in
 * dilates the complement, so erosion with border value 1 is the complement of the result and the background
 * of binary_closing(x, disk(r)) is  dilate(invert_in = 1) of dilate(x).  r <= 31; in != out.
 * cfe_border_flags: for scipy.ndimage.binary_fill_holes -- flag[root] = 0 for every run that touches the
 * border of its slice (bits / pref / parent of the labelled background, labelled as ONE slice of Z*Y rows). */
int cfe_disk_dilate_bits(const uint32_t* in, int64_t Z, int64_t Y, int64_t X, int r, int invert_in, uint32_t* out,
                         cfe_stream_t s);
int cfe_border_flags(const uint32_t* bits, const uint32_t* pref, int64_t Z, int64_t Y, int64_t X,
                     const uint32_t* parent, uint32_t* flag, cfe_stream_t s);

/* ------------------------------------------------------------------ Gaussian smoothing (pipeline.py:146-148) */

/* One pass of scipy.ndimage.correlate1d(input, weights, axis, mode='nearest') with a symmetric kernel of
 * 2 r + 1 float64 weights (device pointer), float64 arithmetic in scipy's order of operations (bit-exact):
 * skimage.filters.gaussian(I, sigma, preserve_range=True) is three such passes (axes 0, 1, 2).  n_total
 * elements; the filtered axis has n_axis entries `stride` elements apart.  in: uint8 (in_is_u8) or float64. */
int cfe_gauss1d(const void* in, int in_is_u8, double* out, int64_t n_total, int64_t n_axis, int64_t stride,
                const double* weights, int r, cfe_stream_t s);

/* image > t for a float64 image: bytes 0 / 1 (preprocess.py:46 after smoothing) */
int cfe_threshold_f64(const double* img, int64_t n, double t, uint8_t* out, cfe_stream_t s);

/* One pass of a separable binary dilation on packed bits: out[q] = OR of src[q + e], e = -a .. +b along `axis`
 * (0 slices, 1 rows, 2 columns; 0 outside the image), src = in or (invert_in) its complement; invert_out
 * complements the result.  Three passes dilate by a cube, three more (on the complement) erode:
 * `morphology.binary_closing(perimask, morphology.cube(closevoxels))`, src/ciclope/utils/preprocess.py:408. */
int cfe_line_or_bits(const uint32_t* in, int64_t Z, int64_t Y, int64_t X, int axis, int a, int b, int invert_in,
                     int invert_out, uint32_t* out, cfe_stream_t s);
/* out[v] = bit v of a packed volume (the inverse of cfe_pack_bits), 0 / 1 bytes */
int cfe_unpack_bits(const uint32_t* bits, int64_t rows, int64_t X, uint8_t* out, cfe_stream_t s);

/* Spline resampling, `ndimage.zoom(image, 1 / resampling_factor, order=2)` of
 * src/ciclope/utils/preprocess.py:49-75 (pipeline.py:150-153).  cfe_spline_prefilter: in-place quadratic
 * B-spline pre-filter of the float64 array [outer][n][inner] along the axis of length n (mirror boundary,
 * what scipy uses for mode='constant').  cfe_spline_zoom: 3 x 3 x 3 taps per output sample from per-axis
 * tables (input indices idx [n_out][3], weights w [n_out][3], out-of-range flags zero [n_out]) computed by
 * the host facade; out = float64 [Zo][Yo][Xo], or uint8 with scipy's rounding (out_u8).
 * Tolerance (not bit-exact): 1e-12 absolute on float64 images of 8-bit range. */
int cfe_spline_prefilter(double* data, int64_t outer, int64_t n, int64_t inner, cfe_stream_t s);
int cfe_spline_zoom(const double* coef, int64_t Yi, int64_t Xi, int64_t Zo, int64_t Yo, int64_t Xo, const int* iz,
                    const int* iy, const int* ix, const double* wz, const double* wy, const double* wx,
                    const uint8_t* zz, const uint8_t* zy, const uint8_t* zx, int out_u8, void* out,
                    cfe_stream_t s);

/* ------------------------------------------------------------------ z-slab collectives over peer memory */

/* All-gather / neighbour shift of a small payload as one kernel over NVLink peer memory (csrc/peer.cu).
 * peer_bufs_dev / peer_sigs_dev: device arrays [world] of the ranks' symmetric buffers and signal pads
 * (torch.distributed._symmetric_memory: buffer_ptrs_dev, signal_pad_ptrs_dev).  Buffer layout
 * [slot][source rank][chunk_cap]; slot_off = byte offset of the slot used by this epoch; sig_base = first
 * u32 of the signal pad used by this slot (world words).  Every rank calls with the same epoch; out =
 * [world][nbytes].  send_mask: bit p = push my chunk to rank p and signal it; wait_mask: bit p = wait for
 * rank p and gather its chunk (all-gather: both full; shift up: send to rank + 1, wait for rank - 1 -- only
 * neighbours synchronise).  counter: 2 zeroed u32 of local device memory. */
int cfe_peer_exchange(const void* src, int64_t nbytes, const void* peer_bufs_dev, const void* peer_sigs_dev,
                      int rank, int world, int64_t slot_off, int64_t chunk_cap, uint32_t epoch, uint32_t send_mask,
                      uint32_t wait_mask, uint32_t sig_base, uint32_t* counter, void* out, cfe_stream_t s);

#ifdef __cplusplus
}
#endif
#endif
