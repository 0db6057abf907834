// ciclope_b200 -- 3-D 6-connected component labelling, keep the largest component (sm_100a).
//
// Replaces skimage.measure.label(bw, None, True, 1) + np.bincount + argmax + (labels == id)
// of ciclope/utils/preprocess.py:119-143.
//
// B200-first formulation: the volume is bit-packed (1 bit / voxel, grid.cu::cfe_pack_bits) and
// the union-find runs over x-RUNS (maximal runs of foreground voxels in a row), not voxels.
// A run's id is its raster rank (exclusive scan of run starts, cfe_scan_run_starts), so
// "smaller id" == "first voxel comes earlier in C order" and union-by-min-id keeps as root the
// run holding the component's first voxel: exactly the numbering rule that breaks size ties
// in the reference (first maximum of np.bincount over labels numbered by first appearance).
// Traffic: 1 B/voxel read once + 1 B/voxel mask written once + ~1 B/voxel of bit/prefix words;
// the parent/size arrays (4 B per run, ~3 % of the voxels) live in the 126 MB L2.
#include "common.cuh"
#include "ciclope_b200.h"

// run id of the run covering bit b of word i (word value `w`, `carry` = bit 31 of the previous
// word of the same row, `pref` = exclusive count of run starts before word i).  When the run
// started in an earlier word the count of starts <= b inside this word is 0 and the id is the
// last run started before the word.
__device__ __forceinline__ u32 run_id(u32 w, u32 carry, u32 pref, int b)
{
    const u32 starts = w & ~((w << 1) | carry);
    return pref + (u32)__popc(starts & mask_le(b)) - 1u;
}

__device__ __forceinline__ u32 uf_find(u32* parent, u32 x)
{
    // path halving; concurrent writers only ever store smaller ancestors
    u32 p = parent[x];
    while (p != x) {
        const u32 gp = parent[p];
        if (gp != p) parent[x] = gp;
        x = p;
        p = gp;
    }
    return x;
}

__device__ __forceinline__ void uf_union(u32* parent, u32 a, u32 b)
{
    while (true) {
        a = uf_find(parent, a);
        b = uf_find(parent, b);
        if (a == b) return;
        if (a < b) { const u32 t = a; a = b; b = t; }  // a > b: hook a under b
        const u32 old = atomicMin(parent + a, b);
        if (old == a) return;
        a = old;  // someone hooked a first: merge their target with b
    }
}

// ------------------------------------------------------------------------------------------
// Labelling chain (cfe_ccl_label): three passes, every one of them parallel over its items.
//
//   k_ccl_tile     one CTA per tile of TS slices x TR rows (full width): the tile's runs get LOCAL indices in
//                  raster order (so min local = min global), every overlapping run pair inside the tile is
//                  first written to a shared-memory queue (cheap, divergent bit loops) and then united by
//                  all threads, one pair each (the expensive part runs dense), in a shared-memory union-find.
//                  The tile's forest is flattened, the voxels of every tile-local component are counted
//                  (shared-memory adds aggregated per warp by match.any) and parent[run] = tile root (as a
//                  global id), size[tile root] = voxels, size[other runs] = 0 are published.  Tiles with more
//                  runs / pairs than the shared-memory capacity publish identity parents + run lengths and
//                  leave their pairs to k_ccl_inner (flag 0).
//   k_ccl_border   the run pairs that cross a tile border (rows r % TR == 0 against r - 1, slices
//                  s % TS == 0 against s - 1): union-find in HBM / L2, atomicMin hooking under the smaller id.
//   k_ccl_finish   per run: parent[i] = final root; a tile root that is not the final root adds its voxel
//                  count to the final root's (warp-aggregated), so size[root] = voxels of the component.
#ifndef CCL_TILE_LOG2
#define CCL_TILE_LOG2 16        // voxels per tile = 2^16 (measured: 2^15 .. 2^17 within 12 %, 2^16 best at 2048^2 rows)
#endif
#ifndef CCL_MINBLOCKS
#define CCL_MINBLOCKS 4         // CTAs per SM the tile kernel is compiled for (shared memory: 50 KB each)
#endif
#define CCL_CAP (1 << (CCL_TILE_LOG2 - 4))            // runs per tile held in shared memory (1 per 16 voxels)
#define CCL_WCAP ((1 << (CCL_TILE_LOG2 - 5)) + 256)   // words per tile (+ one pad word per row)
#define CCL_MAXROWS 256         // TS * TR
#define CCL_THREADS 512
#define CCL_SMEM ((2 * CCL_CAP + 2 * CCL_WCAP) * 4)

struct CclGeom {
    int Z, Y, W;
    int TS, TR, tiles_s, tiles_r;
    FastDiv dW;
};

// global pair kernel body: unions of word i with word j (same column word in the row above / slice below)
// `seen` (may be NULL): a small direct-mapped table in shared memory of the (root, root) pairs this CTA has
// already united.  After k_ccl_tile a run's parent is its tile root, a tile has a handful of roots, and all the
// words of a border row join the same two tiles: the ~10^7 border pairs of a slab connect only ~10^4..10^5
// distinct root pairs, and a pair that was united once does not need its pointer chases again.  (A second,
// device-wide table shared by all CTAs was tried and lost: 1.12 ms against 0.57 ms for this kernel -- the fence
// and the extra L2 round trip per new pair cost more than the repeated unions they save.)
#define CCL_SEEN 256
__device__ __forceinline__ void ccl_pairs_global(const u32* __restrict__ bits, const u32* __restrict__ pref, i64 i,
                                                 i64 j, bool has_prev, u32 cur, u32* parent, u64* seen = nullptr)
{
    const u32 nb = __ldg(bits + j);
    const u32 both = cur & nb;
    if (!both) return;
    const u32 cur_prev = has_prev ? __ldg(bits + i - 1) : 0u;
    const u32 nb_prev = has_prev ? __ldg(bits + j - 1) : 0u;
    const u32 cur_pref = __ldg(pref + i), nb_pref = __ldg(pref + j);
    u32 pair_starts = both & ~((both << 1) | ((cur_prev & nb_prev) >> 31));
    while (pair_starts) {
        const int b = __ffs((int)pair_starts) - 1;
        pair_starts &= pair_starts - 1;
        const u32 ra = run_id(cur, cur_prev >> 31, cur_pref, b), rb = run_id(nb, nb_prev >> 31, nb_pref, b);
        if (!seen) { uf_union(parent, ra, rb); continue; }
        const u32 ta = parent[ra], tb = parent[rb];            // one hop each: the tile roots (or closer ancestors)
        if (ta == tb) continue;
        const u64 key = ta > tb ? (((u64)ta << 32) | tb) : (((u64)tb << 32) | ta);
        const u32 slot = ((ta * 0x9E3779B1u) ^ (tb * 0x85EBCA77u) ^ ((ta ^ tb) >> 15)) & (CCL_SEEN - 1);
        if (seen[slot] == key) continue;
        uf_union(parent, ta, tb);
        seen[slot] = key;
    }
}

__global__ void __launch_bounds__(CCL_THREADS, CCL_MINBLOCKS)
k_ccl_tile(const u32* __restrict__ bits, const u32* __restrict__ pref, const u64* __restrict__ n_runs_p, CclGeom G,
           u32* __restrict__ parent, u32* __restrict__ size, u32* __restrict__ tile_flag)
{
    extern __shared__ __align__(16) u32 dyn[];
    u32* const sp = dyn;                        // [CCL_CAP] local union-find
    u32* const cnt_l = dyn + CCL_CAP;           // [CCL_CAP] voxels per run, then per tile-local component
    u32* const sb = dyn + 2 * CCL_CAP;          // [CCL_WCAP] the tile's bit words (tile row major)
    u32* const spf = sb + CCL_WCAP;             // [CCL_WCAP] LOCAL run-start prefix of every word
    __shared__ u32 rb_g[CCL_MAXROWS + 1];       // global run id at the start of each tile row
    __shared__ u32 rb_l[CCL_MAXROWS + 1];       // local run index at the start of each tile row
    __shared__ u32 s_scan[33];
    const int W = G.W, TR = G.TR, n_trows = G.TS * G.TR;
    const int tile_s = blockIdx.x / G.tiles_r, tile_r = blockIdx.x - tile_s * G.tiles_r;
    const u32 t = threadIdx.x, lane = t & 31u;
    const u32 tile_words = (u32)n_trows * (u32)W;
    // ---- stage the tile's bit words and (still global) run prefixes right away: these loads and the row-start
    // loads of the numbering below are then in flight together (one round of HBM latency per tile, not two)
    if (tile_words <= CCL_WCAP) {
        for (u32 wt = t; wt < tile_words; wt += CCL_THREADS) {
            const u32 tr = fd_div(wt, G.dW), w = wt - tr * (u32)W;
            const int s = tile_s * G.TS + (int)tr / TR, r = tile_r * TR + ((int)tr % TR);
            u32 word = 0, pf = 0;
            if (s < G.Z && r < G.Y) {
                const i64 i = ((i64)s * G.Y + r) * W + w;
                word = __ldg(bits + i);
                pf = __ldg(pref + i);
            }
            sb[wt] = word;
            spf[wt] = pf;
        }
    }
    u32 g0 = 0, cnt = 0;
    if (t < (u32)n_trows) {
        // tile row t  <->  slice s0 + t / TR, row r0 + t % TR
        const int s = tile_s * G.TS + (int)t / TR, r = tile_r * TR + ((int)t % TR);
        const i64 row = (i64)s * G.Y + r;
        const i64 n_rows = (i64)G.Z * G.Y;
        if (s < G.Z && r < G.Y) {
            g0 = __ldg(pref + row * W);
            const u32 g1 = (row + 1 < n_rows) ? __ldg(pref + (row + 1) * W) : (u32)(*n_runs_p);
            cnt = g1 - g0;
        }
        rb_g[t] = g0;
    }
    u32 total;
    const u32 l0 = block_exclusive_scan<u32>(cnt, s_scan, total);
    if (t < (u32)n_trows) rb_l[t] = l0;
    if (t == 0) rb_l[n_trows] = total;
    __syncthreads();
    if (total == 0) {                            // uniform for the CTA
        if (t == 0) tile_flag[blockIdx.x] = 1u;
        return;
    }
    if (total > CCL_CAP || tile_words > CCL_WCAP) {
        // over capacity: identity parents, run lengths as sizes; k_ccl_inner unites the tile's pairs in HBM
        if (t == 0) tile_flag[blockIdx.x] = 0u;
        for (u32 tr = 0; tr < (u32)n_trows; ++tr) {
            const u32 gb = rb_g[tr], n = rb_l[tr + 1] - rb_l[tr];
            for (u32 k = t; k < n; k += CCL_THREADS) { parent[gb + k] = gb + k; size[gb + k] = 0u; }
        }
        __syncthreads();
        for (u32 wt = t; wt < tile_words; wt += CCL_THREADS) {
            const u32 tr = fd_div(wt, G.dW), w = wt - tr * (u32)W;
            const int s = tile_s * G.TS + (int)tr / TR, r = tile_r * TR + ((int)tr % TR);
            if (s >= G.Z || r >= G.Y) continue;
            const i64 i = ((i64)s * G.Y + r) * W + w;
            u32 word = __ldg(bits + i);
            if (!word) continue;
            const u32 orig = word, carry = w ? (__ldg(bits + i - 1) >> 31) : 0u, pf = __ldg(pref + i);
            while (word) {
                const int b = __ffs((int)word) - 1;
                const u32 shifted = word >> b;
                const int len = (~shifted == 0u) ? 32 - b : __ffs((int)~shifted) - 1;
                atomicAdd(size + run_id(orig, carry, pf, b), (u32)len);
                word = (b + len >= 32) ? 0u : (word & ~(((1u << len) - 1u) << b));
            }
        }
        return;
    }
    if (t == 0) tile_flag[blockIdx.x] = 1u;
    // ---- global -> LOCAL run prefixes (local = global - first global id of the row + first local index of the row)
    for (u32 l = t; l < total; l += CCL_THREADS) { sp[l] = l; cnt_l[l] = 0u; }
    for (u32 wt = t; wt < tile_words; wt += CCL_THREADS) {
        const u32 tr = fd_div(wt, G.dW);
        spf[wt] += rb_l[tr] - rb_g[tr];
    }
    __syncthreads();
    // ---- one pass over the staged words: run lengths, and the union of every overlapping run pair with the
    // row above (same slice) and the same row of the slice below, in the shared-memory forest
    const u32 slice_stride = (u32)TR * (u32)W;
    for (u32 wt = t; wt < tile_words; wt += CCL_THREADS) {
        const u32 cur = sb[wt];
        if (!cur) continue;
        const u32 tr = fd_div(wt, G.dW), w = wt - tr * (u32)W;
        const u32 cur_prev = w ? sb[wt - 1] : 0u;
        const u32 cur_pref = spf[wt];
        const u32 carry = cur_prev >> 31;
        {   // voxels of every run: a run's pieces in different words add up in its own counter
            u32 rest = cur;
            while (rest) {
                const int b = __ffs((int)rest) - 1;
                const u32 shifted = rest >> b;
                const int len = (~shifted == 0u) ? 32 - b : __ffs((int)~shifted) - 1;
                CFE_ASSERT(run_id(cur, carry, cur_pref, b) < total);
                atomicAdd(cnt_l + run_id(cur, carry, cur_pref, b), (u32)len);
                rest = (b + len >= 32) ? 0u : (rest & ~(((1u << len) - 1u) << b));
            }
        }
#pragma unroll
        for (int dir = 0; dir < 2; ++dir) {
            if (dir == 0 ? (tr % (u32)TR) == 0 : tr < (u32)TR) continue;
            const u32 j = dir == 0 ? wt - (u32)W : wt - slice_stride;
            const u32 nb = sb[j];
            const u32 both = cur & nb;
            if (!both) continue;
            const u32 nb_prev = w ? sb[j - 1] : 0u;
            const u32 nb_pref = spf[j];
            u32 pair_starts = both & ~((both << 1) | ((cur_prev & nb_prev) >> 31));
            while (pair_starts) {
                const int b = __ffs((int)pair_starts) - 1;
                pair_starts &= pair_starts - 1;
                CFE_ASSERT(run_id(cur, carry, cur_pref, b) < total && run_id(nb, nb_prev >> 31, nb_pref, b) < total);
                uf_union(sp, run_id(cur, carry, cur_pref, b), run_id(nb, nb_prev >> 31, nb_pref, b));
            }
        }
    }
    __syncthreads();
    // ---- flatten (in place: concurrent readers only ever see ancestors)
    for (u32 l = t; l < total; l += CCL_THREADS) {
        u32 x = l, p;
        while ((p = sp[x]) != x) x = p;
        sp[l] = x;
    }
    __syncthreads();
    // ---- voxels per tile-local component: every non-root run adds its length to its root's counter.  Runs that
    // follow each other mostly share the root, so the warp adds up every maximal stretch of lanes with the same
    // root (head flags -> one inclusive scan -> difference of two prefix values) and the last lane of a stretch
    // issues one shared-memory add
    for (u32 base = 0; base < total; base += CCL_THREADS) {
        const u32 l = base + t;
        u32 root = 0xffffffffu, add = 0;
        if (l < total) {
            const u32 r = sp[l];
            if (r != l) { root = r; add = cnt_l[l]; }
        }
        const u32 prev_root = __shfl_up_sync(CFE_FULL_MASK, root, 1);
        const u32 heads = __ballot_sync(CFE_FULL_MASK, lane == 0 || prev_root != root);
        u32 incl = add;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const u32 v = __shfl_up_sync(CFE_FULL_MASK, incl, o);
            if (lane >= (u32)o) incl += v;
        }
        const u32 start = 31u - (u32)__clz((int)(heads & mask_le((int)lane)));     // head lane of my stretch
        const u32 before = __shfl_sync(CFE_FULL_MASK, incl, start ? start - 1u : 0u);
        const bool tail = lane == 31u || ((heads >> (lane + 1u)) & 1u);
        if (tail && root != 0xffffffffu) atomicAdd(cnt_l + root, incl - (start ? before : 0u));
    }
    __syncthreads();
    // ---- publish: the rows of one slice of the tile are consecutive in raster order, so local indices map to
    // global run ids piecewise linearly with one piece per slice (<= 16 pieces)
    for (u32 l = t; l < total; l += CCL_THREADS) {
        const u32 root = sp[l];
        int ka = 0, kb = 0;                               // slice of the tile holding l / root (TS is a power of two)
#pragma unroll 1
        for (int step = G.TS >> 1; step > 0; step >>= 1) {
            if (rb_l[(ka + step) * TR] <= l) ka += step;
            if (rb_l[(kb + step) * TR] <= root) kb += step;
        }
        const u32 g = rb_g[ka * TR] + (l - rb_l[ka * TR]);
        CFE_ASSERT(root <= l && (u64)g < *n_runs_p && rb_g[kb * TR] + (root - rb_l[kb * TR]) <= g);
        parent[g] = rb_g[kb * TR] + (root - rb_l[kb * TR]);
        size[g] = (root == l) ? cnt_l[l] : 0u;
    }
}

// Run pairs across tile borders.  Work items = words of the border rows: first the rows r = k * TR (k >= 1) of
// every slice (against row r - 1), then every row of the slices s = k * TS (k >= 1) (against slice s - 1).
// The items are ordered so that the TS rows of ONE tile border follow each other (a CTA of 512 threads = one border
// of 8 slices x 64 words at 2048^2): the rows of a border mostly join the same two tile roots, and the CTA's table
// of pairs already united then catches the repeats (in slice-major order every CTA re-united them: 0.57 ms).
#define CCL_BORDER_THREADS 512
__global__ void __launch_bounds__(CCL_BORDER_THREADS)
k_ccl_border(const u32* __restrict__ bits, const u32* __restrict__ pref, CclGeom G, i64 n0, i64 n_items,
             u32* __restrict__ parent)
{
    __shared__ u64 s_seen[CCL_SEEN];
    for (u32 k = threadIdx.x; k < CCL_SEEN; k += blockDim.x) s_seen[k] = ~0ull;
    __syncthreads();
    const i64 it = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (it >= n_items) return;
    const int W = G.W;
    i64 row;        // global row index s * Y + r
    int dir;
    i64 q = it / W;
    const int w = (int)(it - q * W);
    if (it < n0) {
        // q = ((tile_s * (tiles_r - 1)) + (k - 1)) * TS + slice of the tile
        const int per_slice = G.tiles_r - 1;
        const i64 b = q / G.TS;
        const int ds = (int)(q - b * G.TS);
        const i64 ts = b / per_slice;
        const int r = ((int)(b - ts * per_slice) + 1) * G.TR;
        const i64 s = ts * G.TS + ds;
        if (s >= G.Z) return;
        row = s * G.Y + r;
        dir = 0;
    } else {
        q -= n0 / W;
        const i64 k = q / G.Y;
        const int r = (int)(q - k * G.Y);
        row = (k + 1) * G.TS * (i64)G.Y + r;
        dir = 1;
    }
    CFE_ASSERT(row >= (dir == 0 ? 1 : G.Y) && row < (i64)G.Z * G.Y);
    const i64 i = row * W + w;
    const u32 cur = __ldg(bits + i);
    if (!cur) return;
    ccl_pairs_global(bits, pref, i, dir == 0 ? i - W : i - (i64)W * G.Y, w != 0, cur, parent, s_seen);
}

// The pairs INSIDE the tiles that did not fit shared memory (flag 0): one CTA per tile, returns at once for
// the others.
__global__ void __launch_bounds__(256)
k_ccl_inner(const u32* __restrict__ bits, const u32* __restrict__ pref, CclGeom G, u32* __restrict__ parent,
            const u32* __restrict__ tile_flag)
{
    if (tile_flag[blockIdx.x]) return;
    const int tile_s = blockIdx.x / G.tiles_r, tile_r = blockIdx.x - tile_s * G.tiles_r;
    const int W = G.W, n_trows = G.TS * G.TR;
    const u32 tile_words = (u32)n_trows * (u32)W;
    for (u32 wt = threadIdx.x; wt < tile_words; wt += blockDim.x) {
        const u32 tr = fd_div(wt, G.dW), w = wt - tr * (u32)W;
        const int s = tile_s * G.TS + (int)tr / G.TR, r = tile_r * G.TR + ((int)tr % G.TR);
        if (s >= G.Z || r >= G.Y) continue;
        const i64 i = ((i64)s * G.Y + r) * W + w;
        const u32 cur = __ldg(bits + i);
        if (!cur) continue;
        if ((tr % (u32)G.TR) != 0) ccl_pairs_global(bits, pref, i, i - W, w != 0, cur, parent);
        if (tr >= (u32)G.TR) ccl_pairs_global(bits, pref, i, i - (i64)W * G.Y, w != 0, cur, parent);
    }
}

__global__ void __launch_bounds__(256)
k_ccl_finish(u32* __restrict__ parent, u32* __restrict__ size, const u64* __restrict__ n_runs)
{
    const u64 n = *n_runs;
    const u32 lane = threadIdx.x & 31u;
    for (u64 base = (u64)blockIdx.x * blockDim.x + threadIdx.x - lane; base < n; base += (u64)gridDim.x * blockDim.x) {
        const u64 i = base + lane;
        u32 r = 0xffffffffu, add = 0;
        if (i < n) {
            const u32 p = parent[i];
            u32 x = p, qn;
            while ((qn = parent[x]) != x) x = qn;
            if (x != p) parent[i] = x;
            const u32 sz = size[i];
            if (sz && x != (u32)i) { r = x; add = sz; size[i] = 0u; }
        }
        if (!__any_sync(CFE_FULL_MASK, add != 0)) continue;
        const u32 peers = __match_any_sync(CFE_FULL_MASK, r);
        const u32 sum = __reduce_add_sync(peers, add);
        if (r != 0xffffffffu && ((u32)__ffs((int)peers) - 1u) == lane) atomicAdd(size + r, sum);
    }
}

// K3b: arg-max over roots of (size, then smallest id)  ->  best = (size << 32) | ~id
__global__ void __launch_bounds__(256)
k_ccl_select(const u32* __restrict__ parent, const u32* __restrict__ size, const u64* __restrict__ n_runs,
             u64* __restrict__ best, const u32* __restrict__ exclude)
{
    __shared__ u64 s_best[8];
    const u64 n = *n_runs;
    u64 key = 0;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        if (parent[i] == (u32)i && !(exclude && exclude[i] != 0xffffffffu)) {
            const u64 k = ((u64)size[i] << 32) | (u64)(0xffffffffu - (u32)i);
            key = k > key ? k : key;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const u64 t = __shfl_xor_sync(CFE_FULL_MASK, key, o);
        key = t > key ? t : key;
    }
    if ((threadIdx.x & 31) == 0) s_best[threadIdx.x >> 5] = key;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) key = s_best[w] > key ? s_best[w] : key;
        if (key) atomicMax(best, key);
    }
}

// K3c: write the byte mask (labels == largest_label_id).  One thread per 16 voxels -> one
// 16-byte store; FLAT requires X % 32 == 0 and a 16-byte aligned output.
__device__ __forceinline__ u32 spread_nibble(u32 n) { return ((n & 15u) * 0x00204081u) & 0x01010101u; }

template <bool FLAT>
__global__ void __launch_bounds__(256)
k_ccl_write_mask(const u32* __restrict__ bits, const u32* __restrict__ pref, i64 rows, int X, int W,
                 const u32* __restrict__ parent, const u64* __restrict__ best, const u32* __restrict__ keep_flag,
                 int drop, const uint8_t* src, u32 fill, uint8_t* out)
{
    const i64 g = (i64)blockIdx.x * blockDim.x + threadIdx.x;  // 16-voxel chunk
    if (g >= rows * 2 * (i64)W) return;
    const u64 key = keep_flag ? 1ull : *best;
    const u32 winner = 0xffffffffu - (u32)(key & 0xffffffffu);
    const i64 i = g >> 1;
    const int half = (int)(g & 1);
    const u32 word = __ldg(bits + i);
    u32 keep = 0;
    u32 rest = word & (half ? 0xffff0000u : 0x0000ffffu);
    if (rest && key) {
        const u32 carry = (i % W) ? (__ldg(bits + i - 1) >> 31) : 0u;
        const u32 pf = __ldg(pref + i);
        while (rest) {
            const int b = __ffs((int)rest) - 1;
            const u32 shifted = rest >> b;
            const int len = (~shifted == 0u) ? 32 - b : __ffs((int)~shifted) - 1;
            const u32 seg = (len >= 32) ? 0xffffffffu : (((1u << len) - 1u) << b);
            const u32 root = __ldg(parent + run_id(word, carry, pf, b));
            if ((keep_flag ? (keep_flag[root] == 0xffffffffu) : (root == winner)) != (drop != 0)) keep |= seg;
            rest &= ~seg;
        }
    }
    const u32 k16 = (keep >> (16 * half)) & 0xffffu;
    uint4 v;
    v.x = spread_nibble(k16);
    v.y = spread_nibble(k16 >> 4);
    v.z = spread_nibble(k16 >> 8);
    v.w = spread_nibble(k16 >> 12);
    // selected voxels get `fill`, the others 0 or (fill_voids) their value in `src`; out may be src
    const u32 f4 = (fill & 0xffu) * 0x01010101u;
    if (FLAT) {
        uint4 o = make_uint4(0u, 0u, 0u, 0u);
        if (src) o = reinterpret_cast<const uint4*>(src)[g];
        const u32 mx = v.x * 0xffu, my = v.y * 0xffu, mz = v.z * 0xffu, mw = v.w * 0xffu;
        o.x = (o.x & ~mx) | (f4 & mx);
        o.y = (o.y & ~my) | (f4 & my);
        o.z = (o.z & ~mz) | (f4 & mz);
        o.w = (o.w & ~mw) | (f4 & mw);
        reinterpret_cast<uint4*>(out)[g] = o;
    } else {
        const i64 row = i / W;
        const int c0 = (int)(i - row * W) * 32 + 16 * half;
        uint8_t* p = out + row * (i64)X + c0;
        const uint8_t* q = src ? src + row * (i64)X + c0 : nullptr;
        const u32 vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 16; ++k)
            if (c0 + k < X)
                p[k] = ((vv[k >> 2] >> (8 * (k & 3))) & 1u) ? (uint8_t)fill : (q ? q[k] : (uint8_t)0);
    }
}

// Flat volumes (X % 32 == 0, 16-byte aligned): one thread per WORD walks the word's runs once (the
// 16-voxel version above visits a run that crosses the half-word twice), then the warp redistributes
// the 32 keep-words so that every store instruction writes 512 contiguous bytes.
__global__ void __launch_bounds__(256)
k_ccl_write_mask_words(const u32* __restrict__ bits, const u32* __restrict__ pref, i64 n_words, int W,
                       const u32* __restrict__ parent, const u64* __restrict__ best,
                       const u32* __restrict__ keep_flag, int drop, const uint8_t* src, u32 fill, uint8_t* out,
                       u32* __restrict__ sel_bits)
{
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    const u32 lane = threadIdx.x & 31u;
    const u64 key = keep_flag ? 1ull : *best;
    const u32 winner = 0xffffffffu - (u32)(key & 0xffffffffu);
    u32 keep = 0;
    if (i < n_words && key) {
        // the three loads go out together (the pass is bound by its chain of dependent loads: word -> prefix ->
        // parent; fetching prefix and carry only for non-empty words added a round of latency to every word)
        const u32 word = __ldg(bits + i);
        const u32 prev = (i % W) ? __ldg(bits + i - 1) : 0u;
        const u32 pf = __ldg(pref + i);
        if (word) {
            const u32 carry = prev >> 31;
            u32 rest = word;
            while (rest) {
                const int b = __ffs((int)rest) - 1;
                const u32 shifted = rest >> b;
                const int len = (~shifted == 0u) ? 32 - b : __ffs((int)~shifted) - 1;
                const u32 seg = (len >= 32) ? 0xffffffffu : (((1u << len) - 1u) << b);
                const u32 root = __ldg(parent + run_id(word, carry, pf, b));
                if ((keep_flag ? (keep_flag[root] == 0xffffffffu) : (root == winner)) != (drop != 0)) keep |= seg;
                rest &= ~seg;
            }
        }
    }
    if (sel_bits && i < n_words) sel_bits[i] = keep;     // the selection, bit-packed like `bits`
    const i64 chunk0 = (i - lane) * 2;            // first 16-byte chunk of the warp's 32 words
    const u32 f4 = (fill & 0xffu) * 0x01010101u;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const u32 c = (u32)k * 32u + lane;        // chunk of the warp handled by this lane
        const u32 kw = __shfl_sync(CFE_FULL_MASK, keep, c >> 1);
        const i64 g = chunk0 + c;
        if (g >= 2 * n_words) continue;
        const u32 k16 = (kw >> (16u * (c & 1u))) & 0xffffu;
        const u32 mx = spread_nibble(k16) * 0xffu, my = spread_nibble(k16 >> 4) * 0xffu,
                  mz = spread_nibble(k16 >> 8) * 0xffu, mw = spread_nibble(k16 >> 12) * 0xffu;
        uint4 o = make_uint4(0u, 0u, 0u, 0u);
        if (src) o = reinterpret_cast<const uint4*>(src)[g];
        o.x = (o.x & ~mx) | (f4 & mx);
        o.y = (o.y & ~my) | (f4 & my);
        o.z = (o.z & ~mz) | (f4 & mz);
        o.w = (o.w & ~mw) | (f4 & mw);
        reinterpret_cast<uint4*>(out)[g] = o;
    }
}

extern "C" size_t cfe_ccl_max_runs(int64_t rows, int64_t X)
{
    // n_runs <= rows * ceil(X/2); the tail holds one flag per block-local CCL tile
    // (ceil(Z/8) * ceil(Y/32) <= rows/256 + rows/8 + rows/32 + 1 for any factorisation rows = Z*Y)
    return (size_t)(rows * ((X + 1) / 2)) + (size_t)(rows / 4 + 8);
}

static unsigned grid_stride_blocks()
{
    int sms = 148, dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return (unsigned)(sms * 8);  // 8 CTAs of 256 threads per SM
}

// Tile shape of k_ccl_tile: about 2^16 voxels per tile (2 * 10^3 runs at the run density of trabecular
// volumes, half the shared-memory capacity), as square as possible in (slices, rows) because the pairs that
// cross a tile border (1 / TS + 1 / TR of them) are united in HBM.
static void ccl_geometry(CclGeom& G, i64 Z, i64 Y, i64 X)
{
    G.Z = (int)Z; G.Y = (int)Y; G.W = (int)((X + 31) / 32);
    G.dW = make_fastdiv((u32)G.W);
    int rows = 8;
    while (rows < CCL_MAXROWS && (i64)(rows * 2) * X <= (1 << CCL_TILE_LOG2)) rows *= 2;
    int ts = 1;
    while (ts * ts < rows) ts *= 2;          // 8 -> 4 x 2, 16 -> 4 x 4, 32 -> 8 x 4, 64 -> 8 x 8, 128 -> 16 x 8, 256 -> 16 x 16
    while (ts > 1 && ts / 2 >= Z) ts /= 2;   // thin slabs: spend the rows on y
    int tr = rows / ts;
    while (tr > 1 && tr / 2 >= Y) tr /= 2;
    G.TS = ts; G.TR = tr;
    G.tiles_s = (int)((Z + ts - 1) / ts);
    G.tiles_r = (int)((Y + tr - 1) / tr);
}

// Phase 1: label the slab.  After this call parent[run] = root run id (min run id of the
// component inside this slab) and size[root] = voxels of the component inside this slab.
extern "C" int cfe_ccl_label(const uint32_t* bits, const uint32_t* pref, const uint64_t* n_runs, int64_t Z,
                             int64_t Y, int64_t X, uint32_t* parent, uint32_t* size, cudaStream_t stream)
{
    const i64 rows = Z * Y;
    if (rows <= 0 || X <= 0) return 0;
    CclGeom G;
    ccl_geometry(G, Z, Y, X);
    const u64* nr = reinterpret_cast<const u64*>(n_runs);
    // tile flags live behind the n_runs <= rows*ceil(X/2) entries of `size` (see cfe_ccl_max_runs)
    u32* tile_flag = size + (size_t)rows * (size_t)((X + 1) / 2);
    static unsigned long long attr_mask = 0;
    if (cfe_first_on_device(&attr_mask))
        CFE_CHECK_CUDA(cudaFuncSetAttribute(k_ccl_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, CCL_SMEM));
    const unsigned n_tiles = (unsigned)(G.tiles_s * G.tiles_r);
    k_ccl_tile<<<n_tiles, CCL_THREADS, CCL_SMEM, stream>>>(bits, pref, nr, G, parent, size, tile_flag);
    CFE_CHECK_LAUNCH("k_ccl_tile");
    k_ccl_inner<<<n_tiles, 256, 0, stream>>>(bits, pref, G, parent, tile_flag);
    CFE_CHECK_LAUNCH("k_ccl_inner");
    const i64 n0 = (i64)G.tiles_s * G.TS * (i64)(G.tiles_r - 1) * G.W;     // rows of the last tile layer beyond Z: skipped
    const i64 n1 = (i64)(G.tiles_s - 1) * Y * G.W;
    if (n0 + n1 > 0) {
        k_ccl_border<<<(unsigned)((n0 + n1 + CCL_BORDER_THREADS - 1) / CCL_BORDER_THREADS), CCL_BORDER_THREADS, 0, stream>>>(
            bits, pref, G, n0, n0 + n1, parent);
        CFE_CHECK_LAUNCH("k_ccl_border");
    }
    k_ccl_finish<<<grid_stride_blocks(), 256, 0, stream>>>(parent, size, nr);
    CFE_CHECK_LAUNCH("k_ccl_finish");
    return 0;
}

// Phase 2: *best = max over roots of (size << 32) | (0xffffffff - root); 0 = no foreground.
extern "C" int cfe_ccl_select(const uint32_t* parent, const uint32_t* size, const uint64_t* n_runs,
                              uint64_t* best, cudaStream_t stream)
{
    CFE_CHECK_CUDA(cudaMemsetAsync(best, 0, sizeof(u64), stream));
    k_ccl_select<<<grid_stride_blocks(), 256, 0, stream>>>(parent, size, reinterpret_cast<const u64*>(n_runs),
                                                           reinterpret_cast<u64*>(best), nullptr);
    CFE_CHECK_LAUNCH("k_ccl_select");
    return 0;
}

// Phase 3: out_mask = 1 where the voxel's run has root (0xffffffff - low32(*best)).
// With `keep` != NULL the test is keep[root] == 0xffffffff instead (z-slab sharding: the winning
// global component may be the union of several slab-local components; the marker cannot be a size).
extern "C" int cfe_ccl_write_mask(const uint32_t* bits, const uint32_t* pref, int64_t Z, int64_t Y, int64_t X,
                                  const uint32_t* parent, const uint64_t* best, const uint32_t* keep,
                                  uint8_t* out_mask, cudaStream_t stream)
{
    return cfe_ccl_write_select(bits, pref, Z, Y, X, parent, best, keep, 0, nullptr, 1, out_mask, stream);
}

// General form of phase 3.  Selected voxels = runs whose root is the winner (drop = 0) or every
// labelled voxel EXCEPT the winner's (drop = 1: remove_largest, preprocess.py:204-206; the voids of
// fill_voids, :176-177).  out = fill on selected voxels, src[voxel] (0 when src is NULL) elsewhere;
// out == src is allowed (in-place fill).
static int write_select(const uint32_t* bits, const uint32_t* pref, int64_t Z, int64_t Y, int64_t X,
                        const uint32_t* parent, const uint64_t* best, const uint32_t* keep, int drop,
                        const uint8_t* src, int fill, uint8_t* out_mask, uint32_t* sel_bits, cudaStream_t stream)
{
    const i64 rows = Z * Y;
    if (rows <= 0 || X <= 0) return 0;
    const int W = (int)((X + 31) / 32);
    const i64 chunks = rows * W * 2;
    const unsigned cb = (unsigned)((chunks + 255) / 256);
    const bool flat = (X % 32 == 0) && ((reinterpret_cast<uintptr_t>(out_mask) & 15u) == 0) &&
                      ((reinterpret_cast<uintptr_t>(src) & 15u) == 0);
    const u64* b = reinterpret_cast<const u64*>(best);
    if (sel_bits && !flat) {
        cfe_set_error("cfe_ccl_*_bits: the packed selection needs X %% 32 == 0 and 16-byte aligned buffers");
        return -1;
    }
    if (flat)
        k_ccl_write_mask_words<<<(unsigned)((rows * W + 255) / 256), 256, 0, stream>>>(
            bits, pref, rows * W, W, parent, b, keep, drop, src, (u32)fill, out_mask, sel_bits);
    else
        k_ccl_write_mask<false><<<cb, 256, 0, stream>>>(bits, pref, rows, (int)X, W, parent, b, keep, drop, src,
                                                        (u32)fill, out_mask);
    CFE_CHECK_LAUNCH("k_ccl_write_mask");
    return 0;
}

extern "C" int cfe_ccl_write_select(const uint32_t* bits, const uint32_t* pref, int64_t Z, int64_t Y, int64_t X,
                                    const uint32_t* parent, const uint64_t* best, const uint32_t* keep,
                                    int drop, const uint8_t* src, int fill, uint8_t* out_mask,
                                    cudaStream_t stream)
{
    return write_select(bits, pref, Z, Y, X, parent, best, keep, drop, src, fill, out_mask, nullptr, stream);
}

// cfe_ccl_write_mask that also leaves the kept voxels bit-packed (sel_bits: u32[Z*Y*ceil(X/32)], the layout of
// cfe_pack_bits): vol2ugrid of the mask then starts from the bits instead of re-reading the byte mask.
// Requires X % 32 == 0 and a 16-byte aligned out_mask.
extern "C" int cfe_ccl_write_mask_bits(const uint32_t* bits, const uint32_t* pref, int64_t Z, int64_t Y, int64_t X,
                                       const uint32_t* parent, const uint64_t* best, const uint32_t* keep,
                                       uint8_t* out_mask, uint32_t* sel_bits, cudaStream_t stream)
{
    return write_select(bits, pref, Z, Y, X, parent, best, keep, 0, nullptr, 1, out_mask, sel_bits, stream);
}

// Labels the slab and keeps the largest component (single-GPU path = phases 1-3 back to back).
//   bits/pref : packed foreground bits and run-start prefixes (cfe_pack_bits, cfe_scan_run_starts)
//   n_runs    : device scalar written by cfe_scan_run_starts
//   parent,size : u32[cfe_ccl_max_runs] scratch
//   best      : device u64, out: (size << 32) | (0xffffffff - root run id); 0 = no foreground
//   out_mask  : uint8 [rows*X], 1 = voxel belongs to the largest component
extern "C" int cfe_ccl_largest(const uint32_t* bits, const uint32_t* pref, const uint64_t* n_runs,
                               int64_t Z, int64_t Y, int64_t X, uint32_t* parent, uint32_t* size,
                               uint64_t* best, uint8_t* out_mask, cudaStream_t stream)
{
    return cfe_ccl_largest_select(bits, pref, n_runs, Z, Y, X, parent, size, best, 0, nullptr, 1, out_mask, stream);
}

// cfe_ccl_largest + the packed selection (see cfe_ccl_write_mask_bits)
extern "C" int cfe_ccl_largest_bits(const uint32_t* bits, const uint32_t* pref, const uint64_t* n_runs,
                                    int64_t Z, int64_t Y, int64_t X, uint32_t* parent, uint32_t* size,
                                    uint64_t* best, uint8_t* out_mask, uint32_t* sel_bits, cudaStream_t stream)
{
    if (Z * Y <= 0 || X <= 0) return 0;
    int rc = cfe_ccl_label(bits, pref, n_runs, Z, Y, X, parent, size, stream);
    if (rc) return rc;
    rc = cfe_ccl_select(parent, size, n_runs, best, stream);
    if (rc) return rc;
    return write_select(bits, pref, Z, Y, X, parent, best, nullptr, 0, nullptr, 1, out_mask, sel_bits, stream);
}

// Same, with the general phase 3 (cfe_ccl_write_select): drop / src / fill.
extern "C" int cfe_ccl_largest_select(const uint32_t* bits, const uint32_t* pref, const uint64_t* n_runs,
                                      int64_t Z, int64_t Y, int64_t X, uint32_t* parent, uint32_t* size,
                                      uint64_t* best, int drop, const uint8_t* src, int fill, uint8_t* out,
                                      cudaStream_t stream)
{
    if (Z * Y <= 0 || X <= 0) return 0;
    int rc = cfe_ccl_label(bits, pref, n_runs, Z, Y, X, parent, size, stream);
    if (rc) return rc;
    rc = cfe_ccl_select(parent, size, n_runs, best, stream);
    if (rc) return rc;
    return cfe_ccl_write_select(bits, pref, Z, Y, X, parent, best, nullptr, drop, src, fill, out, stream);
}

// ------------------------------------------------------------------------------------------
// z-slab sharding: cross-slab merge support.

// roots[k] = root of the k-th run (raster order) of the voxel plane made of rows
// [row0, row0 + Y); *count = number of runs in the plane.
__global__ void __launch_bounds__(256)
k_ccl_plane_roots(const u32* __restrict__ bits, const u32* __restrict__ pref, i64 word0, i64 n_plane_words,
                  int W, const u32* __restrict__ parent, u32* __restrict__ roots, u32* __restrict__ count)
{
    const i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_plane_words) return;
    const i64 i = word0 + t;
    const u32 w = __ldg(bits + i);
    const u32 base = __ldg(pref + word0);
    const u32 pf = __ldg(pref + i);
    const u32 carry = (i % W) ? (__ldg(bits + i - 1) >> 31) : 0u;
    u32 starts = w & ~((w << 1) | carry);
    u32 k = pf - base;
    if (t == n_plane_words - 1) *count = k + (u32)__popc(starts);
    u32 id = pf;
    while (starts) {
        starts &= starts - 1;
        roots[k++] = __ldg(parent + id++);
    }
}

extern "C" int cfe_ccl_plane_roots(const uint32_t* bits, const uint32_t* pref, int64_t plane, int64_t Y,
                                   int64_t X, const uint32_t* parent, uint32_t* roots, uint32_t* count,
                                   cudaStream_t stream)
{
    const int W = (int)((X + 31) / 32);
    const i64 n = Y * W;
    if (n <= 0) return 0;
    k_ccl_plane_roots<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(bits, pref, plane * n, n, W, parent, roots,
                                                                      count);
    CFE_CHECK_LAUNCH("k_ccl_plane_roots");
    return 0;
}

// Interface between the top plane A of the slab below (bitsA, plane-local run prefixes prefA and
// the roots of its runs, all received from the neighbour rank) and this slab's bottom plane B:
// one edge (rootA, rootB) per pair of overlapping runs, appended in arbitrary order.
__global__ void __launch_bounds__(256)
k_ccl_interface_edges(const u32* __restrict__ bitsA, const u32* __restrict__ prefA,
                      const u32* __restrict__ rootsA, const u32* __restrict__ bitsB,
                      const u32* __restrict__ prefB, const u32* __restrict__ parentB, i64 n_plane_words, int W,
                      u32* __restrict__ edges, u32* __restrict__ count, u32 capacity)
{
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_plane_words) return;
    const u32 a = __ldg(bitsA + i), b = __ldg(bitsB + i);
    const u32 both = a & b;
    if (!both) return;
    const bool inner = (i % W) != 0;
    const u32 a_prev = inner ? __ldg(bitsA + i - 1) : 0u;
    const u32 b_prev = inner ? __ldg(bitsB + i - 1) : 0u;
    const u32 pa = __ldg(prefA + i), pb = __ldg(prefB + i);
    u32 pair_starts = both & ~((both << 1) | ((a_prev & b_prev) >> 31));
    while (pair_starts) {
        const int bit = __ffs((int)pair_starts) - 1;
        pair_starts &= pair_starts - 1;
        const u32 ra = __ldg(rootsA + run_id(a, a_prev >> 31, pa, bit));
        const u32 rb = __ldg(parentB + run_id(b, b_prev >> 31, pb, bit));
        const u32 slot = atomicAdd(count, 1u);
        if (slot < capacity) { edges[2 * slot] = ra; edges[2 * slot + 1] = rb; }
    }
}

extern "C" int cfe_ccl_interface_edges(const uint32_t* bitsA, const uint32_t* prefA, const uint32_t* rootsA,
                                       const uint32_t* bitsB, const uint32_t* prefB, const uint32_t* parentB,
                                       int64_t Y, int64_t X, uint32_t* edges, uint32_t* count, int64_t capacity,
                                       cudaStream_t stream)
{
    const int W = (int)((X + 31) / 32);
    const i64 n = Y * W;
    CFE_CHECK_CUDA(cudaMemsetAsync(count, 0, sizeof(u32), stream));
    if (n <= 0) return 0;
    k_ccl_interface_edges<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(
        bitsA, prefA, rootsA, bitsB, prefB, parentB, n, W, edges, count, (u32)capacity);
    CFE_CHECK_LAUNCH("k_ccl_interface_edges");
    return 0;
}

// Same walk over the overlapping run pairs of an interface, for the device-side merge: the runs carry
// the dense boundary-vertex index of their component (vA: the neighbour's numbering, vB: this slab's).
// A trabecular interface has ~10^5 overlapping run pairs but only as many DISTINCT component pairs as
// there are boundary components, so the pairs are united in a scratch forest over the (<= 2 Cv) vertices
// and only the unions that actually hooked two trees are appended: a spanning forest of the interface
// graph, at most nvA + nvB - 1 edges whatever the number of run pairs.
// (compare-and-swap, not atomicMin: atomicMin also LOWERS the parent of a node that another thread hooked a
// moment ago, which links two trees without anybody appending an edge for it -- the thread that did it goes
// on to unite the displaced parent with its target, but a third thread can complete that union with an edge
// that duplicates an appended one.  With CAS a link is created only by the thread that appends its edge.)
__device__ __forceinline__ bool uf_union_hooked(u32* parent, u32 a, u32 b)
{
    while (true) {
        a = uf_find(parent, a);
        b = uf_find(parent, b);
        if (a == b) return false;
        if (a < b) { const u32 t = a; a = b; b = t; }
        const u32 old = atomicCAS(parent + a, a, b);
        if (old == a) return true;
        a = old;
    }
}

__global__ void __launch_bounds__(256)
k_ccl_interface_forest(const u32* __restrict__ bitsA, const u32* __restrict__ prefA, const u32* __restrict__ vA,
                       u32 vA_len, const u32* __restrict__ bitsB, const u32* __restrict__ prefB,
                       const u32* __restrict__ vB, i64 n_plane_words, int W, u32 Cv, u32* __restrict__ forest,
                       u32* __restrict__ edges, u32* __restrict__ count, u32 capacity)
{
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_plane_words) return;
    const u32 a = __ldg(bitsA + i), b = __ldg(bitsB + i);
    const u32 both = a & b;
    if (!both) return;
    const bool inner = (i % W) != 0;
    const u32 a_prev = inner ? __ldg(bitsA + i - 1) : 0u;
    const u32 b_prev = inner ? __ldg(bitsB + i - 1) : 0u;
    const u32 pa = __ldg(prefA + i), pb = __ldg(prefB + i);
    u32 pair_starts = both & ~((both << 1) | ((a_prev & b_prev) >> 31));
    while (pair_starts) {
        const int bit = __ffs((int)pair_starts) - 1;
        pair_starts &= pair_starts - 1;
        const u32 ida = run_id(a, a_prev >> 31, pa, bit);
        // (the neighbour sends the vertices of its first vA_len plane runs only: a plane with more runs than the
        // message holds is reported like a capacity overflow and the step takes the host-driven merge)
        const u32 va = ida < vA_len ? __ldg(vA + ida) : 0xffffffffu;
        const u32 vb = __ldg(vB + run_id(b, b_prev >> 31, pb, bit));
        if (va >= Cv || vb >= Cv) {        // vertex capacity exceeded: the solver sees it in the vertex counts
            atomicMax(count, capacity + 1u);
            continue;
        }
        if (uf_union_hooked(forest, va, Cv + vb)) {
            const u32 slot = atomicAdd(count, 1u);
            CFE_ASSERT(slot < 2u * Cv || slot >= capacity);   // (the overflow flag above lifts the count past capacity)
            if (slot < capacity) { edges[2 * slot] = va; edges[2 * slot + 1] = vb; }
        }
    }
}

__global__ void __launch_bounds__(256) k_uf_iota(u32* __restrict__ label, i64 n);

// forest: u32[2 * Cv] scratch; edges: u32[2 * capacity] (capacity >= 2 * Cv never overflows); *count = edges
extern "C" int cfe_ccl_interface_forest(const uint32_t* bitsA, const uint32_t* prefA, const uint32_t* vA,
                                        int64_t vA_len, const uint32_t* bitsB, const uint32_t* prefB,
                                        const uint32_t* vB, int64_t Y, int64_t X, int64_t Cv, uint32_t* forest,
                                        uint32_t* edges, uint32_t* count, int64_t capacity, cudaStream_t stream)
{
    const int W = (int)((X + 31) / 32);
    const i64 n = Y * W;
    CFE_CHECK_CUDA(cudaMemsetAsync(count, 0, sizeof(u32), stream));
    if (n <= 0 || Cv <= 0) return 0;
    k_uf_iota<<<(unsigned)((2 * Cv + 255) / 256), 256, 0, stream>>>(forest, 2 * Cv);
    CFE_CHECK_LAUNCH("k_uf_iota");
    k_ccl_interface_forest<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(
        bitsA, prefA, vA, (u32)vA_len, bitsB, prefB, vB, n, W, (u32)Cv, forest, edges, count, (u32)capacity);
    CFE_CHECK_LAUNCH("k_ccl_interface_forest");
    return 0;
}

// ------------------------------------------------------------------------------------------
// Boundary graph of the cross-slab merge: vertices = slab-local components that touch an
// interface plane (dense indices, sorted by global key), edges = overlapping run pairs.
// label[v] = smallest vertex index of v's component (same union-find as the voxel kernels).

__global__ void __launch_bounds__(256)
k_uf_edges(u32* __restrict__ label, const i64* __restrict__ ea, const i64* __restrict__ eb, i64 n_edges)
{
    const i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n_edges) uf_union(label, (u32)ea[e], (u32)eb[e]);
}

__global__ void __launch_bounds__(256) k_uf_flatten(u32* __restrict__ label, i64 n)
{
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    u32 r = (u32)i, p;
    while ((p = label[r]) != r) r = p;
    label[i] = r;
}

__global__ void __launch_bounds__(256) k_uf_iota(u32* __restrict__ label, i64 n)
{
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) label[i] = (u32)i;
}

// label: u32 [n_vertices] (initialised here); ea/eb: int64 dense endpoints of the n_edges edges
extern "C" int cfe_uf_solve(uint32_t* label, int64_t n_vertices, const int64_t* ea, const int64_t* eb,
                            int64_t n_edges, cudaStream_t stream)
{
    if (n_vertices <= 0) return 0;
    const unsigned vb = (unsigned)((n_vertices + 255) / 256);
    k_uf_iota<<<vb, 256, 0, stream>>>(label, n_vertices);
    CFE_CHECK_LAUNCH("k_uf_iota");
    if (n_edges > 0) {
        k_uf_edges<<<(unsigned)((n_edges + 255) / 256), 256, 0, stream>>>(label, reinterpret_cast<const i64*>(ea),
                                                                          reinterpret_cast<const i64*>(eb), n_edges);
        CFE_CHECK_LAUNCH("k_uf_edges");
        k_uf_flatten<<<vb, 256, 0, stream>>>(label, n_vertices);
        CFE_CHECK_LAUNCH("k_uf_flatten");
    }
    return 0;
}

// ------------------------------------------------------------------------------------------
// Device-side cross-slab merge (z-slab sharding).  Per rank a fixed-capacity message
//   pack[0] = n_vertices, pack[1] = n_edges, pack[2..3] = best slab-internal component (u64),
//   pack[4 .. 4+Cv) vertex roots, pack[4+Cv .. 4+2Cv) vertex sizes, then Ce (a, b) edge pairs
// is built by three launches, all-gathered over NCCL, and solved identically on every rank by
// one single-CTA kernel that also writes the keep markers; the host reads 4 words at the end.

__global__ void __launch_bounds__(256) k_fill_u32(u32* __restrict__ a, const u64* __restrict__ n, u32 value)
{
    const u64 m = *n;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (u64)gridDim.x * blockDim.x) a[i] = value;
}

// Boundary vertices = slab-local roots owning a run in an interface plane.  vfirst[] (SENT-filled
// over the run ids) receives the smallest plane-run index of every such root; those first
// occurrences are numbered densely (any order: the solver only groups by index) and every plane
// run learns its vertex.  Three grid-wide passes over the (<= 2 * plane runs) entries.
__device__ __forceinline__ u32 bv_root(const u32* roots_bot, const u32* roots_top, u32 n_bot, u32 k)
{
    return k < n_bot ? roots_bot[k] : roots_top[k - n_bot];
}

__global__ void __launch_bounds__(256)
k_bv_first(const u32* __restrict__ roots_bot, const u32* __restrict__ roots_top, const u32* __restrict__ cnt,
           int use_bot, int use_top, u32* __restrict__ vfirst)
{
    const u32 n_bot = use_bot ? cnt[0] : 0u, n = n_bot + (use_top ? cnt[1] : 0u);
    for (u32 k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x)
        atomicMin(vfirst + bv_root(roots_bot, roots_top, n_bot, k), k);
}

__global__ void __launch_bounds__(256)
k_bv_number(const u32* __restrict__ roots_bot, const u32* __restrict__ roots_top, const u32* __restrict__ cnt,
            int use_bot, int use_top, const u32* __restrict__ vfirst, u32* __restrict__ didx,
            const u32* __restrict__ size, u32 Cv, u32* __restrict__ pack)
{
    const u32 n_bot = use_bot ? cnt[0] : 0u, n = n_bot + (use_top ? cnt[1] : 0u);
    const u32 lane = threadIdx.x & 31u;
    for (u32 k0 = blockIdx.x * blockDim.x + threadIdx.x - lane; k0 < n; k0 += gridDim.x * blockDim.x) {
        const u32 k = k0 + lane;
        u32 r = 0;
        bool flag = false;
        if (k < n) {
            r = bv_root(roots_bot, roots_top, n_bot, k);
            flag = vfirst[r] == k;
        }
        const u32 vote = __ballot_sync(CFE_FULL_MASK, flag);
        if (!vote) continue;
        u32 base = 0;
        if (lane == 0) base = atomicAdd(pack, (u32)__popc(vote));
        base = __shfl_sync(CFE_FULL_MASK, base, 0);
        if (flag) {
            const u32 d = base + (u32)__popc(vote & ((1u << lane) - 1u));
            didx[k] = d;
            if (d < Cv) { pack[4 + d] = r; pack[4 + Cv + d] = size[r]; }
        }
    }
}

__global__ void __launch_bounds__(256)
k_bv_assign(const u32* __restrict__ roots_bot, const u32* __restrict__ roots_top, const u32* __restrict__ cnt,
            int use_bot, int use_top, const u32* __restrict__ vfirst, const u32* __restrict__ didx,
            u32* __restrict__ vbot, u32* __restrict__ vtop)
{
    const u32 n_bot = use_bot ? cnt[0] : 0u, n = n_bot + (use_top ? cnt[1] : 0u);
    for (u32 k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const u32 v = didx[vfirst[bv_root(roots_bot, roots_top, n_bot, k)]];
        if (k < n_bot) vbot[k] = v; else vtop[k - n_bot] = v;
    }
}

extern "C" int cfe_ccl_boundary_vertices(const uint32_t* roots_bot, const uint32_t* roots_top, const uint32_t* cnt,
                                         int use_bot, int use_top, const uint64_t* n_runs, uint32_t* vfirst,
                                         uint32_t* didx, const uint32_t* parent, const uint32_t* size,
                                         int64_t Cv, uint32_t* pack, uint32_t* vbot, uint32_t* vtop,
                                         cudaStream_t stream)
{
    const unsigned gs = grid_stride_blocks();
    k_fill_u32<<<gs, 256, 0, stream>>>(vfirst, reinterpret_cast<const u64*>(n_runs), 0xffffffffu);
    CFE_CHECK_LAUNCH("k_fill_u32");
    CFE_CHECK_CUDA(cudaMemsetAsync(pack, 0, sizeof(u32), stream));
    k_bv_first<<<gs, 256, 0, stream>>>(roots_bot, roots_top, cnt, use_bot, use_top, vfirst);
    CFE_CHECK_LAUNCH("k_bv_first");
    k_bv_number<<<gs, 256, 0, stream>>>(roots_bot, roots_top, cnt, use_bot, use_top, vfirst, didx, size, (u32)Cv, pack);
    CFE_CHECK_LAUNCH("k_bv_number");
    k_bv_assign<<<gs, 256, 0, stream>>>(roots_bot, roots_top, cnt, use_bot, use_top, vfirst, didx, vbot, vtop);
    CFE_CHECK_LAUNCH("k_bv_assign");
    // best slab-internal component: roots that are boundary vertices are excluded
    CFE_CHECK_CUDA(cudaMemsetAsync(pack + 2, 0, sizeof(u64), stream));
    k_ccl_select<<<gs, 256, 0, stream>>>(parent, size, reinterpret_cast<const u64*>(n_runs),
                                         reinterpret_cast<u64*>(pack + 2), vfirst);
    CFE_CHECK_LAUNCH("k_ccl_select");
    return 0;
}

struct BCand { u64 size; u64 key; u32 rep; };

__device__ __forceinline__ bool bcand_better(const BCand& a, const BCand& b)
{
    // larger component first; ties -> smaller key (rank << 32 | root) = earlier in raster order
    return a.size > b.size || (a.size == b.size && a.key < b.key);
}

__global__ void __launch_bounds__(1024)
k_boundary_solve(const u32* __restrict__ packs, int N, u32 Cv, u32 Ce, int rank, u32* __restrict__ label,
                 u64* __restrict__ comp, u64* __restrict__ ckey, u32* __restrict__ size, i64* __restrict__ result)
{
    __shared__ u32 s_nr[64], s_ne[64];
    __shared__ int s_ovf;
    __shared__ BCand s_c[32];
    __shared__ BCand s_win;
    const u32 PS = 4u + 2u * Cv + 2u * Ce;
    const u32 t = threadIdx.x;
    if (t == 0) s_ovf = 0;
    __syncthreads();
    if (t < (u32)N) {
        s_nr[t] = packs[(size_t)t * PS];
        s_ne[t] = packs[(size_t)t * PS + 1];
        if (s_nr[t] > Cv || s_ne[t] > Ce) s_ovf = 1;
    }
    __syncthreads();
    if (s_ovf) {
        if (t == 0) { result[0] = 0; result[1] = 0; result[2] = 1; result[3] = 0; }
        return;
    }
    for (int k = 0; k < N; ++k)
        for (u32 i = t; i < s_nr[k]; i += 1024u) {
            const u32 v = (u32)k * Cv + i;
            label[v] = v; comp[v] = 0; ckey[v] = ~0ull;
        }
    __syncthreads();
    for (int k = 1; k < N; ++k) {
        const u32* e = packs + (size_t)k * PS + 4u + 2u * Cv;
        for (u32 i = t; i < s_ne[k]; i += 1024u) {
            const u32 a = e[2 * i], b = e[2 * i + 1];
            if (a < s_nr[k - 1] && b < s_nr[k]) uf_union(label, (u32)(k - 1) * Cv + a, (u32)k * Cv + b);
        }
    }
    __syncthreads();
    for (int k = 0; k < N; ++k) {
        const u32* pk = packs + (size_t)k * PS;
        for (u32 i = t; i < s_nr[k]; i += 1024u) {
            const u32 v = (u32)k * Cv + i;
            u32 r = v, p;
            while ((p = label[r]) != r) r = p;
            atomicAdd(reinterpret_cast<unsigned long long*>(comp + r), (unsigned long long)pk[4 + Cv + i]);
            atomicMin(reinterpret_cast<unsigned long long*>(ckey + r), ((unsigned long long)k << 32) | pk[4 + i]);
        }
    }
    __syncthreads();
    // flatten (second pass so that the accumulation above saw a stable forest)
    for (int k = 0; k < N; ++k)
        for (u32 i = t; i < s_nr[k]; i += 1024u) {
            const u32 v = (u32)k * Cv + i;
            u32 r = v, p;
            while ((p = label[r]) != r) r = p;
            if (r != v) label[v] = r;
        }
    __syncthreads();
    BCand best = {0ull, ~0ull, 0xffffffffu};
    for (int k = 0; k < N; ++k)
        for (u32 i = t; i < s_nr[k]; i += 1024u) {
            const u32 v = (u32)k * Cv + i;
            if (label[v] == v) {
                const BCand c = {comp[v], ckey[v], v};
                if (bcand_better(c, best)) best = c;
            }
        }
    if (t < (u32)N) {
        const u64 b = *reinterpret_cast<const u64*>(packs + (size_t)t * PS + 2);
        const BCand c = {b >> 32, ((u64)t << 32) | (u64)(0xffffffffu - (u32)(b & 0xffffffffu)), 0xffffffffu};
        if (c.size > 0 && bcand_better(c, best)) best = c;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        BCand other;
        other.size = __shfl_xor_sync(CFE_FULL_MASK, best.size, o);
        other.key = __shfl_xor_sync(CFE_FULL_MASK, best.key, o);
        other.rep = __shfl_xor_sync(CFE_FULL_MASK, best.rep, o);
        if (bcand_better(other, best)) best = other;
    }
    if ((t & 31u) == 0) s_c[t >> 5] = best;
    __syncthreads();
    if (t == 0) {
        for (int w = 1; w < 32; ++w) if (bcand_better(s_c[w], best)) best = s_c[w];
        s_win = best;
        result[0] = (i64)best.size; result[1] = (i64)best.key; result[2] = 0; result[3] = (i64)best.rep;
    }
    __syncthreads();
    const BCand w = s_win;
    if (w.size == 0) return;
    if (w.rep != 0xffffffffu) {
        const u32* pk = packs + (size_t)rank * PS;
        for (u32 i = t; i < s_nr[rank]; i += 1024u)
            if (label[(u32)rank * Cv + i] == w.rep) size[pk[4 + i]] = 0xffffffffu;
    } else if ((int)(w.key >> 32) == rank) {
        if (t == 0) size[(u32)(w.key & 0xffffffffu)] = 0xffffffffu;
    }
}

// packs: the all-gathered messages [N][4 + 2 Cv + 2 Ce] u32; label u32 [N*Cv], comp / ckey u64 [N*Cv]
// scratch; size: this rank's size array (keep markers are written into it); result: i64[4] =
// {winner size, winner key (rank << 32 | root), capacity overflow flag, winner vertex or -1}
extern "C" int cfe_ccl_boundary_solve(const uint32_t* packs, int N, int64_t Cv, int64_t Ce, int rank,
                                      uint32_t* label, uint64_t* comp, uint64_t* ckey, uint32_t* size,
                                      int64_t* result, cudaStream_t stream)
{
    if (N > 64) { cfe_set_error("cfe_ccl_boundary_solve: at most 64 slabs"); return -1; }
    k_boundary_solve<<<1, 1024, 0, stream>>>(packs, N, (u32)Cv, (u32)Ce, rank, label, reinterpret_cast<u64*>(comp),
                                             reinterpret_cast<u64*>(ckey), size, reinterpret_cast<i64*>(result));
    CFE_CHECK_LAUNCH("k_boundary_solve");
    return 0;
}
