// ciclope_b200 -- voxel-grid kernels (sm_100a): bit packing, element / node / run numbering
// by decoupled look-back scans over 32-voxel words, list filling, boundary sets.
//
// Replaces the Python loops of ciclope/core/voxelFE.py:135-192 (cell loop) and :195-251 (node
// loop) and np.unique(cells) (:628).  Layout in HBM (one z-slab per GPU):
//   vol    uint8 [Zl][Y][X]                      the caller's volume, read once per pass
//   bits   u32   [Zl*Y][W], W = ceil(X/32)       bit b of word w <-> column 32w+b, 1 = element
//   pref   u32   [Zl*Y*W]                        exclusive count of elements before each word
//   nbits  u32   [NL*(Y+1)][WN], WN=ceil((X+1)/32)  1 = grid node touched by >= 1 element
//   npref  u32   [NL*(Y+1)*WN]                   exclusive count of used nodes before each word
#include "common.cuh"
#include "ciclope_b200.h"

#define SCAN_THREADS 256
#ifndef SCAN_ITEMS
#define SCAN_ITEMS 8
#endif
#define SCAN_TILE (SCAN_THREADS * SCAN_ITEMS)
#define SCAN_CHUNK 512u         // items per entry of the coarse rank index (= nodes per CTA of k_emit_nodes4b)

// ------------------------------------------------------------------------------------------
// K0: pack (v > thr) into bit rows

// 4 bits: byte k of x > thr (unsigned), thr in [0, 254].  y4 = (thr + 1) in every byte; x > thr <=> x >= y:
// the low 7 bits are compared by a borrow-free subtraction ((x | H) - (y & ~H) keeps bit 7 of a byte
// set iff x7 >= y7 in the low 7 bits), the top bits by logic; the four bit-7 flags are then gathered
// with one multiply (7+21, 15+14, 23+7, 31+0 -> bits 28..31, no two partial products collide).
__device__ __forceinline__ u32 cmp_nibble(u32 x, u32 y4, bool all)
{
    const u32 H = 0x80808080u;
    const u32 d = (x | H) - (y4 & ~H);
    const u32 ge = ((x & ~y4) | (~(x ^ y4) & d)) & H;
    return all ? 0xfu : (ge * 0x00204081u) >> 28;
}

#define PACK_UNROLL 4

// FLAT: X % 32 == 0 and vol 16-byte aligned -> the volume is a flat array of 16-byte chunks.
template <bool FLAT>
__global__ void __launch_bounds__(256) k_pack_bits(const uint8_t* __restrict__ vol, i64 rows, int X, int W,
                                                   int thr, int invert, u32* __restrict__ bits)
{
    const i64 nchunks = rows * 2 * (i64)W;
    const bool all = thr < 0;
    const bool none = thr >= 255;                              // no uint8 value exceeds 255
    const u32 thr4 = (u32)((thr < 0 || none) ? 1 : thr + 1) * 0x01010101u;   // y = thr + 1 per byte
    // PACK_UNROLL chunks per thread, 256 chunks apart, all loads issued first (bytes in flight)
    const i64 g0 = (i64)blockIdx.x * (256 * PACK_UNROLL) + threadIdx.x;
    uint4 v[PACK_UNROLL];
    if (FLAT) {
#pragma unroll
        for (int u = 0; u < PACK_UNROLL; ++u) {
            const i64 g = g0 + 256 * u;
            v[u] = g < nchunks ? __ldg(reinterpret_cast<const uint4*>(vol) + g) : make_uint4(0u, 0u, 0u, 0u);
        }
    }
#pragma unroll
    for (int u = 0; u < PACK_UNROLL; ++u) {
        const i64 g = g0 + 256 * u;
        u32 m16 = 0;
        if (g < nchunks) {
            if (FLAT) {
                m16 = cmp_nibble(v[u].x, thr4, all) | (cmp_nibble(v[u].y, thr4, all) << 4) |
                      (cmp_nibble(v[u].z, thr4, all) << 8) | (cmp_nibble(v[u].w, thr4, all) << 12);
                if (none) m16 = 0;
                if (invert) m16 ^= 0xffffu;
            } else {
                const i64 row = g / (2 * W);
                const int c0 = (int)(g - row * 2 * W) * 16;
                const uint8_t* p = vol + row * (i64)X + c0;
#pragma unroll
                for (int k = 0; k < 16; ++k)
                    if (c0 + k < X) m16 |= (u32)(((int)p[k] > thr) != (invert != 0)) << k;   // pad bits stay 0
            }
        }
        const u32 other = __shfl_xor_sync(CFE_FULL_MASK, m16, 1);
        if (g < nchunks && !(threadIdx.x & 1)) bits[g >> 1] = m16 | (other << 16);
    }
}

static int pack_bits(const uint8_t* vol, int64_t rows, int64_t X, int thr, int invert, uint32_t* bits,
                     cudaStream_t stream)
{
    if (rows <= 0 || X <= 0) return 0;
    const int W = (int)((X + 31) / 32);
    const i64 nchunks = rows * 2 * (i64)W;
    const i64 blocks = (nchunks + 256 * PACK_UNROLL - 1) / (256 * PACK_UNROLL);
    if (blocks > 0x7fffffffLL) { cfe_set_error("cfe_pack_bits: volume too large"); return -1; }
    const bool flat = (X % 32 == 0) && ((reinterpret_cast<uintptr_t>(vol) & 15u) == 0);
    if (flat)
        k_pack_bits<true><<<(unsigned)blocks, 256, 0, stream>>>(vol, rows, (int)X, W, thr, invert, bits);
    else
        k_pack_bits<false><<<(unsigned)blocks, 256, 0, stream>>>(vol, rows, (int)X, W, thr, invert, bits);
    CFE_CHECK_LAUNCH("k_pack_bits");
    return 0;
}

extern "C" int cfe_pack_bits(const uint8_t* vol, int64_t rows, int64_t X, int thr, uint32_t* bits,
                             cudaStream_t stream)
{
    return pack_bits(vol, rows, X, thr, 0, bits, stream);
}

// bits of the voxels with v <= thr: the background ~(I > 0) that fill_voids labels (preprocess.py:167)
extern "C" int cfe_pack_bits_le(const uint8_t* vol, int64_t rows, int64_t X, int thr, uint32_t* bits,
                                cudaStream_t stream)
{
    return pack_bits(vol, rows, X, thr, 1, bits, stream);
}

// *out = max(*out, max of vol[0..n)): I.max(), the default fill value of fill_voids (preprocess.py:163)
__global__ void __launch_bounds__(256) k_max_u8(const uint8_t* __restrict__ vol, i64 n, u32* __restrict__ out)
{
    u32 m = 0;
    const i64 n16 = (reinterpret_cast<uintptr_t>(vol) & 15u) ? 0 : n / 16;
    for (i64 g = (i64)blockIdx.x * blockDim.x + threadIdx.x; g < n16; g += (i64)gridDim.x * blockDim.x) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(vol) + g);
        m = __vmaxu4(m, __vmaxu4(__vmaxu4(v.x, v.y), __vmaxu4(v.z, v.w)));
    }
    m = max(max(m & 0xffu, (m >> 8) & 0xffu), max((m >> 16) & 0xffu, m >> 24));
    for (i64 g = n16 * 16 + (i64)blockIdx.x * blockDim.x + threadIdx.x; g < n; g += (i64)gridDim.x * blockDim.x)
        m = max(m, (u32)vol[g]);
    m = __reduce_max_sync(CFE_FULL_MASK, m);
    if ((threadIdx.x & 31) == 0 && m) atomicMax(out, m);
}

extern "C" int cfe_max_u8(const uint8_t* vol, int64_t n, uint32_t* out, cudaStream_t stream)
{
    if (n <= 0) return 0;
    const i64 want = (n / 16 + 255) / 256 + 1;
    k_max_u8<<<(unsigned)(want < 148 * 8 ? want : 148 * 8), 256, 0, stream>>>(vol, n, out);
    CFE_CHECK_LAUNCH("k_max_u8");
    return 0;
}

// out[0] = max(out[0], max v), out[1] = min(out[1], min of the non-zero v): a uint8 image is "binary" for
// measure.label(bw, None, True, 1) (preprocess.py:135, regions of EQUAL value) iff both agree
__global__ void __launch_bounds__(256) k_value_span_u8(const uint8_t* __restrict__ vol, i64 n, u32* __restrict__ out)
{
    u32 m = 0, lo = 0xffffffffu;
    const i64 n16 = (reinterpret_cast<uintptr_t>(vol) & 15u) ? 0 : n / 16;
    for (i64 g = (i64)blockIdx.x * blockDim.x + threadIdx.x; g < n16; g += (i64)gridDim.x * blockDim.x) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(vol) + g);
        m = __vmaxu4(m, __vmaxu4(__vmaxu4(v.x, v.y), __vmaxu4(v.z, v.w)));
        // zero bytes read as 0xff for the minimum
        lo = __vminu4(lo, __vminu4(__vminu4(v.x | __vcmpeq4(v.x, 0), v.y | __vcmpeq4(v.y, 0)),
                                   __vminu4(v.z | __vcmpeq4(v.z, 0), v.w | __vcmpeq4(v.w, 0))));
    }
    m = max(max(m & 0xffu, (m >> 8) & 0xffu), max((m >> 16) & 0xffu, m >> 24));
    lo = min(min(lo & 0xffu, (lo >> 8) & 0xffu), min((lo >> 16) & 0xffu, lo >> 24));
    for (i64 g = n16 * 16 + (i64)blockIdx.x * blockDim.x + threadIdx.x; g < n; g += (i64)gridDim.x * blockDim.x) {
        const u32 v = vol[g];
        m = max(m, v);
        if (v) lo = min(lo, v);
    }
    m = __reduce_max_sync(CFE_FULL_MASK, m);
    lo = __reduce_min_sync(CFE_FULL_MASK, lo);
    if ((threadIdx.x & 31) == 0 && m) {      // (a warp that saw a non-zero value: lo is its smallest one)
        atomicMax(out, m);
        atomicMin(out + 1, lo);
    }
}

extern "C" int cfe_value_span_u8(const uint8_t* vol, int64_t n, uint32_t* out, cudaStream_t stream)
{
    CFE_CHECK_CUDA(cudaMemsetAsync(out, 0, 4, stream));
    CFE_CHECK_CUDA(cudaMemsetAsync(out + 1, 0xff, 4, stream));
    if (n <= 0) return 0;
    const i64 want = (n / 16 + 255) / 256 + 1;
    k_value_span_u8<<<(unsigned)(want < 148 * 8 ? want : 148 * 8), 256, 0, stream>>>(vol, n, out);
    CFE_CHECK_LAUNCH("k_value_span_u8");
    return 0;
}

// ------------------------------------------------------------------------------------------
// Word functors for the scan kernel.

struct NodeGeom {
    int Zl, Y, W, WN;        // local voxel slices, rows, words per voxel row / node row
    int n_layers;            // node layers owned by this slab (Zl, or Zl+1 on the last slab)
    const u32* bits;         // [Zl*Y][W]
    const u32* halo_below;   // [Y][W] element bits of the voxel plane below local slice 0, or null
};

// 1 bit per grid node of node row (layer s, row r), word w: node used <=> any of the <= 8
// adjacent voxels is an element (np.unique(cells), voxelFE.py:628).
__device__ __forceinline__ u32 node_used_word(const NodeGeom& g, int s, int r, int w)
{
    u32 u = 0, up = 0;
#pragma unroll
    for (int ds = 0; ds < 2; ++ds) {
        const int vs = s - ds;
        const u32* plane;
        if (vs >= 0 && vs < g.Zl) plane = g.bits + (size_t)vs * g.Y * g.W;
        else if (vs == -1 && g.halo_below) plane = g.halo_below;
        else continue;
#pragma unroll
        for (int dr = 0; dr < 2; ++dr) {
            const int vr = r - dr;
            if (vr < 0 || vr >= g.Y) continue;
            const u32* row = plane + (size_t)vr * g.W;
            if (w < g.W) u |= __ldg(row + w);
            if (w > 0) up |= __ldg(row + w - 1);
        }
    }
    return u | (u << 1) | (up >> 31);
}

// MODE 0: popc(bits[i])                       element numbering        (voxelFE.py:142)
// MODE 1: popc(run starts of bits[i])         x-run numbering for CCL
// MODE 2: popc(node_used_word) + store word   used-node numbering      (voxelFE.py:628)
// words per thread: the popcount scans (modes 0, 1) are bound by the look-back chain and run faster with
// fatter tiles (measured 0.16 -> 0.12 ms for 2^25 words); the node scan (mode 2) computes its words from up to
// eight voxel-row words each and is better off with more, thinner tiles (0.34 vs 0.38 ms)
template <int MODE> struct ScanItems { static constexpr int value = (MODE == 2) ? SCAN_ITEMS : 2 * SCAN_ITEMS; };

template <int MODE>
__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_words(const u32* __restrict__ bits, i64 n_words, int W, NodeGeom geom, u32* __restrict__ nbits,
             u32* __restrict__ pref, u64* __restrict__ total, u64* __restrict__ ws, u32* __restrict__ chunk_word)
{
    constexpr int ITEMS = ScanItems<MODE>::value;
    constexpr int TILE = SCAN_THREADS * ITEMS;
    __shared__ u32 s_scan[33];
    __shared__ u64 s_excl;
    const u32 tile = scan_next_tile(ws);
    const i64 i0 = (i64)tile * TILE + (i64)threadIdx.x * ITEMS;
    u32 cnt[ITEMS];
    u32 sum = 0;
    if (MODE == 2) {
        // decompose the first word index once, then step
        const i64 words_per_layer = (i64)(geom.Y + 1) * geom.WN;
        int s = (int)(i0 / words_per_layer);
        i64 rem = i0 - (i64)s * words_per_layer;
        int r = (int)(rem / geom.WN);
        int w = (int)(rem - (i64)r * geom.WN);
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) {
            u32 word = 0;
            if (i0 + k < n_words) {
                word = node_used_word(geom, s, r, w);
                nbits[i0 + k] = word;
            }
            cnt[k] = __popc(word);
            sum += cnt[k];
            if (++w == geom.WN) { w = 0; if (++r == geom.Y + 1) { r = 0; ++s; } }
        }
    } else {
        u32 word[ITEMS];
        if (i0 + ITEMS <= n_words && (reinterpret_cast<uintptr_t>(bits) & 15u) == 0) {   // 16-byte loads
#pragma unroll
            for (int k = 0; k < ITEMS; k += 4) {
                const uint4 v = __ldg(reinterpret_cast<const uint4*>(bits + i0 + k));
                word[k] = v.x; word[k + 1] = v.y; word[k + 2] = v.z; word[k + 3] = v.w;
            }
        } else {
#pragma unroll
            for (int k = 0; k < ITEMS; ++k) word[k] = (i0 + k < n_words) ? __ldg(bits + i0 + k) : 0u;
        }
        if (MODE == 1) {
            int w = (int)(i0 % W);                      // word position inside the voxel row
            u32 prev = (w != 0 && i0 < n_words) ? __ldg(bits + i0 - 1) : 0u;
#pragma unroll
            for (int k = 0; k < ITEMS; ++k) {
                const u32 carry = w ? (prev >> 31) : 0u;
                cnt[k] = __popc(word[k] & ~((word[k] << 1) | carry));
                prev = word[k];
                if (++w == W) w = 0;
                sum += cnt[k];
            }
        } else {
#pragma unroll
            for (int k = 0; k < ITEMS; ++k) { cnt[k] = __popc(word[k]); sum += cnt[k]; }
        }
    }
    u32 block_total;
    u32 excl = block_exclusive_scan<u32>(sum, s_scan, block_total);
    if (threadIdx.x < 32) {
        u64 e = scan_lookback(ws, tile, (u64)block_total);
        if (threadIdx.x == 0) {
            s_excl = e;
            if ((i64)(tile + 1) * TILE >= n_words) *total = e + block_total;  // last tile
        }
    }
    __syncthreads();
    u32 run = (u32)s_excl + excl;
    if (chunk_word) {
        // coarse rank index for the list-free emitters: chunk_word[m] = the word that holds the item of rank
        // m * SCAN_CHUNK (a word holds <= 32 items, so at most one chunk starts in it)
        u32 o = run;
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) {
            const u32 m = (o + SCAN_CHUNK - 1u) / SCAN_CHUNK;
            if (cnt[k] && m * SCAN_CHUNK < o + cnt[k]) chunk_word[m] = (u32)(i0 + k);
            o += cnt[k];
        }
    }
    if (i0 + ITEMS <= n_words && (reinterpret_cast<uintptr_t>(pref) & 15u) == 0) {
        u32 o[ITEMS];
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) { o[k] = run; run += cnt[k]; }
#pragma unroll
        for (int k = 0; k < ITEMS; k += 4)
            *reinterpret_cast<uint4*>(pref + i0 + k) = make_uint4(o[k], o[k + 1], o[k + 2], o[k + 3]);
    } else {
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) {
            if (i0 + k < n_words) pref[i0 + k] = run;
            run += cnt[k];
        }
    }
}

extern "C" size_t cfe_scan_workspace_bytes(int64_t n_items)
{
    // one status word per SCAN_TILE items (+ counter); callers may use the same workspace for
    // any tile size >= 256 items.
    return (size_t)((n_items + 255) / 256 + 2) * sizeof(u64);
}

static int launch_scan(int mode, const u32* bits, i64 n_words, int W, const NodeGeom& geom, u32* nbits,
                       u32* pref, u64* total, void* ws, cudaStream_t stream, u32* chunk_word = nullptr)
{
    if (n_words >= (1LL << 32) * 32 / 32) {
        // prefixes are u32: a slab holds < 2^32 voxels / nodes
    }
    const i64 tile_items = (i64)SCAN_THREADS * (mode == 2 ? ScanItems<2>::value : ScanItems<0>::value);
    const i64 tiles = (n_words + tile_items - 1) / tile_items;
    CFE_CHECK_CUDA(cudaMemsetAsync(ws, 0, (size_t)(tiles + 2) * sizeof(u64), stream));
    if (tiles == 0) {
        CFE_CHECK_CUDA(cudaMemsetAsync(total, 0, sizeof(u64), stream));
        return 0;
    }
    if (tiles > 0x7fffffffLL) { cfe_set_error("scan: too many tiles"); return -1; }
    u64* w = reinterpret_cast<u64*>(ws);
    switch (mode) {
    case 0: k_scan_words<0><<<(unsigned)tiles, SCAN_THREADS, 0, stream>>>(bits, n_words, W, geom, nbits, pref, total, w, chunk_word); break;
    case 1: k_scan_words<1><<<(unsigned)tiles, SCAN_THREADS, 0, stream>>>(bits, n_words, W, geom, nbits, pref, total, w, chunk_word); break;
    default: k_scan_words<2><<<(unsigned)tiles, SCAN_THREADS, 0, stream>>>(bits, n_words, W, geom, nbits, pref, total, w, chunk_word); break;
    }
    CFE_CHECK_LAUNCH("k_scan_words");
    return 0;
}

extern "C" int cfe_scan_elements(const uint32_t* bits, int64_t n_words, uint32_t* pref, uint64_t* total,
                                 void* ws, cudaStream_t stream)
{
    NodeGeom g = {};
    return launch_scan(0, bits, n_words, 1, g, nullptr, pref, reinterpret_cast<u64*>(total), ws, stream);
}

extern "C" int cfe_scan_run_starts(const uint32_t* bits, int64_t rows, int64_t X, uint32_t* pref,
                                   uint64_t* total, void* ws, cudaStream_t stream)
{
    NodeGeom g = {};
    const int W = (int)((X + 31) / 32);
    return launch_scan(1, bits, rows * W, W, g, nullptr, pref, reinterpret_cast<u64*>(total), ws, stream);
}

// chunk_word (may be NULL): u32 [n_layers*(Y+1)*(X+1) / 512 + 2], receives for every m the index of the word
// of the node grid that holds the used node of rank 512 m -- the entry point of the list-free *NODE emitter
extern "C" int cfe_scan_nodes_index(const uint32_t* bits, const uint32_t* halo_below, int64_t Zl, int64_t Y,
                                    int64_t X, int64_t n_layers, uint32_t* nbits, uint32_t* npref,
                                    uint64_t* total, uint32_t* chunk_word, void* ws, cudaStream_t stream)
{
    NodeGeom g;
    g.Zl = (int)Zl; g.Y = (int)Y; g.W = (int)((X + 31) / 32); g.WN = (int)((X + 1 + 31) / 32);
    g.n_layers = (int)n_layers; g.bits = bits; g.halo_below = halo_below;
    const i64 n_words = n_layers * (Y + 1) * (i64)g.WN;
    return launch_scan(2, bits, n_words, g.WN, g, nbits, npref, reinterpret_cast<u64*>(total), ws, stream, chunk_word);
}

extern "C" int cfe_scan_nodes(const uint32_t* bits, const uint32_t* halo_below, int64_t Zl, int64_t Y,
                              int64_t X, int64_t n_layers, uint32_t* nbits, uint32_t* npref,
                              uint64_t* total, void* ws, cudaStream_t stream)
{
    return cfe_scan_nodes_index(bits, halo_below, Zl, Y, X, n_layers, nbits, npref, total, nullptr, ws, stream);
}

// ------------------------------------------------------------------------------------------
// K9: expand set bits into an index list: list[pref[i] + k] = row*rowlen + 32*w + b
// (elements: rowlen = X, words per row W; nodes: rowlen = X+1, words per row WN).

// One CTA expands 256 consecutive words (<= 8192 entries) into shared memory at their final
// relative positions, then streams the contiguous span to HBM with coalesced stores.
#define FILL_WORDS 256
// 1: staged lists leave shared memory through the bulk-copy engine (cp.async.bulk.global.shared::cta)
#ifndef CFE_BULK_STORE
#define CFE_BULK_STORE 1
#endif

__global__ void __launch_bounds__(FILL_WORDS)
k_fill_list(const u32* __restrict__ bits, const u32* __restrict__ pref, u32 n_words, FastDiv d_wpr, u32 wpr,
            u32 rowlen, u32* __restrict__ list)
{
    // entries are staged at sm[off + q], off = (destination word index) mod 4, so that 16-byte chunks of the
    // staging area are 16-byte chunks of the destination: the aligned middle leaves in ONE bulk copy
    __shared__ __align__(16) u32 sm[FILL_WORDS * 32 + 4];
    __shared__ u32 s_total;
    const u32 i0 = blockIdx.x * FILL_WORDS;
    const u32 i = i0 + threadIdx.x;
    const u32 base0 = __ldg(pref + i0);
    u32* dst = list + base0;
    const u32 off = (u32)(((uintptr_t)dst >> 2) & 3u);
    u32 word = 0, p = 0;
    if (i < n_words) {
        word = __ldg(bits + i);
        p = __ldg(pref + i) - base0;
    }
    const u32 last = (n_words - i0 < FILL_WORDS ? n_words - i0 : FILL_WORDS) - 1u;
    if (threadIdx.x == last) s_total = p + (u32)__popc(word);
    CFE_ASSERT(p + (u32)__popc(word) <= FILL_WORDS * 32u);      // the scan's prefixes are consistent with the bits
    if (word) {
        const u32 row = fd_div(i, d_wpr);
        const u32 w = i - row * wpr;
        const u32 base = row * rowlen + w * 32u;
        u32* q = sm + off + p;
        while (word) {
            const int b = __ffs((int)word) - 1;
            word &= word - 1;
            *q++ = base + (u32)b;
        }
    }
#if CFE_BULK_STORE
    fence_proxy_async_smem();
#endif
    __syncthreads();
    const u32 total = s_total;
    const u32 head = (4u - off) & 3u;
    const u32 h = head < total ? head : total;
    const u32 n4 = (total - h) >> 2;
    if (threadIdx.x < h) dst[threadIdx.x] = sm[off + threadIdx.x];
#if CFE_BULK_STORE
    if (threadIdx.x == 0 && n4) {
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                     :: "l"(dst + h), "r"((u32)__cvta_generic_to_shared(sm + off + h)), "r"(16u * n4) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
#else
    for (u32 k = threadIdx.x; k < n4; k += FILL_WORDS)
        reinterpret_cast<uint4*>(dst + h)[k] = reinterpret_cast<const uint4*>(sm + off + h)[k];
#endif
    for (u32 k = h + 4 * n4 + threadIdx.x; k < total; k += FILL_WORDS) dst[k] = sm[off + k];
#if CFE_BULK_STORE
    if (threadIdx.x == 0 && n4) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
#endif
}

extern "C" int cfe_fill_list(const uint32_t* bits, const uint32_t* pref, int64_t n_words,
                             int64_t words_per_row, int64_t row_len, uint32_t* list, cudaStream_t stream)
{
    if (n_words <= 0) return 0;
    if (n_words >= (1LL << 32)) { cfe_set_error("cfe_fill_list: too many words"); return -1; }
    const i64 blocks = (n_words + FILL_WORDS - 1) / FILL_WORDS;
    k_fill_list<<<(unsigned)blocks, FILL_WORDS, 0, stream>>>(bits, pref, (u32)n_words,
                                                             make_fastdiv((u32)words_per_row), (u32)words_per_row,
                                                             (u32)row_len, list);
    CFE_CHECK_LAUNCH("k_fill_list");
    return 0;
}

// Element list + line lengths in one pass (raster order): besides the voxel index of every element
// the kernel writes what the *ELEMENT emitter needs to place the element's text line
//   len8[j] = 9 + digits(eid) + sum of the digits of the 8 node ids     (voxelFE.py:647)
//   nd8[j]  = 0x80 | n when all eight node ids have n digits, else the digit count of the smallest
// A 32-voxel word holds x-consecutive voxels, so the digit counts are computed once per word (first
// and last set bit) and only re-derived per element when the ids straddle a power of ten.
#define FILLE_WORDS 128

// decimal digits of a 32-bit value without branches: floor(bits * log10(2)) is the digit count or one less
__constant__ u32 c_pow10[10] = {1u, 10u, 100u, 1000u, 10000u, 100000u, 1000000u, 10000000u, 100000000u, 1000000000u};
__device__ __forceinline__ int digits_u32_flat(u32 v)
{
    const int t = ((32 - __clz((int)(v | 1u))) * 1233) >> 12;
    return t + ((v | 1u) >= c_pow10[t] ? 1 : 0);
}
__device__ __forceinline__ int nd64(u64 v) { return (v >> 32) ? digits_u64(v) : digits_u32_flat((u32)v); }

__global__ void __launch_bounds__(FILLE_WORDS)
k_fill_elements(const u32* __restrict__ bits, const u32* __restrict__ pref, u32 n_words, FastDiv d_wpr, u32 wpr,
                u32 X, FastDiv dY, u32 Y, i64 SN, i64 RN, i64 z0, i64 eid_offset, u32* __restrict__ list,
                uint8_t* __restrict__ len8, uint8_t* __restrict__ nd8, const uint8_t* __restrict__ vol,
                uint8_t* __restrict__ gv8)
{
    // every array is staged at an offset congruent (mod 16 bytes) with its destination, so that the aligned
    // middle of the CTA's span leaves in one bulk copy per array
    __shared__ __align__(16) u32 sm_[FILLE_WORDS * 32 + 4];
    __shared__ __align__(16) uint8_t sl_[FILLE_WORDS * 32 + 16];
    __shared__ __align__(16) uint8_t sn_[FILLE_WORDS * 32 + 16];
    __shared__ u32 s_total;
    const u32 i0 = blockIdx.x * FILLE_WORDS;
    const u32 i = i0 + threadIdx.x;
    const u32 base0 = __ldg(pref + i0);
    const u32 off4 = (u32)(((uintptr_t)(list + base0) >> 2) & 3u);
    const u32 offl = (u32)((uintptr_t)(len8 + base0) & 15u), offn = (u32)((uintptr_t)(nd8 + base0) & 15u);
    u32* const sm = sm_ + off4;
    uint8_t* const sl = sl_ + offl;
    uint8_t* const sn = sn_ + offn;
    u32 word = 0, p = 0;
    if (i < n_words) {
        word = __ldg(bits + i);
        p = __ldg(pref + i) - base0;
    }
    const u32 last = (n_words - i0 < FILLE_WORDS ? n_words - i0 : FILLE_WORDS) - 1u;
    if (threadIdx.x == last) s_total = p + (u32)__popc(word);
    CFE_ASSERT(p + (u32)__popc(word) <= FILLE_WORDS * 32u);
    if (word) {
        const u32 row = fd_div(i, d_wpr);
        const u32 w = i - row * wpr;
        const u32 s = fd_div(row, dY);
        const u32 r = row - s * Y;
        const u32 vbase = row * X + w * 32u;
        const u64 b0 = (u64)(SN * ((i64)s + z0) + RN * (i64)r + (i64)(w * 32u));
        const u64 eid0 = (u64)(eid_offset + (i64)base0 + (i64)p + 1);
        const int f = __ffs((int)word) - 1, l = 31 - __clz((int)word);
        const int nlo = nd64(b0 + f), nhi = nd64(b0 + l + SN + RN + 1);
        const int ne0 = nd64(eid0), ne1 = nd64(eid0 + (u64)__popc(word) - 1);
        if (nlo == nhi && ne0 == ne1) {
            // the usual case: every element of the word has the same line length
            const uint8_t len = (uint8_t)(9 + ne0 + 8 * nlo), flags = (uint8_t)(0x80 | nlo);
            while (word) {
                const int bit = __ffs((int)word) - 1;
                word &= word - 1;
                sm[p] = vbase + (u32)bit;
                sl[p] = len;
                sn[p] = flags;
                ++p;
            }
        } else {
            u32 k = 0;
            while (word) {
                const int bit = __ffs((int)word) - 1;
                word &= word - 1;
                const int nde = nd64(eid0 + k);
                const u64 b = b0 + bit;
                const int n0 = nd64(b), n7 = nd64(b + SN + RN + 1);
                u32 len, flags;
                if (n0 == n7) {
                    len = 9u + (u32)nde + 8u * (u32)n0;
                    flags = 0x80u | (u32)n0;
                } else {
                    len = 9u + (u32)nde + (u32)n0 + (u32)n7 + (u32)nd64(b + 1) + (u32)nd64(b + RN) +
                          (u32)nd64(b + RN + 1) + (u32)nd64(b + SN) + (u32)nd64(b + SN + 1) + (u32)nd64(b + SN + RN);
                    flags = (u32)n0;
                }
                sm[p] = vbase + (u32)bit;
                sl[p] = (uint8_t)len;
                sn[p] = (uint8_t)flags;
                ++p;
                ++k;
            }
        }
    }
#if CFE_BULK_STORE
    fence_proxy_async_smem();
#endif
    __syncthreads();
    const u32 total = s_total;
    u32* dst = list + base0;
#if CFE_BULK_STORE
    // heads / tails by threads, aligned middles by the bulk-copy engine
    const u32 h4 = min((4u - off4) & 3u, total), n4 = (total - h4) >> 2;
    const u32 hl = min((16u - offl) & 15u, total), nl = (total - hl) >> 4;
    const u32 hn = min((16u - offn) & 15u, total), nn = (total - hn) >> 4;
    if (threadIdx.x == 0) {
        if (n4)
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                         :: "l"(dst + h4), "r"((u32)__cvta_generic_to_shared(sm + h4)), "r"(16u * n4) : "memory");
        if (nl)
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                         :: "l"(len8 + base0 + hl), "r"((u32)__cvta_generic_to_shared(sl + hl)), "r"(16u * nl) : "memory");
        if (nn)
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                         :: "l"(nd8 + base0 + hn), "r"((u32)__cvta_generic_to_shared(sn + hn)), "r"(16u * nn) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    auto edges = [&](u32 q) {
        if (q < h4 || q >= h4 + 4u * n4) dst[q] = sm[q];
        if (q < hl || q >= hl + 16u * nl) len8[base0 + q] = sl[q];
        if (q < hn || q >= hn + 16u * nn) nd8[base0 + q] = sn[q];
    };
    if (threadIdx.x < 32u) {
        // heads are shorter than 16 entries, tails too: the first and the last 32 entries cover them
        const u32 qa = threadIdx.x, qb = total - 1u - threadIdx.x;
        if (qa < total) edges(qa);
        if (threadIdx.x < total && qb >= 32u) edges(qb);
    }
    if (gv8) {
        // compact grey values in element order (cell_GV, voxelFE.py:160): the CTA's voxels are a span
        // of 128 words = 4 KB of the volume, so these byte loads hit the sectors once
        for (u32 q = threadIdx.x; q < total; q += FILLE_WORDS) gv8[base0 + q] = __ldg(vol + sm[q]);
    }
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
#else
    for (u32 q = threadIdx.x; q < total; q += FILLE_WORDS) {
        const u32 v = sm[q];
        dst[q] = v;
        len8[base0 + q] = sl[q];
        nd8[base0 + q] = sn[q];
        if (gv8) gv8[base0 + q] = __ldg(vol + v);
    }
#endif
}

extern "C" size_t cfe_emit_elements_workspace_bytes(int64_t n_el);

// Raster-ordered element list AND the emitter's per-line lengths (written into the element
// workspace `ws`, which cfe_element_lengths(have_lengths = 1) / cfe_emit_elements then use).
extern "C" int cfe_fill_elements(const uint32_t* bits, const uint32_t* pref, int64_t n_words, int64_t Y, int64_t X,
                                 int64_t z0, int64_t eid_offset, int64_t n_el, uint32_t* list, void* ws,
                                 cudaStream_t stream)
{
    return cfe_fill_elements_gv(bits, pref, n_words, Y, X, z0, eid_offset, n_el, list, ws, nullptr, nullptr, stream);
}

// Same, and gv8[j] = vol[list[j]]: the grey value of every element in element order (vol = the slab the
// bits were packed from).  The GV histogram and the GV-major partition then read n_el consecutive bytes
// instead of gathering from the volume.
extern "C" int cfe_fill_elements_gv(const uint32_t* bits, const uint32_t* pref, int64_t n_words, int64_t Y,
                                    int64_t X, int64_t z0, int64_t eid_offset, int64_t n_el, uint32_t* list,
                                    void* ws, const uint8_t* vol, uint8_t* gv8, cudaStream_t stream)
{
    if (n_words <= 0 || n_el <= 0) return 0;
    if (n_words >= (1LL << 32)) { cfe_set_error("cfe_fill_elements: too many words"); return -1; }
    const u32 wpr = (u32)((X + 31) / 32);
    uint8_t* len8 = reinterpret_cast<uint8_t*>(ws);
    uint8_t* nd8 = len8 + (((size_t)n_el + 15) & ~(size_t)15);
    const i64 blocks = (n_words + FILLE_WORDS - 1) / FILLE_WORDS;
    k_fill_elements<<<(unsigned)blocks, FILLE_WORDS, 0, stream>>>(
        bits, pref, (u32)n_words, make_fastdiv(wpr), wpr, (u32)X, make_fastdiv((u32)Y), (u32)Y, (X + 1) * (Y + 1),
        X + 1, z0, eid_offset, list, len8, nd8, vol, gv8);
    CFE_CHECK_LAUNCH("k_fill_elements");
    return 0;
}

// ------------------------------------------------------------------------------------------
// K6: boundary sets (voxelFE.py:163-192, 205-251).  Every set is "candidates in a box of the
// node / voxel grid, kept when a predicate holds", enumerated in raster order (ascending id).
// Two passes over the candidates (count per 256-candidate tile -> flat scan -> fill).

struct SetGeom {
    int Zl, Y, X, W;
    i64 z0;                  // global slice index of local slice 0
    i64 eid_offset;          // elements in the slabs below (global element id = offset + local + 1)
    const u32* bits;
    const u32* pref;         // element prefix per word (kind 2)
    const u32* halo_below;   // [Y][W] or null
};

__device__ __forceinline__ bool voxel_bit(const SetGeom& g, int s, int r, int c)
{
    // s is a LOCAL slice index; -1 addresses the halo plane
    if (r < 0 || r >= g.Y || c < 0 || c >= g.X) return false;
    const u32* plane;
    if (s >= 0 && s < g.Zl) plane = g.bits + (size_t)s * g.Y * g.W;
    else if (s == -1 && g.halo_below) plane = g.halo_below;
    else return false;
    return (__ldg(plane + (size_t)r * g.W + (c >> 5)) >> (c & 31)) & 1u;
}

// candidate at (ks, kr, kc) of the box (box-relative coordinates) -> keep?, id
__device__ __forceinline__ bool set_eval(const CfeSetDesc& d, const SetGeom& g, u32 ks, u32 kr, u32 kc, i64& id)
{
    const i64 s = d.s0 + (i64)ks;                 // GLOBAL coordinates
    const int r = (int)d.r0 + (int)kr;
    const int c = (int)d.c0 + (int)kc;
    const int sl = (int)(s - g.z0);
    if (d.kind == 2) {
        if (!voxel_bit(g, sl, r, c)) return false;
        const size_t wi = ((size_t)sl * g.Y + r) * g.W + (c >> 5);
        const u32 word = __ldg(g.bits + wi);
        id = g.eid_offset + (i64)__ldg(g.pref + wi) + __popc(word & mask_lt(c & 31)) + 1;
        return true;
    }
    // node (s, r, c): does one of its (up to) 8 voxels inside the voxel box exist?  The two x-neighbours
    // c - 1, c of a row sit in one word unless c is a multiple of 32: four row probes instead of eight
    // voxel probes.
    id = ((i64)(g.X + 1) * (g.Y + 1)) * s + (i64)(g.X + 1) * r + c;
    const int vlo = (int)(d.vc0 > 0 ? d.vc0 : 0), vhi = (int)(d.vc1 < g.X ? d.vc1 : g.X);   // columns [vlo, vhi)
    const int ca = (c - 1 > vlo) ? c - 1 : vlo;
    const int cb = (c < vhi - 1) ? c : vhi - 1;
    if (ca > cb) return false;
    bool hit = false;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const i64 vs = s - (t >> 1);
        const int vr = r - (t & 1);
        if (vs < d.vs0 || vs >= d.vs1 || vr < d.vr0 || vr >= d.vr1 || vr < 0 || vr >= g.Y) continue;
        const int vsl = (int)(vs - g.z0);
        const u32* plane;
        if (vsl >= 0 && vsl < g.Zl) plane = g.bits + (size_t)vsl * g.Y * g.W;
        else if (vsl == -1 && g.halo_below) plane = g.halo_below;
        else continue;
        const u32* rowp = plane + (size_t)vr * g.W;
        const u32 wa = __ldg(rowp + (ca >> 5));
        u32 m = (wa >> (ca & 31)) & 1u;
        if (cb != ca) {
            const u32 wb = ((cb >> 5) == (ca >> 5)) ? wa : __ldg(rowp + (cb >> 5));
            m |= (wb >> (cb & 31)) & 1u;
        }
        hit |= (m != 0);
    }
    return hit;
}

// Eight candidates (ks, kr, kc .. kc + 7) of one box row at once: the same predicates as set_eval, evaluated on a
// window of the voxel rows' words instead of voxel by voxel (a thread of k_sets_onepass owns eight consecutive
// candidates; one by one they probe the same words 30-60 times).  Returns false when the case is left to set_eval.
__device__ __forceinline__ bool set_eval8(const CfeSetDesc& d, const SetGeom& g, u32 ks, u32 kr, u32 kc,
                                          i64 (&id)[8], u32& keep)
{
    const i64 s = d.s0 + (i64)ks;                 // GLOBAL coordinates
    const int r = (int)d.r0 + (int)kr;
    const int c0 = (int)d.c0 + (int)kc;
    keep = 0;
    if (d.kind == 2) {
        const i64 sl = s - g.z0;
        if (sl < 0 || sl >= g.Zl || c0 < 0) return false;
#pragma unroll
        for (int j = 0; j < 8; ++j) id[j] = 0;
        if (r < 0 || r >= g.Y || c0 >= g.X) return true;            // no voxel there: nothing kept
        const int w = c0 >> 5, sh = c0 & 31;
        const size_t wi = ((size_t)sl * g.Y + r) * g.W + w;
        const u32 w0 = __ldg(g.bits + wi);
        const u32 w1 = (sh > 24 && w + 1 < g.W) ? __ldg(g.bits + wi + 1) : 0u;
        u32 win = __funnelshift_r(w0, w1, sh) & 0xffu;              // voxels c0 .. c0 + 7
        if (c0 + 8 > g.X) win &= (1u << (g.X - c0)) - 1u;           // (pad bits are 0 anyway)
        if (win) {
            const i64 base = g.eid_offset + (i64)__ldg(g.pref + wi) + __popc(w0 & mask_lt(sh)) + 1;
#pragma unroll
            for (int j = 0; j < 8; ++j) id[j] = base + __popc(win & ((1u << j) - 1u));
        }
        keep = win;
        return true;
    }
    // nodes c0 .. c0 + 7 of node row (s, r): node c exists when one of the voxels {c - 1, c} x {r - 1, r} x {s - 1, s}
    // inside the voxel box is set
    const i64 id0 = ((i64)(g.X + 1) * (g.Y + 1)) * s + (i64)(g.X + 1) * r + c0;
#pragma unroll
    for (int j = 0; j < 8; ++j) id[j] = id0 + j;
    const int vlo = (int)(d.vc0 > 0 ? d.vc0 : 0), vhi = (int)(d.vc1 < g.X ? d.vc1 : g.X);   // columns [vlo, vhi)
    const int cl = c0 - 1;                                          // column of window bit 0
    const int a = vlo - cl > 0 ? vlo - cl : 0, b = vhi - cl < 9 ? vhi - cl : 9;
    if (a >= b) return true;
    const u32 colmask = ((1u << b) - 1u) & ~((1u << a) - 1u);
    u32 win = 0;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const i64 vs = s - (t >> 1);
        const int vr = r - (t & 1);
        if (vs < d.vs0 || vs >= d.vs1 || vr < d.vr0 || vr >= d.vr1 || vr < 0 || vr >= g.Y) continue;
        const i64 vsl = vs - g.z0;
        const u32* plane;
        if (vsl >= 0 && vsl < g.Zl) plane = g.bits + (size_t)vsl * g.Y * g.W;
        else if (vsl == -1 && g.halo_below) plane = g.halo_below;
        else continue;
        const u32* rowp = plane + (size_t)vr * g.W;
        u32 x;
        if (cl < 0) {
            x = __ldg(rowp) << 1;                                   // window bit 0 = column -1
        } else {
            const int w = cl >> 5, sh = cl & 31;
            const u32 w0 = w < g.W ? __ldg(rowp + w) : 0u;
            const u32 w1 = (sh > 23 && w + 1 < g.W) ? __ldg(rowp + w + 1) : 0u;
            x = __funnelshift_r(w0, w1, sh);
        }
        win |= x;
    }
    win &= colmask;
    keep = (win | (win >> 1)) & 0xffu;
    return true;
}

// candidate k of the box -> box-relative (ks, kr, kc)
// (a box never has 2^32 candidates: it is a subset of one slab's node grid)
__device__ __forceinline__ void set_decode(const CfeSetDesc& d, i64 k, u32& ks, u32& kr, u32& kc)
{
    const u32 nc = (u32)(d.c1 - d.c0), plane = (u32)(d.r1 - d.r0) * nc;
    const u32 k32 = (u32)k;
    ks = k32 / plane;
    const u32 rem = k32 - ks * plane;
    kr = rem / nc;
    kc = rem - kr * nc;
}

__device__ __forceinline__ bool set_candidate(const CfeSetDesc& d, const SetGeom& g, i64 k, i64& id)
{
    u32 ks, kr, kc;
    set_decode(d, k, ks, kr, kc);
    return set_eval(d, g, ks, kr, kc, id);
}

__device__ __forceinline__ int find_set(const CfeSetDesc* descs, int n_sets, i64 tile)
{
    int q = 0;
    while (q + 1 < n_sets && tile >= descs[q + 1].tile_base) ++q;
    return q;
}

template <bool FILL>
__global__ void __launch_bounds__(256)
k_sets(const CfeSetDesc* __restrict__ descs, int n_sets, SetGeom g, u32* __restrict__ tile_count,
       const u64* __restrict__ tile_off, i64* __restrict__ ids)
{
    __shared__ u32 s_scan[33];
    const i64 tile = blockIdx.x;
    const int q = find_set(descs, n_sets, tile);
    const CfeSetDesc d = descs[q];
    const i64 k = (tile - d.tile_base) * 256 + threadIdx.x;
    i64 id = 0;
    bool keep = false;
    if (k < d.n_cand) keep = set_candidate(d, g, k, id);
    if (!FILL) {
        const int n = __syncthreads_count(keep);
        if (threadIdx.x == 0) tile_count[tile] = (u32)n;
    } else {
        u32 tot;
        const u32 e = block_exclusive_scan<u32>(keep ? 1u : 0u, s_scan, tot);
        if (keep) ids[tile_off[tile] + e] = id;
    }
}

// Single-pass form: every tile (SETS1_TILE candidates, tile_base in those units) evaluates its predicates ONCE, gets its global position
// by decoupled look-back over the tiles (set-major, raster order inside a set) and writes its ids.
// first[q] = position of set q's first id (only written for sets that own at least one tile),
// first[n_sets] = total number of ids.  `ids` must hold sum(n_cand) entries.
#define SETS1_ITEMS 8
#define SETS1_TILE (256 * SETS1_ITEMS)

extern "C" int64_t cfe_sets_onepass_tile(void) { return SETS1_TILE; }

__global__ void __launch_bounds__(256, 4)
k_sets_onepass(const CfeSetDesc* __restrict__ descs, int n_sets, i64 n_tiles, SetGeom g, u64* __restrict__ ws,
               i64* __restrict__ ids, i64* __restrict__ first)
{
    __shared__ u32 s_scan[33];
    __shared__ u64 s_excl;
    const i64 tile = scan_next_tile(ws);
    const int q = find_set(descs, n_sets, tile);
    const CfeSetDesc d = descs[q];
    // thread t owns candidates k0 .. k0+7 of its set (tile_base counts tiles of SETS1_TILE candidates)
    const i64 k0 = (tile - d.tile_base) * SETS1_TILE + (i64)threadIdx.x * SETS1_ITEMS;
    i64 id[SETS1_ITEMS];
    u32 keep = 0;
    if (k0 < d.n_cand) {
        // one decode per thread; the following candidates advance (c, r, s) with carries
        const u32 nc = (u32)(d.c1 - d.c0), nr = (u32)(d.r1 - d.r0);
        u32 ks, kr, kc;
        set_decode(d, k0, ks, kr, kc);
        if (kc + SETS1_ITEMS <= nc && k0 + SETS1_ITEMS <= d.n_cand && set_eval8(d, g, ks, kr, kc, id, keep)) {
            // the eight candidates sit in one row of the box: evaluated together from the rows' words
        } else {
            keep = 0;
#pragma unroll
            for (int j = 0; j < SETS1_ITEMS; ++j) {
                id[j] = 0;
                if (k0 + j < d.n_cand && set_eval(d, g, ks, kr, kc, id[j])) keep |= 1u << j;
                if (++kc == nc) { kc = 0; if (++kr == nr) { kr = 0; ++ks; } }
            }
        }
    }
    u32 tot;
    const u32 e = block_exclusive_scan<u32>((u32)__popc(keep), s_scan, tot);
    if (threadIdx.x < 32) {
        const u64 excl = scan_lookback_d<4>(ws, (u32)tile, (u64)tot);
        if (threadIdx.x == 0) {
            s_excl = excl;
            if (tile == d.tile_base) first[q] = (i64)excl;
            if (tile == n_tiles - 1) first[n_sets] = (i64)(excl + tot);
        }
    }
    __syncthreads();
    i64* dst = ids + s_excl + e;
#pragma unroll
    for (int j = 0; j < SETS1_ITEMS; ++j)
        if (keep & (1u << j)) *dst++ = id[j];
}

extern "C" int cfe_sets_onepass(const CfeSetDesc* descs_dev, int n_sets, int64_t n_tiles, int64_t Zl,
                                int64_t Y, int64_t X, int64_t z0, int64_t eid_offset, const uint32_t* bits,
                                const uint32_t* pref, const uint32_t* halo_below, void* ws, int64_t* ids,
                                int64_t* first, cudaStream_t stream)
{
    CFE_CHECK_CUDA(cudaMemsetAsync(first, 0xff, (size_t)(n_sets + 1) * sizeof(i64), stream));   // -1 = no tile
    if (n_tiles <= 0) {
        CFE_CHECK_CUDA(cudaMemsetAsync(first + n_sets, 0, sizeof(i64), stream));
        return 0;
    }
    if (n_tiles >= (1LL << 31)) { cfe_set_error("cfe_sets_onepass: too many tiles"); return -1; }
    CFE_CHECK_CUDA(cudaMemsetAsync(ws, 0, (size_t)(n_tiles + 2) * sizeof(u64), stream));
    SetGeom g;
    g.Zl = (int)Zl; g.Y = (int)Y; g.X = (int)X; g.W = (int)((X + 31) / 32);
    g.z0 = z0; g.eid_offset = eid_offset; g.bits = bits; g.pref = pref; g.halo_below = halo_below;
    k_sets_onepass<<<(unsigned)n_tiles, 256, 0, stream>>>(descs_dev, n_sets, n_tiles, g, reinterpret_cast<u64*>(ws),
                                                          reinterpret_cast<i64*>(ids), reinterpret_cast<i64*>(first));
    CFE_CHECK_LAUNCH("k_sets_onepass");
    return 0;
}

// Two-pass form, between count and fill: out[0] = stats[0] (n_el), out[1] = stats[1] (n_nd),
// out[2 + q] = number of ids of set q (difference of the scanned tile offsets at the sets' tile bases,
// stats[2] = total), out[2 + n_sets] = stats[3] (field overflow flag): the vector a z-slab rank
// all-gathers, assembled in one launch instead of a chain of tensor ops.
__global__ void k_set_counts(const CfeSetDesc* __restrict__ descs, int n_sets, i64 n_tiles,
                             const u64* __restrict__ tile_off, const u64* __restrict__ stats, i64* __restrict__ out)
{
    const int q = threadIdx.x;
    const u64 total = stats[2];
    if (q < n_sets) {
        const i64 t0 = descs[q].tile_base, t1 = q + 1 < n_sets ? descs[q + 1].tile_base : n_tiles;
        const u64 a = t0 < n_tiles ? tile_off[t0] : total, b = t1 < n_tiles ? tile_off[t1] : total;
        out[2 + q] = (i64)(b - a);
    }
    if (q == 0) { out[0] = (i64)stats[0]; out[1] = (i64)stats[1]; out[2 + n_sets] = (i64)stats[3]; }
}

extern "C" int cfe_set_counts(const CfeSetDesc* descs_dev, int n_sets, int64_t n_tiles, const uint64_t* tile_off,
                              const uint64_t* stats, int64_t* out, cudaStream_t stream)
{
    if (n_sets > 32) { cfe_set_error("cfe_set_counts: at most 32 sets"); return -1; }
    k_set_counts<<<1, 32, 0, stream>>>(descs_dev, n_sets, n_tiles, reinterpret_cast<const u64*>(tile_off),
                                       reinterpret_cast<const u64*>(stats), reinterpret_cast<i64*>(out));
    CFE_CHECK_LAUNCH("k_set_counts");
    return 0;
}

// generic exclusive scan u32 -> u64 (small arrays: tile counts, per-bin histograms)
__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_u32(const u32* __restrict__ in, i64 n, u64* __restrict__ out, u64* __restrict__ total, u64* __restrict__ ws)
{
    __shared__ u64 s_scan[33];
    __shared__ u64 s_excl;
    const u32 tile = scan_next_tile(ws);
    const i64 i0 = (i64)tile * SCAN_TILE + (i64)threadIdx.x * SCAN_ITEMS;
    u32 v[SCAN_ITEMS];
    u64 sum = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        v[k] = (i0 + k < n) ? in[i0 + k] : 0u;
        sum += v[k];
    }
    u64 block_total;
    u64 excl = block_exclusive_scan<u64>(sum, s_scan, block_total);
    if (threadIdx.x < 32) {
        u64 e = scan_lookback(ws, tile, block_total);
        if (threadIdx.x == 0) {
            s_excl = e;
            if ((i64)(tile + 1) * SCAN_TILE >= n) *total = e + block_total;
        }
    }
    __syncthreads();
    u64 run = s_excl + excl;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        if (i0 + k < n) out[i0 + k] = run;
        run += v[k];
    }
}

extern "C" int cfe_scan_u32(const uint32_t* in, int64_t n, uint64_t* out, uint64_t* total, void* ws,
                            cudaStream_t stream)
{
    const i64 tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    CFE_CHECK_CUDA(cudaMemsetAsync(ws, 0, (size_t)(tiles + 2) * sizeof(u64), stream));
    if (tiles == 0) {
        CFE_CHECK_CUDA(cudaMemsetAsync(total, 0, sizeof(u64), stream));
        return 0;
    }
    k_scan_u32<<<(unsigned)tiles, SCAN_THREADS, 0, stream>>>(in, n, reinterpret_cast<u64*>(out),
                                                             reinterpret_cast<u64*>(total),
                                                             reinterpret_cast<u64*>(ws));
    CFE_CHECK_LAUNCH("k_scan_u32");
    return 0;
}

extern "C" int cfe_sets_count(const CfeSetDesc* descs_dev, int n_sets, int64_t n_tiles, int64_t Zl,
                              int64_t Y, int64_t X, int64_t z0, int64_t eid_offset, const uint32_t* bits,
                              const uint32_t* pref, const uint32_t* halo_below, uint32_t* tile_count,
                              cudaStream_t stream)
{
    if (n_tiles <= 0) return 0;
    SetGeom g;
    g.Zl = (int)Zl; g.Y = (int)Y; g.X = (int)X; g.W = (int)((X + 31) / 32);
    g.z0 = z0; g.eid_offset = eid_offset; g.bits = bits; g.pref = pref; g.halo_below = halo_below;
    k_sets<false><<<(unsigned)n_tiles, 256, 0, stream>>>(descs_dev, n_sets, g, tile_count, nullptr, nullptr);
    CFE_CHECK_LAUNCH("k_sets<count>");
    return 0;
}

extern "C" int cfe_sets_fill(const CfeSetDesc* descs_dev, int n_sets, int64_t n_tiles, int64_t Zl,
                             int64_t Y, int64_t X, int64_t z0, int64_t eid_offset, const uint32_t* bits,
                             const uint32_t* pref, const uint32_t* halo_below, const uint64_t* tile_off,
                             int64_t* ids, cudaStream_t stream)
{
    if (n_tiles <= 0) return 0;
    SetGeom g;
    g.Zl = (int)Zl; g.Y = (int)Y; g.X = (int)X; g.W = (int)((X + 31) / 32);
    g.z0 = z0; g.eid_offset = eid_offset; g.bits = bits; g.pref = pref; g.halo_below = halo_below;
    k_sets<true><<<(unsigned)n_tiles, 256, 0, stream>>>(descs_dev, n_sets, g, nullptr,
                                                        reinterpret_cast<const u64*>(tile_off),
                                                        reinterpret_cast<i64*>(ids));
    CFE_CHECK_LAUNCH("k_sets<fill>");
    return 0;
}

// ------------------------------------------------------------------------------------------
// Materialisers for the lazy Mesh (meshio surface): cells (n_el,8) int64, GV, points.

__global__ void __launch_bounds__(256)
k_fill_cells(const u32* __restrict__ elem_vox, i64 n_el, int Y, int X, i64 z0, FastDiv dYX, FastDiv dX,
             i64* __restrict__ cells)
{
    const i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;  // one (element, node) pair per thread
    const i64 e = t >> 3;
    if (e >= n_el) return;
    const int j = (int)(t & 7);
    const u32 v = __ldg(elem_vox + e);
    const u32 s = fd_div(v, dYX);
    const u32 rem = v - s * (u32)(Y * X);
    const u32 r = fd_div(rem, dX);
    const u32 c = rem - r * (u32)X;
    const i64 RN = X + 1, SN = (i64)(X + 1) * (Y + 1);
    const i64 base = SN * ((i64)s + z0) + RN * r + c;
    // C3D8 order (voxelFE.py:145-154): b, b+1, b+RN+1, b+RN, b+SN, b+SN+1, b+SN+RN+1, b+SN+RN
    const i64 off = ((j >> 2) ? SN : 0) + (((j & 3) == 1 || (j & 3) == 2) ? 1 : 0) + (((j & 3) >= 2) ? RN : 0);
    cells[t] = base + off;
}

extern "C" int cfe_fill_cells(const uint32_t* elem_vox, int64_t n_el, int64_t Y, int64_t X, int64_t z0,
                              int64_t* cells, cudaStream_t stream)
{
    if (n_el <= 0) return 0;
    const i64 threads = n_el * 8;
    k_fill_cells<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(
        elem_vox, n_el, (int)Y, (int)X, z0, make_fastdiv((u32)(Y * X)), make_fastdiv((u32)X),
        reinterpret_cast<i64*>(cells));
    CFE_CHECK_LAUNCH("k_fill_cells");
    return 0;
}

__global__ void __launch_bounds__(256)
k_gather_u8(const uint8_t* __restrict__ vol, const u32* __restrict__ idx, i64 n, uint8_t* __restrict__ out)
{
    const i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) out[t] = __ldg(vol + __ldg(idx + t));
}

extern "C" int cfe_gather_gv(const uint8_t* vol, const uint32_t* elem_vox, int64_t n_el, uint8_t* gv,
                             cudaStream_t stream)
{
    if (n_el <= 0) return 0;
    k_gather_u8<<<(unsigned)((n_el + 255) / 256), 256, 0, stream>>>(vol, elem_vox, n_el, gv);
    CFE_CHECK_LAUNCH("k_gather_u8");
    return 0;
}

// points[(s,r,c)] = (xs[c], ys[r], zs[s]) copied as raw ELEM-byte values (float64/int64: 8,
// float32: 4) from the host-computed per-axis tables (voxelFE.py:202).
template <typename T>
__global__ void __launch_bounds__(256)
k_fill_points(const T* __restrict__ xs, const T* __restrict__ ys, const T* __restrict__ zs, i64 n_nodes,
              int RN, i64 SN, T* __restrict__ out)
{
    const i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_nodes) return;
    const i64 s = t / SN;
    const i64 rem = t - s * SN;
    const int r = (int)(rem / RN);
    const int c = (int)(rem - (i64)r * RN);
    out[3 * t + 0] = xs[c];
    out[3 * t + 1] = ys[r];
    out[3 * t + 2] = zs[s];
}

extern "C" int cfe_fill_points(const void* xs, const void* ys, const void* zs, int elem_bytes,
                               int64_t n_layers, int64_t Y, int64_t X, void* out, cudaStream_t stream)
{
    const i64 n = n_layers * (Y + 1) * (X + 1);
    if (n <= 0) return 0;
    const unsigned blocks = (unsigned)((n + 255) / 256);
    if (elem_bytes == 8)
        k_fill_points<u64><<<blocks, 256, 0, stream>>>((const u64*)xs, (const u64*)ys, (const u64*)zs, n,
                                                       (int)(X + 1), (X + 1) * (Y + 1), (u64*)out);
    else if (elem_bytes == 4)
        k_fill_points<u32><<<blocks, 256, 0, stream>>>((const u32*)xs, (const u32*)ys, (const u32*)zs, n,
                                                       (int)(X + 1), (X + 1) * (Y + 1), (u32*)out);
    else { cfe_set_error("cfe_fill_points: elem_bytes must be 4 or 8"); return -1; }
    CFE_CHECK_LAUNCH("k_fill_points");
    return 0;
}

// ------------------------------------------------------------------------------------------
// Reference-node barycentre sums (voxelFE.py:208-223): float64 accumulation in ascending node
// order, one thread per set so that the summation ORDER (hence every rounding) is the
// reference's.  acc[q][0..2] is read-modify-written so z-slabs can be chained.

__global__ void k_refnode_sums(const i64* __restrict__ ids, const i64* __restrict__ first,
                               const i64* __restrict__ count, int n_sets, i64 SN, i64 RN,
                               const double* __restrict__ xs, const double* __restrict__ ys,
                               const double* __restrict__ zs, double* __restrict__ acc)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_sets) return;
    double a0 = acc[3 * q], a1 = acc[3 * q + 1], a2 = acc[3 * q + 2];
    const i64* p = ids + first[q];
    const i64 n = count[q];
    for (i64 i = 0; i < n; ++i) {
        const i64 id = p[i];
        const i64 s = id / SN, rem = id - s * SN;
        const i64 r = rem / RN, c = rem - r * RN;
        a0 = __dadd_rn(a0, xs[c]);
        a1 = __dadd_rn(a1, ys[r]);
        a2 = __dadd_rn(a2, zs[s]);
    }
    acc[3 * q] = a0; acc[3 * q + 1] = a1; acc[3 * q + 2] = a2;
}

extern "C" int cfe_refnode_sums(const int64_t* ids, const int64_t* first, const int64_t* count, int n_sets,
                                int64_t Y, int64_t X, const double* xs, const double* ys, const double* zs,
                                double* acc, cudaStream_t stream)
{
    if (n_sets <= 0) return 0;
    k_refnode_sums<<<1, 32, 0, stream>>>(reinterpret_cast<const i64*>(ids), reinterpret_cast<const i64*>(first),
                                         reinterpret_cast<const i64*>(count), n_sets, (X + 1) * (Y + 1), X + 1,
                                         xs, ys, zs, acc);
    CFE_CHECK_LAUNCH("k_refnode_sums");
    return 0;
}

// ------------------------------------------------------------------------------------------
// ParOSol deck (voxelFE.py:285-472).  The O(N^3) pieces of vol2h5ParOSol:
//   k_proj_bits       the three max-projections of recon_utils.bbox (recon_utils.py:258-272) from packed
//                     bits: "row has a set bit" per voxel row and the OR of all rows; one warp per row
//   k_scale_crop_f32  Image dataset  (sub_block * young_modulus).astype(float32)  (voxelFE.py:350-357):
//                     crop to the bounding box and map every byte through a 256-entry float table that
//                     the host fills with NumPy's own products, so the values are the reference's bits

#define PROJ_COLS 8      // column words kept in registers per lane (X <= 8192); wider rows use atomics directly

__global__ void __launch_bounds__(256)
k_proj_bits(const u32* __restrict__ bits, i64 rows, int W, uint8_t* __restrict__ rowany, u32* __restrict__ colbits)
{
    const u32 lane = threadIdx.x & 31u;
    const i64 warp0 = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const i64 nwarps = ((i64)gridDim.x * blockDim.x) >> 5;
    u32 acc[PROJ_COLS];
#pragma unroll
    for (int k = 0; k < PROJ_COLS; ++k) acc[k] = 0;
    for (i64 row = warp0; row < rows; row += nwarps) {
        const u32* p = bits + row * W;
        u32 any = 0;
#pragma unroll
        for (int k = 0; k < PROJ_COLS; ++k) {
            const int w = (int)lane + 32 * k;
            if (w < W) {
                const u32 word = __ldg(p + w);
                acc[k] |= word;
                any |= word;
            }
        }
        for (int w = (int)lane + 32 * PROJ_COLS; w < W; w += 32) {
            const u32 word = __ldg(p + w);
            if (word) atomicOr(colbits + w, word);
            any |= word;
        }
        const bool row_any = __any_sync(CFE_FULL_MASK, any != 0);
        if (lane == 0) rowany[row] = row_any ? 1 : 0;
    }
#pragma unroll
    for (int k = 0; k < PROJ_COLS; ++k) {
        const int w = (int)lane + 32 * k;
        if (w < W && acc[k]) atomicOr(colbits + w, acc[k]);
    }
}

// Projections of the packed bits of a (rows, X) volume: rowany[row] = 1 when the row has a set bit,
// colbits = OR of all rows (W = ceil(X / 32) words).  bits = cfe_pack_bits of the volume.
extern "C" int cfe_proj_bits(const uint32_t* bits, int64_t rows, int64_t X, uint8_t* rowany, uint32_t* colbits,
                             cudaStream_t stream)
{
    if (X <= 0) return 0;
    const int W = (int)((X + 31) / 32);
    CFE_CHECK_CUDA(cudaMemsetAsync(colbits, 0, (size_t)W * sizeof(u32), stream));
    if (rows <= 0) return 0;
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    i64 blocks = (rows + 7) / 8;
    if (blocks > (i64)sms * 8) blocks = (i64)sms * 8;
    k_proj_bits<<<(unsigned)blocks, 256, 0, stream>>>(bits, rows, W, rowany, colbits);
    CFE_CHECK_LAUNCH("k_proj_bits");
    return 0;
}

// one warp per row of the cropped block; a lane converts 4 consecutive voxels (one 32-bit load when the
// source address allows, one 16-byte store when the output row is 16-byte aligned)
__global__ void __launch_bounds__(256)
k_scale_crop_f32(const uint8_t* __restrict__ vol, i64 sY, i64 sX, i64 z0, i64 y0, i64 x0, i64 rows, FastDiv dny,
                 u32 ny, i64 nx, const float* __restrict__ lut, float* __restrict__ out)
{
    __shared__ float t[256];
    t[threadIdx.x] = __ldg(lut + threadIdx.x);
    __syncthreads();
    const u32 lane = threadIdx.x & 31u;
    const i64 warp0 = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const i64 nwarps = ((i64)gridDim.x * blockDim.x) >> 5;
    const bool vec_out = ((nx & 3) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15u) == 0);
    for (i64 row = warp0; row < rows; row += nwarps) {
        const u32 z = fd_div((u32)row, dny);
        const u32 y = (u32)row - z * ny;
        const uint8_t* src = vol + ((z0 + z) * sY + (y0 + y)) * sX + x0;
        float* dst = out + row * nx;
        const bool vec_in = (reinterpret_cast<uintptr_t>(src) & 3u) == 0;
        const i64 nx4 = nx & ~(i64)3;
#pragma unroll 4
        for (i64 x = (i64)lane * 4; x < nx4; x += 128) {
            u32 w;
            if (vec_in) {
                w = __ldg(reinterpret_cast<const u32*>(src + x));
            } else {
                w = (u32)__ldg(src + x) | ((u32)__ldg(src + x + 1) << 8) | ((u32)__ldg(src + x + 2) << 16) |
                    ((u32)__ldg(src + x + 3) << 24);
            }
            const float4 f = make_float4(t[w & 255u], t[(w >> 8) & 255u], t[(w >> 16) & 255u], t[w >> 24]);
            if (vec_out) {
                *reinterpret_cast<float4*>(dst + x) = f;
            } else {
                dst[x] = f.x; dst[x + 1] = f.y; dst[x + 2] = f.z; dst[x + 3] = f.w;
            }
        }
        for (i64 x = nx4 + lane; x < nx; x += 32) dst[x] = t[__ldg(src + x)];
    }
}

// out[z][y][x] = lut[vol[z0 + z][y0 + y][x0 + x]] for the (nz, ny, nx) block of the (., Y, X) volume
extern "C" int cfe_scale_crop_f32(const uint8_t* vol, int64_t Y, int64_t X, int64_t z0, int64_t y0, int64_t x0,
                                  int64_t nz, int64_t ny, int64_t nx, const float* lut256, float* out,
                                  cudaStream_t stream)
{
    if (nz <= 0 || ny <= 0 || nx <= 0) return 0;
    const i64 rows = nz * ny;
    if (rows >= (1LL << 32) || ny >= (1LL << 32)) { cfe_set_error("cfe_scale_crop_f32: block too large"); return -1; }
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    i64 blocks = (rows + 7) / 8;
    if (blocks > (i64)sms * 8) blocks = (i64)sms * 8;
    k_scale_crop_f32<<<(unsigned)blocks, 256, 0, stream>>>(vol, Y, X, z0, y0, x0, rows, make_fastdiv((u32)ny),
                                                           (u32)ny, nx, lut256, out);
    CFE_CHECK_LAUNCH("k_scale_crop_f32");
    return 0;
}

// ------------------------------------------------------------------------------------------
// segment (preprocess.py:20-46):  image > T  for uint8 images, 16 voxels per thread (borrow-free SWAR
// compare, one 16-byte load and one 16-byte store), bool bytes 0 / 1 out.

__device__ __forceinline__ u32 swar_gt_u8(u32 x, u32 t4)
{
    // per byte: 0x01 where x > t (unsigned); t4 = t replicated.  x > t  <=>  carry out of  x + (255 - t)
    // computed without inter-byte carries: low 7 bits added separately, top bits combined by logic
    const u32 H = 0x80808080u;
    const u32 y = ~t4;                                   // 255 - t per byte
    const u32 low = (x & ~H) + (y & ~H);                 // bit 7 = carry of the low 7 bits
    const u32 carry = ((x & y) | ((x | y) & low)) & H;   // carry out of bit 7
    return carry >> 7;
}

__global__ void __launch_bounds__(256)
k_threshold_u8(const uint8_t* __restrict__ img, i64 n, u32 t4, int t, uint8_t* __restrict__ out)
{
    const i64 n16 = n >> 4;
    for (i64 g = (i64)blockIdx.x * blockDim.x + threadIdx.x; g < n16; g += (i64)gridDim.x * blockDim.x) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(img) + g);
        uint4 r;
        r.x = swar_gt_u8(v.x, t4); r.y = swar_gt_u8(v.y, t4); r.z = swar_gt_u8(v.z, t4); r.w = swar_gt_u8(v.w, t4);
        reinterpret_cast<uint4*>(out)[g] = r;
    }
    for (i64 i = (n16 << 4) + (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x)
        out[i] = (uint8_t)((int)img[i] > t);
}

// out[i] = img[i] > t (bytes 0 / 1).  t outside 0..254 gives the constant answer.  img and out 16-byte aligned.
extern "C" int cfe_threshold_u8(const uint8_t* img, int64_t n, int t, uint8_t* out, cudaStream_t stream)
{
    if (n <= 0) return 0;
    if (t < 0) { CFE_CHECK_CUDA(cudaMemsetAsync(out, 1, (size_t)n, stream)); return 0; }
    if (t >= 255) { CFE_CHECK_CUDA(cudaMemsetAsync(out, 0, (size_t)n, stream)); return 0; }
    if ((reinterpret_cast<uintptr_t>(img) | reinterpret_cast<uintptr_t>(out)) & 15u) {
        cfe_set_error("cfe_threshold_u8: buffers must be 16-byte aligned");
        return -1;
    }
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    i64 blocks = (n / 16 + 255) / 256;
    if (blocks > (i64)sms * 16) blocks = (i64)sms * 16;
    if (blocks < 1) blocks = 1;
    k_threshold_u8<<<(unsigned)blocks, 256, 0, stream>>>(img, n, 0x01010101u * (u32)t, t, out);
    CFE_CHECK_LAUNCH("k_threshold_u8");
    return 0;
}

// ------------------------------------------------------------------------------------------
// embed (preprocess.py:233-355): the embedding mask before the emboss,  box & ~(I > 0)  (:327), and the
// final  I[BW_embedding] = embed_val  (:331-337).  Both are one pass over the volume, 16 voxels per thread.

__global__ void __launch_bounds__(256)
k_box_andnot_u8(const uint8_t* __restrict__ img, i64 n_rows, FastDiv dY, u32 Y, int X, int z0, int z1, int y0,
                int y1, int x0, int x1, uint8_t* __restrict__ out)
{
    // one warp per row; inside the box a voxel is 1 when the image is 0
    const u32 lane = threadIdx.x & 31u;
    const i64 warp0 = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const i64 nwarps = ((i64)gridDim.x * blockDim.x) >> 5;
    for (i64 row = warp0; row < n_rows; row += nwarps) {
        const u32 z = fd_div((u32)row, dY);
        const u32 y = (u32)row - z * Y;
        const bool row_in = (int)z >= z0 && (int)z < z1 && (int)y >= y0 && (int)y < y1;
        const uint8_t* src = img + row * X;
        uint8_t* dst = out + row * X;
        for (int x = (int)lane; x < X; x += 32)
            dst[x] = (row_in && x >= x0 && x < x1 && __ldg(src + x) == 0) ? 1 : 0;
    }
}

// out = 1 where (z, y, x) lies in [z0,z1) x [y0,y1) x [x0,x1) and img == 0, else 0
extern "C" int cfe_box_andnot_u8(const uint8_t* img, int64_t Z, int64_t Y, int64_t X, int64_t z0, int64_t z1,
                                 int64_t y0, int64_t y1, int64_t x0, int64_t x1, uint8_t* out, cudaStream_t stream)
{
    const i64 rows = Z * Y;
    if (rows <= 0 || X <= 0) return 0;
    if (rows >= (1LL << 32) || X >= (1LL << 31)) { cfe_set_error("cfe_box_andnot_u8: volume too large"); return -1; }
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    i64 blocks = (rows + 7) / 8;
    if (blocks > (i64)sms * 8) blocks = (i64)sms * 8;
    k_box_andnot_u8<<<(unsigned)blocks, 256, 0, stream>>>(img, rows, make_fastdiv((u32)Y), (u32)Y, (int)X, (int)z0,
                                                         (int)z1, (int)y0, (int)y1, (int)x0, (int)x1, out);
    CFE_CHECK_LAUNCH("k_box_andnot_u8");
    return 0;
}

// KEEP = false: out = mask ? val : img   (I[mask] = val);   KEEP = true: out = mask ? img : 0   (I[~mask] = 0)
template <bool KEEP>
__global__ void __launch_bounds__(256)
k_masked_set_u8(const uint8_t* __restrict__ img, const uint8_t* __restrict__ mask, i64 n, u32 val4,
                uint8_t* __restrict__ out)
{
    const i64 n16 = n >> 4;
    for (i64 g = (i64)blockIdx.x * blockDim.x + threadIdx.x; g < n16; g += (i64)gridDim.x * blockDim.x) {
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(img) + g);
        const uint4 m = __ldg(reinterpret_cast<const uint4*>(mask) + g);
        uint4 r;
        // mask bytes are 0 / 1: 0xff per selected byte
        const u32 mx = m.x * 0xffu, my = m.y * 0xffu, mz = m.z * 0xffu, mw = m.w * 0xffu;
        if (KEEP) {
            r.x = a.x & mx; r.y = a.y & my; r.z = a.z & mz; r.w = a.w & mw;
        } else {
            r.x = (a.x & ~mx) | (val4 & mx); r.y = (a.y & ~my) | (val4 & my);
            r.z = (a.z & ~mz) | (val4 & mz); r.w = (a.w & ~mw) | (val4 & mw);
        }
        reinterpret_cast<uint4*>(out)[g] = r;
    }
    for (i64 i = (n16 << 4) + (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x)
        out[i] = KEEP ? (mask[i] ? img[i] : (uint8_t)0) : (mask[i] ? (uint8_t)val4 : img[i]);
}

static int masked_launch(bool keep, const uint8_t* img, const uint8_t* mask, int64_t n, int val, uint8_t* out,
                         cudaStream_t stream, const char* who)
{
    if (n <= 0) return 0;
    if ((reinterpret_cast<uintptr_t>(img) | reinterpret_cast<uintptr_t>(mask) | reinterpret_cast<uintptr_t>(out)) & 15u) {
        cfe_set_error("%s: buffers must be 16-byte aligned", who);
        return -1;
    }
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    i64 blocks = (n / 16 + 255) / 256;
    if (blocks > (i64)sms * 16) blocks = (i64)sms * 16;
    if (blocks < 1) blocks = 1;
    if (keep) k_masked_set_u8<true><<<(unsigned)blocks, 256, 0, stream>>>(img, mask, n, 0u, out);
    else k_masked_set_u8<false><<<(unsigned)blocks, 256, 0, stream>>>(img, mask, n, 0x01010101u * (u32)(val & 255), out);
    CFE_CHECK_LAUNCH("k_masked_set_u8");
    return 0;
}

// out[i] = mask[i] ? val : img[i]  (mask bytes 0 / 1; out may alias img; all 16-byte aligned)
extern "C" int cfe_masked_set_u8(const uint8_t* img, const uint8_t* mask, int64_t n, int val, uint8_t* out,
                                 cudaStream_t stream)
{
    return masked_launch(false, img, mask, n, val, out, stream, "cfe_masked_set_u8");
}

// out[i] = mask[i] ? img[i] : 0  -- `masked = I; masked[~L] = 0` of the grey-scale branch (pipeline.py:240-242)
extern "C" int cfe_masked_keep_u8(const uint8_t* img, const uint8_t* mask, int64_t n, uint8_t* out, cudaStream_t stream)
{
    return masked_launch(true, img, mask, n, 0, out, stream, "cfe_masked_keep_u8");
}
