// ciclope_b200 -- CalculiX/Abaqus deck emission kernels (sm_100a).
//
// Replaces the str.format + file.write loops of ciclope/core/voxelFE.py:624-688:
//   *NODE lines      '{0:10d}, {n[0]:12.6f}, {n[1]:12.6f}, {n[2]:12.6f}\n'      (:630)
//   *ELEMENT lines   '{eid},{n0},...,{n7}\n' grouped by grey value               (:639-647)
//   *NSET / *ELSET   id lists, ten per line                                       (:650-688)
// plus the grey-value histogram / stable partition that np.unique(cell_GV) and
// np.where(cell_GV == GV) (:570, :645) stand for.
//
// Every emitter formats a fixed number of items per CTA into a shared-memory staging buffer
// that is congruent (mod 16) with its destination span, then streams the span to HBM with
// 16-byte stores.  Variable-length sections get their byte offsets from a single-pass
// decoupled look-back scan inside the emitting kernel (no separate length pass).
#include "common.cuh"
#include "ciclope_b200.h"

#define EMIT_THREADS 256

// ------------------------------------------------------------------------------------------
// staging helpers

// write the decimal digits of v so that the last digit lands at p[-1]; returns nothing.
__device__ __forceinline__ void put_u32_back(char* p, u32 v, int nd)
{
#pragma unroll 1
    for (int k = 0; k < nd; ++k) {
        const u32 q = v / 10u;
        *--p = (char)('0' + (v - q * 10u));
        v = q;
    }
}

__device__ __forceinline__ void put_u64_back(char* p, u64 v, int nd)
{
    if (v < 4294967296ull) {
        put_u32_back(p, (u32)v, nd);
    } else {
        const u64 hi = v / 1000000000ull;
        const u32 lo = (u32)(v - hi * 1000000000ull);
        put_u32_back(p, lo, 9);
        put_u32_back(p - 9, (u32)hi, nd - 9);
    }
}

// Copy the staged span [pad, pad + n) of `stage` to dst (global).  `stage` is 16-byte aligned
// and pad == (uintptr_t)dst & 15, so stage + 16k and dst - pad + 16k are both 16-byte aligned.
__device__ __forceinline__ void flush_stage(const char* stage, u32 pad, u32 n, char* dst)
{
    char* gbase = dst - pad;
    const u32 end = pad + n;
    const u32 nchunks = (end + 15u) >> 4;
    for (u32 c = threadIdx.x; c < nchunks; c += blockDim.x) {
        const u32 lo = c << 4, hi = lo + 16u;
        if (lo >= pad && hi <= end) {
            *reinterpret_cast<uint4*>(gbase + lo) = *reinterpret_cast<const uint4*>(stage + lo);
        } else {
            const u32 a = lo < pad ? pad : lo, b = hi > end ? end : hi;
            for (u32 p = a; p < b; ++p) gbase[p] = stage[p];
        }
    }
}

// ------------------------------------------------------------------------------------------
// *NODE section: fixed 53-byte lines, one thread per used node.

struct NodeEmitParams {
    const u32* node_list;    // local node-grid linear indices of the used nodes, ascending
    i64 n_nodes;
    const char* xtab;        // [X+1][16]: 12-char '{:12.6f}' field of each column coordinate
    const char* ytab;        // [Y+1][16]
    const char* ztab;        // [n_layers][16] (this slab's node layers)
    FastDiv dSN, dRN;
    u32 SN32, RN32;          // (X+1)(Y+1), X+1
    i64 z0;                  // global index of the slab's first node layer
    char* out;               // destination of the first line of this slab
};

__global__ void __launch_bounds__(EMIT_THREADS) k_emit_nodes(NodeEmitParams P)
{
    __shared__ __align__(16) char stage[EMIT_THREADS * 53 + 32];
    const i64 k0 = (i64)blockIdx.x * EMIT_THREADS;
    const i64 k = k0 + threadIdx.x;
    char* dst = P.out + k0 * 53;
    const u32 pad = (u32)(reinterpret_cast<uintptr_t>(dst) & 15u);
    const u32 n_here = (u32)((P.n_nodes - k0) < EMIT_THREADS ? (P.n_nodes - k0) : EMIT_THREADS);
    if (k < P.n_nodes) {
        const u32 v = __ldg(P.node_list + k);
        const u32 s = fd_div(v, P.dSN);
        const u32 rem = v - s * P.SN32;
        const u32 r = fd_div(rem, P.dRN);
        const u32 c = rem - r * P.RN32;
        const u64 id = (u64)P.SN32 * (u64)((i64)s + P.z0) + (u64)rem;
        char* line = stage + pad + threadIdx.x * 53;
        const int nd = digits_u64(id);
#pragma unroll
        for (int q = 0; q < 10; ++q) line[q] = ' ';
        put_u64_back(line + 10, id, nd);
        const uint4 fx = __ldg(reinterpret_cast<const uint4*>(P.xtab) + c);
        const uint4 fy = __ldg(reinterpret_cast<const uint4*>(P.ytab) + r);
        const uint4 fz = __ldg(reinterpret_cast<const uint4*>(P.ztab) + s);
        const u32 f[9] = {fx.x, fx.y, fx.z, fy.x, fy.y, fy.z, fz.x, fz.y, fz.z};
        char* p = line + 10;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            *p++ = ',';
            *p++ = ' ';
#pragma unroll
            for (int q = 0; q < 12; ++q) *p++ = (char)((f[a * 3 + (q >> 2)] >> (8 * (q & 3))) & 0xffu);
        }
        *p = '\n';
    }
    __syncthreads();
    flush_stage(stage, pad, n_here * 53u, dst);
}

extern "C" int cfe_emit_nodes(const uint32_t* node_list, int64_t n_nodes, const char* xtab, const char* ytab,
                              const char* ztab, int64_t Y, int64_t X, int64_t z0, char* out,
                              cudaStream_t stream)
{
    if (n_nodes <= 0) return 0;
    NodeEmitParams P;
    P.node_list = node_list; P.n_nodes = n_nodes; P.xtab = xtab; P.ytab = ytab; P.ztab = ztab;
    P.SN32 = (u32)((X + 1) * (Y + 1)); P.RN32 = (u32)(X + 1);
    P.dSN = make_fastdiv(P.SN32); P.dRN = make_fastdiv(P.RN32);
    P.z0 = z0; P.out = out;
    const i64 blocks = (n_nodes + EMIT_THREADS - 1) / EMIT_THREADS;
    k_emit_nodes<<<(unsigned)blocks, EMIT_THREADS, 0, stream>>>(P);
    CFE_CHECK_LAUNCH("k_emit_nodes");
    return 0;
}

// ------------------------------------------------------------------------------------------
// *ELEMENT section: variable-length lines, one thread per element, single-pass look-back.

#define ELEM_HDR_SLOT 64      // bytes reserved per '*ELEMENT, TYPE=..., ELSET=SET<gv>\n' header
#define ELEM_MAX_LINE 192     // eid (<= 20) + 8 node ids (<= 20 each) + 9 separators
#define ELEM_STAGE_BYTES (EMIT_THREADS * 100 + 4096 + 32)

struct ElemEmitParams {
    const u32* elem_vox;     // raster-ordered local voxel indices (PERM = false)
    const uint2* perm;       // GV-major order: {local voxel index, raster rank} (PERM = true)
    i64 n_el;
    const uint8_t* vol;      // grey values (only read when PERM)
    const u64* first;        // [256] index (in emission order) of the first element of each GV block
    const char* hdr_text;    // [256][ELEM_HDR_SLOT]
    const uint8_t* hdr_len;  // [256]; 0 = no header for this GV (e.g. slabs other than the first)
    int single_gv;           // PERM = false: the one grey value
    FastDiv dYX, dX;
    u32 YX32, X32;
    i64 SN, RN, z0, eid_offset;
    char* out;               // start of the element lines
    u64* total;              // out: bytes written
    u64* blk_off;            // out [256]: byte offset (from `out`) where each GV block (header included) starts
    u64* ws;                 // look-back workspace
};

template <bool PERM>
__global__ void __launch_bounds__(EMIT_THREADS) k_emit_elements(ElemEmitParams P)
{
    extern __shared__ __align__(16) char stage[];
    __shared__ u32 s_scan[33];
    __shared__ u64 s_base;
    const u32 tile = scan_next_tile(P.ws);
    const i64 j = (i64)tile * EMIT_THREADS + threadIdx.x;
    u64 val[9];
    int nd[9];
    u32 len = 0, hl = 0;
    bool is_first = false;
    int gv = P.single_gv;
    if (j < P.n_el) {
        u32 v, rank;
        if (PERM) { const uint2 e = __ldg(P.perm + j); v = e.x; rank = e.y; gv = (int)__ldg(P.vol + v); }
        else { v = __ldg(P.elem_vox + j); rank = (u32)j; }
        const u32 s = fd_div(v, P.dYX);
        const u32 rem = v - s * P.YX32;
        const u32 r = fd_div(rem, P.dX);
        const u32 c = rem - r * P.X32;
        const u64 b = (u64)(P.SN * ((i64)s + P.z0) + P.RN * (i64)r + (i64)c);
        val[0] = (u64)(P.eid_offset + (i64)rank + 1);
        val[1] = b;                 val[2] = b + 1;
        val[3] = b + P.RN + 1;      val[4] = b + P.RN;
        val[5] = b + P.SN;          val[6] = b + P.SN + 1;
        val[7] = b + P.SN + P.RN + 1; val[8] = b + P.SN + P.RN;
        len = 9;
#pragma unroll
        for (int q = 0; q < 9; ++q) { nd[q] = digits_u64(val[q]); len += (u32)nd[q]; }
        is_first = ((u64)j == __ldg(P.first + gv));
        if (is_first) hl = __ldg(P.hdr_len + gv);
    }
    u32 tile_total;
    const u32 excl = block_exclusive_scan<u32>(len + hl, s_scan, tile_total);
    if (threadIdx.x < 32) {
        const u64 e = scan_lookback(P.ws, tile, (u64)tile_total);
        if (threadIdx.x == 0) {
            s_base = e;
            if ((i64)(tile + 1) * EMIT_THREADS >= P.n_el) *P.total = e + tile_total;
        }
    }
    __syncthreads();
    char* dst = P.out + s_base;
    const u32 pad = (u32)(reinterpret_cast<uintptr_t>(dst) & 15u);
    if (is_first && P.blk_off) P.blk_off[gv] = s_base + excl;
    // A tile normally fits the staging buffer; a pathological tile (hundreds of block headers)
    // falls back to direct global writes.
    const bool staged = (pad + tile_total) <= (u32)ELEM_STAGE_BYTES;
    if (j < P.n_el) {
        char* p = staged ? (stage + pad + excl) : (dst + excl);
        if (hl) {
            const char* h = P.hdr_text + gv * ELEM_HDR_SLOT;
            for (u32 q = 0; q < hl; ++q) p[q] = __ldg(h + q);
            p += hl;
        }
#pragma unroll
        for (int q = 0; q < 9; ++q) {
            p += nd[q];
            put_u64_back(p, val[q], nd[q]);
            *p++ = (q == 8) ? '\n' : ',';
        }
    }
    __syncthreads();
    if (staged) flush_stage(stage, pad, tile_total, dst);
}

extern "C" int cfe_emit_elements(const uint32_t* elem_vox, const void* perm, int64_t n_el, const uint8_t* vol,
                                 const uint64_t* first, const char* hdr_text, const uint8_t* hdr_len,
                                 int single_gv, int64_t Y, int64_t X, int64_t z0, int64_t eid_offset,
                                 char* out, uint64_t* total, uint64_t* blk_off, void* ws, cudaStream_t stream)
{
    const i64 tiles = (n_el + EMIT_THREADS - 1) / EMIT_THREADS;
    CFE_CHECK_CUDA(cudaMemsetAsync(ws, 0, (size_t)(tiles + 2) * sizeof(u64), stream));
    CFE_CHECK_CUDA(cudaMemsetAsync(total, 0, sizeof(u64), stream));
    if (tiles == 0) return 0;
    ElemEmitParams P;
    P.elem_vox = elem_vox; P.perm = reinterpret_cast<const uint2*>(perm); P.n_el = n_el; P.vol = vol;
    P.first = reinterpret_cast<const u64*>(first); P.hdr_text = hdr_text; P.hdr_len = hdr_len;
    P.single_gv = single_gv;
    P.YX32 = (u32)(Y * X); P.X32 = (u32)X;
    P.dYX = make_fastdiv(P.YX32); P.dX = make_fastdiv(P.X32);
    P.RN = X + 1; P.SN = (X + 1) * (Y + 1); P.z0 = z0; P.eid_offset = eid_offset;
    P.out = out; P.total = reinterpret_cast<u64*>(total); P.blk_off = reinterpret_cast<u64*>(blk_off);
    P.ws = reinterpret_cast<u64*>(ws);
    static bool attr_set = false;
    if (!attr_set) {
        CFE_CHECK_CUDA(cudaFuncSetAttribute(k_emit_elements<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            ELEM_STAGE_BYTES));
        CFE_CHECK_CUDA(cudaFuncSetAttribute(k_emit_elements<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            ELEM_STAGE_BYTES));
        attr_set = true;
    }
    if (perm)
        k_emit_elements<true><<<(unsigned)tiles, EMIT_THREADS, ELEM_STAGE_BYTES, stream>>>(P);
    else
        k_emit_elements<false><<<(unsigned)tiles, EMIT_THREADS, ELEM_STAGE_BYTES, stream>>>(P);
    CFE_CHECK_LAUNCH("k_emit_elements");
    return 0;
}

// ------------------------------------------------------------------------------------------
// *NSET / *ELSET sections: a stream of items, each either a literal (section comment, set
// header, the closing '\n' of a set) or one id of a set's list followed by ',' or '\n'.

#define ITEM_STAGE_BYTES (EMIT_THREADS * 24 + 8192 + 32)

__global__ void __launch_bounds__(EMIT_THREADS)
k_emit_items(const CfeItemSeg* __restrict__ segs, int n_segs, i64 n_items, const i64* __restrict__ ids,
             const char* __restrict__ lit, char* out_base, const u64* dyn_off, u64* total, u64* seg_off, u64* ws)
{
    extern __shared__ __align__(16) char stage[];
    __shared__ u32 s_scan[33];
    __shared__ u64 s_base;
    const u32 tile = scan_next_tile(ws);
    const i64 t = (i64)tile * EMIT_THREADS + threadIdx.x;
    u32 len = 0;
    int nd = 0;
    u64 id = 0;
    char sep = 0;
    const char* src = nullptr;
    int seg_first = -1;
    if (t < n_items) {
        int q = 0;
        while (q + 1 < n_segs && t >= segs[q + 1].item0) ++q;
        const CfeItemSeg sg = segs[q];
        if (t == sg.item0) seg_first = q;
        if (sg.kind == 0) {
            src = lit + sg.lit_off;
            len = (u32)sg.lit_len;
        } else {
            const i64 e = t - sg.item0;
            id = (u64)__ldg(ids + sg.id0 + e);
            nd = digits_u64(id);
            // voxelFE.py:657-666: ',' after entries 1-9 of a line, '\n' after the tenth
            sep = (((sg.rank0 + e + 1) % 10) == 0) ? '\n' : ',';
            len = (u32)nd + 1u;
        }
    }
    u32 tile_total;
    const u32 excl = block_exclusive_scan<u32>(len, s_scan, tile_total);
    if (threadIdx.x < 32) {
        const u64 e = scan_lookback(ws, tile, (u64)tile_total);
        if (threadIdx.x == 0) {
            s_base = e;
            if ((i64)(tile + 1) * EMIT_THREADS >= n_items) *total = e + tile_total;
        }
    }
    __syncthreads();
    char* dst = out_base + (dyn_off ? *dyn_off : 0ull) + s_base;
    const u32 pad = (u32)(reinterpret_cast<uintptr_t>(dst) & 15u);
    if (seg_first >= 0 && seg_off) seg_off[seg_first] = s_base + excl;
    const bool staged = (pad + tile_total) <= (u32)ITEM_STAGE_BYTES;
    if (t < n_items) {
        char* p = staged ? (stage + pad + excl) : (dst + excl);
        if (src) {
            for (u32 q = 0; q < len; ++q) p[q] = __ldg(src + q);
        } else {
            put_u64_back(p + nd, id, nd);
            p[nd] = sep;
        }
    }
    __syncthreads();
    if (staged) flush_stage(stage, pad, tile_total, dst);
}

extern "C" int cfe_emit_items(const CfeItemSeg* segs_dev, int n_segs, int64_t n_items, const int64_t* ids,
                              const char* literals, char* out_base, const uint64_t* dyn_off, uint64_t* total,
                              uint64_t* seg_off, void* ws, cudaStream_t stream)
{
    const i64 tiles = (n_items + EMIT_THREADS - 1) / EMIT_THREADS;
    CFE_CHECK_CUDA(cudaMemsetAsync(ws, 0, (size_t)(tiles + 2) * sizeof(u64), stream));
    CFE_CHECK_CUDA(cudaMemsetAsync(total, 0, sizeof(u64), stream));
    if (tiles == 0) return 0;
    static bool attr_set = false;
    if (!attr_set) {
        CFE_CHECK_CUDA(cudaFuncSetAttribute(k_emit_items, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            ITEM_STAGE_BYTES));
        attr_set = true;
    }
    k_emit_items<<<(unsigned)tiles, EMIT_THREADS, ITEM_STAGE_BYTES, stream>>>(
        segs_dev, n_segs, n_items, reinterpret_cast<const i64*>(ids), literals, out_base,
        reinterpret_cast<const u64*>(dyn_off), reinterpret_cast<u64*>(total), reinterpret_cast<u64*>(seg_off),
        reinterpret_cast<u64*>(ws));
    CFE_CHECK_LAUNCH("k_emit_items");
    return 0;
}

// ------------------------------------------------------------------------------------------
// Literal text (header, section comments, PROPERTY cards + template) copied to a destination
// whose offset may depend on device-side totals: dst = out + const_off + *dyn0 + *dyn1.

__global__ void __launch_bounds__(256)
k_put_literal(const char* __restrict__ src, i64 n, char* out, const u64* dyn0, const u64* dyn1)
{
    char* dst = out + (dyn0 ? *dyn0 : 0ull) + (dyn1 ? *dyn1 : 0ull);
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x)
        dst[i] = src[i];
}

extern "C" int cfe_put_literal(const char* src, int64_t n, char* out, const uint64_t* dyn0,
                               const uint64_t* dyn1, cudaStream_t stream)
{
    if (n <= 0) return 0;
    i64 blocks = (n + 255) / 256;
    if (blocks > 1024) blocks = 1024;
    k_put_literal<<<(unsigned)blocks, 256, 0, stream>>>(src, n, out, reinterpret_cast<const u64*>(dyn0),
                                                        reinterpret_cast<const u64*>(dyn1));
    CFE_CHECK_LAUNCH("k_put_literal");
    return 0;
}

// ------------------------------------------------------------------------------------------
// Grey-value histogram of the elements (np.unique(cell_GV), voxelFE.py:570).  Bank-conflict-free
// privatisation: 32 copies of the 256 bins, copy l used by lane l (bin b of lane l lives in
// bank l), so a warp never conflicts with itself even on a binary volume.

__global__ void __launch_bounds__(256)
k_gv_hist(const uint8_t* __restrict__ vol, i64 n_vox, int thr, u64* __restrict__ hist)
{
    __shared__ u32 h[256 * 32];
    for (int i = threadIdx.x; i < 256 * 32; i += blockDim.x) h[i] = 0;
    __syncthreads();
    const u32 lane = threadIdx.x & 31u;
    const i64 n16 = n_vox >> 4;
    const bool aligned = (reinterpret_cast<uintptr_t>(vol) & 15u) == 0;
    if (aligned) {
        for (i64 g = (i64)blockIdx.x * blockDim.x + threadIdx.x; g < n16; g += (i64)gridDim.x * blockDim.x) {
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(vol) + g);
            const u32 w[4] = {v.x, v.y, v.z, v.w};
            int prev = -1;
            u32 run = 0;
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const int b = (int)((w[q >> 2] >> (8 * (q & 3))) & 0xffu);
                if (b == prev) { ++run; continue; }
                if (run && prev > thr) atomicAdd(&h[prev * 32 + lane], run);
                prev = b;
                run = 1;
            }
            if (run && prev > thr) atomicAdd(&h[prev * 32 + lane], run);
        }
    }
    // tail (or everything, when unaligned)
    const i64 start = aligned ? (n16 << 4) : 0;
    for (i64 i = start + (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n_vox; i += (i64)gridDim.x * blockDim.x) {
        const int b = (int)vol[i];
        if (b > thr) atomicAdd(&h[b * 32 + lane], 1u);
    }
    __syncthreads();
    // 256 threads: thread b reduces bin b (rotated start -> conflict-free reads)
    const int b = threadIdx.x;
    u32 sum = 0;
#pragma unroll 8
    for (int l = 0; l < 32; ++l) sum += h[b * 32 + ((l + b) & 31)];
    if (sum) atomicAdd(hist + b, (u64)sum);
}

extern "C" int cfe_gv_histogram(const uint8_t* vol, int64_t n_vox, int thr, uint64_t* hist256,
                                cudaStream_t stream)
{
    CFE_CHECK_CUDA(cudaMemsetAsync(hist256, 0, 256 * sizeof(u64), stream));
    if (n_vox <= 0) return 0;
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    i64 blocks = (n_vox / 16 + 255) / 256;
    if (blocks > sms * 4) blocks = sms * 4;
    if (blocks < 1) blocks = 1;
    k_gv_hist<<<(unsigned)blocks, 256, 0, stream>>>(vol, n_vox, thr, reinterpret_cast<u64*>(hist256));
    CFE_CHECK_LAUNCH("k_gv_hist");
    return 0;
}

// ------------------------------------------------------------------------------------------
// Stable partition of the raster-ordered element list by grey value (one 8-bit radix pass):
//   pass 1: per-tile 256-bin histogram, stored bin-major  mat[gv][tile]
//   flat exclusive scan of mat (cfe_scan_u32_inplace)  -> destination of (gv, tile)
//   pass 2: per-tile stable ranking (warp match.any) and scatter of {voxel, raster rank}.

#define PART_THREADS 256
#define PART_WARPS 8
#define PART_STEPS 16
#define PART_TILE (PART_THREADS * PART_STEPS)   // 4096 elements per CTA

extern "C" int64_t cfe_partition_tiles(int64_t n_el) { return (n_el + PART_TILE - 1) / PART_TILE; }

__global__ void __launch_bounds__(PART_THREADS)
k_part_hist(const uint8_t* __restrict__ vol, const u32* __restrict__ elem_vox, i64 n_el, i64 n_tiles,
            u32* __restrict__ mat)
{
    __shared__ u32 h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const i64 tile = blockIdx.x;
    const u32 warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const i64 w0 = tile * PART_TILE + (i64)warp * (PART_STEPS * 32);
#pragma unroll 4
    for (int t = 0; t < PART_STEPS; ++t) {
        const i64 j = w0 + t * 32 + lane;
        if (j < n_el) atomicAdd(&h[__ldg(vol + __ldg(elem_vox + j))], 1u);
    }
    __syncthreads();
    mat[(i64)threadIdx.x * n_tiles + tile] = h[threadIdx.x];
}

__global__ void __launch_bounds__(PART_THREADS)
k_part_scatter(const uint8_t* __restrict__ vol, const u32* __restrict__ elem_vox, i64 n_el, i64 n_tiles,
               const u64* __restrict__ mat_off, uint2* __restrict__ perm)
{
    __shared__ u32 wh[PART_WARPS][256];
    for (int i = threadIdx.x; i < PART_WARPS * 256; i += PART_THREADS) (&wh[0][0])[i] = 0;
    __syncthreads();
    const i64 tile = blockIdx.x;
    const u32 warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const i64 w0 = tile * PART_TILE + (i64)warp * (PART_STEPS * 32);
    u32 vox[PART_STEPS];
    int gv[PART_STEPS];
#pragma unroll
    for (int t = 0; t < PART_STEPS; ++t) {
        const i64 j = w0 + t * 32 + lane;
        gv[t] = -1;
        vox[t] = 0;
        if (j < n_el) {
            vox[t] = __ldg(elem_vox + j);
            gv[t] = (int)__ldg(vol + vox[t]);
            atomicAdd(&wh[warp][gv[t]], 1u);
        }
    }
    __syncthreads();
    {
        // thread g turns the per-warp counts of bin g into per-warp destination cursors
        const int g = threadIdx.x;
        u64 run = mat_off[(i64)g * n_tiles + tile];
#pragma unroll
        for (int w = 0; w < PART_WARPS; ++w) {
            const u32 c = wh[w][g];
            wh[w][g] = (u32)run;  // destinations are < n_el < 2^32
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int t = 0; t < PART_STEPS; ++t) {
        const i64 j = w0 + t * 32 + lane;
        const u32 peers = __match_any_sync(CFE_FULL_MASK, gv[t]);
        u32 base = 0;
        if (gv[t] >= 0) base = wh[warp][gv[t]];
        __syncwarp();
        if (gv[t] >= 0) {
            const u32 rank = (u32)__popc(peers & ((1u << lane) - 1u));
            if (rank == 0) wh[warp][gv[t]] = base + (u32)__popc(peers);
            perm[base + rank] = make_uint2(vox[t], (u32)j);
        }
        __syncwarp();
    }
}

extern "C" int cfe_partition_hist(const uint8_t* vol, const uint32_t* elem_vox, int64_t n_el, uint32_t* mat,
                                  cudaStream_t stream)
{
    const i64 n_tiles = cfe_partition_tiles(n_el);
    if (n_tiles == 0) return 0;
    k_part_hist<<<(unsigned)n_tiles, PART_THREADS, 0, stream>>>(vol, elem_vox, n_el, n_tiles, mat);
    CFE_CHECK_LAUNCH("k_part_hist");
    return 0;
}

extern "C" int cfe_partition_scatter(const uint8_t* vol, const uint32_t* elem_vox, int64_t n_el,
                                     const uint64_t* mat_off, void* perm, cudaStream_t stream)
{
    const i64 n_tiles = cfe_partition_tiles(n_el);
    if (n_tiles == 0) return 0;
    k_part_scatter<<<(unsigned)n_tiles, PART_THREADS, 0, stream>>>(vol, elem_vox, n_el, n_tiles,
                                                                   reinterpret_cast<const u64*>(mat_off),
                                                                   reinterpret_cast<uint2*>(perm));
    CFE_CHECK_LAUNCH("k_part_scatter");
    return 0;
}
