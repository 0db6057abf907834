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
// 1: staged text leaves shared memory through the bulk-copy engine (cp.async.bulk.global.shared::cta);
// 0: LDS.128 + STG.128 loops (kept for A/B measurements, profiles/tools/build_variant.sh -DEE2_BULK_FLUSH=0)
#ifndef EE2_BULK_FLUSH
#define EE2_BULK_FLUSH 1
#endif

extern "C" size_t cfe_scan_workspace_bytes(int64_t n_items);

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
// SWAR decimal formatting.  raw4(g): the four decimal digits of g < 10000, one per byte, most
// significant digit in the LOW byte (= first in memory); add 0x30303030 for ASCII.
__device__ __forceinline__ u32 raw4(u32 g)
{
    const u32 hi = (g * 5243u) >> 19;                    // g / 100   (exact for g < 10000)
    const u32 lo = g - hi * 100u;
    const u32 x = hi | (lo << 16);
    const u32 t = ((x * 103u) >> 10) & 0x000F000Fu;      // tens digit of each pair
    const u32 o = x - t * 10u;                           // ones digit of each pair
    return t | (o << 8);
}

// 12 zero-padded decimal digits of v (< 10^12) as three raw4 words, first word = most significant.
template <bool WIDE>
__device__ __forceinline__ void raw12(u64 v, u32& r0, u32& r1, u32& r2, u32& low4)
{
    u32 a, rem;
    if (WIDE) {
        const u64 q = v / 100000000ull;
        a = (u32)q;
        rem = (u32)(v - q * 100000000ull);
    } else {
        const u32 vv = (u32)v;
        a = vv / 100000000u;
        rem = vv - a * 100000000u;
    }
    const u32 b = rem / 10000u;
    low4 = rem - b * 10000u;
    if (WIDE) {
        r0 = raw4(a);
    } else {
        // a 32-bit value has at most two digits above 10^8: tens and ones go to bytes 2 and 3
        const u32 t = (a * 103u) >> 10;
        r0 = (t << 16) | ((a - t * 10u) << 24);
    }
    r1 = raw4(b);
    r2 = raw4(low4);
}

// number of significant digits (>= 1) of the 12-digit raw representation
__device__ __forceinline__ int raw12_digits(u32 r0, u32 r1, u32 r2)
{
    int lz;
    if (r0) lz = (__ffs((int)r0) - 1) >> 3;
    else if (r1) lz = 4 + ((__ffs((int)r1) - 1) >> 3);
    else if (r2) lz = 8 + ((__ffs((int)r2) - 1) >> 3);
    else lz = 11;
    return 12 - lz;
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


// Fast path: one thread formats FOUR consecutive node lines (212 bytes = 53 words), so every byte
// position inside the thread's block is a compile-time constant and the staging buffer is filled
// with aligned 32-bit stores (stride 53 words between threads: bank-conflict free).  Requires a
// 4-byte aligned destination (the callers align the *NODE section to 16 bytes).
#define NODE4_THREADS 128
#define NODE4_PER_CTA (NODE4_THREADS * 4)

__device__ __forceinline__ void node_line_words(const NodeEmitParams& P, u32 v, u32 (&L)[14])
{
    const u32 s = fd_div(v, P.dSN);
    const u32 rem = v - s * P.SN32;
    const u32 r = fd_div(rem, P.dRN);
    const u32 c = rem - r * P.RN32;
    const u64 id = (u64)P.SN32 * (u64)((i64)s + P.z0) + (u64)rem;
    u32 r0, r1, r2, low4;
    raw12<true>(id, r0, r1, r2, low4);
    const int nd = raw12_digits(r0, r1, r2);
    // '{:10d}': blank the leading zeros (chars [0, 12-nd)) -> ' ' = '0' - 0x10
    const int blank = 12 - nd;
    u32 I[3] = {r0 + 0x30303030u, r1 + 0x30303030u, r2 + 0x30303030u};
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const int bc = blank - 4 * i;
        const u32 m = bc >= 4 ? 0x10101010u : (bc <= 0 ? 0u : (0x10101010u & ((1u << (8 * bc)) - 1u)));
        I[i] -= m;
    }
    const uint4 fx = __ldg(reinterpret_cast<const uint4*>(P.xtab) + c);
    const uint4 fy = __ldg(reinterpret_cast<const uint4*>(P.ytab) + r);
    const uint4 fz = __ldg(reinterpret_cast<const uint4*>(P.ztab) + s);
    const u32 CS = 0x0000202Cu;   // ", "
    L[0] = __byte_perm(I[0], I[1], 0x5432);   // id chars 2..5
    L[1] = __byte_perm(I[1], I[2], 0x5432);   // id chars 6..9
    L[2] = __byte_perm(I[2], CS, 0x5432);     // id chars 10,11 + ", "
    L[3] = fx.x; L[4] = fx.y; L[5] = fx.z;
    L[6] = __byte_perm(CS, fy.x, 0x5410);     // ", " + y chars 0,1
    L[7] = __byte_perm(fy.x, fy.y, 0x5432);
    L[8] = __byte_perm(fy.y, fy.z, 0x5432);
    L[9] = __byte_perm(fy.z, CS, 0x5432);     // y chars 10,11 + ", "
    L[10] = fz.x; L[11] = fz.y; L[12] = fz.z;
    L[13] = 0x0000000Au;                      // '\n'
}

__global__ void __launch_bounds__(NODE4_THREADS) k_emit_nodes4(NodeEmitParams P)
{
    __shared__ __align__(16) u32 stage[NODE4_THREADS * 53];
    const i64 k0 = (i64)blockIdx.x * NODE4_PER_CTA;           // first node of this CTA
    const i64 k = k0 + (i64)threadIdx.x * 4;
    u32* B = stage + threadIdx.x * 53;
    if (k + 3 < P.n_nodes) {
        const uint4 v4 = __ldg(reinterpret_cast<const uint4*>(P.node_list + k));
        const u32 vv[4] = {v4.x, v4.y, v4.z, v4.w};
        u32 pending = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            u32 L[14];
            node_line_words(P, vv[q], L);
            const int sh = 8 * q;                              // (53 q) mod 4 == q
            B[13 * q] = pending | (L[0] << sh);
#pragma unroll
            for (int i = 1; i < 13; ++i) B[13 * q + i] = sh ? __funnelshift_l(L[i - 1], L[i], sh) : L[i];
            pending = sh ? __funnelshift_l(L[12], L[13], sh) : L[13];
        }
        B[52] = pending;
    }
#if EE2_BULK_FLUSH
    fence_proxy_async_smem();
#endif
    __syncthreads();
    i64 n_here = P.n_nodes - k0;
    if (n_here > NODE4_PER_CTA) n_here = NODE4_PER_CTA;
    const u32 nbytes = (u32)n_here * 53u;                      // n_here is a multiple of 4
    char* dst = P.out + k0 * 53;
    if ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0) {
        const u32 n16 = nbytes >> 4;
#if EE2_BULK_FLUSH
        // the CTA's 27 KB leave through the bulk-copy engine (cp.async.bulk shared -> global)
        if (threadIdx.x == 0 && n16) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                         :: "l"(dst), "r"((u32)__cvta_generic_to_shared(stage)), "r"(16u * n16) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
#else
        for (u32 c = threadIdx.x; c < n16; c += blockDim.x)
            reinterpret_cast<uint4*>(dst)[c] = reinterpret_cast<const uint4*>(stage)[c];
#endif
        for (u32 c = (n16 << 2) + threadIdx.x; c < (nbytes >> 2); c += blockDim.x)
            reinterpret_cast<u32*>(dst)[c] = stage[c];
#if EE2_BULK_FLUSH
        if (threadIdx.x == 0 && n16) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // before the smem goes away
#endif
    } else {
        for (u32 c = threadIdx.x; c < (nbytes >> 2); c += blockDim.x) reinterpret_cast<u32*>(dst)[c] = stage[c];
    }
}

// List-free variant: the CTA's 512 nodes are expanded from the used-node bitmap itself.  chunk_word[b] (written
// by the node scan) is the word of the node grid that holds node 512 b; the CTA walks the words from there, 128
// at a time, and drops the local grid index of every used node with rank in [512 b, 512 b + 512) into a
// shared-memory list -- usually ~45 words, more only across empty stretches of the grid.  No 4 B / node list is
// written by a separate pass and read back here; everything after the expansion is k_emit_nodes4.
#ifndef NODE4B_CTAS_PER_SM
#define NODE4B_CTAS_PER_SM 8      // 27 KB of shared memory each
#endif
struct NodeBitsParams {
    const u32* nbits;        // used-node bitmap, [node rows][wpr] words
    const u32* npref;        // exclusive used-node count per word
    const u32* chunk_word;   // [ceil(n_nodes / 512)]
    u32 n_words, wpr, rowlen;
    FastDiv d_wpr;
};

// PERSISTENT CTAs, software pipelined: while chunk b is expanded and formatted, the bitmap words of the CTA's
// next chunk are already in flight (their position, chunk_word, was fetched one iteration earlier still), so
// the two dependent global loads of a chunk never sit on the critical path.
__global__ void __launch_bounds__(NODE4_THREADS) k_emit_nodes4b(NodeEmitParams P, NodeBitsParams B, u32 n_chunks)
{
    __shared__ __align__(16) u32 stage[NODE4_THREADS * 53];
    u32* const list = stage;          // the 512-entry node list lives in the first 2 KB of the staging area:
                                      // every thread takes its four entries into registers before any line is staged
    const u32 G = gridDim.x;
    u32 b = blockIdx.x;
    if (b >= n_chunks) return;
    u32 cw = __ldg(B.chunk_word + b);
    u32 cw_next = (b + G < n_chunks) ? __ldg(B.chunk_word + b + G) : 0u;
    u32 word = 0, pref = 0;
    if (cw + threadIdx.x < B.n_words) { word = __ldg(B.nbits + cw + threadIdx.x); pref = __ldg(B.npref + cw + threadIdx.x); }
    bool pending_copy = false;
    for (; b < n_chunks; b += G) {
        // ---- prefetch: first 128 words of the next chunk, position of the one after
        u32 nx_word = 0, nx_pref = 0, cw_next2 = 0;
        if (b + G < n_chunks) {
            if (cw_next + threadIdx.x < B.n_words) {
                nx_word = __ldg(B.nbits + cw_next + threadIdx.x);
                nx_pref = __ldg(B.npref + cw_next + threadIdx.x);
            }
            if (b + 2 * G < n_chunks) cw_next2 = __ldg(B.chunk_word + b + 2 * G);
        }
        const i64 k0 = (i64)b * NODE4_PER_CTA;                 // first node of this chunk
        i64 n_here = P.n_nodes - k0;
        if (n_here > NODE4_PER_CTA) n_here = NODE4_PER_CTA;
        const u32 lo = (u32)k0, hi = (u32)(k0 + n_here);        // ranks [lo, hi)
        // the previous chunk's bulk copy must have read the staging area before it is overwritten
        // (the node list below lives inside the staging area, so every thread has to wait for that)
        if (pending_copy) {
            if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            __syncthreads();
        }
        // ---- expand: local grid index of the used nodes with rank in [lo, hi) -> list
        for (u32 wbase = cw;; wbase += NODE4_THREADS) {
            const u32 i = wbase + threadIdx.x;
            u32 w = word, p = pref;
            if (wbase != cw) {                                   // rare: a chunk that spans more than 128 words
                w = 0; p = 0;
                if (i < B.n_words) { w = __ldg(B.nbits + i); p = __ldg(B.npref + i); }
            }
            const u32 pe = p + (u32)__popc(w);
            if (w && pe > lo && p < hi) {
                const u32 row = fd_div(i, B.d_wpr);
                const u32 base = row * B.rowlen + (i - row * B.wpr) * 32u;
                while (w) {
                    const int bit = __ffs((int)w) - 1;
                    w &= w - 1;
                    if (p >= lo && p < hi) list[p - lo] = base + (u32)bit;
                    ++p;
                }
            }
            const bool last = (i < B.n_words && pe >= hi) || wbase + NODE4_THREADS >= B.n_words;
            if (__syncthreads_or(last ? 1 : 0)) break;
        }
        // ---- format four lines per thread (k_emit_nodes4)
        const u32 n4 = (u32)n_here & ~3u;
        const u32 q0 = threadIdx.x * 4u;
        u32* Bw = stage + threadIdx.x * 53;
        uint4 v4 = make_uint4(0u, 0u, 0u, 0u);
        u32 v_tail = 0;
        if (q0 + 3u < n4) v4 = *reinterpret_cast<const uint4*>(list + q0);
        if (threadIdx.x < (u32)n_here - n4) v_tail = list[n4 + threadIdx.x];
        __syncthreads();                                        // the list has been read: the stage may be written
        if (q0 + 3u < n4) {
            const u32 vv[4] = {v4.x, v4.y, v4.z, v4.w};
            u32 pending = 0;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                u32 L[14];
                node_line_words(P, vv[q], L);
                const int sh = 8 * q;                          // (53 q) mod 4 == q
                Bw[13 * q] = pending | (L[0] << sh);
#pragma unroll
                for (int i = 1; i < 13; ++i) Bw[13 * q + i] = sh ? __funnelshift_l(L[i - 1], L[i], sh) : L[i];
                pending = sh ? __funnelshift_l(L[12], L[13], sh) : L[13];
            }
            Bw[52] = pending;
        }
        char* dst = P.out + k0 * 53;                            // 16-byte aligned: 53 * 512 = 16 * 1696
        // the last <= 3 nodes of the section: byte-wise lines
        if (threadIdx.x < (u32)n_here - n4) {
            u32 L[14];
            node_line_words(P, v_tail, L);
            char* line = dst + (size_t)(n4 + threadIdx.x) * 53;
#pragma unroll
            for (int q = 0; q < 53; ++q) line[q] = (char)((L[q >> 2] >> (8 * (q & 3))) & 0xffu);
        }
        fence_proxy_async_smem();
        __syncthreads();
        const u32 nbytes = n4 * 53u;
        const u32 n16 = nbytes >> 4;
        if (threadIdx.x == 0 && n16) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                         :: "l"(dst), "r"((u32)__cvta_generic_to_shared(stage)), "r"(16u * n16) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        pending_copy = true;
        for (u32 c = (n16 << 2) + threadIdx.x; c < (nbytes >> 2); c += blockDim.x) reinterpret_cast<u32*>(dst)[c] = stage[c];
        cw = cw_next; cw_next = cw_next2; word = nx_word; pref = nx_pref;
    }
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // before the smem goes away
}

// *NODE lines straight from the used-node bitmap (nbits / npref / chunk_word of cfe_scan_nodes_index; n_words =
// words of the bitmap, (X + 32) / 32 per node row).  `out` must be 16-byte aligned.
extern "C" int cfe_emit_nodes_bits(const uint32_t* nbits, const uint32_t* npref, const uint32_t* chunk_word,
                                   int64_t n_words, int64_t n_nodes, const char* xtab, const char* ytab,
                                   const char* ztab, int64_t Y, int64_t X, int64_t z0, char* out,
                                   cudaStream_t stream)
{
    if (n_nodes <= 0) return 0;
    if ((reinterpret_cast<uintptr_t>(out) & 15u) || n_words <= 0 || n_words >= (1LL << 32)) {
        cfe_set_error("cfe_emit_nodes_bits: the *NODE section must start 16-byte aligned");
        return -1;
    }
    NodeEmitParams P;
    P.node_list = nullptr; P.n_nodes = n_nodes; P.xtab = xtab; P.ytab = ytab; P.ztab = ztab;
    P.SN32 = (u32)((X + 1) * (Y + 1)); P.RN32 = (u32)(X + 1);
    P.dSN = make_fastdiv(P.SN32); P.dRN = make_fastdiv(P.RN32);
    P.z0 = z0; P.out = out;
    NodeBitsParams B;
    B.nbits = nbits; B.npref = npref; B.chunk_word = chunk_word;
    B.n_words = (u32)n_words; B.wpr = (u32)((X + 32) / 32); B.rowlen = (u32)(X + 1);
    B.d_wpr = make_fastdiv(B.wpr);
    const i64 n_chunks = (n_nodes + NODE4_PER_CTA - 1) / NODE4_PER_CTA;
    static unsigned long long once = 0;
    int sms = 148;
    cfe_first_on_device(&once, &sms);
    i64 blocks = (i64)sms * NODE4B_CTAS_PER_SM;                  // persistent: every CTA strides over the chunks
    if (blocks > n_chunks) blocks = n_chunks;
    k_emit_nodes4b<<<(unsigned)blocks, NODE4_THREADS, 0, stream>>>(P, B, (u32)n_chunks);
    CFE_CHECK_LAUNCH("k_emit_nodes4b");
    return 0;
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
    i64 done = 0;
    const bool aligned = ((reinterpret_cast<uintptr_t>(out) & 3u) == 0) &&
                         ((reinterpret_cast<uintptr_t>(node_list) & 15u) == 0);
    if (aligned && n_nodes >= 4) {
        P.n_nodes = n_nodes & ~3ll;
        const i64 blocks = (P.n_nodes + NODE4_PER_CTA - 1) / NODE4_PER_CTA;
        k_emit_nodes4<<<(unsigned)blocks, NODE4_THREADS, 0, stream>>>(P);
        CFE_CHECK_LAUNCH("k_emit_nodes4");
        done = P.n_nodes;
    }
    if (done < n_nodes) {   // generic byte-wise kernel: unaligned destination or the last n % 4 nodes
        P.node_list = node_list + done;
        P.n_nodes = n_nodes - done;
        P.out = out + done * 53;
        const i64 blocks = (P.n_nodes + EMIT_THREADS - 1) / EMIT_THREADS;
        k_emit_nodes<<<(unsigned)blocks, EMIT_THREADS, 0, stream>>>(P);
        CFE_CHECK_LAUNCH("k_emit_nodes");
    }
    return 0;
}

// ------------------------------------------------------------------------------------------
// *ELEMENT section: variable-length lines, one thread per element, single-pass look-back.

#define ELEM_HDR_SLOT 64      // bytes reserved per '*ELEMENT, TYPE=..., ELSET=SET<gv>\n' header
#define ELEM_MAX_LINE 192     // eid (<= 20) + 8 node ids (<= 20 each) + 9 separators
#define ELEM_STAGE_BYTES (EMIT_THREADS * 100 + 2048 + 32)

struct ElemEmitParams {
    const u32* elem_vox;     // raster-ordered local voxel indices (PERM = false)
    const uint2* perm;       // GV-major order: {local voxel index, raster rank} (PERM = true)
    i64 n_el;
    const uint8_t* vol;      // grey values (only read when PERM)
    const u64* first;        // [256] emission index of the first element with grey value >= g (n_el when none)
    const char* hdr_text;    // [256][ELEM_HDR_SLOT]
    const uint8_t* hdr_len;  // [256]; 0 = no header for this GV (e.g. slabs other than the first)
    int single_gv;           // PERM = false: the one grey value
    FastDiv dYX, dX;
    u32 YX32, X32;
    i64 SN, RN, z0, eid_offset;
    char* out;               // start of the element lines
    u64* blk_off;            // out [256]: byte offset (from `out`) where each GV block (header included) starts
    uint8_t* len8;           // [n_el] line length (block header included) in emission order
    uint8_t* nd8;            // [n_el] digit-count flags (see k_elem_lengths)
    u32* warp_total;         // [n_warp_tiles] bytes of every 32-element tile
    u64* warp_off;           // [n_warp_tiles] exclusive scan of warp_total
    i64 n_warp_tiles;
    u32* tile_counter;       // next chunk of EE2_CHUNK tiles to hand out (k_emit_elements2, dynamic schedule)
};

// Line builder: bytes are appended to a 32-bit accumulator and leave as aligned 32-bit shared-memory
// stores.  A line's first word is stored whole with its low `d` bytes zero (they belong to the previous
// line, which sits in the neighbouring lane); the last partial word of every line is merged into the
// staging area by lw_finish after the warp has synchronised -- no byte stores anywhere.
struct LineWriter {
    char* wp;      // 4-byte aligned staging address of the word being assembled
    u32 acc;       // its low `d` bytes are valid
    u32 d;         // 0..3
};

// append the nd significant digits of a raw12 number followed by `sep`
__device__ __forceinline__ void lw_append(LineWriter& w, u32 r0, u32 r1, u32 r2, int nd, u32 sep)
{
    const u32 W0 = r0 + 0x30303030u, W1 = r1 + 0x30303030u, W2 = r2 + 0x30303030u;
    const int skip = 12 - nd;                 // leading characters to drop (2..11)
    const u32 sb = (u32)(skip & 3) * 8u;
    // x0..x3 = the 16-byte string W0 W1 W2 sep shifted left by (skip / 4) words
    const bool s1 = skip >= 4, s2 = skip >= 8;
    const u32 x0 = s2 ? W2 : (s1 ? W1 : W0);
    const u32 x1 = s2 ? sep : (s1 ? W2 : W1);
    const u32 x2 = s2 ? 0u : (s1 ? sep : W2);
    const u32 x3 = s1 ? 0u : sep;
    // left-aligned field: nd digits, separator, zeros
    const u32 A0 = __funnelshift_r(x0, x1, sb), A1 = __funnelshift_r(x1, x2, sb), A2 = __funnelshift_r(x2, x3, sb);
    const u32 s8 = w.d * 8u;
    const u32 O0 = w.acc | (A0 << s8);
    const u32 O1 = __funnelshift_l(A0, A1, s8);
    const u32 O2 = __funnelshift_l(A1, A2, s8);
    const u32 O3 = __funnelshift_l(A2, 0u, s8);
    const u32 tot = w.d + (u32)nd + 1u;       // bytes now held by O0..O3 (<= 14)
    u32* p = reinterpret_cast<u32*>(w.wp);
    if (tot >= 4u) p[0] = O0;
    if (tot >= 8u) p[1] = O1;
    if (tot >= 12u) p[2] = O2;
    w.acc = tot >= 8u ? (tot >= 12u ? O3 : O2) : (tot >= 4u ? O1 : O0);
    w.wp += tot & ~3u;
    w.d = tot & 3u;
}

// Four consecutive node-id fields of an element line when all eight ids have the same number of
// digits N (the usual case: the ids of one element differ by < 2 slices of the node grid).  Every
// byte position of the 4*(N+1)-byte half-block is then a compile-time constant: it is assembled
// with immediate funnel shifts from the 12-digit words of the fields (F[k] = ASCII words of field
// k, most significant first) and appended with ONE runtime shift.  LAST: the half-block ends the
// line ('\n' after its fourth field).
template <bool LAST>
__device__ __forceinline__ u32 field_word(const u32 (&F)[4][3], int k, int j)
{
    // word j of the 16-byte string "12 digits, separator, zeros" of field k
    return j < 3 ? F[k][j] : (j == 3 ? ((LAST && k == 3) ? 0x0Au : 0x2Cu) : 0u);
}

template <bool LAST>
__device__ __forceinline__ u32 field_chars(const u32 (&F)[4][3], int k, int c)
{
    // characters c .. c+3 of that string (c is a compile-time constant after unrolling)
    const u32 lo = field_word<LAST>(F, k, c >> 2), hi = field_word<LAST>(F, k, (c >> 2) + 1);
    return (c & 3) ? __funnelshift_r(lo, hi, 8 * (c & 3)) : lo;
}

template <int N, bool LAST>
__device__ __forceinline__ void lw_append_block(LineWriter& w, const u32 (&F)[4][3])
{
    constexpr int NB = N + 1;        // bytes per field (digits + separator); 4 * NB bytes = NB words
    const u32 s8 = w.d * 8u;
    u32 prev = 0;
#pragma unroll
    for (int m = 0; m < NB; ++m) {
        u32 Bm = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int fs = k * NB - 4 * m;          // where field k starts relative to this word
            if (fs <= -NB || fs >= 4) continue;     // no overlap
            if (fs <= 0) Bm |= field_chars<LAST>(F, k, 12 - N - fs);
            else Bm |= field_chars<LAST>(F, k, 12 - N) << (8 * fs);
        }
        const u32 outw = (m == 0) ? (w.acc | (Bm << s8)) : __funnelshift_l(prev, Bm, s8);
        reinterpret_cast<u32*>(w.wp)[m] = outw;
        prev = Bm;
    }
    w.acc = __funnelshift_l(prev, 0u, s8);
    w.wp += 4 * NB;
}

template <bool LAST>
__device__ __forceinline__ void lw_append_half(LineWriter& w, const u32 (&F)[4][3], int n, bool same)
{
    if (same) {
        switch (n) {
        case 1: lw_append_block<1, LAST>(w, F); break;
        case 2: lw_append_block<2, LAST>(w, F); break;
        case 3: lw_append_block<3, LAST>(w, F); break;
        case 4: lw_append_block<4, LAST>(w, F); break;
        case 5: lw_append_block<5, LAST>(w, F); break;
        case 6: lw_append_block<6, LAST>(w, F); break;
        case 7: lw_append_block<7, LAST>(w, F); break;
        case 8: lw_append_block<8, LAST>(w, F); break;
        case 9: lw_append_block<9, LAST>(w, F); break;
        default: lw_append_block<10, LAST>(w, F); break;   // ids are < 10^10 (checked by the host)
        }
    } else {
        // ids straddle a power of ten: per-field appends
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const u32 a0 = F[k][0] - 0x30303030u, a1 = F[k][1] - 0x30303030u, a2 = F[k][2] - 0x30303030u;
            lw_append(w, a0, a1, a2, raw12_digits(a0, a1, a2), (LAST && k == 3) ? 0x0Au : 0x2Cu);
        }
    }
}

// The last partial word of a line shares its upper bytes with the first word of the next line, which the
// neighbouring lane has stored by now (the caller synchronises the warp first): merge, keep the upper bytes.
// Behind the last line of a tile the upper bytes are leftovers that the flush never copies.
__device__ __forceinline__ void lw_finish(const LineWriter& w)
{
    if (w.d) {
        u32* p = reinterpret_cast<u32*>(w.wp);
        *p = (*p & (0xFFFFFFFFu << (8u * w.d))) | w.acc;
    }
}

// Copy n staged bytes (stage[0..n), stage 16-byte aligned) to an arbitrarily aligned global dst:
// aligned 16-byte global stores assembled from two aligned 16-byte shared loads and a byte funnel.
// The word / byte shifts are uniform for the whole tile, so the loop is specialised on them.
template <int WSEL>
__device__ __forceinline__ void flush_loop(const uint4* sv, char* gbase, u32 c_first, u32 c_last, u32 bs)
{
#pragma unroll 1
    for (u32 c = c_first + threadIdx.x; c < c_last; c += blockDim.x) {
        const uint4 a = sv[c - 1u];
        const uint4 b = sv[c];
        u32 x0, x1, x2, x3, x4;
        if (WSEL == 0) { x0 = a.x; x1 = a.y; x2 = a.z; x3 = a.w; x4 = b.x; }
        else if (WSEL == 1) { x0 = a.y; x1 = a.z; x2 = a.w; x3 = b.x; x4 = b.y; }
        else if (WSEL == 2) { x0 = a.z; x1 = a.w; x2 = b.x; x3 = b.y; x4 = b.z; }
        else { x0 = a.w; x1 = b.x; x2 = b.y; x3 = b.z; x4 = b.w; }
        uint4 o;
        o.x = __funnelshift_r(x0, x1, bs); o.y = __funnelshift_r(x1, x2, bs);
        o.z = __funnelshift_r(x2, x3, bs); o.w = __funnelshift_r(x3, x4, bs);
        *reinterpret_cast<uint4*>(gbase + 16u * c) = o;
    }
}

__device__ __forceinline__ void flush_unaligned(const char* stage, u32 n, char* dst)
{
    const u32 pad = (u32)(reinterpret_cast<uintptr_t>(dst) & 15u);
    char* gbase = dst - pad;                       // 16-byte aligned; chunk c = [gbase+16c, +16)
    const u32 end = pad + n;
    const u32 c_last = end >> 4;                   // one past the last full chunk
    const uint4* sv = reinterpret_cast<const uint4*>(stage);
    if (pad == 0) {
#pragma unroll 1
        for (u32 c = threadIdx.x; c < c_last; c += blockDim.x) reinterpret_cast<uint4*>(gbase)[c] = sv[c];
    } else {
        // chunk c (>= 1) takes staged bytes [16(c-1) + off, +16), off = 16 - pad
        const u32 off = 16u - pad;
        const u32 bs = (off & 3u) * 8u;
        switch (off >> 2) {
        case 0: flush_loop<0>(sv, gbase, 1u, c_last, bs); break;
        case 1: flush_loop<1>(sv, gbase, 1u, c_last, bs); break;
        case 2: flush_loop<2>(sv, gbase, 1u, c_last, bs); break;
        default: flush_loop<3>(sv, gbase, 1u, c_last, bs); break;
        }
        const u32 hb = off < n ? off : n;          // partial head chunk: bytes
        for (u32 k = threadIdx.x; k < hb; k += blockDim.x) dst[k] = stage[k];
    }
    // partial tail chunk: bytes [16 c_last - pad, n)
    if ((c_last << 4) >= pad && (c_last << 4) < end && (pad == 0 || c_last >= 1u)) {
        const u32 s0 = (c_last << 4) - pad;
        for (u32 k = s0 + threadIdx.x; k < n; k += blockDim.x) dst[k] = stage[k];
    }
}

// The *ELEMENT section is emitted in two kernels so that no CTA ever waits on another:
//   k_elem_lengths  line length of every element (digit counts only) -> len8[j], and the byte
//                   total of every 32-element warp tile -> warp_total[w]
//   (exclusive scan of warp_total: cfe_scan_u32 machinery)
//   k_emit_elements each WARP formats its 32 lines into a private shared-memory region that is
//                   congruent mod 16 with its destination span and streams it out with 16-byte
//                   stores; there is no block-level barrier anywhere.
template <bool WIDE>
__device__ __forceinline__ int id_digits(u64 v)
{
    if (WIDE && (v >> 32)) return digits_u64(v);
    return digits_u32((u32)v);
}

// Grey value of emission index j in GV-major order: first[g] is non-decreasing in g (absent grey values
// carry the start of the next present one, n_el after the last), so the block of j is the largest g
// with first[g] <= j -- eight L1-resident probes instead of a random read of the volume.
__device__ __forceinline__ int gv_of_index(const u64* __restrict__ first, u64 j)
{
    int lo = 0;
#pragma unroll
    for (int step = 128; step > 0; step >>= 1)
        if (__ldg(first + lo + step) <= j) lo += step;
    return lo;
}

template <bool PERM>
__device__ __forceinline__ void elem_ids(const ElemEmitParams& P, i64 j, u64& b, u64& eid, int& gv)
{
    u32 v, rank;
    if (PERM) { const uint2 e = __ldg(P.perm + j); v = e.x; rank = e.y; gv = gv_of_index(P.first, (u64)j); }
    else { v = __ldg(P.elem_vox + j); rank = (u32)j; gv = P.single_gv; }
    const u32 s = fd_div(v, P.dYX);
    const u32 rem = v - s * P.YX32;
    const u32 r = fd_div(rem, P.dX);
    const u32 c = rem - r * P.X32;
    b = (u64)(P.SN * ((i64)s + P.z0) + P.RN * (i64)r + (i64)c);
    eid = (u64)(P.eid_offset + (i64)rank + 1);
}

template <bool PERM, bool WIDE>
__global__ void __launch_bounds__(256) k_elem_lengths(ElemEmitParams P)
{
    const i64 j = (i64)blockIdx.x * 256 + threadIdx.x;
    u32 len = 0;
    if (j < P.n_el) {
        u64 b, eid;
        int gv;
        elem_ids<PERM>(P, j, b, eid, gv);
        const int nd0 = id_digits<WIDE>(eid);
        const int nlo = id_digits<WIDE>(b);
        const int nhi = id_digits<WIDE>(b + P.SN + P.RN + 1);       // the largest of the eight node ids
        if (nlo == nhi) {
            len = 9u + (u32)nd0 + 8u * (u32)nlo;
        } else {                                                  // the ids straddle a power of ten
            len = 9u + (u32)nd0 + (u32)nlo + (u32)nhi + (u32)id_digits<WIDE>(b + 1) +
                  (u32)id_digits<WIDE>(b + P.RN) + (u32)id_digits<WIDE>(b + P.RN + 1) +
                  (u32)id_digits<WIDE>(b + P.SN) + (u32)id_digits<WIDE>(b + P.SN + 1) +
                  (u32)id_digits<WIDE>(b + P.SN + P.RN);
        }
        // flags byte: low nibble = digit count of the smallest node id, bit 7 = all eight counts equal
        P.nd8[j] = (uint8_t)((nlo == nhi ? 0x80u : 0u) | (u32)nlo);
        P.len8[j] = (uint8_t)len;                                  // the line only
        if ((u64)j == __ldg(P.first + gv)) len += __ldg(P.hdr_len + gv);   // + the block header in the tile total
    }
    const u32 tot = __reduce_add_sync(CFE_FULL_MASK, len);
    if ((threadIdx.x & 31u) == 0 && (j >> 5) < P.n_warp_tiles) P.warp_total[j >> 5] = tot;
}

// 32-line tile totals from precomputed lengths (cfe_fill_elements): one thread sums 16 lengths
// (one 16-byte load), two threads make a tile; the first element of a grey-value block also
// carries its '*ELEMENT ...' header.
__global__ void __launch_bounds__(256) k_tile_totals(ElemEmitParams P)
{
    const i64 t = (i64)blockIdx.x * 256 + threadIdx.x;      // 16-element chunk
    const i64 j0 = t * 16;
    u32 sum = 0;
    if (j0 < P.n_el) {
        if (j0 + 16 <= P.n_el) {
            const uint4 v = *reinterpret_cast<const uint4*>(P.len8 + j0);
            const u32 w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) sum += __dp4a(w[q], 0x01010101u, 0u);
        } else {
            for (i64 j = j0; j < P.n_el; ++j) sum += P.len8[j];
        }
        if (!P.perm) {
            const u64 f = __ldg(P.first + P.single_gv);
            if (f >= (u64)j0 && f < (u64)(j0 + 16)) sum += __ldg(P.hdr_len + P.single_gv);
        }
    }
    const u32 other = __shfl_xor_sync(CFE_FULL_MASK, sum, 1);
    if (!(threadIdx.x & 1) && (t >> 1) < P.n_warp_tiles) P.warp_total[t >> 1] = sum + other;
}

// GV-major order: the '*ELEMENT ...' header of every present grey value joins the tile of its first element
__global__ void __launch_bounds__(256) k_tile_headers(ElemEmitParams P)
{
    const int g = threadIdx.x;
    const u64 f = __ldg(P.first + g);
    const u64 nxt = (g < 255) ? __ldg(P.first + g + 1) : (u64)P.n_el;
    const u32 h = __ldg(P.hdr_len + g);
    if (h && f < nxt && f < (u64)P.n_el) atomicAdd(P.warp_total + (f >> 5), h);
}

#define ELEM_WARP_STAGE 3584     // bytes of staging per warp: 32 lines of <= 99 bytes + headers + pad

// ------------------------------------------------------------------------------------------
// Version 2 of the element emitter: PERSISTENT CTAs that keep a 10000-entry table "four decimal
// digits -> four ASCII bytes" in shared memory (40 KB, filled once per CTA), so that a 12-digit id
// costs two divisions and three table reads instead of three SWAR conversions; the ids v, v+1, v+RN,
// v+RN+1 of one node layer share their eight leading digits unless the low four wrap.  Everything else
// (warp-autonomous staging, static-layout half blocks, 16-byte flush) is as in k_emit_elements.
// EE2_LATE_LOADS: the loads of the next tile are issued behind the bulk copy of the current one (their latency
// then runs in parallel with the copy engine reading the staging area) instead of at the top of the iteration
// (where the membar of fence.proxy.async waits for them).  EE2_DYNAMIC: tiles are handed out in chunks of
// EE2_CHUNK consecutive tiles from a global counter instead of a fixed stride, so that no SM idles at the end.
#ifndef EE2_LATE_LOADS
#define EE2_LATE_LOADS 0
#endif
#ifndef EE2_DYNAMIC
#define EE2_DYNAMIC 1
#endif
#ifndef EE2_CHUNK
#define EE2_CHUNK 8u
#endif
#define EE2_WARPS 16
#define EE2_THREADS (32 * EE2_WARPS)
#define EE2_LUT_BYTES 40000
#define EE2_SMEM (EE2_LUT_BYTES + EE2_WARPS * ELEM_WARP_STAGE)

// ASCII words of the 12 zero-padded digits of v (< 10^12), most significant word first
template <bool WIDE>
__device__ __forceinline__ void asc12(const u32* lut, u64 v, u32& a0, u32& a1, u32& a2, u32& low4)
{
    u32 a, rem;
    if (WIDE) {
        const u64 q = v / 100000000ull;
        a = (u32)q;
        rem = (u32)(v - q * 100000000ull);
    } else {
        const u32 vv = (u32)v;
        a = vv / 100000000u;
        rem = vv - a * 100000000u;
    }
    const u32 b = rem / 10000u;
    low4 = rem - b * 10000u;
    a0 = lut[a]; a1 = lut[b]; a2 = lut[low4];
}

// the four node-id fields of one node layer in line order: v, v+1, v+RN+1, v+RN
template <bool WIDE>
__device__ __forceinline__ void layer_fields(const u32* lut, u64 v, u32 RN, u32 (&F)[4][3])
{
    u32 low4;
    asc12<WIDE>(lut, v, F[0][0], F[0][1], F[0][2], low4);
    const u32 l2 = low4 + RN + 1u;
    if (l2 < 10000u) {            // no carry out of the low four digits: the leading eight are shared
#pragma unroll
        for (int k = 1; k < 4; ++k) { F[k][0] = F[0][0]; F[k][1] = F[0][1]; }
        F[1][2] = lut[low4 + 1u];
        F[2][2] = lut[l2];
        F[3][2] = lut[l2 - 1u];
    } else {
        u32 dummy;
        asc12<WIDE>(lut, v + 1u, F[1][0], F[1][1], F[1][2], dummy);
        asc12<WIDE>(lut, v + RN + 1u, F[2][0], F[2][1], F[2][2], dummy);
        asc12<WIDE>(lut, v + RN, F[3][0], F[3][1], F[3][2], dummy);
    }
}

template <int N>
__device__ __forceinline__ void lw_append_line8(LineWriter& w, const u32 (&F0)[4][3], const u32 (&F1)[4][3])
{
    lw_append_block<N, false>(w, F0);
    lw_append_block<N, true>(w, F1);
}

template <bool WIDE> struct ElemId { typedef u32 type; };
template <> struct ElemId<true> { typedef u64 type; };

template <bool PERM, bool WIDE>
__global__ void __launch_bounds__(EE2_THREADS, 2) k_emit_elements2(ElemEmitParams P)
{
    typedef typename ElemId<WIDE>::type id_t;      // ids of a slab below 2^32 are computed in 32 bits
    extern __shared__ __align__(16) char ee2_smem[];
    u32* lut = reinterpret_cast<u32*>(ee2_smem);
    for (u32 i = threadIdx.x; i < 10000u; i += EE2_THREADS) lut[i] = raw4(i) + 0x30303030u;
    __syncthreads();
    const u32 lane = threadIdx.x & 31u;
    char* stage = ee2_smem + EE2_LUT_BYTES + (threadIdx.x >> 5) * ELEM_WARP_STAGE;
    const u32 RN32 = (u32)P.RN;
    // tile and element indices fit 32 bits (elem_vox holds 32-bit voxel indices, so n_el < 2^32)
    const u32 n_el = (u32)P.n_el, n_tiles = (u32)P.n_warp_tiles;
#if EE2_DYNAMIC
    // chunk c = tiles [c * EE2_CHUNK, (c + 1) * EE2_CHUNK); the first gridDim.x * EE2_WARPS chunks are the warps' own,
    // the counter (zeroed by the launcher) hands out the rest
    u32 wt = (blockIdx.x * EE2_WARPS + (threadIdx.x >> 5)) * EE2_CHUNK;
    u32 grabbed = 0;
#else
    const u32 stride = gridDim.x * EE2_WARPS;
    u32 wt = blockIdx.x * EE2_WARPS + (threadIdx.x >> 5);
#endif
    // the low words of first[] are all there is (first[g] <= n_el)
    const u32* first32 = reinterpret_cast<const u32*>(P.first);
    const u32 first_single = PERM ? 0u : __ldg(first32 + 2 * P.single_gv);
    // software pipeline: the global loads of the NEXT tile are in flight while the current one is formatted / flushed
    u32 nx_v = 0, nx_rank = 0, nx_flags = 0, nx_len = 0;
    int nx_gv = P.single_gv;
    u64 nx_base = 0;
    // GV-major order: grey value block of the warp's current tile start.  Within a chunk a warp visits its tiles
    // in ascending order, so the block only moves forward: one (warp-uniform) probe per tile instead of a
    // search per line; lanes look further only in the rare tile that contains a block boundary.
    int g_tile = 0;
    auto load_tile = [&](u32 t) {
        const u32 jj = t * 32u + lane;
        if (PERM) {
            const u32 j0 = t * 32u;
            while (g_tile < 255 && __ldg(first32 + 2 * (g_tile + 1)) <= j0) ++g_tile;
            nx_gv = g_tile;
            if (g_tile < 255 && __ldg(first32 + 2 * (g_tile + 1)) <= j0 + 31u)
                while (nx_gv < 255 && __ldg(first32 + 2 * (nx_gv + 1)) <= jj) ++nx_gv;
        }
        if (jj < n_el) {
            if (PERM) { const uint2 e = __ldg(P.perm + jj); nx_v = e.x; nx_rank = e.y; }
            else { nx_v = __ldg(P.elem_vox + jj); nx_rank = jj; }
            nx_flags = __ldg(P.nd8 + jj);
            nx_len = (u32)__ldg(P.len8 + jj);
        }
        nx_base = __ldg(P.warp_off + t);
    };
    if (wt < n_tiles) {
        if (PERM) g_tile = gv_of_index(P.first, (u64)wt * 32u);
        load_tile(wt);
    }
    while (wt < n_tiles) {
        const u32 j = wt * 32u + lane;
        const bool active = j < n_el;
        const u32 v = nx_v, rank = nx_rank, flags = nx_flags, line_len = nx_len;
        const int gv = nx_gv;
        const u64 base = nx_base;
        // the tile after this one
#if EE2_DYNAMIC
        u32 nxt = wt + 1u;
        if ((wt & (EE2_CHUNK - 1u)) == EE2_CHUNK - 2u) {
            // one tile before the chunk ends: ask for the next chunk (the answer is read an iteration later)
            if (lane == 0) grabbed = atomicAdd(P.tile_counter, 1u);
        } else if ((wt & (EE2_CHUNK - 1u)) == EE2_CHUNK - 1u) {
            nxt = (gridDim.x * EE2_WARPS + __shfl_sync(CFE_FULL_MASK, grabbed, 0)) * EE2_CHUNK;
        }
#else
        const u32 nxt = wt + stride;
#endif
#if !EE2_LATE_LOADS
        if (nxt < n_tiles) load_tile(nxt);
#endif
        const bool is_first = active && j == (PERM ? __ldg(first32 + 2 * gv) : first_single);
        u32 hl = 0;
        if (is_first) hl = __ldg(P.hdr_len + gv);
        const u32 len = active ? line_len + hl : 0u;   // the block header rides along with its first line
        // byte offset of every line in the tile: lane * len when all 32 lengths agree (the rule: one vote
        // and one add-reduction), else a warp scan
        const u32 total = __reduce_add_sync(CFE_FULL_MASK, len);
        u32 excl;
        if (__all_sync(CFE_FULL_MASK, len * 32u == total)) {
            excl = lane * len;
        } else {
            u32 incl = len;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const u32 t = __shfl_up_sync(CFE_FULL_MASK, incl, o);
                if (lane >= (u32)o) incl += t;
            }
            excl = incl - len;
        }
        char* dst = P.out + base;
        const u32 pad = (u32)(reinterpret_cast<uintptr_t>(dst) & 15u);
        const bool staged = pad + total <= (u32)ELEM_WARP_STAGE;
        LineWriter w;
        w.wp = stage; w.acc = 0; w.d = 0;
        if (active) {
            const u32 s = fd_div(v, P.dYX);
            const u32 rem = v - s * P.YX32;
            const u32 r = fd_div(rem, P.dX);
            const u32 c = rem - r * P.X32;
            id_t b, eid;
            if (WIDE) {
                b = (id_t)(P.SN * ((i64)s + P.z0) + P.RN * (i64)r + (i64)c);
                eid = (id_t)(P.eid_offset + (i64)rank + 1);
            } else {
                b = (id_t)((u32)P.SN * (s + (u32)P.z0) + RN32 * r + c);
                eid = (id_t)((u32)P.eid_offset + rank + 1u);
            }
            if (is_first && P.blk_off) P.blk_off[gv] = base + excl;
            if (staged) {
                // eid, b, b+1, b+RN+1, b+RN | b+SN, b+SN+1, b+SN+RN+1, b+SN+RN   (voxelFE.py:145-154)
                u32 o = pad + excl;
                if (hl) {
                    // block header: byte copy; the bytes it leaves in the line's first word start the accumulator
                    const char* h = P.hdr_text + gv * ELEM_HDR_SLOT;
                    for (u32 q = 0; q < hl; ++q) stage[o + q] = __ldg(h + q);
                    o += hl;
                    for (u32 q = 0; q < (o & 3u) && q < hl; ++q)
                        w.acc |= (u32)(unsigned char)__ldg(h + hl - 1u - q) << (8u * ((o & 3u) - 1u - q));
                }
                CFE_ASSERT(o + line_len <= (u32)ELEM_WARP_STAGE && o + line_len <= pad + total);
                w.d = o & 3u;
                w.wp = stage + (o - w.d);
                const bool same = (flags & 0x80u) != 0;
                const int nlo = (int)(flags & 15u);
                u32 e0, e1, e2, low4;
                asc12<WIDE>(lut, eid, e0, e1, e2, low4);
                const int nde = same ? (int)line_len - 9 - 8 * nlo : id_digits<WIDE>(eid);
                lw_append(w, e0 - 0x30303030u, e1 - 0x30303030u, e2 - 0x30303030u, nde, 0x2Cu);
                u32 F0[4][3], F1[4][3];
                layer_fields<WIDE>(lut, b, RN32, F0);
                layer_fields<WIDE>(lut, b + (id_t)P.SN, RN32, F1);
                if (same) {
                    switch (nlo) {
                    case 1: lw_append_line8<1>(w, F0, F1); break;
                    case 2: lw_append_line8<2>(w, F0, F1); break;
                    case 3: lw_append_line8<3>(w, F0, F1); break;
                    case 4: lw_append_line8<4>(w, F0, F1); break;
                    case 5: lw_append_line8<5>(w, F0, F1); break;
                    case 6: lw_append_line8<6>(w, F0, F1); break;
                    case 7: lw_append_line8<7>(w, F0, F1); break;
                    case 8: lw_append_line8<8>(w, F0, F1); break;
                    case 9: lw_append_line8<9>(w, F0, F1); break;
                    default: lw_append_line8<10>(w, F0, F1); break;   // ids are < 10^10 (checked by the host)
                    }
                } else {
                    lw_append_half<false>(w, F0, nlo, false);
                    lw_append_half<true>(w, F1, nlo, false);
                }
            } else {
                // pathological warp tile (dozens of block headers): direct byte writes
                char* p = dst + excl;
                if (hl) {
                    const char* h = P.hdr_text + gv * ELEM_HDR_SLOT;
                    for (u32 q = 0; q < hl; ++q) p[q] = __ldg(h + q);
                    p += hl;
                }
#pragma unroll 1
                for (int f = 0; f < 9; ++f) {
                    // eid, b, b+1, b+RN+1, b+RN, b+SN, b+SN+1, b+SN+RN+1, b+SN+RN
                    u64 val = (u64)eid;
                    if (f) {
                        const int q = f - 1;
                        val = (u64)b + ((q & 4) ? (u64)P.SN : 0ull) + (((q & 3) == 1 || (q & 3) == 2) ? 1ull : 0ull) +
                              (((q & 3) >= 2) ? (u64)P.RN : 0ull);
                    }
                    const int n = digits_u64(val);
                    p += n;
                    put_u64_back(p, val, n);
                    *p++ = (f == 8) ? '\n' : ',';
                }
            }
        }
        __syncwarp();     // every line's whole words are in the staging area
        if (staged) {
            // the formatted line ends where the length pass said it would
            CFE_ASSERT(!active || (u32)(w.wp - stage) + w.d == pad + excl + len);
            if (active) lw_finish(w);
#if EE2_BULK_FLUSH
            fence_proxy_async_smem();     // every lane: its staged bytes, for the bulk copy lane 0 issues below
#endif
            __syncwarp();
            // stream the warp's span [pad, pad + total) to dst: aligned 16-byte chunks, bytes at the edges
            char* gbase = dst - pad;
            const u32 end = pad + total;
            const u32 c0 = pad ? 1u : 0u, c1 = end >> 4;
#if EE2_BULK_FLUSH
            // the aligned middle leaves through the bulk-copy engine (cp.async.bulk shared -> global): one
            // instruction of one lane instead of an LDS.128 + STG.128 loop of the whole warp
            if (lane == 0 && c1 > c0) {
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                             :: "l"(gbase + 16u * c0), "r"((u32)__cvta_generic_to_shared(stage + 16u * c0)),
                                "r"(16u * (c1 - c0)) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
#else
#pragma unroll
            for (u32 it = 0; it < (ELEM_WARP_STAGE / 512u); ++it) {
                const u32 c = c0 + lane + 32u * it;
                if (c < c1)
                    *reinterpret_cast<uint4*>(gbase + 16u * c) = *reinterpret_cast<const uint4*>(stage + 16u * c);
            }
#endif
            // the two partial chunks: lanes 0-15 take the head [pad, 16), lanes 16-31 the tail [16 c1, end)
            {
                const bool head = lane < 16u;
                const u32 k = head ? lane : (c1 << 4) + lane - 16u;
                const bool ok = k < end && (head ? (pad != 0u && k >= pad) : c1 >= c0);
                if (ok) gbase[k] = stage[k];
            }
        }
#if EE2_LATE_LOADS
        if (nxt < n_tiles) load_tile(nxt);
#endif
#if EE2_BULK_FLUSH
        // the bulk copy must have READ the staging region before the next tile overwrites it
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
#endif
        __syncwarp();     // the staging region is reused by the next tile
        wt = nxt;
    }
#if EE2_BULK_FLUSH
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
#endif
}

extern "C" size_t cfe_emit_elements_workspace_bytes(int64_t n_el)
{
    // len8 + nd8 (1 byte each, 16-byte aligned), warp_total (u32), warp_off (u64), scan workspace
    const size_t nw = (size_t)((n_el + 31) / 32);
    const size_t a = ((size_t)n_el + 15) & ~(size_t)15;
    return 2 * a + ((nw * 4 + 15) & ~(size_t)15) + nw * 8 + 16 + cfe_scan_workspace_bytes((int64_t)nw) + 64;
}

static void elem_params(ElemEmitParams& P, const uint32_t* elem_vox, const void* perm, int64_t n_el,
                        const uint8_t* vol, const uint64_t* first, const char* hdr_text, const uint8_t* hdr_len,
                        int single_gv, int64_t Y, int64_t X, int64_t z0, int64_t eid_offset, char* out,
                        uint64_t* blk_off, void* ws, void** scan_ws)
{
    const i64 nw = (n_el + 31) / 32;
    P.elem_vox = elem_vox; P.perm = reinterpret_cast<const uint2*>(perm); P.n_el = n_el; P.vol = vol;
    P.first = reinterpret_cast<const u64*>(first); P.hdr_text = hdr_text; P.hdr_len = hdr_len;
    P.single_gv = single_gv;
    P.YX32 = (u32)(Y * X); P.X32 = (u32)X;
    P.dYX = make_fastdiv(P.YX32); P.dX = make_fastdiv(P.X32);
    P.RN = X + 1; P.SN = (X + 1) * (Y + 1); P.z0 = z0; P.eid_offset = eid_offset;
    P.out = out; P.blk_off = reinterpret_cast<u64*>(blk_off);
    // carve the workspace
    char* w = reinterpret_cast<char*>(ws);
    const size_t a = ((size_t)n_el + 15) & ~(size_t)15;
    P.len8 = reinterpret_cast<uint8_t*>(w); w += a;
    P.nd8 = reinterpret_cast<uint8_t*>(w); w += a;
    P.warp_total = reinterpret_cast<u32*>(w); w += ((size_t)nw * 4 + 15) & ~(size_t)15;
    P.warp_off = reinterpret_cast<u64*>(w); w += (size_t)nw * 8;
    P.tile_counter = reinterpret_cast<u32*>(w); w += 16;
    *scan_ws = w;
    P.n_warp_tiles = nw;
}

static bool elem_wide(const ElemEmitParams& P, int64_t Zl)
{
    // largest id that can appear in this slab's lines
    const unsigned long long max_node = (unsigned long long)P.SN * (unsigned long long)(P.z0 + Zl + 1);
    return (max_node >= 0xffffffffull) || ((unsigned long long)(P.eid_offset + P.n_el) >= 0xffffffffull);
}

// Pass 1 of the *ELEMENT section: line lengths, 32-line tile totals and their exclusive scan.
// *total = bytes of the section (block headers included).  Same arguments as cfe_emit_elements.
extern "C" int cfe_element_lengths(const uint32_t* elem_vox, const void* perm, int64_t n_el, const uint8_t* vol,
                                   const uint64_t* first, const char* hdr_text, const uint8_t* hdr_len,
                                   int single_gv, int64_t Zl, int64_t Y, int64_t X, int64_t z0, int64_t eid_offset,
                                   int have_lengths, uint64_t* total, void* ws, cudaStream_t stream)
{
    CFE_CHECK_CUDA(cudaMemsetAsync(total, 0, sizeof(u64), stream));
    if (n_el <= 0) return 0;
    ElemEmitParams P;
    void* scan_ws;
    elem_params(P, elem_vox, perm, n_el, vol, first, hdr_text, hdr_len, single_gv, Y, X, z0, eid_offset, nullptr,
                nullptr, ws, &scan_ws);
    const bool wide = elem_wide(P, Zl);
    const unsigned g1 = (unsigned)((n_el + 255) / 256);
    if (have_lengths) {
        // len8 / nd8 were written by cfe_fill_elements (same z0 / eid_offset) and, for GV-major order,
        // permuted by cfe_gv_partition: only the tile totals
        k_tile_totals<<<(unsigned)(((n_el + 15) / 16 + 255) / 256), 256, 0, stream>>>(P);
        CFE_CHECK_LAUNCH("k_tile_totals");
        if (perm) {
            k_tile_headers<<<1, 256, 0, stream>>>(P);
            CFE_CHECK_LAUNCH("k_tile_headers");
        }
    } else {
        if (perm) {
            if (wide) k_elem_lengths<true, true><<<g1, 256, 0, stream>>>(P);
            else k_elem_lengths<true, false><<<g1, 256, 0, stream>>>(P);
        } else {
            if (wide) k_elem_lengths<false, true><<<g1, 256, 0, stream>>>(P);
            else k_elem_lengths<false, false><<<g1, 256, 0, stream>>>(P);
        }
        CFE_CHECK_LAUNCH("k_elem_lengths");
    }
    return cfe_scan_u32(P.warp_total, P.n_warp_tiles, reinterpret_cast<uint64_t*>(P.warp_off),
                        total, scan_ws, stream);
}

// Pass 2: format the lines (requires cfe_element_lengths on the same workspace).
extern "C" int cfe_emit_elements(const uint32_t* elem_vox, const void* perm, int64_t n_el, const uint8_t* vol,
                                 const uint64_t* first, const char* hdr_text, const uint8_t* hdr_len,
                                 int single_gv, int64_t Zl, int64_t Y, int64_t X, int64_t z0, int64_t eid_offset,
                                 char* out, uint64_t* blk_off, void* ws, cudaStream_t stream)
{
    if (n_el <= 0) return 0;
    ElemEmitParams P;
    void* scan_ws;
    elem_params(P, elem_vox, perm, n_el, vol, first, hdr_text, hdr_len, single_gv, Y, X, z0, eid_offset, out,
                blk_off, ws, &scan_ws);
    const bool wide = elem_wide(P, Zl);
    static unsigned long long attr_done = 0;
    int sms = 148;
    if (cfe_first_on_device(&attr_done, &sms)) {
        CFE_CHECK_CUDA(cudaFuncSetAttribute(k_emit_elements2<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, EE2_SMEM));
        CFE_CHECK_CUDA(cudaFuncSetAttribute(k_emit_elements2<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, EE2_SMEM));
        CFE_CHECK_CUDA(cudaFuncSetAttribute(k_emit_elements2<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, EE2_SMEM));
        CFE_CHECK_CUDA(cudaFuncSetAttribute(k_emit_elements2<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, EE2_SMEM));
    }
    // persistent CTAs: two per SM (97 KB of shared memory each)
#if EE2_DYNAMIC
    // every warp starts with its own chunk of EE2_CHUNK tiles and takes the following ones from the counter
    CFE_CHECK_CUDA(cudaMemsetAsync(P.tile_counter, 0, sizeof(u32), stream));
    i64 g2 = ((P.n_warp_tiles + EE2_CHUNK - 1) / EE2_CHUNK + EE2_WARPS - 1) / EE2_WARPS;
#else
    i64 g2 = (P.n_warp_tiles + EE2_WARPS - 1) / EE2_WARPS;     // every warp strides over the warp tiles
#endif
    if (g2 > 2 * (i64)sms) g2 = 2 * (i64)sms;
    if (perm) {
        if (wide) k_emit_elements2<true, true><<<(unsigned)g2, EE2_THREADS, EE2_SMEM, stream>>>(P);
        else k_emit_elements2<true, false><<<(unsigned)g2, EE2_THREADS, EE2_SMEM, stream>>>(P);
    } else {
        if (wide) k_emit_elements2<false, true><<<(unsigned)g2, EE2_THREADS, EE2_SMEM, stream>>>(P);
        else k_emit_elements2<false, false><<<(unsigned)g2, EE2_THREADS, EE2_SMEM, stream>>>(P);
    }
    CFE_CHECK_LAUNCH("k_emit_elements2");
    return 0;
}

// ------------------------------------------------------------------------------------------
// *NSET / *ELSET sections: a stream of items, each either a literal (section comment, set
// header, the closing '\n' of a set) or one id of a set's list followed by ',' or '\n'.

#define ITEM_PER_THREAD 8
#define ITEM_TILE (EMIT_THREADS * ITEM_PER_THREAD)
#define ITEM_STAGE_BYTES (ITEM_TILE * 12 + 8192 + 32)

// One CTA = ITEM_TILE consecutive items (8 per thread): few, fat tiles keep the look-back chain short.
__global__ void __launch_bounds__(EMIT_THREADS)
k_emit_items(const CfeItemSeg* __restrict__ segs, int n_segs, i64 n_items, const i64* __restrict__ ids,
             const char* __restrict__ lit, char* out_base, const u64* dyn_off, u64* total, u64* seg_off, u64* ws)
{
    extern __shared__ __align__(16) char stage[];
    __shared__ u32 s_scan[33];
    __shared__ u64 s_base;
    const u32 tile = scan_next_tile(ws);
    const i64 t0 = (i64)tile * ITEM_TILE + (i64)threadIdx.x * ITEM_PER_THREAD;
    u32 len[ITEM_PER_THREAD];
    u64 id[ITEM_PER_THREAD];
    int lit_seg[ITEM_PER_THREAD];       // >= 0: literal of that segment
    u32 first_mask = 0, nl_mask = 0, sum = 0;
    int q = 0;
    if (t0 < n_items)
        while (q + 1 < n_segs && t0 >= segs[q + 1].item0) ++q;
#pragma unroll
    for (int j = 0; j < ITEM_PER_THREAD; ++j) {
        const i64 t = t0 + j;
        len[j] = 0; id[j] = 0; lit_seg[j] = -1;
        if (t < n_items) {
            while (q + 1 < n_segs && t >= segs[q + 1].item0) ++q;
            const i64 item0 = segs[q].item0;
            if (t == item0) first_mask |= 1u << j;
            if (segs[q].kind == 0) {
                lit_seg[j] = q;
                len[j] = (u32)segs[q].lit_len;
            } else {
                const i64 e = t - item0;
                id[j] = (u64)__ldg(ids + segs[q].id0 + e);
                // voxelFE.py:657-666: ',' after entries 1-9 of a line, '\n' after the tenth
                if (((segs[q].rank0 + e + 1) % 10) == 0) nl_mask |= 1u << j;
                len[j] = (u32)digits_u64(id[j]) + 1u;
            }
        }
        sum += len[j];
    }
    u32 tile_total;
    u32 excl = block_exclusive_scan<u32>(sum, s_scan, tile_total);
    if (threadIdx.x < 32) {
        const u64 e = scan_lookback_d<4>(ws, tile, (u64)tile_total);
        if (threadIdx.x == 0) {
            s_base = e;
            if ((i64)(tile + 1) * ITEM_TILE >= n_items) *total = e + tile_total;
        }
    }
    __syncthreads();
    char* dst = out_base + (dyn_off ? *dyn_off : 0ull) + s_base;
    const u32 pad = (u32)(reinterpret_cast<uintptr_t>(dst) & 15u);
    const bool staged = (pad + tile_total) <= (u32)ITEM_STAGE_BYTES;
    char* p = staged ? (stage + pad + excl) : (dst + excl);
#pragma unroll
    for (int j = 0; j < ITEM_PER_THREAD; ++j) {
        if (t0 + j >= n_items) break;
        if ((first_mask >> j) & 1u) {
            // segment index of item j: recompute (rare path)
            int qq = 0;
            while (qq + 1 < n_segs && t0 + j >= segs[qq + 1].item0) ++qq;
            if (seg_off) seg_off[qq] = s_base + excl;
        }
        if (lit_seg[j] >= 0) {
            const char* src = lit + segs[lit_seg[j]].lit_off;
            for (u32 c = 0; c < len[j]; ++c) p[c] = __ldg(src + c);
        } else {
            const int nd = (int)len[j] - 1;
            put_u64_back(p + nd, id[j], nd);
            p[nd] = ((nl_mask >> j) & 1u) ? '\n' : ',';
        }
        p += len[j];
        excl += len[j];
    }
    __syncthreads();
    if (staged) flush_stage(stage, pad, tile_total, dst);
}

extern "C" int cfe_emit_items(const CfeItemSeg* segs_dev, int n_segs, int64_t n_items, const int64_t* ids,
                              const char* literals, char* out_base, const uint64_t* dyn_off, uint64_t* total,
                              uint64_t* seg_off, void* ws, cudaStream_t stream)
{
    const i64 tiles = (n_items + ITEM_TILE - 1) / ITEM_TILE;
    CFE_CHECK_CUDA(cudaMemsetAsync(ws, 0, (size_t)(tiles + 2) * sizeof(u64), stream));
    CFE_CHECK_CUDA(cudaMemsetAsync(total, 0, sizeof(u64), stream));
    if (tiles == 0) return 0;
    static unsigned long long attr_done = 0;
    if (cfe_first_on_device(&attr_done)) {
        CFE_CHECK_CUDA(cudaFuncSetAttribute(k_emit_items, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            ITEM_STAGE_BYTES));
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
// Stable partition of the raster-ordered element list by grey value (np.where(cell_GV == GV),
// voxelFE.py:645): one 8-bit radix pass.
//   k_part_hist     per-tile 256-bin histogram, TILE-major  mat[tile][gv] (coalesced), and the sums of
//                   every chunk of 256 tiles by atomics  chunk[c][gv]
//   k_part_chunks   one CTA: exclusive scan of the chunk sums per grey value, bin totals -> bin starts
//   k_part_offsets  per chunk: mat[tile][gv] <- destination of the first element of (gv, tile)
//   k_part_scatter  the tile is ranked stably in shared memory (warp match.any + per-warp cursors) and
//                   leaves in GV-major order, so consecutive threads write consecutive destinations;
//                   the raster-order line lengths (cfe_fill_elements) travel with the elements.

#define PART_THREADS 256
#define PART_WARPS 8
#define PART_STEPS 16
#define PART_TILE (PART_THREADS * PART_STEPS)   // 4096 elements per CTA
#define PART_CHUNK 256                          // tiles per chunk

static inline i64 part_tiles(i64 n_el) { return (n_el + PART_TILE - 1) / PART_TILE; }
static inline i64 part_chunks(i64 n_el) { return (part_tiles(n_el) + PART_CHUNK - 1) / PART_CHUNK; }

extern "C" size_t cfe_gv_partition_workspace_bytes(int64_t n_el)
{
    if (n_el <= 0) return 64;
    return ((size_t)part_tiles(n_el) + (size_t)part_chunks(n_el) + 1) * 256 * sizeof(u32) + 64;
}

// 16 consecutive bytes per thread into shared memory (one 16-byte load when possible)
__device__ __forceinline__ void part_stage_bytes(const uint8_t* __restrict__ src, i64 tile0, u32 nval,
                                                 uint8_t* __restrict__ dst)
{
    const u32 o = threadIdx.x * 16u;
    const uint8_t* p = src + tile0 + o;
    if (o + 16u <= nval && (reinterpret_cast<uintptr_t>(p) & 15u) == 0) {
        *reinterpret_cast<uint4*>(dst + o) = __ldg(reinterpret_cast<const uint4*>(p));
    } else {
        for (u32 k = 0; k < 16u && o + k < nval; ++k) dst[o + k] = __ldg(p + k);
    }
}

__global__ void __launch_bounds__(PART_THREADS)
k_part_hist(const uint8_t* __restrict__ vol, const u32* __restrict__ elem_vox, i64 n_el,
            u32* __restrict__ mat, u32* __restrict__ chunk, const uint8_t* __restrict__ gv8)
{
    __shared__ u32 h[256];
    __shared__ __align__(16) uint8_t g[PART_TILE];
    h[threadIdx.x] = 0;
    const i64 tile = blockIdx.x;
    const i64 tile0 = tile * PART_TILE;
    const u32 nval = (u32)((n_el - tile0 < PART_TILE) ? (n_el - tile0) : PART_TILE);
    const u32 o = threadIdx.x * 16u;
    if (gv8) {
        part_stage_bytes(gv8, tile0, nval, g);
    } else {
        for (u32 k = 0; k < 16u && o + k < nval; ++k) g[o + k] = __ldg(vol + __ldg(elem_vox + tile0 + o + k));
    }
    __syncthreads();
    // runs of equal grey values (smooth fields) are added once
    int prev = -1;
    u32 run = 0;
    for (u32 k = 0; k < 16u && o + k < nval; ++k) {
        const int b = g[o + k];
        if (b == prev) { ++run; continue; }
        if (run) atomicAdd(&h[prev], run);
        prev = b;
        run = 1;
    }
    if (run) atomicAdd(&h[prev], run);
    __syncthreads();
    const u32 c = h[threadIdx.x];
    mat[tile * 256 + threadIdx.x] = c;
    if (c) atomicAdd(chunk + (tile / PART_CHUNK) * 256 + threadIdx.x, c);
}

// chunk[c][g] <- number of elements with grey value g in the chunks before c; binbase[g] <- number of
// elements with a smaller grey value
__global__ void __launch_bounds__(256) k_part_chunks(u32* __restrict__ chunk, i64 n_chunks, u32* __restrict__ binbase)
{
    __shared__ u32 sc[33];
    const u32 g = threadIdx.x;
    u32 run = 0;
#pragma unroll 8
    for (i64 c = 0; c < n_chunks; ++c) {
        const u32 v = chunk[c * 256 + g];
        chunk[c * 256 + g] = run;
        run += v;
    }
    u32 total;
    binbase[g] = block_exclusive_scan<u32>(run, sc, total);
}

__global__ void __launch_bounds__(256)
k_part_offsets(u32* __restrict__ mat, const u32* __restrict__ chunk, const u32* __restrict__ binbase, i64 n_tiles)
{
    const u32 g = threadIdx.x;
    const i64 c = blockIdx.x;
    u32 run = chunk[c * 256 + g] + binbase[g];
    const i64 t0 = c * PART_CHUNK;
    const i64 t1 = (t0 + PART_CHUNK < n_tiles) ? t0 + PART_CHUNK : n_tiles;
#pragma unroll 8
    for (i64 t = t0; t < t1; ++t) {
        const u32 v = mat[t * 256 + g];
        mat[t * 256 + g] = run;
        run += v;
    }
}

__global__ void __launch_bounds__(PART_THREADS, 4)
k_part_scatter(const uint8_t* __restrict__ vol, const u32* __restrict__ elem_vox, i64 n_el,
               const u32* __restrict__ mat_off, uint2* __restrict__ perm, const uint8_t* __restrict__ gv8,
               const uint8_t* __restrict__ len_src, const uint8_t* __restrict__ nd_src,
               uint8_t* __restrict__ len_dst, uint8_t* __restrict__ nd_dst)
{
    __shared__ __align__(16) u32 voxin[PART_TILE];
    __shared__ u32 wh[PART_WARPS][256];
    __shared__ u32 gbase[256];
    __shared__ u32 sc[33];
    __shared__ unsigned short jl[PART_TILE];
    __shared__ __align__(16) uint8_t gvin[PART_TILE];
    __shared__ __align__(16) uint8_t lenin[PART_TILE];
    __shared__ __align__(16) uint8_t ndin[PART_TILE];

    const i64 tile = blockIdx.x;
    const i64 tile0 = tile * PART_TILE;
    const u32 nval = (u32)((n_el - tile0 < PART_TILE) ? (n_el - tile0) : PART_TILE);
    const u32 warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const u32 my_off = __ldg(mat_off + tile * 256 + threadIdx.x);   // destination of (gv = thread, this tile)

    // ---- stage the tile (16 consecutive elements per thread)
    {
        const u32 o = threadIdx.x * 16u;
        const u32* p = elem_vox + tile0 + o;
        if (o + 16u <= nval && (reinterpret_cast<uintptr_t>(p) & 15u) == 0) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
                *reinterpret_cast<uint4*>(voxin + o + 4 * q) = __ldg(reinterpret_cast<const uint4*>(p) + q);
        } else {
            for (u32 k = 0; k < 16u && o + k < nval; ++k) voxin[o + k] = __ldg(p + k);
        }
        if (gv8) {
            part_stage_bytes(gv8, tile0, nval, gvin);
        } else {
            for (u32 k = 0; k < 16u && o + k < nval; ++k) gvin[o + k] = __ldg(vol + voxin[o + k]);
        }
        if (len_src) {
            part_stage_bytes(len_src, tile0, nval, lenin);
            part_stage_bytes(nd_src, tile0, nval, ndin);
        }
    }
    for (int i = threadIdx.x; i < PART_WARPS * 256; i += PART_THREADS) (&wh[0][0])[i] = 0;
    __syncthreads();

    // ---- per-warp counts; warp w owns the 512 consecutive elements [512 w, 512 w + 512), step t its
    // elements 32 t .. 32 t + 31, so (warp, step, lane) order is raster order
    int gv[PART_STEPS];
    u32 peers[PART_STEPS];
#pragma unroll
    for (int t = 0; t < PART_STEPS; ++t) {
        const u32 idx = warp * (PART_STEPS * 32) + t * 32 + lane;
        gv[t] = (idx < nval) ? (int)gvin[idx] : 256;
        peers[t] = __match_any_sync(CFE_FULL_MASK, gv[t]);
        if (gv[t] < 256 && (peers[t] & ((1u << lane) - 1u)) == 0) atomicAdd(&wh[warp][gv[t]], (u32)__popc(peers[t]));
    }
    __syncthreads();
    {
        // thread g: bin total -> start of bin g in the sorted tile -> per-warp cursors
        const int g = threadIdx.x;
        u32 tot = 0;
#pragma unroll
        for (int w = 0; w < PART_WARPS; ++w) tot += wh[w][g];
        u32 total;
        const u32 start = block_exclusive_scan<u32>(tot, sc, total);
        u32 run = start;
#pragma unroll
        for (int w = 0; w < PART_WARPS; ++w) {
            const u32 c = wh[w][g];
            wh[w][g] = run;
            run += c;
        }
        gbase[g] = my_off - start;      // destination of sorted position i with grey value g = gbase[g] + i (mod 2^32)
    }
    __syncthreads();
#pragma unroll
    for (int t = 0; t < PART_STEPS; ++t) {
        const u32 idx = warp * (PART_STEPS * 32) + t * 32 + lane;
        u32 base = 0;
        if (gv[t] < 256) base = wh[warp][gv[t]];
        __syncwarp();                       // every peer has read the cursor before its leader advances it
        if (gv[t] < 256) {
            const u32 rank = (u32)__popc(peers[t] & ((1u << lane) - 1u));
            CFE_ASSERT(base + rank < nval);
            jl[base + rank] = (unsigned short)idx;
            if (rank == 0) wh[warp][gv[t]] = base + (u32)__popc(peers[t]);
        }
        __syncwarp();
    }
    __syncthreads();
    // ---- the sorted tile leaves: runs of equal grey value go to consecutive destinations
#pragma unroll 4
    for (u32 i = threadIdx.x; i < nval; i += PART_THREADS) {
        const u32 j = jl[i];
        CFE_ASSERT(j < nval);
        const u32 d = gbase[gvin[j]] + i;
        CFE_ASSERT((i64)d < n_el);
        perm[d] = make_uint2(voxin[j], (u32)(tile0 + j));
        if (len_src) {
            len_dst[d] = lenin[j];
            nd_dst[d] = ndin[j];
        }
    }
}

// perm[k] = {local voxel index, raster rank} of the k-th element in GV-major order.  gv8 (optional): grey
// values in element order (cfe_fill_elements_gv); without it the kernels gather vol[elem_vox[j]].
// ws_src / ws_dst (optional, both or neither): element workspaces; the line lengths that cfe_fill_elements
// left in ws_src (raster order) are written to ws_dst in GV-major order, which is what
// cfe_element_lengths(perm, have_lengths = 1) and cfe_emit_elements(perm) then read.
extern "C" int cfe_gv_partition(const uint8_t* vol, const uint32_t* elem_vox, int64_t n_el, const uint8_t* gv8,
                                const void* ws_src, void* ws_dst, void* perm, void* ws_part, cudaStream_t stream)
{
    if (n_el <= 0) return 0;
    if (n_el >= (1LL << 32)) { cfe_set_error("cfe_gv_partition: too many elements"); return -1; }
    if ((ws_src == nullptr) != (ws_dst == nullptr)) { cfe_set_error("cfe_gv_partition: ws_src and ws_dst go together"); return -1; }
    const i64 n_tiles = part_tiles(n_el), n_chunks = part_chunks(n_el);
    u32* mat = reinterpret_cast<u32*>(ws_part);
    u32* chunk = mat + n_tiles * 256;
    u32* binbase = chunk + n_chunks * 256;
    const size_t a = ((size_t)n_el + 15) & ~(size_t)15;      // layout of the element workspace (elem_params)
    const uint8_t* len_src = reinterpret_cast<const uint8_t*>(ws_src);
    uint8_t* len_dst = reinterpret_cast<uint8_t*>(ws_dst);
    CFE_CHECK_CUDA(cudaMemsetAsync(chunk, 0, (size_t)n_chunks * 256 * sizeof(u32), stream));
    k_part_hist<<<(unsigned)n_tiles, PART_THREADS, 0, stream>>>(vol, elem_vox, n_el, mat, chunk, gv8);
    CFE_CHECK_LAUNCH("k_part_hist");
    k_part_chunks<<<1, 256, 0, stream>>>(chunk, n_chunks, binbase);
    CFE_CHECK_LAUNCH("k_part_chunks");
    k_part_offsets<<<(unsigned)n_chunks, 256, 0, stream>>>(mat, chunk, binbase, n_tiles);
    CFE_CHECK_LAUNCH("k_part_offsets");
    k_part_scatter<<<(unsigned)n_tiles, PART_THREADS, 0, stream>>>(
        vol, elem_vox, n_el, mat, reinterpret_cast<uint2*>(perm), gv8, len_src, len_src ? len_src + a : nullptr,
        len_dst, len_dst ? len_dst + a : nullptr);
    CFE_CHECK_LAUNCH("k_part_scatter");
    return 0;
}

// ------------------------------------------------------------------------------------------
// '{:12.6f}'.format(v) of the per-axis node coordinates (voxelFE.py:630) -- exact: the binary
// value times 10^6 is rounded half-to-even in 128-bit integer arithmetic, which is what CPython's
// correctly rounded float formatting does.  One thread per value, 16-byte output slots.
// *overflow is set when a field would be wider than 12 characters (|v| >= 1e5) or v is not finite.

__global__ void __launch_bounds__(128)
k_format_coords(const double* __restrict__ vals, i64 n, char* __restrict__ out16, int* __restrict__ overflow)
{
    const i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const u64 bits = (u64)__double_as_longlong(vals[t]);
    const bool neg = (bits >> 63) != 0;
    const int ex = (int)((bits >> 52) & 0x7ffu);
    u64 m = bits & 0xfffffffffffffull;
    int e;
    if (ex == 0) e = -1074; else { m |= 1ull << 52; e = ex - 1075; }
    char* o = out16 + 16 * t;
    u64 N = 0;
    bool bad = (ex == 0x7ff);
    if (!bad && m) {
        const unsigned __int128 P = (unsigned __int128)m * 1000000u;   // < 2^73
        if (e >= 0) {
            bad = true;                                               // >= 2^52: far beyond 1e5
        } else {
            const int sft = -e;
            if (sft < 128) {
                unsigned __int128 q = P >> sft;
                const unsigned __int128 rem = P - (q << sft);
                const unsigned __int128 half = (unsigned __int128)1 << (sft - 1);
                if (rem > half || (rem == half && ((u64)q & 1ull))) q += 1;
                if (q >= (unsigned __int128)100000000000ull) bad = true; else N = (u64)q;
            }
        }
    }
    if (N >= 100000000000ull) bad = true;                             // 100000.000000 needs 13 chars
    const u64 ip = N / 1000000ull;
    u32 fp = (u32)(N - ip * 1000000ull);
    if (neg && ip >= 10000ull) bad = true;                            // '-' + 5 digits + 7 = 13 chars
    if (bad) { *overflow = 1; return; }
    char buf[12];
#pragma unroll
    for (int k = 11; k >= 6; --k) { buf[k] = (char)('0' + fp % 10u); fp /= 10u; }
    buf[5] = '.';
    u32 iv = (u32)ip;
    int k = 4;
    do { buf[k--] = (char)('0' + iv % 10u); iv /= 10u; } while (iv && k >= 0);
    if (neg) buf[k--] = '-';
    while (k >= 0) buf[k--] = ' ';
#pragma unroll
    for (int q = 0; q < 12; ++q) o[q] = buf[q];
    o[12] = 0; o[13] = 0; o[14] = 0; o[15] = 0;
}

extern "C" int cfe_format_coords(const double* vals, int64_t n, char* out16, int* overflow, cudaStream_t stream)
{
    if (n <= 0) return 0;
    k_format_coords<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(vals, n, out16, overflow);
    CFE_CHECK_LAUNCH("k_format_coords");
    return 0;
}

// ------------------------------------------------------------------------------------------
// Wide coordinate fields: '{:12.6f}' grows beyond 12 characters when |coord| >= 1e5 (voxel sizes in
// micrometres on large scans), which makes the *NODE lines variable-length.  General (slower) path:
// exact formatting into 32-byte slots (slot[0] = length), per-node line lengths, tile offsets by
// cfe_scan_u32, byte-wise staging.

__global__ void __launch_bounds__(128)
k_format_coords_wide(const double* __restrict__ vals, i64 n, char* __restrict__ out32, int* __restrict__ bad)
{
    const i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const u64 bits = (u64)__double_as_longlong(vals[t]);
    const bool neg = (bits >> 63) != 0;
    const int ex = (int)((bits >> 52) & 0x7ffu);
    u64 m = bits & 0xfffffffffffffull;
    int e;
    if (ex == 0) e = -1074; else { m |= 1ull << 52; e = ex - 1075; }
    if (ex == 0x7ff || e > 7) { *bad = 1; return; }                 // inf / nan / >= 2^60
    unsigned __int128 q = (unsigned __int128)m * 1000000u;            // < 2^73
    if (e >= 0) {
        q <<= e;
    } else {
        const int sft = -e;
        if (sft >= 128) q = 0;
        else {
            const unsigned __int128 P = q;
            q = P >> sft;
            const unsigned __int128 rem = P - (q << sft);
            const unsigned __int128 half = (unsigned __int128)1 << (sft - 1);
            if (rem > half || (rem == half && ((u64)q & 1ull))) q += 1;
        }
    }
    const unsigned __int128 lim = (unsigned __int128)1000000000000000000ull * 1000000u;   // 1e24
    if (q >= lim) { *bad = 1; return; }
    u64 ip = (u64)(q / 1000000u);
    u32 fp = (u32)(q - (unsigned __int128)ip * 1000000u);
    char buf[32];
    int p = 32;
    for (int k = 0; k < 6; ++k) { buf[--p] = (char)('0' + fp % 10u); fp /= 10u; }
    buf[--p] = '.';
    do { buf[--p] = (char)('0' + (int)(ip % 10ull)); ip /= 10ull; } while (ip);
    if (neg) buf[--p] = '-';
    while (32 - p < 12) buf[--p] = ' ';
    const int len = 32 - p;                                           // 12 .. 26
    char* o = out32 + 32 * t;
    o[0] = (char)len;
    for (int k = 0; k < len; ++k) o[1 + k] = buf[p + k];
}

extern "C" int cfe_format_coords_wide(const double* vals, int64_t n, char* out32, int* bad, cudaStream_t stream)
{
    if (n <= 0) return 0;
    k_format_coords_wide<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(vals, n, out32, bad);
    CFE_CHECK_LAUNCH("k_format_coords_wide");
    return 0;
}

struct NodeWideParams {
    const u32* node_list;
    i64 n_nodes;
    const char* xtab;        // [X+1][32]
    const char* ytab;        // [Y+1][32]
    const char* ztab;        // [layers][32]
    FastDiv dSN, dRN;
    u32 SN32, RN32;
    i64 z0;
    uint8_t* len8;           // [n_nodes] line lengths
    u32* tile_tot;           // [tiles of 256 nodes]
    const u64* tile_off;
    char* out;
};

__device__ __forceinline__ void node_wide_fields(const NodeWideParams& P, u32 v, u64& id, u32& s, u32& r, u32& c)
{
    s = fd_div(v, P.dSN);
    const u32 rem = v - s * P.SN32;
    r = fd_div(rem, P.dRN);
    c = rem - r * P.RN32;
    id = (u64)P.SN32 * (u64)((i64)s + P.z0) + (u64)rem;
}

__global__ void __launch_bounds__(256) k_node_lengths_wide(NodeWideParams P)
{
    const i64 k = (i64)blockIdx.x * 256 + threadIdx.x;
    u32 len = 0;
    if (k < P.n_nodes) {
        u64 id; u32 s, r, c;
        node_wide_fields(P, __ldg(P.node_list + k), id, s, r, c);
        const int nd = digits_u64(id);
        len = (u32)(nd < 10 ? 10 : nd) + 2u + (u32)(uint8_t)P.xtab[32 * (size_t)c] + 2u +
              (u32)(uint8_t)P.ytab[32 * (size_t)r] + 2u + (u32)(uint8_t)P.ztab[32 * (size_t)s] + 1u;
        P.len8[k] = (uint8_t)len;
    }
    __shared__ u32 s_scan[33];
    u32 tot;
    block_exclusive_scan<u32>(len, s_scan, tot);
    if (threadIdx.x == 0) P.tile_tot[blockIdx.x] = tot;
}

#define NODEW_STAGE (256 * 120 + 32)

__global__ void __launch_bounds__(256) k_emit_nodes_wide(NodeWideParams P)
{
    extern __shared__ __align__(16) char stage[];
    __shared__ u32 s_scan[33];
    const i64 k = (i64)blockIdx.x * 256 + threadIdx.x;
    const u32 len = k < P.n_nodes ? (u32)P.len8[k] : 0u;
    u32 tot;
    const u32 excl = block_exclusive_scan<u32>(len, s_scan, tot);
    char* dst = P.out + P.tile_off[blockIdx.x];
    const u32 pad = (u32)(reinterpret_cast<uintptr_t>(dst) & 15u);
    if (k < P.n_nodes) {
        u64 id; u32 s, r, c;
        node_wide_fields(P, __ldg(P.node_list + k), id, s, r, c);
        char* p = stage + pad + excl;
        const int nd = digits_u64(id), w = nd < 10 ? 10 : nd;
        for (int q = 0; q < w - nd; ++q) *p++ = ' ';
        p += nd;
        put_u64_back(p, id, nd);
        const char* f[3] = {P.xtab + 32 * (size_t)c, P.ytab + 32 * (size_t)r, P.ztab + 32 * (size_t)s};
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            *p++ = ',';
            *p++ = ' ';
            const int fl = (int)(uint8_t)f[a][0];
            for (int q = 0; q < fl; ++q) *p++ = f[a][1 + q];
        }
        *p = '\n';
    }
    __syncthreads();
    flush_stage(stage, pad, tot, dst);
}

static void node_wide_params(NodeWideParams& P, const uint32_t* node_list, int64_t n_nodes, const char* xtab,
                             const char* ytab, const char* ztab, int64_t Y, int64_t X, int64_t z0, uint8_t* len8,
                             uint32_t* tile_tot, const uint64_t* tile_off, char* out)
{
    P.node_list = node_list; P.n_nodes = n_nodes; P.xtab = xtab; P.ytab = ytab; P.ztab = ztab;
    P.SN32 = (u32)((X + 1) * (Y + 1)); P.RN32 = (u32)(X + 1);
    P.dSN = make_fastdiv(P.SN32); P.dRN = make_fastdiv(P.RN32);
    P.z0 = z0; P.len8 = len8; P.tile_tot = tile_tot; P.tile_off = reinterpret_cast<const u64*>(tile_off); P.out = out;
}

// line length of every used node (len8) and byte total of every 256-node tile (tile_tot)
extern "C" int cfe_node_lengths_wide(const uint32_t* node_list, int64_t n_nodes, const char* xtab32,
                                     const char* ytab32, const char* ztab32, int64_t Y, int64_t X, int64_t z0,
                                     uint8_t* len8, uint32_t* tile_tot, cudaStream_t stream)
{
    if (n_nodes <= 0) return 0;
    NodeWideParams P;
    node_wide_params(P, node_list, n_nodes, xtab32, ytab32, ztab32, Y, X, z0, len8, tile_tot, nullptr, nullptr);
    k_node_lengths_wide<<<(unsigned)((n_nodes + 255) / 256), 256, 0, stream>>>(P);
    CFE_CHECK_LAUNCH("k_node_lengths_wide");
    return 0;
}

// the *NODE lines with variable-width fields; tile_off = exclusive scan of tile_tot
extern "C" int cfe_emit_nodes_wide(const uint32_t* node_list, int64_t n_nodes, const char* xtab32,
                                   const char* ytab32, const char* ztab32, int64_t Y, int64_t X, int64_t z0,
                                   const uint8_t* len8, const uint64_t* tile_off, char* out, cudaStream_t stream)
{
    if (n_nodes <= 0) return 0;
    NodeWideParams P;
    node_wide_params(P, node_list, n_nodes, xtab32, ytab32, ztab32, Y, X, z0, const_cast<uint8_t*>(len8), nullptr,
                     tile_off, out);
    static unsigned long long attr_done = 0;
    if (cfe_first_on_device(&attr_done)) {
        CFE_CHECK_CUDA(cudaFuncSetAttribute(k_emit_nodes_wide, cudaFuncAttributeMaxDynamicSharedMemorySize, NODEW_STAGE));
    }
    k_emit_nodes_wide<<<(unsigned)((n_nodes + 255) / 256), 256, NODEW_STAGE, stream>>>(P);
    CFE_CHECK_LAUNCH("k_emit_nodes_wide");
    return 0;
}

// ------------------------------------------------------------------------------------------
// Generic (array-driven) deck emission: mesh2voxelfe accepts ANY hexahedral meshio mesh
// (voxelFE.py:612-688 only touches points, cells, the first cell_data entry and the sets), e.g. one
// read back from VTK or a vol2ugrid mesh whose sets were edited.  Same text, plain kernels:
// lengths -> 256-item tile totals -> cfe_scan_u32 -> byte-wise staging + coalesced flush.

struct GenNodeParams {
    const i64* ids;          // used node ids, ascending (np.unique(cells))
    i64 n;
    const char* slots;       // [n][3] 32-byte coordinate fields (cfe_format_coords_wide of points[ids])
    uint8_t* len8;
    u32* tile_tot;
    const u64* tile_off;
    char* out;
};

__global__ void __launch_bounds__(256) k_node_lengths_gen(GenNodeParams P)
{
    const i64 k = (i64)blockIdx.x * 256 + threadIdx.x;
    u32 len = 0;
    if (k < P.n) {
        const int nd = digits_u64((u64)P.ids[k]);
        const char* f = P.slots + 96 * (size_t)k;
        len = (u32)(nd < 10 ? 10 : nd) + 7u + (u32)(uint8_t)f[0] + (u32)(uint8_t)f[32] + (u32)(uint8_t)f[64];
        P.len8[k] = (uint8_t)len;
    }
    __shared__ u32 s_scan[33];
    u32 tot;
    block_exclusive_scan<u32>(len, s_scan, tot);
    if (threadIdx.x == 0) P.tile_tot[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(256) k_emit_nodes_gen(GenNodeParams P)
{
    extern __shared__ __align__(16) char stage[];
    __shared__ u32 s_scan[33];
    const i64 k = (i64)blockIdx.x * 256 + threadIdx.x;
    const u32 len = k < P.n ? (u32)P.len8[k] : 0u;
    u32 tot;
    const u32 excl = block_exclusive_scan<u32>(len, s_scan, tot);
    char* dst = P.out + P.tile_off[blockIdx.x];
    const u32 pad = (u32)(reinterpret_cast<uintptr_t>(dst) & 15u);
    if (k < P.n) {
        const u64 id = (u64)P.ids[k];
        char* p = stage + pad + excl;
        const int nd = digits_u64(id), w = nd < 10 ? 10 : nd;
        for (int q = 0; q < w - nd; ++q) *p++ = ' ';
        p += nd;
        put_u64_back(p, id, nd);
        const char* f = P.slots + 96 * (size_t)k;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            *p++ = ',';
            *p++ = ' ';
            const int fl = (int)(uint8_t)f[32 * a];
            for (int q = 0; q < fl; ++q) *p++ = f[32 * a + 1 + q];
        }
        *p = '\n';
    }
    __syncthreads();
    flush_stage(stage, pad, tot, dst);
}

extern "C" int cfe_gen_node_lengths(const int64_t* ids, int64_t n, const char* slots, uint8_t* len8,
                                    uint32_t* tile_tot, cudaStream_t stream)
{
    if (n <= 0) return 0;
    GenNodeParams P = {reinterpret_cast<const i64*>(ids), n, slots, len8, tile_tot, nullptr, nullptr};
    k_node_lengths_gen<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(P);
    CFE_CHECK_LAUNCH("k_node_lengths_gen");
    return 0;
}

extern "C" int cfe_gen_emit_nodes(const int64_t* ids, int64_t n, const char* slots, const uint8_t* len8,
                                  const uint64_t* tile_off, char* out, cudaStream_t stream)
{
    if (n <= 0) return 0;
    GenNodeParams P = {reinterpret_cast<const i64*>(ids), n, slots, const_cast<uint8_t*>(len8), nullptr,
                       reinterpret_cast<const u64*>(tile_off), out};
    static unsigned long long attr_done = 0;
    if (cfe_first_on_device(&attr_done)) {
        CFE_CHECK_CUDA(cudaFuncSetAttribute(k_emit_nodes_gen, cudaFuncAttributeMaxDynamicSharedMemorySize, NODEW_STAGE));
    }
    k_emit_nodes_gen<<<(unsigned)((n + 255) / 256), 256, NODEW_STAGE, stream>>>(P);
    CFE_CHECK_LAUNCH("k_emit_nodes_gen");
    return 0;
}

struct GenElemParams {
    const i64* cells;        // [n_cells][8] node ids
    const i64* order;        // [n] element index of every emitted line (grouped by grey value, ascending)
    const int* group;        // [n] group of every emitted line
    const i64* gfirst;       // [n_groups] emission index of each group's first line
    const int* hdr_off;      // [n_groups] '*ELEMENT ...' header of each group in `lit`
    const int* hdr_len;
    const char* lit;
    i64 n;
    unsigned short* len16;
    u32* tile_tot;
    const u64* tile_off;
    char* out;
};

#define GEN_ELEM_STAGE (256 * 200 + 32)

__global__ void __launch_bounds__(256) k_elem_lengths_gen(GenElemParams P)
{
    const i64 t = (i64)blockIdx.x * 256 + threadIdx.x;
    u32 len = 0;
    if (t < P.n) {
        const i64 e = P.order[t];
        len = 9u + (u32)digits_u64((u64)(e + 1));
#pragma unroll
        for (int q = 0; q < 8; ++q) len += (u32)digits_u64((u64)P.cells[8 * e + q]);
        const int g = P.group[t];
        if (P.gfirst[g] == t) len += (u32)P.hdr_len[g];
        P.len16[t] = (unsigned short)len;
    }
    __shared__ u32 s_scan[33];
    u32 tot;
    block_exclusive_scan<u32>(len, s_scan, tot);
    if (threadIdx.x == 0) P.tile_tot[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(256) k_emit_elements_gen(GenElemParams P)
{
    extern __shared__ __align__(16) char stage[];
    __shared__ u32 s_scan[33];
    const i64 t = (i64)blockIdx.x * 256 + threadIdx.x;
    const u32 len = t < P.n ? (u32)P.len16[t] : 0u;
    u32 tot;
    const u32 excl = block_exclusive_scan<u32>(len, s_scan, tot);
    char* dst = P.out + P.tile_off[blockIdx.x];
    const u32 pad = (u32)(reinterpret_cast<uintptr_t>(dst) & 15u);
    const bool staged = pad + tot <= (u32)GEN_ELEM_STAGE;
    if (t < P.n) {
        char* p = staged ? stage + pad + excl : dst + excl;
        const i64 e = P.order[t];
        const int g = P.group[t];
        if (P.gfirst[g] == t) {
            const char* h = P.lit + P.hdr_off[g];
            const int hl = P.hdr_len[g];
            for (int q = 0; q < hl; ++q) *p++ = h[q];
        }
#pragma unroll 1
        for (int q = 0; q < 9; ++q) {
            const u64 v = q == 0 ? (u64)(e + 1) : (u64)P.cells[8 * e + q - 1];
            const int nd = digits_u64(v);
            p += nd;
            put_u64_back(p, v, nd);
            *p++ = q == 8 ? '\n' : ',';
        }
    }
    __syncthreads();
    if (staged) flush_stage(stage, pad, tot, dst);
}

extern "C" int cfe_gen_elem_lengths(const int64_t* cells, const int64_t* order, const int* group,
                                    const int64_t* gfirst, const int* hdr_off, const int* hdr_len, int64_t n,
                                    uint16_t* len16, uint32_t* tile_tot, cudaStream_t stream)
{
    if (n <= 0) return 0;
    GenElemParams P = {reinterpret_cast<const i64*>(cells), reinterpret_cast<const i64*>(order), group,
                       reinterpret_cast<const i64*>(gfirst), hdr_off, hdr_len, nullptr, n, len16, tile_tot, nullptr,
                       nullptr};
    k_elem_lengths_gen<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(P);
    CFE_CHECK_LAUNCH("k_elem_lengths_gen");
    return 0;
}

extern "C" int cfe_gen_emit_elements(const int64_t* cells, const int64_t* order, const int* group,
                                     const int64_t* gfirst, const int* hdr_off, const int* hdr_len, const char* lit,
                                     int64_t n, const uint16_t* len16, const uint64_t* tile_off, char* out,
                                     cudaStream_t stream)
{
    if (n <= 0) return 0;
    GenElemParams P = {reinterpret_cast<const i64*>(cells), reinterpret_cast<const i64*>(order), group,
                       reinterpret_cast<const i64*>(gfirst), hdr_off, hdr_len, lit, n,
                       const_cast<unsigned short*>(len16), nullptr, reinterpret_cast<const u64*>(tile_off), out};
    static unsigned long long attr_done = 0;
    if (cfe_first_on_device(&attr_done)) {
        CFE_CHECK_CUDA(cudaFuncSetAttribute(k_emit_elements_gen, cudaFuncAttributeMaxDynamicSharedMemorySize, GEN_ELEM_STAGE));
    }
    k_emit_elements_gen<<<(unsigned)((n + 255) / 256), 256, GEN_ELEM_STAGE, stream>>>(P);
    CFE_CHECK_LAUNCH("k_emit_elements_gen");
    return 0;
}
