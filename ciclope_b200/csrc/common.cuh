// ciclope_b200 -- shared device helpers (sm_100a).
//
// * decoupled look-back exclusive scan (single pass, dynamic tile ids, 64-bit status words)
// * block-wide exclusive scans
// * magic-number division for voxel/node index decomposition
// * decimal digit counting
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

typedef unsigned long long u64;
typedef unsigned int u32;
typedef long long i64;

#define CFE_FULL_MASK 0xffffffffu

// ------------------------------------------------------------------------------------------
// error plumbing (api.cu owns the storage)
void cfe_set_error(const char* fmt, ...);
void cfe_count_launch();
#define CFE_CHECK_LAUNCH(name)                                                        \
    do {                                                                              \
        cfe_count_launch();                                                           \
        cudaError_t e__ = cudaGetLastError();                                         \
        if (e__ != cudaSuccess) {                                                     \
            cfe_set_error("%s: kernel launch failed: %s", name, cudaGetErrorString(e__)); \
            return -2;                                                                \
        }                                                                             \
    } while (0)
#define CFE_CHECK_CUDA(call)                                                          \
    do {                                                                              \
        cudaError_t e__ = (call);                                                     \
        if (e__ != cudaSuccess) {                                                     \
            cfe_set_error("%s failed: %s", #call, cudaGetErrorString(e__));           \
            return -2;                                                                \
        }                                                                             \
    } while (0)

// Checked builds (profiles/tools/build_variant.sh <out.so> -DCFE_CHECKED): index assertions that trap the
// kernel (compute-sanitizer is not available on every pool; the asserts + the CPU-reference comparison of
// profiles/tools/sanitize_cases.py stand in for memcheck on the kernels that index shared memory by data).
#ifdef CFE_CHECKED
#define CFE_ASSERT(cond) do { if (!(cond)) { printf("CFE_ASSERT failed: %s (%s:%d)\n", #cond, __FILE__, __LINE__); __trap(); } } while (0)
#else
#define CFE_ASSERT(cond) do { } while (0)
#endif

// Bulk copies (cp.async.bulk, the "async proxy") read shared memory that ordinary stores (the generic proxy) wrote:
// EVERY writer fences its own stores before the barrier that precedes the copy, then one thread issues the copy
// (the order of the CUDA programming guide's bulk-copy examples: fence, barrier, elected thread).
__device__ __forceinline__ void fence_proxy_async_smem()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// Function attributes (opt-in dynamic shared memory) are per device: true the first time it is
// evaluated on the current device for this `mask` (one static mask per launch site; devices 0..63).
static inline bool cfe_first_on_device(unsigned long long* mask, int* sm_count = nullptr)
{
    int dev = 0;
    cudaGetDevice(&dev);
    if (sm_count) {
        static int sms[64];
        if (!sms[dev & 63]) {
            int n = 0;
            if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
            sms[dev & 63] = n;
        }
        *sm_count = sms[dev & 63];
    }
    const unsigned long long bit = 1ull << (dev & 63);
    if (*mask & bit) return false;
    *mask |= bit;
    return true;
}

// ------------------------------------------------------------------------------------------
// Decoupled look-back.  status[tile] packs a 2-bit flag and a 62-bit value in ONE 64-bit word,
// so a single volatile store publishes both (no fence needed between flag and value).
//   flag 0 = not ready, 1 = tile aggregate available, 2 = inclusive prefix available.
// The workspace layout is: [0] = dynamic tile counter (u64 slot, used as u32), [1..] = status.
// It must be zeroed (cudaMemsetAsync) before every launch that uses it.
#define SCAN_FLAG_A (1ull << 62)
#define SCAN_FLAG_P (2ull << 62)
#define SCAN_VAL_MASK ((1ull << 62) - 1ull)

__device__ __forceinline__ u32 scan_next_tile(u64* ws)
{
    __shared__ u32 s_tile;
    if (threadIdx.x == 0) s_tile = atomicAdd(reinterpret_cast<u32*>(ws), 1u);
    __syncthreads();
    return s_tile;
}

// Must be called by every lane of ONE warp.  Returns the exclusive prefix of `tile` (sum of the
// aggregates of tiles 0..tile-1) in every lane of that warp and publishes the inclusive prefix.
//
// The window is 32 * SCAN_LB_DEPTH predecessors per round trip to L2: with small, fast tiles the
// nearest published prefix is "number of tiles in flight" (hundreds) behind, and a 32-wide window
// makes the whole kernel wait on (tiles in flight / 32) serial L2 latencies per tile.
template <int SCAN_LB_DEPTH>
__device__ __forceinline__ u64 scan_lookback_d(u64* ws, u32 tile, u64 aggregate)
{
    volatile u64* status = ws + 1;
    const u32 lane = threadIdx.x & 31u;
    if (tile == 0) {
        if (lane == 0) status[0] = SCAN_FLAG_P | aggregate;
        return 0;
    }
    if (lane == 0) status[tile] = SCAN_FLAG_A | aggregate;
    u64 excl = 0;
    int base = (int)tile - 1;
    bool done = false;
    while (!done) {
        u64 v[SCAN_LB_DEPTH];
#pragma unroll
        for (int k = 0; k < SCAN_LB_DEPTH; ++k) {
            const int idx = base - (int)lane - 32 * k;
            v[k] = idx >= 0 ? status[idx] : SCAN_FLAG_P;  // virtual predecessors: inclusive prefix 0
        }
#pragma unroll
        for (int k = 0; k < SCAN_LB_DEPTH; ++k) {
            if (!done) {                                   // warp-uniform
                const int idx = base - (int)lane - 32 * k;
                while ((v[k] >> 62) == 0ull) v[k] = status[idx];   // not published yet: spin
                const u32 pmask = __ballot_sync(CFE_FULL_MASK, (v[k] >> 62) == 2ull);
                u64 val = v[k] & SCAN_VAL_MASK;
                if (pmask) {
                    const u32 first = (u32)__ffs((int)pmask) - 1u;  // nearest predecessor holding a prefix
                    if (lane > first) val = 0;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(CFE_FULL_MASK, val, o);
                excl += val;
                if (pmask) done = true;
            }
        }
        base -= 32 * SCAN_LB_DEPTH;
    }
    if (lane == 0) status[tile] = SCAN_FLAG_P | (excl + aggregate);
    return excl;
}

__device__ __forceinline__ u64 scan_lookback(u64* ws, u32 tile, u64 aggregate)
{
    return scan_lookback_d<1>(ws, tile, aggregate);
}

// ------------------------------------------------------------------------------------------
// Block-wide exclusive scan (blockDim.x a multiple of 32, <= 1024).  Returns the exclusive
// prefix of `v` over threadIdx.x and the block total in `total` (valid in all threads).
// `smem` must hold 33 elements of T.  Contains two __syncthreads().
template <typename T>
__device__ __forceinline__ T block_exclusive_scan(T v, T* smem, T& total)
{
    const u32 lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const u32 nwarps = (blockDim.x + 31u) >> 5;
    T incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        T t = __shfl_up_sync(CFE_FULL_MASK, incl, o);
        if (lane >= (u32)o) incl += t;
    }
    if (lane == 31) smem[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        T w = (lane < nwarps) ? smem[lane] : T(0);
        T wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            T t = __shfl_up_sync(CFE_FULL_MASK, wi, o);
            if (lane >= (u32)o) wi += t;
        }
        smem[lane] = wi - w;              // exclusive prefix of the warp sums
        if (lane == 31) smem[32] = wi;    // block total
    }
    __syncthreads();
    total = smem[32];
    return smem[warp] + incl - v;
}

// ------------------------------------------------------------------------------------------
// Division of a 32-bit numerator by a runtime-constant 32-bit divisor (d >= 1) with one
// mul.hi + shift; exact for all n < 2^32 (round-up method with a 33-bit magic folded in).
struct FastDiv {
    u32 d, mul, shr, add;
};

static inline FastDiv make_fastdiv(u32 d)
{
    FastDiv f;
    f.d = d;
    if (d == 1) { f.mul = 0; f.shr = 0; f.add = 1; return f; }  // handled specially
    u32 l = 0;
    while ((1ull << l) < d) ++l;  // l = ceil(log2 d), 1..32
    u64 m = ((1ull << 32) * ((1ull << l) - d)) / d + 1;  // fits 32 bits
    f.mul = (u32)m;
    f.shr = l - 1;
    f.add = 0;
    return f;
}

__device__ __forceinline__ u32 fd_div(u32 n, const FastDiv& f)
{
    if (f.add) return n;  // d == 1
    u32 t = __umulhi(n, f.mul);
    return (t + ((n - t) >> 1)) >> f.shr;
}

// ------------------------------------------------------------------------------------------
// decimal digits of v (v >= 0); 1 for v == 0.
__device__ __forceinline__ int digits_u64(u64 v)
{
    int d = 1;
    if (v >= 10000000000ull) { v /= 10000000000ull; d += 10; }
    u32 w = (u32)v;  // < 10^10 does not fit 32 bits in general: handle the 10-digit case first
    if (v >= 1000000000ull) return d + 9;
    if (w >= 100000000u) return d + 8;
    if (w >= 10000000u) return d + 7;
    if (w >= 1000000u) return d + 6;
    if (w >= 100000u) return d + 5;
    if (w >= 10000u) return d + 4;
    if (w >= 1000u) return d + 3;
    if (w >= 100u) return d + 2;
    if (w >= 10u) return d + 1;
    return d;
}

// decimal digits of a 32-bit value by binary search (4 compares)
__device__ __forceinline__ int digits_u32(u32 v)
{
    if (v >= 100000u) {
        if (v >= 10000000u) return v >= 1000000000u ? 10 : (v >= 100000000u ? 9 : 8);
        return v >= 1000000u ? 7 : 6;
    }
    if (v >= 100u) return v >= 10000u ? 5 : (v >= 1000u ? 4 : 3);
    return v >= 10u ? 2 : 1;
}

__device__ __forceinline__ u32 mask_lt(int b) { return (b >= 32) ? 0xffffffffu : ((1u << b) - 1u); }
__device__ __forceinline__ u32 mask_le(int b) { return (b >= 31) ? 0xffffffffu : ((2u << b) - 1u); }
