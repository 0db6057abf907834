// Small collectives of the z-slab path as ONE kernel over NVLink peer memory.
//
// The sharded path exchanges little data (counts, one packed bit plane, boundary vertices: bytes to a few
// hundred KB) but does it five to seven times per step, and every NCCL call costs a host enqueue plus a
// kernel whose duration is mostly rank skew.  Here every rank owns a symmetric buffer (allocated and
// exchanged once with torch.distributed._symmetric_memory; this file only sees raw peer pointers):
//
//   buffer : [slot][source rank][chunk_cap bytes]          signal pad : u32 [sig_base + source rank]
//
// k_peer_exchange, called by every rank with the same epoch:
//   1. stores its chunk into slot `epoch % slots` of every peer named in send_mask (remote stores over
//      NVLink; the own copy is a local store), then __threadfence_system();
//   2. the last CTA to finish publishes  signal[peer][slot][my rank] = epoch  on the peers it sent to
//      (release, system scope);
//   3. waits until the signal words of the ranks it receives from show `epoch` (acquire, system scope);
//   4. copies the received chunks to a private output [world][nbytes].
// An all-gather (everybody sends to everybody) synchronises all ranks: a rank can start exchange e + 1 only
// after every rank has entered exchange e, i.e. has consumed e - 1.  A shift (send to the rank above, wait for
// the rank below) only orders neighbours, so the caller runs at most SLOTS - 2 shifts in a row before it makes
// one of them a full synchronisation (sharded.PeerExchange): a slot is then never overwritten before its
// previous content has been gathered.
// A rank that waits longer than ~30 s traps (a peer died or the ranks disagree on the call sequence):
// the step fails loudly instead of hanging the GPU.
#include <cstdint>

#include "common.cuh"

struct PeerArgs {
    void* const* bufs;        // device array [world]: base of every rank's symmetric buffer
    u32* const* sigs;         // device array [world]: base of every rank's signal pad
    int rank, world;
    i64 slot_off;             // byte offset of the slot inside a buffer
    i64 chunk_cap;            // bytes reserved per source rank inside a slot
    u32 epoch, send_mask, sig_base;
    u32 wait_mask;            // bit p = wait for rank p's signal and gather its chunk (all-gather: every rank;
                              // shift up: the rank below only, and only the rank above is signalled)
};

__device__ __forceinline__ void st_release_sys(u32* p, u32 v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ u32 ld_acquire_sys(const u32* p)
{
    u32 v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// VOL: read the source with ld.volatile (no L1): it was written by other GPUs
template <bool VOL>
__device__ __forceinline__ void copy_bytes(uint8_t* dst, const uint8_t* src, i64 n, i64 tid, i64 nthreads)
{
    if ((((uintptr_t)dst | (uintptr_t)src) & 15u) == 0) {
        const i64 n16 = n >> 4;
        for (i64 i = tid; i < n16; i += nthreads) {
            const uint4* q = reinterpret_cast<const uint4*>(src) + i;
            reinterpret_cast<uint4*>(dst)[i] = VOL ? __ldcv(q) : *q;
        }
        for (i64 i = (n16 << 4) + tid; i < n; i += nthreads) dst[i] = VOL ? __ldcv(src + i) : src[i];
    } else {
        for (i64 i = tid; i < n; i += nthreads) dst[i] = VOL ? __ldcv(src + i) : src[i];
    }
}

__global__ void __launch_bounds__(256)
k_peer_exchange(const uint8_t* __restrict__ src, i64 nbytes, PeerArgs A, u32* counter, uint8_t* __restrict__ out)
{
    __shared__ u32 s_last;
    const i64 tid = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    const i64 nthreads = (i64)gridDim.x * blockDim.x;
    // 1. push
    for (int p = 0; p < A.world; ++p) {
        if (!((A.send_mask >> p) & 1u)) continue;
        uint8_t* dst = reinterpret_cast<uint8_t*>(A.bufs[p]) + A.slot_off + (i64)A.rank * A.chunk_cap;
        copy_bytes<false>(dst, src, nbytes, tid, nthreads);
    }
    __threadfence_system();
    __syncthreads();
    // 2. the last CTA publishes the signals, then releases the other CTAs of this grid
    if (threadIdx.x == 0) s_last = (atomicAdd(counter, 1u) == gridDim.x - 1u) ? 1u : 0u;
    __syncthreads();
    if (s_last) {
        __threadfence_system();
        if ((int)threadIdx.x < A.world && ((A.send_mask >> threadIdx.x) & 1u))
            st_release_sys(A.sigs[threadIdx.x] + A.sig_base + A.rank, A.epoch);
        __syncthreads();
        if (threadIdx.x == 0) {
            counter[0] = 0;
            __threadfence();
            atomicExch(counter + 1, A.epoch);
        }
    }
    // 3. wait for the peers this rank receives from (and for the publishing CTA of this grid)
    if ((int)threadIdx.x <= A.world) {
        const bool peer = (int)threadIdx.x < A.world;
        if (!peer || ((A.wait_mask >> threadIdx.x) & 1u)) {
            const u32* w = peer ? (A.sigs[A.rank] + A.sig_base + threadIdx.x) : (counter + 1);
            const long long t0 = clock64();
            while (ld_acquire_sys(w) != A.epoch) {
                if (clock64() - t0 > 60000000000LL) __trap();       // ~30 s at 2 GHz
                __nanosleep(200);
            }
        }
    }
    __syncthreads();
    // 4. gather the received chunks into the private output
    const uint8_t* slot = reinterpret_cast<const uint8_t*>(A.bufs[A.rank]) + A.slot_off;
    for (int p = 0; p < A.world; ++p)
        if ((A.wait_mask >> p) & 1u)
            copy_bytes<true>(out + (i64)p * nbytes, slot + (i64)p * A.chunk_cap, nbytes, tid, nthreads);
}

// All ranks call this with the same nbytes / epoch / slot geometry.  src: this rank's nbytes; out: [world][nbytes]
// (chunks of ranks outside wait_mask are undefined).  counter: 2 zero-initialised u32 of local device memory,
// private to this communicator.  send_mask: bit p = push my chunk to rank p and signal it; wait_mask: bit p =
// wait for rank p's signal and gather its chunk.  Every rank's wait_mask must be the transpose of the others'
// send_masks (all-gather: both full; shift up: send to rank + 1, wait for rank - 1).
extern "C" int cfe_peer_exchange(const void* src, int64_t nbytes, const void* peer_bufs_dev, const void* peer_sigs_dev,
                                 int rank, int world, int64_t slot_off, int64_t chunk_cap, uint32_t epoch,
                                 uint32_t send_mask, uint32_t wait_mask, uint32_t sig_base, uint32_t* counter,
                                 void* out, cudaStream_t stream)
{
    if (world <= 0 || world > 32 || rank < 0 || rank >= world) { cfe_set_error("cfe_peer_exchange: bad rank / world"); return -1; }
    if (nbytes < 0 || nbytes > chunk_cap || (chunk_cap & 15) || (slot_off & 15)) {
        cfe_set_error("cfe_peer_exchange: %lld bytes do not fit the %lld-byte chunk", (long long)nbytes, (long long)chunk_cap);
        return -1;
    }
    PeerArgs A;
    A.bufs = reinterpret_cast<void* const*>(peer_bufs_dev);
    A.sigs = reinterpret_cast<u32* const*>(peer_sigs_dev);
    A.rank = rank; A.world = world; A.slot_off = slot_off; A.chunk_cap = chunk_cap;
    A.epoch = epoch; A.send_mask = send_mask; A.wait_mask = wait_mask; A.sig_base = sig_base;
    // co-resident CTAs only (they wait for each other; the stream holds nothing else): 1 for small payloads,
    // up to 128 for the multi-MB planes of a 2048^2 interface
    int n_send = 0;
    for (int p = 0; p < world; ++p) n_send += (int)((send_mask >> p) & 1u) + (int)((wait_mask >> p) & 1u);
    i64 blocks = (nbytes * (i64)n_send / (16 * 256 * 16)) + 1;
    if (blocks > 128) blocks = 128;
    k_peer_exchange<<<(unsigned)blocks, 256, 0, stream>>>(reinterpret_cast<const uint8_t*>(src), nbytes, A, counter,
                                                         reinterpret_cast<uint8_t*>(out));
    CFE_CHECK_LAUNCH("k_peer_exchange");
    return 0;
}
