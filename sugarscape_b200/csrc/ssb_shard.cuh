// ssb_shard.cuh -- x-slab sharding of one simulation over the GPUs of a box.
//
// All grid and agent arrays are "stitched" (ssb_fabric.cuh): one address space, rank p's slab of
// the map and rank p's slot range live in rank p's HBM.  A rank processes the agents on ITS list
// (initially: the agents standing on its slab; migration at the end of a step keeps it that way)
// and the cells of its slab; whatever a transaction touches across the slab boundary -- candidate
// cells, prey, neighbours, reservation marks -- is reached directly over NVLink.  What remains is
// synchronisation: the commit rounds of the agent step run in lockstep on all ranks, so every
// grid-wide barrier that separates phases touching shared state becomes a box-wide barrier.
//
// Box-wide barrier / small all-gather: every rank owns a control page in the stitched `ctrl`
// array.  Message slot [parity][src] of rank p's page is written only by rank src: 7 payload words,
// then the barrier epoch with release semantics at system scope.  A rank waits until all slots of
// its own page carry the current epoch.  Slots alternate with the epoch's parity: a rank cannot be
// two barriers ahead of a peer, so a slot is never overwritten before it was read.  Spins are
// bounded (SHARD_TIMEOUT_NS): on expiry the rank raises its abort flag, every later barrier falls
// through and the host reports the failure -- a lost peer must not hang the GPU.
#pragma once
#include <cooperative_groups.h>
#include <stdint.h>

namespace ssb {

namespace cg = cooperative_groups;

constexpr int SHARD_MSG_WORDS = 8;                 // 7 payload + epoch
constexpr int SHARD_PAYLOAD = 7;
constexpr int SHARD_CTRL_EPOCH = 0, SHARD_CTRL_ABORT = 1, SHARD_CTRL_MSGS = 64;
constexpr unsigned long long SHARD_TIMEOUT_NS = 20ull * 1000ull * 1000ull * 1000ull;

struct ShardCtx {
    int32_t rank, world;
    long long cell_lo, cell_hi;                    // cells of this rank's slab
    long long slot_lo, slot_hi;                    // slots this rank allocates from
    long long mark_lo, mark_hi;                    // mark words this rank clears
    long long cell_end[SSB_MAX_RANKS];             // rank r owns the cells below cell_end[r] (and not below cell_end[r - 1])
    unsigned long long *ctrl;                      // stitched: rank p's page at ctrl + p * ctrl_stride
    long long ctrl_stride;                         // in 64-bit words
    unsigned long long *gather;                    // stitched scratch, rank p's part at gather + p * gather_stride
    long long gather_stride;                       // in 64-bit words
    unsigned long long *peer_vals;                 // local: [world][SHARD_PAYLOAD] of the last exchange
    unsigned long long *flags;                     // local: [0] = sum of payload word 0, [1] = ok
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long shard_now_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// One full warp of ONE block per rank.  out: this rank's 7 payload words (nullptr = zeros).
// Afterwards s.peer_vals holds every rank's payload, s.flags[0] the sum of word 0 over the ranks,
// s.flags[1] = 1 (or 0 after a timeout / abort).  Returns the ok flag on every lane.
__device__ __forceinline__ bool shard_exchange_warp(const ShardCtx &s, const unsigned long long *out) {
    const int lane = threadIdx.x & 31;
    unsigned long long *mine = s.ctrl + (long long)s.rank * s.ctrl_stride;
    const unsigned long long epoch = ld_relaxed_sys(mine + SHARD_CTRL_EPOCH) + 1ull;
    bool ok = ld_relaxed_sys(mine + SHARD_CTRL_ABORT) == 0ull;
    unsigned long long word0 = 0;
    if (ok && lane < s.world) {
        const int parity = (int)(epoch & 1ull);
        unsigned long long *dst = s.ctrl + (long long)lane * s.ctrl_stride + SHARD_CTRL_MSGS +
                                  (parity * SSB_MAX_RANKS + s.rank) * SHARD_MSG_WORDS;
        for (int w = 0; w < SHARD_PAYLOAD; w++) st_relaxed_sys(dst + w, out ? out[w] : 0ull);
        st_release_sys(dst + SHARD_PAYLOAD, epoch);
        const unsigned long long *src = mine + SHARD_CTRL_MSGS + (parity * SSB_MAX_RANKS + lane) * SHARD_MSG_WORDS;
        const unsigned long long t0 = shard_now_ns();
        while (ld_acquire_sys(src + SHARD_PAYLOAD) != epoch) {
            if (shard_now_ns() - t0 > SHARD_TIMEOUT_NS) { ok = false; break; }
            __nanosleep(64);
        }
        if (ok) {
            for (int w = 0; w < SHARD_PAYLOAD; w++) s.peer_vals[lane * SHARD_PAYLOAD + w] = ld_relaxed_sys(src + w);
            word0 = s.peer_vals[lane * SHARD_PAYLOAD];
        }
    }
    ok = __all_sync(0xffffffffu, ok);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) word0 += __shfl_xor_sync(0xffffffffu, word0, o);
    if (lane == 0) {
        if (!ok) st_relaxed_sys(mine + SHARD_CTRL_ABORT, 1ull);
        st_relaxed_sys(mine + SHARD_CTRL_EPOCH, epoch);
        s.flags[0] = word0;
        s.flags[1] = ok ? 1ull : 0ull;
    }
    __threadfence();
    return ok;
}

// Box-wide barrier inside a cooperative kernel; *payload (read after the grid barrier) is summed over the ranks.
// Every thread of the grid must call it.  Returns false once the box-wide barrier is broken.
__device__ __forceinline__ bool shard_grid_sync(cg::grid_group &grid, const ShardCtx &s, const int32_t *payload,
                                                unsigned long long *sum_out) {
    if (s.world <= 1) {
        grid.sync();
        if (sum_out) *sum_out = payload ? (unsigned long long)*(const volatile int32_t *)payload : 0ull;
        return true;
    }
    // Everything this block wrote (also into the peers' HBM) is performed system-wide before the
    // block arrives at the grid barrier: block barrier, then ONE system-scope fence, which is
    // cumulative over the writes it observed through the barrier (the pattern of NCCL's
    // primitives: workers -> bar.sync -> the posting thread fences and stores the flag).
    __syncthreads();
    if (threadIdx.x == 0) __threadfence_system();
    grid.sync();
    if (blockIdx.x == 0 && threadIdx.x < 32) {
        unsigned long long out[SHARD_PAYLOAD] = {payload ? (unsigned long long)*(const volatile int32_t *)payload : 0ull, 0, 0, 0, 0, 0, 0};
        shard_exchange_warp(s, out);
    }
    grid.sync();
    if (sum_out) *sum_out = ((volatile unsigned long long *)s.flags)[0];
    return ((volatile unsigned long long *)s.flags)[1] != 0ull;
}

// Between kernels: orders everything this rank's stream has done before against everything the
// peers' streams do afterwards, and all-gathers 7 words (device pointer, may be null).
__global__ void k_shard_barrier(ShardCtx s, const unsigned long long *out, long long *overflow) {
    __threadfence_system();
    if (!shard_exchange_warp(s, out) && threadIdx.x == 0 && overflow) *overflow = 7;
}

}  // namespace ssb
