// ssb_sort.cuh -- keys-only LSD radix sort of 64-bit keys (8-bit digits), used to order agent
// wealths for the Lorenz / Gini reduction (reference sugarscape.py:913-934 sorts the wealth list).
// Non-negative doubles sort correctly as their raw bit patterns.
//
// Per pass: k_radix_hist (per-tile digit histograms, digit-major) -> k_radix_scan (one block per
// digit row: exclusive scan of the row + row total) -> k_radix_scatter (stable: warp match-any ranking in striped
// order, warps of a tile combined by a per-digit prefix).  n lives in device memory so that no
// host synchronisation is needed between the compaction and the sort.
#pragma once
#include <stdint.h>

namespace ssb {

constexpr int RADIX_THREADS = 256;
constexpr int RADIX_ITEMS = 16;
constexpr int RADIX_TILE = RADIX_THREADS * RADIX_ITEMS;   // 4096 keys per block
constexpr int RADIX_WARPS = RADIX_THREADS / 32;

// varying[0] = OR of all keys, varying[1] = AND of all keys: a digit whose bits are all equal in
// OR and AND is the same in every key, so its pass degenerates to a copy.
__device__ __forceinline__ bool digit_is_constant(const unsigned long long *varying, int shift) {
    return (((varying[0] ^ varying[1]) >> shift) & 255ull) == 0ull;
}

// needed (may be null): the sort only runs if *needed != 0 (ssb_kernels.cuh: the Lorenz histogram covers integral wealths)
__global__ void __launch_bounds__(RADIX_THREADS) k_radix_hist(const uint64_t *keys, const int64_t *n_ptr, int shift,
                                                              uint32_t *hist, const unsigned long long *varying, const unsigned long long *needed = nullptr) {
    if (needed && *needed == 0ull) return;
    if (digit_is_constant(varying, shift)) return;
    const long long n = *n_ptr;
    const int n_tiles = (int)((n + RADIX_TILE - 1) / RADIX_TILE);
    __shared__ uint32_t h[256];
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {      // the grid is sized for the SMs, not for n
        h[threadIdx.x] = 0;
        __syncthreads();
        const long long tile0 = (long long)tile * RADIX_TILE;
#pragma unroll
        for (int k = 0; k < RADIX_ITEMS; k++) {
            const long long i = tile0 + k * RADIX_THREADS + threadIdx.x;
            if (i < n) atomicAdd(&h[(keys[i] >> shift) & 255u], 1u);
        }
        __syncthreads();
        hist[(long long)threadIdx.x * n_tiles + tile] = h[threadIdx.x];
    }
}

// Scan of the digit-major histogram table hist[256][n_tiles], one block per digit row: the row is
// scanned in place (exclusive) and its total goes to row_total[digit]; the scatter kernel adds the
// exclusive prefix of the row totals.  Launch with 256 blocks.
__global__ void __launch_bounds__(RADIX_THREADS) k_radix_scan(uint32_t *hist, const int64_t *n_ptr, int shift,
                                                              const unsigned long long *varying, uint32_t *row_total, const unsigned long long *needed = nullptr) {
    if (needed && *needed == 0ull) return;
    if (digit_is_constant(varying, shift)) return;
    const long long n = *n_ptr;
    const int n_tiles = (int)((n + RADIX_TILE - 1) / RADIX_TILE);
    uint32_t *row = hist + (long long)blockIdx.x * n_tiles;
    const int chunk = (n_tiles + RADIX_THREADS - 1) / RADIX_THREADS;
    const int lo = min((int)threadIdx.x * chunk, n_tiles), hi = min(lo + chunk, n_tiles);
    uint32_t sum = 0;
    for (int i = lo; i < hi; i++) sum += row[i];
    __shared__ uint32_t warp_sums[RADIX_WARPS];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += u;
    }
    if (lane == 31) warp_sums[w] = inc;
    __syncthreads();
    uint32_t base = 0;
    for (int j = 0; j < w; j++) base += warp_sums[j];
    uint32_t run = base + inc - sum;
    for (int i = lo; i < hi; i++) { const uint32_t v = row[i]; row[i] = run; run += v; }
    if (threadIdx.x == RADIX_THREADS - 1) row_total[blockIdx.x] = run;
}

__global__ void __launch_bounds__(RADIX_THREADS) k_radix_scatter(const uint64_t *keys, uint64_t *out, const int64_t *n_ptr,
                                                                 int shift, const uint32_t *hist,
                                                                 const unsigned long long *varying, const uint32_t *row_total, const unsigned long long *needed = nullptr) {
    if (needed && *needed == 0ull) return;
    const long long n = *n_ptr;
    const int n_tiles = (int)((n + RADIX_TILE - 1) / RADIX_TILE);
    if (digit_is_constant(varying, shift)) {       // nothing to reorder: keep the ping-pong going
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const long long tile0 = (long long)tile * RADIX_TILE;
            for (int k = 0; k < RADIX_ITEMS; k++) {
                const long long i = tile0 + k * RADIX_THREADS + threadIdx.x;
                if (i < n) out[i] = keys[i];
            }
        }
        return;
    }
    __shared__ uint32_t cnt[RADIX_WARPS][256];
    __shared__ uint32_t digit_base[256];
    __shared__ uint32_t wsum[RADIX_WARPS];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    {   // exclusive prefix of the row totals: where the keys of each digit start
        const uint32_t mine = row_total[threadIdx.x];
        uint32_t inc = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += u;
        }
        if (lane == 31) wsum[w] = inc;
        __syncthreads();
        uint32_t b = 0;
        for (int j = 0; j < w; j++) b += wsum[j];
        digit_base[threadIdx.x] = b + inc - mine;
    }
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        __syncthreads();
        for (int i = threadIdx.x; i < RADIX_WARPS * 256; i += RADIX_THREADS) (&cnt[0][0])[i] = 0;
        __syncthreads();
        const long long warp0 = (long long)tile * RADIX_TILE + (long long)w * (32 * RADIX_ITEMS);
        uint64_t key[RADIX_ITEMS];
        uint16_t rank[RADIX_ITEMS];
#pragma unroll
        for (int r = 0; r < RADIX_ITEMS; r++) {
            const long long i = warp0 + r * 32 + lane;
            const bool valid = i < n;
            key[r] = valid ? keys[i] : 0ull;
            const uint32_t d = (uint32_t)(key[r] >> shift) & 255u;
            const unsigned vmask = __ballot_sync(0xffffffffu, valid);
            uint32_t pre = 0, peers = 0;
            if (valid) {
                peers = __match_any_sync(vmask, d);
                pre = cnt[w][d];
            }
            __syncwarp();
            if (valid && lane == __ffs(peers) - 1) cnt[w][d] = pre + __popc(peers);
            __syncwarp();
            rank[r] = (uint16_t)(pre + __popc(peers & ((1u << lane) - 1u)));
        }
        __syncthreads();
        {   // per digit: exclusive prefix over the warps of this tile, on top of the global offset
            const int d = threadIdx.x;
            uint32_t running = digit_base[d] + hist[(long long)d * n_tiles + tile];
#pragma unroll
            for (int j = 0; j < RADIX_WARPS; j++) {
                const uint32_t t = cnt[j][d];
                cnt[j][d] = running;
                running += t;
            }
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < RADIX_ITEMS; r++) {
            const long long i = warp0 + r * 32 + lane;
            if (i < n) out[cnt[w][(uint32_t)(key[r] >> shift) & 255u] + rank[r]] = key[r];
        }
    }
}

}  // namespace ssb
