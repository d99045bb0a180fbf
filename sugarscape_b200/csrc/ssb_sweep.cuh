// ssb_sweep.cuh -- the kernels of the ordered sweep (ssb_sweep_core.cuh has the algorithm and its proof sketch).
//
// Per step, on one stream:
//   cudaMemsetAsync   done bitmap (one bit per cell), gates (one 64-bit word per list entry)
//   k_sweep_prepare   cooperative, three phases separated by grid barriers: histogram of the priority buckets,
//                     exclusive scan, scatter.  LEAN: walks the SLOTS (coalesced reads of the ABI's arrays), writes
//                     record A by slot and record B at the agent's list index; otherwise walks the agent list and
//                     writes `entries`.  Either way every agent leaves its descriptor on its cell and flags its tile.
//   k_sweep_build     one CTA per populated 64 x 64 tile (dynamic tile cursor): descriptors of tile + halo staged
//                     in shared memory as 4-byte words, one occupancy bit-plane per slab (shared-memory atomics), one
//                     thread per agent scans the bit-planes of its two near slabs and stores its gate (only agents
//                     WITH near predecessors store anything).
//   k_sweep_clear     zeroes the descriptors of the populated tiles and their flags (the arrays are clean again).
//   k_sweep_lean /    persistent cooperative grid (co-residency is what the progress argument needs): a lane draws a
//   k_sweep_generic   ticket for the next list entry whenever it is free; the entry runs once the watermark has
//                     passed the slabs before its near slabs and the done bits of its near predecessors are set.
//                     Lanes never wait inside the transaction: a lane whose agent is not ready simply retries in the
//                     next iteration of its warp, so lanes of one warp can depend on each other.
//                     Memory ordering: writes -> red.release.gpu (done bit) -> warp barrier -> atom.release.gpu (bucket
//                     count; the thread that completes a bucket advances the watermark behind a fence.sc);
//                     consumer: ld.relaxed.gpu (watermark, done words) -> reads past L1 (lean) / acquire fence (generic).
//   k_lean_unpack     LEAN: mutable fields of record A back to the ABI's arrays, slot order.
#pragma once
#include "ssb_agent_step.cuh"
#include "ssb_sweep_core.cuh"

namespace ssb {

constexpr int SWEEP_MAX_KEY_BITS = 13;                 // bucket keys = priority bucket [<< 3 | range on the lean path]
constexpr int SWEEP_MAX_KEYS = 1 << SWEEP_MAX_KEY_BITS;
constexpr int SWEEP_LEAN_SUB_BITS = 3;                 // lean path: entries of one priority bucket are grouped by range, so that the
                                                       // 32 agents a warp works on have the same number of candidate cells
constexpr int SWEEP_PREPARE_THREADS = 256;
constexpr int SWEEP_BUILD_THREADS = 512;
constexpr int SWEEP_OVERFLOW_STALL = 10;    // counters[SSB_C_OVERFLOW] code: no agent of a warp became ready for SWEEP_STALL_NS
constexpr int SWEEP_OVERFLOW_LIVE = 11;     // ... the live slots differ from the agent list (lean path: every live slot must be a list entry)
constexpr int SWEEP_OVERFLOW_BUCKET = 12;   // ... a priority bucket is longer than the number of concurrent transactions
constexpr unsigned long long SWEEP_STALL_NS = 4000000000ull;
constexpr int LEAN_THREADS = 256;
constexpr int SWEEP_STRIPES = 32;                      // ticket counters of the sweep (sweep_claim)
constexpr int SWEEP_CTL_WORDS = 16 + SWEEP_STRIPES * 16;            // ctl[16 + 16 k]: ticket counter of stripe k (own 128-byte line)

struct SweepCtx {
    unsigned long long *desc;        // [n_cells] descriptor of the agent that starts the step on the cell (0: none)
    int4 *entries;                   // [capacity] generic path, by list index: {slot, cell, priority, r | k << 8}
    unsigned long long *gate;        // [capacity] near predecessors by list index (ssb_sweep_core.cuh); 0: none
    uint32_t *preds;                 // [capacity * SWEEP_COLS] mask lines of the entries whose gate is SWEEP_GATE_LINE
    unsigned long long *done;        // [sweep_done_words] the done bitmap (ssb_sweep_core.cuh)
    int32_t *bucket_count, *bucket_start, *bucket_fill;      // [SWEEP_MAX_KEYS + 1]
    int32_t *bucket_fin;             // [SWEEP_MAX_KEYS + 1] finished agents per priority bucket
    int32_t *tile_pop;               // [tiles_x * tiles_y] != 0: agents stand on the tile
    int32_t *tile_need;              // [tiles_x * tiles_y] == stamp: the tile or a neighbour is populated (lean path: its cells are packed)
    int32_t stamp;                   // changes with every agent step
    unsigned long long *ctl;         // [0] tile cursor of the build, [1] abort flag, [2] live agents seen by prepare, [3] overflowing gates, [4..7] diagnostics,
                                     // [8] the watermark (own 64-byte half line), [16 + 16 k] tickets
    const int8_t *tables;            // conflict tables `any` then `pair` (sweep_fill_tables)
    unsigned long long *diag;        // optional (SSB_SWEEP_DIAG=1): [0] warp iterations, [1] entries checked, [2] entries run, [3] entries behind the watermark
    int32_t bucket_bits;             // priority buckets = 1 << bucket_bits
    int32_t sub_bits;                // bucket keys per priority bucket = 1 << sub_bits (lean path: by range)
    int32_t slab_bits;               // slabs = 1 << min(bucket_bits, slab_bits) (sweep_slab_bits)
    int32_t lines_only;              // few slabs (most predecessors are near ones): every agent with predecessors gets a mask line
    // x-slab sharding (ssb_shard.cuh; world 1: slot0 = 0, all tiles, no peers): this rank walks ITS slots and tiles; the
    // descriptors, the done bitmap and the packed cells are stitched arrays, the ranks publish their watermarks in their
    // control pages and every rank waits on the smallest
    int32_t slot0;                   // first slot of this rank
    int32_t tile_x_lo, tile_x_hi;    // tile columns of this rank's slab
    int32_t pack_all;                // pack / unpack every tile of the slab (agents of a peer may walk in)
    int32_t world, rank;
    unsigned long long *wm_pages;    // stitched control pages: rank r's watermark at wm_pages[r * wm_stride]
    long long wm_stride;
    int32_t n_keys;                  // bucket keys (= buckets << sub bits); bucket_start[n_keys] = live agents
    int32_t rb, kb, halo;            // largest range / dilation of the step, halo = flow_halo(rb, kb)
    int32_t tiles_x, tiles_y;
    int32_t concurrent;              // T: transactions in flight in the sweep kernel
    uint64_t key;
};

// sharded: a rank also waits for its peers' kernels to START (their hosts enqueue them on their own clocks): the limit of the
// box-wide barriers (SHARD_TIMEOUT_NS, 20 s) applies
__device__ __forceinline__ unsigned long long sweep_stall_limit(const SweepCtx &w) { return w.world > 1 ? SHARD_TIMEOUT_NS : SWEEP_STALL_NS; }

__device__ __forceinline__ unsigned long long sweep_ld_relaxed(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// the same under a predicate (all ones when off): straight-line code, so that a batch of these loads is in flight together
__device__ __forceinline__ unsigned long long sweep_ld_relaxed_if(const unsigned long long *p, uint32_t on) {
    unsigned long long v;
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\tmov.u64 %0, 0xffffffffffffffff;\n\t@p ld.relaxed.gpu.global.u64 %0, [%1];\n\t}"
                 : "=l"(v) : "l"(p), "r"(on) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long sweep_ld_relaxed_sys_if(const unsigned long long *p, uint32_t on) {
    unsigned long long v;
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\tmov.u64 %0, 0xffffffffffffffff;\n\t@p ld.relaxed.sys.global.u64 %0, [%1];\n\t}"
                 : "=l"(v) : "l"(p), "r"(on) : "memory");
    return v;
}
// 32-bit halves of the done words (little endian: bit b of word k is bit b & 31 of half 2 k + (b >> 5)), under a predicate
template <bool SYS>
__device__ __forceinline__ uint32_t sweep_ld32_relaxed_if(const uint32_t *p, uint32_t on) {
    uint32_t v;
    if (SYS) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\tmov.u32 %0, 0xffffffff;\n\t@p ld.relaxed.sys.global.u32 %0, [%1];\n\t}" : "=r"(v) : "l"(p), "r"(on) : "memory");
    else asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\tmov.u32 %0, 0xffffffff;\n\t@p ld.relaxed.gpu.global.u32 %0, [%1];\n\t}" : "=r"(v) : "l"(p), "r"(on) : "memory");
    return v;
}
__device__ __forceinline__ void sweep_fence() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
// release: MEMBAR.ALL.GPU (my writes are in L2) + the reduction that sets my done bit, WITHOUT the L1 invalidation of a
// full fence; RELEASE = false: after an explicit fence
// SYS: sharded -- a peer GPU reads the bit over NVLink, and the transaction may have written into the peer's memory
template <bool RELEASE, bool SYS = false>
__device__ __forceinline__ void sweep_set_done(const SweepCtx &w, int x, int y, int wpc) {
    unsigned long long *p = w.done + sweep_done_word(x, y, wpc);
    if (SYS) asm volatile("red.release.sys.global.or.b64 [%0], %1;" ::"l"(p), "l"(1ull << (y & 63)) : "memory");
    else if (RELEASE) asm volatile("red.release.gpu.global.or.b64 [%0], %1;" ::"l"(p), "l"(1ull << (y & 63)) : "memory");
    else asm volatile("red.relaxed.gpu.global.or.b64 [%0], %1;" ::"l"(p), "l"(1ull << (y & 63)) : "memory");
}

// ---------------------------------------------------------------------------------------------
// prepare
// ---------------------------------------------------------------------------------------------
__host__ __device__ constexpr int sweep_prepare_smem_bytes(int n_keys) { return n_keys * 4; }

template <bool LEAN>
__global__ void __launch_bounds__(SWEEP_PREPARE_THREADS) k_sweep_prepare(DevCtx c, SweepCtx w, LeanCtx L) {
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) unsigned char prepare_smem[];
    __shared__ int32_t sh_scan[SWEEP_PREPARE_THREADS];
    constexpr int SUB = LEAN ? SWEEP_LEAN_SUB_BITS : 0;
    const int E = 1 << w.bucket_bits, NK = E << SUB, shift = c.p.id_bits - w.bucket_bits;
    int32_t *sh_count = reinterpret_cast<int32_t *>(prepare_smem);
    const int tid = threadIdx.x, H = c.p.height;
    const long long n = LEAN ? c.a.counters[SSB_C_N_SLOTS] - w.slot0 : c.a.counters[SSB_C_N_LIST];      // (LEAN: this rank's slots)
    long long chunk = (n + gridDim.x - 1) / gridDim.x;
    chunk = (chunk + SWEEP_PREPARE_THREADS - 1) / SWEEP_PREPARE_THREADS * SWEEP_PREPARE_THREADS;
    const long long lo = min((long long)blockIdx.x * chunk, n), hi = min(lo + chunk, n);
    for (int e = tid; e < NK; e += blockDim.x) sh_count[e] = 0;
    if (blockIdx.x == 0) {
        for (int e = tid; e <= NK; e += blockDim.x) { w.bucket_count[e] = 0; w.bucket_fill[e] = 0; }
        for (int e = tid; e <= E; e += blockDim.x) w.bucket_fin[e] = 0;
        for (int k = tid; k < SWEEP_STRIPES; k += blockDim.x) w.ctl[16 + 16 * k] = 0;
        if (tid == 0) { w.ctl[0] = 0; w.ctl[1] = 0; w.ctl[2] = 0; w.ctl[3] = 0; w.ctl[8] = 0; c.a.counters[SSB_C_N_BORN] = 0; c.a.counters[SSB_C_ROUNDS] = 1; }
    }
    grid.sync();
    // phase A: histogram of the bucket keys (the per-block work assignment is repeated exactly in phase B)
    int mine = 0;
    for (long long i = lo + tid; i < hi; i += blockDim.x) {
        const int s = LEAN ? w.slot0 + (int)i : c.a.order[i];
        const bool live = c.a.pos[s] >= 0 && (!LEAN || c.a.cod[s] == SSB_ALIVE);
        if (!live) continue;
        const uint32_t pr = priority_of(w.key, (uint32_t)c.a.id[s], c.p.id_bits);
        const int key = (int)((pr >> shift) << SUB) | (LEAN ? lean_range(c, s) : 0);
        atomicAdd(&sh_count[key], 1);
        mine++;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
    if ((tid & 31) == 0 && mine) atomicAdd(&w.ctl[2], (unsigned long long)mine);
    __syncthreads();
    for (int e = tid; e < NK; e += blockDim.x) if (sh_count[e]) atomicAdd(&w.bucket_count[e], sh_count[e]);
    grid.sync();
    // exclusive scan of the key histogram (block 0), sanity checks
    if (blockIdx.x == 0) {
        const int per = (NK + SWEEP_PREPARE_THREADS - 1) / SWEEP_PREPARE_THREADS;
        int sum = 0;
        for (int j = 0; j < per; j++) { const int e = tid * per + j; if (e < NK) sum += w.bucket_count[e]; }
        sh_scan[tid] = sum;
        __syncthreads();
        if (tid == 0) { int run = 0; for (int j = 0; j < SWEEP_PREPARE_THREADS; j++) { const int v = sh_scan[j]; sh_scan[j] = run; run += v; } w.bucket_start[NK] = run; }
        __syncthreads();
        int run = sh_scan[tid];
        for (int j = 0; j < per; j++) { const int e = tid * per + j; if (e < NK) { w.bucket_start[e] = run; run += w.bucket_count[e]; } }
        // a thread meets entries `concurrent` apart: they must lie in different PRIORITY buckets (ssb_sweep_core.cuh)
        int longest = 0;
        for (int b = tid; b < E; b += blockDim.x) {
            int v = 0;
            for (int j = 0; j < (1 << SUB); j++) v += w.bucket_count[(b << SUB) + j];
            longest = max(longest, v);
        }
        if (longest > w.concurrent) atomicExch((unsigned long long *)&c.a.counters[SSB_C_OVERFLOW], (unsigned long long)SWEEP_OVERFLOW_BUCKET);
        // the watermark starts behind the leading EMPTY buckets (nobody would ever complete them)
        __syncthreads();
        if (tid == 0) {
            int wb = 0;
            while (wb < E && w.bucket_start[(wb + 1) << SUB] == w.bucket_start[wb << SUB]) wb++;
            w.ctl[8] = w.world > 1 ? 0ull : (unsigned long long)wb;
            w.ctl[9] = (unsigned long long)wb;                       // this rank's own watermark
            if (w.world > 1) st_release_sys(w.wm_pages + (long long)w.rank * w.wm_stride, (unsigned long long)wb);
        }
        if (LEAN && tid == 0 && (long long)w.ctl[2] != c.a.counters[SSB_C_N_LIST])
            atomicExch((unsigned long long *)&c.a.counters[SSB_C_OVERFLOW], (unsigned long long)SWEEP_OVERFLOW_LIVE);
    }
    grid.sync();
    // phase B: one range reservation per block and key, then scatter
    for (int e = tid; e < NK; e += blockDim.x)          // (the counter becomes the cursor of this block's range of the key)
        sh_count[e] = sh_count[e] ? w.bucket_start[e] + atomicAdd(&w.bucket_fill[e], sh_count[e]) : 0;
    __syncthreads();
    for (long long i = lo + tid; i < hi; i += blockDim.x) {
        const int s = LEAN ? w.slot0 + (int)i : c.a.order[i];
        bool live;
        if (LEAN) live = lean_pack_a(c, L, s);
        else live = c.a.pos[s] >= 0;
        if (!live) continue;
        const int pos = c.a.pos[s];
        const uint32_t pr = priority_of(w.key, (uint32_t)c.a.id[s], c.p.id_bits);
        const int r = lean_range(c, s);
        const int key = (int)((pr >> shift) << SUB) | (LEAN ? r : 0);
        const int index = atomicAdd(&sh_count[key], 1);
        const int k = LEAN ? 0 : footprint_dilation(c, s);
        w.desc[pos] = sweep_desc((uint32_t)index, sweep_lo(pr, r, k));
        const int x = pos / H, y = pos - x * H;
        w.tile_pop[(x / FLOW_TX) * w.tiles_y + y / FLOW_TY] = 1;      // (a flag: plain stores of the same value)
        if (LEAN) lean_pack_b(c, L, s, index, r);
        else w.entries[index] = make_int4(s, pos, (int)pr, r | (k << 8));
    }
}

// ---------------------------------------------------------------------------------------------
// build
// ---------------------------------------------------------------------------------------------
__host__ __device__ inline int sweep_build_smem_bytes(int rb, int kb, int halo, int n_slabs) {
    return (FLOW_TX + 2 * halo) * (FLOW_TY + 2 * halo) * 4 + n_slabs * (FLOW_TX + 2 * halo) * 16 + FLOW_TX * FLOW_TY * 2 + sweep_tables_bytes(rb, kb, halo) + 16;
}

__global__ void __launch_bounds__(SWEEP_BUILD_THREADS, 3) k_sweep_build(DevCtx c, SweepCtx w) {
    extern __shared__ __align__(16) unsigned char sweep_smem[];
    __shared__ int sh_tile, sh_n_agents, sh_class[16];
    const int R = w.halo, sw = FLOW_TX + 2 * R, sh = FLOW_TY + 2 * R;
    const int n_slabs = sweep_n_slabs(w.bucket_bits, w.slab_bits);
    const int slab_prio_shift = c.p.id_bits - w.bucket_bits + sweep_slab_shift(w.bucket_bits, w.slab_bits);
    unsigned long long *smask = reinterpret_cast<unsigned long long *>(sweep_smem);
    uint32_t *cell = reinterpret_cast<uint32_t *>(smask + (size_t)n_slabs * sw * 2);
    uint16_t *agents = reinterpret_cast<uint16_t *>(cell + sw * sh);
    int8_t *tables = reinterpret_cast<int8_t *>(agents + FLOW_TX * FLOW_TY);
    const int tid = threadIdx.x;
    const int W = c.p.width, H = c.p.height;
    const int n_tiles = w.tiles_x * w.tiles_y;
    for (int i = tid; i < sweep_tables_bytes(w.rb, w.kb, R); i += SWEEP_BUILD_THREADS) tables[i] = w.tables[i];
    SweepTile tile;
    tile.cell = cell; tile.smask = smask; tile.sw = sw; tile.sh = sh; tile.halo = R; tile.slab_prio_shift = slab_prio_shift;
    tile.any = tables; tile.pair = tables + sweep_any_size(w.rb, w.kb, R);
    for (;;) {
        __syncthreads();                                        // the previous tile's shared state is dead
        if (tid == 0) { sh_tile = (int)atomicAdd(&w.ctl[0], 1ull); sh_n_agents = 0; }
        if (tid < 16) sh_class[tid] = 0;
        __syncthreads();
        const int t = sh_tile;
        if (t >= n_tiles) break;
        if (w.tile_pop[t] == 0) continue;
        const int tx = t / w.tiles_y, ty = t - tx * w.tiles_y;
        const int x0 = tx * FLOW_TX - R, y0 = ty * FLOW_TY - R;
        for (int i = tid; i < n_slabs * sw * 2; i += SWEEP_BUILD_THREADS) smask[i] = 0ull;
        __syncthreads();
        // stage the low descriptor words (a column of the staged tile is a contiguous run of the x-major grid, up to the
        // torus seam) and set every agent's bit in the plane of its slab
        const bool inside = x0 >= 0 && y0 >= 0 && x0 + sw <= W && y0 + sh <= H;
        const unsigned long long *src = w.desc + (size_t)(inside ? x0 : 0) * H + (inside ? y0 : 0);
#pragma unroll 4
        for (int job = tid; job < sw * sh; job += SWEEP_BUILD_THREADS) {
            const int col = job / sh, row = job - col * sh;
            const uint32_t lo = inside ? (uint32_t)__ldg(src + (size_t)col * H + row)
                                       : (uint32_t)__ldg(w.desc + (size_t)wrapi(x0 + col, W) * H + wrapi(y0 + row, H));
            cell[job] = lo;
            if (lo != 0u) atomicOr(smask + ((size_t)(sweep_prio(lo) >> slab_prio_shift) * sw + col) * 2 + (row >> 6), 1ull << (row & 63));
        }
        __syncthreads();
        for (int word = tid; word < sw * 2; word += SWEEP_BUILD_THREADS) sweep_fold_planes(smask, n_slabs, sw, word);    // plane g: slabs g - 1 and g
        __syncthreads();
        // the tile's own agents (interior cells that lie on the map), grouped by range: the 32 agents a warp scans have
        // conflict regions of the same size
        for (int idx = tid; idx < FLOW_TX * FLOW_TY; idx += SWEEP_BUILD_THREADS) {
            const int ix = idx / FLOW_TY, iy = idx - ix * FLOW_TY;
            if (tx * FLOW_TX + ix >= W || ty * FLOW_TY + iy >= H) continue;
            const uint32_t lo = cell[(R + ix) * sh + R + iy];
            if (lo != 0u) atomicAdd(&sh_class[sweep_r(lo)], 1);
        }
        __syncthreads();
        if (tid == 0) {
            int run = 0;
            for (int k = 15; k >= 0; k--) { const int v = sh_class[k]; sh_class[k] = run; run += v; }     // longest scans first
            sh_n_agents = run;
        }
        __syncthreads();
        for (int idx = tid; idx < FLOW_TX * FLOW_TY; idx += SWEEP_BUILD_THREADS) {
            const int ix = idx / FLOW_TY, iy = idx - ix * FLOW_TY;
            if (tx * FLOW_TX + ix >= W || ty * FLOW_TY + iy >= H) continue;
            const int sc = (R + ix) * sh + R + iy;
            const uint32_t lo = cell[sc];
            if (lo != 0u) agents[atomicAdd(&sh_class[sweep_r(lo)], 1)] = (uint16_t)sc;
        }
        __syncthreads();
        const int n_agents = sh_n_agents;
        if (w.lines_only) {      // one pass: the column masks stream into the agent's line, 16 bytes at a time
            for (int a = tid; a < n_agents; a += SWEEP_BUILD_THREADS) {
                const int sc = agents[a], lx = sc / sh, ly = sc - lx * sh;
                const uint32_t gcell = (uint32_t)((tx * FLOW_TX + lx - R) * H + (ty * FLOW_TY + ly - R));
                const uint32_t index = (uint32_t)(__ldg(w.desc + gcell) >> 32);
                uint4 *out = reinterpret_cast<uint4 *>(w.preds + (size_t)index * SWEEP_COLS);
                uint4 acc = make_uint4(0u, 0u, 0u, 0u);
                uint32_t any_pred = 0u;
                sweep_near_masks(tile, lx, ly, cell[sc], w.rb, w.kb, [&](int col, uint32_t m) {
                    const int u = col & 3;
                    any_pred |= m;
                    if (u == 0) acc.x = m; else if (u == 1) acc.y = m; else if (u == 2) acc.z = m; else acc.w = m;
                    if (u == 3) { out[col >> 2] = acc; acc = make_uint4(0u, 0u, 0u, 0u); }
                });
                out[(2 * R + 1) >> 2] = acc;                    // the last, partial group (2 R + 1 columns; the words behind them are not read)
                if (any_pred) w.gate[index] = SWEEP_GATE_LINE;
            }
            continue;
        }
        for (int a = tid; a < n_agents; a += SWEEP_BUILD_THREADS) {
            const int sc = agents[a], lx = sc / sh, ly = sc - lx * sh;
            unsigned long long gate = 0ull;
            int n_near = 0;
            sweep_near_preds(tile, lx, ly, cell[sc], w.rb, w.kb, [&](int dx, int dy) {
                if (n_near < SWEEP_GATE_PREDS) gate |= sweep_gate_code(dx, dy, R) << (10 * n_near);
                n_near++;
            });
            if (n_near == 0) continue;                           // (the gates were zeroed before the launch)
            const uint32_t gcell = (uint32_t)((tx * FLOW_TX + lx - R) * H + (ty * FLOW_TY + ly - R));
            const uint32_t index = (uint32_t)(__ldg(w.desc + gcell) >> 32);
            if (n_near <= SWEEP_GATE_PREDS) { w.gate[index] = gate; continue; }
            // more near predecessors than a gate holds (rare): a line of column masks
            uint32_t *line = w.preds + (size_t)index * SWEEP_COLS;
            for (int k = 0; k < SWEEP_COLS; k++) line[k] = 0u;
            sweep_near_preds(tile, lx, ly, cell[sc], w.rb, w.kb, [&](int dx, int dy) { line[dx + R] |= 1u << (dy + R); });
            w.gate[index] = SWEEP_GATE_LINE;
            atomicAdd(&w.ctl[3], 1ull);
        }
    }
}

__global__ void __launch_bounds__(256) k_sweep_clear(DevCtx c, SweepCtx w) {
    const int W = c.p.width, H = c.p.height, n_tiles = w.tiles_x * w.tiles_y;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        if (w.tile_pop[t] == 0) continue;                       // (uniform: read before anyone resets it)
        __syncthreads();
        const int tx = t / w.tiles_y, ty = t - tx * w.tiles_y;
        if (threadIdx.x == 0) w.tile_pop[t] = 0;
        if (threadIdx.x < 9) {                                  // agents reach (and move) into the neighbouring tiles
            const int nx = wrapi(tx + (int)threadIdx.x / 3 - 1, w.tiles_x), ny = wrapi(ty + (int)threadIdx.x % 3 - 1, w.tiles_y);
            w.tile_need[nx * w.tiles_y + ny] = w.stamp;
        }
        for (int idx = threadIdx.x; idx < FLOW_TX * FLOW_TY; idx += blockDim.x) {
            const int ix = idx / FLOW_TY, iy = idx - ix * FLOW_TY;
            const int x = tx * FLOW_TX + ix, y = ty * FLOW_TY + iy;
            if (x < W && y < H) w.desc[(size_t)x * H + y] = 0ull;
        }
    }
}

// Lean path: {cell.agent, cell.sugar} of the needed tiles between the ABI's x-major arrays and the tiled cell array
// (ssb_sweep_core.cuh, lean_tiled).  One thread per 2 x 2 sector; a 64 x 64 tile is sixteen 2 KB runs of the tiled array.
template <bool PACK>
__global__ void __launch_bounds__(256) k_lean_cells(DevCtx c, SweepCtx w, LeanCtx L) {
    const int W = c.p.width, H = c.p.height, n_tiles = w.tiles_x * w.tiles_y;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int tx = t / w.tiles_y, ty = t - tx * w.tiles_y;
        if (w.pack_all ? (tx < w.tile_x_lo || tx >= w.tile_x_hi) : (w.tile_need[t] != w.stamp)) continue;
        static_assert(FLOW_TX == 64 && FLOW_TY == 64, "sector enumeration below");
        for (int sct = threadIdx.x; sct < (FLOW_TX / 2) * (FLOW_TY / 2); sct += blockDim.x) {
            // consecutive sectors of the tiled array: 4 x 4 tile (xq, yq), then the 2 x 2 sector (sx, sy) inside it
            const int sy = sct & 1, sx = (sct >> 1) & 1, yq = (sct >> 2) & 15, xq = sct >> 6;
            const int x = tx * FLOW_TX + 4 * xq + 2 * sx, y = ty * FLOW_TY + 4 * yq + 2 * sy;
            if (x >= W || y >= H) continue;
            int2 *p = L.pcell + lean_tiled(x, y, L.ty4);         // (x, y), (x, y + 1), (x + 1, y), (x + 1, y + 1)
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int cx = x + (u >> 1), cy = y + (u & 1);
                if (cx >= W || cy >= H) continue;
                const size_t cell = (size_t)cx * H + cy;
                if (PACK) p[u] = make_int2(c.g.occ[cell], c.g.sugar[cell]);
                else { const int2 v = p[u]; c.g.occ[cell] = v.x; c.g.sugar[cell] = v.y; }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// sweep
// ---------------------------------------------------------------------------------------------
// May list entry i run?  The caller has compared the watermark; this is the individual part: the done bits of the near
// predecessors in the entry's gate.  Straight-line code: the (at most six) done words are loaded under predicates and
// are in flight together.  pos: the entry's start cell.
// ... of a mask line: straight-line code like the gate -- the eight 16-byte loads of the line are in flight together, then
// the done words of its columns in batches of eight (the 2 halo + 1 rows of a column lie in one or two 64-bit words)
template <bool SYS>
__device__ __noinline__ bool sweep_line_ready_seam(const SweepCtx &w, int i, int x, int y, int W, int H) {
    const uint32_t *line = w.preds + (size_t)i * SWEEP_COLS;
    uint32_t masks[SWEEP_COLS];
    for (int k = 0; k <= 2 * w.halo; k++) masks[k] = __ldg(line + k);
    return sweep_line_done(masks, [&](size_t wd) { return SYS ? ld_relaxed_sys(w.done + wd) : sweep_ld_relaxed(w.done + wd); }, x, y, W, H, w.halo);
}
template <bool SYS>
__device__ __forceinline__ bool sweep_line_ready(const SweepCtx &w, int i, int x, int y, int W, int H) {
    const int R = w.halo, span = 2 * R + 1, y0 = y - R;
    if (y0 < 0 || y0 + span > H) return sweep_line_ready_seam<SYS>(w, i, x, y, W, H);      // the window crosses the torus seam in y
    const uint4 *pm = reinterpret_cast<const uint4 *>(w.preds + (size_t)i * SWEEP_COLS);
    uint4 v[8];
#pragma unroll
    for (int g = 0; g < 8; g++) v[g] = __ldg(pm + g);
    const int hpc = 2 * sweep_done_wpc(H), shift = y0 & 31;        // 32-bit halves per column; the window lies in one or two of them
    const bool two = shift + span > 32;
    const uint32_t *col0 = reinterpret_cast<const uint32_t *>(w.done) + (y0 >> 5);
    bool ok = true;
#pragma unroll
    for (int g = 0; g < 8; g++) {
        if (4 * g >= span) break;
        const uint32_t m[4] = {v[g].x, v[g].y, v[g].z, v[g].w};
        uint32_t lo[4], hi[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int col = 4 * g + u;
            const uint32_t on = col < span ? m[u] : 0u;
            int cx = x - R + col;
            if (cx < 0) cx += W; else if (cx >= W) cx -= W;
            const uint32_t *p = col0 + (size_t)cx * hpc;
            lo[u] = sweep_ld32_relaxed_if<SYS>(p, on);
            hi[u] = sweep_ld32_relaxed_if<SYS>(p + 1, two ? on : 0u);
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const uint32_t on = 4 * g + u < span ? m[u] : 0u;
            ok &= (__funnelshift_r(lo[u], hi[u], shift) & on) == on;
        }
    }
    return ok;
}
template <bool SYS = false>
__device__ __forceinline__ bool sweep_gate_ready(const SweepCtx &w, unsigned long long gate, int i, int pos, int W, int H) {
    if (gate == 0ull) return true;
    const int x = pos / H, y = pos - x * H;
    if (gate == SWEEP_GATE_LINE) return sweep_line_ready<SYS>(w, i, x, y, W, H);
    const int wpc = sweep_done_wpc(H), R = w.halo;
    unsigned long long dw[SWEEP_GATE_PREDS];
    int bit[SWEEP_GATE_PREDS];
#pragma unroll
    for (int k = 0; k < SWEEP_GATE_PREDS; k++) {
        const int code = (int)(gate >> (10 * k)) & 1023;
        int cx, cy;
        sweep_offset_cell(x, y, (code & 31) - 1 - R, (code >> 5) - R, W, H, &cx, &cy);       // (an empty slot decodes to a cell that is not loaded)
        const unsigned long long *p = w.done + sweep_done_word(cx, cy, wpc);
        dw[k] = SYS ? sweep_ld_relaxed_sys_if(p, (uint32_t)(code & 31)) : sweep_ld_relaxed_if(p, (uint32_t)(code & 31));
        bit[k] = cy & 63;
    }
    bool ok = true;
#pragma unroll
    for (int k = 0; k < SWEEP_GATE_PREDS; k++) ok &= (bool)((dw[k] >> bit[k]) & 1ull);
    return ok;
}

// The watermark (ctl[8]): the number of leading priority buckets whose agents have all finished.  The lanes of a warp
// that finished an agent in this iteration count themselves in their buckets (one fire-and-forget reduction per
// distinct bucket, released); ONE warp of the grid (sweep_watermark_warp) does nothing but walk the watermark over the
// complete buckets.  Everybody else keeps a copy of the watermark in a register and re-reads it only when one of its
// agents fails the comparison (the watermark only grows: a stale copy is merely conservative).
// all lanes of the warp; ran: this lane finished an agent of `bucket` in this iteration (its done bit is set: its writes
// are released)
template <bool SYS = false>
__device__ __forceinline__ void sweep_count_finished(const SweepCtx &w, bool ran, int bucket) {
    const unsigned ranm = __ballot_sync(0xffffffffu, ran);
    if (!ran) return;
    __syncwarp(ranm);
    const unsigned peers = __match_any_sync(ranm, bucket);
    if ((int)(threadIdx.x & 31) != __ffs(peers) - 1) return;
    if (SYS) asm volatile("red.release.sys.global.add.s32 [%0], %1;" ::"l"(w.bucket_fin + bucket), "r"(__popc(peers)) : "memory");
    else asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(w.bucket_fin + bucket), "r"(__popc(peers)) : "memory");
}
// Sharded (w.world > 1): the walk yields THIS rank's watermark (ctl[9], published in its control page); ctl[8] -- what the
// agents compare with -- is the smallest watermark of the box, read from the peers' pages over NVLink.
__device__ __noinline__ void sweep_watermark_warp(const DevCtx &c, const SweepCtx &w) {
    if ((threadIdx.x & 31) != 0) return;
    const int E = 1 << w.bucket_bits, SUB = w.sub_bits;
    const bool sharded = w.world > 1;
    int mine = (int)sweep_ld_relaxed(w.ctl + 9), box = (int)sweep_ld_relaxed(w.ctl + 8);
    unsigned long long last = global_ns();
    while (box < E) {
        bool moved = false;
        while (mine < E) {
            const int size = w.bucket_start[(mine + 1) << SUB] - w.bucket_start[mine << SUB];
            int fin;
            asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(fin) : "l"(w.bucket_fin + mine) : "memory");
            if (fin != size) break;
            mine++;
            moved = true;
        }
        int all = mine;
        if (sharded) {
            if (moved) st_release_sys(w.wm_pages + (long long)w.rank * w.wm_stride, (unsigned long long)mine);
            for (int r = 0; r < w.world; r++)
                if (r != w.rank) { const int v = (int)ld_acquire_sys(w.wm_pages + (long long)r * w.wm_stride); if (v < all) all = v; }
        }
        if (all > box) {
            box = all;
            asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(w.ctl + 8), "l"((unsigned long long)box) : "memory");
            last = global_ns();
            continue;
        }
        __nanosleep(100);
        if (*(const volatile unsigned long long *)(w.ctl + 1) != 0ull) return;
        if (global_ns() - last > sweep_stall_limit(w)) {
            atomicExch((unsigned long long *)&c.a.counters[SSB_C_OVERFLOW], (unsigned long long)SWEEP_OVERFLOW_STALL);
            atomicExch(w.ctl + 1, 1ull);
            return;
        }
    }
}

// Entries are handed out in list order by ticket counters, one entry per lane (group) whenever it has finished its
// previous one: the agents in flight are always the most recently started ones -- a thin slice of the priority order,
// however long individual transactions take.  ONE counter would serialise every warp of the grid on one L2 address
// (measured: the kernel ran at exactly the rate of that atomic), so the list is dealt out in blocks of 32 entries to
// SWEEP_STRIPES counters (block b belongs to stripe b mod SWEEP_STRIPES, warp w draws from stripe w mod SWEEP_STRIPES).
// Progress: every lane meets its entries in ascending priority bucket; the unfinished agent with the smallest priority
// is either held by a lane (and runs) or is the next entry of its stripe, and fewer than a bucket's worth of that
// stripe's lanes can be waiting for it (k_sweep_prepare checks that a bucket is shorter than the grid).
__device__ __forceinline__ void sweep_claim(const SweepCtx &w, int n, bool leader_lane, int &i, bool &exhausted) {
    const bool need = leader_lane && i < 0 && !exhausted;
    const unsigned needm = __ballot_sync(0xffffffffu, need);
    if (needm == 0u) return;
    const int lane = threadIdx.x & 31, first = __ffs(needm) - 1;
    const int stripe = (int)((blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) % SWEEP_STRIPES);
    unsigned long long base = 0;
    if (lane == first) base = atomicAdd(w.ctl + 16 + 16 * stripe, (unsigned long long)__popc(needm));
    base = __shfl_sync(0xffffffffu, base, first);
    if (need) {
        const long long t = (long long)base + __popc(needm & ((1u << lane) - 1u));
        const long long mine = ((t >> 5) * SWEEP_STRIPES + stripe) * 32 + (t & 31);
        if (mine < n) i = (int)mine;
        else exhausted = true;
    }
}

// a warp without a ready agent backs off; one that starves for SWEEP_STALL_NS fails loudly instead of hanging the GPU
__device__ __forceinline__ bool sweep_idle(const DevCtx &c, const SweepCtx &w, unsigned long long &idle_since) {
    __nanosleep(64);
    const unsigned long long now = global_ns();
    if (idle_since == 0) idle_since = now;
    else if (now - idle_since > sweep_stall_limit(w)) {
        if ((threadIdx.x & 31) == 0) {
            atomicExch((unsigned long long *)&c.a.counters[SSB_C_OVERFLOW], (unsigned long long)SWEEP_OVERFLOW_STALL);
            atomicExch(w.ctl + 1, 1ull);
        }
        return true;
    }
    return *(const volatile unsigned long long *)(w.ctl + 1) != 0ull;
}

// per thread: [maxc + 1] 16-bit sugar words, then (not in the plain instantiation) [maxc] 32-bit occupants
__host__ __device__ constexpr int lean_smem_bytes(int maxc, int threads, bool slim) { return threads * (((maxc + 2) & ~1) * 2 + (slim ? 0 : maxc * 4)); }

// The lean transaction reads everything mutable past L1 (lean_ld), so the consumer needs no acquire fence (whose
// CCTL.IVALL would invalidate the whole L1 once per agent); the producer's release rides on the reduction that sets
// the done bit.  MODE: see lean_transaction.
#ifndef SSB_LEAN_BLOCKS
#define SSB_LEAN_BLOCKS 5
#endif
#ifndef SSB_LEAN_PREFETCH
#define SSB_LEAN_PREFETCH 0      /* measured on S2: requesting the cross ahead of the gate check costs 7 ms (L2 request rate) */
#endif
// SYS: the sharded sweep (x-slabs over the GPUs of a box): done bits, bucket counts and watermarks at system scope
template <int MODE, bool SYS = false>
__global__ void __launch_bounds__(LEAN_THREADS, MODE == 0 ? SSB_LEAN_BLOCKS : 3) k_sweep_lean(DevCtx c, SweepCtx w, LeanCtx L) {
    static_assert(!SYS || !(MODE & LEAN_M_FIGHT), "the sharded sweep keeps the records A private: no instantiation that reads a neighbour's record");
    constexpr bool PREFETCH = SSB_LEAN_PREFETCH != 0;
    extern __shared__ __align__(16) unsigned char lean_smem[];
    const int T = blockDim.x, tid = threadIdx.x;
    int16_t *sug_st = reinterpret_cast<int16_t *>(lean_smem);
    int32_t *occ_st = reinterpret_cast<int32_t *>(sug_st + ((L.maxc + 2) & ~1) * T);
    const LeanStage st = {occ_st + tid, sug_st + tid, T};
    const int n = w.bucket_start[w.n_keys];
    const int W = c.p.width, H = c.p.height, wpc = sweep_done_wpc(H);
    const int prio_shift = c.p.id_bits - w.bucket_bits, slab_shift = sweep_slab_shift(w.bucket_bits, w.slab_bits);
    if (blockIdx.x == 0 && tid < 32) { sweep_watermark_warp(c, w); return; }
    int i = -1;
    bool exhausted = false;
    unsigned long long idle_since = 0;
    int watermark = 0;
    for (;;) {
        const int held = i;
        sweep_claim(w, n, true, i, exhausted);
        if (!__any_sync(0xffffffffu, i >= 0)) break;
        bool ran = false, ready = false;
        int bucket = 0, need = 0;
        LeanB B;
        unsigned long long gate = 0ull;
        if (i >= 0) {
            B = L.b[i];                                          // (read-only: in flight together with the gate)
            gate = __ldg(w.gate + i);
            if (PREFETCH && i != held) lean_prefetch_agent(L, B.slot, B.pos, (int)((B.idf >> LEANB_RANGE_SHIFT) & 15u), W, H);
            bucket = (int)(priority_of(w.key, B.idf & LEANB_ID_MASK, c.p.id_bits) >> prio_shift);
            need = sweep_watermark_needed(bucket, slab_shift);
        }
        if (__any_sync(0xffffffffu, i >= 0 && watermark < need)) watermark = (int)sweep_ld_relaxed(w.ctl + 8);
        const bool passed = i >= 0 && watermark >= need;
        LeanA A0;
        if (passed) {
            if (!(MODE & LEAN_M_FIGHT)) A0 = lean_load_a(L.a + B.slot);     // (nobody else writes it: in flight with the done words)
            ready = sweep_gate_ready<SYS>(w, gate, i, B.pos, W, H);
        }
        const unsigned readym = __ballot_sync(0xffffffffu, ready);      // (also re-converges the warp after the check)
        if (ready) {
            lean_transaction<MODE>(c, L, B, w.key, st, readym, (MODE & LEAN_M_FIGHT) ? nullptr : &A0);
            const int x = B.pos / H;
            sweep_set_done<true, SYS>(w, x, B.pos - x * H, wpc); // release: my writes happen before my done bit
            i = -1;
            ran = true;
        }
        sweep_count_finished<SYS>(w, ran, bucket);
        if (w.diag) {
            const unsigned held = __ballot_sync(0xffffffffu, i >= 0 || ran), did = __ballot_sync(0xffffffffu, ran);
            const unsigned late = __ballot_sync(0xffffffffu, i >= 0 && !passed);
            if ((tid & 31) == 0) { atomicAdd(w.diag, 1ull); atomicAdd(w.diag + 1, (unsigned long long)__popc(held)); atomicAdd(w.diag + 2, (unsigned long long)__popc(did)); atomicAdd(w.diag + 3, (unsigned long long)__popc(late)); }
        }
        if (__any_sync(0xffffffffu, ran)) idle_since = 0;
        else if (sweep_idle(c, w, idle_since)) break;
    }
}

// the generic transaction (ssb_rules.cuh, GL lanes per agent) under the same scheduler
__global__ void __launch_bounds__(AGENT_THREADS, 2) k_sweep_generic(DevCtx c, SweepCtx w) {
    extern __shared__ __align__(8) unsigned char scratch_mem[];
    const WarpScratch ws = make_warp_scratch(scratch_mem + (threadIdx.x / GL) * warp_scratch_bytes(MAXC_SCALABLE), MAXC_SCALABLE);
    const int n = w.bucket_start[w.n_keys];
    const int W = c.p.width, H = c.p.height, lane = Group<GL>::lane(), wpc = sweep_done_wpc(H);
    const int prio_shift = c.p.id_bits - w.bucket_bits, slab_shift = sweep_slab_shift(w.bucket_bits, w.slab_bits);
    if (blockIdx.x == 0 && threadIdx.x < 32) { sweep_watermark_warp(c, w); return; }
    int i = -1;
    bool exhausted = false;
    unsigned long long idle_since = 0;
    int watermark = 0;
    for (;;) {
        sweep_claim(w, n, lane == 0, i, exhausted);             // lane 0 of every group claims, the group follows
        i = Group<GL>::bcast(i, 0);
        if (!__any_sync(0xffffffffu, i >= 0)) break;
        bool ran = false;
        int bucket = 0, need = 0;
        int4 en = make_int4(0, 0, 0, 0);
        if (i >= 0) {                                            // (every lane of the group reads the same words)
            en = w.entries[i];
            bucket = (int)((uint32_t)en.z >> prio_shift);
            need = sweep_watermark_needed(bucket, slab_shift);
        }
        if (__any_sync(0xffffffffu, i >= 0 && watermark < need)) watermark = (int)sweep_ld_relaxed(w.ctl + 8);
        if (i >= 0) {
            if (watermark >= need && sweep_gate_ready(w, __ldg(w.gate + i), i, en.y, W, H)) {
                sweep_fence();
                const int s = en.x;
                AgentRec rec;
                rec.load(c, s);
                RngStream rng{w.key, (uint32_t)rec.id, 0u, 0u, nullptr};
                if (transaction_move_phase<GL, 4>(c, s, rec, rng, ws, MAXC_SCALABLE)) {
                    if (lane == 0) transaction_interact_phase<RngStream, false>(c, s, rec, rng);
                }
                Group<GL>::sync();
                sweep_fence();
                const int x = en.y / H;
                if (lane == 0) sweep_set_done<false>(w, x, en.y - x * H, wpc);
                i = -1;
                ran = true;
            }
        }
        sweep_count_finished(w, ran && lane == 0, bucket);
        if (__any_sync(0xffffffffu, ran)) idle_since = 0;
        else if (sweep_idle(c, w, idle_since)) break;
    }
}

__global__ void __launch_bounds__(256) k_lean_unpack(DevCtx c, LeanCtx L, int slot0) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthreads = gridDim.x * blockDim.x;
    const int n_slots = (int)c.a.counters[SSB_C_N_SLOTS];
    for (int s = slot0 + tid; s < n_slots; s += nthreads) lean_unpack_slot(c, L, s);
}

}  // namespace ssb
