// ssb_sweep.cuh -- the kernels of the ordered sweep (ssb_sweep_core.cuh has the algorithm and its proof sketch).
//
// Per step, on one stream:
//   cudaMemsetAsync   done bitmap (one bit per cell)
//   k_sweep_prepare   cooperative, three phases separated by grid barriers: histogram of the priority buckets,
//                     exclusive scan, scatter.  LEAN: walks the SLOTS (coalesced reads of the ABI's arrays), writes
//                     record A by slot and record B at the agent's list index; otherwise walks the agent list and
//                     writes `entries`.  Either way every agent leaves its descriptor on its cell and flags its tile.
//   k_sweep_build     one CTA per populated 64 x 64 tile (dynamic tile cursor): descriptors of tile + halo staged
//                     in shared memory as 4-byte words, occupancy masks by ballot, one thread per agent computes
//                     the predecessor masks and stores them as one 128-byte line (word 31 = the start cell).
//   k_sweep_clear     zeroes the descriptors of the populated tiles and their flags (the arrays are clean again).
//   k_sweep_lean /    persistent cooperative grid (co-residency is what the progress argument needs): thread /
//   k_sweep_generic   lane group j takes entries j, j + T, ...; an entry runs once its predecessors' done bits
//                     are set.  Lanes never wait inside the transaction: a lane whose agent is not ready simply
//                     retries in the next iteration of its warp, so lanes of one warp can depend on each other.
//                     Memory ordering: writes -> fence.acq_rel.gpu -> red.or (done bit);  consumer: ld.relaxed.gpu
//                     (done words) -> fence.acq_rel.gpu -> reads.
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
constexpr int SWEEP_BUILD_THREADS = 256;
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
    uint32_t *preds;                 // [capacity * SWEEP_COLS] predecessor masks by list index; word 31 = start cell
    unsigned long long *done;        // [sweep_done_words] two offset copies of the done bitmap (ssb_sweep_core.cuh)
    int32_t *bucket_count, *bucket_start, *bucket_fill;      // [SWEEP_MAX_KEYS + 1]
    int32_t *tile_pop;               // [tiles_x * tiles_y] != 0: agents stand on the tile
    int32_t *tile_need;              // [tiles_x * tiles_y] == stamp: the tile or a neighbour is populated (lean path: its cells are packed)
    int32_t stamp;                   // changes with every agent step
    unsigned long long *ctl;         // [0] tile cursor of the build, [1] abort flag, [2] live agents seen by prepare, [4..6] diagnostics, [16 + 16 k] tickets
    const int8_t *tables;            // conflict tables `any` then `pair` (sweep_fill_tables)
    unsigned long long *diag;        // optional (SSB_SWEEP_DIAG=1): [0] warp iterations, [1] entries checked, [2] entries run
    int32_t bucket_bits;             // priority buckets = 1 << bucket_bits
    int32_t n_keys;                  // bucket keys (= buckets << sub bits); bucket_start[n_keys] = live agents
    int32_t rb, kb, halo;            // largest range / dilation of the step, halo = flow_halo(rb, kb)
    int32_t tiles_x, tiles_y;
    int32_t concurrent;              // T: transactions in flight in the sweep kernel
    uint64_t key;
};

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
__device__ __forceinline__ void sweep_fence() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
// release: MEMBAR.ALL.GPU (my writes are in L2) + the reductions that set my done bit in both copies, WITHOUT the L1
// invalidation of a full fence; RELEASE = false: after an explicit fence
template <bool RELEASE>
__device__ __forceinline__ void sweep_set_done(const SweepCtx &w, int x, int y, int W, int wpc) {
    bool first = RELEASE;
    sweep_done_set(x, y, W, wpc, [&](size_t wd, int bit) {
        if (first) asm volatile("red.release.gpu.global.or.b64 [%0], %1;" ::"l"(w.done + wd), "l"(1ull << bit) : "memory");
        else asm volatile("red.relaxed.gpu.global.or.b64 [%0], %1;" ::"l"(w.done + wd), "l"(1ull << bit) : "memory");
        first = false;
    });
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
    const long long n = LEAN ? c.a.counters[SSB_C_N_SLOTS] : c.a.counters[SSB_C_N_LIST];
    long long chunk = (n + gridDim.x - 1) / gridDim.x;
    chunk = (chunk + SWEEP_PREPARE_THREADS - 1) / SWEEP_PREPARE_THREADS * SWEEP_PREPARE_THREADS;
    const long long lo = min((long long)blockIdx.x * chunk, n), hi = min(lo + chunk, n);
    for (int e = tid; e < NK; e += blockDim.x) sh_count[e] = 0;
    if (blockIdx.x == 0) {
        for (int e = tid; e <= NK; e += blockDim.x) { w.bucket_count[e] = 0; w.bucket_fill[e] = 0; }
        for (int k = tid; k < SWEEP_STRIPES; k += blockDim.x) w.ctl[16 + 16 * k] = 0;
        if (tid == 0) { w.ctl[0] = 0; w.ctl[1] = 0; w.ctl[2] = 0; c.a.counters[SSB_C_N_BORN] = 0; c.a.counters[SSB_C_ROUNDS] = 1; }
    }
    grid.sync();
    // phase A: histogram of the bucket keys (the per-block work assignment is repeated exactly in phase B)
    int mine = 0;
    for (long long i = lo + tid; i < hi; i += blockDim.x) {
        const int s = LEAN ? (int)i : c.a.order[i];
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
        if (LEAN && tid == 0 && (long long)w.ctl[2] != c.a.counters[SSB_C_N_LIST])
            atomicExch((unsigned long long *)&c.a.counters[SSB_C_OVERFLOW], (unsigned long long)SWEEP_OVERFLOW_LIVE);
    }
    grid.sync();
    // phase B: one range reservation per block and key, then scatter
    for (int e = tid; e < NK; e += blockDim.x)          // (the counter becomes the cursor of this block's range of the key)
        sh_count[e] = sh_count[e] ? w.bucket_start[e] + atomicAdd(&w.bucket_fill[e], sh_count[e]) : 0;
    __syncthreads();
    for (long long i = lo + tid; i < hi; i += blockDim.x) {
        const int s = LEAN ? (int)i : c.a.order[i];
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
__host__ __device__ inline int sweep_build_smem_bytes(int rb, int kb, int halo) {
    return (FLOW_TX + 2 * halo) * (FLOW_TY + 2 * halo) * 4 + (FLOW_TX + 2 * halo) * 16 + FLOW_TX * FLOW_TY * 2 + sweep_tables_bytes(rb, kb, halo) + 16;
}

__global__ void __launch_bounds__(SWEEP_BUILD_THREADS) k_sweep_build(DevCtx c, SweepCtx w) {
    extern __shared__ __align__(16) unsigned char sweep_smem[];
    __shared__ int sh_tile, sh_n_agents, sh_class[16];
    const int R = w.halo, sw = FLOW_TX + 2 * R, sh = FLOW_TY + 2 * R;
    unsigned long long *mask = reinterpret_cast<unsigned long long *>(sweep_smem);
    uint32_t *cell = reinterpret_cast<uint32_t *>(mask + sw * 2);
    uint16_t *agents = reinterpret_cast<uint16_t *>(cell + sw * sh);
    int8_t *tables = reinterpret_cast<int8_t *>(agents + FLOW_TX * FLOW_TY);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int W = c.p.width, H = c.p.height;
    const int n_tiles = w.tiles_x * w.tiles_y;
    for (int i = tid; i < sweep_tables_bytes(w.rb, w.kb, R); i += SWEEP_BUILD_THREADS) tables[i] = w.tables[i];
    SweepTile tile;
    tile.cell = cell; tile.mask = mask; tile.sw = sw; tile.sh = sh; tile.halo = R;
    tile.any = tables; tile.pair = tables + sweep_any_size(w.rb, w.kb, R);
    for (;;) {
        __syncthreads();                                        // the previous tile's shared state is dead
        if (tid == 0) { sh_tile = (int)atomicAdd(&w.ctl[0], 1ull); sh_n_agents = 0; }
        __syncthreads();
        const int t = sh_tile;
        if (t >= n_tiles) break;
        if (w.tile_pop[t] == 0) continue;
        const int tx = t / w.tiles_y, ty = t - tx * w.tiles_y;
        const int x0 = tx * FLOW_TX - R, y0 = ty * FLOW_TY - R;
        // stage the low descriptor words: a column of the staged tile is a contiguous run of the x-major grid (up to the torus seam)
        const bool inside = x0 >= 0 && y0 >= 0 && x0 + sw <= W && y0 + sh <= H;
        if (inside) {
            const unsigned long long *src = w.desc + (size_t)x0 * H + y0;
#pragma unroll 4
            for (int job = tid; job < sw * sh; job += SWEEP_BUILD_THREADS) {
                const int col = job / sh, row = job - col * sh;
                cell[job] = (uint32_t)__ldg(src + (size_t)col * H + row);
            }
        } else {
            for (int job = tid; job < sw * sh; job += SWEEP_BUILD_THREADS) {
                const int col = job / sh, row = job - col * sh;
                cell[job] = (uint32_t)__ldg(w.desc + (size_t)wrapi(x0 + col, W) * H + wrapi(y0 + row, H));
            }
        }
        __syncthreads();
        // occupancy masks: one ballot per 32 rows
        for (int job = warp; job < sw * 4; job += SWEEP_BUILD_THREADS / 32) {
            const int col = job >> 2, part = job & 3, row = part * 32 + lane;
            const unsigned bits = __ballot_sync(0xffffffffu, row < sh && cell[col * sh + row] != 0u);
            if (lane == 0) reinterpret_cast<unsigned *>(mask + col * 2)[part] = bits;
        }
        // the tile's own agents (interior cells that lie on the map), grouped by range: the 32 agents a warp scans have
        // conflict regions of the same size
        if (tid < 16) sh_class[tid] = 0;
        __syncthreads();
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
        for (int a = tid; a < n_agents; a += SWEEP_BUILD_THREADS) {
            const int sc = agents[a], lx = sc / sh, ly = sc - lx * sh;
            const uint32_t gcell = (uint32_t)((tx * FLOW_TX + lx - R) * H + (ty * FLOW_TY + ly - R));
            const uint32_t index = (uint32_t)(__ldg(w.desc + gcell) >> 32);
            uint4 *out = reinterpret_cast<uint4 *>(w.preds + (size_t)index * SWEEP_COLS);
            uint4 acc = make_uint4(0u, 0u, 0u, 0u);
            sweep_pred_masks(tile, lx, ly, cell[sc], w.rb, w.kb, [&](int col, uint32_t m) {
                const int u = col & 3;
                if (u == 0) acc.x = m; else if (u == 1) acc.y = m; else if (u == 2) acc.z = m; else acc.w = m;
                if (u == 3) { out[col >> 2] = acc; acc = make_uint4(0u, 0u, 0u, 0u); }
            });
            // columns 2 R + 1 .. 30 stay empty; word 31 carries the start cell
            const int first = (2 * R + 1) >> 2;
            if (first == 7) { acc.w = gcell; out[7] = acc; }
            else { out[first] = acc; out[7] = make_uint4(0u, 0u, 0u, gcell); }
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
        if (w.tile_need[t] != w.stamp) continue;
        const int tx = t / w.tiles_y, ty = t - tx * w.tiles_y;
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
// All predecessors of list entry i done?  Straight-line code: the eight 16-byte loads of the agent's mask line are in
// flight together, then the done words of its columns in batches of eight (one 64-bit word per column, see the
// bitmap layout in ssb_sweep_core.cuh).  Also returns the start cell of the entry.
__device__ __noinline__ bool sweep_entry_ready_seam(const SweepCtx &w, const uint32_t *pm, int x, int y, int W, int H) {
    uint32_t masks[SWEEP_COLS];
    for (int k = 0; k < SWEEP_COLS; k++) masks[k] = __ldg(pm + k);
    return sweep_preds_done(masks, [&](size_t wd) { return sweep_ld_relaxed(w.done + wd); }, x, y, W, H, w.halo);
}
__device__ __forceinline__ bool sweep_entry_ready(const SweepCtx &w, int i, int W, int H, int *sx, int *sy) {
    const uint32_t *pm = w.preds + (size_t)i * SWEEP_COLS;
    uint4 v[8];
#pragma unroll
    for (int g = 0; g < 8; g++) v[g] = __ldg(reinterpret_cast<const uint4 *>(pm) + g);
    const uint32_t cell = v[7].w;                                      // (word 31 is the start cell, not a mask)
    const int x = (int)(cell / (uint32_t)H), y = (int)(cell - (uint32_t)x * (uint32_t)H);
    *sx = x; *sy = y;
    const SweepWindow win = sweep_window(x, y, W, H, w.halo);
    if (!win.straight) return sweep_entry_ready_seam(w, pm, x, y, W, H);
    int shift;
    const unsigned long long *col0 = w.done + sweep_done_window(0, win.y0, W, win.wpc, &shift);
    bool ok = true;
#pragma unroll
    for (int c0 = 0; c0 < 32; c0 += 8) {
        if (c0 >= win.span) break;
        unsigned long long dw[8];
        uint32_t m[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const int col = c0 + u;
            const uint4 q = v[col >> 2];
            m[u] = (col & 3) == 0 ? q.x : ((col & 3) == 1 ? q.y : ((col & 3) == 2 ? q.z : q.w));
            if (col >= win.span || col == 31) m[u] = 0u;
            int cx = win.x + col;
            if (cx < 0) cx += W; else if (cx >= W) cx -= W;
            dw[u] = sweep_ld_relaxed_if(col0 + (size_t)cx * win.wpc, m[u]);
        }
#pragma unroll
        for (int u = 0; u < 8; u++) ok &= ((uint32_t)(dw[u] >> shift) & m[u]) == m[u];
    }
    return ok;
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
    else if (now - idle_since > SWEEP_STALL_NS) {
        if ((threadIdx.x & 31) == 0) {
            atomicExch((unsigned long long *)&c.a.counters[SSB_C_OVERFLOW], (unsigned long long)SWEEP_OVERFLOW_STALL);
            atomicExch(w.ctl + 1, 1ull);
        }
        return true;
    }
    return *(const volatile unsigned long long *)(w.ctl + 1) != 0ull;
}

__host__ __device__ constexpr int lean_smem_bytes(int maxc, int threads) { return threads * (maxc * 4 + (maxc + 1) * 2); }

// The lean transaction reads everything mutable past L1 (lean_ld), so the consumer needs no acquire fence (whose
// CCTL.IVALL would invalidate the whole L1 once per agent); the producer's release rides on the reduction that sets
// the done bit.  MODE: see lean_transaction.
template <int MODE>
__global__ void __launch_bounds__(LEAN_THREADS, MODE == 0 ? 5 : 3) k_sweep_lean(DevCtx c, SweepCtx w, LeanCtx L) {
    extern __shared__ __align__(16) unsigned char lean_smem[];
    const int T = blockDim.x, tid = threadIdx.x;
    int32_t *occ_st = reinterpret_cast<int32_t *>(lean_smem);
    int16_t *sug_st = reinterpret_cast<int16_t *>(occ_st + L.maxc * T);
    const LeanStage st = {occ_st + tid, sug_st + tid, T};
    const int n = w.bucket_start[w.n_keys];
    const int W = c.p.width, H = c.p.height, wpc = sweep_done_wpc(H);
    int i = -1;
    bool exhausted = false;
    unsigned long long idle_since = 0;
    for (;;) {
        sweep_claim(w, n, true, i, exhausted);
        if (!__any_sync(0xffffffffu, i >= 0)) break;
        bool ran = false;
        int x = 0, y = 0;
        LeanB B;
        if (i >= 0) B = L.b[i];                                  // (read-only: in flight together with the mask line)
        const bool ready = i >= 0 && sweep_entry_ready(w, i, W, H, &x, &y);
        const unsigned readym = __ballot_sync(0xffffffffu, ready);      // (also re-converges the warp after the check)
        if (ready) {
            lean_transaction<MODE>(c, L, B, w.key, st, readym);
            sweep_set_done<true>(w, x, y, W, wpc);               // release: my writes happen before my done bit
            i = -1;
            ran = true;
        }
        if (w.diag) {
            const unsigned held = __ballot_sync(0xffffffffu, i >= 0 || ran), did = __ballot_sync(0xffffffffu, ran);
            if ((tid & 31) == 0) { atomicAdd(w.diag, 1ull); atomicAdd(w.diag + 1, (unsigned long long)__popc(held)); atomicAdd(w.diag + 2, (unsigned long long)__popc(did)); }
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
    int i = -1;
    bool exhausted = false;
    unsigned long long idle_since = 0;
    for (;;) {
        sweep_claim(w, n, lane == 0, i, exhausted);             // lane 0 of every group claims, the group follows
        i = Group<GL>::bcast(i, 0);
        if (!__any_sync(0xffffffffu, i >= 0)) break;
        bool ran = false;
        int x = 0, y = 0;
        if (i >= 0 && sweep_entry_ready(w, i, W, H, &x, &y)) {    // (every lane of the group reads the same words)
            sweep_fence();
            const int4 en = w.entries[i];
            const int s = en.x;
            AgentRec rec;
            rec.load(c, s);
            RngStream rng{w.key, (uint32_t)rec.id, 0u, 0u, nullptr};
            if (transaction_move_phase<GL, 4>(c, s, rec, rng, ws, MAXC_SCALABLE)) {
                if (lane == 0) transaction_interact_phase<RngStream, false>(c, s, rec, rng);
            }
            Group<GL>::sync();
            sweep_fence();
            if (lane == 0) sweep_set_done<false>(w, x, y, W, wpc);
            i = -1;
            ran = true;
        }
        if (__any_sync(0xffffffffu, ran)) idle_since = 0;
        else if (sweep_idle(c, w, idle_since)) break;
    }
}

__global__ void __launch_bounds__(256) k_lean_unpack(DevCtx c, LeanCtx L) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthreads = gridDim.x * blockDim.x;
    const int n_slots = (int)c.a.counters[SSB_C_N_SLOTS];
    for (int s = tid; s < n_slots; s += nthreads) lean_unpack_slot(c, L, s);
}

}  // namespace ssb
