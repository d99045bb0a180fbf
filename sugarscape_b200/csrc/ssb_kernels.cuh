// ssb_kernels.cuh -- environment stencil, list compaction and the runtimeStats reductions.
#pragma once
#include "ssb_rules.cuh"

namespace ssb {

// ---------------------------------------------------------------------------------------------
// K1 env_step: growback with season gating (environment.py:61-96, cell.py:138-142), closed-form
// schedules of SURVEY.md Appendix A.3.  HBM-bound: 12 B per cell per resource (level read,
// capacity read, level write); the write is skipped for vectors that did not change.
// ---------------------------------------------------------------------------------------------
struct GrowArgs {
    int32_t *level;
    const int32_t *cap;
    int32_t rate;
    int32_t height, equator;
    int32_t seasons, flipped, dry_grows;
    long long n_cells;
    long long cell_lo, cell_hi;     // the cells this rank grows (the whole map unless sharded)
};

__device__ __forceinline__ int grow_one(int level, int cap, int rate, bool allowed) {
    return allowed ? min(level + rate, cap) : level;
}

// 128-bit path: requires height % 4 == 0 so that a vector never straddles two x-lines
__global__ void __launch_bounds__(256) k_growback_v4(GrowArgs g) {
    const long long n4 = g.cell_hi >> 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long v = (g.cell_lo >> 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; v < n4; v += stride) {
        int4 lv = reinterpret_cast<const int4 *>(g.level)[v];
        const int4 cp = __ldg(reinterpret_cast<const int4 *>(g.cap) + v);
        bool a0 = true, a1 = true, a2 = true, a3 = true;
        if (g.seasons) {
            const int y = (int)((v << 2) % g.height);
            a0 = ((y + 0 < g.equator) != (bool)g.flipped) || g.dry_grows;
            a1 = ((y + 1 < g.equator) != (bool)g.flipped) || g.dry_grows;
            a2 = ((y + 2 < g.equator) != (bool)g.flipped) || g.dry_grows;
            a3 = ((y + 3 < g.equator) != (bool)g.flipped) || g.dry_grows;
        }
        int4 nv;
        nv.x = grow_one(lv.x, cp.x, g.rate, a0);
        nv.y = grow_one(lv.y, cp.y, g.rate, a1);
        nv.z = grow_one(lv.z, cp.z, g.rate, a2);
        nv.w = grow_one(lv.w, cp.w, g.rate, a3);
        if (nv.x != lv.x || nv.y != lv.y || nv.z != lv.z || nv.w != lv.w)
            reinterpret_cast<int4 *>(g.level)[v] = nv;
    }
}

__global__ void __launch_bounds__(256) k_growback_scalar(GrowArgs g) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = g.cell_lo + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < g.cell_hi; i += stride) {
        bool allowed = true;
        if (g.seasons) {
            const int y = (int)(i % g.height);
            allowed = ((y < g.equator) != (bool)g.flipped) || g.dry_grows;
        }
        const int lv = g.level[i], nv = grow_one(lv, g.cap[i], g.rate, allowed);
        if (nv != lv) g.level[i] = nv;
    }
}

// Pollution diffusion (cell.py:104-109,24-25): Jacobi mean of N, S, E, W from the OLD field, the
// centre excluded, summed in that order starting from 0.  16 B per cell (one read, one write):
// the result goes to the second buffer and the two buffers swap roles (no copy back).
// x-major layout: N / S are the neighbours in memory, E / W sit one x-line (H cells) away, so the
// three x-lines a thread block touches are re-read from L1 / L2, not from HBM.
__global__ void __launch_bounds__(256) k_diffuse_v2(const double *__restrict__ p, double *__restrict__ out,
                                                    int W, int H, long long cell_lo, long long cell_hi) {
    // two cells (one 128-bit vector) per thread along y; requires H % 2 == 0
    const long long n2 = cell_hi >> 1;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const int H2 = H >> 1;
    for (long long v = (cell_lo >> 1) + (long long)blockIdx.x * blockDim.x + threadIdx.x; v < n2; v += stride) {
        const int x = (int)(v / H2), y = (int)(v - (long long)x * H2) << 1;
        const long long row = (long long)x * H;
        const double2 e = *reinterpret_cast<const double2 *>(p + (long long)(x == W - 1 ? 0 : x + 1) * H + y);
        const double2 w = *reinterpret_cast<const double2 *>(p + (long long)(x == 0 ? W - 1 : x - 1) * H + y);
        const double2 c = *reinterpret_cast<const double2 *>(p + row + y);
        const double n0 = p[row + (y == 0 ? H - 1 : y - 1)];
        const double s1 = p[row + (y + 2 == H ? 0 : y + 2)];
        double m0 = 0.0, m1 = 0.0;
        m0 += n0;  m0 += c.y; m0 += e.x; m0 += w.x;      // cell y:     N = y-1, S = y+1
        m1 += c.x; m1 += s1;  m1 += e.y; m1 += w.y;      // cell y + 1: N = y,   S = y+2
        *reinterpret_cast<double2 *>(out + row + y) = make_double2(m0 / 4, m1 / 4);
    }
}

// special: SSB_F_NOWRAP | SSB_F_MOORE bits -- the mean runs over cell.neighbors as the reference builds it (cell.py:61-89, 47-121,
// 104-109): N, S, E, W, then NE, NW, SE, SW with the Moore neighbourhood; without wraparound only the ones that exist
__global__ void __launch_bounds__(256) k_diffuse(const double *__restrict__ p, double *__restrict__ flux,
                                                 int W, int H, long long cell_lo, long long cell_hi, int special) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = cell_lo + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < cell_hi; i += stride) {
        const int x = (int)(i / H), y = (int)(i - (long long)x * H);
        const long long row = (long long)x * H;
        double m = 0.0;
        if (special) {
            const bool wrap = !(special & SSB_F_NOWRAP);
            const bool hn = wrap || y - 1 >= 0, hs = wrap || !(y + 1 < H - 1), he = wrap || !(x + 1 > W - 1), hw = wrap || x - 1 >= 0;
            const long long yn = y == 0 ? H - 1 : y - 1, ys = y == H - 1 ? 0 : y + 1;
            const long long re = (long long)(x == W - 1 ? 0 : x + 1) * H, rw = (long long)(x == 0 ? W - 1 : x - 1) * H;
            int n = 0;
            if (hn) { m += p[row + yn]; n++; }
            if (hs) { m += p[row + ys]; n++; }
            if (he) { m += p[re + y]; n++; }
            if (hw) { m += p[rw + y]; n++; }
            if (special & SSB_F_MOORE) {
                if (hn && he) { m += p[re + yn]; n++; }
                if (hn && hw) { m += p[rw + yn]; n++; }
                if (hs && he) { m += p[re + ys]; n++; }
                if (hs && hw) { m += p[rw + ys]; n++; }
            }
            flux[i] = m / n;
            continue;
        }
        m += p[row + (y == 0 ? H - 1 : y - 1)];
        m += p[row + (y == H - 1 ? 0 : y + 1)];
        m += p[(long long)(x == W - 1 ? 0 : x + 1) * H + y];
        m += p[(long long)(x == 0 ? W - 1 : x - 1) * H + y];
        flux[i] = m / 4;
    }
}

// ---------------------------------------------------------------------------------------------
// K5 death_compact: removeDeadAgents (sugarscape.py:843-853), order preserving.
// count -> scan -> scatter over tiles of COMPACT_TILE list entries.
// ---------------------------------------------------------------------------------------------
constexpr int COMPACT_THREADS = 256;
constexpr int COMPACT_ITEMS = 4;
constexpr int COMPACT_TILE = COMPACT_THREADS * COMPACT_ITEMS;

__device__ __forceinline__ bool list_entry_survives(const DevCtx &c, int s) {
    return agent_alive(c, s) && c.a.pos[s] >= 0;
}

__device__ __forceinline__ int block_exclusive_scan(int v, int *total) {
    // 256-thread exclusive scan: warp shuffles + one shared level
    __shared__ int warp_sums[COMPACT_THREADS / 32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += u;
    }
    if (lane == 31) warp_sums[w] = inc;
    __syncthreads();
    if (w == 0) {
        int ws = lane < COMPACT_THREADS / 32 ? warp_sums[lane] : 0;
#pragma unroll
        for (int o = 1; o < COMPACT_THREADS / 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, ws, o);
            if (lane >= o) ws += u;
        }
        if (lane < COMPACT_THREADS / 32) warp_sums[lane] = ws;
    }
    __syncthreads();
    const int base = w > 0 ? warp_sums[w - 1] : 0;
    *total = warp_sums[COMPACT_THREADS / 32 - 1];
    __syncthreads();
    return base + inc - v;
}

// the COMPACT_ITEMS consecutive list entries of this thread (-1 beyond the end of the list): one 16-byte load
static_assert(COMPACT_ITEMS == 4, "compact_load_entries reads one int4 per thread");
__device__ __forceinline__ void compact_load_entries(const DevCtx &c, long long tile0, long long n, int slots[COMPACT_ITEMS]) {
    const long long i0 = tile0 + (long long)threadIdx.x * COMPACT_ITEMS;
    if (i0 + COMPACT_ITEMS <= n) {
        const int4 v = *reinterpret_cast<const int4 *>(c.a.order + i0);
        slots[0] = v.x; slots[1] = v.y; slots[2] = v.z; slots[3] = v.w;
    } else {
#pragma unroll
        for (int k = 0; k < COMPACT_ITEMS; k++) slots[k] = i0 + k < n ? c.a.order[i0 + k] : -1;
    }
}

__global__ void __launch_bounds__(COMPACT_THREADS) k_compact_count(DevCtx c, int32_t *tile_alive) {
    const long long n = c.a.counters[SSB_C_N_LIST];
    const long long tile0 = (long long)blockIdx.x * COMPACT_TILE;
    if (tile0 >= n) { if (threadIdx.x == 0) tile_alive[blockIdx.x] = 0; return; }
    int slots[COMPACT_ITEMS];
    compact_load_entries(c, tile0, n, slots);
    int cnt = 0;
#pragma unroll
    for (int k = 0; k < COMPACT_ITEMS; k++)
        if (slots[k] >= 0 && list_entry_survives(c, slots[k])) cnt++;
    int total;
    block_exclusive_scan(cnt, &total);
    if (threadIdx.x == 0) tile_alive[blockIdx.x] = total;
}

// single block: exclusive scan of per-tile counts (in place); writes the grand total to *total
__global__ void __launch_bounds__(COMPACT_THREADS) k_scan_tiles(int32_t *tile_counts, const int64_t *n_items_ptr,
                                                                int items_per_tile, int64_t *total_out) {
    const long long n_items = *n_items_ptr;
    const int n_tiles = (int)((n_items + items_per_tile - 1) / items_per_tile);
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n_tiles; base += COMPACT_THREADS) {
        const int i = base + threadIdx.x;
        const int v = i < n_tiles ? tile_counts[i] : 0;
        int total;
        const int ex = block_exclusive_scan(v, &total);
        if (i < n_tiles) tile_counts[i] = carry + ex;
        __syncthreads();
        if (threadIdx.x == 0) carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total_out = carry;
}

// EXT: an ABI 2 extension is bound (doDeath then also settles the agent's loans / infections)
// (__grid_constant__: do_death takes the context by reference and is not inlined -- without it every thread would first copy
// the 850-byte parameter struct to its local memory: 8 GB of DRAM writes per S2 step, measured)
template <bool EXT>
__global__ void __launch_bounds__(COMPACT_THREADS) k_compact_scatter(const __grid_constant__ DevCtx c, const int32_t *tile_offset,
                                                                     int32_t *new_order) {
    const long long n = c.a.counters[SSB_C_N_LIST];
    const long long tile0 = (long long)blockIdx.x * COMPACT_TILE;
    if (tile0 >= n) return;
    __shared__ int32_t staged[COMPACT_TILE];     // the survivors of this tile in list order: written out coalesced
    int slots[COMPACT_ITEMS];
    bool alive[COMPACT_ITEMS];
    compact_load_entries(c, tile0, n, slots);
    int cnt = 0;
#pragma unroll
    for (int k = 0; k < COMPACT_ITEMS; k++) {
        alive[k] = slots[k] >= 0 && list_entry_survives(c, slots[k]);
        cnt += alive[k];
    }
    int total;
    int at = block_exclusive_scan(cnt, &total);
    // survivors keep list order; the dead of this tile go to dead_list[(tile0 - offset) ...]
    const int alive_base = tile_offset[blockIdx.x];
    int dead_at = (int)(tile0 - alive_base) + (threadIdx.x * COMPACT_ITEMS - at);
#pragma unroll
    for (int k = 0; k < COMPACT_ITEMS; k++) {
        if (slots[k] < 0) continue;
        if (alive[k]) staged[at++] = slots[k];
        else {
            const int s = slots[k];
            if (c.a.cod[s] == SSB_ALIVE) do_death<EXT>(c, s, SSB_OTHER);   // removeDeadAgents -> doDeath()
            else if (c.a.pos[s] >= 0) do_death<EXT>(c, s, c.a.cod[s]);
            c.a.dead_list[dead_at++] = s;
        }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < total; j += COMPACT_THREADS) new_order[alive_base + j] = staged[j];
}

// finish: publish the new list, recycle dead slots
__global__ void __launch_bounds__(256) k_compact_finish(DevCtx c, const int32_t *new_order, const int64_t *n_alive_ptr) {
    const long long n_old = c.a.counters[SSB_C_N_LIST];
    const long long n_alive = *n_alive_ptr, n_dead = n_old - n_alive;
    const long long n_free = c.a.counters[SSB_C_N_FREE];
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (long long i = tid; i < n_alive; i += stride) c.a.order[i] = new_order[i];
    for (long long i = tid; i < n_dead; i += stride) c.a.free_slots[n_free + i] = c.a.dead_list[i];
}

__global__ void k_compact_counters(DevCtx c, const int64_t *n_alive_ptr) {
    const long long n_old = c.a.counters[SSB_C_N_LIST];
    const long long n_alive = *n_alive_ptr;
    c.a.counters[SSB_C_N_DEAD] = n_old - n_alive;
    c.a.counters[SSB_C_N_FREE] += n_old - n_alive;
    c.a.counters[SSB_C_N_LIST] = n_alive;
}

__global__ void k_kin_reclaim(DevCtx c) { kin_reclaim(c); }

// Mode S: this step's surviving newborn are numbered and appended in CELL order (the oracle does
// the same; DESIGN.md "mode S").  Same count / scan / scatter pattern, over the cells.
__device__ __forceinline__ bool cell_has_unnumbered_newborn(const DevCtx &c, long long cell) {
    const int s = c.g.occ[cell];
    return s >= 0 && c.a.id[s] < 0;
}

// Sharded: a rank numbers the newborn standing on ITS slab (cell order = rank order, so the ranks'
// counts, all-gathered with the box-wide barrier, give every rank its first id); they join its list.
__device__ __forceinline__ long long newborn_cell_lo(const DevCtx &c) { return c.sh.world > 1 ? c.sh.cell_lo : 0; }
__device__ __forceinline__ long long newborn_cell_hi(const DevCtx &c) { return c.sh.world > 1 ? c.sh.cell_hi : c.g.n_cells; }

__global__ void __launch_bounds__(COMPACT_THREADS) k_newborn_count(DevCtx c, int32_t *tile_counts) {
    const long long n = newborn_cell_hi(c);
    const long long tile0 = newborn_cell_lo(c) + (long long)blockIdx.x * COMPACT_TILE;
    int cnt = 0;
#pragma unroll
    for (int k = 0; k < COMPACT_ITEMS; k++) {
        const long long i = tile0 + (long long)threadIdx.x * COMPACT_ITEMS + k;
        if (i < n && cell_has_unnumbered_newborn(c, i)) cnt++;
    }
    int total;
    block_exclusive_scan(cnt, &total);
    if (threadIdx.x == 0) tile_counts[blockIdx.x] = total;
}

__global__ void k_newborn_pack(const int64_t *n_new_ptr, unsigned long long *msg) {
    msg[0] = (unsigned long long)*n_new_ptr;
    for (int w = 1; w < SHARD_PAYLOAD; w++) msg[w] = 0;
}

__global__ void __launch_bounds__(COMPACT_THREADS) k_newborn_scatter(DevCtx c, const int32_t *tile_offset) {
    const long long n = newborn_cell_hi(c);
    const long long tile0 = newborn_cell_lo(c) + (long long)blockIdx.x * COMPACT_TILE;
    int slots[COMPACT_ITEMS];
    int cnt = 0;
#pragma unroll
    for (int k = 0; k < COMPACT_ITEMS; k++) {
        const long long i = tile0 + (long long)threadIdx.x * COMPACT_ITEMS + k;
        slots[k] = (i < n && cell_has_unnumbered_newborn(c, i)) ? c.g.occ[i] : -1;
        cnt += slots[k] >= 0;
    }
    int total;
    int at = block_exclusive_scan(cnt, &total) + tile_offset[blockIdx.x];
    const long long n_list = c.a.counters[SSB_C_N_LIST];
    long long next_id = c.a.counters[SSB_C_NEXT_ID];
    for (int r = 0; r < c.sh.rank && c.sh.world > 1; r++) next_id += (long long)c.sh.peer_vals[r * SHARD_PAYLOAD];   // lower slabs first
#pragma unroll
    for (int k = 0; k < COMPACT_ITEMS; k++) {
        if (slots[k] < 0) continue;
        if (n_list + at < c.a.capacity) c.a.order[n_list + at] = slots[k];
        else c.a.counters[SSB_C_OVERFLOW] = 1;
        c.a.id[slots[k]] = (int32_t)(next_id + at);
        at++;
    }
}

__global__ void k_newborn_counters(DevCtx c, const int64_t *n_new_ptr) {
    long long box = *n_new_ptr;
    if (c.sh.world > 1) {
        box = 0;
        for (int r = 0; r < c.sh.world; r++) box += (long long)c.sh.peer_vals[r * SHARD_PAYLOAD];
    }
    c.a.counters[SSB_C_N_LIST] += *n_new_ptr;
    c.a.counters[SSB_C_NEXT_ID] += box;
}

// ---------------------------------------------------------------------------------------------
// K6 stats_reduce: raw sums of updateRuntimeStatsPerGroup (sugarscape.py:1005-1426).
// Two deterministic stages: per-block partials in list order, then one block folds them.
// ---------------------------------------------------------------------------------------------
struct StatsPartial {
    long long population, sum_sugar_metab, sum_spice_metab, sum_movement, sum_vision, sum_age;
    long long sum_neighbors, sum_valid_moves, n_traders, trade_volume;
    long long n_dead, n_dead_moved, dead_starvation, dead_aging, dead_combat, sum_age_at_death;
    long long env_wealth_total, env_capacity_total, n_acted;
    double sum_wealth, max_wealth, min_wealth, sum_wealth_collected, sum_burn_rate, sum_ttl_age_limited;
    double sum_trade_price, sum_happiness, sum_wealth_happiness, sum_family_happiness;
    double dead_wealth_collected;
};

__device__ __forceinline__ void partial_zero(StatsPartial &p) {
    memset(&p, 0, sizeof(p));
    p.max_wealth = -1.0e300; p.min_wealth = 1.0e300;
}

__device__ __forceinline__ void partial_merge(StatsPartial &a, const StatsPartial &b) {
    long long *ai = &a.population; const long long *bi = &b.population;
    for (int i = 0; i < 19; i++) ai[i] += bi[i];
    a.sum_wealth += b.sum_wealth; a.max_wealth = fmax(a.max_wealth, b.max_wealth);
    a.min_wealth = fmin(a.min_wealth, b.min_wealth);
    a.sum_wealth_collected += b.sum_wealth_collected; a.sum_burn_rate += b.sum_burn_rate;
    a.sum_ttl_age_limited += b.sum_ttl_age_limited; a.sum_trade_price += b.sum_trade_price;
    a.sum_happiness += b.sum_happiness; a.sum_wealth_happiness += b.sum_wealth_happiness;
    a.sum_family_happiness += b.sum_family_happiness; a.dead_wealth_collected += b.dead_wealth_collected;
}

constexpr int STATS_THREADS = 128;

// block-level fold through shared memory, fixed tree order -> deterministic
__device__ void block_fold(StatsPartial &mine, StatsPartial *smem) {
    smem[threadIdx.x] = mine;
    __syncthreads();
    for (int o = STATS_THREADS / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) partial_merge(smem[threadIdx.x], smem[threadIdx.x + o]);
        __syncthreads();
    }
}

// tribe census (tribes dict, sugarscape.py:1171-1174; max(dict, key=get) needs first-seen order)
struct TribeCensus { unsigned long long count[32]; unsigned long long first[32]; unsigned long long key_or, key_and; };

__global__ void k_stats_reset(TribeCensus *tc) {
    if (threadIdx.x < 32) { tc->count[threadIdx.x] = 0; tc->first[threadIdx.x] = ~0ull; }
    if (threadIdx.x == 0) { tc->key_or = 0ull; tc->key_and = ~0ull; }
}

__global__ void __launch_bounds__(STATS_THREADS) k_stats_partials(DevCtx c, StatsPartial *out, double *wealth_keys,
                                                                  TribeCensus *tc) {
    __shared__ StatsPartial smem[STATS_THREADS];
    __shared__ unsigned long long tr_count[32], tr_first[32];
    if (threadIdx.x < 32) { tr_count[threadIdx.x] = 0; tr_first[threadIdx.x] = ~0ull; }
    __syncthreads();
    StatsPartial p;
    partial_zero(p);
    unsigned long long key_or = 0ull, key_and = ~0ull;
    const long long n = c.a.counters[SSB_C_N_LIST], n_dead = c.a.counters[SSB_C_N_DEAD];
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (long long i = tid; i < n; i += stride) {
        const int s = c.a.order[i];
        const double w = c.a.sugar[s] + c.a.spice[s];
        if (wealth_keys) wealth_keys[i] = w;
        key_or |= (unsigned long long)__double_as_longlong(w); key_and &= (unsigned long long)__double_as_longlong(w);
        p.population++;
        p.sum_sugar_metab += c.a.sugar_metab[s]; p.sum_spice_metab += c.a.spice_metab[s];
        p.sum_movement += c.a.movement[s]; p.sum_vision += c.a.vision[s]; p.sum_age += c.a.age[s];
        p.n_acted += c.a.age[s] > 0;
        p.sum_wealth += w; p.max_wealth = fmax(p.max_wealth, w); p.min_wealth = fmin(p.min_wealth, w);
        p.sum_wealth_collected += w - (c.a.last_sugar[s] + c.a.last_spice[s]);
        // findTimeToLive (agent.py:1113-1140), sys.maxsize when a metabolism is zero
        const double sm = (double)max(c.a.sugar_metab[s], 0), pm = (double)max(c.a.spice_metab[s], 0);
        const double st = sm > 0 ? c.a.sugar[s] / sm : 9223372036854775807.0;
        const double pt = pm > 0 ? c.a.spice[s] / pm : 9223372036854775807.0;
        const double ttl = fmin(st, pt);
        p.sum_burn_rate += ttl;
        p.sum_ttl_age_limited += fmin(ttl, (double)(c.a.max_age[s] - c.a.age[s]));
        p.sum_neighbors += c.a.n_neighborhood[s]; p.sum_valid_moves += c.a.n_valid_moves[s];
        if (c.a.trade_volume[s] > 0) {
            p.n_traders++; p.trade_volume += c.a.trade_volume[s]; p.sum_trade_price += c.a.trade_price[s];
        }
        p.sum_happiness += c.a.happiness[s]; p.sum_wealth_happiness += c.a.wealth_happiness[s];
        p.sum_family_happiness += c.a.family_happiness[s];
        const int tr = c.a.tribe[s];
        if (tr >= 0 && tr < 32) { atomicAdd(&tr_count[tr], 1ull); atomicMin(&tr_first[tr], (unsigned long long)i); }
    }
    for (long long i = tid; i < n_dead; i += stride) {      // this step's dead (1213-1265)
        const int s = c.a.dead_list[i];
        p.n_dead++;
        p.n_dead_moved += c.a.last_moved[s] == c.t;
        p.sum_age_at_death += c.a.age[s];
        p.dead_wealth_collected += (c.a.sugar[s] + c.a.spice[s]) - (c.a.last_sugar[s] + c.a.last_spice[s]);
        p.dead_starvation += c.a.cod[s] == SSB_STARVATION;
        p.dead_aging += c.a.cod[s] == SSB_AGING;
        p.dead_combat += c.a.cod[s] == SSB_COMBAT;
    }
    const bool spicy = (c.p.flags & SSB_F_SPICE) != 0;      // spice arrays are all zero otherwise
    const long long cell_lo = c.sh.world > 1 ? c.sh.cell_lo : 0, cell_hi = c.sh.world > 1 ? c.sh.cell_hi : c.g.n_cells;
    for (long long i = cell_lo + tid; i < cell_hi; i += stride)   // environment totals (1044-1051)
        p.env_wealth_total += c.g.sugar[i] + (spicy ? c.g.spice[i] : 0);
    if (c.t <= 1)                                           // capacities are constant: only t = 1 logs them
        for (long long i = cell_lo + tid; i < cell_hi; i += stride)
            p.env_capacity_total += c.g.sugar_cap[i] + (spicy ? c.g.spice_cap[i] : 0);
    block_fold(p, smem);
    if (threadIdx.x == 0) out[blockIdx.x] = smem[0];
    if (threadIdx.x < 32 && tr_count[threadIdx.x]) {
        atomicAdd(&tc->count[threadIdx.x], tr_count[threadIdx.x]);
        atomicMin(&tc->first[threadIdx.x], tr_first[threadIdx.x]);
    }
    // OR / AND of the wealth keys (which digits of the radix sort are constant)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        key_or |= __shfl_xor_sync(0xffffffffu, key_or, o);
        key_and &= __shfl_xor_sync(0xffffffffu, key_and, o);
    }
    if ((threadIdx.x & 31) == 0 && n > 0) { atomicOr(&tc->key_or, key_or); atomicAnd(&tc->key_and, key_and); }
}

__global__ void __launch_bounds__(STATS_THREADS) k_stats_final(DevCtx c, const StatsPartial *partials, int n_partials,
                                                               const TribeCensus *tc, ssb_stats_t *out) {
    __shared__ StatsPartial smem[STATS_THREADS];
    StatsPartial p;
    partial_zero(p);
    for (int i = threadIdx.x; i < n_partials; i += STATS_THREADS) partial_merge(p, partials[i]);
    block_fold(p, smem);
    if (threadIdx.x != 0) return;
    const StatsPartial &r = smem[0];
    out->timestep = c.t;
    out->population = r.population;
    out->sum_sugar_metab = r.sum_sugar_metab; out->sum_spice_metab = r.sum_spice_metab;
    out->sum_movement = r.sum_movement; out->sum_vision = r.sum_vision; out->sum_age = r.sum_age;
    out->sum_wealth = r.sum_wealth;
    out->max_wealth = r.population ? r.max_wealth : 0; out->min_wealth = r.population ? r.min_wealth : 0;
    out->sum_wealth_collected = r.sum_wealth_collected; out->sum_burn_rate = r.sum_burn_rate;
    out->sum_ttl_age_limited = r.sum_ttl_age_limited;
    out->sum_neighbors = r.sum_neighbors; out->sum_valid_moves = r.sum_valid_moves;
    out->n_traders = r.n_traders; out->trade_volume = r.trade_volume; out->sum_trade_price = r.sum_trade_price;
    out->sum_happiness = r.sum_happiness; out->sum_wealth_happiness = r.sum_wealth_happiness;
    out->sum_family_happiness = r.sum_family_happiness;
    for (int i = 0; i < 32; i++) {
        out->tribe_count[i] = (int64_t)tc->count[i];
        out->tribe_first[i] = tc->count[i] ? (int64_t)tc->first[i] : -1;
    }
    out->n_dead = r.n_dead; out->n_dead_moved = r.n_dead_moved; out->dead_starvation = r.dead_starvation;
    out->dead_aging = r.dead_aging; out->dead_combat = r.dead_combat; out->sum_age_at_death = r.sum_age_at_death;
    out->dead_wealth_collected = r.dead_wealth_collected;
    out->n_born = c.a.counters[SSB_C_N_BORN];
    out->n_replaced = c.a.counters[SSB_C_N_REPLACED];
    out->n_acted = r.n_acted;
    out->env_wealth_total = r.env_wealth_total;
    out->env_capacity_total = r.env_capacity_total;
    out->env_wealth_created = 0;   // closed form, filled in by the host
    out->rounds = c.a.counters[SSB_C_ROUNDS];
    out->gini_area = -1.0;
}

// sharded: {population, OR, AND of the wealth keys} of this rank for the all-gather
__global__ void k_stats_pack(DevCtx c, const TribeCensus *tc, unsigned long long *msg) {
    msg[0] = (unsigned long long)c.a.counters[SSB_C_N_LIST];
    msg[1] = tc->key_or; msg[2] = tc->key_and;
    msg[3] = msg[4] = msg[5] = msg[6] = 0;
}

// sharded: all ranks' wealth keys (rank r's start at word r * gather_stride + 4 * local_slots of the
// stitched scratch) into one local array; vals[0] = box population, vals[1..2] = OR / AND of all keys
__global__ void __launch_bounds__(256) k_keys_gather(DevCtx c, uint64_t *keys, long long capacity, long long local_slots,
                                                     unsigned long long *vals) {
    const long long stride = (long long)gridDim.x * blockDim.x, tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long base = 0;
    unsigned long long k_or = 0ull, k_and = ~0ull;
    for (int r = 0; r < c.sh.world; r++) {
        const long long n = (long long)c.sh.peer_vals[r * SHARD_PAYLOAD];
        k_or |= c.sh.peer_vals[r * SHARD_PAYLOAD + 1]; k_and &= c.sh.peer_vals[r * SHARD_PAYLOAD + 2];
        const unsigned long long *src = c.sh.gather + (long long)r * c.sh.gather_stride + 4 * local_slots;
        for (long long i = tid; i < n; i += stride)
            if (base + i < capacity) keys[base + i] = src[i];
        base += n;
    }
    if (tid == 0) { vals[0] = (unsigned long long)(base < capacity ? base : capacity); vals[1] = k_or; vals[2] = k_and; }
}

// Mode E: the reference accumulates its float sums agent by agent in list order and rounds the
// means to 2 decimals, so a last-bit difference can flip a logged value.  One thread re-folds the
// order-sensitive fp64 sums sequentially (example scale: <= a few thousand agents).
__global__ void k_stats_ordered(DevCtx c, ssb_stats_t *out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const long long n = c.a.counters[SSB_C_N_LIST], n_dead = c.a.counters[SSB_C_N_DEAD];
    double sum_wealth = 0, collected = 0, burn = 0, ttl_lim = 0, price = 0, hap = 0, whap = 0, fhap = 0;
    for (long long i = 0; i < n; i++) {
        const int s = c.a.order[i];
        const double w = c.a.sugar[s] + c.a.spice[s];
        sum_wealth += w;
        collected += w - (c.a.last_sugar[s] + c.a.last_spice[s]);
        const double sm = (double)max(c.a.sugar_metab[s], 0), pm = (double)max(c.a.spice_metab[s], 0);
        const double st = sm > 0 ? c.a.sugar[s] / sm : 9223372036854775807.0;
        const double pt = pm > 0 ? c.a.spice[s] / pm : 9223372036854775807.0;
        const double ttl = fmin(st, pt);
        burn += ttl;
        ttl_lim += fmin(ttl, (double)(c.a.max_age[s] - c.a.age[s]));
        if (c.a.trade_volume[s] > 0) price += c.a.trade_price[s];
        hap += c.a.happiness[s]; whap += c.a.wealth_happiness[s]; fhap += c.a.family_happiness[s];
    }
    double dead_collected = 0;
    for (long long i = 0; i < n_dead; i++) {
        const int s = c.a.dead_list[i];
        dead_collected += (c.a.sugar[s] + c.a.spice[s]) - (c.a.last_sugar[s] + c.a.last_spice[s]);
    }
    out->sum_wealth = sum_wealth; out->sum_wealth_collected = collected; out->sum_burn_rate = burn;
    out->sum_ttl_age_limited = ttl_lim; out->sum_trade_price = price; out->sum_happiness = hap;
    out->sum_wealth_happiness = whap; out->sum_family_happiness = fhap;
    out->dead_wealth_collected = dead_collected;
}

// Lorenz numerator WITHOUT a sort, for populations whose wealths are all small non-negative integers (S2: integral
// holdings and metabolisms): a histogram over the values [0, LORENZ_BINS) gives every value its rank range
// [start, start + count), and  sum_j w_(j) * (n - 1/2 - j)  is then a sum over the bins in exact integer arithmetic
// (128 bits), rounded once at the end.  lz[0] != 0: some wealth is not such an integer -- the radix sort below takes over
// (every kernel of that path returns at once otherwise).  lz[1 ..]: the histogram.
constexpr int LORENZ_BINS = 8192;
constexpr int LORENZ_THREADS = 256;
__global__ void __launch_bounds__(LORENZ_THREADS) k_lorenz_hist(const double *keys, const int64_t *n_ptr, unsigned long long *lz) {
    __shared__ uint32_t h[LORENZ_BINS];
    for (int i = threadIdx.x; i < LORENZ_BINS; i += LORENZ_THREADS) h[i] = 0u;
    __syncthreads();
    const long long n = *n_ptr, stride = (long long)gridDim.x * blockDim.x;
    bool bad = false;
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
        const double w = keys[j];
        const int v = (w >= 0.0 && w < (double)LORENZ_BINS) ? (int)w : -1;
        if (v < 0 || (double)v != w) { bad = true; continue; }
        atomicAdd(&h[v], 1u);
    }
    if (bad) lz[0] = 1ull;
    __syncthreads();
    for (int i = threadIdx.x; i < LORENZ_BINS; i += LORENZ_THREADS) if (h[i]) atomicAdd(lz + 1 + i, (unsigned long long)h[i]);
}
__global__ void k_lorenz_reset(unsigned long long *lz) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i <= LORENZ_BINS; i += gridDim.x * blockDim.x) lz[i] = 0ull;
}
// one block: 2 * numerator = sum_v v * count_v * (2 n - 1 - 2 start_v - (count_v - 1))
__global__ void __launch_bounds__(LORENZ_THREADS) k_lorenz_from_hist(const unsigned long long *lz, const int64_t *n_ptr, ssb_stats_t *out) {
    if (lz[0] != 0ull) return;
    constexpr int PER = LORENZ_BINS / LORENZ_THREADS;
    __shared__ unsigned long long scan[LORENZ_THREADS], lo_sum[LORENZ_THREADS], hi_sum[LORENZ_THREADS];
    const int tid = threadIdx.x;
    unsigned long long mine = 0;
    for (int k = 0; k < PER; k++) mine += lz[1 + tid * PER + k];
    scan[tid] = mine;
    __syncthreads();
    if (tid == 0) { unsigned long long run = 0; for (int i = 0; i < LORENZ_THREADS; i++) { const unsigned long long v = scan[i]; scan[i] = run; run += v; } }
    __syncthreads();
    const unsigned long long n = (unsigned long long)*n_ptr;
    unsigned long long start = scan[tid];
    unsigned __int128 acc = 0;
    for (int k = 0; k < PER; k++) {
        const unsigned long long v = (unsigned long long)(tid * PER + k), cnt = lz[1 + tid * PER + k];
        if (cnt) acc += (unsigned __int128)(v * cnt) * (unsigned __int128)(2ull * n - 1ull - 2ull * start - (cnt - 1ull));
        start += cnt;
    }
    lo_sum[tid] = (unsigned long long)acc; hi_sum[tid] = (unsigned long long)(acc >> 64);
    __syncthreads();
    if (tid != 0) return;
    unsigned __int128 total = 0;
    for (int i = 0; i < LORENZ_THREADS; i++) total += ((unsigned __int128)hi_sum[i] << 64) | lo_sum[i];
    out->lorenz_weighted_sum = ((double)(unsigned long long)(total >> 64) * 18446744073709551616.0 + (double)(unsigned long long)total) * 0.5;
}

// Lorenz numerator from the ascending wealth keys (updateGiniCoefficient, sugarscape.py:913-934):
// sum_{i<n-1} cum_i + cum_{n-1}/2 = sum_j w_j * (n - 1 - j) + w_j / 2 ... folded per block.
__global__ void __launch_bounds__(256) k_lorenz(const double *sorted, const int64_t *n_ptr, double *block_sums, const unsigned long long *needed) {
    if (needed && *needed == 0ull) return;       // (the histogram path has the result)
    const long long n = *n_ptr;
    const long long stride = (long long)gridDim.x * blockDim.x;
    double acc = 0;
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride)
        acc += sorted[j] * ((double)(n - 1 - j) + 0.5);
    __shared__ double sm[256];
    sm[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) block_sums[blockIdx.x] = sm[0];
}

// Exact mode: the Lorenz area in the reference's own order of operations (updateGiniCoefficient, sugarscape.py:913-934) over
// the ascending wealth keys -- one thread, like everything that is ordered in that mode.
__global__ void k_gini_ordered(const double *sorted, const int64_t *n_ptr, ssb_stats_t *out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const long long n = *n_ptr;
    if (n <= 0) { out->gini_area = -1.0; return; }
    double total = 0;
    for (long long i = 0; i < n; i++) total += sorted[i];
    double lorenz = 0;
    if (total != 0) {
        double cum = 0;
        for (long long i = 0; i + 1 < n; i++) { cum += sorted[i]; lorenz += cum / total; }
        cum += sorted[n - 1];
        lorenz += (cum / 2) / total;
        lorenz /= (double)n;
    }
    out->gini_area = lorenz;
}

__global__ void k_lorenz_final(const double *block_sums, int n_blocks, ssb_stats_t *out, ssb_stats_t *ring_slot, const unsigned long long *needed) {
    if (!needed || *needed != 0ull) {
        double acc = 0;
        for (int i = 0; i < n_blocks; i++) acc += block_sums[i];
        out->lorenz_weighted_sum = acc;
    }
    *ring_slot = *out;      // the step's stats are complete: publish them in the history ring
}

}  // namespace ssb
