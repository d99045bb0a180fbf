// ssb_dataflow.cuh -- mode S agent step as a dataflow over the static conflict DAG (ssb_flow_core.cuh).
//
// Three launches per step, no grid barrier inside any of them:
//   k_flow_prepare  one thread per list entry: priority, range r, dilation k; entries[i] = {slot, cell, priority,
//                   r | k << 8}; descriptor word of the agent's cell; population of its tile.
//   k_flow_build    one CTA per 64 x 64 tile: stages the tile + halo of descriptor words in shared memory
//                   (coalesced 8-byte loads, torus wrap), builds per-column occupancy bit masks, and lets one
//                   thread per agent walk the set bits of its conflict region: predecessors are counted
//                   (dep[i]), successors go to a CSR pool (one cursor bump per 256 agents); agents without
//                   predecessors are pushed to their ready queue.
//   k_flow_execute  persistent (cooperative launch for co-residency only): every lane group owns ONE
//                   single-consumer ready queue; it pops a ready agent, runs the whole transaction
//                   (ssb_rules.cuh, unchanged), and decrements the counters of its successors; whoever
//                   brings a counter to zero pushes that agent to ITS queue (list index mod queues, so the
//                   number of entries every queue will ever receive is known and no termination protocol
//                   is needed).  Memory ordering: writes -> fence.acq_rel.gpu -> atomic decrement ->
//                   fence.acq_rel.gpu -> queue store; consumer ld.acquire.gpu on the queue slot.
// Queues, tails, descriptors and tile populations are self-cleaning (whoever consumes an entry zeroes it),
// so a step needs no memset.
#pragma once
#include "ssb_agent_step.cuh"

namespace ssb {

constexpr int FLOW_BUILD_THREADS = 256;
constexpr int FLOW_BUF = 40;              // successors buffered per thread (staged cell ids); longer lists are re-scanned
constexpr int FLOW_OVERFLOW_POOL = 9;     // counters[SSB_C_OVERFLOW] code: successor pool exhausted
constexpr int FLOW_OVERFLOW_STALL = 10;   // ... a ready queue stayed empty for FLOW_STALL_NS: fail loudly instead of hanging the GPU
constexpr unsigned long long FLOW_STALL_NS = 4000000000ull;

struct FlowCtx {
    unsigned long long *desc;        // [n_cells] descriptor of the agent on the cell at the start of the step
    int4 *entries;                   // [capacity] by list index: {slot, cell, priority, r | k << 8}
    int32_t *dep;                    // [capacity] unfinished predecessors
    int2 *span;                      // [capacity] {begin, count} of the successor list in `pool`
    int32_t *pool;                   // successor lists (list indices)
    long long pool_capacity;
    unsigned long long *cursor;      // [0] pool cursor, [1] abort flag
    int32_t *queue;                  // [n_queues * queue_cap] list index + 1; 0 = empty
    int32_t *tail;                   // [n_queues]
    int32_t *tile_pop;               // [tiles_x * tiles_y] != 0: agents stand on the tile
    int32_t n_queues, queue_cap;
    int32_t rb, kb, halo;            // largest range / dilation of the step, halo = flow_halo(rb, kb)
    int32_t tiles_x, tiles_y;
    uint64_t key;
};

__device__ __forceinline__ int flow_ld_acquire(const int32_t *p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void flow_st_relaxed(int32_t *p, int v) {
    asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void flow_fence() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }

// agent i is ready: append it to the queue that owns it
__device__ __forceinline__ void flow_push(const FlowCtx &f, int i) {
    const int q = i % f.n_queues;
    const int p = atomicAdd(&f.tail[q], 1);
    flow_st_relaxed(&f.queue[(size_t)q * f.queue_cap + p], i + 1);
}

__global__ void __launch_bounds__(256) k_flow_prepare(DevCtx c, FlowCtx f) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthreads = gridDim.x * blockDim.x;
    const int n_list = (int)c.a.counters[SSB_C_N_LIST];
    if (tid == 0) { c.a.counters[SSB_C_N_BORN] = 0; c.a.counters[SSB_C_ROUNDS] = 1; f.cursor[0] = 0; f.cursor[1] = 0; }
    const int H = c.p.height;
    for (int i = tid; i < n_list; i += nthreads) {
        const int s = c.a.order[i];
        const uint32_t pr = priority_of(f.key, (uint32_t)c.a.id[s], c.p.id_bits);
        const int r = min(min(max(c.a.vision[s], 0), max(c.a.movement[s], 0)), max(c.max_dx, c.max_dy));
        const int k = footprint_dilation(c, s);
        const int pos = c.a.pos[s];
        f.entries[i] = make_int4(s, pos, (int)pr, r | (k << 8));
        if (pos >= 0) {
            f.desc[pos] = flow_pack((uint32_t)i, pr, r, k);
            const int x = pos / H, y = pos - x * H;
            f.tile_pop[(x / FLOW_TX) * f.tiles_y + y / FLOW_TY] = 1;      // (a flag: plain stores of the same value, no atomics)
        } else {        // not on the map (cannot happen after a compaction): nothing to order, runs as a no-op
            f.dep[i] = 0;
            f.span[i] = make_int2(0, 0);
            flow_push(f, i);
        }
    }
}

__host__ __device__ constexpr int flow_build_smem_bytes(int halo) {
    return ((FLOW_TX + 2 * halo) * (FLOW_TY + 2 * halo) + (FLOW_TX + 2 * halo) * 2) * 8 + FLOW_TX * FLOW_TY * 2 +
           FLOW_BUILD_THREADS * FLOW_BUF * 2 + 64;
}

__global__ void __launch_bounds__(FLOW_BUILD_THREADS) k_flow_build(DevCtx c, FlowCtx f) {
    extern __shared__ __align__(16) unsigned char flow_smem[];
    __shared__ int sh_n_agents, sh_warp_sum[FLOW_BUILD_THREADS / 32];
    __shared__ long long sh_pool_base;
    const int R = f.halo, sw = FLOW_TX + 2 * R, sh = FLOW_TY + 2 * R;
    unsigned long long *cell = reinterpret_cast<unsigned long long *>(flow_smem);
    unsigned long long *mask = cell + sw * sh;
    uint16_t *agents = reinterpret_cast<uint16_t *>(mask + sw * 2);
    uint16_t *buf = agents + FLOW_TX * FLOW_TY;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int W = c.p.width, H = c.p.height;
    const int n_tiles = f.tiles_x * f.tiles_y;
    FlowTile tile;
    tile.cell = cell; tile.mask = mask; tile.sw = sw; tile.sh = sh; tile.halo = R;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        if (f.tile_pop[t] == 0) continue;                       // (uniform: read before anyone resets it)
        __syncthreads();                                        // the previous tile's shared state is dead
        if (tid == 0) { f.tile_pop[t] = 0; sh_n_agents = 0; }
        const int tx = t / f.tiles_y, ty = t - tx * f.tiles_y;
        const int x0 = tx * FLOW_TX - R, y0 = ty * FLOW_TY - R;
        // stage: a column of the staged tile is a contiguous run of the x-major grid (up to the torus seam)
        for (int col = warp; col < sw; col += FLOW_BUILD_THREADS / 32) {
            const size_t gcol = (size_t)wrapi(x0 + col, W) * H;
            for (int row = lane; row < sh; row += 32) cell[col * sh + row] = f.desc[gcol + wrapi(y0 + row, H)];
        }
        __syncthreads();
        // occupancy masks: one ballot per 32 rows
        for (int job = warp; job < sw * 4; job += FLOW_BUILD_THREADS / 32) {
            const int col = job >> 2, part = job & 3, row = part * 32 + lane;
            const unsigned bits = __ballot_sync(0xffffffffu, row < sh && cell[col * sh + row] != 0ull);
            if (lane == 0) reinterpret_cast<unsigned *>(mask + col * 2)[part] = bits;
        }
        // the tile's own agents (interior cells that lie on the map)
        for (int idx = tid; idx < FLOW_TX * FLOW_TY; idx += FLOW_BUILD_THREADS) {
            const int ix = idx / FLOW_TY, iy = idx - ix * FLOW_TY;
            if (tx * FLOW_TX + ix >= W || ty * FLOW_TY + iy >= H) continue;
            const int sc = (R + ix) * sh + R + iy;
            if (cell[sc] != 0ull) agents[atomicAdd(&sh_n_agents, 1)] = (uint16_t)sc;
        }
        __syncthreads();
        const int n_agents = sh_n_agents;
        for (int base = 0; base < n_agents; base += FLOW_BUILD_THREADS) {
            const int a = base + tid;
            int npred = 0, nsucc = 0, sc = 0;
            unsigned long long da = 0;
            if (a < n_agents) {
                sc = agents[a];
                da = cell[sc];
                const uint32_t pa = flow_prio(da);
                flow_scan_agent(tile, sc / sh, sc % sh, da, f.rb, f.kb, [&](int cb, unsigned long long db) {
                    const uint32_t pb = flow_prio(db);
                    if (pb < pa) npred++;
                    else if (pb > pa) {
                        if (nsucc < FLOW_BUF) buf[nsucc * FLOW_BUILD_THREADS + tid] = (uint16_t)cb;
                        nsucc++;
                    }
                });
            }
            // exclusive scan of the list lengths over the CTA, one pool reservation for all of them
            int incl = nsucc;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
            if (lane == 31) sh_warp_sum[warp] = incl;
            __syncthreads();
            int warp_off = 0, total = 0;
#pragma unroll
            for (int w = 0; w < FLOW_BUILD_THREADS / 32; w++) { if (w < warp) warp_off += sh_warp_sum[w]; total += sh_warp_sum[w]; }
            if (tid == 0) sh_pool_base = (long long)atomicAdd(f.cursor, (unsigned long long)total);
            __syncthreads();
            const long long begin = sh_pool_base + warp_off + incl - nsucc;
            if (sh_pool_base + total > f.pool_capacity) {        // uniform: fail loudly, the execute kernel bails out
                if (tid == 0) { atomicExch((unsigned long long *)&c.a.counters[SSB_C_OVERFLOW], (unsigned long long)FLOW_OVERFLOW_POOL); atomicExch(f.cursor + 1, 1ull); }
                continue;
            }
            if (a < n_agents) {
                const int ia = (int)flow_index(da);
                if (nsucc <= FLOW_BUF) {
                    for (int j = 0; j < nsucc; j++) f.pool[begin + j] = (int)flow_index(cell[buf[j * FLOW_BUILD_THREADS + tid]]);
                } else {                                         // long list: walk the region again
                    const uint32_t pa = flow_prio(da);
                    int j = 0;
                    flow_scan_agent(tile, sc / sh, sc % sh, da, f.rb, f.kb, [&](int, unsigned long long db) {
                        if (flow_prio(db) > pa) f.pool[begin + j++] = (int)flow_index(db);
                    });
                }
                f.dep[ia] = npred;
                f.span[ia] = make_int2((int)begin, nsucc);
                if (npred == 0) flow_push(f, ia);
            }
        }
    }
}

// One persistent launch: every lane group of GL lanes owns the ready queue with its global group number.
__global__ void __launch_bounds__(AGENT_THREADS, 2) k_flow_execute(DevCtx c, FlowCtx f) {
    extern __shared__ __align__(8) unsigned char scratch_mem[];
    const WarpScratch ws = make_warp_scratch(scratch_mem + (threadIdx.x / GL) * warp_scratch_bytes(MAXC_SCALABLE), MAXC_SCALABLE);
    const int q = (blockIdx.x * blockDim.x + threadIdx.x) / GL, lane = Group<GL>::lane();
    const int n_list = (int)c.a.counters[SSB_C_N_LIST];
    const int n_q = q < n_list && q < f.n_queues ? (n_list - q + f.n_queues - 1) / f.n_queues : 0;
    int32_t *my_queue = f.queue + (size_t)q * f.queue_cap;
    const volatile unsigned long long *abort_flag = f.cursor + 1;
    int head = 0;
    unsigned long long idle_since = 0;
    for (;;) {
        int idx = 0;
        if (head < n_q) idx = flow_ld_acquire(my_queue + head);
        const bool have = Group<GL>::all(idx != 0);              // group-uniform (an entry is written once)
        if (have) {
            const int i = idx - 1;
            if (lane == 0) my_queue[head] = 0;
            head++;
            const int4 en = f.entries[i];
            const int s = en.x;
            AgentRec rec;
            rec.load(c, s);
            RngStream rng{f.key, (uint32_t)rec.id, 0u, 0u, nullptr};
            if (transaction_move_phase<GL, 4>(c, s, rec, rng, ws, MAXC_SCALABLE)) {
                if (lane == 0) transaction_interact_phase<RngStream, false>(c, s, rec, rng);
            }
            if (lane == 0 && en.y >= 0) f.desc[en.y] = 0ull;
            Group<GL>::sync();                                   // lane 0's writes happen before the group's releases
            flow_fence();
            const int2 sp = f.span[i];
            for (int j0 = 0; j0 < sp.y; j0 += GL * 4) {          // group-uniform trip count
                int b[4], old[4];
#pragma unroll
                for (int u = 0; u < 4; u++) { const int j = j0 + u * GL + lane; b[u] = j < sp.y ? f.pool[sp.x + j] : -1; }
#pragma unroll
                for (int u = 0; u < 4; u++) old[u] = b[u] >= 0 ? atomicSub(&f.dep[b[u]], 1) : 0;
#pragma unroll
                for (int u = 0; u < 4; u++)
                    if (old[u] == 1) { flow_fence(); flow_push(f, b[u]); }   // acquire the other predecessors' releases, publish
            }
        }
        __syncwarp();
        if (!__any_sync(0xffffffffu, head < n_q)) break;
        if (*abort_flag) break;
        if (__any_sync(0xffffffffu, have)) idle_since = 0;
        else {                                                   // the whole warp is waiting for predecessors elsewhere
            __nanosleep(100);
            const unsigned long long now = global_ns();
            if (idle_since == 0) idle_since = now;
            else if (now - idle_since > FLOW_STALL_NS) {
                if ((threadIdx.x & 31) == 0) { atomicExch((unsigned long long *)&c.a.counters[SSB_C_OVERFLOW], (unsigned long long)FLOW_OVERFLOW_STALL); atomicExch(f.cursor + 1, 1ull); }
                break;
            }
        }
    }
    if (lane == 0 && q < f.n_queues) f.tail[q] = 0;
}


}  // namespace ssb
