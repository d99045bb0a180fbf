// ssb_agent_step.cuh -- the agent loop of Sugarscape.doTimestep (reference sugarscape.py:400-407).
//
// Mode E (exact): one CTA, agents strictly in list order, the CPython MT19937 stream in shared
//   memory.  Data-dependent consumption of the single global stream makes exact replay
//   inherently ordered; this mode serves the shipped 50x50
//   examples.
//
// Mode S (scalable): the step is DEFINED as "agents act one at a time in ascending priority"
//   (priority = keyed bijection of the agent id, ssb_rng.cuh).  The kernel below reproduces that
//   sequential result in parallel with deterministic reservations over a SLIDING PREFIX of the
//   priority order (after Blelloch et al., "Internally deterministic parallel algorithms can be
//   fast", PPoPP'12):
//     prepare   agents are bucketed by the top bits of their priority into E epochs;
//     round k   the active set is epoch k plus the agents of earlier epochs that have not
//               committed yet;
//       MARK    every active agent atomicMin's (tag_k | priority) onto every cell of its
//               footprint F (all cells its transaction may read or write),
//       SELECT  an agent whose own value survived on ALL of F is first among every uncommitted
//               agent it could interact with -> ready (agents of later epochs have later
//               priorities and are not acting yet, so they cannot matter),
//       EXECUTE ready agents run their whole transaction concurrently (their footprints are
//               pairwise disjoint); everyone else stays active.
//   Two agents whose footprints intersect are therefore always committed in priority order, which
//   is the sequential semantics; activating only a slice of the order per round keeps the
//   conflict rate low, so an agent marks its footprint ~2 times per step instead of ~50.
//   Marks carry a descending round tag in their high bits: stale marks lose every atomicMin and
//   are cleared only when the 6-bit tag wraps.
//   Footprints and the two commit phases.  A transaction has a MOVE part (inspect the cross of
//   radius r, move, collect, metabolise: touches cross + own cell only) and an INTERACT part
//   (tagging, trade, reproduction, aging around the NEW cell p': touches the diamond D_k(p'),
//   k = 1 for tagging / trading, 2 for reproduction, 0 if none of them applies to the agent).
//     phase 0 agent  marks  M0 = D_k(cross + own cell)   (everything it may ever touch)
//                    needs  C0 = cross + own cell        -> commits its move part; if it also
//                           owns D_k(p') (a subset of M0) it finishes in the same round,
//     phase 1 agent  marks and needs D_k(p')             -> commits the interact part.
//   An earlier agent that could still touch a cell of C0 (or of D_k(p')) would have its smaller
//   value there, so owning the set means every earlier interferer has committed; later agents are
//   kept out by the marks, which always cover whatever the agent may still touch.  Checking only
//   the 4r + 1 cells of the cross instead of the dilated footprint cuts the conflict rate of
//   reproduction rule sets by an order of magnitude.
//   In the x-major grid a footprint is one contiguous y-interval per x-column.
//   All rounds of a step run inside ONE persistent cooperative kernel (grid.sync between phases).
#pragma once
#include <cooperative_groups.h>

#include "ssb_rules.cuh"

namespace ssb {
namespace cg = cooperative_groups;

// ---------------------------------------------------------------------------------------------
// mode E
// ---------------------------------------------------------------------------------------------
constexpr int MAXC_EXACT = 128;   // up to 4 * 32 candidate cells

__global__ void __launch_bounds__(32, 1) k_agent_step_exact(DevCtx c, uint32_t *mt_global) {
    __shared__ uint32_t mt[625];
    __shared__ __align__(8) unsigned char scratch_mem[warp_scratch_bytes(MAXC_EXACT)];
    const int lane = threadIdx.x;
    for (int i = lane; i < 625; i += 32) mt[i] = mt_global[i];
    __syncwarp();
    RngExact rng{mt};
    const WarpScratch ws = make_warp_scratch(scratch_mem, MAXC_EXACT);
    volatile unsigned long long *cn = (volatile unsigned long long *)c.a.counters;
    if (lane == 0) {
        cn[SSB_C_N_BORN] = 0;
        shuffle(rng, c.a.order, (int)cn[SSB_C_N_LIST]);        // random.shuffle(self.agents), sugarscape.py:400
    }
    __syncwarp();
    // the list grows while it is iterated: children are appended, visited and skipped
    for (long long i = 0;; i++) {
        const long long n_list = __shfl_sync(0xffffffffu, (long long)cn[SSB_C_N_LIST], 0);
        if (i >= n_list) break;
        const int s = c.a.order[i];
        AgentRec rec;
        rec.load(c, s);
        const bool go_on = transaction_move_phase<32, 4>(c, s, rec, rng, ws, MAXC_EXACT);
        if (go_on && lane == 0) transaction_interact_phase<RngExact, true>(c, s, rec, rng);
        __threadfence_block();
        __syncwarp();
    }
    for (int i = lane; i < 625; i += 32) mt_global[i] = mt[i];
}

// ---------------------------------------------------------------------------------------------
// mode S
// ---------------------------------------------------------------------------------------------
constexpr int MAXC_SCALABLE = 32;     // 4 * 8 candidate cells; the host refuses larger ranges
constexpr int TAG_BITS = 6;
constexpr int MAX_EPOCHS = 1024;
constexpr int MAX_ROUNDS = 4096;
constexpr int AGENT_THREADS = 256;
#ifndef SSB_GROUP_LANES
#define SSB_GROUP_LANES 4
#endif
#ifndef SSB_MARK_LANES
#define SSB_MARK_LANES 8
#endif
constexpr int GL = SSB_GROUP_LANES;          // lanes cooperating on one transaction (EXECUTE)
constexpr int GM = SSB_MARK_LANES;           // lanes cooperating on one footprint (MARK / SELECT)
constexpr int AGENT_GROUPS = AGENT_THREADS / GL;
constexpr int AGENT_SMEM_BYTES = warp_scratch_bytes(MAXC_SCALABLE) * AGENT_GROUPS;
constexpr int PHASE1_BIT = 0x40000000;   // list entries: slot | PHASE1_BIT once the move part has committed

struct RoundCtx {
    uint32_t *mark;          // one word per cell
    uint32_t *prio;          // per slot, this step's priority
    int32_t *bucketed;       // the agent list grouped by epoch
    int32_t *list_a, *list_b, *ready;
    int32_t *cnt;            // [0..1] carry-list lengths (ping-pong), [2..3] ready counts (ping-pong)
    int32_t *epoch_count;    // [MAX_EPOCHS]
    int32_t *epoch_start;    // [MAX_EPOCHS + 1]
    int32_t *epoch_fill;     // [MAX_EPOCHS]
    int32_t epochs, epoch_shift;
    uint64_t key;
    unsigned long long *trace;   // optional: per round {n_active, n_ready, ns at start / after mark / select / execute}
    int32_t trace_rounds;
};

__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// Dilation of the footprint required by the rule set, for agent s (agent.py:585-588).
// Reproduction (k = 2) is only possible if the agent can be fertile when it gets there: fertile
// age, and holdings that can still reach the starting endowment -- other agents can only LOWER
// them (mate costs), the agent's own move adds at most one cell's capacity (plus combat loot)
// and subtracts its metabolism.  `moved` = the move part has committed (holdings are final).
__device__ __forceinline__ int footprint_dilation(const DevCtx &c, int s, bool moved) {
    int k = 0;
    if (c.p.flags & (SSB_F_TAGGING | SSB_F_TRADE)) k = 1;
    if ((c.p.flags & SSB_F_REPRO) && c.a.fert_factor[s] > 0 &&
        c.a.age[s] >= c.a.fert_age[s] && c.a.age[s] < c.a.infert_age[s]) {
        bool can = true;
        if (!(c.p.flags & SSB_F_TRADE)) {
            double gain_s = 0, gain_p = 0;
            if (!moved) {
                const double loot = c.a.aggression[s] > 0 ? c.p.max_combat_loot : 0.0;
                gain_s = (double)c.p.max_sugar_capacity + loot - (double)max(c.a.sugar_metab[s], 0);
                gain_p = (double)c.p.max_spice_capacity + loot - (double)max(c.a.spice_metab[s], 0);
            }
            can = c.a.sugar[s] + gain_s >= c.a.start_sugar[s] && c.a.spice[s] + gain_p >= c.a.start_spice[s];
        }
        if (can) k = 2;
    }
    return k;
}

// the same from the register copy of the record, after the move part (holdings are final)
__device__ __forceinline__ int footprint_dilation_moved(const DevCtx &c, const AgentRec &rec) {
    int k = 0;
    if (c.p.flags & (SSB_F_TAGGING | SSB_F_TRADE)) k = 1;
    if ((c.p.flags & SSB_F_REPRO) && rec.fert_factor > 0 && rec.age >= rec.fert_age && rec.age < rec.infert_age &&
        ((c.p.flags & SSB_F_TRADE) || (rec.sugar >= rec.start_sugar && rec.spice >= rec.start_spice))) k = 2;
    return k;
}

// Footprint = union over the cells c of (cross of radius r) + {pos} of the diamond of radius k
// around c.  Per x-column offset i it is the y-interval [-h(i), h(i)] with
//   h(i) = max( |i| <= k ? ry + k - |i| : -1 ,                  (north-south arm, dilated)
//               |i| <= rx ? k : k - (|i| - rx) )                (east-west arm, dilated)
// f(base_index_of_column, y_lo, y_hi) is called once per column with the (unwrapped) interval.
struct Footprint {
    int x, y, rx, ry, k, W, H;
    __device__ __forceinline__ void init(const DevCtx &c, int pos, int r, int k_) {
        W = c.p.width; H = c.p.height;
        x = pos / H; y = pos - x * H;
        rx = min(r, c.max_dx); ry = min(r, c.max_dy); k = k_;
        // on tiny maps the intervals may wrap onto themselves: cells are then visited twice,
        // which is harmless (atomicMin and the ownership test are idempotent)
    }
    __device__ __forceinline__ int half_height(int i) const {
        const int ai = abs(i);
        const int ns = ai <= k ? ry + k - ai : -1;
        const int ew = ai <= rx ? k : k - (ai - rx);
        return max(ns, ew);
    }
};

__device__ __forceinline__ uint32_t round_tag(const DevCtx &c, int round) {
    constexpr int per_epoch = (1 << TAG_BITS) - 1;          // tag 63 is the cleared value
    return (uint32_t)(per_epoch - 1 - (round % per_epoch)) << c.p.id_bits;
}
constexpr int ROUNDS_PER_TAG_EPOCH = (1 << TAG_BITS) - 1;

// The footprint D_k(cross of radius (rx, ry) + centre), visited by a lane group of G lanes: one
// x-column per iteration, the lanes cover its y-interval (at most 2 (ry + k) + 1 <= 21 cells).
template <int G>
__device__ __forceinline__ void group_mark_region(const DevCtx &c, const RoundCtx &rc, int pos, int r, int k, uint32_t v) {
    const int lane = Group<G>::lane();
    Footprint f;
    f.init(c, pos, r, k);
    for (int i = -(f.rx + f.k); i <= f.rx + f.k; i++) {
        const int h = f.half_height(i);
        uint32_t *col = rc.mark + (size_t)wrap1(f.x + i, f.W) * f.H;
        for (int j = lane - h; j <= h; j += G) atomicMin(col + wrap1(f.y + j, f.H), v);
    }
}

// true (on every lane of the group) iff every cell of the region holds v
template <int G>
__device__ __forceinline__ bool group_owns_region(const DevCtx &c, const RoundCtx &rc, int pos, int r, int k, uint32_t v) {
    const int lane = Group<G>::lane();
    Footprint f;
    f.init(c, pos, r, k);
    unsigned bad = 0;
    for (int i = -(f.rx + f.k); i <= f.rx + f.k; i++) {
        const int h = f.half_height(i);
        const uint32_t *col = rc.mark + (size_t)wrap1(f.x + i, f.W) * f.H;
        for (int j = lane - h; j <= h; j += G) bad |= __ldcg(col + wrap1(f.y + j, f.H)) ^ v;
    }
    return !Group<G>::any(bad != 0);
}

// warp-aggregated append: every lane of the warp calls this (full mask)
__device__ __forceinline__ void warp_append(bool want, int32_t *list, int32_t *counter, int value) {
    const unsigned mask = __ballot_sync(0xffffffffu, want);
    if (mask == 0) return;
    const int lane = threadIdx.x & 31, leader = __ffs(mask) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(counter, __popc(mask));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (want) list[base + __popc(mask & ((1u << lane) - 1u))] = value;
}

// One persistent cooperative kernel per step.
__global__ void __launch_bounds__(AGENT_THREADS, 2) k_agent_step_scalable(DevCtx c, RoundCtx rc) {
    cg::grid_group grid = cg::this_grid();
    __shared__ int32_t sh_count[MAX_EPOCHS], sh_base[MAX_EPOCHS];
    extern __shared__ __align__(8) unsigned char scratch_mem[];
    const WarpScratch ws = make_warp_scratch(scratch_mem + (threadIdx.x / GL) * warp_scratch_bytes(MAXC_SCALABLE), MAXC_SCALABLE);
    const int warp_id = (blockIdx.x * blockDim.x + threadIdx.x) / GL;      // lane-group index
    const int n_warps = (gridDim.x * blockDim.x) / GL;
    const int glane = Group<GL>::lane();
    const int mgroup_id = (blockIdx.x * blockDim.x + threadIdx.x) / GM, n_mgroups = (gridDim.x * blockDim.x) / GM;
    const int mlane = Group<GM>::lane();
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int nthreads = gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31;
    const long long n_cells = c.g.n_cells;
    const int n_list = (int)c.a.counters[SSB_C_N_LIST];
    const int E = rc.epochs;

    // ---- prepare: priorities, epoch histogram ------------------------------------------------------
    for (int i = tid; i < E; i += nthreads) { rc.epoch_count[i] = 0; rc.epoch_fill[i] = 0; }
    for (long long i = tid; i < n_cells; i += nthreads) rc.mark[i] = 0xffffffffu;
    if (tid == 0) { rc.cnt[0] = 0; rc.cnt[1] = 0; rc.cnt[2] = 0; rc.cnt[3] = 0; c.a.counters[SSB_C_N_BORN] = 0; }
    grid.sync();
    for (int e = threadIdx.x; e < E; e += blockDim.x) sh_count[e] = 0;
    __syncthreads();
    for (int i = tid; i < n_list; i += nthreads) {
        const int s = c.a.order[i];
        const uint32_t pr = priority_of(rc.key, (uint32_t)c.a.id[s], c.p.id_bits);
        rc.prio[s] = pr;
        atomicAdd(&sh_count[pr >> rc.epoch_shift], 1);
    }
    __syncthreads();
    for (int e = threadIdx.x; e < E; e += blockDim.x) if (sh_count[e]) atomicAdd(&rc.epoch_count[e], sh_count[e]);
    grid.sync();
    if (blockIdx.x == 0) {   // exclusive scan of the epoch histogram (E <= 1024)
        if (threadIdx.x == 0) {
            int run = 0;
            for (int e = 0; e < E; e++) { rc.epoch_start[e] = run; run += rc.epoch_count[e]; }
            rc.epoch_start[E] = run;
        }
    }
    grid.sync();
    // scatter into epoch segments: one range reservation per block and epoch
    for (int e = threadIdx.x; e < E; e += blockDim.x) sh_count[e] = 0;
    __syncthreads();
    for (int i = tid; i < n_list; i += nthreads) atomicAdd(&sh_count[rc.prio[c.a.order[i]] >> rc.epoch_shift], 1);
    __syncthreads();
    for (int e = threadIdx.x; e < E; e += blockDim.x) {
        sh_base[e] = sh_count[e] ? rc.epoch_start[e] + atomicAdd(&rc.epoch_fill[e], sh_count[e]) : 0;
        sh_count[e] = 0;
    }
    __syncthreads();
    for (int i = tid; i < n_list; i += nthreads) {
        const int s = c.a.order[i];
        const int e = rc.prio[s] >> rc.epoch_shift;
        rc.bucketed[sh_base[e] + atomicAdd(&sh_count[e], 1)] = s;
    }
    grid.sync();

    // ---- commit rounds -------------------------------------------------------------------------------
    int32_t *cur = rc.list_a, *nxt = rc.list_b;
    int round = 0;
    for (;;) {
        const int n_carry = ((volatile int32_t *)rc.cnt)[round & 1];
        const int e0 = round < E ? rc.epoch_start[round] : 0;
        const int n_new = round < E ? rc.epoch_start[round + 1] - e0 : 0;
        const int n = n_carry + n_new;
        if (round >= E && n == 0) break;
        if (round >= MAX_ROUNDS) {               // cannot happen (the earliest active agent always commits): fail loudly
            if (tid == 0) c.a.counters[SSB_C_OVERFLOW] = 2;
            break;
        }
        if (round > 0 && round % ROUNDS_PER_TAG_EPOCH == 0) {   // tags exhausted: clear and restart
            for (long long i = tid; i < n_cells; i += nthreads) rc.mark[i] = 0xffffffffu;
            grid.sync();
        }
        int32_t *n_next = &rc.cnt[(round + 1) & 1], *n_ready = &rc.cnt[2 + (round & 1)];
        const bool tr = rc.trace && tid == 0 && round < rc.trace_rounds;
        if (tid == 0) { *n_next = 0; *n_ready = 0; }
        if (tr) { rc.trace[round * 6 + 0] = n; rc.trace[round * 6 + 2] = global_ns(); }
        const uint32_t tag = round_tag(c, round);
        const int32_t *fresh = rc.bucketed + e0;
        // MARK: one agent per lane group, lanes spread over the footprint cells
        for (int i = mgroup_id; i < n; i += n_mgroups) {
            const int entry = i < n_carry ? cur[i] : fresh[i - n_carry];
            const int s = entry & ~PHASE1_BIT;
            const int pos = c.a.pos[s];
            if (pos < 0) continue;                   // killed earlier this step: nothing to reserve
            const bool moved = (entry & PHASE1_BIT) != 0;
            const int k = footprint_dilation(c, s, moved);
            const int r = moved ? 0 : agent_range(c, s);
            group_mark_region<GM>(c, rc, pos, r, k, tag | rc.prio[s]);
        }
        grid.sync();
        if (tr) rc.trace[round * 6 + 3] = global_ns();
        // SELECT
        for (int i = mgroup_id; i < n; i += n_mgroups) {
            const int entry = i < n_carry ? cur[i] : fresh[i - n_carry];
            const int s = entry & ~PHASE1_BIT;
            const int pos = c.a.pos[s];
            bool ok = true;
            if (pos >= 0) {
                if (entry & PHASE1_BIT) ok = group_owns_region<GM>(c, rc, pos, 0, footprint_dilation(c, s, true), tag | rc.prio[s]);
                else ok = group_owns_region<GM>(c, rc, pos, agent_range(c, s), 0, tag | rc.prio[s]);
            }
            if (mlane == 0) {
                if (ok) rc.ready[atomicAdd(n_ready, 1)] = entry;
                else nxt[atomicAdd(n_next, 1)] = entry;
            }
        }
        grid.sync();
        const int nr = *(volatile int32_t *)n_ready;
        if (tr) { rc.trace[round * 6 + 1] = nr; rc.trace[round * 6 + 4] = global_ns(); }
        // EXECUTE: one transaction per lane group
        for (int i = warp_id; i < nr; i += n_warps) {
            const int entry = rc.ready[i];
            const int s = entry & ~PHASE1_BIT;
            AgentRec rec;
            rec.load(c, s);                                  // one latency level for the whole record
            RngStream rng{rc.key, (uint32_t)rec.id, 0u, 0u, nullptr};
            if (entry & PHASE1_BIT) {
                if (glane == 0) transaction_interact_phase<RngStream, false>(c, s, rec, rng);
            } else if (transaction_move_phase<GL, 4>(c, s, rec, rng, ws, MAXC_SCALABLE)) {
                // the interact part needs D_k(new cell), k from the now final holdings (never
                // larger than the k the marks were placed with); the marks of this round still
                // stand, so it runs right away if the agent owns that diamond too.  rec is the
                // same on every lane (broadcast from lane 0), so the branch is group-uniform.
                const int k = footprint_dilation_moved(c, rec);
                const bool now = k == 0 || group_owns_region<GL>(c, rc, rec.pos, 0, k, tag | rc.prio[s]);
                if (glane == 0) {
                    if (now) transaction_interact_phase<RngStream, false>(c, s, rec, rng);
                    else nxt[atomicAdd(n_next, 1)] = s | PHASE1_BIT;
                }
            }
            Group<GL>::sync();
        }
        grid.sync();
        if (tr) rc.trace[round * 6 + 5] = global_ns();
        int32_t *tmp = cur; cur = nxt; nxt = tmp;
        round++;
    }
    if (tid == 0) c.a.counters[SSB_C_ROUNDS] = round;
}

}  // namespace ssb
