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

#include "ssb_exact.cuh"
#include "ssb_flow_core.cuh"      // footprint_dilation (shared with the dataflow scheduler)

namespace ssb {
namespace cg = cooperative_groups;

// ---------------------------------------------------------------------------------------------
// mode E
// ---------------------------------------------------------------------------------------------
// EXT = false: the plain Agent of the shipped book examples; EXT = true: an ABI 2 extension is bound
// (decision models) -- a separate instantiation, so the plain kernel's code does not change.
template <bool EXT>
__global__ void __launch_bounds__(32, 1) k_agent_step_exact(DevCtx c, uint32_t *mt_global) {
    __shared__ uint32_t mt[625];
    __shared__ __align__(8) unsigned char scratch_mem[warp_scratch_bytes(MAXC_EXACT)];
    const int lane = threadIdx.x;
    for (int i = lane; i < 625; i += 32) mt[i] = mt_global[i];
    __syncwarp();
    agent_step_exact_body<32, EXT>(c, mt, make_warp_scratch(scratch_mem, MAXC_EXACT));
    __syncwarp();
    for (int i = lane; i < 625; i += 32) mt_global[i] = mt[i];
}

// decision-model reductions of the log, list order (mode E)
__global__ void k_ethics_stats(DevCtx c, ssb_ethics_stats_t *out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) ethics_stats_body(c, out);
}
// the experimental-group split of the log (mode E)
__global__ void k_group_stats(DevCtx c, ssb_group_stats_t *out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) group_stats_body(c, out);
}
// lending / friends reductions of the log, list order (mode E)
__global__ void k_ext_stats(DevCtx c, ssb_ext_stats_t *out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) ext_stats_body(c, out);
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

// One entry of the active lists: everything MARK and SELECT need, so that those phases read only
// the (coalesced) list and the marks -- no per-agent gathers.
//   x = slot | PHASE1_BIT, y = cell the footprint is centred on, z = priority, w = r | k << 8
struct RoundCtx {
    uint32_t *mark;          // one word per mark cell (cells coarsened by 2^mark_shift per side)
    int32_t mark_shift, wc, hc;
    int4 *bucketed;          // the agent list grouped by epoch
    int4 *list_a, *list_b, *ready;
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
// Marks live on a grid coarsened by 2^sh cells per side (sh = 0: one mark per cell); a coarse
// column takes the largest interval of the fine columns it contains -- a superset is always safe.
struct Footprint {
    int x, y, rx, ry, k;
    __device__ __forceinline__ void init(const DevCtx &c, int pos, int r, int k_) {
        const int H = c.p.height;
        x = pos / H; y = pos - x * H;
        rx = min(r, c.max_dx); ry = min(r, c.max_dy); k = k_;
    }
    __device__ __forceinline__ int half_height(int i) const {
        const int ai = abs(i);
        const int ns = ai <= k ? ry + k - ai : -1;
        const int ew = ai <= rx ? k : k - (ai - rx);
        return max(ns, ew);
    }
    // largest half height among the fine columns x + i that fall into coarse column cx
    __device__ __forceinline__ int coarse_half_height(int cx, int sh) const {
        const int c0 = cx * (1 << sh);
        const int lo = max(c0 - x, -(rx + k)), hi = min(c0 + (1 << sh) - 1 - x, rx + k);
        // h(i) is non-increasing in |i|: the maximum sits at the offset closest to 0
        const int i = lo > 0 ? lo : (hi < 0 ? hi : 0);
        return half_height(i);
    }
};

__device__ __forceinline__ uint32_t round_tag(const DevCtx &c, int round) {
    constexpr int per_epoch = (1 << TAG_BITS) - 1;          // tag 63 is the cleared value
    return (uint32_t)(per_epoch - 1 - (round % per_epoch)) << c.p.id_bits;
}
constexpr int ROUNDS_PER_TAG_EPOCH = (1 << TAG_BITS) - 1;

// Visit the mark cells of a footprint with a lane group of G lanes: one (coarse) x-column per
// iteration, the lanes cover its rows.  f(mark index) is called once per lane and cell.
template <int G, class F>
__device__ __forceinline__ void for_each_mark_cell(const DevCtx &c, const RoundCtx &rc, int pos, int r, int k, F f) {
    const int lane = Group<G>::lane(), sh = rc.mark_shift;
    Footprint fp;
    fp.init(c, pos, r, k);
    const int cx0 = (fp.x - fp.rx - fp.k) >> sh, cx1 = (fp.x + fp.rx + fp.k) >> sh;     // >> floors
    for (int cx = cx0; cx <= cx1; cx++) {
        const int h = fp.coarse_half_height(cx, sh);
        const int cy0 = (fp.y - h) >> sh, cy1 = (fp.y + h) >> sh;
        const size_t col = (size_t)wrap1(cx, rc.wc) * rc.hc;
        for (int cy = cy0 + lane; cy <= cy1; cy += G) f(col + wrap1(cy, rc.hc));
    }
}

template <int G>
__device__ __forceinline__ void group_mark_region(const DevCtx &c, const RoundCtx &rc, int pos, int r, int k, uint32_t v) {
    // sharded: the mark field is NVLink-stitched memory that several GPUs update concurrently -- device-scope atomics of
    // different GPUs are not morally strong against each other, so the reservation needs system scope there
    if (c.sh.world > 1) for_each_mark_cell<G>(c, rc, pos, r, k, [&](size_t m) { atomicMin_system(rc.mark + m, v); });
    else for_each_mark_cell<G>(c, rc, pos, r, k, [&](size_t m) { atomicMin(rc.mark + m, v); });
}

// true (on every lane of the group) iff every mark cell of the region holds v
template <int G>
__device__ __forceinline__ bool group_owns_region(const DevCtx &c, const RoundCtx &rc, int pos, int r, int k, uint32_t v) {
    unsigned bad = 0;
    for_each_mark_cell<G>(c, rc, pos, r, k, [&](size_t m) { bad |= __ldcg(rc.mark + m) ^ v; });
    return !Group<G>::any(bad != 0);
}

// One persistent cooperative kernel per step.
__global__ void __launch_bounds__(AGENT_THREADS, 2) k_agent_step_scalable(DevCtx c, RoundCtx rc) {
    cg::grid_group grid = cg::this_grid();
    __shared__ int32_t sh_count[MAX_EPOCHS], sh_base[MAX_EPOCHS];
    extern __shared__ __align__(8) unsigned char scratch_mem[];
    const WarpScratch ws = make_warp_scratch(scratch_mem + (threadIdx.x / GL) * warp_scratch_bytes(MAXC_SCALABLE), MAXC_SCALABLE);
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int nthreads = gridDim.x * blockDim.x;
    const int xgroup_id = tid / GL, n_xgroups = nthreads / GL, xlane = Group<GL>::lane();   // EXECUTE groups
    const int mgroup_id = tid / GM, n_mgroups = nthreads / GM, mlane = Group<GM>::lane();   // MARK / SELECT groups
    const long long n_marks = (long long)rc.wc * rc.hc;
    const int n_list = (int)c.a.counters[SSB_C_N_LIST];
    const int E = rc.epochs;

    // ---- prepare: priorities, footprint parameters, epoch buckets ---------------------------------
    for (int i = tid; i < E; i += nthreads) { rc.epoch_count[i] = 0; rc.epoch_fill[i] = 0; }
    const bool sharded = c.sh.world > 1;       // x-slab sharding: this rank clears its part of the marks
    const long long mark_lo = sharded ? c.sh.mark_lo : 0, mark_hi = sharded ? c.sh.mark_hi : n_marks;
    for (long long i = mark_lo + tid; i < mark_hi; i += nthreads) rc.mark[i] = 0xffffffffu;
    if (tid == 0) { rc.cnt[0] = 0; rc.cnt[1] = 0; rc.cnt[2] = 0; rc.cnt[3] = 0; c.a.counters[SSB_C_N_BORN] = 0; }
    for (int e = threadIdx.x; e < E; e += blockDim.x) sh_count[e] = 0;
    grid.sync();
    // pass A: histogram of epochs (the per-block work assignment is repeated exactly in pass B)
    for (int i = tid; i < n_list; i += nthreads) {
        const uint32_t pr = priority_of(rc.key, (uint32_t)c.a.id[c.a.order[i]], c.p.id_bits);
        atomicAdd(&sh_count[pr >> rc.epoch_shift], 1);
    }
    __syncthreads();
    for (int e = threadIdx.x; e < E; e += blockDim.x) if (sh_count[e]) atomicAdd(&rc.epoch_count[e], sh_count[e]);
    grid.sync();
    if (tid == 0) {   // exclusive scan of the epoch histogram (E <= 1024)
        int run = 0;
        for (int e = 0; e < E; e++) { rc.epoch_start[e] = run; run += rc.epoch_count[e]; }
        rc.epoch_start[E] = run;
    }
    grid.sync();
    // pass B: one range reservation per block and epoch, then scatter the entries
    for (int e = threadIdx.x; e < E; e += blockDim.x) {
        sh_base[e] = sh_count[e] ? rc.epoch_start[e] + atomicAdd(&rc.epoch_fill[e], sh_count[e]) : 0;
        sh_count[e] = 0;
    }
    __syncthreads();
    for (int i = tid; i < n_list; i += nthreads) {
        const int s = c.a.order[i];
        const uint32_t pr = priority_of(rc.key, (uint32_t)c.a.id[s], c.p.id_bits);
        const int e = pr >> rc.epoch_shift;
        const int r = min(min(max(c.a.vision[s], 0), max(c.a.movement[s], 0)), max(c.max_dx, c.max_dy));
        rc.bucketed[sh_base[e] + atomicAdd(&sh_count[e], 1)] =
            make_int4(s, c.a.pos[s], (int)pr, r | (footprint_dilation(c, s) << 8));
    }
    // every rank has cleared its marks before any rank places one
    bool box_ok = shard_grid_sync(grid, c.sh, nullptr, nullptr);

    // ---- commit rounds -------------------------------------------------------------------------------
    int4 *cur = rc.list_a, *nxt = rc.list_b;
    int round = 0;
    unsigned long long box_carry = 0;            // uncommitted agents carried over, summed over the ranks
    for (;;) {
        const int n_carry = ((volatile int32_t *)rc.cnt)[round & 1];
        const int e0 = round < E ? rc.epoch_start[round] : 0;
        const int n_new = round < E ? rc.epoch_start[round + 1] - e0 : 0;
        const int n = n_carry + n_new;
        if (round >= E && box_carry == 0) break;
        if (round >= MAX_ROUNDS || !box_ok) {    // cannot happen (the earliest active agent always commits) / lost peer: fail loudly
            if (tid == 0) c.a.counters[SSB_C_OVERFLOW] = box_ok ? 2 : 7;
            break;
        }
        if (round > 0 && round % ROUNDS_PER_TAG_EPOCH == 0) {   // tags exhausted: clear and restart
            for (long long i = mark_lo + tid; i < mark_hi; i += nthreads) rc.mark[i] = 0xffffffffu;
            box_ok = shard_grid_sync(grid, c.sh, nullptr, nullptr);
        }
        int32_t *n_next = &rc.cnt[(round + 1) & 1], *n_ready = &rc.cnt[2 + (round & 1)];
        const bool tr = rc.trace && tid == 0 && round < rc.trace_rounds;
        if (tid == 0) { *n_next = 0; *n_ready = 0; }
        if (tr) { rc.trace[round * 6 + 0] = n; rc.trace[round * 6 + 2] = global_ns(); }
        const uint32_t tag = round_tag(c, round);
        const int4 *fresh = rc.bucketed + e0;
        // MARK: one agent per lane group, lanes spread over the footprint cells
        for (int i = mgroup_id; i < n; i += n_mgroups) {
            const int4 en = i < n_carry ? cur[i] : fresh[i - n_carry];
            const bool moved = (en.x & PHASE1_BIT) != 0;
            group_mark_region<GM>(c, rc, en.y, moved ? 0 : (en.w & 255), en.w >> 8, tag | (uint32_t)en.z);
        }
        box_ok = shard_grid_sync(grid, c.sh, nullptr, nullptr) && box_ok;     // all marks of the box are placed
        if (tr) rc.trace[round * 6 + 3] = global_ns();
        // SELECT: phase 0 needs the cross + own cell (k = 0), phase 1 the diamond D_k(new cell)
        for (int i = mgroup_id; i < n; i += n_mgroups) {
            const int4 en = i < n_carry ? cur[i] : fresh[i - n_carry];
            const bool moved = (en.x & PHASE1_BIT) != 0;
            const bool ok = group_owns_region<GM>(c, rc, en.y, moved ? 0 : (en.w & 255), moved ? (en.w >> 8) : 0,
                                                  tag | (uint32_t)en.z);
            if (mlane == 0) {
                if (ok) rc.ready[atomicAdd(n_ready, 1)] = en;
                else nxt[atomicAdd(n_next, 1)] = en;
            }
        }
        grid.sync();
        const int nr = *(volatile int32_t *)n_ready;
        if (tr) { rc.trace[round * 6 + 1] = nr; rc.trace[round * 6 + 4] = global_ns(); }
        // EXECUTE: one transaction per lane group
        for (int i = xgroup_id; i < nr; i += n_xgroups) {
            const int4 en = rc.ready[i];
            const int s = en.x & ~PHASE1_BIT;
            AgentRec rec;
            rec.load(c, s);                                  // one latency level for the whole record
            RngStream rng{rc.key, (uint32_t)rec.id, 0u, 0u, nullptr};
            if (en.x & PHASE1_BIT) {
                if (xlane == 0) transaction_interact_phase<RngStream, false>(c, s, rec, rng);
            } else if (transaction_move_phase<GL, 4>(c, s, rec, rng, ws, MAXC_SCALABLE)) {
                // the interact part needs D_k(new cell), k from the now final holdings (never
                // larger than the k the marks were placed with); the marks of this round still
                // stand, so it runs right away if the agent owns that diamond too.  rec is the
                // same on every lane (broadcast from lane 0), so the branch is group-uniform.
                const int k = min(footprint_dilation_moved(c, rec), en.w >> 8);
                const bool now = k == 0 || group_owns_region<GL>(c, rc, rec.pos, 0, k, tag | (uint32_t)en.z);
                if (xlane == 0) {
                    if (now) transaction_interact_phase<RngStream, false>(c, s, rec, rng);
                    else nxt[atomicAdd(n_next, 1)] = make_int4(s | PHASE1_BIT, rec.pos, en.z, k << 8);
                }
            }
            Group<GL>::sync();
        }
        // all transactions of the box are done (their marks may be overwritten now); the carry-over counts add up
        box_ok = shard_grid_sync(grid, c.sh, n_next, &box_carry) && box_ok;
        if (tr) rc.trace[round * 6 + 5] = global_ns();
        int4 *tmp = cur; cur = nxt; nxt = tmp;
        round++;
    }
    if (tid == 0) c.a.counters[SSB_C_ROUNDS] = round;
}

}  // namespace ssb
