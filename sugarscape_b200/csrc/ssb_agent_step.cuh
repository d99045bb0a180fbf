// ssb_agent_step.cuh -- the agent loop of Sugarscape.doTimestep (reference sugarscape.py:400-407).
//
// Mode E (exact): one CTA, agents strictly in list order, the CPython MT19937 stream in shared
//   memory.  Data-dependent consumption of the single global stream makes exact replay
//   inherently ordered; this mode serves the shipped 50x50
//   examples.
//
// Mode S (scalable): the step is DEFINED as "agents act one at a time in ascending priority"
//   (priority = keyed bijection of the agent id, ssb_rng.cuh).  The kernel below reproduces that
//   sequential result in parallel with deterministic reservations:
//     round k:  MARK    every uncommitted agent atomicMin's (tag_k | priority) onto every mark
//                       cell of its footprint F (all cells its transaction may read or write),
//               SELECT  an agent whose own value survived on ALL of F is first among every
//                       uncommitted agent it could interact with -> ready,
//               EXECUTE ready agents run their whole transaction concurrently (their footprints
//                       are pairwise disjoint), everyone else retries next round.
//   Two agents whose footprints intersect are therefore always committed in priority order, which
//   is the sequential semantics.  Marks carry a descending round tag in their high bits, so stale
//   marks of earlier rounds lose every atomicMin and never need clearing inside a step.
//   Footprint: the cross of radius r around the agent (cells it may inspect / move to) dilated by
//   k = 0 (movement / combat only), 1 (tagging or trading touch the 4 neighbours of the new cell)
//   or 2 (reproduction reads the mate's empty neighbours and places a child there); marks live on
//   a grid coarsened by 2^mark_shift cells per side (a superset is always safe).
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
    for (int i = threadIdx.x; i < 625; i += 32) mt[i] = mt_global[i];
    __syncwarp();
    if (threadIdx.x == 0) {
        RngExact rng{mt};
        unsigned long long *cn = (unsigned long long *)c.a.counters;
        cn[SSB_C_N_BORN] = 0;
        // random.shuffle(self.agents), sugarscape.py:400
        shuffle(rng, c.a.order, (int)cn[SSB_C_N_LIST]);
        // the list grows while it is iterated: children are appended, visited and skipped
        for (long long i = 0; i < (long long)((volatile unsigned long long *)cn)[SSB_C_N_LIST]; i++)
            agent_transaction<RngExact, true, MAXC_EXACT>(c, c.a.order[i], rng);
    }
    __syncwarp();
    for (int i = threadIdx.x; i < 625; i += 32) mt_global[i] = mt[i];
}

// ---------------------------------------------------------------------------------------------
// mode S
// ---------------------------------------------------------------------------------------------
constexpr int MAXC_SCALABLE = 32;     // 4 * 8 candidate cells; the host refuses larger ranges
constexpr int TAG_BITS = 6;

struct RoundCtx {
    uint32_t *mark;          // coarse mark grid, wc * hc words
    int32_t mark_shift, wc, hc;
    uint32_t *prio;          // per slot, this step's priority
    int32_t *list_a, *list_b, *ready;
    int32_t *cnt;            // [0..1] list lengths (ping-pong), [2..3] ready counts (ping-pong)
    uint64_t key;
    unsigned long long *trace;   // optional: per round {n_active, n_ready, ns after mark, select, execute}
    int32_t trace_rounds;
};

__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// dilation of the footprint required by the rule set, for agent s (agent.py:585-588)
__device__ __forceinline__ int footprint_dilation(const DevCtx &c, int s) {
    int k = 0;
    if (c.p.flags & (SSB_F_TAGGING | SSB_F_TRADE)) k = 1;
    if ((c.p.flags & SSB_F_REPRO) && c.a.fert_factor[s] > 0 &&
        c.a.age[s] >= c.a.fert_age[s] && c.a.age[s] < c.a.infert_age[s]) k = 2;
    return k;
}

// Visit every coarse mark cell covering the two dilated arms of the cross, as rectangles:
// horizontal band [x-rx-k, x+rx+k] x [y-k, y+k], vertical band [x-k, x+k] x [y-ry-k, y+ry+k]
// (only its parts above and below the horizontal band).  A cell may be visited twice on tiny
// maps (harmless: atomicMin / equality are idempotent).  f returns false to stop early.
template <class F>
__device__ __forceinline__ bool for_each_mark(const DevCtx &c, const RoundCtx &rc, int pos, int r, int k, F f) {
    const int H = c.p.height, sh = rc.mark_shift;
    const int x = pos / H, y = pos - x * H;
    const int rx = min(r, c.max_dx), ry = min(r, c.max_dy);
    const int ax0 = (x - rx - k) >> sh, ax1 = min((x + rx + k) >> sh, ax0 + rc.wc - 1);   // >> floors
    const int ay0 = (y - k) >> sh, ay1 = min((y + k) >> sh, ay0 + rc.hc - 1);
    for (int cx = ax0; cx <= ax1; cx++) {
        const int base = wrapi(cx, rc.wc) * rc.hc;
        for (int cy = ay0; cy <= ay1; cy++)
            if (!f(base + wrapi(cy, rc.hc))) return false;
    }
    const int bx0 = (x - k) >> sh, bx1 = min((x + k) >> sh, bx0 + rc.wc - 1);
    const int by0 = max((y - ry - k) >> sh, ay0 - rc.hc), by1 = min((y + ry + k) >> sh, ay1 + rc.hc);
    for (int cx = bx0; cx <= bx1; cx++) {
        const int base = wrapi(cx, rc.wc) * rc.hc;
        for (int cy = by0; cy < ay0; cy++)
            if (!f(base + wrapi(cy, rc.hc))) return false;
        for (int cy = ay1 + 1; cy <= by1; cy++)
            if (!f(base + wrapi(cy, rc.hc))) return false;
    }
    return true;
}

__device__ __forceinline__ void phase_prepare(const DevCtx &c, const RoundCtx &rc, int tid, int nthreads) {
    const int n = (int)c.a.counters[SSB_C_N_LIST];
    for (int i = tid; i < n; i += nthreads) {
        const int s = c.a.order[i];
        rc.prio[s] = priority_of(rc.key, (uint32_t)c.a.id[s], c.p.id_bits);
        rc.list_a[i] = s;
    }
    if (tid == 0) {
        rc.cnt[0] = n; rc.cnt[1] = 0; rc.cnt[2] = 0; rc.cnt[3] = 0;
        c.a.counters[SSB_C_N_BORN] = 0;
    }
}

constexpr int ROUNDS_PER_EPOCH = (1 << TAG_BITS) - 1;   // tag 63 is the cleared value

__device__ __forceinline__ uint32_t round_tag(const DevCtx &c, int round) {
    return (uint32_t)(ROUNDS_PER_EPOCH - 1 - (round % ROUNDS_PER_EPOCH)) << c.p.id_bits;
}

__device__ __forceinline__ void phase_mark(const DevCtx &c, const RoundCtx &rc, const int32_t *list, int n,
                                           int round, int tid, int nthreads) {
    const uint32_t tag = round_tag(c, round);
    for (int i = tid; i < n; i += nthreads) {
        const int s = list[i];
        const int pos = c.a.pos[s];
        if (pos < 0) continue;                       // killed earlier this step: nothing to reserve
        const uint32_t v = tag | rc.prio[s];
        for_each_mark(c, rc, pos, agent_range(c, s), footprint_dilation(c, s),
                      [&](int m) { atomicMin(&rc.mark[m], v); return true; });
    }
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

__device__ __forceinline__ void phase_select(const DevCtx &c, const RoundCtx &rc, const int32_t *list, int n,
                                             int32_t *next, int32_t *n_next, int32_t *n_ready,
                                             int round, int tid, int nthreads) {
    const uint32_t tag = round_tag(c, round);
    const int lane = threadIdx.x & 31;
    for (int base = tid - lane; base < n; base += nthreads) {     // warp-uniform trip count
        const int i = base + lane;
        int s = -1;
        bool ok = false;
        if (i < n) {
            s = list[i];
            const int pos = c.a.pos[s];
            ok = true;
            if (pos >= 0) {
                const uint32_t v = tag | rc.prio[s];
                ok = for_each_mark(c, rc, pos, agent_range(c, s), footprint_dilation(c, s),
                                   [&](int m) { return rc.mark[m] == v; });
            }
        }
        __syncwarp();
        warp_append(i < n && ok, rc.ready, n_ready, s);
        warp_append(i < n && !ok, next, n_next, s);
    }
}

__device__ __forceinline__ void phase_execute(const DevCtx &c, const RoundCtx &rc, int n_ready, int tid, int nthreads) {
    for (int i = tid; i < n_ready; i += nthreads) {
        const int s = rc.ready[i];
        RngStream rng{rc.key, (uint32_t)c.a.id[s], 0u};
        agent_transaction<RngStream, false, MAXC_SCALABLE>(c, s, rng);
    }
}

// One persistent cooperative kernel per step: prepare, then commit rounds until every agent of
// the list has acted.  Per round: MARK | sync | SELECT | sync | EXECUTE | sync.
__global__ void __launch_bounds__(256) k_agent_step_scalable(DevCtx c, RoundCtx rc) {
    cg::grid_group grid = cg::this_grid();
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int nthreads = gridDim.x * blockDim.x;
    const int n_marks = rc.wc * rc.hc;
    phase_prepare(c, rc, tid, nthreads);
    for (int i = tid; i < n_marks; i += nthreads) rc.mark[i] = 0xffffffffu;
    grid.sync();
    int32_t *cur = rc.list_a, *nxt = rc.list_b;
    int round = 0;
    for (;;) {
        const int n = ((volatile int32_t *)rc.cnt)[round & 1];
        if (n == 0) break;
        if (round > 0 && round % ROUNDS_PER_EPOCH == 0) {   // tags exhausted: clear and restart
            for (int i = tid; i < n_marks; i += nthreads) rc.mark[i] = 0xffffffffu;
            grid.sync();
        }
        int32_t *n_next = &rc.cnt[(round + 1) & 1], *n_ready = &rc.cnt[2 + (round & 1)];
        if (tid == 0) { *n_next = 0; *n_ready = 0; }
        const bool tr = rc.trace && tid == 0 && round < rc.trace_rounds;
        if (tr) { rc.trace[round * 6 + 0] = n; rc.trace[round * 6 + 2] = global_ns(); }
        phase_mark(c, rc, cur, n, round, tid, nthreads);
        grid.sync();
        if (tr) rc.trace[round * 6 + 3] = global_ns();
        phase_select(c, rc, cur, n, nxt, n_next, n_ready, round, tid, nthreads);
        grid.sync();
        if (tr) { rc.trace[round * 6 + 1] = *(volatile int32_t *)n_ready; rc.trace[round * 6 + 4] = global_ns(); }
        phase_execute(c, rc, *(volatile int32_t *)n_ready, tid, nthreads);
        grid.sync();
        if (tr) rc.trace[round * 6 + 5] = global_ns();
        int32_t *tmp = cur; cur = nxt; nxt = tmp;
        round++;
    }
    if (tid == 0) c.a.counters[SSB_C_ROUNDS] = round;
}

}  // namespace ssb
