// ssb_flow_core.cuh -- the static conflict DAG of one mode S agent step (host- and device-compilable part).
//
// Mode S defines a step as "agents act one at a time in ascending priority" (ssb_rng.cuh; the reference's
// random.shuffle(self.agents) + agent loop, sugarscape.py:400-407).  A transaction only touches its
// FOOTPRINT F = D_k(cross of radius r around the agent's cell at the start of the step), with k = 0 (move,
// combat, collect, metabolise: agent.py:568-584), 1 (tagging / trading: the von Neumann neighbours of the new
// cell, agent.py:556-566, 600-706) or 2 (reproduction: a mate's empty neighbours, agent.py:500-554).  Two agents
// whose footprints are disjoint commute, so ANY execution that orders every pair with intersecting
// footprints by priority equals the sequential one.  That partial order is a DAG that depends only on the
// positions, ranges and priorities at the start of the step:
//
//     A -> B   iff   priority(A) < priority(B)  and  dist_L1(cross_A, cross_B) <= k_A + k_B.
//
// The engine builds it ONCE per step (k_flow_build: one descriptor word per cell, staged tile by tile in
// shared memory) and then runs it as a dataflow (k_flow_execute: an agent is pushed to a ready queue when its
// last predecessor has finished) -- no commit rounds, no grid barriers, every agent marks nothing and runs
// exactly once.  Everything in this file is plain C++ so that tests/hostemu compiles the SAME predicate and
// the SAME tile scan with g++ and replays the golden fixtures through an adversarial topological order.
#pragma once
#include "ssb_rules.cuh"

namespace ssb {

// ---- descriptor of the agent standing on a cell at the start of the step (0 = empty) ---------------------
//   bit 63 valid | 62..36 list index (27 bits) | 35..10 priority (26 bits, = id_bits limit) | 9..6 r | 5..4 k
constexpr int FLOW_INDEX_BITS = 27;
constexpr unsigned long long FLOW_VALID = 1ull << 63;
__host__ __device__ __forceinline__ unsigned long long flow_pack(uint32_t index, uint32_t prio, int r, int k) {
    return FLOW_VALID | ((unsigned long long)index << 36) | ((unsigned long long)prio << 10) |
           ((unsigned long long)(r & 15) << 6) | ((unsigned long long)(k & 3) << 4);
}
__host__ __device__ __forceinline__ uint32_t flow_index(unsigned long long d) { return (uint32_t)(d >> 36) & ((1u << FLOW_INDEX_BITS) - 1u); }
__host__ __device__ __forceinline__ uint32_t flow_prio(unsigned long long d) { return (uint32_t)(d >> 10) & ((1u << 26) - 1u); }
__host__ __device__ __forceinline__ int flow_r(unsigned long long d) { return (int)(d >> 6) & 15; }
__host__ __device__ __forceinline__ int flow_k(unsigned long long d) { return (int)(d >> 4) & 3; }

// Dilation of the footprint required by the rule set, for agent s (agent.py:585-588), evaluated
// BEFORE the agent moves.  Reproduction (k = 2) is only possible if the agent can be fertile when
// it gets there: fertile age, and holdings that can still reach the starting endowment -- other
// agents can only LOWER them (mate costs), the agent's own move adds at most one cell's capacity
// (plus combat loot) and subtracts its metabolism.  Computed once per step: a value that was
// valid at the start of the step stays a superset later.
__device__ __forceinline__ int footprint_dilation(const DevCtx &c, int s) {
    int k = 0;
    if (c.p.flags & (SSB_F_TAGGING | SSB_F_TRADE)) k = 1;
    if ((c.p.flags & SSB_F_REPRO) && c.a.fert_factor[s] > 0 &&
        c.a.age[s] >= c.a.fert_age[s] && c.a.age[s] < c.a.infert_age[s]) {
        bool can = true;
        if (!(c.p.flags & SSB_F_TRADE)) {
            const double loot = c.a.aggression[s] > 0 ? c.p.max_combat_loot : 0.0;
            const double gain_s = (double)c.p.max_sugar_capacity + loot - (double)max(c.a.sugar_metab[s], 0);
            const double gain_p = (double)c.p.max_spice_capacity + loot - (double)max(c.a.spice_metab[s], 0);
            can = c.a.sugar[s] + gain_s >= c.a.start_sugar[s] && c.a.spice[s] + gain_p >= c.a.start_spice[s];
        }
        if (can) k = 2;
    }
    return k;
}

// largest dilation any agent of this rule set can have
__host__ __device__ __forceinline__ int flow_max_dilation(int flags) {
    return (flags & SSB_F_REPRO) ? 2 : ((flags & (SSB_F_TAGGING | SSB_F_TRADE)) ? 1 : 0);
}

__host__ __device__ __forceinline__ int flow_gap(int a, int r) { return a > r ? a - r : 0; }

// Do the footprints of A (range ra at the origin) and B (range rb at offset (dx, dy)) intersect?
// K = k_A + k_B.  The L1 distance between the two crosses is the smallest distance between an arm of A
// and an arm of B (axis-aligned segments): parallel arms, and A's north-south arm against B's east-west
// arm and vice versa.  Symmetric: conflict(dx, dy, ra, rb) == conflict(-dx, -dy, rb, ra).
__host__ __device__ __forceinline__ bool flow_conflict(int adx, int ady, int ra, int rb, int K) {
    const int s = ra + rb;
    return adx + flow_gap(ady, s) <= K || ady + flow_gap(adx, s) <= K ||
           flow_gap(adx, rb) + flow_gap(ady, ra) <= K || flow_gap(adx, ra) + flow_gap(ady, rb) <= K;
}

// Largest |dy| at column offset |dx| = adx for which flow_conflict can hold against ANY partner with
// range <= RB and dilation <= KB (-1: none).  Each bound is the predicate's term solved for ady with
// rb = RB, K = ka + KB (all four are monotone in rb and K).
__host__ __device__ __forceinline__ int flow_column_reach(int adx, int ra, int ka, int RB, int KB) {
    const int K = ka + KB;
    int h = -1;
    if (adx <= K) h = ra + RB + K - adx;
    const int b2 = K - flow_gap(adx, ra + RB);
    const int b3 = ra + K - flow_gap(adx, RB);
    const int b4 = RB + K - flow_gap(adx, ra);
    if (b2 > h) h = b2;
    if (b3 > h) h = b3;
    if (b4 > h) h = b4;
    return h;
}

// halo of a staged tile: no conflict reaches further than this in either axis
__host__ __device__ __forceinline__ int flow_halo(int RB, int KB) { return 2 * RB + 2 * KB; }

// ---- the staged tile ------------------------------------------------------------------------------------
// A tile of FLOW_TX x FLOW_TY cells plus a halo of R cells on every side, as descriptor words in column
// (x) major order like the grid, rows padded to `sh`; mask[col * 2 + w] holds the occupancy bits of rows
// 64 w .. 64 w + 63 of that column (sh <= 128).
constexpr int FLOW_TX = 64, FLOW_TY = 64;
constexpr int FLOW_MAX_HALO = 28;                            // 2 * halo + 1 <= 64: one column window is one 64-bit word
constexpr int FLOW_MAX_SH = FLOW_TY + 2 * FLOW_MAX_HALO;      // <= 128 rows: two 64-bit mask words per column

struct FlowTile {
    const unsigned long long *cell;      // [sw * sh]
    const unsigned long long *mask;      // [sw * 2]
    int sw, sh, halo;
};

// bits y0 .. y0 + n - 1 (n <= 64) of a column's 128-bit occupancy mask; y0 >= 0, y0 + n <= 128
__host__ __device__ __forceinline__ unsigned long long flow_mask_window(const unsigned long long *m, int y0, int n) {
    unsigned long long w;
    if (y0 >= 64) w = m[1] >> (y0 - 64);
    else if (y0 == 0) w = m[0];
    else w = (m[0] >> y0) | (m[1] << (64 - y0));
    return n >= 64 ? w : (w & ((1ull << n) - 1ull));
}

__host__ __device__ __forceinline__ int flow_ctz64(unsigned long long v) {
#if defined(__CUDA_ARCH__)
    return __ffsll((long long)v) - 1;
#else
    return __builtin_ctzll(v);
#endif
}

// Visit every agent B of the staged tile whose footprint intersects that of the agent A standing on
// staged cell (lx, ly): f(staged cell index of B, descriptor of B).  A itself is skipped.  RB / KB: the
// largest range / dilation of any agent in this step.
template <class F>
__host__ __device__ __forceinline__ void flow_scan_agent(const FlowTile &tile, int lx, int ly, unsigned long long da,
                                                        int RB, int KB, F f) {
    const int ra = flow_r(da), ka = flow_k(da);
    for (int dx = -tile.halo; dx <= tile.halo; dx++) {
        const int adx = dx < 0 ? -dx : dx;
        int h = flow_column_reach(adx, ra, ka, RB, KB);
        if (h < 0) continue;
        if (h > tile.halo) h = tile.halo;
        const int col = lx + dx, y0 = ly - h;
        unsigned long long bits = flow_mask_window(tile.mask + col * 2, y0, 2 * h + 1);
        if (dx == 0) bits &= ~(1ull << h);
        while (bits) {
            const int j = flow_ctz64(bits);
            bits &= bits - 1;
            const int cell = col * tile.sh + y0 + j;
            const unsigned long long db = tile.cell[cell];
            const int dy = j - h;
            if (flow_conflict(adx, dy < 0 ? -dy : dy, ra, flow_r(db), ka + flow_k(db))) f(cell, db);
        }
    }
}

}  // namespace ssb
