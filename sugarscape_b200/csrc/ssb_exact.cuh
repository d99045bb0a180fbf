// ssb_exact.cuh -- the ordered agent loop of mode E (reference sugarscape.py:400-407), as a device
// function so that the one-warp kernel (ssb_agent_step.cuh) and the host-compiled emulation of the
// device rule code (tests/hostemu, test infrastructure: G = 1, plain g++) run the SAME source.
#pragma once
#include "ssb_rules.cuh"

namespace ssb {

constexpr int MAXC_EXACT = 160;   // up to 4 * 40 candidate cells of a cross, or the 144 of a radial range of 6
static_assert(MAXC_EXACT == MAXC_EXACT_RULES, "ssb_rules.cuh sizes its ranking arrays for MAXC_EXACT candidates");

// random.shuffle(self.agents), then every agent of the list in order; the list grows while it is
// iterated: children are appended, visited and skipped (their lastMovedTimestep is this step).
// All G lanes of the (single) group call this; `mt` is the 625-word MT19937 state (index last).
template <int G, bool EXT = false>
__device__ void agent_step_exact_body(const DevCtx &c, uint32_t *mt, const WarpScratch &ws) {
    typedef Group<G> Gr;
    const int lane = Gr::lane();
    RngExact rng{mt};
    volatile unsigned long long *cn = (volatile unsigned long long *)c.a.counters;
    if (lane == 0) {
        cn[SSB_C_N_BORN] = 0;
        shuffle(rng, c.a.order, (int)cn[SSB_C_N_LIST]);        // random.shuffle(self.agents), sugarscape.py:400
        if constexpr (EXT) { if (has_disease(c)) disease_add_remaining(c); }       // self.addRemainingDiseases(), sugarscape.py:401
    }
    Gr::sync();
    if constexpr (EXT) {
        // robots next to the Bentham family: the agents were (re)configured since the last step -- every agent's stored
        // neighbourhood is rebuilt where it stands (sugarscape.py:265-267; the host marked the lists with -1)
        if (keeps_hoods(c) && lane == 0) {
            const long long n0 = (long long)cn[SSB_C_N_LIST];
            for (long long i = 0; i < n0; i++) {       // -1: after a replacement; -2: the initial population (before its diseases)
                const int mark = c.eth->hood_n[c.a.order[i]];
                if (mark == -1 || mark == -2) hood_store_here(c, c.a.order[i], mark == -2);
            }
        }
        Gr::sync();
    }
    for (long long i = 0;; i++) {
        const long long n_list = Gr::bcast((long long)cn[SSB_C_N_LIST], 0);
        if (i >= n_list) break;
        const int s = c.a.order[i];
        AgentRec rec;
        rec.load(c, s);
        const bool go_on = transaction_move_phase<G, 4, RngExact, EXT>(c, s, rec, rng, ws, MAXC_EXACT);
        if (go_on && lane == 0) transaction_interact_phase<RngExact, true, EXT>(c, s, rec, rng);
        __threadfence_block();
        Gr::sync();
    }
    if constexpr (EXT) {
        if (group_dynamic(c)) group_resplit_neighbors<G>(c);
    }
}

// Decision-model reductions of the log (sugarscape.py:1141-1160, 1218-1222): one lane, list order (the
// difference sum is order-sensitive fp64); the dead of this step count towards the move optimality --
// doTimestep(t) ran for every one of them.
__device__ inline void ethics_stats_body(const DevCtx &c, ssb_ethics_stats_t *out, int pass = 0) {
    ssb_ethics_stats_t r;
    r.sum_selfishness = 0; r.optimal_moves = 0; r.moves = 0; r.sum_move_rank = 0; r.sum_move_diff = 0;
    const long long n = c.a.counters[SSB_C_N_LIST], n_dead = c.a.counters[SSB_C_N_DEAD];
    for (long long i = 0; i < n; i++) {
        const int s = c.a.order[i];
        if (!in_pass(c, s, pass)) continue;
        r.sum_selfishness += c.eth->selfishness[s];
        r.optimal_moves += c.eth->last_move_optimal[s]; r.moves += 1;
        r.sum_move_rank += c.eth->move_rank[s]; r.sum_move_diff += c.eth->move_diff[s];
    }
    for (long long i = 0; i < n_dead; i++)
        if (in_pass(c, c.a.dead_list[i], pass)) { r.optimal_moves += c.eth->last_move_optimal[c.a.dead_list[i]]; r.moves += 1; }
    *out = r;
}

// The experimental-group split of the log: the group (pass 1) and its complement (pass 2) in full, then the
// everybody pass (0), which only adds the neighbour sums here -- its statistics come from the regular reductions.
// The order matters: passes 1 and 2 reset the interaction counters of their members.
__device__ inline void group_stats_body(const DevCtx &c, ssb_group_stats_t out[3]) {
    for (int pass = 1; pass <= 2; pass++) {
        stats_pass_body(c, pass, &out[pass].stats);
        if (c.eth != nullptr) ethics_stats_body(c, &out[pass].ethics, pass);
        ext_stats_body(c, &out[pass].ext, pass);
        group_counters_body(c, pass, &out[pass]);
    }
    group_counters_body(c, 0, &out[0]);
}

}  // namespace ssb
