// ssb_exact.cuh -- the ordered agent loop of mode E (reference sugarscape.py:400-407), as a device
// function so that the one-warp kernel (ssb_agent_step.cuh) and the host-compiled emulation of the
// device rule code (tests/hostemu, test infrastructure: G = 1, plain g++) run the SAME source.
#pragma once
#include "ssb_rules.cuh"

namespace ssb {

constexpr int MAXC_EXACT = 128;   // up to 4 * 32 candidate cells

// random.shuffle(self.agents), then every agent of the list in order; the list grows while it is
// iterated: children are appended, visited and skipped (their lastMovedTimestep is this step).
// All G lanes of the (single) group call this; `mt` is the 625-word MT19937 state (index last).
template <int G>
__device__ void agent_step_exact_body(const DevCtx &c, uint32_t *mt, const WarpScratch &ws) {
    typedef Group<G> Gr;
    const int lane = Gr::lane();
    RngExact rng{mt};
    volatile unsigned long long *cn = (volatile unsigned long long *)c.a.counters;
    if (lane == 0) {
        cn[SSB_C_N_BORN] = 0;
        shuffle(rng, c.a.order, (int)cn[SSB_C_N_LIST]);        // random.shuffle(self.agents), sugarscape.py:400
    }
    Gr::sync();
    for (long long i = 0;; i++) {
        const long long n_list = Gr::bcast((long long)cn[SSB_C_N_LIST], 0);
        if (i >= n_list) break;
        const int s = c.a.order[i];
        AgentRec rec;
        rec.load(c, s);
        const bool go_on = transaction_move_phase<G, 4>(c, s, rec, rng, ws, MAXC_EXACT);
        if (go_on && lane == 0) transaction_interact_phase<RngExact, true>(c, s, rec, rng);
        __threadfence_block();
        Gr::sync();
    }
}

}  // namespace ssb
