// ssb_lean.cuh -- the agent transaction of mode S on packed 32-byte records, ONE THREAD PER AGENT.
//
// The generic transaction (ssb_rules.cuh) walks ~17 struct-of-array fields per agent and runs its scalar part
// on one lane of a lane group; on a B200 it costs 40-90 us per agent and 128 registers (25 % occupancy).  For
// the rule sets whose transaction stays inside the cross (no tagging / trading / reproduction: footprint
// dilation k = 0 -- S2's combat_limited rules, the growback / pollution / combat / replacement examples) this
// file restates Agent.doTimestep (reference agent.py:568-598) on three sector-sized records per slot:
//
//   LeanA  read + written  holdings, position, tribe, cause of death, lookahead      (what neighbours read)
//   LeanB  read only       id, age, max age, metabolisms, range, aggression          (endowments)
//   LeanC  written only    lastSugar / lastSpice, neighbourhood and valid-move counts (by-products for the log)
//
// k_lean_pack fills A and B from the ABI's arrays in slot order (coalesced) together with the DAG descriptors,
// k_lean_unpack writes the mutable fields back (and evaluates happiness, agent.py:1522-1530) -- the ABI arrays
// stay the canonical state between steps.  The candidate cells (occupant + sugar of the 4 r cells of the cross
// and of the own cell) are staged in shared memory with cp.async, issued BEFORE the Fisher-Yates shuffle of
// agent.py:1400-1401 so that the loads fly while the random words are computed.
//
// Everything up to the kernels is plain C++: tests/hostemu compiles the same transaction with g++.
#pragma once
#include "ssb_flow_core.cuh"

namespace ssb {

enum { LEAN_COD_MASK = 15, LEAN_ACTED = 16, LEAN_RANGED = 32, LEAN_AGED = 64, LEAN_PACKED = 128, LEAN_SKIP = 256,
       LEAN_RANGE_SHIFT = 12 };     // bits 12..15 of LeanA.state: the agent's range r (<= 8)

struct __align__(32) LeanA {
    double sugar, spice, lookahead;
    int32_t pos;
    int16_t tribe;
    uint16_t state;            // cause of death | LEAN_* flags | range << LEAN_RANGE_SHIFT
};
struct __align__(32) LeanB {
    int32_t id, age, max_age, sugar_metab, spice_metab;
    int32_t start;             // the cell the agent stands on at the start of the step (its DAG descriptor lives there)
    double aggression;
};
struct __align__(32) LeanC {
    double last_sugar, last_spice;
    int32_t n_neighborhood, n_valid_moves;
    int32_t pad[2];
};
static_assert(sizeof(LeanA) == 32 && sizeof(LeanB) == 32 && sizeof(LeanC) == 32, "one 32-byte sector per record");

struct LeanCtx {
    LeanA *a;
    LeanB *b;
    LeanC *c;
    int32_t *expect;           // [n_queues] entries each ready queue will receive in this step
    unsigned long long *live;  // [0] live slots seen by the pack pass (must equal the list length)
    int32_t maxc;              // 4 * largest range: candidate cells staged per agent
};

// The rule sets the lean transaction covers (anything else runs the generic transaction): scalable RNG mode,
// footprint dilation 0, no tag preferences, and a map on which every distance of the cross has four distinct
// cells (environment.py:111-122 merges them when 2 d equals the map side).
__host__ __device__ inline bool lean_supported(const ssb_params_t &p) {
    if (p.rng_mode != SSB_RNG_SCALABLE) return false;
    if (p.flags & (SSB_F_TAGGING | SSB_F_TRADE | SSB_F_REPRO | SSB_F_TAGPREF)) return false;
    const int side = p.width < p.height ? p.width : p.height;
    return p.max_agent_range >= 0 && 2 * (p.max_agent_range < 8 ? p.max_agent_range : 8) + 1 < side;
}

// Per-thread staging arrays (shared memory on the device, element i of this thread at [i * stride]).
struct LeanStage {
    int32_t *occ, *sugar;      // [maxc + 1]: the candidates in enumeration order, then the own cell
    uint8_t *perm, *rank;      // [maxc]
    int stride;
};

__device__ __forceinline__ void lean_stage_load(int32_t *dst, const int32_t *src) {
#if defined(__CUDA_ARCH__)
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(src) : "memory");
#else
    *dst = *src;
#endif
}
__device__ __forceinline__ void lean_stage_wait() {
#if defined(__CUDA_ARCH__)
    asm volatile("cp.async.wait_all;" ::: "memory");
#endif
}

// Candidate i of Agent.findCellsInRange in the insertion order of Environment.findCardinalCellRanges
// (environment.py:111-122): distance d = i / 4 + 1; within a distance the order is
// W N E S in the interior, N E S W if the west cell wraps (x < d), W E S N if the north cell wraps (y < d),
// E S N W if both.  (2 d < map side: the four cells are distinct, lean_supported.)
__device__ __forceinline__ int lean_candidate(int i, int x, int y, int W, int H) {
    const int d = (i >> 2) + 1, j = i & 3;
    const unsigned order = (x < d) ? ((y < d) ? 0x1Eu : 0x39u) : ((y < d) ? 0x78u : 0xE4u);
    const int dir = (order >> (2 * j)) & 3;                       // 0 W, 1 N, 2 E, 3 S
    int cx = x, cy = y;
    if (dir == 0) { cx = x - d; if (cx < 0) cx += W; }
    else if (dir == 1) { cy = y - d; if (cy < 0) cy += H; }
    else if (dir == 2) { cx = x + d; if (cx >= W) cx -= W; }
    else { cy = y + d; if (cy >= H) cy -= H; }
    return cx * H + cy;
}

// k_lean_pack for one slot.  Returns the descriptor priority / range through the record; live = the slot
// holds an agent of the list (every agent that is alive and on the map is on the list after a compaction).
__device__ __forceinline__ bool lean_pack_slot(const DevCtx &c, const LeanCtx &L, int s) {
    const ssb_agents_t &a = c.a;
    const int pos = a.pos[s], cod = a.cod[s];
    const bool live = cod == SSB_ALIVE && pos >= 0;
    LeanA A;
    A.sugar = a.sugar[s]; A.spice = a.spice[s]; A.lookahead = a.lookahead[s];
    A.pos = pos; A.tribe = (int16_t)a.tribe[s];
    const int r = live ? min(min(max(a.vision[s], 0), max(a.movement[s], 0)), max(c.max_dx, c.max_dy)) : 0;
    A.state = live ? (uint16_t)(LEAN_PACKED | (a.last_moved[s] == c.t ? LEAN_SKIP : 0) | (r << LEAN_RANGE_SHIFT)) : (uint16_t)0;
    L.a[s] = A;
    if (!live) return false;
    LeanB B;
    B.id = a.id[s]; B.age = a.age[s]; B.max_age = a.max_age[s];
    B.sugar_metab = a.sugar_metab[s]; B.spice_metab = a.spice_metab[s];
    B.start = pos;
    B.aggression = a.aggression[s];
    L.b[s] = B;
    return true;
}

// k_lean_unpack for one slot: the mutable fields go back to the ABI's arrays; happiness of the agents that are
// still alive after aging (updateHappiness, agent.py:1522-1530: family 0 in mode S, health 1, wealth erf).
__device__ __forceinline__ void lean_unpack_slot(const DevCtx &c, const LeanCtx &L, int s) {
    const LeanA A = L.a[s];
    if (!(A.state & LEAN_PACKED)) return;
    const ssb_agents_t &a = c.a;
    const int cod = A.state & LEAN_COD_MASK;
    a.sugar[s] = A.sugar; a.spice[s] = A.spice; a.pos[s] = A.pos; a.cod[s] = cod;
    if (!(A.state & LEAN_ACTED)) return;
    const LeanC C = L.c[s];
    a.last_moved[s] = c.t;
    a.last_sugar[s] = C.last_sugar; a.last_spice[s] = C.last_spice;
    if (A.state & LEAN_RANGED) { a.n_neighborhood[s] = C.n_neighborhood; a.n_valid_moves[s] = C.n_valid_moves; }
    if (!(A.state & LEAN_AGED)) return;
    a.age[s] = L.b[s].age + 1;
    if (cod != SSB_ALIVE) return;
    const double wealth_h = erf((A.sugar + A.spice) - c.prev_mean_wealth);
    const double family_h = erf(0.0);
    a.wealth_happiness[s] = wealth_h;
    a.family_happiness[s] = family_h;
    a.happiness[s] = family_h + 1.0 + wealth_h;
}

// Agent.doTimestep (agent.py:568-598) for a rule set with footprint dilation 0: findBestCell with the shuffled
// tie-break (1395-1443, 1479-1490), doCombat (274-294), gotoCell, collectResourcesAtCell (249-259), doMetabolism
// (485-498) with the pollution terms (cell.py:27-45), starvation, doAging (266-272).  Mirrors
// transaction_move_phase / transaction_interact_phase of ssb_rules.cuh statement by statement.
__device__ inline void lean_transaction(const DevCtx &c, const LeanCtx &L, int s, const LeanB &B, uint64_t key, const LeanStage &st) {
    LeanA A = L.a[s];
    if (!((A.state & LEAN_COD_MASK) == SSB_ALIVE && A.sugar >= 0 && A.spice >= 0) || (A.state & LEAN_SKIP)) return;
    const int W = c.p.width, H = c.p.height, pos = A.pos, x = pos / H, y = pos - x * H;
    const int r = A.state >> LEAN_RANGE_SHIFT, n = 4 * r, S = st.stride;
    const bool polluted = (c.p.flags & SSB_F_POLLUTION) != 0, spicy = (c.p.flags & SSB_F_SPICE) != 0;
    const double aggression = fmax(B.aggression, 0.0);
    const bool fighter = aggression > 0;
    // the loads of the cross (and of the own cell, in case the agent stays) fly during the shuffle
    for (int i = 0; i < n; i++) {
        const int cell = lean_candidate(i, x, y, W, H);
        lean_stage_load(st.occ + i * S, c.g.occ + cell);
        lean_stage_load(st.sugar + i * S, c.g.sugar + cell);
    }
    lean_stage_load(st.sugar + n * S, c.g.sugar + pos);
    // findWelfare's per-agent part (agent.py:1170-1201, no tag preferences)
    Welfare w;
    {
        const double sm = (double)max(B.sugar_metab, 0), pm = (double)max(B.spice_metab, 0);
        w.sugar = A.sugar; w.spice = A.spice;
        w.s_look = sm * A.lookahead; w.p_look = pm * A.lookahead;
        const double total = sm + pm;
        w.e_s = 0; w.e_p = 0;
        if (total != 0) { w.e_s = sm / total; w.e_p = pm / total; }
    }
    const double my_wealth = A.sugar + A.spice, loot = c.p.max_combat_loot;
    int b_cell = pos, b_dist = 1 << 30, b_rank = 1 << 30, b_cs = 0, b_cp = 0, b_occ = -1, hood = 0, m = 0;
    double b_welfare = -1.0, b_pl = 0.0;
    bool b_valid = false;
    if (n > 0) {
        // random.shuffle(cellsInRange), agent.py:1400-1401: rank[i] = shuffled position of candidate i
        RngStream rng{key, (uint32_t)B.id, 0u, 0u, nullptr};
        rng.at(SITE_MOVE);
        for (int i = 0; i < n; i++) st.perm[i * S] = (uint8_t)i;
        for (int i = n - 1; i >= 1; i--) {
            const uint32_t j = rand_below(rng, (uint32_t)i + 1u);
            const uint8_t t = st.perm[i * S]; st.perm[i * S] = st.perm[j * S]; st.perm[j * S] = t;
        }
        for (int k = 0; k < n; k++) st.rank[st.perm[k * S] * S] = (uint8_t)k;
        lean_stage_wait();
        // findRetaliatorsInVision (agent.py:1088-1098): the richest visible agent of every tribe
        double retaliator[27];
        if (fighter) {
            for (int i = 0; i < 27; i++) retaliator[i] = 0.0;
            for (int i = 0; i < n; i++) {
                const int o = st.occ[i * S];
                if (o < 0) continue;
                const LeanA O = L.a[o];
                const double ow = fmax(O.sugar + O.spice, 0.0);
                const int slot = min((int)O.tribe + 1, 26);
                if (ow > retaliator[slot]) retaliator[slot] = ow;
            }
        }
        for (int i = 0; i < n; i++) {
            const int o = st.occ[i * S], cs = st.sugar[i * S];
            const bool occupied = o >= 0;
            int otribe = -1;
            double ps = 0, pp = 0;
            if (occupied) {
                // findNeighborhood (agent.py:1052-1065): an agent found through a cell's occupant pointer is alive
                // (every death of this rule family clears the pointer in the same transaction)
                hood++;
                if (!fighter) continue;
                const LeanA O = L.a[o];
                otribe = O.tribe;
                // isNeighborValidPrey (agent.py:1309-1314)
                if (!(A.tribe != otribe && my_wealth >= O.sugar + O.spice)) continue;
                ps = O.sugar; pp = O.spice;
            }
            const int cell = lean_candidate(i, x, y, W, H);
            const int cp = spicy ? c.g.spice[cell] : 0;
            const double pl = polluted ? c.g.pollution[cell] : 0.0;
            const double wps = aggression * fmin(loot, ps), wpp = aggression * fmin(loot, pp);
            const double welfare = w.eval(((double)cs + wps) / (1 + pl), ((double)cp + wpp) / (1 + pl));
            if (occupied && retaliator[min(otribe + 1, 26)] > my_wealth + welfare) continue;
            const int dist = (i >> 2) + 1, rk = st.rank[i * S];
            if (!b_valid || welfare > b_welfare || (welfare == b_welfare && (dist < b_dist || (dist == b_dist && rk < b_rank)))) {
                b_cell = cell; b_welfare = welfare; b_dist = dist; b_rank = rk; b_valid = true;
                b_cs = cs; b_cp = cp; b_pl = pl; b_occ = o;
            }
            m++;
        }
        if (m == 0) m = 1;
    } else {
        lean_stage_wait();
    }
    LeanC C;
    C.last_sugar = A.sugar; C.last_spice = A.spice; C.n_neighborhood = hood + 1; C.n_valid_moves = m; C.pad[0] = 0; C.pad[1] = 0;
    uint16_t state = (uint16_t)(LEAN_PACKED | LEAN_ACTED | (n > 0 ? LEAN_RANGED : 0) | (r << LEAN_RANGE_SHIFT));
    double sugar = A.sugar, spice = A.spice;
    const int best = b_valid ? b_cell : pos;
    if (!b_valid) {          // stays on its own cell: its contents were not among the candidates
        b_cs = st.sugar[n * S];
        b_cp = spicy ? c.g.spice[pos] : 0;
        b_pl = polluted ? c.g.pollution[pos] : 0.0;
    }
    // doCombat (agent.py:274-294) then gotoCell (agent.py:1213-1219)
    if (fighter && b_valid && b_occ >= 0 && b_occ != s) {
        LeanA P = L.a[b_occ];
        const double sl = fmin(loot, P.sugar), pl = fmin(loot, P.spice);
        sugar += sl; spice += pl;
        P.sugar -= sl; P.spice -= pl;
        P.state = (uint16_t)((P.state & ~LEAN_COD_MASK) | SSB_COMBAT);       // doDeath("combat"): the killer takes the cell below
        P.pos = -1;
        L.a[b_occ] = P;
    }
    c.g.occ[pos] = -1;
    c.g.occ[best] = s;
    // collectResourcesAtCell (agent.py:249-259) + production pollution (cell.py:32-45)
    sugar += b_cs;
    spice += b_cp;
    const bool polluting = c.p.pollution_start <= c.t && c.t <= c.p.pollution_end;
    double cell_poll = b_pl;
    if (polluting) {
        cell_poll += c.p.sugar_prod_pollution * b_cs;
        cell_poll += c.p.spice_prod_pollution * b_cp;
    }
    c.g.sugar[best] = 0;
    if (spicy) c.g.spice[best] = 0;
    // doMetabolism (agent.py:485-498) + consumption pollution (cell.py:27-40)
    const double sm = (double)max(B.sugar_metab, 0), pm = (double)max(B.spice_metab, 0);
    sugar -= sm;
    spice -= pm;
    if (polluting) {
        cell_poll += c.p.sugar_cons_pollution * sm;
        cell_poll += c.p.spice_cons_pollution * pm;
        c.g.pollution[best] = cell_poll;
    }
    A.sugar = sugar; A.spice = spice; A.pos = best;
    if (sugar < 0 || spice < 0 || (sugar <= 0 && sm > 0) || (spice <= 0 && pm > 0)) {
        state |= SSB_STARVATION;                                              // doDeath("starvation") (agent.py:296-313)
        c.g.occ[best] = -1;
        A.pos = -1;
    } else {
        // doAging (agent.py:266-272); isAlive() holds here: both holdings are positive or their metabolism is 0
        if (sugar >= 0 && spice >= 0) {
            state |= LEAN_AGED;
            const int age = B.age + 1;
            if (age >= B.max_age && B.max_age != -1) {
                state |= SSB_AGING;
                c.g.occ[best] = -1;
                A.pos = -1;
            }
        }
    }
    A.state = state;
    L.a[s] = A;
    L.c[s] = C;
}

}  // namespace ssb
