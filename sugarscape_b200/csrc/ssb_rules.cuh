// ssb_rules.cuh -- one agent's timestep transaction on the device SoA.
//
// Restates reference agent.py:568-598 (Agent.doTimestep) and what it calls; each function cites
// the lines it follows.  One WARP runs one transaction: the move ranking (the reference's hot
// spot, 54-58 % of its time) is lane-parallel over the candidate cells, the rest runs on lane 0.
// In mode E the single ordered warp calls it for every agent in list order; in mode S every agent
// that won its reservation in the current commit round runs it concurrently (footprints of
// concurrent agents are disjoint, see ssb_agent_step.cuh), so no atomics are needed on cells or
// neighbours -- only on the shared slot / birth counters.
#pragma once
#include <math.h>
#include <stdint.h>

#include "../../include/ssb.h"
#include "ssb_pow.cuh"
#include "ssb_rng.cuh"
#ifndef SSB_HOSTEMU
#include "ssb_shard.cuh"
#endif                 // (tests/hostemu/cuda_shim.h supplies a plain ShardCtx for the host-compiled rule code)

namespace ssb {

// child endowment bit positions inside endow_bits[t] (sugarscape_b200/steptables.py)
enum { EB_AGGRESSION = 0, EB_FERT_AGE, EB_FERT_FACTOR, EB_INFERT_AGE, EB_LOOKAHEAD, EB_MAX_AGE,
       EB_MOVEMENT, EB_SPICE_METAB, EB_SUGAR_METAB, EB_SEX, EB_TRADE_FACTOR, EB_VISION,
       EB_DECISION_MODEL = 19 };

struct DevCtx {
    ssb_params_t p;
    ssb_grid_t g;
    ssb_agents_t a;
    int32_t max_dx, max_dy;          // per-axis memoised range (environment.py:138-139)
    int32_t t;
    double prev_mean_wealth;         // runtimeStats["meanWealth"] of the previous step (agent.py:1166)
    const uint32_t *endow_bits;      // per-timestep child endowment choices (agent.py:839-851)
    const uint32_t *tag_bits;        // per-timestep tag mismatch draws (agent.py:859-875)
    int32_t n_step_tables;
    // mode S
    int32_t *born_list;              // slots of this step's newborn (unordered)
    ShardCtx sh;                     // world <= 1: the whole simulation lives on this GPU
    // ABI 2, mode E (null otherwise; device copies of the structs bound with ssb_bind_ethics / ssb_bind_ext -- pointers,
    // so that the context every kernel carries does not grow with the families)
    const ssb_ethics_t *eth;         // decision models (null: every agent is a plain Agent)
    const ssb_ext_t *x;              // lending, friends, diseases ... (a family is off when its first array is null)
};

// ssb_ext.cuh (included below, once the basic predicates exist)
__device__ __forceinline__ bool has_lending(const DevCtx &c);
__device__ __forceinline__ bool has_disease(const DevCtx &c);
__device__ __forceinline__ double dis_mod(const DevCtx &c, int s, int which);
__device__ __forceinline__ double pos_part(double v);
__device__ inline void lending_note_death(const DevCtx &c, int s);
__device__ inline void disease_note_death(const DevCtx &c, int s);

__device__ __forceinline__ int wrapi(int v, int n) { v %= n; return v < 0 ? v + n : v; }
// wrap for |offset| < n (every neighbourhood access on maps at least as wide as the range)
__device__ __forceinline__ int wrap1(int v, int n) { return v < 0 ? v + n : (v >= n ? v - n : v); }

// Lane groups: G consecutive lanes of a warp (G = 32, 16, 8, 4) cooperate on one agent.
constexpr int MAXC_EXACT_RULES = 160;   // = MAXC_EXACT (ssb_exact.cuh): candidate cells of one ranking in mode E

template <class Rng> struct RngTraits { static constexpr bool ordered = false; };
template <> struct RngTraits<RngExact> { static constexpr bool ordered = true; };      // mode E: the one ordered stream

template <int G> struct Group {
    static __device__ __forceinline__ int lane() { return threadIdx.x & (G - 1); }
    static __device__ __forceinline__ unsigned mask() {
        return G == 32 ? 0xffffffffu : (((1u << (G & 31)) - 1u) << (threadIdx.x & 31 & ~(G - 1)));
    }
    static __device__ __forceinline__ void sync() { __syncwarp(mask()); }
    template <class T> static __device__ __forceinline__ T bcast(T v, int src) { return __shfl_sync(mask(), v, src, G); }
    template <class T> static __device__ __forceinline__ T xor_(T v, int o) { return __shfl_xor_sync(mask(), v, o, G); }
    static __device__ __forceinline__ bool any(bool p) { return __any_sync(mask(), p) != 0; }
    static __device__ __forceinline__ bool all(bool p) { return __any_sync(mask(), !p) == 0; }
};

// Per-group scratch of the move ranking, in SHARED memory.
struct WarpScratch {
    int32_t *cells;          // [maxc] candidate cell index, reference enumeration order
    uint8_t *perm, *rank, *dist;   // [maxc] shuffle permutation, its inverse, distance
    uint32_t *words;         // [64] prefetched random words of the shuffle (mode S)
    unsigned long long *retaliator;   // [27] richest visible agent per tribe, as ordered bits
};
__host__ __device__ constexpr int warp_scratch_bytes(int maxc) { return maxc * 8 + 64 * 4 + 27 * 8; }
__device__ __forceinline__ WarpScratch make_warp_scratch(unsigned char *base, int maxc) {
    WarpScratch ws;   // base must be 8-byte aligned and maxc a multiple of 8
    ws.retaliator = reinterpret_cast<unsigned long long *>(base);
    ws.cells = reinterpret_cast<int32_t *>(base + 27 * 8);
    ws.words = reinterpret_cast<uint32_t *>(base + 27 * 8 + maxc * 4);
    ws.perm = base + 27 * 8 + maxc * 4 + 64 * 4;
    ws.rank = ws.perm + maxc;
    ws.dist = ws.rank + maxc;
    return ws;
}

// Agent.isAlive (agent.py:1221-1224)
__device__ __forceinline__ bool agent_alive(const DevCtx &c, int s) {
    return c.a.cod[s] == SSB_ALIVE && c.a.sugar[s] >= 0 && c.a.spice[s] >= 0;
}

// Agent.findTribe (agent.py:1142-1151)
__device__ __forceinline__ int find_tribe(const DevCtx &c, uint32_t tags) {
    const int L = c.p.tag_len, T = c.p.max_tribes;
    if (L <= 0 || T <= 0) return -1;
    const int zeroes = L - __popc(tags);
    const double tribe_size = (double)(L + 1) / (double)T;
    const int tribe = (int)ceil((double)(zeroes + 1) / tribe_size) - 1;
    return min(tribe, T - 1);
}

// Python float ** : glibc's pow restated bit for bit (ssb_pow.cuh) -- candidate cells that are
// mathematically tied are separated by pow's rounding alone, so "accurate" is not enough.
__device__ __forceinline__ double py_pow(double x, double y) {
    int exact;
    return ssb_py_pow(x, y, &exact);
}

// Agent.findWelfare (agent.py:1170-1201).  The per-agent part (holdings, lookahead, exponents)
// is loaded once into registers; eval() is then pure arithmetic, so the candidate loop of the
// move ranking issues no dependent loads.
struct Welfare {
    double sugar, spice, s_look, p_look, e_s, e_p;

    template <bool EXT = false>
    __device__ __forceinline__ void load(const DevCtx &c, int s) {
        double sm = (double)max(c.a.sugar_metab[s], 0), pm = (double)max(c.a.spice_metab[s], 0);
        if (EXT && has_disease(c)) {     // findSugarMetabolism / findSpiceMetabolism with the disease modifiers
            sm = pos_part(c.a.sugar_metab[s] + dis_mod(c, s, SSB_DM_SUGAR_METAB));
            pm = pos_part(c.a.spice_metab[s] + dis_mod(c, s, SSB_DM_SPICE_METAB));
        }
        const double look = c.a.lookahead[s];
        sugar = c.a.sugar[s]; spice = c.a.spice[s];
        s_look = sm * look; p_look = pm * look;
        const double total = sm + pm;
        e_s = 0; e_p = 0;
        if (total != 0) { e_s = sm / total; e_p = pm / total; }
        if ((c.p.flags & SSB_F_TAGPREF) && (c.p.flags & SSB_F_TAGS) && c.p.tag_len > 0) {
            // tag preferences replace the exponents (agent.py:1190-1200); findWelfare also
            // refreshes self.tribe as a side effect (1192)
            const int L = c.p.tag_len;
            c.a.tribe[s] = find_tribe(c, c.a.tags[s]);
            const int zeroes = L - __popc(c.a.tags[s]);
            const double fz = (double)zeroes / (double)L, fo = 1 - fz;
            double pref = sm * fz + pm * fo;
            if (pref <= 0) pref = 1;
            e_s = (sm / pref) * fz; e_p = (pm / pref) * fo;
        }
    }
    // same, from values already in registers (tags / tribe refresh included)
    template <bool EXT = false>
    __device__ __forceinline__ void set(const DevCtx &c, int s, double sugar_, double spice_, int sugar_metab,
                                        int spice_metab, double look, uint32_t tags) {
        double sm = (double)max(sugar_metab, 0), pm = (double)max(spice_metab, 0);
        if (EXT && has_disease(c)) {
            sm = pos_part(sugar_metab + dis_mod(c, s, SSB_DM_SUGAR_METAB));
            pm = pos_part(spice_metab + dis_mod(c, s, SSB_DM_SPICE_METAB));
        }
        sugar = sugar_; spice = spice_;
        s_look = sm * look; p_look = pm * look;
        const double total = sm + pm;
        e_s = 0; e_p = 0;
        if (total != 0) { e_s = sm / total; e_p = pm / total; }
        if ((c.p.flags & SSB_F_TAGPREF) && (c.p.flags & SSB_F_TAGS) && c.p.tag_len > 0) {
            const int L = c.p.tag_len;
            c.a.tribe[s] = find_tribe(c, tags);
            const int zeroes = L - __popc(tags);
            const double fz = (double)zeroes / (double)L, fo = 1 - fz;
            double pref = sm * fz + pm * fo;
            if (pref <= 0) pref = 1;
            e_s = (sm / pref) * fz; e_p = (pm / pref) * fo;
        }
    }
    __device__ __forceinline__ double eval(double sugar_reward, double spice_reward) const {
        double ts = (sugar + sugar_reward) - s_look;
        double tp = (spice + spice_reward) - p_look;
        if (ts < 0) ts = 0;
        if (tp < 0) tp = 0;
        return py_pow(ts, e_s) * py_pow(tp, e_p);
    }
};

template <bool EXT = false>
__device__ __forceinline__ double find_welfare(const DevCtx &c, int s, double sugar_reward, double spice_reward) {
    Welfare w;
    w.template load<EXT>(c, s);
    return w.eval(sugar_reward, spice_reward);
}

// Kin lists (mode E): socialNetwork["children"] / ["mates"] of the INITIATING parent
// (agent.py:521-527) as singly linked lists in the kin pool.
__device__ __forceinline__ void kin_push(const DevCtx &c, int32_t *head, int owner, int other) {
    unsigned long long *cn = (unsigned long long *)c.a.counters;
    long long e;
    if (cn[SSB_C_KIN_FREE] != 0ull) {            // recycled entry of a removed owner (mode E: one ordered warp, no race)
        e = (long long)cn[SSB_C_KIN_FREE] - 1;
        cn[SSB_C_KIN_FREE] = (unsigned long long)(c.a.kin_next[e] + 1);
    } else {
        e = (long long)atomicAdd(&cn[SSB_C_N_KIN], 1ull);
        if (e >= c.a.kin_capacity) { atomicExch(&cn[SSB_C_OVERFLOW], 3ull); return; }
    }
    c.a.kin_slot[e] = other; c.a.kin_id[e] = c.a.id[other]; c.a.kin_next[e] = head[owner];
    head[owner] = (int32_t)e;
}

// removeDeadAgents: the kin lists of a removed agent die with it -- their pool entries go to the free chain
// (one thread, after the compaction; the lists of an agent are only ever read through its own record -- except
// with lending, where a dead creditor's children inherit its loans (agent.py:1362-1377): children lists stay then)
__device__ inline void kin_reclaim(const DevCtx &c) {
    const ssb_agents_t &a = c.a;
    if (!a.children_head || !a.kin_next) return;
    const bool keep_children = has_lending(c);
    long long free_head = a.counters[SSB_C_KIN_FREE] - 1;
    const long long n_dead = a.counters[SSB_C_N_DEAD];
    for (long long i = 0; i < n_dead; i++) {
        const int s = a.dead_list[i];
        for (int list = keep_children ? 1 : 0; list < 2; list++) {
            int32_t *head = list == 0 ? a.children_head : a.mates_head;
            int e = head[s];
            while (e >= 0) { const int nx = a.kin_next[e]; a.kin_next[e] = (int32_t)free_head; free_head = e; e = nx; }
            head[s] = -1;
        }
    }
    a.counters[SSB_C_KIN_FREE] = free_head + 1;
}

// the relative of pool entry e, or -1 if it is no longer alive (dead, removed, slot recycled)
__device__ __forceinline__ int kin_alive(const DevCtx &c, int e) {
    const int o = c.a.kin_slot[e];
    return (c.a.id[o] == c.a.kin_id[e] && agent_alive(c, o)) ? o : -1;
}

// Agent.doInheritance (agent.py:373-429): policies children / sons / daughters.  Heirs can be
// anywhere on the map, so inheritance only exists in the ordered mode E.
__device__ __noinline__ void do_inheritance(const DevCtx &c, int s) {
    const int policy = c.p.inheritance_policy;
    if (policy == 0 || c.p.rng_mode != SSB_RNG_EXACT) return;
    const ssb_agents_t &a = c.a;
    if (a.sugar[s] < 0) a.sugar[s] = 0;
    if (a.spice[s] < 0) a.spice[s] = 0;
    int heirs = 0;
    for (int e = a.children_head[s]; e >= 0; e = a.kin_next[e]) {
        const int o = kin_alive(c, e);
        if (o < 0) continue;
        if (policy == 1 || (policy == 2 && a.sex[o] == SSB_SEX_MALE) || (policy == 3 && a.sex[o] == SSB_SEX_FEMALE)) heirs++;
    }
    if (heirs == 0) return;
    const double si = a.sugar[s] / heirs, pi = a.spice[s] / heirs;
    for (int e = a.children_head[s]; e >= 0; e = a.kin_next[e]) {
        const int o = kin_alive(c, e);
        if (o < 0) continue;
        if (!(policy == 1 || (policy == 2 && a.sex[o] == SSB_SEX_MALE) || (policy == 3 && a.sex[o] == SSB_SEX_FEMALE))) continue;
        a.sugar[o] = a.sugar[o] + si;
        a.spice[o] = a.spice[o] + pi;
        a.sugar[s] -= si;
        a.spice[s] -= pi;
    }
}

// Agent.doDeath (agent.py:296-313)
template <bool EXT = false>
__device__ __forceinline__ void do_death(const DevCtx &c, int s, int cause) {
    if (EXT && has_lending(c)) lending_note_death(c, s);
    if (EXT && has_disease(c)) disease_note_death(c, s);
    c.a.cod[s] = cause;
    const int pos = c.a.pos[s];
    if (pos >= 0 && c.g.occ[pos] == s) c.g.occ[pos] = -1;
    c.a.pos[s] = -1;
    if (c.p.inheritance_policy != 0) do_inheritance(c, s);
}

// Agent.findMarginalRateOfSubstitution (agent.py:1023-1029)
template <bool EXT = false>
__device__ __forceinline__ void find_mrs(const DevCtx &c, int s) {
    double pm = (double)max(c.a.spice_metab[s], 0), sm = (double)max(c.a.sugar_metab[s], 0);
    if (EXT && has_disease(c)) {
        pm = pos_part(c.a.spice_metab[s] + dis_mod(c, s, SSB_DM_SPICE_METAB));
        sm = pos_part(c.a.sugar_metab[s] + dis_mod(c, s, SSB_DM_SUGAR_METAB));
    }
    const double spice_need = pm > 0 ? c.a.spice[s] / pm : 1;
    const double sugar_need = sm > 0 ? c.a.sugar[s] / sm : 1;
    c.a.mrs[s] = c.a.trade_factor[s] * (spice_need / sugar_need);
}

// Agent.findNewMarginalRateOfSubstitution (agent.py:1067-1080)
template <bool EXT = false>
__device__ __forceinline__ double find_new_mrs(const DevCtx &c, int s, double sugar, double spice) {
    double pm = (double)max(c.a.spice_metab[s], 0), sm = (double)max(c.a.sugar_metab[s], 0);
    if (EXT && has_disease(c)) {
        pm = pos_part(c.a.spice_metab[s] + dis_mod(c, s, SSB_DM_SPICE_METAB));
        sm = pos_part(c.a.sugar_metab[s] + dis_mod(c, s, SSB_DM_SUGAR_METAB));
    }
    const double spice_need = pm > 0 ? spice / pm : 1;
    const double sugar_need = sm > 0 ? sugar / sm : 1;
    if (spice_need == 1 && sugar_need == 1) return 1;
    if (spice_need == 0) return pm;
    if (sugar_need == 0) return 1 / sm;
    return spice_need / sugar_need;
}

// Agent.isFertile (agent.py:1244-1247)
__device__ __forceinline__ bool is_fertile(const DevCtx &c, int s) {
    return c.a.sugar[s] >= c.a.start_sugar[s] && c.a.spice[s] >= c.a.start_spice[s] &&
           c.a.age[s] >= c.a.fert_age[s] && c.a.age[s] < c.a.infert_age[s] && c.a.fert_factor[s] > 0;
}

// cell.neighbors in the reference's dict order (cell.py:61-89): N, S, E, W, then -- neighborhoodMode "moore", SSB_F_MOORE --
// NE, NW, SE, SW (the east / west neighbours of the north and south cells); returns how many there are (out holds up to
// SSB_MAX_NEIGHBOURS).  environmentWraparound = false (SSB_F_NOWRAP) drops the neighbours across the map's edge
// (cell.py:47-50, 98-121) -- with the reference's own test for the south neighbour, `y + 1 < height - 1`, which leaves a
// south neighbour only to the last two rows (the last row's wraps to row 0).
constexpr int SSB_MAX_NEIGHBOURS = 8;
__device__ __forceinline__ int neighbours4(const DevCtx &c, int pos, int32_t *out) {
    const int W = c.p.width, H = c.p.height, x = pos / H, y = pos - x * H;
    const int yn = y == 0 ? H - 1 : y - 1, ys = y == H - 1 ? 0 : y + 1, xe = x == W - 1 ? 0 : x + 1, xw = x == 0 ? W - 1 : x - 1;
    if (!(c.p.flags & (SSB_F_NOWRAP | SSB_F_MOORE))) {
        out[0] = x * H + yn; out[1] = x * H + ys; out[2] = xe * H + y; out[3] = xw * H + y;
        return 4;
    }
    const bool wrap = !(c.p.flags & SSB_F_NOWRAP);
    const bool hn = wrap || y - 1 >= 0, hs = wrap || !(y + 1 < H - 1), he = wrap || !(x + 1 > W - 1), hw = wrap || x - 1 >= 0;
    int n = 0;
    if (hn) out[n++] = x * H + yn;
    if (hs) out[n++] = x * H + ys;
    if (he) out[n++] = xe * H + y;
    if (hw) out[n++] = xw * H + y;
    if (c.p.flags & SSB_F_MOORE) {
        if (hn && he) out[n++] = xe * H + yn;
        if (hn && hw) out[n++] = xw * H + yn;
        if (hs && he) out[n++] = xe * H + ys;
        if (hs && hw) out[n++] = xw * H + ys;
    }
    return n;
}

}  // namespace ssb
#include "ssb_ext.cuh"
namespace ssb {

// collectResourcesAtCell (agent.py:249-259) + production pollution (cell.py:32-45)
template <bool EXT = false>
__device__ __forceinline__ void collect(const DevCtx &c, int s, int pos) {
    const int32_t cs = c.g.sugar[pos];
    const int32_t cp = (c.p.flags & SSB_F_SPICE) ? c.g.spice[pos] : 0;
    c.a.sugar[s] += cs;
    c.a.spice[s] += cp;
    if (EXT && has_lending(c)) lending_note_income(c, s, cs, cp);
    if (c.p.pollution_start <= c.t && c.t <= c.p.pollution_end) {
        double pl = c.g.pollution[pos];
        pl += c.p.sugar_prod_pollution * cs;
        pl += c.p.spice_prod_pollution * cp;
        c.g.pollution[pos] = pl;
    }
    c.g.sugar[pos] = 0;
    if (c.p.flags & SSB_F_SPICE) c.g.spice[pos] = 0;
}

// ---------------------------------------------------------------------------------------------
// agentVisionMode = agentMovementMode = "radial" (SSB_F_RADIAL; environment.py:146-173, exact mode): the
// cells in range of an agent are the disc floor(sqrt(dx^2 + dy^2)) <= r of the cells with dx <= max_dx and dy <= max_dy
// (wraparound distances).  Environment.findRadialCellRanges walks the pairs (i, j > i) of the x-major cell list, so within
// one ring g = floor(distance) a cell's partners enter its dict in ascending order of THEIR x * height + y; findCellsInRange
// concatenates the rings 1 .. r.  The travel distance sqrt(dx^2 + dy^2) is only ever compared (agent.py:1430, 1484), so
// the engine keeps dx^2 + dy^2 in its place (same order; < 256 for the ranges the candidate arrays hold).
// environmentWraparound = false (SSB_F_NOWRAP): plain distances (findWraparoundDistance, environment.py:175-179), per-axis
// caps of width - 1 / height - 1 (environment.py:141-144) -- the disc is clipped at the map's edge.
// ---------------------------------------------------------------------------------------------
constexpr int SSB_MAX_RADIAL = 15;      // (r + 1)^2 <= 256: squared distances fit the 8-bit distance field
__device__ __forceinline__ bool radial(const DevCtx &c) { return (c.p.flags & SSB_F_RADIAL) != 0; }
__device__ __forceinline__ int isqrt_small(int v) { int g = 0; while ((g + 1) * (g + 1) <= v) g++; return g; }
// Environment.maxCellDistance (environment.py:140-146): the cap of findCellsInRange (agent.py:770)
__device__ __forceinline__ int max_cell_distance(const DevCtx &c) {
    if (!radial(c)) return max(c.max_dx, c.max_dy);
    const bool nowrap = (c.p.flags & SSB_F_NOWRAP) != 0;
    const int hw = nowrap ? c.p.width - 1 : c.p.width / 2, hh = nowrap ? c.p.height - 1 : c.p.height / 2;
    return min(c.p.max_agent_range, isqrt_small(hw * hw + hh * hh));
}
// the coordinates within wraparound distance md of `centre` on an axis of n cells, ascending: at most 2 * md + 1 (n when
// 2 * md = n: the two ends coincide); out_v / out_d = coordinate and distance
__device__ __forceinline__ int axis_span(int centre, int md, int n, bool nowrap, int *out_v, int *out_d) {
    const int lo = centre - md, hi = (!nowrap && 2 * md >= n) ? lo + n - 1 : centre + md;
    int k = 0;
    if (nowrap) {                                                       // clipped at the edge, nothing wraps
        for (int p = max(lo, 0); p <= min(hi, n - 1); p++) { out_v[k] = p; out_d[k++] = abs(p - centre); }
        return k;
    }
    for (int p = max(lo, n); p <= hi; p++) { out_v[k] = p - n; out_d[k++] = abs(p - centre); }                 // wrapped past the end
    for (int p = max(lo, 0); p <= min(hi, n - 1); p++) { out_v[k] = p; out_d[k++] = abs(p - centre); }
    for (int p = lo; p <= min(hi, -1); p++) { out_v[k] = p + n; out_d[k++] = abs(p - centre); }                // wrapped before 0
    return k;
}
// f(cell, squared distance, ring) over the disc of range r around (x, y), cells in ascending x * height + y
template <class F>
__device__ __forceinline__ void for_disc(const DevCtx &c, int x, int y, int r, F f) {
    int xs[2 * SSB_MAX_RADIAL + 1], dxs[2 * SSB_MAX_RADIAL + 1], ys[2 * SSB_MAX_RADIAL + 1], dys[2 * SSB_MAX_RADIAL + 1];
    const int H = c.p.height;
    r = min(r, SSB_MAX_RADIAL);                                         // (enumerate_disc flags longer ranges)
    const bool nowrap = (c.p.flags & SSB_F_NOWRAP) != 0;            // (then the per-axis caps are min(range, width - 1): never below r)
    const int nx = axis_span(x, nowrap ? r : min(c.max_dx, r), c.p.width, nowrap, xs, dxs);       // dx, dy <= r inside the disc
    const int ny = axis_span(y, nowrap ? r : min(c.max_dy, r), H, nowrap, ys, dys);
    for (int i = 0; i < nx; i++)
        for (int j = 0; j < ny; j++) {
            const int d2 = dxs[i] * dxs[i] + dys[j] * dys[j];
            if (d2 == 0 || d2 >= (r + 1) * (r + 1)) continue;
            f(xs[i] * H + ys[j], d2, isqrt_small(d2));
        }
}
__device__ int enumerate_disc(const DevCtx &c, int x, int y, int r, const WarpScratch &ws, int maxc) {
    int start[SSB_MAX_RADIAL + 2] = {};
    const bool too_far = r > SSB_MAX_RADIAL;
    if (!too_far) {
        for_disc(c, x, y, r, [&](int, int, int g) { start[g + 1]++; });
        for (int g = 1; g <= r + 1; g++) start[g] += start[g - 1];      // start[g]: first slot of ring g
    }
    if (too_far || start[r + 1] > maxc) {                                 // the host refuses such ranges; disease bonuses can still get here
        atomicExch((unsigned long long *)&c.a.counters[SSB_C_OVERFLOW], 15ull);
        return 0;
    }
    for_disc(c, x, y, r, [&](int cell, int d2, int g) { const int i = start[g]++; ws.cells[i] = cell; ws.dist[i] = (uint8_t)d2; });
    return start[r];
}
__device__ int disc_size(const DevCtx &c, int centre, int r) {           // len(cellsInRange) (on the torus the same wherever the agent stands)
    int n = 0;
    const int x = centre / c.p.height;
    for_disc(c, x, centre - x * c.p.height, r, [&](int, int, int) { n++; });
    return n;
}

// Candidate cells of Agent.findCellsInRange (agent.py:766-779) in the dictionary insertion
// order produced by Environment.findCardinalCellRanges (environment.py:111-122) -- SURVEY.md
// Appendix A.5.  Per distance d the (up to) four cells appear in the order in which the table
// builder inserted them: the west and north cells were inserted when THEY were the source cell
// (time x'*H + y'), the east and south cells when this cell was the source.  That gives
//   W, N, E, S in the interior;  N, E, S, W if x < d;  W, E, S, N if y < d;  E, S, N, W if both;
// when 2d equals the width (height) the east and west (north and south) cells coincide and the
// single entry sits at the earlier of the two positions.
__device__ int enumerate_cross(const DevCtx &c, int x, int y, int r, const WarpScratch &ws, int maxc) {
    const int W = c.p.width, H = c.p.height;
    int n = 0;
    for (int d = 1; d <= r && n + 4 <= maxc; d++) {
        const bool has_x = d <= c.max_dx, has_y = d <= c.max_dy;
        const bool wx = x < d, wy = y < d;                  // west / north neighbour wraps around
        const int xw = wx ? x - d + W : x - d, xe = x + d >= W ? x + d - W : x + d;
        const int yn = wy ? y - d + H : y - d, ys = y + d >= H ? y + d - H : y + d;
        const bool one_x = xw == xe, one_y = yn == ys;      // 2d == W / 2d == H
        const int west = xw * H + y, east = xe * H + y, north = x * H + yn, south = x * H + ys;
        // slot order: [W if !wx] [N if !wy] [E] [S] [N if wy] [W if wx]
        const int n0 = n;
        if (has_x && !wx) ws.cells[n++] = west;
        if (has_y && !wy) ws.cells[n++] = north;
        if (has_x && !(one_x && !wx)) ws.cells[n++] = east;   // a coinciding west already took the earlier slot
        if (has_y && !(one_y && !wy)) ws.cells[n++] = south;
        if (has_y && wy && !one_y) ws.cells[n++] = north;
        if (has_x && wx && !one_x) ws.cells[n++] = west;
        for (int i = n0; i < n; i++) ws.dist[i] = (uint8_t)d;
    }
    return n;
}

__device__ __forceinline__ int agent_range(const DevCtx &c, int s) {
    const int vision = max(c.a.vision[s], 0), movement = max(c.a.movement[s], 0);
    return min(min(vision, movement), max_cell_distance(c));
}

// ---------------------------------------------------------------------------------------------
// ethics.Bentham (reference ethics.py:105-244).  Mode E, lane 0: the
// ethical ranking walks the WHOLE welfare-sorted candidate list and, per candidate, every agent in
// view -- sequential like the rest of the ordered mode.
// ---------------------------------------------------------------------------------------------
// Agent.findTimeToLive (agent.py:1113-1140)
__device__ __forceinline__ double time_to_live(const DevCtx &c, int s, bool age_limited) {
    const double sm = (double)max(c.a.sugar_metab[s], 0), pm = (double)max(c.a.spice_metab[s], 0);
    const double big = 9223372036854775807.0;            // sys.maxsize
    const double st = sm > 0 ? c.a.sugar[s] / sm : big;
    const double pt = pm > 0 ? c.a.spice[s] / pm : big;
    double ttl = st < pt ? st : pt;
    if (age_limited) { const double lim = (double)(c.a.max_age[s] - c.a.age[s]); if (lim < ttl) ttl = lim; }
    return ttl;
}

// len(cellsInRange) of an agent with range r: the cross does not depend on where it stands
__device__ __forceinline__ int cross_size(const DevCtx &c, int centre, int r) {
    if (radial(c)) return disc_size(c, centre, r);
    int n = 0;
    for (int d = 1; d <= r; d++) {
        if (d <= c.max_dx) n += (2 * d == c.p.width) ? 1 : 2;
        if (d <= c.max_dy) n += (2 * d == c.p.height) ? 1 : 2;
    }
    return n;
}

// live agents standing on the cross of radius r around `centre` (findNeighborhood, agent.py:1052-1065)
__device__ int cross_live_count(const DevCtx &c, int centre, int r) {
    const int W = c.p.width, H = c.p.height, x = centre / H, y = centre - x * H;
    int n = 0;
    if (radial(c)) {
        for_disc(c, x, y, r, [&](int cell, int, int) { const int o = c.g.occ[cell]; if (o >= 0 && agent_alive(c, o)) n++; });
        return n;
    }
    for (int d = 1; d <= r; d++) {
        if (d <= c.max_dx) {
            const int xw = wrapi(x - d, W), xe = wrapi(x + d, W);
            int o = c.g.occ[xw * H + y];
            if (o >= 0 && agent_alive(c, o)) n++;
            if (xe != xw) { o = c.g.occ[xe * H + y]; if (o >= 0 && agent_alive(c, o)) n++; }
        }
        if (d <= c.max_dy) {
            const int yn = wrapi(y - d, H), ys = wrapi(y + d, H);
            int o = c.g.occ[x * H + yn];
            if (o >= 0 && agent_alive(c, o)) n++;
            if (ys != yn) { o = c.g.occ[x * H + ys]; if (o >= 0 && agent_alive(c, o)) n++; }
        }
    }
    return n;
}

// is `cell` one of the cells in range of an agent standing on `centre` with range r
__device__ __forceinline__ bool cell_in_cross(const DevCtx &c, int centre, int r, int cell) {
    const int W = c.p.width, H = c.p.height;
    const int x = centre / H, y = centre - x * H, cx = cell / H, cy = cell - cx * H;
    if (cell == centre || r <= 0) return false;
    if (radial(c)) {
        int dx = abs(cx - x), dy = abs(cy - y);
        if (c.p.flags & SSB_F_NOWRAP) return dx * dx + dy * dy < (r + 1) * (r + 1);      // (caps of width - 1 / height - 1 never bind)
        if (W - dx < dx) dx = W - dx;
        if (H - dy < dy) dy = H - dy;
        return dx <= c.max_dx && dy <= c.max_dy && dx * dx + dy * dy < (r + 1) * (r + 1);
    }
    if (cx == x) { int d = abs(cy - y); if (H - d < d) d = H - d; return d <= min(r, c.max_dy); }
    if (cy == y) { int d = abs(cx - x); if (W - d < d) d = W - d; return d <= min(r, c.max_dx); }
    return false;
}

// Bentham.findEthicalValueOfCell (ethics.py:141-244) for the agents in view; selfishness < 0 splits the
// value into happiness and unhappiness (negativeBentham).  hood[] = live agents in range + the agent.
__device__ __noinline__ double ethical_value(const DevCtx &c, int s, int my_r, const int32_t *hood, int n_hood, int cell,
                                             double *happy, double *unhappy) {
    const ssb_agents_t &a = c.a;
    const bool spicy = (c.p.flags & SSB_F_SPICE) != 0;
    double site = (double)(c.g.sugar[cell] + (spicy ? c.g.spice[cell] : 0));
    const double max_loot = c.p.max_combat_loot * 2;
    double max_site = (double)(c.g.sugar_cap[cell] + (spicy ? c.g.spice_cap[cell] : 0));
    const int occupant = c.g.occ[cell];
    if (occupant >= 0) {
        const double w = a.sugar[occupant] + a.spice[occupant];
        site += w < max_loot ? w : max_loot;
        max_site += w < max_loot ? w : max_loot;
    }
    int32_t nb4[SSB_MAX_NEIGHBOURS];
    const int n_nb = neighbours4(c, cell, nb4);
    double neighbor_wealth = 0;
    for (int i = 0; i < n_nb; i++) neighbor_wealth += c.g.sugar[nb4[i]] + (spicy ? c.g.spice[nb4[i]] : 0);
    const double lookahead = c.eth->lookahead_factor[s];
    // len(self.findNeighborhood(cell)): live agents on MY cross moved to `cell`, plus myself
    const double future_hood = lookahead != 0 ? (double)(1 + (my_r > 0 ? cross_live_count(c, cell, my_r) : 0)) : 1.0;
    const double selfish = c.eth->selfishness[s];
    double value = 0, h = 0, u = 0;
    for (int k = 0; k < n_hood; k++) {
        const int nb = hood[k];
        if (nb < 0 || !agent_alive(c, nb)) continue;      // (nb < 0: a stored list's member whose slot has a new tenant -- dead)
        const int rn = min(min((int)eff_vision(c, nb), (int)eff_movement(c, nb)), max_cell_distance(c));   // findCellsInRange with the modifiers
        if (!(cell == a.pos[nb] || cell_in_cross(c, a.pos[nb], rn, cell))) continue;      // canReachCell
        const double metab = (double)(a.sugar_metab[nb] + a.spice_metab[nb]);               // raw, not the find* accessors
        const double cell_duration = metab > 0 ? site / metab : 0;
        const double proximity = 1.0 / 1;
        const double intensity = (1 / (1 + time_to_live_x(c, nb, false)) / (1 + c.g.pollution[cell]));
        const double duration = max_site > 0 ? cell_duration / max_site : 0;
        const double discount = c.eth->lookahead_factor[nb] != 0 ? c.eth->lookahead_discount[nb] : 0;
        double future_duration = metab > 0 ? (site - metab) / metab : site;
        future_duration = max_site > 0 ? future_duration / max_site : 0;
        int32_t nb_cells[SSB_MAX_NEIGHBOURS];
        const double future_intensity = neighbor_wealth / (c.eth->global_max_wealth * neighbours4(c, a.pos[nb], nb_cells));  // len(neighbor.cell.neighbors)
        const int in_range = rn > 0 ? cross_size(c, a.pos[nb], rn) : 0;                                // len(neighbor.cellsInRange)
        const double extent = in_range > 0 ? (double)n_hood / in_range : 1;
        const double future_extent = (in_range > 0 && lookahead != 0) ? future_hood / in_range : 1;
        const double current_reward = extent * (intensity + duration);
        const double future_reward = future_extent * (future_intensity + future_duration);
        double v = (1 * proximity) * (current_reward + (discount * future_reward));
        if (nb != s && selfish < 1) {
            v = -1 * v;
            if (cell == a.pos[nb] && v > -1) v = -1;
        }
        if (selfish >= 0) v *= nb == s ? selfish : 1 - selfish;
        else { if (v > 0) h += v; else u += v; }
        value += v;
    }
    *happy = h; *unhappy = u;
    return value;
}

// The agent's own record, loaded ONCE per transaction by every lane of the group (uniform
// addresses: one memory transaction each, all independent -> a single latency level).
struct AgentRec {
    int id, pos, cod, last_moved, vision, movement, sugar_metab, spice_metab, tribe, age, max_age, fert_age, infert_age;
    uint32_t tags;
    double sugar, spice, aggression, lookahead, fert_factor, start_sugar, start_spice;
    __device__ __forceinline__ void load(const DevCtx &c, int s) {
        const ssb_agents_t &a = c.a;
        id = a.id[s]; pos = a.pos[s]; cod = a.cod[s]; last_moved = a.last_moved[s];
        vision = a.vision[s]; movement = a.movement[s];
        sugar_metab = a.sugar_metab[s]; spice_metab = a.spice_metab[s];
        tribe = a.tribe[s]; age = a.age[s]; max_age = a.max_age[s];
        sugar = a.sugar[s]; spice = a.spice[s];
        aggression = a.aggression[s]; lookahead = a.lookahead[s];
        const bool repro = (c.p.flags & SSB_F_REPRO) != 0;
        fert_age = repro ? a.fert_age[s] : 0; infert_age = repro ? a.infert_age[s] : 0;
        fert_factor = repro ? a.fert_factor[s] : 0.0;
        start_sugar = repro ? a.start_sugar[s] : 0.0; start_spice = repro ? a.start_spice[s] : 0.0;
        tags = (c.p.flags & SSB_F_TAGS) ? a.tags[s] : 0u;
    }
    __device__ __forceinline__ bool alive() const { return cod == SSB_ALIVE && sugar >= 0 && spice >= 0; }
    __device__ __forceinline__ bool fertile() const {       // Agent.isFertile (agent.py:1244-1247)
        return sugar >= start_sugar && spice >= start_spice && age >= fert_age && age < infert_age && fert_factor > 0;
    }
};

// rankCellsInRange in full: the valid candidates in welfare order (stable: welfare descending, range ascending,
// shuffled position).  Returns their number; lane 0, after the group's ranking pass.
__device__ __noinline__ int rank_candidates(const DevCtx &c, int s, int n, const Welfare &w, int my_tribe, double aggression,
                                            const WarpScratch &ws, int32_t *cand_cell, uint8_t *cand_dist, double *cand_w) {
    const ssb_agents_t &a = c.a;
    const bool polluted = (c.p.flags & SSB_F_POLLUTION) != 0, spicy = (c.p.flags & SSB_F_SPICE) != 0;
    const bool fighter = aggression > 0;
    const double loot = c.p.max_combat_loot, my_wealth = w.sugar + w.spice;
    int m = 0;
    for (int k = 0; k < n && k < MAXC_EXACT_RULES; k++) {            // shuffled order
        const int i = ws.perm[k], cell = ws.cells[i];
        const int prey = c.g.occ[cell];
        double ps = 0, pp = 0;
        if (prey >= 0) {                                           // isNeighborValidPrey (agent.py:1309-1314)
            if (!(fighter && my_tribe != a.tribe[prey] && my_wealth >= a.sugar[prey] + a.spice[prey])) continue;
            ps = a.sugar[prey]; pp = a.spice[prey];
        }
        const double wps = aggression * fmin(loot, ps), wpp = aggression * fmin(loot, pp);
        const double pl = polluted ? c.g.pollution[cell] : 0.0;
        const double welfare = w.eval(((double)c.g.sugar[cell] + wps) / (1 + pl),
                                      ((double)(spicy ? c.g.spice[cell] : 0) + wpp) / (1 + pl));
        if (prey >= 0 && __longlong_as_double((long long)ws.retaliator[min(a.tribe[prey] + 1, 26)]) > my_wealth + welfare) continue;
        // sortCellsByWealth (agent.py:1479-1490): stable, welfare descending then range ascending
        int j = m++;
        const uint8_t dist = ws.dist[i];
        while (j > 0 && (cand_w[j - 1] < welfare || (cand_w[j - 1] == welfare && cand_dist[j - 1] > dist))) {
            cand_cell[j] = cand_cell[j - 1]; cand_dist[j] = cand_dist[j - 1]; cand_w[j] = cand_w[j - 1];
            j--;
        }
        cand_cell[j] = cell; cand_dist[j] = dist; cand_w[j] = welfare;
    }
    return m;
}

// Bentham.findBestEthicalCell (ethics.py:105-139) on top of the full welfare-sorted list, in three phases so that the
// scoring can be spread over the lanes of the group (build with -DSSB_PARALLEL_ETHICS; by default lane 0 scores
// every candidate -- same code, same result: the candidates are independent, the only write of a score is the
// idempotent agent.timeToLive side effect).  Scratch in shared memory, one per (single) mode E group.
struct EthScratch {
    int32_t cand_cell[MAXC_EXACT_RULES];
    uint8_t cand_dist[MAXC_EXACT_RULES];
    double cand_w[MAXC_EXACT_RULES];
    int32_t hood[MAXC_EXACT_RULES + 1];
    int32_t m, nh, scored;
    double v[MAXC_EXACT_RULES], h[MAXC_EXACT_RULES], u[MAXC_EXACT_RULES];
};
// phase 1, lane 0: the valid candidates in welfare order and the agents in view (live agents in range + the agent)
__device__ __noinline__ void ethical_prepare(const DevCtx &c, int s, int n, const Welfare &w, int my_tribe, double aggression,
                                             const WarpScratch &ws, EthScratch &es) {
    es.m = rank_candidates(c, s, n, w, my_tribe, aggression, ws, es.cand_cell, es.cand_dist, es.cand_w);
    int nh = 0;
    for (int i = 0; i < n && i < MAXC_EXACT_RULES; i++) {
        const int o = c.g.occ[ws.cells[i]];
        if (o >= 0 && agent_alive(c, o)) es.hood[nh++] = o;
    }
    es.hood[nh++] = s;
    es.nh = nh;
    es.scored = es.m;
}
// phase 2: candidates first, first + stride, ...  The reference scores EVERY cell before it picks (ethics.py:113-114);
// the scores past the pick only matter through findTimeToLive's side effect on the agents in view (agent.timeToLive,
// reported by the per-agent log), so a lane that walks the list alone stops at the pick unless that log is on.  Same
// for the own cell when no candidate is valid.
__device__ __noinline__ void ethical_score(const DevCtx &c, int s, int r, EthScratch &es, int first, int stride) {
    const bool all_cells = has_agentlog(c) || (c.eth->top_ordering[s] & 1) != 0, positive = c.eth->selfishness[s] >= 0;
    if (es.m == 0) {
        double h, u;
        if (all_cells && first == 0) ethical_value(c, s, r, es.hood, es.nh, c.a.pos[s], &h, &u);
        return;
    }
    for (int i = first; i < es.m; i += stride) {
        es.v[i] = ethical_value(c, s, r, es.hood, es.nh, es.cand_cell[i], &es.h[i], &es.u[i]);
        if (stride == 1 && positive && es.v[i] > 0 && !all_cells) { es.scored = i + 1; break; }
    }
}
// phase 3, lane 0: the pick.  Returns false when no candidate is valid (the agent stays put; rank 0 by definition).
__device__ __noinline__ bool ethical_choose(const DevCtx &c, int s, const EthScratch &es, int *out_cell, int *out_rank,
                                            double *out_diff) {
    if (es.m == 0) return false;
    const bool positive = c.eth->selfishness[s] >= 0;
    int chosen = -1;
    double best_u = 0, best_h = 0;
    for (int i = 0; i < es.scored; i++) {
        if (positive) { if (es.v[i] > 0) { chosen = i; break; } }
        else if (chosen < 0 || es.u[i] > best_u || (es.u[i] == best_u && es.h[i] > best_h)) { chosen = i; best_u = es.u[i]; best_h = es.h[i]; }   // stable sort, reverse
    }
    int rank = chosen >= 0 ? chosen : 0;
    if ((c.eth->top_ordering[s] & 1) != 0 && positive) {
        // "Top" variants (ethics.py:124-129): the cells re-sorted by ethical value (stable: value descending, range
        // ascending, welfare order); the first one wins whatever its sign
        rank = 0;
        for (int i = 1; i < es.m; i++)
            if (es.v[i] > es.v[rank] || (es.v[i] == es.v[rank] && es.cand_dist[i] < es.cand_dist[rank])) rank = i;
    }
    *out_cell = es.cand_cell[rank]; *out_rank = rank; *out_diff = es.cand_w[0] - es.cand_w[rank];
    return true;
}
// all lanes of the group (uniform call); the results are meaningful on lane 0
template <int G>
__device__ bool bentham_rank(const DevCtx &c, int s, int r, int n, const Welfare &w, int my_tribe, double aggression,
                             const WarpScratch &ws, int *out_cell, int *out_rank, double *out_diff) {
    typedef Group<G> Gr;
    __shared__ EthScratch es;
    const int lane = Gr::lane();
#ifdef SSB_PARALLEL_ETHICS
    const int first = lane, stride = G;
#else
    const int first = lane == 0 ? 0 : MAXC_EXACT_RULES, stride = 1;
#endif
    if (lane == 0) ethical_prepare(c, s, n, w, my_tribe, aggression, ws, es);
    Gr::sync();
    ethical_score(c, s, r, es, first, stride);
    Gr::sync();
    return lane == 0 ? ethical_choose(c, s, es, out_cell, out_rank, out_diff) : false;
}

// Temperance.findBestEthicalCell, simple variant (ethics.py:412-434, 464-465, 493-506): every candidate is scored by
// how much it would change the agent's time to live, |findTimeToLive(potentialCell) - agent.timeToLive|; a virtue roll
// on the MAIN stream picks the largest or the smallest change and moves the agent's temperance factor.  Runs even
// when the only "candidate" is the own cell (the roll is still drawn).  Lane 0.
template <class Rng>
__device__ __noinline__ void temperance_pick(const DevCtx &c, int s, int n, const Welfare &w, int my_tribe, double aggression,
                                             const WarpScratch &ws, Rng &rng, int *out_cell, int *out_rank, double *out_diff) {
    const ssb_agents_t &a = c.a;
    int32_t cand_cell[MAXC_EXACT_RULES], order[MAXC_EXACT_RULES];
    uint8_t cand_dist[MAXC_EXACT_RULES];
    double cand_w[MAXC_EXACT_RULES], score[MAXC_EXACT_RULES];
    int m = n > 0 ? rank_candidates(c, s, n, w, my_tribe, aggression, ws, cand_cell, cand_dist, cand_w) : 0;
    if (m == 0) { cand_cell[0] = a.pos[s]; cand_dist[0] = 0; cand_w[0] = 0; m = 1; }      // rankCellsInRange: [own cell]
    const double ttl = c.x->agentlog.time_to_live[s];
    for (int i = 0; i < m; i++) {          // sortCellsByWealth on the scores, stable from the welfare order
        const double sc = fabs(time_to_live_potential(c, s, cand_cell[i]) - ttl);
        int j = i;
        while (j > 0 && (score[j - 1] < sc || (score[j - 1] == sc && cand_dist[order[j - 1]] > cand_dist[i]))) {
            score[j] = score[j - 1]; order[j] = order[j - 1];
            j--;
        }
        score[j] = sc; order[j] = i;
    }
    const double roll = rand_double(rng);                          // virtueRoll = random.random()
    const double df = c.eth->decision_factor[s], dyn = c.eth->dynamic_decision_factor[s];
    int pick;
    if (roll < df) { pick = order[0]; const double nf = rint((df + dyn) * 100.0) / 100.0; c.eth->decision_factor[s] = nf <= 1 ? nf : 1; }
    else { pick = order[m - 1]; const double nf = rint((df - dyn) * 100.0) / 100.0; c.eth->decision_factor[s] = nf >= 0 ? nf : 0; }
    *out_cell = cand_cell[pick]; *out_rank = pick; *out_diff = cand_w[0] - cand_w[pick];
}

// Temperance.findEthicalValueOfCell with pecs=True (ethics.py:436-491): physical + emotional + cognitive + social score of one
// cell.  The emotional score counts the cells it has seen: timesOverharvested moves with every call.
__device__ __forceinline__ double pecs_value(const DevCtx &c, int s, int cell) {
    if (c.eth->total_metabolism[s] == 0) return 0;
    const double ttl = c.x->agentlog.time_to_live[s];
    const int32_t *rules = c.eth->pecs_rules + (long long)s * 5;
    int32_t *counts = c.eth->pecs_counts + (long long)s * 3;
    const double delta = time_to_live_potential(c, s, cell) - ttl;
    const double physical = ttl > 0 ? erf(1 / ttl) : 1;
    double e = 0;                                                   // findCellEmotionalScore
    if (delta > 1) { e = e - counts[2]; counts[2] += 1; }
    const double emotional = erf(e);
    double cognitive;                                               // findCellCognitiveScore
    if (delta < 1) cognitive = -1;
    else {
        double k = 0;
        if (delta >= 1 && delta < 2 && rules[0]) k += rules[0];
        else if (delta >= 2 && delta < 3 && rules[1]) { k += rules[1]; if (rules[2]) k -= rules[2]; }
        else if (delta >= 3 && rules[3]) { k -= rules[3]; if (rules[4]) k -= rules[4]; }
        cognitive = erf(k);
    }
    double so = 0;                                                  // findCellSocialScore
    if (delta <= 1) so = 1;
    else if (delta > 1 && delta <= 2) so -= counts[0];
    else if (delta > 2) so -= counts[1];
    so *= c.eth->social_pressure[s];
    return physical + emotional + cognitive + erf(so);
}

// Temperance with pecs=True (ethics.py:412-491): physical + emotional + cognitive + social score per candidate, the best
// score wins (no virtue roll).  The emotional score counts the cells it has seen (timesOverharvested moves while the
// candidates are scored, in welfare order).  Lane 0.
__device__ __noinline__ void pecs_pick(const DevCtx &c, int s, int n, const Welfare &w, int my_tribe, double aggression,
                                       const WarpScratch &ws, int *out_cell, int *out_rank, double *out_diff) {
    const ssb_agents_t &a = c.a;
    int32_t cand_cell[MAXC_EXACT_RULES], order[MAXC_EXACT_RULES];
    uint8_t cand_dist[MAXC_EXACT_RULES];
    double cand_w[MAXC_EXACT_RULES], score[MAXC_EXACT_RULES];
    int m = n > 0 ? rank_candidates(c, s, n, w, my_tribe, aggression, ws, cand_cell, cand_dist, cand_w) : 0;
    if (m == 0) { cand_cell[0] = a.pos[s]; cand_dist[0] = 0; cand_w[0] = 0; m = 1; }      // rankCellsInRange: [own cell]
    for (int i = 0; i < m; i++) {
        const double sc = pecs_value(c, s, cand_cell[i]);
        int j = i;                                                          // sortCellsByWealth on the scores
        while (j > 0 && (score[j - 1] < sc || (score[j - 1] == sc && cand_dist[order[j - 1]] > cand_dist[i]))) {
            score[j] = score[j - 1]; order[j] = order[j - 1];
            j--;
        }
        score[j] = sc; order[j] = i;
    }
    const int pick = order[0];
    *out_cell = cand_cell[pick]; *out_rank = pick; *out_diff = cand_w[0] - cand_w[pick];
}

// ethics.Asimov (ethics.py:8-98) among plain agents and Temperance agents (the host refuses configurations that mix it
// with the Bentham family: the second law would ask such a neighbour's own findEthicalValueOfCell, which reads the
// neighbourhood list the neighbour stored at ITS last move).  Lane 0.
//   findEthicalValueOfCell: (third-law score + number of agents in view that cannot reach the cell) * (site wealth + loot), or
//     -sys.maxsize as soon as the first law is broken -- the cell's occupant is a human, or a human who can reach the cell
//     would starve on it (in a sugar-only run `0 + 0 - 0 <= 0` holds for every human: the reference's own arithmetic);
//   findBestEthicalCell: the cells re-sorted by that value; walking them, the LAST cell some human in view orders the robot to
//     (second law: a plain agent's own welfare for the cell when another robot stands on it) wins; before any order, the
//     first cell with a positive value is taken at once; nothing found -> the robot stays where it is.
// *out_rank = -1: stays although candidates exist (lastMoveOptimal False, no rank / difference entry in the log).
constexpr int SSB_MODEL_ASIMOV = 4;
// agent.neighborhood as stored lists (ssb_ethics_t.hood_*): only when robots live next to the Bentham family
__device__ __forceinline__ bool keeps_hoods(const DevCtx &c) { return c.eth != nullptr && c.eth->hood_stride > 0; }
// findNeighborhood (agent.py:1052-1065): the live agents on `cells` (the agent's cells in range, in their order), then the agent
__device__ inline void hood_store(const DevCtx &c, int s, const int32_t *cells, int n) {
    int32_t *ls = c.eth->hood_list + (long long)s * c.eth->hood_stride, *li = c.eth->hood_id + (long long)s * c.eth->hood_stride;
    int k = 0;
    for (int i = 0; i < n && k < (int)c.eth->hood_stride - 1; i++) {
        const int o = c.g.occ[cells[i]];
        if (o >= 0 && agent_alive(c, o)) { ls[k] = o; li[k++] = c.a.id[o]; }
    }
    ls[k] = s; li[k++] = c.a.id[s];
    c.eth->hood_n[s] = k;
}
// ... for an agent that is not ranking: a newborn (agent.py:525-526), everybody after a (re)placement (sugarscape.py:265-267)
// (bare: the lists of the initial population are made by configureAgents BEFORE configureDiseases hands out the first
// infections, sugarscape.py:73-75 -- with the endowed vision and movement, whatever the diseases do to them afterwards)
__device__ __noinline__ void hood_store_here(const DevCtx &c, int s, bool bare = false) {
    int32_t cells[MAXC_EXACT_RULES];
    uint8_t dist[MAXC_EXACT_RULES];
    WarpScratch ws = {};
    ws.cells = cells; ws.dist = dist;
    const int r = bare ? min(min(max(c.a.vision[s], 0), max(c.a.movement[s], 0)), max_cell_distance(c))
                       : min(min((int)eff_vision(c, s), (int)eff_movement(c, s)), max_cell_distance(c));
    const int H = c.p.height, x = c.a.pos[s] / H, y = c.a.pos[s] - x * H;
    int n = 0;
    if (r > 0) n = radial(c) ? enumerate_disc(c, x, y, r, ws, MAXC_EXACT_RULES) : enumerate_cross(c, x, y, r, ws, MAXC_EXACT_RULES);
    hood_store(c, s, cells, n);
}
__device__ __forceinline__ bool asimov_can_reach(const DevCtx &c, int nb, int cell) {      // canReachCell (agent.py:213-216)
    const int rn = min(min((int)eff_vision(c, nb), (int)eff_movement(c, nb)), max_cell_distance(c));
    return cell == c.a.pos[nb] || cell_in_cross(c, c.a.pos[nb], rn, cell);
}
__device__ __noinline__ void asimov_pick(const DevCtx &c, int s, int n, const Welfare &w, int my_tribe, double aggression,
                                         const WarpScratch &ws, int *out_cell, int *out_rank, double *out_diff) {
    const ssb_agents_t &a = c.a;
    const bool polluted = (c.p.flags & SSB_F_POLLUTION) != 0, spicy = (c.p.flags & SSB_F_SPICE) != 0;
    int32_t cand_cell[MAXC_EXACT_RULES], order[MAXC_EXACT_RULES], hood[MAXC_EXACT_RULES + 1];
    uint8_t cand_dist[MAXC_EXACT_RULES];
    double cand_w[MAXC_EXACT_RULES], score[MAXC_EXACT_RULES];
    int m = n > 0 ? rank_candidates(c, s, n, w, my_tribe, aggression, ws, cand_cell, cand_dist, cand_w) : 0;
    const bool own_only = m == 0;
    if (own_only) { cand_cell[0] = a.pos[s]; cand_dist[0] = 0; cand_w[0] = 0; m = 1; }      // rankCellsInRange: [own cell]
    int nh = 0;                                                    // self.neighborhood: live agents in range, then the agent
    for (int i = 0; i < n && i < MAXC_EXACT_RULES; i++) {
        const int o = c.g.occ[ws.cells[i]];
        if (o >= 0 && agent_alive(c, o)) hood[nh++] = o;
    }
    hood[nh++] = s;
    const double loot = c.p.max_combat_loot, never = -9223372036854775807.0;      // -1 * sys.maxsize
    const double my_sm = eff_sugar_metab(c, s), my_pm = eff_spice_metab(c, s);
    for (int i = 0; i < m; i++) {                                  // findEthicalValueOfCell (ethics.py:41-56), then sortCellsByWealth
        const int cell = cand_cell[i], occupant = c.g.occ[cell];
        const double cs = (double)c.g.sugar[cell], cp = spicy ? (double)c.g.spice[cell] : 0.0;
        double value = cs + cp;
        if (occupant >= 0) value += fmin(a.sugar[occupant] + a.spice[occupant], loot * 2);
        const bool spice_up = cp + a.spice[s] - my_pm > 0, sugar_up = cs + a.sugar[s] - my_sm > 0;      // scoreLawThree
        double modifier = (spice_up && sugar_up) ? 1 : ((!spice_up && !sugar_up) ? -1 : 0);
        bool broken = false;
        for (int k = 0; k < nh && !broken; k++) {                  // scoreLawOne (ethics.py:58-69)
            const int nb = hood[k];
            const bool human = c.eth->model[nb] != SSB_MODEL_ASIMOV;
            const bool starves = cp + a.spice[nb] - eff_spice_metab(c, nb) <= 0 || cs + a.sugar[nb] - eff_sugar_metab(c, nb) <= 0;
            if (occupant >= 0 && nb == occupant && human) broken = true;
            else if (!asimov_can_reach(c, nb, cell)) modifier += 1;
            else if (human && starves) broken = true;
        }
        const double sc = broken ? never : modifier * value;
        int j = i;
        while (j > 0 && (score[j - 1] < sc || (score[j - 1] == sc && cand_dist[order[j - 1]] > cand_dist[i]))) {
            score[j] = score[j - 1]; order[j] = order[j - 1];
            j--;
        }
        score[j] = sc; order[j] = i;
    }
    int best = -1;
    bool temperance_around = false;       // (or Bentham: somebody whose own value of a cell is an order)
    for (int k = 0; k < nh; k++) {
        const int mk = c.eth->model[hood[k]];
        temperance_around |= (has_agentlog(c) && (mk == 2 || mk == 3)) || (mk == 1 && keeps_hoods(c));
    }
    for (int oi = 0; oi < m; oi++) {                               // findBestEthicalCell (ethics.py:19-33)
        const int i = order[oi], cell = cand_cell[i], robot = c.g.occ[cell];
        const bool robot_there = robot >= 0 && c.eth->model[robot] == SSB_MODEL_ASIMOV;
        if (robot_there || temperance_around)                      // scoreLawTwo (ethics.py:71-86)
            for (int k = 0; k < nh; k++) {
                const int nb = hood[k], nb_model = c.eth->model[nb];
                if (nb_model == SSB_MODEL_ASIMOV || !asimov_can_reach(c, nb, cell)) continue;
                if (nb_model == 1) {
                    // a Bentham neighbour with a live decision factor answers with ITS findEthicalValueOfCell (ethics.py:141-244)
                    // over the neighbourhood it last stored; members that died since are skipped but still counted
                    if (c.eth->decision_factor[nb] > 0 && keeps_hoods(c)) {
                        int32_t theirs[SSB_GROUP_HOOD_MAX];
                        const int32_t *ls = c.eth->hood_list + (long long)nb * c.eth->hood_stride, *li = c.eth->hood_id + (long long)nb * c.eth->hood_stride;
                        const int n_theirs = max(c.eth->hood_n[nb], 0);
                        for (int q = 0; q < n_theirs; q++) theirs[q] = c.a.id[ls[q]] == li[q] ? ls[q] : -1;
                        const int r_nb = min(min((int)eff_vision(c, nb), (int)eff_movement(c, nb)), max_cell_distance(c));
                        double h, u;
                        if (ethical_value(c, nb, r_nb, theirs, n_theirs, cell, &h, &u) > 0) best = i;
                    }
                    continue;
                }
                if (nb_model == 2 || nb_model == 3) {
                    // a Temperance agent with a live decision factor answers with its own findEthicalValueOfCell: |time to live on
                    // the cell - agent.timeToLive| (ethics.py:464-465, 479-480), with pecs=True its four scores (its
                    // timesOverharvested counter moves, as in the reference); with factor 0 it says nothing
                    if (c.eth->decision_factor[nb] > 0) {
                        const double v = nb_model == 3 ? pecs_value(c, nb, cell)
                                                       : fabs(time_to_live_potential(c, nb, cell) - c.x->agentlog.time_to_live[nb]);
                        if (v > 0) best = i;
                    }
                    continue;
                }
                if (!robot_there) continue;                        // a plain agent only orders the robot onto another robot
                const double agg = eff_aggression(c, nb);
                const double rs = agg * fmin(loot, a.sugar[robot]), rp = agg * fmin(loot, a.spice[robot]);
                const double pl = polluted ? c.g.pollution[cell] : 0.0;
                Welfare wn = {};
                wn.template load<true>(c, nb);                     // neighbor.findValueOfCell (agent.py:1153-1159)
                const double order_value = wn.eval(((double)c.g.sugar[cell] + rs) / (1 + pl), ((double)(spicy ? c.g.spice[cell] : 0) + rp) / (1 + pl));
                if (order_value > 0) best = i;
            }
        if (best < 0 && score[oi] > 0) { best = i; break; }
    }
    if (best < 0 || own_only) { *out_cell = a.pos[s]; *out_rank = own_only ? 0 : -1; *out_diff = 0; return; }
    *out_cell = cand_cell[best]; *out_rank = best; *out_diff = cand_w[0] - cand_w[best];
}

// MOVE part of Agent.doTimestep (agent.py:568-584), executed by the G lanes of a group for ONE
// agent: lastSugar/lastSpice, rankCellsInRange + findBestCell + moveToBestCell (1395-1443,
// 719-746, 1321-1328), collectResourcesAtCell (249-259), doMetabolism (485-498).
//
// Only the head of the stable sort (agent.py:1479-1490) is needed: the winner is the candidate
// with the highest welfare, then the smallest distance, then the earliest shuffled position --
// a strict total order, so the lanes evaluate candidates independently (lane l takes candidates
// l, l + G, ...) in chunks of CH: all cell loads of a chunk are issued together, then all occupant
// loads, then the arithmetic.  The contents of the winning cell stay in registers for the collect
// step.  Dependent memory levels of the whole move part: own record, cells, occupants.
// Returns true (on every lane) if the agent is still alive and has to run the interact part;
// rec is updated (lane 0's copy is authoritative and written back to memory).
template <int G, int CH, class Rng, bool EXT = false>
__device__ bool transaction_move_phase(const DevCtx &c, int s, AgentRec &rec, Rng &rng, const WarpScratch &ws, int maxc) {
    typedef Group<G> Gr;
    const int lane = Gr::lane();
    const ssb_agents_t &a = c.a;
    if (!rec.alive() || rec.last_moved == c.t) return false;                 // uniform
    const double last_sugar = rec.sugar, last_spice = rec.spice;
    const int pos = rec.pos;
    const int H = c.p.height, x = pos / H, y = pos - x * H;
    const bool diseased = EXT && has_disease(c);             // uniform: the find* accessors carry disease modifiers
    // (radial ranges exist in mode E only: the mode S instantiations keep the cross, at compile time)
    const int rcap = RngTraits<Rng>::ordered ? max_cell_distance(c) : max(c.max_dx, c.max_dy);
    const int r = diseased ? min(min((int)eff_vision(c, s), (int)eff_movement(c, s)), rcap)
                           : min(min(max(rec.vision, 0), max(rec.movement, 0)), rcap);
    // environmentWraparound = false: the reference memoises ranges up to width - 1 instead of width // 2 (environment.py:141-144);
    // an agent that sees beyond half the map (disease bonuses) would read them -- not reproduced: fail loudly (code 14)
    if ((c.p.flags & SSB_F_NOWRAP) && !radial(c) && diseased && min((int)eff_vision(c, s), (int)eff_movement(c, s)) > max(c.max_dx, c.max_dy))
        atomicExch((unsigned long long *)&c.a.counters[SSB_C_OVERFLOW], 14ull);
    const double aggression = diseased ? eff_aggression(c, s) : fmax(rec.aggression, 0.0);
    Welfare w = {};       // (value-initialised only to keep the front end quiet: set() assigns every field)
    w.template set<EXT>(c, s, rec.sugar, rec.spice, rec.sugar_metab, rec.spice_metab, rec.lookahead, rec.tags);
    const int my_tribe = (c.p.flags & SSB_F_TAGPREF) ? find_tribe(c, rec.tags) : rec.tribe;
    const bool polluted = (c.p.flags & SSB_F_POLLUTION) != 0;
    const bool spicy = (c.p.flags & SSB_F_SPICE) != 0;
    const bool fighter = aggression > 0;
    int n = 0;
    if (lane == 0 && r > 0) {
        if constexpr (RngTraits<Rng>::ordered) n = radial(c) ? enumerate_disc(c, x, y, r, ws, maxc) : enumerate_cross(c, x, y, r, ws, maxc);
        else n = enumerate_cross(c, x, y, r, ws, maxc);
    }
    n = Gr::bcast(n, 0);
    int b_cell = pos, b_dist = 1 << 30, b_rank = 1 << 30, b_cs = 0, b_cp = 0;
    double b_welfare = -1.0, b_pl = 0.0;
    bool b_valid = false;
    int hood = 0, m = 0;
    if (n > 0) {                                               // group-uniform branch
        // random.shuffle(cellsInRange), agent.py:1400-1401: rank[i] = shuffled position of cell i.
        // The draws are sequential (rejection sampling); mode S prefetches its counter-based
        // words with all lanes first.
        rng.at(SITE_MOVE);
        rng.template prefetch<G>(ws.words, lane);
        if (RngTraits<Rng>::ordered) {
            for (int i = lane; i < n; i += G) ws.perm[i] = (uint8_t)i;
            Gr::sync();
            if (lane == 0) {
                for (int i = n - 1; i >= 1; i--) {
                    const uint32_t j = rand_below(rng, (uint32_t)i + 1u);
                    const uint8_t t = ws.perm[i]; ws.perm[i] = ws.perm[j]; ws.perm[j] = t;
                }
                rng.end_prefetch();
            }
            Gr::sync();
            for (int k = lane; k < n; k += G) ws.rank[ws.perm[k]] = (uint8_t)k;
        } else {
            // mode S: the KEYED shuffle (tests/golden/make_golden_scalable.py) -- element i goes to the position its
            // word i has among the n words (ties by i); the words are the ones just prefetched (n <= 64)
            Gr::sync();
            for (int i = lane; i < n; i += G) {
                const uint32_t ki = ws.words[i];
                int rk = 0;
                for (int j = 0; j < n; j++) { const uint32_t kj = ws.words[j]; rk += (kj < ki || (kj == ki && j < i)) ? 1 : 0; }
                ws.rank[i] = (uint8_t)rk;
                ws.perm[rk] = (uint8_t)i;
            }
            if (lane == 0) rng.end_prefetch();
        }
        const double loot = c.p.max_combat_loot;
        const double my_wealth = w.sugar + w.spice;
        // findRetaliatorsInVision (agent.py:1088-1098): every occupant must be seen before any
        // prey cell can be judged, so fighters take one extra pass
        if (fighter) {
            for (int i = lane; i < 27; i += G) ws.retaliator[i] = 0ull;   // wealth >= 0: bit patterns order like values
            Gr::sync();
            for (int i = lane; i < n; i += G) {
                const int o = c.g.occ[ws.cells[i]];
                if (o < 0) continue;
                const double ow = a.sugar[o] + a.spice[o];
                atomicMax(&ws.retaliator[min(a.tribe[o] + 1, 26)], (unsigned long long)__double_as_longlong(fmax(ow, 0.0)));
            }
        }
        Gr::sync();
        for (int base = 0; base < n; base += G * CH) {         // group-uniform trip count
            int cell[CH], o[CH], cs[CH], cp[CH];
            double pl[CH];
            bool valid[CH];
#pragma unroll
            for (int q = 0; q < CH; q++) {                     // level 1: the cells
                const int i = base + q * G + lane;
                valid[q] = i < n;
                cell[q] = valid[q] ? ws.cells[i] : pos;
                o[q] = c.g.occ[cell[q]];
                cs[q] = c.g.sugar[cell[q]];
                cp[q] = spicy ? c.g.spice[cell[q]] : 0;
                pl[q] = polluted ? c.g.pollution[cell[q]] : 0.0;
            }
            int ocod[CH], otribe[CH];
            double osug[CH], ospi[CH];
#pragma unroll
            for (int q = 0; q < CH; q++) {                     // level 2: their occupants
                const bool has = valid[q] && o[q] >= 0;
                const int oo = has ? o[q] : s;
                ocod[q] = a.cod[oo]; osug[q] = a.sugar[oo]; ospi[q] = a.spice[oo];
                otribe[q] = (has && fighter) ? a.tribe[oo] : -1;
            }
#pragma unroll
            for (int q = 0; q < CH; q++) {                     // arithmetic
                if (!valid[q]) continue;
                const int i = base + q * G + lane;
                const bool occupied = o[q] >= 0;
                double ps = 0, pp = 0;
                if (occupied) {
                    // findNeighborhood (agent.py:1052-1065): live agents in range
                    if (ocod[q] == SSB_ALIVE && osug[q] >= 0 && ospi[q] >= 0) hood++;
                    // isNeighborValidPrey (agent.py:1309-1314)
                    if (!(fighter && my_tribe != otribe[q] && my_wealth >= osug[q] + ospi[q])) continue;
                    ps = osug[q]; pp = ospi[q];
                }
                const double wps = aggression * fmin(loot, ps), wpp = aggression * fmin(loot, pp);
                const double welfare = w.eval(((double)cs[q] + wps) / (1 + pl[q]), ((double)cp[q] + wpp) / (1 + pl[q]));
                if (occupied && __longlong_as_double((long long)ws.retaliator[min(otribe[q] + 1, 26)]) > my_wealth + welfare) continue;
                const int dist = ws.dist[i], rk = ws.rank[i];
                if (!b_valid || welfare > b_welfare ||
                    (welfare == b_welfare && (dist < b_dist || (dist == b_dist && rk < b_rank)))) {
                    b_cell = cell[q]; b_welfare = welfare; b_dist = dist; b_rank = rk; b_valid = true;
                    b_cs = cs[q]; b_cp = cp[q]; b_pl = pl[q];
                }
                m++;
            }
        }
        // group reductions: counts, then arg-best under the same strict order
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) {
            hood += Gr::xor_(hood, o);
            m += Gr::xor_(m, o);
            const int o_cell = Gr::xor_(b_cell, o), o_dist = Gr::xor_(b_dist, o), o_rank = Gr::xor_(b_rank, o);
            const int o_cs = Gr::xor_(b_cs, o), o_cp = Gr::xor_(b_cp, o);
            const double o_welfare = Gr::xor_(b_welfare, o), o_pl = Gr::xor_(b_pl, o);
            const bool o_valid = Gr::xor_((int)b_valid, o) != 0;
            const bool take = o_valid && (!b_valid || o_welfare > b_welfare ||
                (o_welfare == b_welfare && (o_dist < b_dist || (o_dist == b_dist && o_rank < b_rank))));
            if (take) {
                b_cell = o_cell; b_welfare = o_welfare; b_dist = o_dist; b_rank = o_rank; b_valid = true;
                b_cs = o_cs; b_cp = o_cp; b_pl = o_pl;
            }
        }
        if (m == 0) m = 1;
    }
    bool stays_unranked = false;      // lane 0: an Asimov agent that stays although it has candidates (lastMoveRank = len(potentialCells))
    if constexpr (EXT) {
    if (c.eth != nullptr) {                                      // uniform
        // decision models: the ethical ranking may override the greedy choice (ethics.py:105-139); the
        // optimality of the move is logged for every agent (agent.py:734-745)
        const bool engaged = c.eth->decision_factor[s] > 0;
        const bool bentham = n > 0 && engaged && c.eth->model[s] == 1 && (b_valid || has_agentlog(c));     // uniform
        int e_cell = b_cell, e_rank = 0;
        double e_diff = 0;
        bool e_picked = false;
        if (bentham) e_picked = bentham_rank<G>(c, s, r, n, w, my_tribe, aggression, ws, &e_cell, &e_rank, &e_diff);
        if (lane == 0) {
            if (n > 0) {
                int rank = e_rank, cell = e_cell;
                double diff = e_diff;
                bool picked = e_picked;
                if (c.eth->model[s] == 2 && engaged && has_agentlog(c)) {
                    temperance_pick(c, s, n, w, my_tribe, aggression, ws, rng, &cell, &rank, &diff);
                    picked = b_valid;              // (without a valid candidate the pick is the own cell: stays)
                } else if (c.eth->model[s] == 3 && engaged && has_agentlog(c)) {
                    pecs_pick(c, s, n, w, my_tribe, aggression, ws, &cell, &rank, &diff);
                    picked = b_valid;
                } else if (c.eth->model[s] == SSB_MODEL_ASIMOV && engaged) {
                    asimov_pick(c, s, n, w, my_tribe, aggression, ws, &cell, &rank, &diff);
                    picked = b_valid;              // (no valid candidate: [own cell] is ranked, the robot stays)
                }
                if (picked && cell != b_cell) {
                    b_cell = cell;
                    b_cs = c.g.sugar[cell];
                    b_cp = spicy ? c.g.spice[cell] : 0;
                    b_pl = polluted ? c.g.pollution[cell] : 0.0;
                }
                // (rank < 0: an Asimov agent that stays although it has candidates -- not the greedy cell, and agent.cell is not among
                // validMoves when the log looks for it, sugarscape.py:1154-1160)
                c.eth->last_move_optimal[s] = rank == 0; c.eth->move_rank[s] = rank < 0 ? 0 : rank; c.eth->move_diff[s] = diff;
                stays_unranked = rank < 0;
            } else {
                c.eth->last_move_optimal[s] = 1;      // [own cell] is its own best; rank and difference keep their values
                if (has_agentlog(c) && c.eth->model[s] == 1 && engaged) {
                    int32_t self_only[1] = {s};       // the own cell is still scored (findTimeToLive side effect, see ethical_score)
                    double h, u;
                    ethical_value(c, s, r, self_only, 1, pos, &h, &u);
                }
                if (has_agentlog(c) && c.eth->model[s] == 2 && engaged) {
                    int cell, rank;                   // the virtue roll is drawn for [own cell], too
                    double diff;
                    temperance_pick(c, s, 0, w, my_tribe, aggression, ws, rng, &cell, &rank, &diff);
                }
                if (has_agentlog(c) && c.eth->model[s] == 3 && engaged) {
                    int cell, rank;                   // [own cell] is scored (timesOverharvested may move)
                    double diff;
                    pecs_pick(c, s, 0, w, my_tribe, aggression, ws, &cell, &rank, &diff);
                }
            }
        }
        b_cell = Gr::bcast(b_cell, 0);
    }
    }
    if (EXT && has_agentlog(c) && lane == 0) {                         // lastValidMoves / lastMoveRank (agent.py:744-745)
        c.x->agentlog.valid_moves[s] = n > 0 ? m : 1;
        c.x->agentlog.move_rank[s] = stays_unranked ? m : ((c.eth != nullptr && n > 0) ? c.eth->move_rank[s] : 0);
    }
    int alive = 1;
    double sugar = rec.sugar, spice = rec.spice;
    const int best = b_valid ? b_cell : pos;
    if (lane == 0) {
        if constexpr (EXT) { if (keeps_hoods(c)) hood_store(c, s, ws.cells, n); }      // self.neighborhood, as of this ranking
        if (n > 0) { a.n_neighborhood[s] = hood + 1; a.n_valid_moves[s] = m; }   // + self
        if (EXT && has_group(c) && n > 0) {      // movementNeighborhood split by group; the agent itself is part of it
            int exp_n = is_member(c, s) ? 1 : 0, ctrl_n = 1 - exp_n;
            for (int i = 0; i < n; i++) {
                const int o = c.g.occ[ws.cells[i]];
                if (o >= 0 && agent_alive(c, o)) { if (is_member(c, o)) exp_n++; else ctrl_n++; }
            }
            c.x->group.control_neighbors[s] = ctrl_n; c.x->group.experimental_neighbors[s] = exp_n;
            if (group_dynamic(c)) {               // the list itself: split again at the end of the step (group_resplit_neighbors)
                int32_t *list = c.x->group.hood_list + (long long)s * SSB_GROUP_HOOD_MAX;
                int k = 0;
                for (int i = 0; i < n && k < SSB_GROUP_HOOD_MAX - 1; i++) {
                    const int o = c.g.occ[ws.cells[i]];
                    if (o >= 0 && agent_alive(c, o)) list[k++] = o;
                }
                list[k++] = s;
                c.x->group.hood_n[s] = k;
            }
        }
        if (EXT && n == 0 && group_dynamic(c) && c.x->group.hood_n[s] > 0) {
            // an empty range leaves movementNeighborhood as it was (agent.py:1398-1399 returns before updateMovementStats): a list
            // from an earlier step, whose dead have left their slots -- not reproduced: fail loudly (code 16)
            atomicExch((unsigned long long *)&c.a.counters[SSB_C_OVERFLOW], 16ull);
        }
        // lastValidMoves: findBestCell records len([own cell]) when the range is empty (agent.py:1398-1399, 745)
        if (diseased) c.x->diseases.last_valid_moves[s] = n > 0 ? m : 1;
        if (!b_valid) {     // stays on its own cell: its contents were not among the candidates
            b_cs = c.g.sugar[pos];
            b_cp = spicy ? c.g.spice[pos] : 0;
            b_pl = polluted ? c.g.pollution[pos] : 0.0;
        }
        // doCombat (agent.py:274-294) then gotoCell (agent.py:1213-1219)
        if (fighter) {
            const int prey = c.g.occ[best];
            if (prey >= 0 && prey != s) {
                const double loot = c.p.max_combat_loot;
                const double sl = fmin(loot, a.sugar[prey]), pl = fmin(loot, a.spice[prey]);
                sugar += sl; spice += pl;
                a.sugar[prey] -= sl; a.spice[prey] -= pl;
                if (EXT && has_social(c)) c.x->social.last_combat[s] = c.t;
                if (EXT && has_agentlog(c)) { c.x->agentlog.last_combat[s] = c.t; c.x->agentlog.prey_wealth[s] = sl + pl; }
                // the prey's heirs may include the killer itself (its own parent fell): keep memory current
                // around the inheritance (agent.py:274-294 updates self.sugar before prey.doDeath)
                const bool heirs = RngTraits<Rng>::ordered && c.p.inheritance_policy != 0;     // (kin lists exist in mode E only)
                if (heirs) { a.sugar[s] = sugar; a.spice[s] = spice; }
                do_death<EXT>(c, prey, SSB_COMBAT);
                if (heirs) { sugar = a.sugar[s]; spice = a.spice[s]; }
                if (EXT) group_note(c, s, SSB_GI_COMBAT, prey);
            }
        }
        if (EXT && has_agentlog(c)) c.x->agentlog.last_pollution[s] = c.g.pollution[pos];      // gotoCell (agent.py:1216)
        c.g.occ[pos] = -1;
        c.g.occ[best] = s;
        if (EXT && has_agentlog(c)) agentlog_note_neighbors(c, s, best);
        if (EXT && has_social(c)) update_neighbors(c, s, best);   // (updateFriends reads tags and ids only: the holdings are still in registers)
        if constexpr (EXT) {
            if (c.eth != nullptr && c.eth->model[s] == 3 && has_agentlog(c)) {
                // Temperance.collectResourcesAtCell / doMetabolism (ethics.py:537-543, 508-516): what the new cell adds to the
                // time to live, before collecting; the social pressure grows with every meal
                c.eth->last_delta_ttl[s] = time_to_live_potential_of(c, s, best, sugar, spice) - c.x->agentlog.time_to_live[s];
                c.eth->social_pressure[s] += c.eth->dynamic_social_pressure[s];
            }
        }
        // collectResourcesAtCell (agent.py:249-259) + production pollution (cell.py:32-45)
        sugar += b_cs;
        spice += b_cp;
        if (EXT && has_lending(c)) lending_note_income(c, s, b_cs, b_cp);
        if (EXT && has_misc(c)) {        // doUniversalIncome (agent.py:708-714): spice first
            const ssb_misc_t &M = c.x->misc;
            if ((c.t - M.last_income_spice[s]) >= M.income_interval_spice) { spice += M.universal_spice[s]; M.last_income_spice[s] = c.t; }
            if ((c.t - M.last_income_sugar[s]) >= M.income_interval_sugar) { sugar += M.universal_sugar[s]; M.last_income_sugar[s] = c.t; }
        }
        const bool polluting = c.p.pollution_start <= c.t && c.t <= c.p.pollution_end;
        double cell_poll = b_pl;
        if (polluting) {
            cell_poll += c.p.sugar_prod_pollution * b_cs;
            cell_poll += c.p.spice_prod_pollution * b_cp;
        }
        c.g.sugar[best] = 0;
        if (spicy) c.g.spice[best] = 0;
        // doMetabolism (agent.py:485-498) + consumption pollution (cell.py:27-40)
        const double sm = diseased ? eff_sugar_metab(c, s) : (double)max(rec.sugar_metab, 0);
        const double pm = diseased ? eff_spice_metab(c, s) : (double)max(rec.spice_metab, 0);
        sugar -= sm;
        spice -= pm;
        if (polluting) {
            cell_poll += c.p.sugar_cons_pollution * sm;
            cell_poll += c.p.spice_cons_pollution * pm;
            c.g.pollution[best] = cell_poll;
        }
        a.last_sugar[s] = last_sugar;
        a.last_spice[s] = last_spice;
        a.last_moved[s] = c.t;
        a.sugar[s] = sugar;
        a.spice[s] = spice;
        a.pos[s] = best;
        if (sugar < 0 || spice < 0 || (sugar <= 0 && sm > 0) || (spice <= 0 && pm > 0)) {
            // doDeath("starvation") (agent.py:296-313)
            if (EXT && has_lending(c)) lending_note_death(c, s);
            if (EXT && has_disease(c)) disease_note_death(c, s);
            a.cod[s] = SSB_STARVATION;
            c.g.occ[best] = -1;
            a.pos[s] = -1;
            if (c.p.inheritance_policy != 0) do_inheritance(c, s);
            alive = 0;
        }
    }
    Gr::sync();      // lane 0's writes are ordered before anything the group reads next
    alive = Gr::bcast(alive, 0);
    rec.sugar = Gr::bcast(sugar, 0);
    rec.spice = Gr::bcast(spice, 0);
    rec.pos = alive ? best : -1;
    rec.last_moved = c.t;
    return alive != 0;
}

// Agent.doTagging (agent.py:556-566)
template <class Rng>
__device__ __noinline__ void do_tagging(const DevCtx &c, int s, Rng &rng) {
    int32_t nb[SSB_MAX_NEIGHBOURS];
    const int n_nb = neighbours4(c, c.a.pos[s], nb);
    rng.at(SITE_TAGGING);
    shuffle(rng, nb, n_nb);
    for (int i = 0; i < n_nb; i++) {
        const int o = c.g.occ[nb[i]];
        if (o < 0) continue;
        const uint32_t bit = 1u << rand_below(rng, (uint32_t)c.p.tag_len);
        c.a.tags[o] = (c.a.tags[o] & ~bit) | (c.a.tags[s] & bit);
        c.a.tribe[o] = find_tribe(c, c.a.tags[o]);
    }
}

// Agent.doTrading (agent.py:600-706)
template <class Rng, bool EXT = false>
__device__ __noinline__ void do_trading(const DevCtx &c, int s, Rng &rng) {
    const ssb_agents_t &a = c.a;
    if (a.trade_factor[s] == 0) return;
    a.trade_volume[s] = 0;
    double sum_sugar_price = 0, sum_spice_price = 0;
    find_mrs<EXT>(c, s);
    int32_t nb[SSB_MAX_NEIGHBOURS], traders[SSB_MAX_NEIGHBOURS];
    int n = 0;
    const int n_nb = neighbours4(c, a.pos[s], nb);
    for (int i = 0; i < n_nb; i++) {
        const int o = c.g.occ[nb[i]];
        if (o >= 0 && agent_alive(c, o) && a.mrs[o] != a.mrs[s]) traders[n++] = o;
    }
    rng.at(SITE_TRADING);
    shuffle(rng, traders, n);
    double sugar_price = 0, spice_price = 0;
    for (int k = 0; k < n; k++) {
        const int tr = traders[k];
        int spice_seller = -1, sugar_seller = -1;
        for (;;) {
            const double tm = a.mrs[tr], my = a.mrs[s];
            if ((tm >= 1 && my >= 1) || (tm < 1 && my < 1) || tm == my) break;   // canTradeWithNeighbor
            if (tm > my) { spice_seller = tr; sugar_seller = s; }
            else { spice_seller = s; sugar_seller = tr; }
            const double ssm = a.mrs[spice_seller], sum = a.mrs[sugar_seller];
            if (ssm < 0 || sum < 0) { spice_seller = sugar_seller = -1; break; }
            const double price = sqrt(ssm * sum);
            if (price < 1) { spice_price = 1; sugar_price = price; }
            else { spice_price = price; sugar_price = 1; }
            if (a.spice[spice_seller] - spice_price < (double)a.spice_metab[spice_seller] ||
                a.sugar[sugar_seller] - sugar_price < (double)a.sugar_metab[sugar_seller]) break;
            const double ss_new = find_new_mrs<EXT>(c, spice_seller, a.sugar[spice_seller] + sugar_price,
                                               a.spice[spice_seller] - spice_price);
            const double su_new = find_new_mrs<EXT>(c, sugar_seller, a.sugar[sugar_seller] - sugar_price,
                                               a.spice[sugar_seller] + spice_price);
            const bool better_ss_mrs = fabs(1 - ssm) > fabs(1 - ss_new);
            const bool better_su_mrs = fabs(1 - sum) > fabs(1 - su_new);
            const bool better_ss_w = find_welfare<EXT>(c, spice_seller, sugar_price, -1 * spice_price) >=
                                     find_welfare<EXT>(c, spice_seller, 0, 0);
            const bool better_su_w = find_welfare<EXT>(c, sugar_seller, -1 * sugar_price, spice_price) >=
                                     find_welfare<EXT>(c, sugar_seller, 0, 0);
            const bool crossing = ss_new < su_new;
            if ((better_ss_mrs || better_ss_w) && (better_su_mrs || better_su_w) && !crossing) {
                a.sugar[spice_seller] += sugar_price; a.spice[spice_seller] -= spice_price;
                a.sugar[sugar_seller] -= sugar_price; a.spice[sugar_seller] += spice_price;
                find_mrs<EXT>(c, spice_seller);
                find_mrs<EXT>(c, sugar_seller);
            } else break;
        }
        if (spice_seller >= 0 && sugar_seller >= 0) {   // "a trade occurred" as the reference logs it
            a.trade_volume[s] += 1;
            sum_sugar_price += sugar_price;
            sum_spice_price += spice_price;
            if (EXT) group_note(c, s, SSB_GI_TRADE, tr);
        }
    }
    a.trade_price[s] = fmax(sum_spice_price, sum_sugar_price);
}

__device__ __forceinline__ int empty_neighbours(const DevCtx &c, int pos, int32_t *out) {
    int32_t nb[SSB_MAX_NEIGHBOURS];
    int n = 0;
    const int n_nb = neighbours4(c, pos, nb);
    for (int i = 0; i < n_nb; i++) if (c.g.occ[nb[i]] < 0) out[n++] = nb[i];
    return n;
}

// slot for a newborn: recycled slot if any, else bump the high-water mark
__device__ __forceinline__ int alloc_slot(const DevCtx &c) {
    unsigned long long *cn = (unsigned long long *)c.a.counters;
    long long k = (long long)atomicAdd(&cn[SSB_C_N_FREE], (unsigned long long)-1LL) - 1;
    if (k >= 0) return c.a.free_slots[k];
    atomicAdd(&cn[SSB_C_N_FREE], 1ull);   // undo: the free list was empty
    const long long slot = (long long)atomicAdd(&cn[SSB_C_N_SLOTS], 1ull);
    if (slot >= (c.sh.world > 1 ? c.sh.slot_hi : c.a.capacity)) {
        atomicAdd(&cn[SSB_C_N_SLOTS], (unsigned long long)-1LL);
        atomicExch(&cn[SSB_C_OVERFLOW], 1ull);
        return -1;
    }
    return (int)slot;
}

// Agent.doReproduction (agent.py:500-554) with findChildEndowment (781-910) and addChildToCell
// (164-182).  EXACT = mode E (children join the list immediately, IDs in birth order).
// The reference walks the four neighbours one after the other, each time re-reading the mate and
// the mate's empty neighbours.  Here both rings (the 4 neighbours and their 16 neighbours) are
// loaded up front in two batches of independent requests; nothing else can change those cells
// while this agent holds its reservation, so the only updates are this call's own child
// placements, which are tracked in the local copies.
template <class Rng, bool EXACT, bool EXT = false>
__device__ __noinline__ void do_reproduction(const DevCtx &c, int s, Rng &rng) {
    const ssb_agents_t &a = c.a;
    const bool diseased = EXT && has_disease(c);
    if (!agent_alive(c, s) || !(diseased ? is_fertile_x(c, s) : is_fertile(c, s))) return;
    const int pos = a.pos[s];
    // NB: neighbours an agent can have -- 4, or 8 with the Moore neighbourhood, an option of the ordered mode only (the
    // scalable mode's instantiations keep their arrays)
    constexpr int NB = EXACT ? SSB_MAX_NEIGHBOURS : 4;
    int32_t nb0[NB], nb[NB];
    const int n_nb = neighbours4(c, pos, nb0);               // (4 unless environmentWraparound is off / neighborhoodMode is "moore")
    for (int i = 0; i < NB; i++) nb[i] = i < n_nb ? nb0[i] : pos;
    rng.at(SITE_REPRODUCTION);
    shuffle(rng, nb, n_nb);
    // batch 1: occupancy of the neighbours (shuffled order) and of each neighbour's ring
    int32_t o[NB], ring[NB][NB], ring_occ[NB][NB], ring_n[NB];
#pragma unroll
    for (int k = 0; k < NB; k++) {
        o[k] = k < n_nb ? c.g.occ[nb[k]] : -1;
        ring_n[k] = k < n_nb ? neighbours4(c, nb[k], ring[k]) : 0;
#pragma unroll
        for (int j = 0; j < NB; j++) ring_occ[k][j] = j < ring_n[k] ? c.g.occ[ring[k][j]] : 0;
    }
    // batch 2: the mates' records
    int m_cod[NB], m_sex[NB], m_age[NB], m_fa[NB], m_ia[NB];
    double m_sugar[NB], m_spice[NB], m_ss[NB], m_sp[NB], m_ff[NB], m_fmod[NB];      // m_fmod: the mate's fertility modifier (diseases)
#pragma unroll
    for (int k = 0; k < NB; k++) {
        const bool has = o[k] >= 0;
        const int m = has ? o[k] : s;
        m_cod[k] = a.cod[m]; m_sex[k] = a.sex[m]; m_age[k] = a.age[m];
        m_fa[k] = a.fert_age[m]; m_ia[k] = a.infert_age[m];
        m_sugar[k] = a.sugar[m]; m_spice[k] = a.spice[m];
        m_ss[k] = a.start_sugar[m]; m_sp[k] = a.start_spice[m]; m_ff[k] = a.fert_factor[m];
        m_fmod[k] = diseased ? dis_mod(c, m, SSB_DM_FERTILITY) : 0.0;
    }
    // own empty neighbours, N S E W order, computed ONCE (agent.py:506): stale by design
    int32_t own_empty[NB];
    int n_own = 0;
    for (int i = 0; i < n_nb; i++)
        for (int k = 0; k < n_nb; k++)
            if (nb[k] == nb0[i] && o[k] < 0) { own_empty[n_own++] = nb0[i]; break; }
    // own record
    const int my_sex = a.sex[s], my_age = a.age[s], my_fa = a.fert_age[s], my_ia = a.infert_age[s];
    const double my_ss = a.start_sugar[s], my_sp = a.start_spice[s], my_ff = a.fert_factor[s];
    double my_sugar = a.sugar[s], my_spice = a.spice[s];
    const double my_fmod = diseased ? dis_mod(c, s, SSB_DM_FERTILITY) : 0.0;
    const bool age_ok = my_age >= my_fa && my_age < my_ia && (diseased ? (my_ff + my_fmod) > 0 : my_ff > 0);
    unsigned long long *cn = (unsigned long long *)a.counters;
    bool newborn_at[NB];                                   // neighbour cell now holds a child of this call
    for (int k = 0; k < NB; k++) newborn_at[k] = false;
    int births = 0;                                        // = len(mates): one child per neighbour at most
    for (int k = 0; k < n_nb; k++) {
        if (o[k] < 0) continue;
        const int m = o[k];
        // neighbor.isAlive(): a newborn placed by this call is alive and never fertile at age 0
        if (!newborn_at[k] && !(m_cod[k] == SSB_ALIVE && m_sugar[k] >= 0 && m_spice[k] >= 0)) continue;
        bool mate_fertile = !newborn_at[k] && m_sugar[k] >= m_ss[k] && m_spice[k] >= m_sp[k] &&
                            m_age[k] >= m_fa[k] && m_age[k] < m_ia[k] && (diseased ? (m_ff[k] + m_fmod[k]) > 0 : m_ff[k] > 0);
        int mate_sex = m_sex[k];
        if (newborn_at[k]) {     // rare: evaluate the child exactly as the reference would
            mate_sex = a.sex[m];
            mate_fertile = diseased ? is_fertile_x(c, m) : is_fertile(c, m);
            m_sugar[k] = a.sugar[m]; m_spice[k] = a.spice[m];
            m_ss[k] = a.start_sugar[m]; m_sp[k] = a.start_spice[m]; m_ff[k] = a.fert_factor[m];
        }
        const bool compatible = ((my_sex == SSB_SEX_FEMALE && mate_sex == SSB_SEX_MALE) ||
                                 (my_sex == SSB_SEX_MALE && mate_sex == SSB_SEX_FEMALE)) && mate_fertile;
        int32_t pool[2 * NB];
        int np = 0;
        for (int i = 0; i < n_own; i++) pool[np++] = own_empty[i];
        for (int j = 0; j < ring_n[k]; j++) if (ring_occ[k][j] < 0) pool[np++] = ring[k][j];
        shuffle(rng, pool, np);    // drawn before the compatibility test (reference quirk)
        const bool fertile_now = age_ok && my_sugar >= my_ss && my_spice >= my_sp;
        if (!(fertile_now && compatible && np != 0)) continue;
        // current occupancy of a pool cell: neighbour cells and ring cells are all tracked locally
        auto occupied_now = [&](int cell) {
            for (int q = 0; q < n_nb; q++) {
                if (nb[q] == cell) return o[q] >= 0;
                for (int j = 0; j < ring_n[q]; j++) if (ring[q][j] == cell) return ring_occ[q][j] >= 0;
            }
            return c.g.occ[cell] >= 0;
        };
        int cell = pool[--np];
        while (occupied_now(cell) && np != 0) cell = pool[--np];
        if (occupied_now(cell)) continue;
        const int ch = alloc_slot(c);
        if (ch < 0) return;
        const uint32_t eb = (c.endow_bits && c.t < c.n_step_tables) ? c.endow_bits[c.t] : 0u;
#define SSB_PICK(field, bit) (((eb >> (bit)) & 1u) ? a.field[m] : a.field[s])
        a.id[ch] = EXACT ? (int32_t)atomicAdd(&cn[SSB_C_NEXT_ID], 1ull) : -1;
        a.born[ch] = c.t; a.cod[ch] = SSB_ALIVE; a.age[ch] = 0;
        a.aggression[ch] = SSB_PICK(aggression, EB_AGGRESSION);
        a.fert_age[ch] = SSB_PICK(fert_age, EB_FERT_AGE);
        a.fert_factor[ch] = SSB_PICK(fert_factor, EB_FERT_FACTOR);
        a.infert_age[ch] = SSB_PICK(infert_age, EB_INFERT_AGE);
        a.lookahead[ch] = SSB_PICK(lookahead, EB_LOOKAHEAD);
        a.max_age[ch] = SSB_PICK(max_age, EB_MAX_AGE);
        a.movement[ch] = SSB_PICK(movement, EB_MOVEMENT);
        a.spice_metab[ch] = SSB_PICK(spice_metab, EB_SPICE_METAB);
        a.sugar_metab[ch] = SSB_PICK(sugar_metab, EB_SUGAR_METAB);
        a.sex[ch] = SSB_PICK(sex, EB_SEX);
        a.trade_factor[ch] = SSB_PICK(trade_factor, EB_TRADE_FACTOR);
        a.vision[ch] = SSB_PICK(vision, EB_VISION);
#undef SSB_PICK
        const double sugar_cost = my_ss / (my_ff * 2), spice_cost = my_sp / (my_ff * 2);
        const double mate_sugar_cost = m_ss[k] / (m_ff[k] * 2), mate_spice_cost = m_sp[k] / (m_ff[k] * 2);
        a.start_sugar[ch] = a.sugar[ch] = sugar_cost + mate_sugar_cost;
        a.start_spice[ch] = a.spice[ch] = spice_cost + mate_spice_cost;
        uint32_t tags = 0;
        if (c.p.flags & SSB_F_TAGS) {
            const uint32_t tb = (c.tag_bits && c.t < c.n_step_tables) ? c.tag_bits[c.t] : 0u;
            const uint32_t ts = a.tags[s], tm = a.tags[m];
            int draw = 0;
            for (int i = 0; i < c.p.tag_len; i++) {
                const uint32_t bs = (ts >> i) & 1u, bm = (tm >> i) & 1u;
                const uint32_t b = bs == bm ? bs : ((tb >> draw++) & 1u);
                tags |= b << i;
            }
        }
        a.tags[ch] = tags;
        a.tribe[ch] = find_tribe(c, tags);
        a.mrs[ch] = 1;
        a.last_moved[ch] = c.t; a.last_sugar[ch] = 0; a.last_spice[ch] = 0;
        a.n_neighborhood[ch] = 0; a.n_valid_moves[ch] = 0;
        a.happiness[ch] = 0; a.wealth_happiness[ch] = 0; a.family_happiness[ch] = 0;
        a.trade_volume[ch] = 0; a.trade_price[ch] = 0; a.last_repro[ch] = -1;
        a.children_head[ch] = -1; a.mates_head[ch] = -1;
        const int my_id = a.id[s], mate_id = a.id[m];
        a.father[ch] = my_sex == SSB_SEX_FEMALE ? mate_id : my_id;
        a.mother[ch] = my_sex == SSB_SEX_FEMALE ? my_id : mate_id;
        if (EXT && c.eth != nullptr) {
            // the paired endowments follow ONE draw (agent.py:845-851); the child's class is that parent's
            const int from = ((eb >> EB_DECISION_MODEL) & 1u) ? m : s;
            // (Temperance.spawnChild builds the child with the default pecs=False, ethics.py:549-550)
            c.eth->model[ch] = c.eth->model[from] == 3 ? 2 : c.eth->model[from]; c.eth->selfishness[ch] = c.eth->selfishness[from];
            c.eth->top_ordering[ch] = c.eth->top_ordering[from];
            c.eth->decision_factor[ch] = c.eth->decision_factor[from];
            c.eth->lookahead_factor[ch] = c.eth->lookahead_factor[from];
            c.eth->lookahead_discount[ch] = c.eth->lookahead_discount[from];
            c.eth->dynamic_selfishness[ch] = c.eth->dynamic_selfishness[from];
            c.eth->dynamic_decision_factor[ch] = c.eth->dynamic_decision_factor[from];
            c.eth->dynamic_social_pressure[ch] = c.eth->dynamic_social_pressure[from];
            for (int k = 0; k < 5; k++) c.eth->pecs_rules[(long long)ch * 5 + k] = 0;
            for (int k = 0; k < 3; k++) c.eth->pecs_counts[(long long)ch * 3 + k] = 0;
            c.eth->social_pressure[ch] = 0; c.eth->last_delta_ttl[ch] = 0;
            c.eth->last_move_optimal[ch] = 1; c.eth->move_rank[ch] = 0; c.eth->move_diff[ch] = 0;      // lastMoveOptimal starts True
        }
        if (EXT && has_group(c)) group_reset_slot(c, ch);
        if (EXT && has_agentlog(c)) agentlog_reset_slot(c, ch);
        if (diseased) disease_child(c, ch, s, m, eb);
        if (EXT && has_misc(c)) misc_child(c, ch, s, m, eb);
        if (EXT && has_social(c)) {
            social_reset_slot(c, ch);
            c.x->social.max_friends[ch] = ((eb >> EB_MAX_FRIENDS) & 1u) ? c.x->social.max_friends[m] : c.x->social.max_friends[s];
        }
        if (EXT && has_lending(c)) {
            const ssb_lending_t &L = c.x->lending;
            lending_reset_slot(c, ch);
            L.base_interest_rate[ch] = ((eb >> EB_BASE_INTEREST) & 1u) ? L.base_interest_rate[m] : L.base_interest_rate[s];
            L.lending_factor[ch] = ((eb >> EB_LENDING_FACTOR) & 1u) ? L.lending_factor[m] : L.lending_factor[s];
            L.loan_duration[ch] = ((eb >> EB_LOAN_DURATION) & 1u) ? L.loan_duration[m] : L.loan_duration[s];
        }
        if (EXT && has_misc(c)) {
            // the child is depressed with the configured chance, decided by the private stream (agent.py:889-895)
            const double draw = c.t < c.x->misc.n_tables ? c.x->misc.depression_draw[c.t] : 1.0;
            if (draw <= c.x->misc.depression_percentage) depress(c, ch);
        }
        if constexpr (EXT) {       // Temperance.totalMetabolism, taken at construction (after a depression trigger)
            if (c.eth != nullptr) c.eth->total_metabolism[ch] = (double)(max(a.sugar_metab[ch], 0) + max(a.spice_metab[ch], 0));
        }
        a.pos[ch] = cell;
        c.g.occ[cell] = ch;
        for (int q = 0; q < n_nb; q++) {    // keep the local copies current
            if (nb[q] == cell) { o[q] = ch; newborn_at[q] = true; }
            for (int j = 0; j < ring_n[q]; j++) if (ring[q][j] == cell) ring_occ[q][j] = ch;
        }
        const long long nth = (long long)atomicAdd(&cn[SSB_C_N_BORN], 1ull);
        if (EXACT) {   // sugarscape.addAgent: the list grows while it is being iterated
            const long long at = (long long)atomicAdd(&cn[SSB_C_N_LIST], 1ull);
            if (at < a.capacity) a.order[at] = ch; else atomicExch(&cn[SSB_C_OVERFLOW], 1ull);
        } else {
            if (nth < a.capacity) c.born_list[nth] = ch; else atomicExch(&cn[SSB_C_OVERFLOW], 1ull);
        }
        collect<EXT>(c, ch, cell);      // the child collects at birth (agent.py:175)
        if constexpr (EXT) { if (keeps_hoods(c)) hood_store_here(c, ch); }      // child.findCellsInRange(); child.findNeighborhood() (agent.py:525-526)
        if (EXACT) {               // kin bookkeeping of the initiating parent (agent.py:521-527)
            bool known = false;
            for (int e = a.mates_head[s]; e >= 0; e = a.kin_next[e])
                if (a.kin_slot[e] == m && a.kin_id[e] == mate_id) known = true;
            if (!known) kin_push(c, a.mates_head, s, m);
            kin_push(c, a.children_head, s, ch);
        }
        my_sugar -= sugar_cost; my_spice -= spice_cost;
        a.sugar[s] = my_sugar; a.spice[s] = my_spice;
        m_sugar[k] -= mate_sugar_cost; m_spice[k] -= mate_spice_cost;
        a.sugar[m] = m_sugar[k];
        a.spice[m] = m_spice[k];
        a.last_repro[s] = c.t;
        births++;
        if (EXT) group_note(c, s, SSB_GI_REPRODUCTION, m);
    }
    if (EXT && has_agentlog(c)) c.x->agentlog.last_mates[s] = births;
}

// INTERACT part of Agent.doTimestep (agent.py:585-598): tagging, trading, reproduction, aging,
// happiness.  Lane 0 only.  `rec` holds the agent's record after the move part.
template <class Rng, bool EXACT, bool EXT = false>
__device__ void transaction_interact_phase(const DevCtx &c, int s, const AgentRec &rec, Rng &rng) {
    const ssb_agents_t &a = c.a;
    if ((c.p.flags & SSB_F_TAGS) && (c.p.flags & SSB_F_TAGGING)) do_tagging(c, s, rng);
    if (c.p.flags & SSB_F_TRADE) do_trading<Rng, EXT>(c, s, rng);
    // doReproduction starts with isFertile(): only trading can have raised the holdings since the
    // move part, so without trade the record decides and most agents skip the call
    const bool diseased = EXT && has_disease(c);
    if ((c.p.flags & SSB_F_REPRO) && ((c.p.flags & SSB_F_TRADE) || diseased || rec.fertile())) do_reproduction<Rng, EXACT, EXT>(c, s, rng);
    const bool lending = EXT && has_lending(c);
    if (lending) do_lending(c, s, rng);
    if (diseased) do_disease(c, s, rng);
    // doAging (agent.py:266-272); reproduction and trading leave the agent alive (isAlive looks at
    // the holdings, which they may have changed)
    double sugar = rec.sugar, spice = rec.spice;
    if ((c.p.flags & (SSB_F_TRADE | SSB_F_REPRO)) || lending) { sugar = a.sugar[s]; spice = a.spice[s]; }
    if (!(sugar >= 0 && spice >= 0)) return;     // isAlive() false: no aging, no happiness update
    const int age = rec.age + 1;
    a.age[s] = age;
    if (age >= rec.max_age && rec.max_age != -1) { do_death<EXT>(c, s, SSB_AGING); return; }
    // updateHappiness (agent.py:1522-1530): conflict (0: no aggression in the supported examples'
    // logs) + family + health (+1: no diseases) + social (0: no friends) + wealth
    const double unit = EXT ? happiness_unit(c, s) : 1.0;       // happinessUnit: 0.5763 when depressed
    double diff_wealth = (sugar + spice) - c.prev_mean_wealth;  // findWealthHappiness (agent.py:1164-1168)
    if (EXT && has_misc(c)) diff_wealth *= unit;
    const double wealth_h = erf(diff_wealth);
    // findFamilyHappiness (agent.py:940-958): mode E only (mode S keeps no kin lists: the relatives
    // can be anywhere and change concurrently)
    double family = 0;
    if (EXACT && (c.p.flags & SSB_F_REPRO)) {
        for (int e = a.children_head[s]; e >= 0; e = a.kin_next[e]) {
            const int o = kin_alive(c, e);
            if (o >= 0) { family += unit; if (diseased && is_sick(c, o)) family -= unit * 0.5; if (a.born[o] == c.t) family += unit; } else family -= unit;
        }
        for (int e = a.mates_head[s]; e >= 0; e = a.kin_next[e]) {
            const int o = kin_alive(c, e);
            if (o >= 0) { family += unit; if (diseased && is_sick(c, o)) family -= unit * 0.5; } else family -= unit;
        }
    }
    const double family_h = erf(family);
    a.wealth_happiness[s] = wealth_h;
    a.family_happiness[s] = family_h;
    if (EXT && (has_social(c) || diseased || has_misc(c))) {
        // findConflictHappiness (agent.py:912-918), findSocialHappiness (1099-1105), findHealthHappiness (1017-1021)
        double conflict_h = 0, social_h = 0;
        if (has_social(c)) {
            const ssb_social_t &S = c.x->social;
            if (S.last_combat[s] == c.t) conflict_h = eff_aggression(c, s) > 1 ? unit : unit * -1;
            if (S.max_friends[s] != 0) social_h = ((S.n_friends[s] * (2.0 / S.max_friends[s])) - 1) * unit;
            S.conflict_happiness[s] = conflict_h; S.social_happiness[s] = social_h;
        }
        const double health_h = is_sick(c, s) ? unit * -1 : unit;
        if (diseased) c.x->diseases.health_happiness[s] = health_h;
        else if (has_misc(c)) c.x->misc.health_happiness[s] = health_h;
        a.happiness[s] = conflict_h + family_h + health_h + social_h + wealth_h;
    } else {
        a.happiness[s] = family_h + 1.0 + wealth_h;
    }
    if (EXT && has_agentlog(c)) agentlog_record(c, s);        // updateRuntimeStats (agent.py:1567-1620)
}

}  // namespace ssb
