/* ss_oracle.c -- CPU restatement of the reference's per-timestep path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (sugarscape_b200/, sugarscape.py) may
 * import, link or execute this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and only as the checker / CPU baseline.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py replays every supported shipped example
 * of the reference (nkremerh/sugarscape) step by step and compares agents, grid and the
 * CPython MT19937 state with fixtures generated from the unmodified reference
 * (tests/golden/make_golden.py).
 *
 * What is restated (reference file:line):
 *   environment step   environment.py:61-109,192-211; cell.py:24-25,104-109,138-142
 *   list shuffle       sugarscape.py:400 ; random.py:350 (shuffle), random.py:242 (_randbelow)
 *   agent transaction  agent.py:568-598 and callees:
 *       findCellsInRange 766-779 (+ table order of environment.py:111-122)
 *       rankCellsInRange 1395-1443, findWelfare 1170-1201, sortCellsByWealth 1479-1490
 *       doCombat 274-294, gotoCell 1213-1219, collectResourcesAtCell 249-259
 *       doMetabolism 485-498, doTagging 556-566, doTrading 600-706
 *       doReproduction 500-554, findChildEndowment 781-910, doAging 266-272, doDeath 296-313
 *   removal            sugarscape.py:843-853
 *   reductions         sugarscape.py:913-934,1005-1426
 *
 * Plain, scalar, single-threaded C on the structs of include/ssb.h (host pointers).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/ssb.h"

/* ------------------------------------------------------------------------------------------
 * RNG.  Mode E: CPython's MT19937 (Modules/_randommodule.c genrand_uint32) and the Python
 * level algorithms random._randbelow_with_getrandbits / random.shuffle.
 * Mode S: counter-based words keyed by (seed, timestep, agent id, call site, draw index).
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    uint32_t mt[624];
    int32_t idx;
    /* mode S */
    int32_t mode;
    uint64_t step_key;
    uint32_t agent_id, site, counter;
} orc_rng_t;

/* call sites of the transaction that draw (agent.py:1401, 560-564, 616, 505-513): one substream each */
enum { SITE_MOVE = 0, SITE_TAGGING = 1, SITE_TRADING = 2, SITE_REPRODUCTION = 3, /* 4 = new tribe tags */
       SITE_LENDING = 5, SITE_DISEASE = 6 };

static uint32_t mt_next(orc_rng_t *r) {
    uint32_t *mt = r->mt, y;
    if (r->idx >= 624) {
        int kk;
        for (kk = 0; kk < 624 - 397; kk++) {
            y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
            mt[kk] = mt[kk + 397] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        for (; kk < 623; kk++) {
            y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
            mt[kk] = mt[kk + (397 - 624)] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        y = (mt[623] & 0x80000000u) | (mt[0] & 0x7fffffffu);
        mt[623] = mt[396] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        r->idx = 0;
    }
    y = mt[r->idx++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}

static uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

uint64_t orc_step_key(uint64_t seed, int32_t t) {
    return mix64(seed + 0x9E3779B97F4A7C15ull * (uint64_t)(uint32_t)(t + 1));
}

static uint32_t stream_word(uint64_t step_key, uint32_t id, uint32_t site, uint32_t j) {
    return (uint32_t)(mix64(step_key ^ (((uint64_t)id << 32) | ((uint64_t)site << 24) | j)) >> 32);
}

/* keyed bijection on [0, 2^bits): 4-round balanced Feistel; bits must be even */
uint32_t orc_priority(uint64_t step_key, uint32_t id, int bits) {
    int half = bits / 2;
    uint32_t mask = (1u << half) - 1u;
    uint32_t L = (id >> half) & mask, R = id & mask;
    for (uint32_t round = 0; round < 4; round++) {
        uint32_t f = (uint32_t)(mix64(step_key ^ 0xA5A5A5A5ull ^ ((uint64_t)(round + 1) << 56) ^ R) >> 40) & mask;
        uint32_t nl = R;
        R = L ^ f;
        L = nl;
    }
    return (L << half) | R;
}

static uint32_t rng_next32(orc_rng_t *r) {
    if (r->mode == SSB_RNG_EXACT) return mt_next(r);
    return stream_word(r->step_key, r->agent_id, r->site, r->counter++);
}

static void rng_at(orc_rng_t *r, uint32_t site) { r->site = site; r->counter = 0; }

static int bit_length(uint32_t n) { int k = 0; while (n) { k++; n >>= 1; } return k; }

/* random._randbelow_with_getrandbits: k = n.bit_length(); r = getrandbits(k) until r < n */
static uint32_t rng_below(orc_rng_t *r, uint32_t n) {
    int k = bit_length(n);
    uint32_t v = rng_next32(r) >> (32 - k);
    while (v >= n) v = rng_next32(r) >> (32 - k);
    return v;
}

/* random.random(): (a >> 5, b >> 6) -> 53 bits (Modules/_randommodule.c) */
static double rng_random(orc_rng_t *r) {
    uint32_t a = rng_next32(r) >> 5, b = rng_next32(r) >> 6;
    return (a * 67108864.0 + b) * (1.0 / 9007199254740992.0);
}

static void rng_shuffle_i32(orc_rng_t *r, int32_t *x, int n) {
    for (int i = n - 1; i >= 1; i--) {
        uint32_t j = rng_below(r, (uint32_t)i + 1u);
        int32_t tmp = x[i]; x[i] = x[j]; x[j] = tmp;
    }
}

/* Mode S, move ranking only: the KEYED shuffle -- the permutation that sorts the list by one substream word per element
 * (ties by position); mode E keeps random.shuffle.  (tests/golden/make_golden_scalable.py, Substreams.shuffle) */
static void rng_move_shuffle_i32(orc_rng_t *r, int32_t *x, int n) {
    if (r->mode == SSB_RNG_EXACT) { rng_shuffle_i32(r, x, n); return; }
    uint32_t key[256];
    int32_t out[256];
    for (int i = 0; i < n; i++) key[i] = rng_next32(r);
    for (int i = 0; i < n; i++) {
        int rank = 0;
        for (int j = 0; j < n; j++) if (key[j] < key[i] || (key[j] == key[i] && j < i)) rank++;
        out[rank] = x[i];
    }
    for (int i = 0; i < n; i++) x[i] = out[i];
}

/* ------------------------------------------------------------------------------------------
 * context
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    const ssb_params_t *p;
    ssb_grid_t g;
    ssb_agents_t a;
    orc_rng_t *rng;
    int32_t t;
    double prev_mean_wealth;
    const uint32_t *endow_bits; /* per-timestep child endowment choices (agent.py:839-851) */
    const uint32_t *tag_bits;   /* per-timestep tag mismatch draws (agent.py:859-875)     */
    int32_t max_dx, max_dy;
} ctx_t;

/* child endowment bit positions inside endow_bits[t] (host: sugarscape_b200/steptables.py) */
enum { EB_AGGRESSION = 0, EB_FERT_AGE, EB_FERT_FACTOR, EB_INFERT_AGE, EB_LOOKAHEAD, EB_MAX_AGE,
       EB_MOVEMENT, EB_SPICE_METAB, EB_SUGAR_METAB, EB_SEX, EB_TRADE_FACTOR, EB_VISION };


/* ------------------------------------------------------------------------------------------
 * diseases (condition.py:36-86; agent.py:184-188 addDisease, 227-247 catchDisease, 315-371 doDisease /
 * doInfectionAttempt, 1034-1050 nearest Hamming distance, 716-717 / 1031-1032 / 1107-1111 / 1161-1162
 * the "find" accessors with their modifiers).  ORACLE-PRIVATE state like lending and friends: the
 * engine refuses disease configurations; exact RNG mode only.
 * ---------------------------------------------------------------------------------------- */
enum { DM_AGGRESSION = 0, DM_FERTILITY, DM_FRIENDLINESS, DM_HAPPINESS, DM_MOVEMENT, DM_SPICE_METAB, DM_SUGAR_METAB,
       DM_VISION, DM_COUNT };
typedef struct {
    int32_t id, incubation, start_timestep, tag_len;
    uint32_t tags, pad;                       /* bit j = diseaseTags[j] */
    double transmission;
    double penalty[DM_COUNT];
    int64_t infected;                         /* len(disease.infected) */
    int64_t listed;                           /* how often sugarscape.diseases holds it (appended once per initial infection) */
} orc_disease_t;
typedef struct { int32_t disease, start, end, caught, incubation, infector_slot, infector_id; } orc_sick_t;
typedef struct { orc_sick_t *v; int n, cap; } sick_list_t;
typedef struct {
    int64_t capacity;
    int32_t immune_len, n_diseases;
    orc_disease_t *diseases;                  /* host array */
    uint64_t *immune, *start_immune;          /* per slot, bit i = immuneSystem[i] */
    double *protection;                       /* diseaseProtectionChance */
    double *modifier;                         /* [slot][DM_COUNT] */
    int32_t *disease_death, *last_spread, *n_spread;
    double *health_happiness;                 /* healthHappiness as of the agent's last updateHappiness */
    int32_t *last_valid_moves;                /* lastValidMoves: 1 when the range is empty (vision / movement pushed to 0) */
    const uint64_t *immune_bits;              /* per timestep: draws for the child's mismatching immune positions */
    sick_list_t *sick;                        /* owned here */
} orc_dis_t;
static orc_dis_t g_dis_store, *g_dis = NULL;
enum { EB_PROTECTION = 16 };                  /* steptables.ENDOWMENT_BITS */

static inline double dis_mod(int s, int which) { return g_dis ? g_dis->modifier[(size_t)s * DM_COUNT + which] : 0.0; }
static inline double pos_part(double v) { return v > 0 ? v : 0; }
/* findSugarMetabolism / findSpiceMetabolism / findVision / findMovement / findAggression */
#define EFF_SUGAR_METAB(c, s) pos_part((c)->a.sugar_metab[s] + dis_mod(s, DM_SUGAR_METAB))
#define EFF_SPICE_METAB(c, s) pos_part((c)->a.spice_metab[s] + dis_mod(s, DM_SPICE_METAB))
#define EFF_VISION(c, s)      pos_part((c)->a.vision[s] + dis_mod(s, DM_VISION))
#define EFF_MOVEMENT(c, s)    pos_part((c)->a.movement[s] + dis_mod(s, DM_MOVEMENT))
#define EFF_AGGRESSION(c, s)  pos_part((c)->a.aggression[s] + dis_mod(s, DM_AGGRESSION))
static inline int is_sick(int s) { return g_dis && g_dis->sick[s].n > 0; }


/* ------------------------------------------------------------------------------------------
 * universal income (agent.py:708-714, 1127-1134), racial tags (1082-1086; inheritance 879-887) and
 * depression (condition.py:13-34; agent.py:137-141, 889-895).  ORACLE-PRIVATE, exact RNG mode.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int64_t capacity;
    double *universal_sugar, *universal_spice;
    int32_t *last_income_sugar, *last_income_spice;
    int32_t income_interval_sugar, income_interval_spice;
    uint32_t *racial_tags;                    /* 2 bits per position */
    int32_t *race;                            /* -1 = None */
    int32_t racial_len, n_races;
    int32_t *depressed;
    double *happiness_unit;
    double depression_percentage;
    const uint32_t *racial_picks;             /* per timestep */
    const double *depression_draw;            /* per timestep */
} orc_misc_t;
static orc_misc_t g_misc_store, *g_misc = NULL;
enum { EB_UNIVERSAL_SPICE = 17, EB_UNIVERSAL_SUGAR = 18 };

void orc_bind_misc(const orc_misc_t *m) {
    if (!m) { g_misc = NULL; return; }
    g_misc_store = *m;
    g_misc = &g_misc_store;
}
int orc_sizeof_misc(void) { return (int)sizeof(orc_misc_t); }
static inline double happiness_unit(int s) { return g_misc ? g_misc->happiness_unit[s] : 1.0; }

/* experimental group = "depressed" (agent.py:1256-1284 isInGroup; sugarscape.py:998-1003): the log repeats
 * every statistic for the group and for its complement ("control").  g_pass selects which agents the
 * statistics functions below see: 0 = everybody, 1 = the group, 2 = the control group. */
enum { GI_COMBAT = 0, GI_DISEASE, GI_LENDING, GI_REPRODUCTION, GI_TRADE, GI_KINDS };
typedef struct {
    int64_t capacity;
    int32_t *with_control, *with_experimental;    /* [slot][GI_KINDS]: interactions since the last statistics pass */
    int32_t *control_neighbors, *experimental_neighbors;   /* movementNeighborhood split, as of the agent's last ranking */
} orc_group_t;
static orc_group_t g_group_store, *g_group = NULL;
static int g_pass = 0;
void orc_bind_group(const orc_group_t *g) { if (!g) { g_group = NULL; g_pass = 0; return; } g_group_store = *g; g_group = &g_group_store; }
int orc_sizeof_group(void) { return (int)sizeof(orc_group_t); }
void orc_set_pass(int32_t pass) { g_pass = pass; }
static inline int is_member(int s) { return g_misc && g_misc->depressed[s]; }
static inline int in_pass(int s) { return g_pass == 0 || (g_pass == 1) == is_member(s); }
static inline void group_note(int s, int kind, int partner) {      /* <kind>With{Experimental,Control}Group += 1 */
    if (!g_group) return;
    if (is_member(partner)) g_group->with_experimental[(size_t)s * GI_KINDS + kind]++;
    else g_group->with_control[(size_t)s * GI_KINDS + kind]++;
}


/* ------------------------------------------------------------------------------------------
 * ethics.Bentham decision models (ethics.py:105-244; spawn rules sugarscape.py:219-233).
 * ORACLE-PRIVATE, exact RNG mode; the in-group factors (ageism / racism / sexism / tribal) and the
 * dynamic selfishness are not restated (disabled, -1 / 0, in every shipped example).
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int64_t capacity;
    int32_t *model;                           /* 0 = plain Agent, 1 = Bentham class (altruist / bentham / egoist / negativeBentham) */
    double *selfishness, *decision_factor, *lookahead_factor, *lookahead_discount;
    int32_t *last_move_optimal, *move_rank;
    double *move_diff;                        /* welfare of the top cell minus welfare of the chosen one */
    double global_max_wealth;                 /* environment.globalMaxSugar + globalMaxSpice */
} orc_eth_t;
static orc_eth_t g_eth_store, *g_eth = NULL;
enum { EB_DECISION_MODEL = 19 };

void orc_bind_ethics(const orc_eth_t *e) {
    if (!e) { g_eth = NULL; return; }
    g_eth_store = *e;
    g_eth = &g_eth_store;
}
int orc_sizeof_ethics(void) { return (int)sizeof(orc_eth_t); }

static inline int wrapi(int v, int n) { v %= n; return v < 0 ? v + n : v; }

static int agent_alive(const ctx_t *c, int s) {
    return c->a.cod[s] == SSB_ALIVE && c->a.sugar[s] >= 0 && c->a.spice[s] >= 0;
}

static int popcount32(uint32_t v) { int n = 0; while (v) { n += v & 1u; v >>= 1; } return n; }

/* agent.py:1142-1151 */
static int find_tribe(const ctx_t *c, uint32_t tags) {
    int L = c->p->tag_len, T = c->p->max_tribes;
    if (L <= 0 || T <= 0) return -1;
    int zeroes = L - popcount32(tags);
    double tribe_size = (double)(L + 1) / (double)T;
    int tribe = (int)ceil((double)(zeroes + 1) / tribe_size) - 1;
    return tribe < T - 1 ? tribe : T - 1;
}

/* agent.py:1170-1201 */
static double find_welfare(const ctx_t *c, int s, double sugar_reward, double spice_reward) {
    double sm = EFF_SUGAR_METAB(c, s);
    double pm = EFF_SPICE_METAB(c, s);
    double total = sm + pm, sprop = 0, pprop = 0;
    if (total != 0) { sprop = sm / total; pprop = pm / total; }
    double look = c->a.lookahead[s];
    double ts = (c->a.sugar[s] + sugar_reward) - sm * look;
    double tp = (c->a.spice[s] + spice_reward) - pm * look;
    if (ts < 0) ts = 0;
    if (tp < 0) tp = 0;
    double welfare = pow(ts, sprop) * pow(tp, pprop);
    if ((c->p->flags & SSB_F_TAGPREF) && (c->p->flags & SSB_F_TAGS) && c->p->tag_len > 0) {
        int L = c->p->tag_len;
        c->a.tribe[s] = find_tribe(c, c->a.tags[s]);
        int zeroes = L - popcount32(c->a.tags[s]);
        double fz = (double)zeroes / (double)L, fo = 1 - fz;
        double pref = sm * fz + pm * fo;
        if (pref <= 0) pref = 1;
        welfare = pow(ts, (sm / pref) * fz) * pow(tp, (pm / pref) * fo);
    }
    return welfare;
}

/* kin lists: socialNetwork["children"] / ["mates"] of the initiating parent (agent.py:521-527) */
static void kin_push(ctx_t *c, int32_t *head, int owner, int other) {
    int64_t *cn = c->a.counters;
    int e;
    if (cn[SSB_C_KIN_FREE] != 0) {               /* recycled entry of a removed owner */
        e = (int)cn[SSB_C_KIN_FREE] - 1;
        cn[SSB_C_KIN_FREE] = c->a.kin_next[e] + 1;
    } else {
        if (cn[SSB_C_N_KIN] >= c->a.kin_capacity) { cn[SSB_C_OVERFLOW] = 3; return; }
        e = (int)cn[SSB_C_N_KIN]++;
    }
    c->a.kin_slot[e] = other; c->a.kin_id[e] = c->a.id[other]; c->a.kin_next[e] = head[owner];
    head[owner] = e;
}

static int agent_alive(const ctx_t *c, int s);

/* the relative of pool entry e, or -1 if it is no longer alive (dead, removed, slot recycled) */
static int kin_alive(const ctx_t *c, int e) {
    int o = c->a.kin_slot[e];
    return (c->a.id[o] == c->a.kin_id[e] && agent_alive(c, o)) ? o : -1;
}

/* Agent.doInheritance (agent.py:373-429): policies children / sons / daughters */
static void do_inheritance(ctx_t *c, int s) {
    int policy = c->p->inheritance_policy;
    if (policy == 0) return;
    if (c->a.sugar[s] < 0) c->a.sugar[s] = 0;
    if (c->a.spice[s] < 0) c->a.spice[s] = 0;
    int heirs = 0;
    for (int e = c->a.children_head[s]; e >= 0; e = c->a.kin_next[e]) {
        int o = kin_alive(c, e);
        if (o < 0) continue;
        if (policy == 1 || (policy == 2 && c->a.sex[o] == SSB_SEX_MALE) || (policy == 3 && c->a.sex[o] == SSB_SEX_FEMALE)) heirs++;
    }
    if (heirs == 0) return;
    double si = c->a.sugar[s] / heirs, pi = c->a.spice[s] / heirs;
    for (int e = c->a.children_head[s]; e >= 0; e = c->a.kin_next[e]) {
        int o = kin_alive(c, e);
        if (o < 0) continue;
        if (!(policy == 1 || (policy == 2 && c->a.sex[o] == SSB_SEX_MALE) || (policy == 3 && c->a.sex[o] == SSB_SEX_FEMALE))) continue;
        c->a.sugar[o] = c->a.sugar[o] + si;
        c->a.spice[o] = c->a.spice[o] + pi;
        c->a.sugar[s] -= si;
        c->a.spice[s] -= pi;
    }
}

static void lending_note_death(const ctx_t *c, int s);
static void disease_note_death(int s);
static void do_death(ctx_t *c, int s, int cause) {
    lending_note_death(c, s);
    disease_note_death(s);
    c->a.cod[s] = cause;
    int pos = c->a.pos[s];
    if (pos >= 0 && c->g.occ[pos] == s) c->g.occ[pos] = -1;
    c->a.pos[s] = -1;
    if (c->p->rng_mode == SSB_RNG_EXACT) do_inheritance(c, s);
}

/* agent.py:1023-1029 */
static void find_mrs(ctx_t *c, int s) {
    double pm = EFF_SPICE_METAB(c, s);
    double sm = EFF_SUGAR_METAB(c, s);
    double spice_need = pm > 0 ? c->a.spice[s] / pm : 1;
    double sugar_need = sm > 0 ? c->a.sugar[s] / sm : 1;
    c->a.mrs[s] = c->a.trade_factor[s] * (spice_need / sugar_need);
}

/* agent.py:1067-1080 */
static double find_new_mrs(const ctx_t *c, int s, double sugar, double spice) {
    double pm = EFF_SPICE_METAB(c, s);
    double sm = EFF_SUGAR_METAB(c, s);
    double spice_need = pm > 0 ? spice / pm : 1;
    double sugar_need = sm > 0 ? sugar / sm : 1;
    if (spice_need == 1 && sugar_need == 1) return 1;
    if (spice_need == 0) return pm;
    if (sugar_need == 0) return 1 / sm;
    return spice_need / sugar_need;
}

/* agent.py:1244-1247 */
static int is_fertile(const ctx_t *c, int s) {
    return c->a.sugar[s] >= c->a.start_sugar[s] && c->a.spice[s] >= c->a.start_spice[s] &&
           c->a.age[s] >= c->a.fert_age[s] && c->a.age[s] < c->a.infert_age[s] &&
           (c->a.fert_factor[s] + dis_mod(s, DM_FERTILITY)) > 0;
}

/* cell.neighbors in the reference's dict order (cell.py:61-89): N, S, E, W, then with neighborhoodMode "moore" (SSB_F_MOORE)
 * NE, NW, SE, SW = the east / west neighbours of the north and south cells; returns their number (out holds up to 8).
 * environmentWraparound = false (SSB_F_NOWRAP): cell.py:47-50, 98-121 -- the neighbours across the edge are missing, and the
 * reference's test for the south neighbour (`y + 1 < height - 1` -> None) leaves one to the last two rows only */
#define MAX_NEIGHBOURS 8
static int neighbours4(const ctx_t *c, int pos, int32_t *out) {
    int W = c->p->width, H = c->p->height, x = pos / H, y = pos % H;
    const int nowrap = (c->p->flags & SSB_F_NOWRAP) != 0;
    int north = -1, south = -1, n = 0;
    if (!(nowrap && y - 1 < 0)) north = x * H + (y - 1 + H) % H;
    if (!(nowrap && y + 1 < H - 1)) south = x * H + (y + 1 + H) % H;
    const int has_east = !(nowrap && x + 1 > W - 1), has_west = !(nowrap && x - 1 < 0);
    const int xe = (x + 1 + W) % W, xw = (x - 1 + W) % W;
    if (north >= 0) out[n++] = north;
    if (south >= 0) out[n++] = south;
    if (has_east) out[n++] = xe * H + y;
    if (has_west) out[n++] = xw * H + y;
    if (c->p->flags & SSB_F_MOORE) {      /* north.findEastNeighbor() ...: the east / west test is the same for the whole column */
        if (north >= 0 && has_east) out[n++] = xe * H + north % H;
        if (north >= 0 && has_west) out[n++] = xw * H + north % H;
        if (south >= 0 && has_east) out[n++] = xe * H + south % H;
        if (south >= 0 && has_west) out[n++] = xw * H + south % H;
    }
    return n;
}

static int alloc_slot(ctx_t *c) {
    int64_t *cn = c->a.counters;
    int slot;
    if (cn[SSB_C_N_FREE] > 0) {
        slot = c->a.free_slots[--cn[SSB_C_N_FREE]];
    } else {
        if (cn[SSB_C_N_SLOTS] >= c->a.capacity) { cn[SSB_C_OVERFLOW] = 1; return -1; }
        slot = (int)cn[SSB_C_N_SLOTS]++;
    }
    return slot;
}


/* findTimeToLive (agent.py:1113-1140) with the disease modifiers and the universal income term */
static double time_to_live(const ssb_agents_t *a, int s, int age_limited) {
    const double sm = pos_part(a->sugar_metab[s] + dis_mod(s, DM_SUGAR_METAB)), pm = pos_part(a->spice_metab[s] + dis_mod(s, DM_SPICE_METAB));
    const double big = 9223372036854775807.0;            /* sys.maxsize */
    double st = sm > 0 ? a->sugar[s] / sm : big;
    double pt = pm > 0 ? a->spice[s] / pm : big;
    if (g_misc) {
        if (g_misc->universal_sugar[s] != 0) {
            const double income = (st * g_misc->universal_sugar[s]) / g_misc->income_interval_sugar;
            st = sm > 0 ? (a->sugar[s] + income) / sm : big;
        }
        if (g_misc->universal_spice[s] != 0) {
            const double income = (pt * g_misc->universal_spice[s]) / g_misc->income_interval_spice;
            pt = pm > 0 ? (a->spice[s] + income) / pm : big;
        }
    }
    double ttl = st < pt ? st : pt;
    if (age_limited) { const double lim = (double)(a->max_age[s] - a->age[s]); if (lim < ttl) ttl = lim; }
    return ttl;
}

static int32_t race_of(uint32_t tags, int len) {          /* findRace: most common tag, the smallest among ties */
    int cnt[4] = {0, 0, 0, 0};
    for (int i = 0; i < len; i++) cnt[(tags >> (2 * i)) & 3u]++;
    int best = 0;
    for (int r = 1; r < 4; r++) if (cnt[r] > cnt[best]) best = r;
    return best;
}

/* races census of the log (sugarscape.py:1167-1170, 1287-1288, 1312): out = {largestRace, size, remaining} */
void orc_race_census(const ssb_agents_t *a, int64_t out[3]) {
    out[0] = -1; out[1] = 0; out[2] = 0;
    if (!g_misc || g_misc->racial_len == 0) return;
    int64_t cnt[4] = {0, 0, 0, 0}, first[4] = {-1, -1, -1, -1};
    for (int64_t i = 0; i < a->counters[SSB_C_N_LIST]; i++) {
        const int r = g_misc->race[a->order[i]];
        if (r < 0 || !in_pass(a->order[i])) continue;
        if (cnt[r]++ == 0) first[r] = i;
    }
    for (int r = 0; r < 4; r++) {
        if (!cnt[r]) continue;
        out[2]++;
        /* max(races, key=races.get): the first key, in insertion order, with the largest count */
        if (cnt[r] > out[1] || (cnt[r] == out[1] && first[r] < first[out[0]])) { out[0] = r; out[1] = cnt[r]; }
    }
}

/* ------------------------------------------------------------------------------------------
 * friends + conflict happiness (agent.py:1499-1520 updateFriends, 1563-1565 / 1622-1632
 * updateNeighbors -> updateSocialNetwork, 1099-1105 findSocialHappiness, 912-918
 * findConflictHappiness).  ORACLE-PRIVATE like lending below: they only move the happiness keys of
 * the log; the engine reports social / conflict happiness as 0 and refuses "friends" configurations.
 * ---------------------------------------------------------------------------------------- */
typedef struct { int32_t slot, id, hd; } orc_friend_t;
typedef struct { orc_friend_t *v; int n, cap; } friend_list_t;
typedef struct {
    int64_t capacity;
    int32_t *max_friends, *last_combat;               /* per slot, host arrays */
    double *social_happiness, *conflict_happiness;
    friend_list_t *friends;                           /* owned here */
} orc_social_t;
static orc_social_t g_social_store, *g_social = NULL;
enum { EB_MAX_FRIENDS = 15 };                         /* steptables.ENDOWMENT_BITS */

void orc_bind_social(int32_t *max_friends, int32_t *last_combat, double *social_happiness, double *conflict_happiness,
                     int64_t capacity) {
    if (g_social) {
        for (int64_t i = 0; i < g_social->capacity; i++) free(g_social->friends[i].v);
        free(g_social->friends);
        g_social = NULL;
    }
    if (!max_friends) return;
    g_social = &g_social_store;
    g_social->capacity = capacity;
    g_social->max_friends = max_friends; g_social->last_combat = last_combat;
    g_social->social_happiness = social_happiness; g_social->conflict_happiness = conflict_happiness;
    g_social->friends = (friend_list_t *)calloc((size_t)capacity, sizeof(friend_list_t));
}

/* sums over the agent list for meanSocialHappiness / meanConflictHappiness (sugarscape.py:1147-1160) */
void orc_social_sums(const ssb_agents_t *a, double out[2]) {
    out[0] = out[1] = 0;
    if (!g_social) return;
    for (int64_t i = 0; i < a->counters[SSB_C_N_LIST]; i++) {
        int s = a->order[i];
        if (!in_pass(s)) continue;
        out[0] += g_social->social_happiness[s];
        out[1] += g_social->conflict_happiness[s];
    }
}

static void social_reset_slot(int s) {
    g_social->friends[s].n = 0;
    g_social->last_combat[s] = -1;
    g_social->social_happiness[s] = 0; g_social->conflict_happiness[s] = 0;
}
static void friend_append(friend_list_t *L, orc_friend_t x) {
    if (L->n == L->cap) { L->cap = L->cap ? 2 * L->cap : 8; L->v = (orc_friend_t *)realloc(L->v, sizeof(orc_friend_t) * (size_t)L->cap); }
    L->v[L->n++] = x;
}
static void friend_remove_first(friend_list_t *L, orc_friend_t x) {     /* list.remove: first entry with the same agent and distance */
    for (int q = 0; q < L->n; q++)
        if (L->v[q].slot == x.slot && L->v[q].id == x.id && L->v[q].hd == x.hd) {
            memmove(&L->v[q], &L->v[q + 1], sizeof(orc_friend_t) * (size_t)(L->n - q - 1)); L->n--; return;
        }
}
/* updateFriends (agent.py:1499-1520) */
static void update_friends(ctx_t *c, int s, int nb) {
    friend_list_t *L = &g_social->friends[s];
    const int hd = (c->p->flags & SSB_F_TAGS) ? popcount32(c->a.tags[s] ^ c->a.tags[nb]) : 0;
    const orc_friend_t entry = {nb, c->a.id[nb], hd};
    if (L->n < g_social->max_friends[s]) { friend_append(L, entry); return; }
    int max_hd = 0, have_max = 0;
    orc_friend_t worst = {0, 0, 0};
    for (int p = 0; p < L->n; p++) {
        const orc_friend_t f = L->v[p];
        if (f.id == entry.id) { friend_remove_first(L, f); friend_append(L, entry); return; }
        if (f.hd > max_hd) { worst = f; max_hd = f.hd; have_max = 1; }
    }
    if (max_hd > hd && have_max) { friend_remove_first(L, worst); friend_append(L, entry); }
}

/* Depression.trigger (condition.py:27-33), at the creation of a depressed agent (agent.py:137-141) */
void orc_depress(const ssb_agents_t *a, int32_t s) {
    g_misc->depressed[s] = 1;
    a->aggression[s] *= 1.145;
    if (g_social) g_social->max_friends[s] = (int32_t)ceil(g_social->max_friends[s] * 0.6333);
    g_misc->happiness_unit[s] *= 0.5763;
    /* v *= ceil(v * f), saturated at 2^30 (Python ints leave the i32 range within two depressed generations of
     * agentMovement 1000; the rules only compare these values with far smaller ones) */
    { int32_t *v[3] = {&a->movement[s], &a->spice_metab[s], &a->sugar_metab[s]}; const double f[3] = {0.429, 1.544, 1.544};
      for (int k = 0; k < 3; k++) { const double p = (double)*v[k] * ceil(*v[k] * f[k]); *v[k] = p >= 1073741824.0 ? 1073741824 : (int32_t)p; } }
}


/* ------------------------------------------------------------------------------------------
 * lending (agent.py:431-483 doLending, 190-212 addLoan*, 920-930 current debt, 1226-1245,
 * 1285-1296, 1330-1360 payDebt, 1445-1449 removeDebt, 1532-1541 updateLoans, 1549-1553 mean income).
 * ORACLE-PRIVATE state: the CUDA engine has no lending yet and refuses such
 * configurations; this restatement is the checker its port will be held against.  Exact RNG mode.
 * A loan is a record in two Python lists: the debtor's "creditors" and the creditor's "debtors";
 * both lists are kept per slot with Python's list semantics (remove = first equal element,
 * iteration by index over a list that changes underneath).  An agent is named by (slot, id):
 * a slot that was reused by a newborn no longer is the agent the loan refers to (= dead).
 * ---------------------------------------------------------------------------------------- */
typedef struct { int32_t other_slot, other_id; double sugar, spice; int32_t duration, origin; } orc_loan_t;
typedef struct { orc_loan_t *v; int n, cap; } loan_list_t;
typedef struct {
    int64_t capacity;
    double *lending_factor, *base_interest_rate;      /* per slot, host arrays */
    int32_t *loan_duration;
    double *sugar_mean_income, *spice_mean_income;
    int32_t *last_lended, *last_loans;
    loan_list_t *creditors, *debtors;                 /* per slot, owned here */
} orc_lending_t;
static orc_lending_t g_lend_store, *g_lend = NULL;
/* a creditor that died keeps its children list in Python (the loan holds the object): remember the
 * head of its kin list, which lives on in the pool after the slot was reused */
typedef struct { int32_t slot, id, children_head; } dead_creditor_t;
static dead_creditor_t *g_dead_creditors = NULL;
static int g_n_dead_creditors = 0, g_cap_dead_creditors = 0;
enum { EB_BASE_INTEREST = 12, EB_LENDING_FACTOR = 13, EB_LOAN_DURATION = 14 };   /* steptables.ENDOWMENT_BITS */

void orc_bind_lending(double *lending_factor, double *base_interest_rate, int32_t *loan_duration,
                      double *sugar_mean_income, double *spice_mean_income, int32_t *last_lended,
                      int32_t *last_loans, int64_t capacity) {
    if (g_lend) {
        for (int64_t i = 0; i < g_lend->capacity; i++) { free(g_lend->creditors[i].v); free(g_lend->debtors[i].v); }
        free(g_lend->creditors); free(g_lend->debtors);
        g_lend = NULL;
    }
    free(g_dead_creditors); g_dead_creditors = NULL; g_n_dead_creditors = g_cap_dead_creditors = 0;
    if (!lending_factor) return;                      /* unbind */
    g_lend = &g_lend_store;
    g_lend->capacity = capacity;
    g_lend->lending_factor = lending_factor; g_lend->base_interest_rate = base_interest_rate;
    g_lend->loan_duration = loan_duration;
    g_lend->sugar_mean_income = sugar_mean_income; g_lend->spice_mean_income = spice_mean_income;
    g_lend->last_lended = last_lended; g_lend->last_loans = last_loans;
    g_lend->creditors = (loan_list_t *)calloc((size_t)capacity, sizeof(loan_list_t));
    g_lend->debtors = (loan_list_t *)calloc((size_t)capacity, sizeof(loan_list_t));
}

/* loanVolume of the log (sugarscape.py:1130-1131): sum of lastLoans over the agents that lent in step t */
int64_t orc_loan_volume(const ssb_agents_t *a, int32_t t) {
    int64_t v = 0;
    if (!g_lend) return 0;
    for (int64_t i = 0; i < a->counters[SSB_C_N_LIST]; i++) {
        int s = a->order[i];
        if (in_pass(s) && g_lend->last_lended[s] == t) v += g_lend->last_loans[s];
    }
    return v;
}

static void loan_append(loan_list_t *L, orc_loan_t x) {
    if (L->n == L->cap) { L->cap = L->cap ? 2 * L->cap : 4; L->v = (orc_loan_t *)realloc(L->v, sizeof(orc_loan_t) * (size_t)L->cap); }
    L->v[L->n++] = x;
}
static int loan_equal(const orc_loan_t *x, const orc_loan_t *y) {
    return x->other_slot == y->other_slot && x->other_id == y->other_id && x->sugar == y->sugar && x->spice == y->spice &&
           x->duration == y->duration && x->origin == y->origin;
}
static void loan_remove_first(loan_list_t *L, const orc_loan_t *x) {      /* list.remove(x) */
    for (int q = 0; q < L->n; q++)
        if (loan_equal(&L->v[q], x)) { memmove(&L->v[q], &L->v[q + 1], sizeof(orc_loan_t) * (size_t)(L->n - q - 1)); L->n--; return; }
}
static void lending_reset_slot(int s) {               /* a new agent object starts with empty lists */
    g_lend->creditors[s].n = 0; g_lend->debtors[s].n = 0;
    g_lend->sugar_mean_income[s] = 1; g_lend->spice_mean_income[s] = 1;
    g_lend->last_lended[s] = -1; g_lend->last_loans[s] = 0;
}
static int loan_other_alive(const ctx_t *c, const orc_loan_t *l) {
    return c->a.id[l->other_slot] == l->other_id && agent_alive(c, l->other_slot);
}
static int is_fertile(const ctx_t *c, int s);

/* addLoanToAgent (agent.py:199-211) incl. addLoanFromAgent (190-197) */
static void add_loan(ctx_t *c, int creditor, int debtor, int origin, double sugar_principal, double sugar_loan,
                     double spice_principal, double spice_loan, int duration) {
    ssb_agents_t *a = &c->a;
    orc_loan_t owed = {creditor, a->id[creditor], sugar_loan, spice_loan, duration, origin};
    orc_loan_t due = {debtor, a->id[debtor], sugar_loan, spice_loan, duration, origin};
    loan_append(&g_lend->creditors[debtor], owed);
    loan_append(&g_lend->debtors[creditor], due);
    a->sugar[creditor] -= sugar_principal;
    a->spice[creditor] -= spice_principal;
    a->sugar[debtor] = a->sugar[debtor] + sugar_principal;
    a->spice[debtor] = a->spice[debtor] + spice_principal;
}

static void lending_note_death(const ctx_t *c, int s) {     /* called by do_death */
    if (!g_lend || g_lend->debtors[s].n == 0) return;
    if (g_n_dead_creditors == g_cap_dead_creditors) {
        g_cap_dead_creditors = g_cap_dead_creditors ? 2 * g_cap_dead_creditors : 64;
        g_dead_creditors = (dead_creditor_t *)realloc(g_dead_creditors, sizeof(dead_creditor_t) * (size_t)g_cap_dead_creditors);
    }
    const dead_creditor_t d = {s, c->a.id[s], c->a.children_head[s]};
    g_dead_creditors[g_n_dead_creditors++] = d;
}

static void add_loan(ctx_t *c, int creditor, int debtor, int origin, double sugar_principal, double sugar_loan,
                     double spice_principal, double spice_loan, int duration);

/* payDebtToCreditorChildren (agent.py:1362-1377): the loan is split among the dead creditor's living
 * children (the debtor itself excluded), each as a new one-step loan, in the children list's order */
static void pay_debt_to_creditor_children(ctx_t *c, int s, orc_loan_t loan) {
    ssb_agents_t *a = &c->a;
    int head = -1;
    if (a->id[loan.other_slot] == loan.other_id) head = a->children_head[loan.other_slot];
    else for (int i = g_n_dead_creditors - 1; i >= 0; i--)
        if (g_dead_creditors[i].slot == loan.other_slot && g_dead_creditors[i].id == loan.other_id) { head = g_dead_creditors[i].children_head; break; }
    int32_t kids[256]; int n = 0;
    for (int e = head; e >= 0 && n < 256; e = a->kin_next[e]) {          /* newest first (kin_push prepends) */
        const int o = a->kin_slot[e];
        if (a->id[o] == a->kin_id[e] && agent_alive(c, o) && o != s) kids[n++] = o;
    }
    if (n > 0) {
        const double sugar_rep = loan.sugar / n, spice_rep = loan.spice / n;
        for (int i = n - 1; i >= 0; i--) add_loan(c, kids[i], s, a->last_moved[s], 0, sugar_rep, 0, spice_rep, 1);   /* append order */
    }
    loan_remove_first(&g_lend->creditors[s], &loan);
    if (a->id[loan.other_slot] == loan.other_id) {
        const orc_loan_t mirror = {s, a->id[s], loan.sugar, loan.spice, loan.duration, loan.origin};
        loan_remove_first(&g_lend->debtors[loan.other_slot], &mirror);
    }
}

/* payDebt (agent.py:1330-1360); `loan` is an element of the creditors list of s */
static void pay_debt(ctx_t *c, int s, orc_loan_t loan) {
    ssb_agents_t *a = &c->a;
    const int cr = loan.other_slot;
    const orc_loan_t mirror = {s, a->id[s], loan.sugar, loan.spice, loan.duration, loan.origin};
    const int same_object = a->id[cr] == loan.other_id;       /* the creditor's record still exists */
    if (!loan_other_alive(c, &loan)) {
        if (c->p->inheritance_policy == 1) { pay_debt_to_creditor_children(c, s, loan); return; }
        loan_remove_first(&g_lend->creditors[s], &loan);
        if (same_object) loan_remove_first(&g_lend->debtors[cr], &mirror);
    } else if (a->sugar[s] - loan.sugar > 0 && a->spice[s] - loan.spice > 0) {
        a->sugar[s] -= loan.sugar;
        a->spice[s] -= loan.spice;
        a->sugar[cr] = a->sugar[cr] + loan.sugar;
        a->spice[cr] = a->spice[cr] + loan.spice;
        loan_remove_first(&g_lend->creditors[s], &loan);
        loan_remove_first(&g_lend->debtors[cr], &mirror);
    } else {
        const double sugar_payout = a->sugar[s] / 2, spice_payout = a->spice[s] / 2;
        const double sugar_left = loan.sugar - sugar_payout, spice_left = loan.spice - spice_payout;
        a->sugar[s] -= sugar_payout;
        a->spice[s] -= spice_payout;
        a->sugar[cr] = a->sugar[cr] + sugar_payout;
        a->spice[cr] = a->spice[cr] + spice_payout;
        const double rate = g_lend->lending_factor[cr] * g_lend->base_interest_rate[cr];
        const double new_sugar = sugar_left + (rate * sugar_left), new_spice = spice_left + (rate * spice_left);
        loan_remove_first(&g_lend->creditors[s], &loan);
        loan_remove_first(&g_lend->debtors[cr], &mirror);
        /* compounded on the previous loan, no new principal */
        add_loan(c, cr, s, a->last_moved[s], 0, new_sugar, 0, new_spice, g_lend->loan_duration[cr]);
    }
}

/* updateLoans (agent.py:1532-1541): Python for-loops over lists that shrink / grow underneath */
static void update_loans(ctx_t *c, int s) {
    loan_list_t *D = &g_lend->debtors[s], *C = &g_lend->creditors[s];
    for (int p = 0; p < D->n; p++) {
        orc_loan_t l = D->v[p];
        if (!loan_other_alive(c, &l)) loan_remove_first(D, &l);
    }
    for (int p = 0; p < C->n; p++) {
        orc_loan_t l = C->v[p];
        const int remaining = (c->a.last_moved[s] - l.origin) - l.duration;
        if (remaining == 0) pay_debt(c, s, l);
    }
}

static double current_debt(const loan_list_t *C, int spice) {       /* agent.py:920-930 */
    double d = 0;
    for (int i = 0; i < C->n; i++) d += (spice ? C->v[i].spice : C->v[i].sugar) / C->v[i].duration;
    return d;
}

/* doLending (agent.py:431-483) */
static void do_lending(ctx_t *c, int s) {
    ssb_agents_t *a = &c->a;
    update_loans(c, s);
    /* isLender (1285-1296) */
    if (g_lend->lending_factor[s] == 0) return;
    if (is_fertile(c, s) && (a->sugar[s] <= a->start_sugar[s] || a->spice[s] <= a->start_spice[s])) return;
    if (a->age[s] < a->fert_age[s]) return;
    double rate = g_lend->lending_factor[s] * g_lend->base_interest_rate[s];
    if (rate > 1) rate = 1;
    int32_t nb[MAX_NEIGHBOURS], borrowers[MAX_NEIGHBOURS];
    int nbw = 0;
    const int n_nb = neighbours4(c, a->pos[s], nb);
    for (int i = 0; i < n_nb; i++) {
        int o = c->g.occ[nb[i]];
        if (o < 0 || !agent_alive(c, o)) continue;
        /* isBorrower (1226-1229) */
        if (a->age[o] >= a->fert_age[o] && a->age[o] < a->infert_age[o] && !is_fertile(c, o)) borrowers[nbw++] = o;
    }
    rng_at(c->rng, SITE_LENDING);
    rng_shuffle_i32(c->rng, borrowers, nbw);
    const double pm = EFF_SPICE_METAB(c, s), sm = EFF_SUGAR_METAB(c, s);
    int loans = 0;
    for (int k = 0; k < nbw; k++) {
        const int b = borrowers[k];
        double max_sugar = a->sugar[s] / 2, max_spice = a->spice[s] / 2;
        if (is_fertile(c, s)) {
            max_sugar = a->sugar[s] - a->start_sugar[s] > 0 ? a->sugar[s] - a->start_sugar[s] : 0;
            max_spice = a->spice[s] - a->start_spice[s] > 0 ? a->spice[s] - a->start_spice[s] : 0;
        }
        if (max_sugar == 0 && max_spice == 0) return;           /* (skips the lastLoans bookkeeping, like the reference) */
        const double sugar_need = a->start_sugar[b] - a->sugar[b] > 0 ? a->start_sugar[b] - a->sugar[b] : 0;
        const double spice_need = a->start_spice[b] - a->spice[b] > 0 ? a->start_spice[b] - a->spice[b] : 0;
        const double sugar_principal = max_sugar < sugar_need ? max_sugar : sugar_need;
        const double spice_principal = max_spice < spice_need ? max_spice : spice_need;
        const double sugar_amount = sugar_principal + (sugar_principal * rate);
        const double spice_amount = spice_principal + (spice_principal * rate);
        if ((sugar_need == 0 && spice_need == 0) || (sugar_amount == 0 && spice_amount == 0)) continue;
        if (a->sugar[s] - sugar_principal <= sm || a->spice[s] - spice_principal <= pm) continue;
        /* isCreditWorthy (1231-1242) */
        const int duration = g_lend->loan_duration[s];
        if (duration == 0) continue;
        const double bpm = EFF_SPICE_METAB(c, b), bsm = EFF_SUGAR_METAB(c, b);
        const double sugar_cost = sugar_amount / duration, spice_cost = spice_amount / duration;
        const double sugar_income = ((g_lend->sugar_mean_income[b] - bsm) - current_debt(&g_lend->creditors[b], 0)) - sugar_cost;
        const double spice_income = ((g_lend->spice_mean_income[b] - bpm) - current_debt(&g_lend->creditors[b], 1)) - spice_cost;
        if (!(sugar_income >= 0 && spice_income >= 0)) continue;
        add_loan(c, s, b, a->last_moved[s], sugar_principal, sugar_amount, spice_principal, spice_amount, duration);
        loans++;
        group_note(s, GI_LENDING, b);
    }
    if (loans > 0) { g_lend->last_lended[s] = c->t; g_lend->last_loans[s] = loans; }
}


/* ---- disease runtime ------------------------------------------------------------------------ */
void orc_bind_disease(orc_disease_t *diseases, int32_t n_diseases, int32_t immune_len, uint64_t *immune,
                      uint64_t *start_immune, double *protection, double *modifier, int32_t *disease_death,
                      int32_t *last_spread, int32_t *n_spread, double *health_happiness, int32_t *last_valid_moves,
                      const uint64_t *immune_bits, int64_t capacity) {
    if (g_dis) {
        for (int64_t i = 0; i < g_dis->capacity; i++) free(g_dis->sick[i].v);
        free(g_dis->sick);
        g_dis = NULL;
    }
    if (!immune) return;
    g_dis = &g_dis_store;
    g_dis->capacity = capacity; g_dis->immune_len = immune_len; g_dis->n_diseases = n_diseases;
    g_dis->diseases = diseases; g_dis->immune = immune; g_dis->start_immune = start_immune;
    g_dis->protection = protection; g_dis->modifier = modifier; g_dis->disease_death = disease_death;
    g_dis->last_spread = last_spread; g_dis->n_spread = n_spread; g_dis->immune_bits = immune_bits;
    g_dis->health_happiness = health_happiness; g_dis->last_valid_moves = last_valid_moves;
    g_dis->sick = (sick_list_t *)calloc((size_t)capacity, sizeof(sick_list_t));
}
int orc_sizeof_disease(void) { return (int)sizeof(orc_disease_t); }
int32_t orc_disease_count(int32_t slot) { return g_dis ? g_dis->sick[slot].n : 0; }

static void sick_append(sick_list_t *L, orc_sick_t x) {
    if (L->n == L->cap) { L->cap = L->cap ? 2 * L->cap : 8; L->v = (orc_sick_t *)realloc(L->v, sizeof(orc_sick_t) * (size_t)L->cap); }
    L->v[L->n++] = x;
}
static void disease_reset_slot(int s) {
    g_dis->sick[s].n = 0;
    for (int k = 0; k < DM_COUNT; k++) g_dis->modifier[(size_t)s * DM_COUNT + k] = 0;
    g_dis->disease_death[s] = 0; g_dis->last_spread[s] = -1; g_dis->n_spread[s] = 0;
    g_dis->health_happiness[s] = 0; g_dis->last_valid_moves[s] = 0;
}
static void disease_trigger(int s, const orc_disease_t *d) {      /* condition.py:57-65 */
    for (int k = 0; k < DM_COUNT; k++) g_dis->modifier[(size_t)s * DM_COUNT + k] += d->penalty[k];
}
static void disease_recover(int s, orc_disease_t *d) {            /* condition.py:67-76 */
    for (int k = 0; k < DM_COUNT; k++) g_dis->modifier[(size_t)s * DM_COUNT + k] -= d->penalty[k];
    d->infected--;
}
static int infected_with(int s, int disease_id) {                 /* agent.py:1249-1254 */
    const sick_list_t *L = &g_dis->sick[s];
    for (int i = 0; i < L->n; i++) if (g_dis->diseases[L->v[i].disease].id == disease_id) return 1;
    return 0;
}
/* findNearestHammingDistanceInDisease (agent.py:1034-1050) */
int32_t orc_disease_nearest(int32_t s, int32_t di, int32_t *start, int32_t *end) {
    const orc_disease_t *d = &g_dis->diseases[di];
    const int len = d->tag_len;
    int best = len, b0 = 0, b1 = len - 1;
    for (int i = 0; i < g_dis->immune_len - len; i++) {
        const uint32_t window = (uint32_t)((g_dis->immune[s] >> i) & ((len >= 32 ? 0xffffffffull : ((1ull << len) - 1ull))));
        const int hd = popcount32(window ^ d->tags);
        if (hd < best) { best = hd; b0 = i; b1 = i + (len - 1); }
    }
    if (start) *start = b0;
    if (end) *end = b1;
    return best;
}
/* catchDisease (agent.py:227-247); infector slot -1 = none (initial infections, timestep given) */
void orc_disease_catch(int32_t s, int32_t di, int32_t infector, int32_t infector_id, int32_t timestep, int32_t initial,
                       orc_rng_t *rng) {
    orc_disease_t *d = &g_dis->diseases[di];
    if (infected_with(s, d->id)) return;
    orc_sick_t rec = {di, -1, -1, timestep, d->incubation, infector, infector_id};
    {   /* generated diseases always carry a tag list (possibly empty: distance 0 = immune); tags None is the premade zombie virus only */
        int32_t st, en;
        if (orc_disease_nearest(s, di, &st, &en) == 0) return;        /* immune */
        rec.start = st; rec.end = en;
    }
    int caught = initial;
    if (!initial) {                            /* doInfectionAttempt (agent.py:363-371): two random.uniform(0, 1) */
        const double attack = 0.0 + (1.0 - 0.0) * rng_random(rng), defense = 0.0 + (1.0 - 0.0) * rng_random(rng);
        const int attack_ok = attack <= d->transmission && d->transmission != 0.0;
        const int defense_ok = defense <= g_dis->protection[s] && g_dis->protection[s] != 0.0;
        caught = attack_ok && !defense_ok;
    }
    if (caught) {                              /* addDisease (agent.py:184-188) */
        sick_append(&g_dis->sick[s], rec);
        d->infected++;
        if (rec.incubation == 0) disease_trigger(s, d);
    }
}
void orc_disease_listed(int32_t di) { g_dis->diseases[di].listed++; }

/* doDisease (agent.py:315-361) */
static void do_disease(ctx_t *c, int s) {
    sick_list_t *L = &g_dis->sick[s];
    rng_at(c->rng, SITE_DISEASE);          /* scalable mode: every draw of doDisease, infection attempts included */
    /* random.shuffle(self.diseases) */
    for (int i = L->n - 1; i >= 1; i--) {
        uint32_t j = rng_below(c->rng, (uint32_t)i + 1u);
        orc_sick_t tmp = L->v[i]; L->v[i] = L->v[j]; L->v[j] = tmp;
    }
    for (int p = 0; p < L->n; p++) {           /* Python for-loop over a list that may lose the current element */
        orc_sick_t *rec = &L->v[p];
        orc_disease_t *d = &g_dis->diseases[rec->disease];
        if (rec->caught != c->t && rec->incubation > 0) rec->incubation -= 1;
        if (rec->incubation == 0) disease_trigger(s, d);           /* (every step: the penalties pile up) */
        if (d->tag_len > 0) {
            const int st = rec->start;
            int en = rec->end + 1;
            if (en > g_dis->immune_len) en = g_dis->immune_len;
            const int len = en - st;
            const uint64_t mask = len >= 64 ? ~0ull : ((1ull << len) - 1ull);
            const uint64_t response = (g_dis->immune[s] >> st) & mask;     /* the slice, taken before the flip */
            for (int i = 0; i < len; i++)
                if (((response >> i) & 1ull) != ((d->tags >> i) & 1u)) {
                    g_dis->immune[s] = (g_dis->immune[s] & ~(1ull << (st + i))) | ((uint64_t)((d->tags >> i) & 1u) << (st + i));
                    break;
                }
            if (len == d->tag_len && response == (uint64_t)d->tags) {      /* diseaseTags == immuneResponse */
                const orc_sick_t gone = *rec;
                /* list.remove(record): first equal record -- records are unique per disease */
                memmove(&L->v[p], &L->v[p + 1], sizeof(orc_sick_t) * (size_t)(L->n - p - 1));
                L->n--;
                disease_recover(s, &g_dis->diseases[gone.disease]);
            }
        }
    }
    const int count = L->n;
    if (count == 0) return;
    int32_t nb4[MAX_NEIGHBOURS], neighbors[MAX_NEIGHBOURS];
    int nn = 0;
    const int n_nb = neighbours4(c, c->a.pos[s], nb4);
    for (int i = 0; i < n_nb; i++) { int o = c->g.occ[nb4[i]]; if (o >= 0 && agent_alive(c, o)) neighbors[nn++] = o; }
    int spread = 0;
    rng_shuffle_i32(c->rng, neighbors, nn);
    for (int k = 0; k < nn; k++) {
        const int di = L->v[rng_below(c->rng, (uint32_t)count)].disease;
        orc_disease_catch(neighbors[k], di, s, c->a.id[s], c->t, 0, c->rng);
        if (infected_with(neighbors[k], g_dis->diseases[di].id)) spread++;
        group_note(s, GI_DISEASE, neighbors[k]);
    }
    if (spread > 0) { g_dis->last_spread[s] = c->t; g_dis->n_spread[s] = spread; }
}

static void disease_note_death(int s) {        /* doDeath (agent.py:302-313) */
    if (!g_dis) return;
    if (g_dis->sick[s].n > 0) g_dis->disease_death[s] = 1;
    for (int i = 0; i < g_dis->sick[s].n; i++) disease_recover(s, &g_dis->diseases[g_dis->sick[s].v[i].disease]);
    g_dis->sick[s].n = 0;
}

int64_t orc_move_space(const ssb_agents_t *a) {      /* moveSpace (sugarscape.py:1152) */
    int64_t v = 0;
    if (g_dis) for (int64_t i = 0; i < a->counters[SSB_C_N_LIST]; i++) if (in_pass(a->order[i])) v += g_dis->last_valid_moves[a->order[i]];
    return v;
}

/* out = {sum selfishness, optimal moves, moves, sum move rank, sum move difference} (sugarscape.py:1141-1160, 1218-1222) */
void orc_ethics_stats(const ssb_agents_t *a, int32_t t, double out[5]) {
    for (int k = 0; k < 5; k++) out[k] = 0;
    if (!g_eth) return;
    for (int64_t i = 0; i < a->counters[SSB_C_N_LIST]; i++) {
        const int s = a->order[i];
        if (!in_pass(s)) continue;
        out[0] += g_eth->selfishness[s];
        out[1] += g_eth->last_move_optimal[s]; out[2] += 1;
        out[3] += g_eth->move_rank[s]; out[4] += g_eth->move_diff[s];
    }
    /* this step's dead: doTimestep(t) was called on every one of them (it sets agent.timestep even when the
     * agent is already dead), so the reference's `agent.timestep == self.timestep` holds for all */
    (void)t;
    for (int64_t i = 0; i < a->counters[SSB_C_N_DEAD]; i++)
        if (in_pass(a->dead_list[i])) { out[1] += g_eth->last_move_optimal[a->dead_list[i]]; out[2] += 1; }
}

/* Interaction counters of the current pass (sugarscape.py:1175-1198, 1230-1250): out[k] = to control,
 * out[GI_KINDS + k] = to experimental, summed over the pass's live agents (then reset) and this
 * step's dead; out[10], out[11] = control / experimental members of the movement neighbourhoods. */
void orc_group_counters(const ssb_agents_t *a, int64_t out[12]) {
    for (int k = 0; k < 12; k++) out[k] = 0;
    if (!g_group) return;
    for (int64_t i = 0; i < a->counters[SSB_C_N_LIST]; i++) {
        const int s = a->order[i];
        if (!in_pass(s)) continue;
        for (int k = 0; k < GI_KINDS; k++) {
            out[k] += g_group->with_control[(size_t)s * GI_KINDS + k];
            out[GI_KINDS + k] += g_group->with_experimental[(size_t)s * GI_KINDS + k];
            if (g_pass != 0) { g_group->with_control[(size_t)s * GI_KINDS + k] = 0; g_group->with_experimental[(size_t)s * GI_KINDS + k] = 0; }
        }
        out[10] += g_group->control_neighbors[s]; out[11] += g_group->experimental_neighbors[s];
    }
    for (int64_t i = 0; i < a->counters[SSB_C_N_DEAD]; i++) {
        const int s = a->dead_list[i];
        if (!in_pass(s)) continue;
        for (int k = 0; k < GI_KINDS; k++) {
            out[k] += g_group->with_control[(size_t)s * GI_KINDS + k];
            out[GI_KINDS + k] += g_group->with_experimental[(size_t)s * GI_KINDS + k];
        }
    }
}

/* a slot handed to a brand-new agent by the HOST (exact-mode replacement): empty lists, default fields */
void orc_private_reset(int32_t s) {
    if (g_lend) lending_reset_slot(s);
    if (g_social) social_reset_slot(s);
    if (g_dis) disease_reset_slot(s);
    if (g_group) {
        for (int k = 0; k < GI_KINDS; k++) { g_group->with_control[(size_t)s * GI_KINDS + k] = 0; g_group->with_experimental[(size_t)s * GI_KINDS + k] = 0; }
        g_group->control_neighbors[s] = 0; g_group->experimental_neighbors[s] = 0;
    }
    if (g_eth) { g_eth->last_move_optimal[s] = 1; g_eth->move_rank[s] = 0; g_eth->move_diff[s] = 0; }
    if (g_misc) {
        g_misc->last_income_sugar[s] = 0; g_misc->last_income_spice[s] = 0;
        g_misc->depressed[s] = 0; g_misc->happiness_unit[s] = 1; g_misc->race[s] = -1; g_misc->racial_tags[s] = 0;
    }
}

double orc_health_happiness_sum(const ssb_agents_t *a) {
    double v = 0;
    if (g_dis) for (int64_t i = 0; i < a->counters[SSB_C_N_LIST]; i++) if (in_pass(a->order[i])) v += g_dis->health_happiness[a->order[i]];
    return v;
}

/* log keys (sugarscape.py:1162, 1200-1211, 1228, 1254-1277, 1316-1317): out = {sickAgents, diseaseIncidence,
 * distinct infectors, diseasePrevalence, agentDiseaseDeaths} */
void orc_disease_stats(const ssb_agents_t *a, int32_t t, int64_t out[5]) {
    for (int k = 0; k < 5; k++) out[k] = 0;
    if (!g_dis) return;
    const int64_t n = a->counters[SSB_C_N_LIST], n_dead = a->counters[SSB_C_N_DEAD];
    int64_t *infectors = (int64_t *)malloc(sizeof(int64_t) * (size_t)(4 * (n + n_dead) + 16));
    int64_t ni = 0;
    /* isDiseaseExperimentalGroup() is false for every disease unless the group names one, so the pass over the
     * experimental group skips all incidence / prevalence bookkeeping (sugarscape.py:1200-1211, 1266-1277) */
    const int skip_diseases = g_pass == 1;
    for (int64_t i = 0; i < n; i++) {
        const int s = a->order[i];
        if (!in_pass(s)) continue;
        const sick_list_t *L = &g_dis->sick[s];
        if (L->n > 0) out[0]++;
        for (int k = 0; k < L->n && !skip_diseases; k++)
            if (L->v[k].caught == t) {
                out[1]++;
                if (t != 0) infectors[ni++] = ((int64_t)L->v[k].infector_slot << 32) | (uint32_t)L->v[k].infector_id;
            }
    }
    for (int64_t i = 0; i < n_dead; i++) if (in_pass(a->dead_list[i])) out[4] += g_dis->disease_death[a->dead_list[i]];    /* (their disease lists are empty) */
    /* distinct infectors */
    for (int64_t i = 0; i < ni; i++) {
        int seen = 0;
        for (int64_t j = 0; j < i && !seen; j++) seen = infectors[j] == infectors[i];
        out[2] += !seen;
    }
    free(infectors);
    for (int d = 0; d < g_dis->n_diseases && !skip_diseases; d++) out[3] += g_dis->diseases[d].listed * g_dis->diseases[d].infected;
}

/* collectResourcesAtCell, agent.py:249-259 + cell.py:32-45 */
static void collect(ctx_t *c, int s, int pos) {
    int32_t cs = c->g.sugar[pos], cp = c->g.spice[pos];
    c->a.sugar[s] += cs;
    c->a.spice[s] += cp;
    if (g_lend) {      /* updateMeanIncome (agent.py:1549-1553) */
        const double alpha = 0.05;
        g_lend->sugar_mean_income[s] = (alpha * cs) + ((1 - alpha) * g_lend->sugar_mean_income[s]);
        g_lend->spice_mean_income[s] = (alpha * cp) + ((1 - alpha) * g_lend->spice_mean_income[s]);
    }
    if (c->p->pollution_start <= c->t && c->t <= c->p->pollution_end) {
        c->g.pollution[pos] += c->p->sugar_prod_pollution * cs;
        c->g.pollution[pos] += c->p->spice_prod_pollution * cp;
    }
    c->g.sugar[pos] = 0;
    c->g.spice[pos] = 0;
}

/* candidate record of rankCellsInRange */
typedef struct { int32_t cell; int32_t dist; double welfare; } cand_t;

/* candidate arrays: a cross holds at most 4 * 64 cells; a radial range can hold a whole small map -- no silent truncation in
 * the checker (the engine stops such a run, overflow code 15).  Sized per call (C11 variable-length arrays), so that the
 * cross keeps the small frames the CPU baseline of bench.py is timed with. */
#define MAX_CAND_CROSS 256
#define MAX_CAND_RADIAL 4096
#define MAX_CAND(c) (((c)->p->flags & SSB_F_RADIAL) ? MAX_CAND_RADIAL : MAX_CAND_CROSS)

/* agent.py:766-779 with the insertion order of environment.py:111-122 */
static int enumerate_disc(const ctx_t *c, int pos, int r, int32_t *cells, int32_t *dists);
static int enumerate_cross(const ctx_t *c, int pos, int r, int32_t *cells, int32_t *dists) {
    int W = c->p->width, H = c->p->height, x = pos / H, y = pos % H, n = 0;
    if (c->p->flags & SSB_F_RADIAL) return enumerate_disc(c, pos, r, cells, dists);
    for (int d = 1; d <= r; d++) {
        int64_t key[4]; int32_t cell[4]; int m = 0;
        if (d <= c->max_dx) {
            int xw = wrapi(x - d, W), xe = wrapi(x + d, W);
            key[m] = ((int64_t)xw * H + y) * 4 + 0; cell[m++] = xw * H + y;      /* west  */
            if (xe != xw) { key[m] = ((int64_t)x * H + y) * 4 + 1; cell[m++] = xe * H + y; } /* east */
            else if ((int64_t)x * H + y < (int64_t)xw * H + y) key[m - 1] = ((int64_t)x * H + y) * 4 + 1;
        }
        if (d <= c->max_dy) {
            int yn = wrapi(y - d, H), ys = wrapi(y + d, H);
            key[m] = ((int64_t)x * H + yn) * 4 + 0; cell[m++] = x * H + yn;      /* north */
            if (ys != yn) { key[m] = ((int64_t)x * H + y) * 4 + 2; cell[m++] = x * H + ys; } /* south */
            else if ((int64_t)x * H + y < (int64_t)x * H + yn) key[m - 1] = ((int64_t)x * H + y) * 4 + 2;
        }
        for (int i = 1; i < m; i++)      /* insertion sort by key */
            for (int j = i; j > 0 && key[j - 1] > key[j]; j--) {
                int64_t tk = key[j]; key[j] = key[j - 1]; key[j - 1] = tk;
                int32_t tc = cell[j]; cell[j] = cell[j - 1]; cell[j - 1] = tc;
            }
        for (int i = 0; i < m && n < MAX_CAND(c); i++) { cells[n] = cell[i]; dists[n++] = d; }
    }
    return n;
}


/* agentVisionMode = agentMovementMode = "radial" (SSB_F_RADIAL): Environment.findRadialCellRanges (environment.py:157-173)
 * walks every pair (i, j > i) of the x-major cell list and files each cell under ranges[floor(distance)] of the other, so
 * ranges[g] of a cell lists its partners in ascending order of their own list index; findCellsInRange (agent.py:766-779)
 * concatenates ranges[1 .. r].  Restated literally: one pass over the whole cell list per ring.  The travel distance
 * sqrt(dx^2 + dy^2) is only compared by its users (agent.py:1430, 1484): dists[] carries dx^2 + dy^2. */
static int wrap_distance(const ctx_t *c, int delta, int border) {          /* findWraparoundDistance (environment.py:175-179) */
    delta = abs(delta);
    if (!(c->p->flags & SSB_F_NOWRAP) && 2 * delta > border) delta = border - delta;
    return delta;
}
/* maxDeltaX / maxDeltaY of findCellRanges (environment.py:138-144): width - 1 instead of width // 2 without wraparound */
static int radial_cap(const ctx_t *c, int border) {
    const int far = (c->p->flags & SSB_F_NOWRAP) ? border - 1 : border / 2;
    return c->p->max_agent_range < far ? c->p->max_agent_range : far;
}
static int enumerate_disc(const ctx_t *c, int pos, int r, int32_t *cells, int32_t *dists) {
    const int W = c->p->width, H = c->p->height, x1 = pos / H, y1 = pos % H;
    int n = 0;
    for (int g = 1; g <= r; g++)
        for (int j = 0; j < W * H; j++) {
            const int dx = wrap_distance(c, x1 - j / H, W), dy = wrap_distance(c, y1 - j % H, H);
            if (j == pos || dx > radial_cap(c, W) || dy > radial_cap(c, H)) continue;
            if ((int)floor(sqrt((double)(dx * dx + dy * dy))) != g) continue;
            if (n < MAX_CAND(c)) { cells[n] = j; dists[n++] = dx * dx + dy * dy; }
        }
    return n;
}
/* Environment.maxCellDistance (environment.py:140-146) */
static int max_cell_distance(const ctx_t *c) {
    if (c->p->flags & SSB_F_RADIAL) {
        const int nowrap = (c->p->flags & SSB_F_NOWRAP) != 0;
        const int hw = nowrap ? c->p->width - 1 : c->p->width / 2, hh = nowrap ? c->p->height - 1 : c->p->height / 2;
        const int lim = (int)floor(sqrt((double)(hw * hw + hh * hh)));
        return c->p->max_agent_range < lim ? c->p->max_agent_range : lim;
    }
    return c->max_dx > c->max_dy ? c->max_dx : c->max_dy;
}

/* ---- Bentham: findBestEthicalCell / findEthicalValueOfCell ------------------------------------ */
static int agent_range(const ctx_t *c, int s) {           /* findCellsInRange (agent.py:766-779) */
    int vision = (int)EFF_VISION(c, s), movement = (int)EFF_MOVEMENT(c, s);
    int maxd = max_cell_distance(c);
    int r = vision < movement ? vision : movement;
    return r > maxd ? maxd : r;
}
static int cell_in_cross(const ctx_t *c, int centre, int r, int cell) {
    const int W = c->p->width, H = c->p->height;
    const int x = centre / H, y = centre % H, cx = cell / H, cy = cell % H;
    if (cell == centre || r <= 0) return 0;
    if (c->p->flags & SSB_F_RADIAL) {
        const int dx = wrap_distance(c, cx - x, W), dy = wrap_distance(c, cy - y, H);
        return dx <= radial_cap(c, W) && dy <= radial_cap(c, H) && (int)floor(sqrt((double)(dx * dx + dy * dy))) <= r;
    }
    if (cx == x) { int d = abs(cy - y); if (H - d < d) d = H - d; return d <= (r < c->max_dy ? r : c->max_dy); }
    if (cy == y) { int d = abs(cx - x); if (W - d < d) d = W - d; return d <= (r < c->max_dx ? r : c->max_dx); }
    return 0;
}
/* value of `cell` for the agents in view (ethics.py:141-244); selfishness < 0: happiness / unhappiness split */
static double ethical_value(ctx_t *c, int s, int my_r, const int32_t *hood, int n_hood, int cell, double *happy, double *unhappy) {
    const ssb_agents_t *a = &c->a;
    double site = (double)(c->g.sugar[cell] + c->g.spice[cell]);
    const double max_loot = c->p->max_combat_loot * 2;
    double max_site = (double)(c->g.sugar_cap[cell] + c->g.spice_cap[cell]);
    const int occupant = c->g.occ[cell];
    if (occupant >= 0) {
        const double w = a->sugar[occupant] + a->spice[occupant];
        site += w < max_loot ? w : max_loot;
        max_site += w < max_loot ? w : max_loot;
    }
    int32_t nb4[MAX_NEIGHBOURS];
    const int n_nb = neighbours4(c, cell, nb4);
    double neighbor_wealth = 0;
    for (int i = 0; i < n_nb; i++) neighbor_wealth += c->g.sugar[nb4[i]] + c->g.spice[nb4[i]];
    const double lookahead = g_eth->lookahead_factor[s];
    /* len(self.findNeighborhood(cell)): live agents in MY cross moved to `cell`, plus myself */
    double future_hood = 1;
    if (lookahead != 0) {
        int32_t fc[MAX_CAND(c)], fd[MAX_CAND(c)];
        const int fn = my_r > 0 ? enumerate_cross(c, cell, my_r, fc, fd) : 0;
        int cnt = 1;
        for (int i = 0; i < fn; i++) { int o = c->g.occ[fc[i]]; if (o >= 0 && agent_alive(c, o)) cnt++; }
        future_hood = cnt;
    }
    const double selfish = g_eth->selfishness[s];
    double value = 0, h = 0, u = 0;
    for (int k = 0; k < n_hood; k++) {
        const int nb = hood[k];
        if (!agent_alive(c, nb)) continue;
        const int rn = agent_range(c, nb);
        if (!(cell == a->pos[nb] || cell_in_cross(c, a->pos[nb], rn, cell))) continue;      /* canReachCell */
        const double metab = (double)(a->sugar_metab[nb] + a->spice_metab[nb]);            /* raw, not the find* accessors */
        const double cell_duration = metab > 0 ? site / metab : 0;
        const double proximity = 1.0 / 1;
        const double intensity = (1 / (1 + time_to_live(a, nb, 0)) / (1 + c->g.pollution[cell]));
        const double duration = max_site > 0 ? cell_duration / max_site : 0;
        const double discount = g_eth->lookahead_factor[nb] != 0 ? g_eth->lookahead_discount[nb] : 0;
        double future_duration = metab > 0 ? (site - metab) / metab : site;
        future_duration = max_site > 0 ? future_duration / max_site : 0;
        int32_t nb_cells[MAX_NEIGHBOURS];
        const double future_intensity = neighbor_wealth / (g_eth->global_max_wealth * neighbours4(c, a->pos[nb], nb_cells));   /* len(neighbor.cell.neighbors) */
        int32_t tc[MAX_CAND(c)], td[MAX_CAND(c)];
        const int in_range = rn > 0 ? enumerate_cross(c, a->pos[nb], rn, tc, td) : 0;          /* len(neighbor.cellsInRange) */
        const double extent = in_range > 0 ? (double)n_hood / in_range : 1;
        const double future_extent = (in_range > 0 && lookahead != 0) ? future_hood / in_range : 1;
        const double current_reward = extent * (intensity + duration);
        const double future_reward = future_extent * (future_intensity + future_duration);
        double v = (1 * proximity) * (current_reward + (discount * future_reward));
        if (nb != s && selfish < 1) {
            v = -1 * v;
            if (cell == a->pos[nb] && v > -1) v = -1;
        }
        if (selfish >= 0) v *= nb == s ? selfish : 1 - selfish;
        else { if (v > 0) h += v; else u += v; }
        value += v;
    }
    *happy = h; *unhappy = u;
    return value;
}

/* agent.py:1395-1443 + 719-746 + 1321-1328: returns the chosen cell */
static int move_to_best_cell(ctx_t *c, int s) {
    const ssb_agents_t *a = &c->a;
    int pos = a->pos[s];
    int vision = (int)EFF_VISION(c, s), movement = (int)EFF_MOVEMENT(c, s);
    int maxd = max_cell_distance(c);
    int r = vision < movement ? vision : movement;
    if (r > maxd) r = maxd;
    int32_t cells[MAX_CAND(c)], dists[MAX_CAND(c)];
    int n = r > 0 ? enumerate_cross(c, pos, r, cells, dists) : 0;
    double aggression = EFF_AGGRESSION(c, s);
    int best = pos;
    if (n == 0) {
        /* empty range: stays put; validMoves / movementNeighborhood keep their old values, but
         * findBestCell still records lastValidMoves = len([own cell]) (agent.py:1398-1399, 745) */
        if (g_dis) g_dis->last_valid_moves[s] = 1;
        if (g_eth) g_eth->last_move_optimal[s] = 1;      /* [own cell] is its own best; validMoves (rank, difference) keep their values */
    } else {
        /* findNeighborhood: live agents in range + self (agent.py:1052-1065) */
        int hood = 1;
        int hood_exp = is_member(s), hood_ctrl = !is_member(s);      /* self is part of the neighbourhood */
        for (int i = 0; i < n; i++) {
            int o = c->g.occ[cells[i]];
            if (o >= 0 && agent_alive(c, o)) { hood++; if (is_member(o)) hood_exp++; else hood_ctrl++; }
        }
        if (g_group) { g_group->control_neighbors[s] = hood_ctrl; g_group->experimental_neighbors[s] = hood_exp; }
        /* random.shuffle(list(cellsInRange.items())) -- shuffle an index permutation */
        int32_t perm[MAX_CAND(c)];
        for (int i = 0; i < n; i++) perm[i] = i;
        rng_at(c->rng, SITE_MOVE);
        rng_move_shuffle_i32(c->rng, perm, n);
        /* findRetaliatorsInVision (agent.py:1088-1098): richest visible member per tribe */
        double retaliator[34]; int have[34];
        for (int i = 0; i < 34; i++) { retaliator[i] = 0; have[i] = 0; }
        if (aggression > 0)
            for (int i = 0; i < n; i++) {
                int o = c->g.occ[cells[i]];
                if (o < 0) continue;
                int tr = a->tribe[o] + 1; /* -1 (None) -> 0 */
                double w = a->sugar[o] + a->spice[o];
                if (!have[tr] || retaliator[tr] < w) { retaliator[tr] = w; have[tr] = 1; }
            }
        cand_t cand[MAX_CAND(c)]; int m = 0;
        double loot = c->p->max_combat_loot;
        double my_wealth = a->sugar[s] + a->spice[s];
        for (int k = 0; k < n; k++) {
            int cell = cells[perm[k]], dist = dists[perm[k]];
            int prey = c->g.occ[cell];
            double ps = 0, pp = 0;
            if (prey >= 0) {
                /* isNeighborValidPrey (agent.py:1309-1314) */
                if (!(aggression > 0 && a->tribe[s] != a->tribe[prey] &&
                      my_wealth >= a->sugar[prey] + a->spice[prey])) continue;
                ps = a->sugar[prey]; pp = a->spice[prey];
            }
            double wps = aggression * (loot < ps ? loot : ps);
            double wpp = aggression * (loot < pp ? loot : pp);
            double poll = c->g.pollution[cell];
            double welfare = find_welfare(c, s, (c->g.sugar[cell] + wps) / (1 + poll),
                                          (c->g.spice[cell] + wpp) / (1 + poll));
            if (prey >= 0 && retaliator[a->tribe[prey] + 1] > my_wealth + welfare) continue;
            cand[m].cell = cell; cand[m].dist = dist; cand[m].welfare = welfare; m++;
        }
        if (m == 0) { cand[0].cell = pos; cand[0].dist = 0; cand[0].welfare = 0; m = 1; }
        /* sortCellsByWealth: stable insertion sort, welfare desc then range asc (1479-1490) */
        for (int i = 0; i < m; i++)
            for (int j = i; j > 0 && (cand[j - 1].welfare < cand[j].welfare ||
                 (cand[j - 1].welfare == cand[j].welfare && cand[j - 1].dist > cand[j].dist)); j--) {
                cand_t tmp = cand[j]; cand[j] = cand[j - 1]; cand[j - 1] = tmp;
            }
        int rank = 0;
        if (g_eth && g_eth->model[s] == 1 && g_eth->decision_factor[s] > 0) {
            /* findBestEthicalCell (ethics.py:105-139): re-score the welfare-sorted cells */
            int32_t hoodv[MAX_CAND(c) + 1]; int nh = 0;
            for (int i = 0; i < n; i++) { int o = c->g.occ[cells[i]]; if (o >= 0 && agent_alive(c, o)) hoodv[nh++] = o; }
            hoodv[nh++] = s;
            int chosen = -1;
            double best_u = 0, best_h = 0;
            for (int i = 0; i < m; i++) {
                double h, u;
                const double v = ethical_value(c, s, r, hoodv, nh, cand[i].cell, &h, &u);
                if (g_eth->selfishness[s] >= 0) { if (v > 0) { chosen = i; break; } }
                else if (chosen < 0 || u > best_u || (u == best_u && h > best_h)) { chosen = i; best_u = u; best_h = h; }   /* stable sort, reverse */
            }
            if (chosen >= 0) rank = chosen;
        }
        best = cand[rank].cell;
        if (g_eth) {
            g_eth->last_move_optimal[s] = rank == 0; g_eth->move_rank[s] = rank;
            g_eth->move_diff[s] = cand[0].welfare - cand[rank].welfare;
        }
        a->n_neighborhood[s] = hood;
        a->n_valid_moves[s] = m;
        if (g_dis) g_dis->last_valid_moves[s] = m;
    }
    /* moveToBestCell: doCombat when aggressive (agent.py:274-294), then gotoCell (1213-1219) */
    if (aggression > 0) {
        int prey = c->g.occ[best];
        if (prey >= 0 && prey != s) {
            double loot = c->p->max_combat_loot;
            double sl = loot < a->sugar[prey] ? loot : a->sugar[prey];
            double pl = loot < a->spice[prey] ? loot : a->spice[prey];
            a->sugar[s] += sl; a->spice[s] += pl;
            a->sugar[prey] -= sl; a->spice[prey] -= pl;
            if (g_social) g_social->last_combat[s] = c->t;
            do_death(c, prey, SSB_COMBAT);
            group_note(s, GI_COMBAT, prey);
        }
    }
    a->last_moved[s] = c->t;
    c->g.occ[pos] = -1;
    a->pos[s] = best;
    c->g.occ[best] = s;
    return best;
}

/* agent.py:556-566 */
static void do_tagging(ctx_t *c, int s) {
    if (!(c->p->flags & SSB_F_TAGS) || !(c->p->flags & SSB_F_TAGGING)) return;
    int32_t nb[MAX_NEIGHBOURS];
    const int n_nb = neighbours4(c, c->a.pos[s], nb);
    rng_at(c->rng, SITE_TAGGING);
    rng_shuffle_i32(c->rng, nb, n_nb);
    for (int i = 0; i < n_nb; i++) {
        int o = c->g.occ[nb[i]];
        if (o < 0) continue;
        uint32_t position = rng_below(c->rng, (uint32_t)c->p->tag_len);
        uint32_t bit = 1u << position;
        c->a.tags[o] = (c->a.tags[o] & ~bit) | (c->a.tags[s] & bit);
        c->a.tribe[o] = find_tribe(c, c->a.tags[o]);
    }
}

/* agent.py:600-706 */
static void do_trading(ctx_t *c, int s) {
    ssb_agents_t *a = &c->a;
    if (a->trade_factor[s] == 0) return;
    a->trade_volume[s] = 0;
    double sum_sugar_price = 0, sum_spice_price = 0;
    find_mrs(c, s);
    int32_t nb[MAX_NEIGHBOURS], traders[MAX_NEIGHBOURS]; int n = 0;
    const int n_nb = neighbours4(c, a->pos[s], nb);
    for (int i = 0; i < n_nb; i++) {
        int o = c->g.occ[nb[i]];
        if (o >= 0 && agent_alive(c, o) && a->mrs[o] != a->mrs[s]) traders[n++] = o;
    }
    rng_at(c->rng, SITE_TRADING);
    rng_shuffle_i32(c->rng, traders, n);
    double sugar_price = 0, spice_price = 0;
    for (int k = 0; k < n; k++) {
        int tr = traders[k];
        int spice_seller = -1, sugar_seller = -1;
        for (;;) {
            double tm = a->mrs[tr], my = a->mrs[s];
            if ((tm >= 1 && my >= 1) || (tm < 1 && my < 1) || tm == my) break;
            if (tm > my) { spice_seller = tr; sugar_seller = s; }
            else { spice_seller = s; sugar_seller = tr; }
            double ssm = a->mrs[spice_seller], sum = a->mrs[sugar_seller];
            if (ssm < 0 || sum < 0) { spice_seller = sugar_seller = -1; break; }
            double price = sqrt(ssm * sum);
            if (price < 1) { spice_price = 1; sugar_price = price; }
            else { spice_price = price; sugar_price = 1; }
            if (a->spice[spice_seller] - spice_price < a->spice_metab[spice_seller] ||
                a->sugar[sugar_seller] - sugar_price < a->sugar_metab[sugar_seller]) break;
            double ss_new = find_new_mrs(c, spice_seller, a->sugar[spice_seller] + sugar_price,
                                         a->spice[spice_seller] - spice_price);
            double su_new = find_new_mrs(c, sugar_seller, a->sugar[sugar_seller] - sugar_price,
                                         a->spice[sugar_seller] + spice_price);
            int better_ss_mrs = fabs(1 - ssm) > fabs(1 - ss_new);
            int better_su_mrs = fabs(1 - sum) > fabs(1 - su_new);
            int better_ss_w = find_welfare(c, spice_seller, sugar_price, -1 * spice_price) >=
                              find_welfare(c, spice_seller, 0, 0);
            int better_su_w = find_welfare(c, sugar_seller, -1 * sugar_price, spice_price) >=
                              find_welfare(c, sugar_seller, 0, 0);
            int crossing = ss_new < su_new;
            if ((better_ss_mrs || better_ss_w) && (better_su_mrs || better_su_w) && !crossing) {
                a->sugar[spice_seller] += sugar_price; a->spice[spice_seller] -= spice_price;
                a->sugar[sugar_seller] -= sugar_price; a->spice[sugar_seller] += spice_price;
                find_mrs(c, spice_seller);
                find_mrs(c, sugar_seller);
            } else break;
        }
        if (spice_seller >= 0 && sugar_seller >= 0) {
            a->trade_volume[s] += 1;
            sum_sugar_price += sugar_price;
            sum_spice_price += spice_price;
            group_note(s, GI_TRADE, traders[k]);
        }
    }
    a->trade_price[s] = sum_spice_price > sum_sugar_price ? sum_spice_price : sum_sugar_price;
}

static int empty_neighbours(const ctx_t *c, int pos, int32_t *out) {
    int32_t nb[MAX_NEIGHBOURS]; int n = 0;
    const int n_nb = neighbours4(c, pos, nb);
    for (int i = 0; i < n_nb; i++) if (c->g.occ[nb[i]] < 0) out[n++] = nb[i];
    return n;
}

/* agent.py:500-554 with findChildEndowment 781-910 and addChildToCell 164-182 */
static void do_reproduction(ctx_t *c, int s) {
    ssb_agents_t *a = &c->a;
    if (!agent_alive(c, s) || !is_fertile(c, s)) return;
    int32_t nb[MAX_NEIGHBOURS], own_empty[MAX_NEIGHBOURS];
    const int n_nb = neighbours4(c, a->pos[s], nb);
    rng_at(c->rng, SITE_REPRODUCTION);
    rng_shuffle_i32(c->rng, nb, n_nb);
    int n_own = empty_neighbours(c, a->pos[s], own_empty);
    for (int k = 0; k < n_nb; k++) {
        int m = c->g.occ[nb[k]];
        if (m < 0 || !agent_alive(c, m)) continue;
        int compatible = ((a->sex[s] == SSB_SEX_FEMALE && a->sex[m] == SSB_SEX_MALE) ||
                          (a->sex[s] == SSB_SEX_MALE && a->sex[m] == SSB_SEX_FEMALE)) && is_fertile(c, m);
        int32_t pool[2 * MAX_NEIGHBOURS]; int np = 0;
        for (int i = 0; i < n_own; i++) pool[np++] = own_empty[i];
        np += empty_neighbours(c, a->pos[m], pool + np);
        rng_shuffle_i32(c->rng, pool, np);   /* drawn before the compatibility test (quirk C.7) */
        if (!(is_fertile(c, s) && compatible && np != 0)) continue;
        int cell = pool[--np];
        while (c->g.occ[cell] >= 0 && np != 0) cell = pool[--np];
        if (c->g.occ[cell] >= 0) continue;
        int ch = alloc_slot(c);
        if (ch < 0) return;
        uint32_t eb = c->endow_bits ? c->endow_bits[c->t] : 0;
#define PICK(field, bit) (((eb >> (bit)) & 1u) ? a->field[m] : a->field[s])
        int64_t *cn = a->counters;
        a->id[ch] = c->p->rng_mode == SSB_RNG_EXACT ? (int32_t)cn[SSB_C_NEXT_ID]++ : -1;
        a->born[ch] = c->t; a->cod[ch] = SSB_ALIVE; a->age[ch] = 0;
        a->aggression[ch] = PICK(aggression, EB_AGGRESSION);
        a->fert_age[ch] = PICK(fert_age, EB_FERT_AGE);
        a->fert_factor[ch] = PICK(fert_factor, EB_FERT_FACTOR);
        a->infert_age[ch] = PICK(infert_age, EB_INFERT_AGE);
        a->lookahead[ch] = PICK(lookahead, EB_LOOKAHEAD);
        a->max_age[ch] = PICK(max_age, EB_MAX_AGE);
        a->movement[ch] = PICK(movement, EB_MOVEMENT);
        a->spice_metab[ch] = PICK(spice_metab, EB_SPICE_METAB);
        a->sugar_metab[ch] = PICK(sugar_metab, EB_SUGAR_METAB);
        a->sex[ch] = PICK(sex, EB_SEX);
        a->trade_factor[ch] = PICK(trade_factor, EB_TRADE_FACTOR);
        a->vision[ch] = PICK(vision, EB_VISION);
        if (g_dis) {     /* immune system: equal bits copied, mismatches drawn from the per-timestep stream (agent.py:897-908) */
            disease_reset_slot(ch);
            /* startingImmuneSystem IS immuneSystem in the reference (agent.py:33,50 bind the same list), so the
             * parents pass on what their immune responses have learnt */
            const uint64_t mine = g_dis->immune[s], other = g_dis->immune[m], draws = g_dis->immune_bits ? g_dis->immune_bits[c->t] : 0;
            uint64_t child = 0; int draw = 0;
            for (int i = 0; i < g_dis->immune_len; i++) {
                const uint64_t bs = (mine >> i) & 1ull, bm = (other >> i) & 1ull;
                const uint64_t b = bs == bm ? bs : ((draws >> draw++) & 1ull);
                child |= b << i;
            }
            g_dis->immune[ch] = g_dis->start_immune[ch] = child;
            g_dis->protection[ch] = ((eb >> EB_PROTECTION) & 1u) ? g_dis->protection[m] : g_dis->protection[s];
        }
        if (g_group) {
            for (int k = 0; k < GI_KINDS; k++) { g_group->with_control[(size_t)ch * GI_KINDS + k] = 0; g_group->with_experimental[(size_t)ch * GI_KINDS + k] = 0; }
            g_group->control_neighbors[ch] = 0; g_group->experimental_neighbors[ch] = 0;
        }
        if (g_eth) {     /* the paired endowments follow ONE draw (agent.py:845-851); the child's class is that parent's */
            const int from = ((eb >> EB_DECISION_MODEL) & 1u) ? m : s;
            g_eth->model[ch] = g_eth->model[from]; g_eth->selfishness[ch] = g_eth->selfishness[from];
            g_eth->decision_factor[ch] = g_eth->decision_factor[from];
            g_eth->lookahead_factor[ch] = g_eth->lookahead_factor[from];
            g_eth->lookahead_discount[ch] = g_eth->lookahead_discount[from];
            g_eth->last_move_optimal[ch] = 1; g_eth->move_rank[ch] = 0; g_eth->move_diff[ch] = 0;      /* lastMoveOptimal starts True */
        }
        if (g_misc) {
            g_misc->universal_spice[ch] = ((eb >> EB_UNIVERSAL_SPICE) & 1u) ? g_misc->universal_spice[m] : g_misc->universal_spice[s];
            g_misc->universal_sugar[ch] = ((eb >> EB_UNIVERSAL_SUGAR) & 1u) ? g_misc->universal_sugar[m] : g_misc->universal_sugar[s];
            g_misc->last_income_sugar[ch] = 0; g_misc->last_income_spice[ch] = 0;
            g_misc->depressed[ch] = 0; g_misc->happiness_unit[ch] = 1;
            g_misc->race[ch] = -1; g_misc->racial_tags[ch] = 0;
            if (g_misc->racial_len > 0) {      /* random.choice([mine, mate's]) per position (agent.py:885-886) */
                const uint32_t picks = g_misc->racial_picks[c->t];
                uint32_t rt = 0;
                for (int i = 0; i < g_misc->racial_len; i++) {
                    const uint32_t src = ((picks >> i) & 1u) ? g_misc->racial_tags[m] : g_misc->racial_tags[s];
                    rt |= ((src >> (2 * i)) & 3u) << (2 * i);
                }
                g_misc->racial_tags[ch] = rt;
                g_misc->race[ch] = race_of(rt, g_misc->racial_len);
            }
        }
        if (g_social) {
            social_reset_slot(ch);
            g_social->max_friends[ch] = ((eb >> EB_MAX_FRIENDS) & 1u) ? g_social->max_friends[m] : g_social->max_friends[s];
        }
        if (g_lend) {
            lending_reset_slot(ch);
            g_lend->base_interest_rate[ch] = ((eb >> EB_BASE_INTEREST) & 1u) ? g_lend->base_interest_rate[m] : g_lend->base_interest_rate[s];
            g_lend->lending_factor[ch] = ((eb >> EB_LENDING_FACTOR) & 1u) ? g_lend->lending_factor[m] : g_lend->lending_factor[s];
            g_lend->loan_duration[ch] = ((eb >> EB_LOAN_DURATION) & 1u) ? g_lend->loan_duration[m] : g_lend->loan_duration[s];
        }
#undef PICK
        double sugar_cost = a->start_sugar[s] / (a->fert_factor[s] * 2);
        double spice_cost = a->start_spice[s] / (a->fert_factor[s] * 2);
        double mate_sugar_cost = a->start_sugar[m] / (a->fert_factor[m] * 2);
        double mate_spice_cost = a->start_spice[m] / (a->fert_factor[m] * 2);
        a->start_sugar[ch] = a->sugar[ch] = sugar_cost + mate_sugar_cost;
        a->start_spice[ch] = a->spice[ch] = spice_cost + mate_spice_cost;
        /* tags: equal bits copied, mismatches drawn in order from the per-timestep stream */
        uint32_t tags = 0;
        if (c->p->flags & SSB_F_TAGS) {
            uint32_t tb = c->tag_bits ? c->tag_bits[c->t] : 0; int draw = 0;
            for (int i = 0; i < c->p->tag_len; i++) {
                uint32_t bs = (a->tags[s] >> i) & 1u, bm = (a->tags[m] >> i) & 1u;
                uint32_t b = bs == bm ? bs : ((tb >> draw++) & 1u);
                tags |= b << i;
            }
        }
        a->tags[ch] = tags;
        a->tribe[ch] = find_tribe(c, tags);
        a->mrs[ch] = 1;
        a->last_moved[ch] = c->t; a->last_sugar[ch] = 0; a->last_spice[ch] = 0;
        a->n_neighborhood[ch] = 0; a->n_valid_moves[ch] = 0;
        a->happiness[ch] = 0; a->wealth_happiness[ch] = 0; a->family_happiness[ch] = 0;
        a->trade_volume[ch] = 0; a->trade_price[ch] = 0; a->last_repro[ch] = -1;
        a->children_head[ch] = -1; a->mates_head[ch] = -1;
        a->father[ch] = a->sex[s] == SSB_SEX_FEMALE ? a->id[m] : a->id[s];
        a->mother[ch] = a->sex[s] == SSB_SEX_FEMALE ? a->id[s] : a->id[m];
        /* the child is depressed with the configured chance, decided by the private stream (agent.py:889-895) */
        if (g_misc && g_misc->depression_draw[c->t] <= g_misc->depression_percentage) orc_depress(a, ch);
        a->pos[ch] = cell;
        c->g.occ[cell] = ch;
        if (cn[SSB_C_N_LIST] < a->capacity) a->order[cn[SSB_C_N_LIST]++] = ch;   /* addAgent */
        else cn[SSB_C_OVERFLOW] = 1;
        cn[SSB_C_N_BORN]++;
        collect(c, ch, cell);                                                    /* child collects */
        if (c->p->rng_mode == SSB_RNG_EXACT) {   /* kin bookkeeping of the initiating parent (agent.py:521-527) */
            int known = 0;
            for (int e = a->mates_head[s]; e >= 0; e = a->kin_next[e])
                if (a->kin_slot[e] == m && a->kin_id[e] == a->id[m]) known = 1;
            if (!known) kin_push(c, a->mates_head, s, m);
            kin_push(c, a->children_head, s, ch);
        }
        a->sugar[s] -= sugar_cost; a->spice[s] -= spice_cost;
        a->sugar[m] = a->sugar[m] - mate_sugar_cost;
        a->spice[m] = a->spice[m] - mate_spice_cost;
        a->last_repro[s] = c->t;
        group_note(s, GI_REPRODUCTION, m);
    }
}

/* agent.py:568-598 */
static void agent_transaction(ctx_t *c, int s) {
    ssb_agents_t *a = &c->a;
    if (!agent_alive(c, s) || a->last_moved[s] == c->t) return;
    a->last_sugar[s] = a->sugar[s];
    a->last_spice[s] = a->spice[s];
    int pos = move_to_best_cell(c, s);
    if (g_social) {      /* updateNeighbors -> updateSocialNetwork -> updateFriends, neighbours in N, S, E, W order */
        int32_t nb4[MAX_NEIGHBOURS];
        const int n_nb = neighbours4(c, pos, nb4);
        for (int i = 0; i < n_nb; i++) { int o = c->g.occ[nb4[i]]; if (o >= 0) update_friends(c, s, o); }
    }
    collect(c, s, pos);
    if (g_misc) {        /* doUniversalIncome (agent.py:708-714): spice first */
        if ((c->t - g_misc->last_income_spice[s]) >= g_misc->income_interval_spice) {
            a->spice[s] += g_misc->universal_spice[s]; g_misc->last_income_spice[s] = c->t;
        }
        if ((c->t - g_misc->last_income_sugar[s]) >= g_misc->income_interval_sugar) {
            a->sugar[s] += g_misc->universal_sugar[s]; g_misc->last_income_sugar[s] = c->t;
        }
    }
    /* doMetabolism (agent.py:485-498) + consumption pollution (cell.py:27-40) */
    double sm = EFF_SUGAR_METAB(c, s);
    double pm = EFF_SPICE_METAB(c, s);
    a->sugar[s] -= sm;
    a->spice[s] -= pm;
    if (c->p->pollution_start <= c->t && c->t <= c->p->pollution_end) {
        c->g.pollution[pos] += c->p->sugar_cons_pollution * sm;
        c->g.pollution[pos] += c->p->spice_cons_pollution * pm;
    }
    if (a->sugar[s] < 0 || a->spice[s] < 0 || (a->sugar[s] <= 0 && sm > 0) || (a->spice[s] <= 0 && pm > 0)) {
        do_death(c, s, SSB_STARVATION);
        return;
    }
    do_tagging(c, s);
    do_trading(c, s);
    do_reproduction(c, s);
    if (g_lend) do_lending(c, s);
    if (g_dis) do_disease(c, s);
    /* doAging (agent.py:266-272) */
    if (agent_alive(c, s)) {
        a->age[s] += 1;
        if (a->age[s] >= a->max_age[s] && a->max_age[s] != -1) { do_death(c, s, SSB_AGING); return; }
    }
    /* updateHappiness (agent.py:1522-1530): conflict + family + health + social + wealth.
     * family/social/conflict bookkeeping is not restated yet -> 0 */
    const double unit = happiness_unit(s);                /* happinessUnit: 1, 0.5763 when depressed */
    /* findWealthHappiness (agent.py:1164-1168) */
    double diff_wealth = (a->sugar[s] + a->spice[s]) - c->prev_mean_wealth;
    if (g_misc) diff_wealth *= unit;
    double wealth_h = erf(diff_wealth);
    /* findFamilyHappiness (agent.py:940-958), mode E only (the lists are not kept in mode S) */
    double family = 0;
    if (c->p->rng_mode == SSB_RNG_EXACT) {
        for (int e = a->children_head[s]; e >= 0; e = a->kin_next[e]) {
            int o = kin_alive(c, e);
            if (o >= 0) { family += unit; if (is_sick(o)) family -= unit * 0.5; if (a->born[o] == c->t) family += unit; } else family -= unit;
        }
        for (int e = a->mates_head[s]; e >= 0; e = a->kin_next[e]) {
            int o = kin_alive(c, e);
            if (o >= 0) { family += unit; if (is_sick(o)) family -= unit * 0.5; } else family -= unit;
        }
    }
    double family_h = erf(family);
    double conflict_h = 0, social_h = 0;
    if (g_social) {
        if (g_social->last_combat[s] == c->t) conflict_h = EFF_AGGRESSION(c, s) > 1 ? unit : unit * -1;
        if (g_social->max_friends[s] != 0) social_h = ((g_social->friends[s].n * (2.0 / g_social->max_friends[s])) - 1) * unit;
        g_social->conflict_happiness[s] = conflict_h; g_social->social_happiness[s] = social_h;
    }
    a->wealth_happiness[s] = wealth_h;
    a->family_happiness[s] = family_h;
    const double health_h = is_sick(s) ? unit * -1 : unit;       /* findHealthHappiness (agent.py:1017-1021) */
    if (g_dis) g_dis->health_happiness[s] = health_h;
    a->happiness[s] = conflict_h + family_h + health_h + social_h + wealth_h;
}

/* ------------------------------------------------------------------------------------------
 * exported entry points (ctypes; tests/oracle_binding.py)
 * ---------------------------------------------------------------------------------------- */
static void make_ctx(ctx_t *c, const ssb_params_t *p, const ssb_grid_t *g, const ssb_agents_t *a,
                     orc_rng_t *rng, int32_t t) {
    memset(c, 0, sizeof(*c));
    c->p = p; c->g = *g; c->a = *a; c->rng = rng; c->t = t;
    /* environment.py:132-146: memoised range per axis (cardinal modes, wraparound);
     * max_agent_range carries max(maxVision, maxMovement) capped per axis below */
    int rmax = p->max_agent_range;
    c->max_dx = rmax < p->width / 2 ? rmax : p->width / 2;
    c->max_dy = rmax < p->height / 2 ? rmax : p->height / 2;
}

int orc_sizeof_rng(void) { return (int)sizeof(orc_rng_t); }

void orc_rng_init(orc_rng_t *r, const uint32_t state[624], int32_t index, int32_t mode) {
    memset(r, 0, sizeof(*r));
    if (state) memcpy(r->mt, state, sizeof(r->mt));
    r->idx = index; r->mode = mode;
}

void orc_rng_get(const orc_rng_t *r, uint32_t state[624], int32_t *index) {
    memcpy(state, r->mt, sizeof(r->mt));
    *index = r->idx;
}

/* environment.py:105-109 with the closed-form schedules of SURVEY Appendix A.3 */
void orc_env_step(const ssb_params_t *p, const ssb_grid_t *g, int32_t t) {
    int W = p->width, H = p->height;
    int seasons = p->season_interval > 0;
    int flipped = seasons ? (((t - 1) / p->season_interval) & 1) : 0;
    int dry_grows = seasons && p->seasonal_growback_delay > 0 ? (t % p->seasonal_growback_delay == 0) : 0;
    for (int x = 0; x < W; x++)
        for (int y = 0; y < H; y++) {
            int i = x * H + y, grow = 1;
            if (seasons) {
                int wet = (y < p->equator) != flipped;   /* north starts wet (environment.py:186-189) */
                grow = wet || dry_grows;
            }
            if (grow) {
                int32_t s = g->sugar[i] + p->sugar_regrow, q = g->spice[i] + p->spice_regrow;
                g->sugar[i] = s < g->sugar_cap[i] ? s : g->sugar_cap[i];
                g->spice[i] = q < g->spice_cap[i] ? q : g->spice_cap[i];
            }
        }
    /* the reference's countdown (environment.py:192-197) is decremented on every step inside the timeframe, the first
     * step being t = 1: it diffuses on the k-th such step when k is a multiple of the delay */
    if (p->diffusion_delay > 0 && p->diffusion_start <= t && t <= p->diffusion_end &&
        ((t - (p->diffusion_start > 1 ? p->diffusion_start : 1) + 1) % p->diffusion_delay) == 0) {
        for (int x = 0; x < W; x++)
            for (int y = 0; y < H; y++) {
                double m = 0;
                if (p->flags & (SSB_F_NOWRAP | SSB_F_MOORE)) {       /* cell.py:104-109 over cell.neighbors as built by cell.py:61-89 */
                    ctx_t cc; cc.p = p;
                    int32_t nbc[MAX_NEIGHBOURS];
                    const int n = neighbours4(&cc, x * H + y, nbc);
                    for (int k = 0; k < n; k++) m += g->pollution[nbc[k]];
                    g->pollution_flux[x * H + y] = m / n;
                    continue;
                }
                m += g->pollution[x * H + wrapi(y - 1, H)];
                m += g->pollution[x * H + wrapi(y + 1, H)];
                m += g->pollution[wrapi(x + 1, W) * H + y];
                m += g->pollution[wrapi(x - 1, W) * H + y];
                g->pollution_flux[x * H + y] = m / 4;
            }
        memcpy(g->pollution, g->pollution_flux, sizeof(double) * (size_t)W * H);
    }
}

static int cmp_u64(const void *x, const void *y) {
    uint64_t a = *(const uint64_t *)x, b = *(const uint64_t *)y;
    return a < b ? -1 : a > b;
}

/* sugarscape.py:400-407 */
void orc_agent_step(const ssb_params_t *p, const ssb_grid_t *g, const ssb_agents_t *a, orc_rng_t *rng,
                    int32_t t, double prev_mean_wealth, const uint32_t *endow_bits, const uint32_t *tag_bits) {
    ctx_t c;
    make_ctx(&c, p, g, a, rng, t);
    c.prev_mean_wealth = prev_mean_wealth;
    c.endow_bits = endow_bits; c.tag_bits = tag_bits;
    int64_t *cn = a->counters;
    cn[SSB_C_N_BORN] = 0;
    if (p->rng_mode == SSB_RNG_EXACT) {
        rng->mode = SSB_RNG_EXACT;
        rng_shuffle_i32(rng, a->order, (int)cn[SSB_C_N_LIST]);
        /* addRemainingDiseases (sugarscape.py:139-151, called right after the shuffle): a disease whose start has come goes to
         * the first agent of the list that is not immune to it -- no infection attempt, no infector -- and joins
         * sugarscape.diseases; one that finds nobody waits for the next step.  start_timestep > 0 marks the waiting ones. */
        if (g_dis)
            for (int d = 0; d < g_dis->n_diseases; d++) {
                if (g_dis->diseases[d].start_timestep <= 0 || g_dis->diseases[d].start_timestep > t) continue;
                for (int64_t i = 0; i < cn[SSB_C_N_LIST]; i++) {
                    if (orc_disease_nearest(a->order[i], d, NULL, NULL) == 0) continue;
                    orc_disease_catch(a->order[i], d, -1, -1, t, 1, NULL);
                    g_dis->diseases[d].listed++;
                    g_dis->diseases[d].start_timestep = 0;
                    break;
                }
            }
        /* the list may grow while iterating (children are appended and skipped) */
        for (int64_t i = 0; i < cn[SSB_C_N_LIST]; i++) agent_transaction(&c, a->order[i]);
    } else {
        int n = (int)cn[SSB_C_N_LIST];
        uint64_t key = orc_step_key(p->seed, t);
        uint64_t *sorted = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)(n > 0 ? n : 1));
        for (int i = 0; i < n; i++) {
            int s = a->order[i];
            sorted[i] = ((uint64_t)orc_priority(key, (uint32_t)a->id[s], p->id_bits) << 32) | (uint32_t)s;
        }
        qsort(sorted, (size_t)n, sizeof(uint64_t), cmp_u64);
        rng->mode = SSB_RNG_SCALABLE; rng->step_key = key;
        for (int i = 0; i < n; i++) {
            int s = (int)(sorted[i] & 0xffffffffu);
            rng->agent_id = (uint32_t)a->id[s]; rng->counter = 0;
            agent_transaction(&c, s);
        }
        free(sorted);
    }
}

static int cmp_i64(const void *x, const void *y) {
    int64_t a = *(const int64_t *)x, b = *(const int64_t *)y;
    return a < b ? -1 : a > b;
}

/* removeDeadAgents, sugarscape.py:843-853: order-preserving; dead slots are recycled.
 * Mode S additionally numbers this step's newborn in cell order ("mode S"). */
/* tests/hostemu: lending lives in the host-compiled device source there, and so do the dead creditors' children lists */
static int g_keep_children_lists = 0;
void orc_keep_children_lists(int on) { g_keep_children_lists = on; }

void orc_compact(const ssb_params_t *p, const ssb_grid_t *g, const ssb_agents_t *a, int32_t t) {
    ctx_t c;
    make_ctx(&c, p, g, a, NULL, t);
    int64_t *cn = a->counters;
    int64_t n = cn[SSB_C_N_LIST], w = 0, nd = 0;
    for (int64_t i = 0; i < n; i++) {
        int s = a->order[i];
        if (!agent_alive(&c, s) || a->pos[s] < 0) {
            if (a->cod[s] == SSB_ALIVE) do_death(&c, s, SSB_OTHER);
            a->dead_list[nd++] = s;
        } else a->order[w++] = s;
    }
    for (int64_t i = 0; i < nd; i++) a->free_slots[cn[SSB_C_N_FREE]++] = a->dead_list[i];
    cn[SSB_C_N_LIST] = w;
    cn[SSB_C_N_DEAD] = nd;
    if (p->rng_mode == SSB_RNG_EXACT && a->children_head && a->kin_next) {
        /* the kin lists of a removed agent die with it: their pool entries go to the free chain */
        int64_t free_head = cn[SSB_C_KIN_FREE] - 1;
        for (int64_t i = 0; i < nd; i++) {
            int s = a->dead_list[i];
            for (int list = (g_lend || g_keep_children_lists) ? 1 : 0; list < 2; list++) {      /* a dead creditor's children inherit its loans: those lists stay */
                int32_t *head = list == 0 ? a->children_head : a->mates_head;
                int e = head[s];
                while (e >= 0) { int nx = a->kin_next[e]; a->kin_next[e] = (int32_t)free_head; free_head = e; e = nx; }
                head[s] = -1;
            }
        }
        cn[SSB_C_KIN_FREE] = free_head + 1;
    }
    if (p->rng_mode == SSB_RNG_SCALABLE) {
        /* newborn survivors: re-ordered and numbered by cell index */
        int64_t first = w;
        while (first > 0 && a->born[a->order[first - 1]] == t && a->id[a->order[first - 1]] < 0) first--;
        int64_t k = w - first;
        if (k > 0) {
            int64_t *keys = (int64_t *)malloc(sizeof(int64_t) * (size_t)k);
            for (int64_t i = 0; i < k; i++) keys[i] = ((int64_t)a->pos[a->order[first + i]] << 32) | (uint32_t)a->order[first + i];
            qsort(keys, (size_t)k, sizeof(int64_t), cmp_i64);
            for (int64_t i = 0; i < k; i++) {
                int s = (int)(keys[i] & 0xffffffff);
                a->order[first + i] = s;
                a->id[s] = (int32_t)cn[SSB_C_NEXT_ID]++;
            }
            free(keys);
        }
    }
}

/* replaceDeadAgents -> configureAgents (sugarscape.py:855-859,171-267) with the mode S sources of
 * randomness: the empty cells of a quadrant are ordered by a hash of the cell, the quadrants by a
 * hash of their position, a new agent's tribe tags come from the substream of its own id
 * (tests/golden/make_golden_scalable.py applies the same substitutions to the reference). */
#define SALT_CELL 0x52455043454C4Cull
#define SALT_QUAD 0x5245505155414Dull
enum { SITE_NEWTAGS = 4 };

static int quadrant_of_cell(const ssb_params_t *p, int x, int y) {
    /* position (0-based) of the cell's quadrant among the ACTIVE quadrants, or -1 */
    int W = p->width, H = p->height, qw = p->quad_w, qh = p->quad_h, q = 0;
    if (x < qw && y < qh) q = 1;
    else if (x >= W - qw && y < qh) q = 2;
    else if (x >= W - qw && y >= H - qh) q = 3;
    else if (x < qw && y >= H - qh) q = 4;
    if (q == 0 || !((p->quad_mask >> (q - 1)) & 1)) return -1;
    int pos = 0;
    for (int k = 1; k < q; k++) pos += (p->quad_mask >> (k - 1)) & 1;
    return pos;
}

void orc_replace(const ssb_params_t *p, const ssb_grid_t *g, const ssb_agents_t *a, const ssb_endowments_t *en, int32_t t) {
    ctx_t c;
    orc_rng_t rng;
    memset(&rng, 0, sizeof(rng));
    make_ctx(&c, p, g, a, &rng, t);
    int64_t *cn = a->counters;
    cn[SSB_C_N_REPLACED] = 0;
    int64_t need = (int64_t)p->agent_replacements - cn[SSB_C_N_LIST];
    if (need <= 0 || en->count <= 0) return;
    int W = p->width, H = p->height, Q = 0;
    for (int k = 0; k < 4; k++) Q += (p->quad_mask >> k) & 1;
    if (Q == 0) return;
    uint64_t key = orc_step_key(p->seed, t);
    uint64_t *cells[4]; int64_t n_empty[4] = {0, 0, 0, 0}, total = 0;
    for (int q = 0; q < Q; q++) cells[q] = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)(p->quad_w > 0 && p->quad_h > 0 ? (int64_t)p->quad_w * p->quad_h : 1));
    for (int x = 0; x < W; x++)
        for (int y = 0; y < H; y++) {
            int q = quadrant_of_cell(p, x, y), cell = x * H + y;
            if (q < 0 || g->occ[cell] >= 0) continue;
            cells[q][n_empty[q]++] = ((mix64(key ^ SALT_CELL ^ (uint64_t)cell) >> 32) << 32) | (uint32_t)cell;
            total++;
        }
    if (total > 0) {
        if (need > total) need = total;
        for (int q = 0; q < Q; q++) qsort(cells[q], (size_t)n_empty[q], sizeof(uint64_t), cmp_u64);
        /* quadrant order: ascending (hash, position) */
        int qorder[4]; uint64_t qk[4];
        for (int q = 0; q < Q; q++) { qorder[q] = q; qk[q] = mix64(key ^ SALT_QUAD ^ (uint64_t)q); }
        for (int i = 1; i < Q; i++)
            for (int j = i; j > 0 && (qk[qorder[j - 1]] > qk[qorder[j]] || (qk[qorder[j - 1]] == qk[qorder[j]] && qorder[j - 1] > qorder[j])); j--) {
                int tq = qorder[j]; qorder[j] = qorder[j - 1]; qorder[j - 1] = tq;
            }
        /* slots: recycled ones first, but not those of this step's dead (the stats still read them) */
        int64_t n_dead = cn[SSB_C_N_DEAD], older = cn[SSB_C_N_FREE] - n_dead, n_free0 = cn[SSB_C_N_FREE];
        rng.mode = SSB_RNG_SCALABLE; rng.step_key = key;
        int64_t made = 0;
        for (int64_t i = 0; i < need; i++) {
            int q = qorder[i % Q];
            if (n_empty[q] == 0) { cn[SSB_C_OVERFLOW] = 4; break; }   /* the reference would raise IndexError here */
            int cell = (int)(cells[q][--n_empty[q]] & 0xffffffffu);
            int slot;
            if (older > 0) slot = a->free_slots[--older];
            else { if (cn[SSB_C_N_SLOTS] >= a->capacity) { cn[SSB_C_OVERFLOW] = 1; break; } slot = (int)cn[SSB_C_N_SLOTS]++; }
            int64_t ei = (cn[SSB_C_ENDOW_INDEX] + i) % en->count;
            int id = (int)(cn[SSB_C_NEXT_ID] + i);
            a->id[slot] = id; a->pos[slot] = cell; a->born[slot] = t; a->cod[slot] = SSB_ALIVE;
            a->sugar[slot] = a->start_sugar[slot] = en->sugar[ei];
            a->spice[slot] = a->start_spice[slot] = en->spice[ei];
            a->sugar_metab[slot] = en->sugar_metab[ei]; a->spice_metab[slot] = en->spice_metab[ei];
            a->vision[slot] = en->vision[ei]; a->movement[slot] = en->movement[ei];
            a->age[slot] = 0; a->max_age[slot] = en->max_age[ei]; a->sex[slot] = en->sex[ei];
            a->fert_age[slot] = en->fert_age[ei]; a->infert_age[slot] = en->infert_age[ei];
            a->fert_factor[slot] = en->fert_factor[ei]; a->aggression[slot] = en->aggression[ei];
            a->lookahead[slot] = en->lookahead[ei]; a->trade_factor[slot] = en->trade_factor[ei];
            a->mrs[slot] = 1;
            uint32_t tags = en->tags[ei];
            if (p->tribe_per_quadrant && p->tag_len > 0 && p->max_tribes > 0) {
                /* generateTribeTags(tribe = quadrant position), sugarscape.py:548-559 */
                int L = p->tag_len;
                double tribe_size = (double)(L + 1) / (double)p->max_tribes;
                int min_z = (int)floor(q * tribe_size), max_z = (int)floor((q + 1) * tribe_size) - 1;
                if (max_z > L) max_z = L;
                rng.agent_id = (uint32_t)id; rng_at(&rng, SITE_NEWTAGS);
                int zeroes = min_z + (int)rng_below(&rng, (uint32_t)(max_z - min_z + 1));
                int32_t bits[32];
                for (int k = 0; k < L; k++) bits[k] = k < zeroes ? 0 : 1;
                rng_shuffle_i32(&rng, bits, L);
                tags = 0;
                for (int k = 0; k < L; k++) tags |= (uint32_t)bits[k] << k;
            }
            a->tags[slot] = tags;
            a->tribe[slot] = find_tribe(&c, tags);
            a->last_moved[slot] = -1; a->last_sugar[slot] = 0; a->last_spice[slot] = 0;
            a->n_neighborhood[slot] = 0; a->n_valid_moves[slot] = 0;
            a->happiness[slot] = 0; a->wealth_happiness[slot] = 0; a->family_happiness[slot] = 0;
            a->trade_volume[slot] = 0; a->trade_price[slot] = 0; a->last_repro[slot] = -1;
            a->children_head[slot] = -1; a->mates_head[slot] = -1; a->father[slot] = -1; a->mother[slot] = -1;
            g->occ[cell] = slot;
            a->order[cn[SSB_C_N_LIST] + i] = slot;
            made++;
        }
        /* close the gap in the free stack left by the recycled slots */
        memmove(a->free_slots + older, a->free_slots + (n_free0 - n_dead), sizeof(int32_t) * (size_t)n_dead);
        cn[SSB_C_N_FREE] = older + n_dead;
        cn[SSB_C_N_LIST] += made; cn[SSB_C_NEXT_ID] += made; cn[SSB_C_ENDOW_INDEX] += made; cn[SSB_C_N_REPLACED] = made;
    }
    for (int q = 0; q < Q; q++) free(cells[q]);
}

static int cmp_f64(const void *x, const void *y) {
    double a = *(const double *)x, b = *(const double *)y;
    return a < b ? -1 : a > b;
}

/* raw reductions of updateRuntimeStatsPerGroup (sugarscape.py:1005-1426) */
void orc_stats(const ssb_params_t *p, const ssb_grid_t *g, const ssb_agents_t *a, int32_t t, ssb_stats_t *out) {
    memset(out, 0, sizeof(*out));
    ctx_t c;
    make_ctx(&c, p, g, a, NULL, t);
    int64_t *cn = a->counters, n = cn[SSB_C_N_LIST];
    out->timestep = t;
    out->population = n;
    out->min_wealth = 0; out->max_wealth = 0;
    for (int i = 0; i < 32; i++) out->tribe_first[i] = -1;
    double *wealths = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    int64_t seen = 0;                 /* agents of the current statistics pass (g_pass: all / group / control) */
    for (int64_t i = 0; i < n; i++) {
        int s = a->order[i];
        if (!in_pass(s)) continue;
        double w = a->sugar[s] + a->spice[s];
        wealths[seen] = w;
        out->sum_sugar_metab += a->sugar_metab[s]; out->sum_spice_metab += a->spice_metab[s];
        out->sum_movement += a->movement[s]; out->sum_vision += a->vision[s]; out->sum_age += a->age[s];
        out->n_acted += a->age[s] > 0;
        out->sum_wealth += w;
        if (seen == 0 || w > out->max_wealth) out->max_wealth = w;
        if (seen == 0 || w < out->min_wealth) out->min_wealth = w;
        seen++;
        out->sum_wealth_collected += w - (a->last_sugar[s] + a->last_spice[s]);
        /* findTimeToLive (agent.py:1113-1140) */
        out->sum_burn_rate += time_to_live(a, s, 0);
        out->sum_ttl_age_limited += time_to_live(a, s, 1);
        out->sum_neighbors += a->n_neighborhood[s];
        out->sum_valid_moves += a->n_valid_moves[s];
        if (a->trade_volume[s] > 0) {
            out->n_traders++; out->trade_volume += a->trade_volume[s]; out->sum_trade_price += a->trade_price[s];
        }
        out->sum_happiness += a->happiness[s];
        out->sum_wealth_happiness += a->wealth_happiness[s];
        out->sum_family_happiness += a->family_happiness[s];
        int tr = a->tribe[s];
        if (tr >= 0 && tr < 32) { if (out->tribe_count[tr]++ == 0) out->tribe_first[tr] = i; }
        if (a->born[s] == t && t > 0) { /* counted below from N_BORN */ }
    }
    n = seen;
    out->population = n;
    /* Lorenz area numerator (sugarscape.py:913-934): sum of cumulative wealth over the first
     * n-1 sorted agents plus half the total */
    qsort(wealths, (size_t)n, sizeof(double), cmp_f64);
    double cum = 0, area = 0;
    for (int64_t i = 0; i + 1 < n; i++) { cum += wealths[i]; area += cum; }
    if (n > 0) { cum += wealths[n - 1]; area += cum / 2; }
    out->lorenz_weighted_sum = area;
    /* exact mode: updateGiniCoefficient's own arithmetic (sugarscape.py:913-934), operation by operation */
    out->gini_area = -1.0;
    if (p->rng_mode == SSB_RNG_EXACT && n > 0) {
        double total = 0;
        for (int64_t i = 0; i < n; i++) total += wealths[i];          /* sum(agentWealths), sorted order */
        double lorenz = 0;
        if (total != 0) {
            double c = 0;
            for (int64_t i = 0; i + 1 < n; i++) { c += wealths[i]; lorenz += c / total; }
            c += wealths[n - 1];
            lorenz += (c / 2) / total;
            lorenz /= (double)n;
        }
        out->gini_area = lorenz;                                       /* total == 0: area 0 -> coefficient 1 */
    }
    free(wealths);
    /* this step's dead (sugarscape.py:1213-1265) */
    out->n_dead = 0;
    for (int64_t i = 0; i < cn[SSB_C_N_DEAD]; i++) {
        int s = a->dead_list[i];
        if (!in_pass(s)) continue;
        out->n_dead++;
        if (a->last_moved[s] == t) out->n_dead_moved++;
        out->sum_age_at_death += a->age[s];
        out->dead_wealth_collected += (a->sugar[s] + a->spice[s]) - (a->last_sugar[s] + a->last_spice[s]);
        out->dead_starvation += a->cod[s] == SSB_STARVATION;
        out->dead_aging += a->cod[s] == SSB_AGING;
        out->dead_combat += a->cod[s] == SSB_COMBAT;
    }
    out->n_born = cn[SSB_C_N_BORN];
    out->n_replaced = cn[SSB_C_N_REPLACED];
    if (g_pass != 0) {                /* bornAgents of the pass (sugarscape.py:1359-1362): this step's newborn, alive or not */
        out->n_born = 0;
        for (int64_t i = 0; i < cn[SSB_C_N_LIST]; i++) { int s = a->order[i]; if (a->born[s] == t && t > 0 && in_pass(s)) out->n_born++; }
        for (int64_t i = 0; i < cn[SSB_C_N_DEAD]; i++) { int s = a->dead_list[i]; if (a->born[s] == t && t > 0 && in_pass(s)) out->n_born++; }
    }
    /* environment totals (sugarscape.py:1044-1051) with the closed-form LastProduced */
    int W = p->width, H = p->height;
    int seasons = p->season_interval > 0;
    int flipped = seasons && t > 0 ? (((t - 1) / p->season_interval) & 1) : 0;
    int dry_grows = seasons && p->seasonal_growback_delay > 0 ? (t % p->seasonal_growback_delay == 0) : 0;
    for (int x = 0; x < W; x++)
        for (int y = 0; y < H; y++) {
            int i = x * H + y;
            out->env_wealth_total += g->sugar[i] + g->spice[i];
            if (t <= 1) out->env_capacity_total += g->sugar_cap[i] + g->spice_cap[i];   /* only logged at t == 1 */
        }
    (void)flipped; (void)dry_grows;
}
