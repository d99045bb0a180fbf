/* ssb.h -- C ABI of the B200 Sugarscape timestep engine (libssb.so).
 *
 * The reference (nkremerh/sugarscape) is pure Python and has NO plugin / FFI interface
 * (SURVEY.md section 8b), so this ABI is new.  The seam it replaces is
 *
 *     Sugarscape.doTimestep()                      reference sugarscape.py:389-421
 *       environment.doTimestep(t)                  reference environment.py:105-109
 *       random.shuffle(self.agents)                reference sugarscape.py:400
 *       for agent in self.agents: agent.doTimestep reference sugarscape.py:404-407, agent.py:568-598
 *       removeDeadAgents()                         reference sugarscape.py:843-853
 *       updateRuntimeStats() (the reductions)      reference sugarscape.py:1005-1426
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no torch / C++ types.
 *   - Every call returns 0 on success, <0 on error; ssb_last_error() gives the text.
 *   - The library never allocates or frees SIMULATION buffers.  The host owns them (PyTorch
 *     CUDA tensors in the shipped host code) and passes raw device pointers in ssb_grid_t /
 *     ssb_agents_t.  The engine owns only scratch (RNG state, marks, work lists, partials).
 *   - Grid arrays are x-major like the reference's grid[x][y] (environment.py:4,44):
 *     idx = x * height + y.
 *   - Agent arrays are slot-indexed struct-of-arrays with `capacity` slots.  `order[0..n_list)`
 *     holds the reference's agent LIST (sugarscape.py:61) as slot numbers, in list order.
 *   - All calls on one engine must come from one host thread; work is enqueued on the CUDA
 *     stream passed in (`stream` is a cudaStream_t cast to void*); only ssb_stats,
 *     ssb_get_* and ssb_sync block.
 *
 * The same structs (with HOST pointers) drive the CPU oracle in oracle/ (test infrastructure).
 */
#ifndef SSB_H
#define SSB_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* 3: ssb_stats_t.gini_area; SSB_F_NOWRAP / SSB_F_MOORE; the sweep scheduler's entry points (ssb_set_scheduler ...,
 * ssb_set_shard_sweep).  The optional per-slot state of the rule families keeps its name "ABI 2" below.
 * Added since without a layout change (same version): SSB_F_RADIAL, ssb_ethics_t.model 4 (ethics.Asimov), bit 1 of
 * ssb_ethics_t.top_ordering, SSB_GROUP_RACE / SSB_GROUP_MODEL, overflow code 15. */
/* 4: ssb_group_t.hood_list / hood_n (experimental groups whose membership changes during the run). */
#define SSB_ABI_VERSION 4
#define SSB_GROUP_HOOD_MAX  161 /* entries per slot of a stored neighbourhood list: the candidate cells of mode E + the agent itself */

/* RNG modes */
#define SSB_RNG_EXACT    0 /* mode E: CPython MT19937 word-for-word replay, ordered execution   */
#define SSB_RNG_SCALABLE 1 /* mode S: counter-based per-agent substreams, parallel execution    */

/* cause of death codes (agent.py:296-313 causeOfDeath strings) */
#define SSB_ALIVE       0
#define SSB_STARVATION  1
#define SSB_AGING       2
#define SSB_COMBAT      3
#define SSB_OTHER       4

/* sexes (agent.py:47) */
#define SSB_SEX_NONE   0
#define SSB_SEX_MALE   1
#define SSB_SEX_FEMALE 2

/* rule flags (params.flags): which optional rule families are live in this run */
#define SSB_F_SPICE     (1 << 0) /* any spice capacity / metabolism / holdings               */
#define SSB_F_POLLUTION (1 << 1) /* pollution timeframe configured                           */
#define SSB_F_REPRO     (1 << 2) /* some agent has fertilityFactor > 0                       */
#define SSB_F_TAGS      (1 << 3) /* tag strings present                                      */
#define SSB_F_TAGGING   (1 << 4) /* agentTagging                                             */
#define SSB_F_COMBAT    (1 << 5) /* some agent has aggression > 0                            */
#define SSB_F_TRADE     (1 << 6) /* some agent has tradeFactor != 0                          */
#define SSB_F_TAGPREF   (1 << 7) /* agentTagPreferences                                      */
#define SSB_F_MOORE     (1 << 9) /* neighborhoodMode = "moore": cell.neighbors has eight entries (cell.py:77-89; exact mode) */
#define SSB_F_NOWRAP    (1 << 8) /* environmentWraparound = false: no neighbours across the map's edge (cell.py:47-121; exact mode) */
#define SSB_F_RADIAL    (1 << 10) /* agentVisionMode = agentMovementMode = "radial": discs instead of crosses (environment.py:146-173; exact mode) */

typedef struct ssb_params {
    int32_t abi_version;
    int32_t width, height;          /* environmentWidth / environmentHeight                    */
    int32_t rng_mode;               /* SSB_RNG_*                                               */
    uint64_t seed;                  /* configuration["seed"]                                   */
    int32_t flags;                  /* SSB_F_*                                                 */
    int32_t max_agent_range;        /* max(maxVision, maxMovement), environment.py:135-137     */
    /* growback + seasons (environment.py:61-96,199-211) */
    int32_t sugar_regrow, spice_regrow;
    int32_t season_interval, seasonal_growback_delay, equator;
    /* pollution (environment.py:97-103,192-197; cell.py:27-45) */
    int32_t pollution_start, pollution_end;
    int32_t diffusion_start, diffusion_end, diffusion_delay;
    double sugar_prod_pollution, spice_prod_pollution;
    double sugar_cons_pollution, spice_cons_pollution;
    /* agents */
    double max_combat_loot;         /* environmentMaxCombatLoot                                */
    int32_t tag_len, max_tribes;    /* agentTagStringLength, environmentMaxTribes              */
    int32_t max_timestep;           /* configuration["timesteps"]                              */
    int32_t id_bits;                /* mode S: agent IDs must stay below 1 << id_bits (<= 26)   */
    int32_t max_sugar_capacity, max_spice_capacity; /* environmentMaxSugar / environmentMaxSpice: upper bound of any cell */
    int32_t inheritance_policy;     /* agentInheritancePolicy: 0 none, 1 children, 2 sons, 3 daughters (agent.py:373-429) */
    /* replacement in mode S (sugarscape.py:171-267,475-497,855-859) */
    int32_t agent_replacements;     /* agentReplacements: population is topped up to this value  */
    int32_t quad_w, quad_h;         /* quadrant size: floor(W/2 * factor), floor(H/2 * factor)    */
    int32_t quad_mask;              /* bit q-1 set if quadrant q (1..4) is a starting quadrant    */
    int32_t tribe_per_quadrant;     /* environmentTribePerQuadrant                                */
    int32_t _pad;
} ssb_params_t;

typedef struct ssb_grid {
    int64_t n_cells;                /* width * height                                          */
    int32_t *sugar, *sugar_cap;     /* cell.sugar / cell.maxSugar  (cell.py:8,19)              */
    int32_t *spice, *spice_cap;     /* cell.spice / cell.maxSpice                              */
    double  *pollution;             /* cell.pollution (cell.py:14)                             */
    double  *pollution_flux;        /* scratch for the Jacobi pass (cell.py:104-109)           */
    int32_t *occ;                   /* cell.agent as a slot number, -1 when empty              */
} ssb_grid_t;

/* counters[] layout (int64 each, device memory owned by the host) */
#define SSB_C_N_LIST   0  /* len(sugarscape.agents)                                           */
#define SSB_C_N_SLOTS  1  /* high-water mark of used slots                                    */
#define SSB_C_NEXT_ID  2  /* sugarscape.nextAgentID                                           */
#define SSB_C_N_FREE   3  /* entries in free_slots[]                                          */
#define SSB_C_N_BORN   4  /* births during the current step                                   */
#define SSB_C_N_DEAD   5  /* agents removed at the end of the current step                    */
#define SSB_C_ROUNDS   6  /* mode S: commit rounds used by the last agent step                */
#define SSB_C_OVERFLOW 7  /* != 0 if a capacity was exceeded (births dropped) -> host raises  */
#define SSB_C_N_KIN    8  /* used entries of the kin pool                                     */
#define SSB_C_ENDOW_INDEX 9 /* sugarscape.agentEndowmentIndex                                  */
#define SSB_C_N_REPLACED 10 /* agents created by the last replacement                           */
#define SSB_C_N_MIGRATED 11 /* sharded: agents this rank handed to other ranks in the last step */
#define SSB_C_KIN_FREE 12  /* kin pool: first entry of the free chain + 1 (0: empty); entries of removed owners are recycled */
#define SSB_C_COUNT    16

typedef struct ssb_agents {
    int64_t capacity;               /* slots allocated in every array below                    */
    int64_t *counters;              /* SSB_C_COUNT int64 values                                */
    int32_t *order;                 /* agent list as slots, list order                         */
    int32_t *free_slots;            /* recycled slots                                          */
    int32_t *dead_list;             /* slots removed at the end of the current step            */
    /* identity and position */
    int32_t *id;                    /* agent.ID                                                */
    int32_t *pos;                   /* cell index x*H+y, -1 once dead                          */
    int32_t *born;                  /* agent.born                                              */
    int32_t *cod;                   /* SSB_ALIVE or cause of death                             */
    /* endowments (agent.py:14-62) */
    double  *sugar, *spice;
    int32_t *sugar_metab, *spice_metab;
    int32_t *vision, *movement;
    int32_t *age, *max_age;
    int32_t *sex;
    int32_t *fert_age, *infert_age;
    double  *fert_factor;
    double  *start_sugar, *start_spice;
    double  *aggression;
    double  *lookahead;
    double  *trade_factor;
    double  *mrs;                   /* marginalRateOfSubstitution (agent.py:102)               */
    uint32_t *tags;                 /* bit i = tags[i]                                         */
    int32_t *tribe;
    /* per-transaction by-products that feed the log (agent.py:734-745,1555-1561) */
    int32_t *last_moved;            /* lastMovedTimestep                                       */
    double  *last_sugar, *last_spice;
    int32_t *n_neighborhood;        /* len(movementNeighborhood)                               */
    int32_t *n_valid_moves;         /* len(validMoves)                                         */
    double  *happiness;             /* agent.happiness (sum of the five components)            */
    double  *wealth_happiness;
    double  *family_happiness;
    int32_t *trade_volume;          /* agent.tradeVolume of its last transaction               */
    double  *trade_price;           /* max(spicePrice, sugarPrice) of its last transaction     */
    int32_t *last_repro;            /* lastReproducedTimestep                                  */
    /* kin lists (mode E): socialNetwork["children"] / ["mates"] of the INITIATING parent
     * (agent.py:521-527), as singly linked lists in the kin pool below */
    int32_t *children_head, *mates_head;         /* per slot: first pool entry, -1 = empty      */
    int32_t *father, *mother;       /* parent IDs (-1 none)                                    */
    /* kin pool: entry i = (slot, id) of the relative and the next entry of the same list */
    int64_t kin_capacity;
    int32_t *kin_slot, *kin_id, *kin_next;
} ssb_agents_t;

/* The endowment list that replacements reuse cyclically (sugarscape.agentEndowments,
 * sugarscape.py:192-194,207-208), struct-of-arrays in creation order.  Mode S only: in mode E the
 * host keeps the list and places replacements itself. */
typedef struct ssb_endowments {
    int64_t count;
    double  *sugar, *spice;
    int32_t *sugar_metab, *spice_metab, *vision, *movement, *max_age, *sex, *fert_age, *infert_age;
    double  *fert_factor, *aggression, *lookahead, *trade_factor;
    uint32_t *tags;                 /* endowment tags (used unless tribes are per quadrant)     */
} ssb_endowments_t;

/* Raw reductions for the reference's runtimeStats (sugarscape.py:1005-1426).  The host turns
 * them into the 59 log keys (means, Python round(), JSON). */
typedef struct ssb_stats {
    int64_t timestep;
    int64_t population;
    int64_t sum_sugar_metab, sum_spice_metab, sum_movement, sum_vision, sum_age;
    double  sum_wealth, max_wealth, min_wealth;
    double  sum_wealth_collected;       /* live agents: wealth - (lastSugar + lastSpice)        */
    double  sum_burn_rate;              /* sum findTimeToLive()                                 */
    double  sum_ttl_age_limited;        /* sum findTimeToLive(True)                             */
    int64_t sum_neighbors, sum_valid_moves;
    int64_t n_traders, trade_volume;
    double  sum_trade_price;
    double  sum_happiness, sum_wealth_happiness, sum_family_happiness;
    int64_t tribe_count[32];            /* agents per tribe (tribes dict, 1171-1174)            */
    int64_t tribe_first[32];            /* list index of the first agent of each tribe          */
    /* dead agents removed this step (1213-1265) */
    int64_t n_dead, n_dead_moved, dead_starvation, dead_aging, dead_combat, sum_age_at_death;
    double  dead_wealth_collected;
    int64_t n_born, n_replaced;
    int64_t n_acted;                    /* live agents with age > 0: healthHappiness == 1 (agent.py:1017-1021) */
    /* environment (1044-1051) */
    int64_t env_wealth_total;           /* sum sugar + spice                                    */
    int64_t env_wealth_created;         /* sum sugarLastProduced + spiceLastProduced            */
    int64_t env_capacity_total;         /* sum maxSugar + maxSpice (added once at t == 1)       */
    double  lorenz_weighted_sum;        /* sum_i (cumulative wealth after i sorted agents), trapezoid form */
    int64_t rounds;                     /* mode S commit rounds of this step                    */
    double  gini_area;                  /* exact mode: the Lorenz curve area of updateGiniCoefficient (sugarscape.py:913-934)
                                           accumulated in the reference's own order -- total = sum of the sorted wealths, one
                                           division per element, / numAgents -- so that round(..., 3) can never differ;
                                           scalable mode: -1 (the host uses lorenz_weighted_sum / sum_wealth / population)  */
} ssb_stats_t;

typedef struct ssb_engine ssb_engine_t;

/* ---- ABI 2: decision models (reference ethics.py:105-244, the Bentham family: bentham / altruist /
 * egoist / negativeBentham and their lookahead variants; spawn rules sugarscape.py:219-233; SURVEY.md
 * section 8 row a21).  Optional per-slot state, exact RNG mode; without it every agent is the plain
 * rawSugarscape Agent.  The host owns the arrays (capacity = ssb_agents_t.capacity slots). */
typedef struct ssb_ethics {
    int64_t capacity;
    int32_t *model;                 /* 0 = Agent (greedy findBestCell), 1 = ethics.Bentham, 2 = ethics.Temperance
                                       (simple variant: ethics.py:412-434, 464-465, 493-506; needs ssb_agentlog_t bound),
                                       3 = ethics.Temperance with pecs=True (ethics.py:436-491, 508-547),
                                       4 = ethics.Asimov (ethics.py:8-98; among plain agents only)                       */
    int32_t *top_ordering;          /* bit 0: a "Top" Bentham variant -- the cell with the best ethical value wins (ethics.py:124-129);
                                       bit 1: the agent's decisionModel string is the experimental group (SSB_GROUP_MODEL) */
    double  *selfishness;           /* selfishnessFactor after the spawn rules (-1 .. 1)              */
    double  *decision_factor;       /* decisionModelFactor: the ethical ranking runs when > 0         */
    double  *lookahead_factor;      /* decisionModelLookaheadFactor                                   */
    double  *lookahead_discount;    /* decisionModelLookaheadDiscount                                 */
    double  *dynamic_decision_factor; /* dynamicDecisionModelFactor: how a Temperance agent's factor moves with its virtue rolls */
    /* Temperance PECS state (zero for a new agent): rules = agentConsumedAdequate / Ample, communityDisapproval,
     * agentOverconsumed, communityDisdain; counts = timeSeenOverconsuming, timesSeenIndulging, timesOverharvested */
    int32_t *pecs_rules;            /* [slot][5]                                                      */
    int32_t *pecs_counts;           /* [slot][3]                                                      */
    double  *social_pressure, *last_delta_ttl, *dynamic_social_pressure, *total_metabolism;
    double  *dynamic_selfishness;   /* dynamicSelfishnessFactor (the "Dynamic" variants; 0 = fixed selfishness; needs
                                       ssb_agentlog_t bound: the update compares agent.timeToLive values)   */
    /* by-products of the agent's last ranking that feed the log (agent.py:734-745) */
    int32_t *last_move_optimal;     /* lastMoveOptimal                                                */
    int32_t *move_rank;             /* index of the chosen cell in the welfare-sorted list            */
    double  *move_diff;             /* welfare of the top cell minus welfare of the chosen one        */
    double  global_max_wealth;      /* environment.globalMaxSugar + globalMaxSpice                    */
    /* ABI 4, Asimov robots next to the Bentham family: a robot's second law asks a Bentham neighbour for ITS value of a cell
     * (ethics.py:74-75), and that value walks agent.neighborhood as the neighbour last stored it -- at its own ranking, at
     * birth, or when the agents were (re)configured (agent.py:1052-1065, 526; sugarscape.py:265-267).  Kept per slot as
     * (slot, id) pairs in list order when hood_stride > 0; hood_n = -1: to be rebuilt at the start of the next agent step
     * (the host sets it for every slot when it places agents; -2 for the initial population, whose lists the reference
     * makes before the first infections: ranges without the disease modifiers). */
    int32_t *hood_list, *hood_id;   /* [slot][hood_stride]                                            */
    int32_t *hood_n;
    int64_t hood_stride;            /* 0: not kept; else SSB_GROUP_HOOD_MAX                           */
} ssb_ethics_t;
/* raw sums of sugarscape.py:1141-1160, 1218-1222 over the agent list (+ this step's dead for the moves) */
typedef struct ssb_ethics_stats {
    double  sum_selfishness;
    int64_t optimal_moves, moves;   /* agentLastMoveOptimalityPercentage = optimal / moves            */
    int64_t sum_move_rank;
    double  sum_move_diff;
} ssb_ethics_stats_t;
/* ---- ABI 2: the optional rule families of the reference beyond the book examples, exact RNG mode.  Every
 * family is a set of per-slot arrays (capacity = ssb_agents_t.capacity) owned by the host; a family is off
 * when its first array pointer is NULL.  Python lists of the reference (loans, friends) keep their list
 * semantics: order of insertion, remove = first equal element. ---- */

/* lending: agent.py:431-483 doLending, 190-212 addLoan*, 920-930, 1226-1245, 1285-1296, 1330-1377 payDebt*,
 * 1532-1541 updateLoans, 1549-1553 mean income.  A loan is one record in the debtor's "creditors" list and
 * one in the creditor's "debtors" list; lists are chains through the loan pool in list order. */
#define SSB_LC_POOL_USED   0  /* high-water mark of the loan pool                                         */
#define SSB_LC_FREE_HEAD   1  /* recycled pool entries: index of the first one + 1 (0 = none)             */
#define SSB_LC_N_DEAD      2  /* dead creditors recorded so far (the table is a ring: entry i lives at i % dead_capacity) */
#define SSB_LC_OVERFLOW    3  /* != 0: a pool was exhausted -> host raises                                 */
#define SSB_LC_COUNT       8
typedef struct ssb_lending {
    double  *lending_factor, *base_interest_rate;
    int32_t *loan_duration;
    double  *sugar_mean_income, *spice_mean_income;   /* updateMeanIncome, start at 1                      */
    int32_t *last_lended, *last_loans;                /* timestep of the last lending / loans made then    */
    int32_t *creditors_head, *debtors_head;           /* per slot: first pool entry of the list, -1 empty  */
    int64_t pool_capacity;
    int32_t *loan_other_slot, *loan_other_id;         /* the other party, named by (slot, id)              */
    double  *loan_sugar, *loan_spice;
    int32_t *loan_duration_v, *loan_origin, *loan_next;
    /* a dead creditor's children list outlives its slot (payDebtToCreditorChildren, agent.py:1362-1377).  The
     * table is a ring: an entry matters only while one of the creditor's loans is outstanding, i.e. for at most
     * the longest loan duration after the death (its loans fall due at origin + duration, and only a living
     * creditor's loans are ever renewed); overwriting a younger entry raises the overflow flag. */
    int64_t dead_capacity;
    int64_t max_loan_duration;                        /* upper end of agentLoanDuration: how long a dead creditor can matter */
    int32_t *dead_slot, *dead_id, *dead_children_head, *dead_time;
    int64_t *counters;                                /* SSB_LC_COUNT values                               */
} ssb_lending_t;

/* friends + conflict happiness: agent.py:1499-1520 updateFriends, 1099-1105, 912-918 */
#define SSB_MAX_FRIENDS 16
typedef struct ssb_social {
    int32_t *max_friends, *last_combat;
    double  *social_happiness, *conflict_happiness;   /* as of the agent's last updateHappiness            */
    int32_t *n_friends;
    int32_t *friend_slot, *friend_id, *friend_hd;     /* [slot][SSB_MAX_FRIENDS], list order               */
} ssb_social_t;

/* diseases: condition.py:36-86; agent.py:184-188 addDisease, 227-247 catchDisease, 315-371 doDisease /
 * doInfectionAttempt, 1034-1050 nearest Hamming distance, and the find* accessors with their modifiers
 * (716-717, 1031-1032, 1107-1111, 1161-1162).  An agent carries a disease at most once: its list of
 * infections is a [slot][max_sick] table of records in list order. */
#define SSB_DM_AGGRESSION   0
#define SSB_DM_FERTILITY    1
#define SSB_DM_FRIENDLINESS 2
#define SSB_DM_HAPPINESS    3
#define SSB_DM_MOVEMENT     4
#define SSB_DM_SPICE_METAB  5
#define SSB_DM_SUGAR_METAB  6
#define SSB_DM_VISION       7
#define SSB_DM_COUNT        8
typedef struct ssb_disease_def {
    int32_t id, incubation, start_timestep, tag_len;
    uint32_t tags, _pad;                              /* bit j = diseaseTags[j]                             */
    double  transmission;
    double  penalty[SSB_DM_COUNT];
    int64_t infected;                                 /* len(disease.infected)                              */
    int64_t listed;                                   /* how often sugarscape.diseases holds it             */
} ssb_disease_def_t;
typedef struct ssb_diseases {
    uint64_t *immune, *start_immune;                  /* bit i = immuneSystem[i]                            */
    double  *protection;                              /* diseaseProtectionChance                            */
    double  *modifier;                                /* [slot][SSB_DM_COUNT]: sum of the triggered penalties */
    int32_t *disease_death, *last_spread, *n_spread;
    double  *health_happiness;                        /* as of the agent's last updateHappiness             */
    int32_t *last_valid_moves;                        /* lastValidMoves (1 when the range is empty)         */
    int32_t *stat_mark;                               /* scratch of the reductions (distinct infectors)     */
    int32_t *n_sick;
    int32_t *sick_disease, *sick_start, *sick_end, *sick_caught, *sick_incubation, *sick_infector_slot, *sick_infector_id;
    ssb_disease_def_t *defs;                          /* [n_diseases]                                       */
    const uint64_t *immune_bits;                      /* per timestep: draws for a child's mismatching immune positions */
    int32_t max_sick, immune_len, n_diseases, n_immune_bits;
} ssb_diseases_t;

/* universal income (agent.py:708-714, 1127-1134), racial tags (1082-1086; inheritance 879-887) and depression
 * (condition.py:13-34; agent.py:137-141, 889-895) */
typedef struct ssb_misc {
    double  *universal_sugar, *universal_spice;
    int32_t *last_income_sugar, *last_income_spice;
    uint32_t *racial_tags;                            /* 2 bits per position                                */
    int32_t *race;                                    /* -1 = None                                          */
    int32_t *depressed;
    double  *happiness_unit;                          /* 1, 0.5763 when depressed                           */
    double  *health_happiness;                        /* as of the agent's last updateHappiness (used when diseases are off) */
    const uint32_t *racial_picks;                     /* per timestep: the parent each racial tag comes from */
    const double *depression_draw;                    /* per timestep: the draw a child's depression is decided with */
    double  depression_percentage;
    int32_t income_interval_sugar, income_interval_spice, racial_len, n_races, n_tables, _pad;
} ssb_misc_t;

/* experimental group (agent.py:1256-1284 isInGroup; sugarscape.py:998-1003, 1175-1198, 1230-1250): the log repeats
 * every statistic for the group and for its complement ("control") and counts the interactions between them.
 * Supported groups (ssb_group_t.kind): "depressed" (ssb_misc_t.depressed), "female" / "male" (ssb_agents_t.sex), "race<N>"
 * (ssb_misc_t.race == N, N in bits 8.. of kind), a decision model's name (agent.decisionModel == group: bit 1 of
 * ssb_ethics_t.top_ordering, set by the host and handed on to the children with the parent's class) --
 * memberships that are fixed at birth -- and "sick" (ssb_diseases_t.n_sick > 0), whose membership changes during the run:
 * the reference re-evaluates the
 * membership of every agent of agent.movementNeighborhood at statistics time (sugarscape.py:1145-1151), so each agent's
 * movement neighbourhood is kept as a list of slots (hood_list, written at its ranking) and split again at the end of the
 * agent step, before the dead leave their slots.  ("disease<N>" has statistics rules of its own, sugarscape.py:1203-1272,
 * never has a member in the reference (a string compared with integer ids) and is not built; "raceInGroup" and, at the first death, "ageRange<N>" raise in the reference's isInGroup.) */
#define SSB_GROUP_DEPRESSED 1
#define SSB_GROUP_FEMALE    2
#define SSB_GROUP_MALE      3
#define SSB_GROUP_RACE      4   /* kind = SSB_GROUP_RACE | (race id << 8)                                       */
#define SSB_GROUP_MODEL     5   /* needs ssb_ethics_t                                                           */
#define SSB_GROUP_SICK      6   /* from here on: membership changes during the run (hood_list is written)       */
#define SSB_GI_COMBAT 0
#define SSB_GI_DISEASE 1
#define SSB_GI_LENDING 2
#define SSB_GI_REPRODUCTION 3
#define SSB_GI_TRADE 4
#define SSB_GI_KINDS 5
typedef struct ssb_group {
    int32_t *with_control, *with_experimental;        /* [slot][SSB_GI_KINDS]: interactions since the last statistics pass */
    int32_t *control_neighbors, *experimental_neighbors;   /* movementNeighborhood split, as of the agent's last ranking */
    int64_t kind;                                     /* SSB_GROUP_*                                        */
    int32_t *hood_list;                               /* [slot][SSB_GROUP_HOOD_MAX] slots of agent.movementNeighborhood (kinds >= SSB_GROUP_SICK) */
    int32_t *hood_n;                                  /* its length                                         */
} ssb_group_t;

/* per-agent log (configuration["agentLogfile"]; agent.py:1567-1620 updateRuntimeStats, sugarscape.py:415-421): one
 * record per agent that completed its transaction, in action order.  Raw values; the host rounds and formats. */
#define SSB_AL_TIMESTEP 0
#define SSB_AL_ID 1
#define SSB_AL_AGE 2
#define SSB_AL_SUGAR 3
#define SSB_AL_SPICE 4
#define SSB_AL_SUGAR_GAINED 5
#define SSB_AL_SPICE_GAINED 6
#define SSB_AL_MOVEMENT 7
#define SSB_AL_TIME_TO_LIVE 8
#define SSB_AL_DEPRESSION 9
#define SSB_AL_HAPPINESS 10
#define SSB_AL_PREY_KILLED 11
#define SSB_AL_PREY_WEALTH 12
#define SSB_AL_TRADE_PARTNERS 13
#define SSB_AL_DISEASES_SPREAD 14
#define SSB_AL_MATES 15
#define SSB_AL_NEIGHBORS 16
#define SSB_AL_VALID_MOVES 17
#define SSB_AL_MOVE_RANK 18
#define SSB_AL_LENDING_PARTNERS 19
#define SSB_AL_POLLUTION_DIFF 20
#define SSB_AL_TTL_DIFF 21
#define SSB_AL_NEIGHBORS_IN_TRIBE 22
#define SSB_AL_SAME_RACE_NEIGHBORS 23
#define SSB_AL_EXPERIMENTAL_NEIGHBORS 24
#define SSB_AL_CONTROL_NEIGHBORS 25
#define SSB_AL_FIELDS 26
typedef struct ssb_agentlog {
    double  *time_to_live;                            /* agent.timeToLive: what the last findTimeToLive() call left  */
    double  *last_pollution;                          /* pollution of the cell the agent left at its last move       */
    double  *prey_wealth;                             /* lastPreyWealth                                              */
    int32_t *last_combat;                             /* lastCombatTimestep                                          */
    int32_t *last_mates;                              /* lastMates                                                   */
    int32_t *valid_moves, *move_rank;                 /* lastValidMoves, lastMoveRank                                */
    int32_t *n_neighbors, *neighbor_slot;             /* self.neighbors as of updateNeighbors: [slot][4]             */
    double  *records;                                 /* [record_capacity][SSB_AL_FIELDS]                            */
    int64_t record_capacity;
    int64_t *n_records;                               /* records of the current step (the host resets it)            */
    int64_t write_records;                            /* 0: only agent.timeToLive is tracked (dynamic selfishness)   */
} ssb_agentlog_t;

typedef struct ssb_ext {
    int64_t capacity;
    ssb_lending_t lending;
    ssb_social_t social;
    ssb_diseases_t diseases;
    ssb_misc_t misc;
    ssb_group_t group;
    ssb_agentlog_t agentlog;
} ssb_ext_t;
/* raw sums for the log keys of the families (sugarscape.py:1130-1160) */
typedef struct ssb_ext_stats {
    int64_t loan_volume;                              /* sum of lastLoans over the agents that lent this step */
    double  sum_social_happiness, sum_conflict_happiness;
    /* diseases (sugarscape.py:1162, 1200-1211, 1228, 1254-1277, 1316-1317) */
    int64_t sick_agents, disease_incidence, distinct_infectors, disease_prevalence, disease_deaths;
    double  sum_health_happiness;
    int64_t move_space;                               /* sum of lastValidMoves                              */
    double  sum_burn_rate, sum_ttl_age_limited;       /* findTimeToLive with the disease modifiers, list order */
    int64_t race_count[4], race_first[4];             /* agents per race, list index of the first one (-1 none) */
} ssb_ext_stats_t;
/* the statistics of one pass over the experimental group (pass 1) or the control group (pass 2) */
typedef struct ssb_group_stats {
    ssb_stats_t stats;                                /* without the Lorenz sum and the environment totals */
    ssb_ethics_stats_t ethics;
    ssb_ext_stats_t ext;
    int64_t to_control[SSB_GI_KINDS], to_experimental[SSB_GI_KINDS];
    int64_t control_neighbors, experimental_neighbors;
} ssb_group_stats_t;
/* out[0] = everybody (only the two neighbour sums are filled), out[1] = experimental group, out[2] = control group,
 * of the last ssb_reduce_stats; synchronises */
int  ssb_group_stats(ssb_engine_t *e, ssb_group_stats_t out[3]);
int  ssb_sizeof_group_stats(void);
/* call after ssb_bind, before the first step (mode E only) */
int  ssb_bind_ext(ssb_engine_t *e, const ssb_ext_t *ext);
int  ssb_ext_stats(ssb_engine_t *e, ssb_ext_stats_t *host_out);   /* of the last ssb_reduce_stats; synchronises */
int  ssb_sizeof_ext(void);

/* call after ssb_bind, before the first step (mode E only; mode S and sharded engines refuse) */
int  ssb_bind_ethics(ssb_engine_t *e, const ssb_ethics_t *ethics);
/* reductions of the last ssb_reduce_stats; synchronises */
int  ssb_ethics_stats(ssb_engine_t *e, ssb_ethics_stats_t *host_out);
int  ssb_sizeof_ethics(void);

/* ---- sharding one simulation over the GPUs of a box (x-slabs) -------------
 * A "stitched" array is world consecutive chunks, chunk p backed by rank p's HBM (one process per
 * GPU), mapped at consecutive addresses on every rank.  ssb_fabric_alloc() creates this rank's chunk
 * and exports it as a POSIX file descriptor; the host passes the descriptors between the ranks and
 * ssb_fabric_map() imports the peers' chunks and maps the array.  The host binds stitched arrays as
 * ssb_grid_t / ssb_agents_t, so a kernel reaches the neighbouring slab's cells and agents over NVLink
 * with ordinary loads, stores and atomics.  All ranks issue the same sequence of alloc calls. */
#define SSB_MAX_RANKS 8
typedef struct ssb_fabric ssb_fabric_t;
int  ssb_fabric_create(int32_t rank, int32_t world, int32_t device, ssb_fabric_t **out);
int  ssb_fabric_alloc(ssb_fabric_t *f, uint64_t chunk_bytes, int32_t *fd_out);    /* returns the array index (>= 0) */
/* fds[world], chunk_bytes[world] (multiples of the granularity; chunk p follows chunk p - 1); closes the peers' fds */
int  ssb_fabric_map(ssb_fabric_t *f, int32_t index, const int32_t *fds, const uint64_t *chunk_bytes, void **base_out);
/* synchronous copies to / from stitched memory: kind 0 = host -> device, 1 = device -> host,
 * 2 = zero `bytes` at dst (src ignored).  Callers split ranges at chunk boundaries. */
int  ssb_fabric_copy(ssb_fabric_t *f, void *dst, const void *src, uint64_t bytes, int32_t kind);
uint64_t ssb_fabric_granularity(ssb_fabric_t *f);
int  ssb_fabric_destroy(ssb_fabric_t *f);
const char *ssb_fabric_last_error(ssb_fabric_t *f);

typedef struct ssb_shard_t {
    int32_t rank, world;
    int64_t cell_end[SSB_MAX_RANKS];/* rank r owns the cells [cell_end[r-1], cell_end[r]) -- whole x-lines; the host
                                       balances the slabs by population                                   */
    int64_t slot_lo, slot_hi;       /* the agent slots this rank allocates from                      */
    int64_t mark_lo, mark_hi;       /* the mark words this rank clears (the ranks' ranges partition the array) */
    void   *ctrl;                   /* stitched control pages, zeroed: chunk p = rank p's page        */
    uint64_t ctrl_stride;           /* bytes between pages (>= 4096)                                  */
    void   *marks;                  /* stitched reservation marks: u32 per (coarsened) cell, see ssb_set_mark_shift */
    void   *gather;                 /* stitched scratch for candidate / wealth-key all-gathers        */
    uint64_t gather_stride;         /* bytes per rank chunk: >= 8 * (4 + 1) * slots per rank          */
} ssb_shard_t;
/* Scalable mode, rule sets without births only (sharded births are experimental and refused
 * unless SSB_EXPERIMENTAL_SHARDED_BIRTHS is set).  Call after ssb_bind, before the
 * first step; grid and agent arrays bound must be stitched.  Per-rank state: order[], free_slots[],
 * dead_list[], counters[] (N_LIST / N_SLOTS / N_FREE are per rank; NEXT_ID / ENDOW_INDEX global). */
int  ssb_set_shard(ssb_engine_t *e, const ssb_shard_t *shard);
/* The sweep scheduler on a sharded engine (after ssb_set_shard; collective: all ranks or none).  Possible for the rule sets
 * of the plain lean transaction (no combat / spice / pollution / tagging / trade / births) when every slab consists of
 * whole 64-column tiles; ssb_shard_sweep_possible tells.  desc: u64 per cell, done: u64 x width x ceil(height / 64),
 * pcell: 8 bytes per cell of the 4 x 4-tiled cell array (width and height rounded up to 4) -- stitched like the grid
 * (rank r backs the columns of its slab) and zero-filled.  Without it a sharded engine runs the reservation rounds. */
int  ssb_shard_sweep_possible(const ssb_engine_t *e);
int  ssb_set_shard_sweep(ssb_engine_t *e, void *desc, void *done, void *pcell);
/* last phase of a sharded step (ssb_step calls it): records of agents that walked onto another
 * rank's slab move to that rank (slot of its range, its list); collective, no-op unless sharded */
int  ssb_migrate(ssb_engine_t *e, int32_t t, void *stream);
/* box-wide barrier on `stream` (orders the peers' later kernels after this rank's earlier ones) */
int  ssb_shard_barrier(ssb_engine_t *e, void *stream);
/* 1 if a box-wide barrier timed out (a peer is gone); the engine is unusable afterwards */
int  ssb_shard_broken(ssb_engine_t *e);

/* lifecycle */
int  ssb_create(const ssb_params_t *params, int device, ssb_engine_t **out);
int  ssb_bind(ssb_engine_t *e, const ssb_grid_t *grid, const ssb_agents_t *agents);
void ssb_destroy(ssb_engine_t *e);
const char *ssb_last_error(const ssb_engine_t *e);
int  ssb_abi_version(void);

/* mode E: the CPython MT19937 main stream; maps 1:1 to random.getstate()[1] */
int  ssb_set_mt_state(ssb_engine_t *e, const uint32_t state[624], int32_t index);
int  ssb_get_mt_state(ssb_engine_t *e, uint32_t state[624], int32_t *index);

/* Per-timestep child-endowment tables (host-precomputed; reference agent.py:827-875 draws them
 * from private streams that are a pure function of the timestep -- SURVEY.md Appendix A.9).
 * endow_bits[t] bit k = inherit endowment k from the mate (1) or the acting parent (0);
 * tag_bits[t] bit k = k-th draw for tag positions where the parents differ. */
int  ssb_set_step_tables(ssb_engine_t *e, const uint32_t *endow_bits, const uint32_t *tag_bits, int32_t n);

/* Mode S scheduler.  Both reproduce "agents act one at a time in ascending priority" (the reference's
 * random.shuffle + agent loop, sugarscape.py:400-407) bit for bit:
 *   SSB_SCHED_SWEEP (default on one GPU): the conflict DAG of the step is built once from the agents' start
 *     positions (csrc/ssb_flow_core.cuh) as per-agent predecessor masks; a persistent grid walks the agents in
 *     priority-bucket order and runs each one as soon as its predecessors' done bits are set
 *     (csrc/ssb_sweep_core.cuh).  Rule sets without tagging / trading / reproduction run the lean transaction (one
 *     thread per agent on packed 32-byte records); SSB_SCHED_SWEEP_GENERIC forces the generic one;
 *   SSB_SCHED_DATAFLOW: the same DAG executed through ready queues and dependency counters (csrc/ssb_dataflow.cuh);
 *   SSB_SCHED_ROUNDS: deterministic reservation rounds over a sliding priority prefix (round 1 kernel; sharded
 *     engines use it).  Environment override at bind time: SSB_SCHEDULER=rounds|dataflow|sweep|sweep-generic. */
#define SSB_SCHED_ROUNDS   0
#define SSB_SCHED_DATAFLOW 1
#define SSB_SCHED_SWEEP    3
#define SSB_SCHED_SWEEP_GENERIC 4
int  ssb_set_scheduler(ssb_engine_t *e, int32_t scheduler);
int  ssb_get_scheduler(const ssb_engine_t *e);
/* 1 if the agent step runs the lean transaction (csrc/ssb_sweep_core.cuh: one thread per agent on packed 32-byte
 * records; rule sets without tagging / trading / reproduction / tag preferences under SSB_SCHED_SWEEP) */
int  ssb_uses_lean_transaction(const ssb_engine_t *e);
/* sub-phases of the last timed agent step under the sweep scheduler, milliseconds:
 * {prepare, build + clear, cell pack, sweep kernel, cell + record unpack}; -1 if none was timed */
int  ssb_agent_phases(ssb_engine_t *e, float out[5]);
/* diagnostics of the sweep kernel since bind (SSB_SWEEP_DIAG=1 at bind time; -1 otherwise):
 * {warp iterations, entries held at a readiness check, entries run, entries held back by the watermark} */
int  ssb_sweep_diag(ssb_engine_t *e, unsigned long long out[4]);

/* mode S tuning (SSB_SCHED_ROUNDS): number of priority epochs of the sliding-prefix commit (power of two, default 128) */
int  ssb_set_epochs(ssb_engine_t *e, int32_t epochs);
/* mode S tuning: reservation marks on a grid coarsened by 2^shift cells per side (default 1 when the map allows it) */
int  ssb_set_mark_shift(ssb_engine_t *e, int32_t shift);
/* mode S tuning: keep the mark array resident in L2 (persisting access window on `stream`) */
int  ssb_set_l2_persist(ssb_engine_t *e, void *stream, int32_t on);

/* test hook: out[i] = x[i] ** y[i] with the engine's bit-exact restatement of glibc's pow
 * (device pointers; exact[i] = 0 where the argument fell outside the restated path) */
int  ssb_test_pow(const double *x, const double *y, double *out, int32_t *exact, int32_t n);

/* debug: per-round trace of the last mode S agent step.  enable_rounds > 0 allocates the trace;
 * out (host, may be NULL) receives per round {active, ready, ns at start, after mark, after
 * select, after execute} */
int  ssb_round_trace(ssb_engine_t *e, int32_t enable_rounds, uint64_t *out, int32_t out_rounds);

/* sizeof() of the ABI structs, for binding self-checks */
int  ssb_sizeof_params(void);
int  ssb_sizeof_grid(void);
int  ssb_sizeof_agents(void);
int  ssb_sizeof_stats(void);

/* one timestep, in the reference's phase order (sugarscape.py:395-410) */
int  ssb_env_step(ssb_engine_t *e, int32_t t, void *stream);     /* environment.doTimestep     */
int  ssb_agent_step(ssb_engine_t *e, int32_t t, double prev_mean_wealth, void *stream);
                                                                  /* shuffle + agent loop       */
int  ssb_compact(ssb_engine_t *e, int32_t t, void *stream);      /* removeDeadAgents + births  */
int  ssb_bind_endowments(ssb_engine_t *e, const ssb_endowments_t *endowments);
int  ssb_replace(ssb_engine_t *e, int32_t t, void *stream);      /* replaceDeadAgents, mode S  */
int  ssb_reduce_stats(ssb_engine_t *e, int32_t t, void *stream); /* device reductions          */
int  ssb_step(ssb_engine_t *e, int32_t t, double prev_mean_wealth, void *stream);
                                                                  /* the four above, in order   */
int  ssb_stats(ssb_engine_t *e, ssb_stats_t *host_out);          /* synchronises               */
#define SSB_STATS_RING 1024
int  ssb_stats_history(ssb_engine_t *e, int32_t t, ssb_stats_t *host_out); /* stats of a recent step t (device ring of SSB_STATS_RING entries); synchronises */
int  ssb_sync(ssb_engine_t *e, void *stream);
/* Diffusion steps write the new field into the other pollution buffer and swap roles: returns 1
 * if the field currently lives in the buffer bound as grid->pollution_flux, else 0. */
int  ssb_pollution_buffer(const ssb_engine_t *e);

/* instrumentation: number of kernels this engine has launched, device time of the last
 * ssb_env_step / ssb_agent_step in ms (CUDA events on the caller's stream; ssb_timing syncs) */
int64_t ssb_launch_count(const ssb_engine_t *e);
int  ssb_timing(ssb_engine_t *e, float *env_ms, float *agent_ms, float *compact_ms, float *stats_ms);
int  ssb_enable_timing(ssb_engine_t *e, int32_t on);

#ifdef __cplusplus
}
#endif
#endif /* SSB_H */
