// ssb_ext.cuh -- the reference's rule families beyond the book examples (ABI 2, include/ssb.h: ssb_ext_t),
// mode E.  Included by ssb_rules.cuh after DevCtx and the basic predicates.  Everything here runs on lane 0
// of the single ordered warp: the families walk Python lists (loans, friends) whose semantics are
// sequential -- order of insertion, remove = first equal element, index loops over lists that change
// underneath -- and they only exist at example scale.
#pragma once

namespace ssb {

// child endowment bits of the families (sugarscape_b200/steptables.py ENDOWMENT_BITS)
enum { EB_BASE_INTEREST = 12, EB_LENDING_FACTOR = 13, EB_LOAN_DURATION = 14, EB_MAX_FRIENDS = 15 };

enum { EB_PROTECTION = 16, EB_UNIVERSAL_SPICE = 17, EB_UNIVERSAL_SUGAR = 18 };

__device__ __forceinline__ bool has_lending(const DevCtx &c) { return c.x != nullptr && c.x->lending.lending_factor != nullptr; }
__device__ __forceinline__ bool has_social(const DevCtx &c) { return c.x != nullptr && c.x->social.max_friends != nullptr; }
__device__ __forceinline__ bool has_disease(const DevCtx &c) { return c.x != nullptr && c.x->diseases.immune != nullptr; }
__device__ __forceinline__ bool has_misc(const DevCtx &c) { return c.x != nullptr && c.x->misc.universal_sugar != nullptr; }
__device__ __forceinline__ bool has_agentlog(const DevCtx &c) { return c.x != nullptr && c.x->agentlog.time_to_live != nullptr; }
__device__ __forceinline__ bool has_group(const DevCtx &c) { return c.x != nullptr && c.x->group.with_control != nullptr; }
// Agent.isInGroup (agent.py:1256-1284) for the configured experimental group
__device__ __forceinline__ bool is_member(const DevCtx &c, int s) {
    if (!has_group(c)) return false;
    const long long kind = c.x->group.kind;
    switch ((int)(kind & 0xff)) {
        case SSB_GROUP_DEPRESSED: return has_misc(c) && c.x->misc.depressed[s] != 0;
        case SSB_GROUP_FEMALE: return c.a.sex[s] == SSB_SEX_FEMALE;
        case SSB_GROUP_MALE: return c.a.sex[s] == SSB_SEX_MALE;
        case SSB_GROUP_RACE: return has_misc(c) && c.x->misc.race[s] == (int)(kind >> 8);      // self.race == raceID (agent.py:1273-1275)
        case SSB_GROUP_MODEL: return c.eth != nullptr && (c.eth->top_ordering[s] & 2) != 0;    // group == self.decisionModel (agent.py:1262)
        case SSB_GROUP_SICK: return has_disease(c) && c.x->diseases.n_sick[s] > 0;             // isSick (a dead agent's list is empty, agent.py:309-313)
    }
    return false;
}
// membership can change during the run: agent.movementNeighborhood is kept as a list and split at the end of the step
__device__ __forceinline__ bool group_dynamic(const DevCtx &c) { return has_group(c) && (c.x->group.kind & 0xff) >= SSB_GROUP_SICK; }
// statistics pass: 0 = everybody, 1 = the experimental group, 2 = the control group
__device__ __forceinline__ bool in_pass(const DevCtx &c, int s, int pass) { return pass == 0 || (pass == 1) == is_member(c, s); }
// <kind>With{Experimental,Control}Group += 1
__device__ __forceinline__ void group_note(const DevCtx &c, int s, int kind, int partner) {
    if (!has_group(c)) return;
    if (is_member(c, partner)) c.x->group.with_experimental[(long long)s * SSB_GI_KINDS + kind]++;
    else c.x->group.with_control[(long long)s * SSB_GI_KINDS + kind]++;
}
__device__ __forceinline__ void group_reset_slot(const DevCtx &c, int s) {
    for (int k = 0; k < SSB_GI_KINDS; k++) { c.x->group.with_control[(long long)s * SSB_GI_KINDS + k] = 0; c.x->group.with_experimental[(long long)s * SSB_GI_KINDS + k] = 0; }
    c.x->group.control_neighbors[s] = 0; c.x->group.experimental_neighbors[s] = 0;
    if (group_dynamic(c)) c.x->group.hood_n[s] = 0;
}
// sugarscape.py:1145-1151 for the groups whose membership moves: every agent's stored movement neighbourhood split by the
// membership its members have at the END of the agent step (the dead are still in their slots and, like the reference's
// dead agent objects, no longer sick).  All lanes of the group; called once after the last agent of the step.
template <int G>
__device__ inline void group_resplit_neighbors(const DevCtx &c) {
    const ssb_group_t &Gp = c.x->group;
    const long long n_list = (long long)((volatile unsigned long long *)c.a.counters)[SSB_C_N_LIST];      // (lane 0 appended the newborn)
    for (long long i = threadIdx.x & (G - 1); i < n_list; i += G) {
        const int s = c.a.order[i];
        int exp_n = 0, ctrl_n = 0;
        for (int k = 0; k < Gp.hood_n[s]; k++) { if (is_member(c, Gp.hood_list[(long long)s * SSB_GROUP_HOOD_MAX + k])) exp_n++; else ctrl_n++; }
        Gp.control_neighbors[s] = ctrl_n; Gp.experimental_neighbors[s] = exp_n;
    }
}
// happinessUnit: 1, 0.5763 when depressed
__device__ __forceinline__ double happiness_unit(const DevCtx &c, int s) { return has_misc(c) ? c.x->misc.happiness_unit[s] : 1.0; }

// the find* accessors (agent.py:716-717, 1031-1032, 1107-1111, 1161-1162): endowment + disease modifier, floored at 0
__device__ __forceinline__ double dis_mod(const DevCtx &c, int s, int which) {
    return has_disease(c) ? c.x->diseases.modifier[(long long)s * SSB_DM_COUNT + which] : 0.0;
}
__device__ __forceinline__ double pos_part(double v) { return v > 0 ? v : 0; }
__device__ __forceinline__ double eff_sugar_metab(const DevCtx &c, int s) { return pos_part(c.a.sugar_metab[s] + dis_mod(c, s, SSB_DM_SUGAR_METAB)); }
__device__ __forceinline__ double eff_spice_metab(const DevCtx &c, int s) { return pos_part(c.a.spice_metab[s] + dis_mod(c, s, SSB_DM_SPICE_METAB)); }
__device__ __forceinline__ double eff_vision(const DevCtx &c, int s) { return pos_part(c.a.vision[s] + dis_mod(c, s, SSB_DM_VISION)); }
__device__ __forceinline__ double eff_movement(const DevCtx &c, int s) { return pos_part(c.a.movement[s] + dis_mod(c, s, SSB_DM_MOVEMENT)); }
__device__ __forceinline__ double eff_aggression(const DevCtx &c, int s) { return pos_part(c.a.aggression[s] + dis_mod(c, s, SSB_DM_AGGRESSION)); }
__device__ __forceinline__ bool is_sick(const DevCtx &c, int s) { return has_disease(c) && c.x->diseases.n_sick[s] > 0; }
// Agent.isFertile with the fertility modifier (agent.py:1244-1247, 1007-1008)
__device__ __forceinline__ bool is_fertile_x(const DevCtx &c, int s) {
    return c.a.sugar[s] >= c.a.start_sugar[s] && c.a.spice[s] >= c.a.start_spice[s] &&
           c.a.age[s] >= c.a.fert_age[s] && c.a.age[s] < c.a.infert_age[s] && (c.a.fert_factor[s] + dis_mod(c, s, SSB_DM_FERTILITY)) > 0;
}
// Agent.findTimeToLive (agent.py:1113-1140) with the modifiers
__device__ __forceinline__ double time_to_live_x(const DevCtx &c, int s, bool age_limited) {
    const double sm = eff_sugar_metab(c, s), pm = eff_spice_metab(c, s);
    const double big = 9223372036854775807.0;            // sys.maxsize
    double st = sm > 0 ? c.a.sugar[s] / sm : big;
    double pt = pm > 0 ? c.a.spice[s] / pm : big;
    if (has_misc(c)) {                                   // the universal income term (agent.py:1127-1134)
        const ssb_misc_t &M = c.x->misc;
        if (M.universal_sugar[s] != 0) {
            const double income = (st * M.universal_sugar[s]) / M.income_interval_sugar;
            st = sm > 0 ? (c.a.sugar[s] + income) / sm : big;
        }
        if (M.universal_spice[s] != 0) {
            const double income = (pt * M.universal_spice[s]) / M.income_interval_spice;
            pt = pm > 0 ? (c.a.spice[s] + income) / pm : big;
        }
    }
    double ttl = st < pt ? st : pt;
    if (age_limited) { const double lim = (double)(c.a.max_age[s] - c.a.age[s]); if (lim < ttl) ttl = lim; }
    // findTimeToLive leaves its result in agent.timeToLive (agent.py:1137-1139), whoever called it: the agent's
    // record in the per-agent log reports the difference to the previous call
    if (has_agentlog(c)) c.x->agentlog.time_to_live[s] = ttl;
    return ttl;
}

// Agent.findTimeToLive(potentialCell=cell) (agent.py:1113-1140): the holdings plus what the cell and its occupant
// would yield; the universal income term uses the current holdings; agent.timeToLive is left alone
// (my_sugar / my_spice: the agent's holdings as they stand -- the move part keeps them in registers)
__device__ inline double time_to_live_potential_of(const DevCtx &c, int s, int cell, double my_sugar, double my_spice) {
    const double sm = eff_sugar_metab(c, s), pm = eff_spice_metab(c, s);
    const double big = 9223372036854775807.0;
    double sugar = my_sugar, spice = my_spice;
    const int o = c.g.occ[cell];
    if (o >= 0) {          // (the agent's own cell counts as occupied: by itself)
        sugar += fmin(c.p.max_combat_loot, o == s ? my_sugar : c.a.sugar[o]);
        spice += fmin(c.p.max_combat_loot, o == s ? my_spice : c.a.spice[o]);
    }
    sugar += c.g.sugar[cell];
    spice += (c.p.flags & SSB_F_SPICE) ? c.g.spice[cell] : 0;
    double st = sm > 0 ? sugar / sm : big, pt = pm > 0 ? spice / pm : big;
    if (has_misc(c)) {
        const ssb_misc_t &M = c.x->misc;
        if (M.universal_sugar[s] != 0) { const double income = (st * M.universal_sugar[s]) / M.income_interval_sugar; st = sm > 0 ? (my_sugar + income) / sm : big; }
        if (M.universal_spice[s] != 0) { const double income = (pt * M.universal_spice[s]) / M.income_interval_spice; pt = pm > 0 ? (my_spice + income) / pm : big; }
    }
    return st < pt ? st : pt;
}
__device__ inline double time_to_live_potential(const DevCtx &c, int s, int cell) {
    return time_to_live_potential_of(c, s, cell, c.a.sugar[s], c.a.spice[s]);
}

// ---------------------------------------------------------------------------------------------
// universal income, racial tags, depression
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int race_of(uint32_t tags, int len) {          // findRace: most common tag, the smallest among ties
    int cnt[4] = {0, 0, 0, 0};
    for (int i = 0; i < len; i++) cnt[(tags >> (2 * i)) & 3u]++;
    int best = 0;
    for (int r = 1; r < 4; r++) if (cnt[r] > cnt[best]) best = r;
    return best;
}
// v *= ceil(v * f): the reference squares the endowment, so a depressed lineage with agentMovement 1000 leaves the
// i32 range within two generations (Python ints do not).  Saturated at 2^30: every rule compares these values with
// ranges / holdings far below, so the trajectory is the reference's; only the logged means differ once saturated.
__device__ __forceinline__ int32_t depress_scale(int32_t v, double f) {
    const double p = (double)v * ceil(v * f);
    return p >= 1073741824.0 ? 1073741824 : (int32_t)p;
}
// Depression.trigger (condition.py:27-33), at the creation of a depressed agent (agent.py:137-141)
__device__ inline void depress(const DevCtx &c, int s) {
    const ssb_agents_t &a = c.a;
    c.x->misc.depressed[s] = 1;
    a.aggression[s] *= 1.145;
    if (has_social(c)) c.x->social.max_friends[s] = (int32_t)ceil(c.x->social.max_friends[s] * 0.6333);
    c.x->misc.happiness_unit[s] *= 0.5763;
    a.movement[s] = depress_scale(a.movement[s], 0.429);
    a.spice_metab[s] = depress_scale(a.spice_metab[s], 1.544);
    a.sugar_metab[s] = depress_scale(a.sugar_metab[s], 1.544);
}
// the child's share of the family (agent.py:845-851, 879-895); called once its endowments are in place
__device__ inline void misc_child(const DevCtx &c, int ch, int s, int m, uint32_t eb) {
    const ssb_misc_t &M = c.x->misc;
    M.universal_spice[ch] = ((eb >> EB_UNIVERSAL_SPICE) & 1u) ? M.universal_spice[m] : M.universal_spice[s];
    M.universal_sugar[ch] = ((eb >> EB_UNIVERSAL_SUGAR) & 1u) ? M.universal_sugar[m] : M.universal_sugar[s];
    M.last_income_sugar[ch] = 0; M.last_income_spice[ch] = 0;
    M.depressed[ch] = 0; M.happiness_unit[ch] = 1; M.health_happiness[ch] = 0;
    M.race[ch] = -1; M.racial_tags[ch] = 0;
    if (M.racial_len > 0) {      // random.choice([mine, mate's]) per position (agent.py:885-886)
        const uint32_t picks = c.t < M.n_tables ? M.racial_picks[c.t] : 0u;
        uint32_t rt = 0;
        for (int i = 0; i < M.racial_len; i++) {
            const uint32_t src = ((picks >> i) & 1u) ? M.racial_tags[m] : M.racial_tags[s];
            rt |= ((src >> (2 * i)) & 3u) << (2 * i);
        }
        M.racial_tags[ch] = rt;
        M.race[ch] = race_of(rt, M.racial_len);
    }
}

// ---------------------------------------------------------------------------------------------
// diseases
// ---------------------------------------------------------------------------------------------
struct SickRec { int disease, start, end, caught, incubation, infector_slot, infector_id; };
__device__ __forceinline__ SickRec sick_load(const ssb_diseases_t &D, long long at) {
    SickRec r;
    r.disease = D.sick_disease[at]; r.start = D.sick_start[at]; r.end = D.sick_end[at]; r.caught = D.sick_caught[at];
    r.incubation = D.sick_incubation[at]; r.infector_slot = D.sick_infector_slot[at]; r.infector_id = D.sick_infector_id[at];
    return r;
}
__device__ __forceinline__ void sick_store(const ssb_diseases_t &D, long long at, const SickRec &r) {
    D.sick_disease[at] = r.disease; D.sick_start[at] = r.start; D.sick_end[at] = r.end; D.sick_caught[at] = r.caught;
    D.sick_incubation[at] = r.incubation; D.sick_infector_slot[at] = r.infector_slot; D.sick_infector_id[at] = r.infector_id;
}
__device__ inline void disease_reset_slot(const DevCtx &c, int s) {
    const ssb_diseases_t &D = c.x->diseases;
    D.n_sick[s] = 0;
    for (int k = 0; k < SSB_DM_COUNT; k++) D.modifier[(long long)s * SSB_DM_COUNT + k] = 0;
    D.disease_death[s] = 0; D.last_spread[s] = -1; D.n_spread[s] = 0; D.health_happiness[s] = 0; D.last_valid_moves[s] = 0;
}
__device__ inline void disease_trigger(const ssb_diseases_t &D, int s, int di) {      // condition.py:57-65
    for (int k = 0; k < SSB_DM_COUNT; k++) D.modifier[(long long)s * SSB_DM_COUNT + k] += D.defs[di].penalty[k];
}
__device__ inline void disease_recover(const ssb_diseases_t &D, int s, int di) {      // condition.py:67-76
    for (int k = 0; k < SSB_DM_COUNT; k++) D.modifier[(long long)s * SSB_DM_COUNT + k] -= D.defs[di].penalty[k];
    D.defs[di].infected--;
}
__device__ inline bool infected_with(const ssb_diseases_t &D, int s, int disease_id) {   // agent.py:1249-1254
    const long long b = (long long)s * D.max_sick;
    for (int i = 0; i < D.n_sick[s]; i++) if (D.defs[D.sick_disease[b + i]].id == disease_id) return true;
    return false;
}
// findNearestHammingDistanceInDisease (agent.py:1034-1050)
__device__ inline int disease_nearest(const ssb_diseases_t &D, int s, int di, int *start, int *end) {
    const int len = D.defs[di].tag_len;
    int best = len, b0 = 0, b1 = len - 1;
    for (int i = 0; i < D.immune_len - len; i++) {
        const uint32_t window = (uint32_t)((D.immune[s] >> i) & (len >= 32 ? 0xffffffffull : ((1ull << len) - 1ull)));
        const int hd = __popc(window ^ D.defs[di].tags);
        if (hd < best) { best = hd; b0 = i; b1 = i + (len - 1); }
    }
    *start = b0; *end = b1;
    return best;
}
// catchDisease (agent.py:227-247) through doInfectionAttempt (363-371): two random.uniform(0, 1)
template <class Rng>
__device__ __noinline__ void disease_catch(const DevCtx &c, int s, int di, int infector, Rng &rng) {
    const ssb_diseases_t &D = c.x->diseases;
    if (infected_with(D, s, D.defs[di].id)) return;
    SickRec rec = {di, -1, -1, c.t, D.defs[di].incubation, infector, c.a.id[infector]};
    if (disease_nearest(D, s, di, &rec.start, &rec.end) == 0) return;        // immune
    const double attack = 0.0 + (1.0 - 0.0) * rand_double(rng), defense = 0.0 + (1.0 - 0.0) * rand_double(rng);
    const bool attack_ok = attack <= D.defs[di].transmission && D.defs[di].transmission != 0.0;
    const bool defense_ok = defense <= D.protection[s] && D.protection[s] != 0.0;
    if (attack_ok && !defense_ok) {                       // addDisease (agent.py:184-188)
        if (D.n_sick[s] >= D.max_sick) return;            // cannot happen: one record per disease
        sick_store(D, (long long)s * D.max_sick + D.n_sick[s], rec);
        D.n_sick[s] += 1;
        D.defs[di].infected++;
        if (rec.incubation == 0) disease_trigger(D, s, di);
    }
}
// addRemainingDiseases (sugarscape.py:139-151; called right after the list shuffle of a step): a disease whose start has come
// goes to the first agent of the list that is not immune to it -- catchDisease(initial=True): no infection attempt, no
// infector, caught = this step -- and joins sugarscape.diseases; one that finds nobody waits for the next step.
// ssb_disease_def_t.start_timestep > 0 marks the waiting ones.  One lane.
__device__ __noinline__ void disease_add_remaining(const DevCtx &c) {
    const ssb_diseases_t &D = c.x->diseases;
    const long long n_list = c.a.counters[SSB_C_N_LIST];
    for (int di = 0; di < D.n_diseases; di++) {
        if (D.defs[di].start_timestep <= 0 || D.defs[di].start_timestep > c.t) continue;
        for (long long i = 0; i < n_list; i++) {
            const int s = c.a.order[i];
            SickRec rec = {di, -1, -1, c.t, D.defs[di].incubation, -1, -1};
            if (disease_nearest(D, s, di, &rec.start, &rec.end) == 0) continue;        // immune: the next agent
            if (!infected_with(D, s, D.defs[di].id) && D.n_sick[s] < D.max_sick) {
                sick_store(D, (long long)s * D.max_sick + D.n_sick[s], rec);
                D.n_sick[s] += 1;
                D.defs[di].infected++;
                if (rec.incubation == 0) disease_trigger(D, s, di);
            }
            D.defs[di].listed++;
            D.defs[di].start_timestep = 0;
            break;
        }
    }
}
// doDisease (agent.py:315-361)
template <class Rng>
__device__ __noinline__ void do_disease(const DevCtx &c, int s, Rng &rng) {
    const ssb_diseases_t &D = c.x->diseases;
    const long long b = (long long)s * D.max_sick;
    rng.at(SITE_DISEASE);
    for (int i = D.n_sick[s] - 1; i >= 1; i--) {          // random.shuffle(self.diseases)
        const uint32_t j = rand_below(rng, (uint32_t)i + 1u);
        const SickRec x = sick_load(D, b + i), y = sick_load(D, b + j);
        sick_store(D, b + i, y); sick_store(D, b + j, x);
    }
    for (int p = 0; p < D.n_sick[s]; p++) {               // Python for-loop over a list that may lose the current element
        SickRec rec = sick_load(D, b + p);
        const int di = rec.disease;
        if (rec.caught != c.t && rec.incubation > 0) { rec.incubation -= 1; D.sick_incubation[b + p] = rec.incubation; }
        if (rec.incubation == 0) disease_trigger(D, s, di);           // (every step: the penalties pile up)
        const int tag_len = D.defs[di].tag_len;
        if (tag_len > 0) {
            const uint32_t dtags = D.defs[di].tags;
            const int st = rec.start;
            int en = rec.end + 1;
            if (en > D.immune_len) en = D.immune_len;
            const int len = en - st;
            const uint64_t mask = len >= 64 ? ~0ull : ((1ull << len) - 1ull);
            const uint64_t response = (D.immune[s] >> st) & mask;     // the slice, taken before the flip
            for (int i = 0; i < len; i++)
                if (((response >> i) & 1ull) != ((dtags >> i) & 1u)) {
                    D.immune[s] = (D.immune[s] & ~(1ull << (st + i))) | ((uint64_t)((dtags >> i) & 1u) << (st + i));
                    break;
                }
            if (len == tag_len && response == (uint64_t)dtags) {      // diseaseTags == immuneResponse
                // list.remove(record): records are unique per disease
                const int n = D.n_sick[s];
                for (int q = p; q + 1 < n; q++) sick_store(D, b + q, sick_load(D, b + q + 1));
                D.n_sick[s] = n - 1;
                disease_recover(D, s, di);
            }
        }
    }
    const int count = D.n_sick[s];
    if (count == 0) return;
    int32_t nb4[SSB_MAX_NEIGHBOURS], neighbors[SSB_MAX_NEIGHBOURS];
    int nn = 0;
    const int n_nb = neighbours4(c, c.a.pos[s], nb4);
    for (int i = 0; i < n_nb; i++) { const int o = c.g.occ[nb4[i]]; if (o >= 0 && agent_alive(c, o)) neighbors[nn++] = o; }
    int spread = 0;
    shuffle(rng, neighbors, nn);
    for (int k = 0; k < nn; k++) {
        const int di = D.sick_disease[b + rand_below(rng, (uint32_t)count)];
        disease_catch(c, neighbors[k], di, s, rng);
        if (infected_with(D, neighbors[k], D.defs[di].id)) spread++;
        group_note(c, s, SSB_GI_DISEASE, neighbors[k]);
    }
    if (spread > 0) { D.last_spread[s] = c.t; D.n_spread[s] = spread; }
}
__device__ inline void disease_note_death(const DevCtx &c, int s) {        // doDeath (agent.py:302-313)
    const ssb_diseases_t &D = c.x->diseases;
    const long long b = (long long)s * D.max_sick;
    if (D.n_sick[s] > 0) D.disease_death[s] = 1;
    for (int i = 0; i < D.n_sick[s]; i++) disease_recover(D, s, D.sick_disease[b + i]);
    D.n_sick[s] = 0;
}
// a child's immune system: equal bits copied, mismatches drawn from the per-timestep stream (agent.py:897-908).
// startingImmuneSystem IS immuneSystem in the reference (agent.py:33,50 bind the same list): the parents pass on
// what their immune responses have learnt.
__device__ inline void disease_child(const DevCtx &c, int ch, int s, int m, uint32_t eb) {
    const ssb_diseases_t &D = c.x->diseases;
    disease_reset_slot(c, ch);
    const uint64_t mine = D.immune[s], other = D.immune[m];
    const uint64_t draws = (D.immune_bits && c.t < D.n_immune_bits) ? D.immune_bits[c.t] : 0;
    uint64_t child = 0;
    int draw = 0;
    for (int i = 0; i < D.immune_len; i++) {
        const uint64_t bs = (mine >> i) & 1ull, bm = (other >> i) & 1ull;
        const uint64_t bit = bs == bm ? bs : ((draws >> draw++) & 1ull);
        child |= bit << i;
    }
    D.immune[ch] = D.start_immune[ch] = child;
    D.protection[ch] = ((eb >> EB_PROTECTION) & 1u) ? D.protection[m] : D.protection[s];
}

// ---------------------------------------------------------------------------------------------
// friends + conflict happiness (agent.py:1499-1520 updateFriends; 1099-1105, 912-918)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void social_reset_slot(const DevCtx &c, int s) {
    const ssb_social_t &S = c.x->social;
    S.n_friends[s] = 0; S.last_combat[s] = -1; S.social_happiness[s] = 0; S.conflict_happiness[s] = 0;
}

__device__ inline void friend_remove_at(const ssb_social_t &S, int s, int q) {
    const long long b = (long long)s * SSB_MAX_FRIENDS;
    const int n = S.n_friends[s];
    for (int i = q; i + 1 < n; i++) {
        S.friend_slot[b + i] = S.friend_slot[b + i + 1]; S.friend_id[b + i] = S.friend_id[b + i + 1];
        S.friend_hd[b + i] = S.friend_hd[b + i + 1];
    }
    S.n_friends[s] = n - 1;
}
__device__ inline void friend_append(const ssb_social_t &S, int s, int slot, int id, int hd) {
    const long long b = (long long)s * SSB_MAX_FRIENDS;
    const int n = S.n_friends[s];
    if (n >= SSB_MAX_FRIENDS) return;        // the host refuses agentMaxFriends > SSB_MAX_FRIENDS
    S.friend_slot[b + n] = slot; S.friend_id[b + n] = id; S.friend_hd[b + n] = hd;
    S.n_friends[s] = n + 1;
}
// list.remove: the first entry with the same agent and distance
__device__ inline void friend_remove_first(const ssb_social_t &S, int s, int slot, int id, int hd) {
    const long long b = (long long)s * SSB_MAX_FRIENDS;
    for (int q = 0; q < S.n_friends[s]; q++)
        if (S.friend_slot[b + q] == slot && S.friend_id[b + q] == id && S.friend_hd[b + q] == hd) { friend_remove_at(S, s, q); return; }
}
// updateFriends (agent.py:1499-1520)
__device__ __noinline__ void update_friends(const DevCtx &c, int s, int nb) {
    const ssb_social_t &S = c.x->social;
    const long long b = (long long)s * SSB_MAX_FRIENDS;
    const int hd = (c.p.flags & SSB_F_TAGS) ? __popc(c.a.tags[s] ^ c.a.tags[nb]) : 0;
    const int id = c.a.id[nb];
    if (S.n_friends[s] < S.max_friends[s]) { friend_append(S, s, nb, id, hd); return; }
    int max_hd = 0, w_slot = 0, w_id = 0, w_hd = 0;
    bool have_max = false;
    for (int p = 0; p < S.n_friends[s]; p++) {
        const int f_slot = S.friend_slot[b + p], f_id = S.friend_id[b + p], f_hd = S.friend_hd[b + p];
        if (f_id == id) { friend_remove_first(S, s, f_slot, f_id, f_hd); friend_append(S, s, nb, id, hd); return; }
        if (f_hd > max_hd) { w_slot = f_slot; w_id = f_id; w_hd = f_hd; max_hd = f_hd; have_max = true; }
    }
    if (max_hd > hd && have_max) { friend_remove_first(S, s, w_slot, w_id, w_hd); friend_append(S, s, nb, id, hd); }
}
// updateNeighbors -> updateSocialNetwork -> updateFriends, neighbours in N, S, E, W order (agent.py:1563-1565, 1622-1632)
__device__ inline void update_neighbors(const DevCtx &c, int s, int pos) {
    int32_t nb4[SSB_MAX_NEIGHBOURS];
    const int n_nb = neighbours4(c, pos, nb4);
    for (int i = 0; i < n_nb; i++) { const int o = c.g.occ[nb4[i]]; if (o >= 0) update_friends(c, s, o); }
}

// ---------------------------------------------------------------------------------------------
// lending
// ---------------------------------------------------------------------------------------------
struct Loan { int other_slot, other_id; double sugar, spice; int duration, origin; };

__device__ __forceinline__ Loan loan_load(const ssb_lending_t &L, int e) {
    Loan l;
    l.other_slot = L.loan_other_slot[e]; l.other_id = L.loan_other_id[e]; l.sugar = L.loan_sugar[e]; l.spice = L.loan_spice[e];
    l.duration = L.loan_duration_v[e]; l.origin = L.loan_origin[e];
    return l;
}
__device__ __forceinline__ bool loan_equal(const Loan &x, const Loan &y) {
    return x.other_slot == y.other_slot && x.other_id == y.other_id && x.sugar == y.sugar && x.spice == y.spice &&
           x.duration == y.duration && x.origin == y.origin;
}
// pool entry p-th of a list, -1 past the end (Python's index loop over a list that changes underneath)
__device__ inline int loan_at(const ssb_lending_t &L, int head, int p) {
    int e = head;
    for (int i = 0; i < p && e >= 0; i++) e = L.loan_next[e];
    return e;
}
__device__ inline void loan_append(const ssb_lending_t &L, int32_t *head, int owner, const Loan &x) {
    int e;
    if (L.counters[SSB_LC_FREE_HEAD] > 0) { e = (int)L.counters[SSB_LC_FREE_HEAD] - 1; L.counters[SSB_LC_FREE_HEAD] = L.loan_next[e] + 1; }
    else {
        if (L.counters[SSB_LC_POOL_USED] >= L.pool_capacity) { L.counters[SSB_LC_OVERFLOW] = 1; return; }
        e = (int)L.counters[SSB_LC_POOL_USED]++;
    }
    L.loan_other_slot[e] = x.other_slot; L.loan_other_id[e] = x.other_id; L.loan_sugar[e] = x.sugar; L.loan_spice[e] = x.spice;
    L.loan_duration_v[e] = x.duration; L.loan_origin[e] = x.origin; L.loan_next[e] = -1;
    if (head[owner] < 0) { head[owner] = e; return; }
    int t = head[owner];
    while (L.loan_next[t] >= 0) t = L.loan_next[t];
    L.loan_next[t] = e;
}
__device__ __forceinline__ void loan_free(const ssb_lending_t &L, int e) {
    L.loan_next[e] = (int)L.counters[SSB_LC_FREE_HEAD] - 1;
    L.counters[SSB_LC_FREE_HEAD] = e + 1;
}
// list.remove(x): the first equal element
__device__ inline void loan_remove_first(const ssb_lending_t &L, int32_t *head, int owner, const Loan &x) {
    int prev = -1;
    for (int e = head[owner]; e >= 0; prev = e, e = L.loan_next[e]) {
        if (!loan_equal(loan_load(L, e), x)) continue;
        if (prev < 0) head[owner] = L.loan_next[e]; else L.loan_next[prev] = L.loan_next[e];
        loan_free(L, e);
        return;
    }
}
__device__ inline void loan_clear(const ssb_lending_t &L, int32_t *head, int owner) {
    for (int e = head[owner]; e >= 0;) { const int nx = L.loan_next[e]; loan_free(L, e); e = nx; }
    head[owner] = -1;
}
// a new agent object starts with empty lists
__device__ inline void lending_reset_slot(const DevCtx &c, int s) {
    const ssb_lending_t &L = c.x->lending;
    loan_clear(L, L.creditors_head, s); loan_clear(L, L.debtors_head, s);
    L.sugar_mean_income[s] = 1; L.spice_mean_income[s] = 1;
    L.last_lended[s] = -1; L.last_loans[s] = 0;
}
__device__ __forceinline__ bool loan_other_alive(const DevCtx &c, const Loan &l) {
    return c.a.id[l.other_slot] == l.other_id && agent_alive(c, l.other_slot);
}
// updateMeanIncome (agent.py:1549-1553), called by collectResourcesAtCell
__device__ __forceinline__ void lending_note_income(const DevCtx &c, int s, int cs, int cp) {
    const ssb_lending_t &L = c.x->lending;
    const double alpha = 0.05;
    L.sugar_mean_income[s] = (alpha * cs) + ((1 - alpha) * L.sugar_mean_income[s]);
    L.spice_mean_income[s] = (alpha * cp) + ((1 - alpha) * L.spice_mean_income[s]);
}
// doDeath: a creditor that dies keeps its children list in Python (the loan holds the object); remember the
// head of its kin list, which lives on in the pool after the slot was reused
__device__ inline void lending_note_death(const DevCtx &c, int s) {
    const ssb_lending_t &L = c.x->lending;
    if (L.debtors_head[s] < 0) return;
    // (atomic: removeDeadAgents' doDeath runs in the parallel compaction kernel)
    const long long n = (long long)atomicAdd((unsigned long long *)&L.counters[SSB_LC_N_DEAD], 1ull);
    const long long at = n % L.dead_capacity;
    // ring: the entry overwritten must be older than any loan of that creditor can still be outstanding
    if (n >= L.dead_capacity && c.t - L.dead_time[at] <= L.max_loan_duration + 1) L.counters[SSB_LC_OVERFLOW] = 2;
    L.dead_slot[at] = s; L.dead_id[at] = c.a.id[s]; L.dead_children_head[at] = c.a.children_head[s]; L.dead_time[at] = c.t;
}
// addLoanToAgent (agent.py:199-211) incl. addLoanFromAgent (190-197)
__device__ inline void add_loan(const DevCtx &c, int creditor, int debtor, int origin, double sugar_principal, double sugar_loan,
                                double spice_principal, double spice_loan, int duration) {
    const ssb_lending_t &L = c.x->lending;
    const ssb_agents_t &a = c.a;
    const Loan owed = {creditor, a.id[creditor], sugar_loan, spice_loan, duration, origin};
    const Loan due = {debtor, a.id[debtor], sugar_loan, spice_loan, duration, origin};
    loan_append(L, L.creditors_head, debtor, owed);
    loan_append(L, L.debtors_head, creditor, due);
    a.sugar[creditor] -= sugar_principal;
    a.spice[creditor] -= spice_principal;
    a.sugar[debtor] = a.sugar[debtor] + sugar_principal;
    a.spice[debtor] = a.spice[debtor] + spice_principal;
}
// payDebtToCreditorChildren (agent.py:1362-1377): the loan is split among the dead creditor's living children
// (the debtor itself excluded), each as a new one-step loan, in the children list's order
__device__ __noinline__ void pay_debt_to_creditor_children(const DevCtx &c, int s, const Loan &loan) {
    const ssb_lending_t &L = c.x->lending;
    const ssb_agents_t &a = c.a;
    int head = -1;
    if (a.id[loan.other_slot] == loan.other_id) head = a.children_head[loan.other_slot];
    else {
        const long long n = L.counters[SSB_LC_N_DEAD], lo = max(n - (long long)L.dead_capacity, 0LL);
        for (long long i = n - 1; i >= lo; i--) {          // newest first
            const long long at = i % L.dead_capacity;
            if (L.dead_slot[at] == loan.other_slot && L.dead_id[at] == loan.other_id) { head = L.dead_children_head[at]; break; }
        }
    }
    int32_t kids[256];
    int n = 0;
    for (int e = head; e >= 0 && n < 256; e = a.kin_next[e]) {          // newest first (kin_push prepends)
        const int o = a.kin_slot[e];
        if (a.id[o] == a.kin_id[e] && agent_alive(c, o) && o != s) kids[n++] = o;
    }
    if (n > 0) {
        const double sugar_rep = loan.sugar / n, spice_rep = loan.spice / n;
        for (int i = n - 1; i >= 0; i--) add_loan(c, kids[i], s, a.last_moved[s], 0, sugar_rep, 0, spice_rep, 1);   // append order
    }
    loan_remove_first(L, L.creditors_head, s, loan);
    if (a.id[loan.other_slot] == loan.other_id) {
        const Loan mirror = {s, a.id[s], loan.sugar, loan.spice, loan.duration, loan.origin};
        loan_remove_first(L, L.debtors_head, loan.other_slot, mirror);
    }
}
// payDebt (agent.py:1330-1360); `loan` is an element of the creditors list of s
__device__ __noinline__ void pay_debt(const DevCtx &c, int s, const Loan &loan) {
    const ssb_lending_t &L = c.x->lending;
    const ssb_agents_t &a = c.a;
    const int cr = loan.other_slot;
    const Loan mirror = {s, a.id[s], loan.sugar, loan.spice, loan.duration, loan.origin};
    const bool same_object = a.id[cr] == loan.other_id;       // the creditor's record still exists
    if (!loan_other_alive(c, loan)) {
        if (c.p.inheritance_policy == 1) { pay_debt_to_creditor_children(c, s, loan); return; }
        loan_remove_first(L, L.creditors_head, s, loan);
        if (same_object) loan_remove_first(L, L.debtors_head, cr, mirror);
    } else if (a.sugar[s] - loan.sugar > 0 && a.spice[s] - loan.spice > 0) {
        a.sugar[s] -= loan.sugar;
        a.spice[s] -= loan.spice;
        a.sugar[cr] = a.sugar[cr] + loan.sugar;
        a.spice[cr] = a.spice[cr] + loan.spice;
        loan_remove_first(L, L.creditors_head, s, loan);
        loan_remove_first(L, L.debtors_head, cr, mirror);
    } else {
        const double sugar_payout = a.sugar[s] / 2, spice_payout = a.spice[s] / 2;
        const double sugar_left = loan.sugar - sugar_payout, spice_left = loan.spice - spice_payout;
        a.sugar[s] -= sugar_payout;
        a.spice[s] -= spice_payout;
        a.sugar[cr] = a.sugar[cr] + sugar_payout;
        a.spice[cr] = a.spice[cr] + spice_payout;
        const double rate = L.lending_factor[cr] * L.base_interest_rate[cr];
        const double new_sugar = sugar_left + (rate * sugar_left), new_spice = spice_left + (rate * spice_left);
        loan_remove_first(L, L.creditors_head, s, loan);
        loan_remove_first(L, L.debtors_head, cr, mirror);
        // compounded on the previous loan, no new principal
        add_loan(c, cr, s, a.last_moved[s], 0, new_sugar, 0, new_spice, L.loan_duration[cr]);
    }
}
// updateLoans (agent.py:1532-1541): Python for-loops over lists that shrink / grow underneath
__device__ inline void update_loans(const DevCtx &c, int s) {
    const ssb_lending_t &L = c.x->lending;
    for (int p = 0;; p++) {
        const int e = loan_at(L, L.debtors_head[s], p);
        if (e < 0) break;
        const Loan l = loan_load(L, e);
        if (!loan_other_alive(c, l)) loan_remove_first(L, L.debtors_head, s, l);
    }
    for (int p = 0;; p++) {
        const int e = loan_at(L, L.creditors_head[s], p);
        if (e < 0) break;
        const Loan l = loan_load(L, e);
        const int remaining = (c.a.last_moved[s] - l.origin) - l.duration;
        if (remaining == 0) pay_debt(c, s, l);
    }
}
__device__ inline double current_debt(const ssb_lending_t &L, int head, bool spice) {       // agent.py:920-930
    double d = 0;
    for (int e = head; e >= 0; e = L.loan_next[e]) d += (spice ? L.loan_spice[e] : L.loan_sugar[e]) / L.loan_duration_v[e];
    return d;
}
// doLending (agent.py:431-483)
template <class Rng>
__device__ __noinline__ void do_lending(const DevCtx &c, int s, Rng &rng) {
    const ssb_lending_t &L = c.x->lending;
    const ssb_agents_t &a = c.a;
    update_loans(c, s);
    // isLender (agent.py:1285-1296)
    if (L.lending_factor[s] == 0) return;
    if (is_fertile_x(c, s) && (a.sugar[s] <= a.start_sugar[s] || a.spice[s] <= a.start_spice[s])) return;
    if (a.age[s] < a.fert_age[s]) return;
    double rate = L.lending_factor[s] * L.base_interest_rate[s];
    if (rate > 1) rate = 1;
    int32_t nb[SSB_MAX_NEIGHBOURS], borrowers[SSB_MAX_NEIGHBOURS];
    int nbw = 0;
    const int n_nb = neighbours4(c, a.pos[s], nb);
    for (int i = 0; i < n_nb; i++) {
        const int o = c.g.occ[nb[i]];
        if (o < 0 || !agent_alive(c, o)) continue;
        // isBorrower (agent.py:1226-1229)
        if (a.age[o] >= a.fert_age[o] && a.age[o] < a.infert_age[o] && !is_fertile_x(c, o)) borrowers[nbw++] = o;
    }
    rng.at(SITE_LENDING);
    shuffle(rng, borrowers, nbw);
    const double pm = eff_spice_metab(c, s), sm = eff_sugar_metab(c, s);
    int loans = 0;
    for (int k = 0; k < nbw; k++) {
        const int b = borrowers[k];
        double max_sugar = a.sugar[s] / 2, max_spice = a.spice[s] / 2;
        if (is_fertile_x(c, s)) {
            max_sugar = a.sugar[s] - a.start_sugar[s] > 0 ? a.sugar[s] - a.start_sugar[s] : 0;
            max_spice = a.spice[s] - a.start_spice[s] > 0 ? a.spice[s] - a.start_spice[s] : 0;
        }
        if (max_sugar == 0 && max_spice == 0) return;           // (skips the lastLoans bookkeeping, like the reference)
        const double sugar_need = a.start_sugar[b] - a.sugar[b] > 0 ? a.start_sugar[b] - a.sugar[b] : 0;
        const double spice_need = a.start_spice[b] - a.spice[b] > 0 ? a.start_spice[b] - a.spice[b] : 0;
        const double sugar_principal = max_sugar < sugar_need ? max_sugar : sugar_need;
        const double spice_principal = max_spice < spice_need ? max_spice : spice_need;
        const double sugar_amount = sugar_principal + (sugar_principal * rate);
        const double spice_amount = spice_principal + (spice_principal * rate);
        if ((sugar_need == 0 && spice_need == 0) || (sugar_amount == 0 && spice_amount == 0)) continue;
        if (a.sugar[s] - sugar_principal <= sm || a.spice[s] - spice_principal <= pm) continue;
        // isCreditWorthy (agent.py:1231-1242)
        const int duration = L.loan_duration[s];
        if (duration == 0) continue;
        const double bpm = eff_spice_metab(c, b), bsm = eff_sugar_metab(c, b);
        const double sugar_cost = sugar_amount / duration, spice_cost = spice_amount / duration;
        const double sugar_income = ((L.sugar_mean_income[b] - bsm) - current_debt(L, L.creditors_head[b], false)) - sugar_cost;
        const double spice_income = ((L.spice_mean_income[b] - bpm) - current_debt(L, L.creditors_head[b], true)) - spice_cost;
        if (!(sugar_income >= 0 && spice_income >= 0)) continue;
        add_loan(c, s, b, a.last_moved[s], sugar_principal, sugar_amount, spice_principal, spice_amount, duration);
        loans++;
        group_note(c, s, SSB_GI_LENDING, b);
    }
    if (loans > 0) { L.last_lended[s] = c.t; L.last_loans[s] = loans; }
}

// ---------------------------------------------------------------------------------------------
// per-agent log
// ---------------------------------------------------------------------------------------------
__device__ inline void agentlog_reset_slot(const DevCtx &c, int s) {
    const ssb_agentlog_t &A = c.x->agentlog;
    A.time_to_live[s] = 0; A.last_pollution[s] = 0; A.prey_wealth[s] = 0; A.last_combat[s] = -1; A.last_mates[s] = 0;
    A.valid_moves[s] = 0; A.move_rank[s] = 0; A.n_neighbors[s] = 0;
}
// updateNeighbors (agent.py:1563-1565): the occupants of the von Neumann neighbours, N S E W
__device__ inline void agentlog_note_neighbors(const DevCtx &c, int s, int pos) {
    const ssb_agentlog_t &A = c.x->agentlog;
    int32_t nb4[SSB_MAX_NEIGHBOURS];      // (the record buffer holds four per agent: Moore + agentLogfile is refused by the host)
    int n = 0;
    const int n_nb = neighbours4(c, pos, nb4);
    for (int i = 0; i < n_nb; i++) { const int o = c.g.occ[nb4[i]]; if (o >= 0) A.neighbor_slot[(long long)s * 4 + n++] = o; }
    A.n_neighbors[s] = n;
}
// Agent.updateRuntimeStats (agent.py:1567-1620), at the end of a completed transaction
__device__ __noinline__ void agentlog_record(const DevCtx &c, int s) {
    const ssb_agentlog_t &A = c.x->agentlog;
    const ssb_agents_t &a = c.a;
    const double last_ttl = A.time_to_live[s];
    const double ttl = time_to_live_x(c, s, false);          // (stores agent.timeToLive)
    // Bentham.updateValues -> updateSelfishnessFactor (ethics.py:246-256) runs right after updateRuntimeStats and
    // compares exactly these two values; nothing in the record depends on the selfishness
    if (c.eth != nullptr && c.eth->model[s] == 1 && c.eth->dynamic_selfishness[s] != 0) {
        double sf = c.eth->selfishness[s];
        if (ttl < last_ttl && sf < 1.0) sf += c.eth->dynamic_selfishness[s];
        else if (ttl > last_ttl && sf > 0.0) sf -= c.eth->dynamic_selfishness[s];
        c.eth->selfishness[s] = rint(sf * 100.0) / 100.0;    // round(x, 2)
    }
    // Temperance.updateValues -> updateAgentTemperanceRules (ethics.py:518-547); the neighbourhood always holds the agent
    // itself, so its "seen by the community" branches are always taken
    if (c.eth != nullptr && c.eth->model[s] == 3) {
        int32_t *rules = c.eth->pecs_rules + (long long)s * 5, *counts = c.eth->pecs_counts + (long long)s * 3;
        const double d = c.eth->last_delta_ttl[s];
        if (d <= 1) rules[0] += 1;
        else if (d > 1 && d <= 2) { rules[1] += 1; counts[0] += 1; rules[2] += 1; }
        else if (d > 2) { rules[3] += 1; counts[1] += 1; rules[4] += 1; }
    }
    if (!A.write_records) return;
    const long long at = A.n_records[0];
    if (at >= A.record_capacity) return;
    A.n_records[0] = at + 1;
    double *r = A.records + at * SSB_AL_FIELDS;
    const int t = c.t;
    int in_tribe = 0, same_race = 0, exp_n = 0, ctrl_n = 0;
    const int my_race = has_misc(c) ? c.x->misc.race[s] : -1;
    for (int i = 0; i < A.n_neighbors[s]; i++) {
        const int o = A.neighbor_slot[(long long)s * 4 + i];
        if (a.tribe[o] == a.tribe[s]) in_tribe++;
        if ((has_misc(c) ? c.x->misc.race[o] : -1) == my_race) same_race++;
        if (has_group(c)) { if (is_member(c, o)) exp_n++; else ctrl_n++; }
    }
    r[SSB_AL_TIMESTEP] = t; r[SSB_AL_ID] = a.id[s]; r[SSB_AL_AGE] = a.age[s];
    r[SSB_AL_SUGAR] = a.sugar[s]; r[SSB_AL_SPICE] = a.spice[s];
    r[SSB_AL_SUGAR_GAINED] = a.sugar[s] - a.last_sugar[s]; r[SSB_AL_SPICE_GAINED] = a.spice[s] - a.last_spice[s];
    r[SSB_AL_MOVEMENT] = eff_movement(c, s);
    r[SSB_AL_TIME_TO_LIVE] = ttl;
    r[SSB_AL_DEPRESSION] = has_misc(c) ? c.x->misc.depressed[s] : 0;
    r[SSB_AL_HAPPINESS] = a.happiness[s];
    r[SSB_AL_PREY_KILLED] = A.last_combat[s] == t;
    r[SSB_AL_PREY_WEALTH] = A.last_combat[s] == t ? A.prey_wealth[s] : 0;
    r[SSB_AL_TRADE_PARTNERS] = ((c.p.flags & SSB_F_TRADE) && a.trade_factor[s] != 0) ? a.trade_volume[s] : 0;
    r[SSB_AL_DISEASES_SPREAD] = (has_disease(c) && c.x->diseases.last_spread[s] == t) ? c.x->diseases.n_spread[s] : 0;
    r[SSB_AL_MATES] = a.last_repro[s] == t ? A.last_mates[s] : 0;
    r[SSB_AL_NEIGHBORS] = A.n_neighbors[s];
    r[SSB_AL_VALID_MOVES] = A.valid_moves[s]; r[SSB_AL_MOVE_RANK] = A.move_rank[s];
    r[SSB_AL_LENDING_PARTNERS] = (has_lending(c) && c.x->lending.last_lended[s] == t) ? c.x->lending.last_loans[s] : 0;
    r[SSB_AL_POLLUTION_DIFF] = c.g.pollution[a.pos[s]] - A.last_pollution[s];
    r[SSB_AL_TTL_DIFF] = ttl - last_ttl;
    r[SSB_AL_NEIGHBORS_IN_TRIBE] = in_tribe; r[SSB_AL_SAME_RACE_NEIGHBORS] = same_race;
    r[SSB_AL_EXPERIMENTAL_NEIGHBORS] = exp_n; r[SSB_AL_CONTROL_NEIGHBORS] = ctrl_n;
}

// reductions of the families' log keys, list order (sugarscape.py:1130-1160, 1162, 1200-1211, 1228, 1254-1277)
__device__ inline void ext_stats_body(const DevCtx &c, ssb_ext_stats_t *out, int pass = 0) {
    ssb_ext_stats_t r;
    r.loan_volume = 0; r.sum_social_happiness = 0; r.sum_conflict_happiness = 0;
    r.sick_agents = 0; r.disease_incidence = 0; r.distinct_infectors = 0; r.disease_prevalence = 0; r.disease_deaths = 0;
    bool none_infector = false;
    r.sum_health_happiness = 0; r.move_space = 0; r.sum_burn_rate = 0; r.sum_ttl_age_limited = 0;
    for (int k = 0; k < 4; k++) { r.race_count[k] = 0; r.race_first[k] = -1; }
    const long long n = c.a.counters[SSB_C_N_LIST], n_dead = c.a.counters[SSB_C_N_DEAD];
    const ssb_diseases_t &D = c.x->diseases;
    // isDiseaseExperimentalGroup() is false for every disease unless the group names one, so the pass over the
    // experimental group skips all incidence / prevalence bookkeeping (sugarscape.py:1200-1211, 1266-1277)
    const bool skip_diseases = pass == 1;
    const int mark = c.t * 4 + pass;
    for (long long i = 0; i < n; i++) {
        const int s = c.a.order[i];
        if (!in_pass(c, s, pass)) continue;
        if (has_lending(c) && c.x->lending.last_lended[s] == c.t) r.loan_volume += c.x->lending.last_loans[s];
        if (has_social(c)) { r.sum_social_happiness += c.x->social.social_happiness[s]; r.sum_conflict_happiness += c.x->social.conflict_happiness[s]; }
        if (has_misc(c) && c.x->misc.racial_len > 0) {
            const int race = c.x->misc.race[s];
            if (race >= 0 && race < 4 && r.race_count[race]++ == 0) r.race_first[race] = i;
        }
        if (!has_disease(c) && has_misc(c)) r.sum_health_happiness += c.x->misc.health_happiness[s];
        // (sugarscape.py:1111-1112 leaves the age-limited value in agent.timeToLive)
        if (has_agentlog(c) && !(has_disease(c) || has_misc(c))) { time_to_live_x(c, s, false); time_to_live_x(c, s, true); }
        if (has_disease(c) || has_misc(c)) {      // findTimeToLive with the modifiers / the universal income term
            const double ttl = time_to_live_x(c, s, false);
            r.sum_burn_rate += ttl;
            r.sum_ttl_age_limited += fmin(ttl, (double)(c.a.max_age[s] - c.a.age[s]));
            if (has_agentlog(c)) time_to_live_x(c, s, true);
        }
        if (has_disease(c)) {
            const long long b = (long long)s * D.max_sick;
            if (D.n_sick[s] > 0) r.sick_agents++;
            for (int k = 0; k < D.n_sick[s] && !skip_diseases; k++)
                if (D.sick_caught[b + k] == c.t) {
                    r.disease_incidence++;
                    const int inf = D.sick_infector_slot[b + k];
                    // distinct infectors: within one step a slot names one acting agent
                    if (c.t != 0 && inf >= 0 && D.stat_mark[inf] != mark) { D.stat_mark[inf] = mark; r.distinct_infectors++; }
                    // (a disease that started this step came from nobody: `None` is one more member of the reference's set)
                    if (c.t != 0 && inf < 0 && !none_infector) { none_infector = true; r.distinct_infectors++; }
                }
            r.sum_health_happiness += D.health_happiness[s];
            r.move_space += D.last_valid_moves[s];
        }
    }
    if (has_disease(c)) {
        for (long long i = 0; i < n_dead; i++) if (in_pass(c, c.a.dead_list[i], pass)) r.disease_deaths += D.disease_death[c.a.dead_list[i]];    // (their disease lists are empty)
        for (int d = 0; d < D.n_diseases && !skip_diseases; d++) r.disease_prevalence += D.defs[d].listed * D.defs[d].infected;
    }
    *out = r;
}

// Interaction counters of a pass (sugarscape.py:1175-1198, 1230-1250): summed over the pass's live agents (then reset)
// and this step's dead; plus the control / experimental members of the movement neighbourhoods.
__device__ inline void group_counters_body(const DevCtx &c, int pass, ssb_group_stats_t *out) {
    const ssb_group_t &G = c.x->group;
    for (int k = 0; k < SSB_GI_KINDS; k++) { out->to_control[k] = 0; out->to_experimental[k] = 0; }
    out->control_neighbors = 0; out->experimental_neighbors = 0;
    const long long n = c.a.counters[SSB_C_N_LIST], n_dead = c.a.counters[SSB_C_N_DEAD];
    for (long long i = 0; i < n; i++) {
        const int s = c.a.order[i];
        if (!in_pass(c, s, pass)) continue;
        for (int k = 0; k < SSB_GI_KINDS; k++) {
            const long long at = (long long)s * SSB_GI_KINDS + k;
            out->to_control[k] += G.with_control[at];
            out->to_experimental[k] += G.with_experimental[at];
            if (pass != 0) { G.with_control[at] = 0; G.with_experimental[at] = 0; }
        }
        out->control_neighbors += G.control_neighbors[s]; out->experimental_neighbors += G.experimental_neighbors[s];
    }
    for (long long i = 0; i < n_dead; i++) {
        const int s = c.a.dead_list[i];
        if (!in_pass(c, s, pass)) continue;
        for (int k = 0; k < SSB_GI_KINDS; k++) {
            out->to_control[k] += G.with_control[(long long)s * SSB_GI_KINDS + k];
            out->to_experimental[k] += G.with_experimental[(long long)s * SSB_GI_KINDS + k];
        }
    }
}

// The reductions of updateRuntimeStatsPerGroup (sugarscape.py:1005-1426) for the agents of one pass, list order.
// The Lorenz sum and the environment totals stay 0: a group pass does not log them (sugarscape.py:1403-1417).
__device__ inline void stats_pass_body(const DevCtx &c, int pass, ssb_stats_t *out) {
    const ssb_agents_t &a = c.a;
    ssb_stats_t r;
    memset(&r, 0, sizeof(r));
    const long long n = a.counters[SSB_C_N_LIST], n_dead = a.counters[SSB_C_N_DEAD];
    const int t = c.t;
    r.timestep = t;
    for (int i = 0; i < 32; i++) r.tribe_first[i] = -1;
    long long seen = 0;
    for (long long i = 0; i < n; i++) {
        const int s = a.order[i];
        if (!in_pass(c, s, pass)) continue;
        const double w = a.sugar[s] + a.spice[s];
        r.sum_sugar_metab += a.sugar_metab[s]; r.sum_spice_metab += a.spice_metab[s];
        r.sum_movement += a.movement[s]; r.sum_vision += a.vision[s]; r.sum_age += a.age[s];
        r.n_acted += a.age[s] > 0;
        r.sum_wealth += w;
        if (seen == 0 || w > r.max_wealth) r.max_wealth = w;
        if (seen == 0 || w < r.min_wealth) r.min_wealth = w;
        seen++;
        r.sum_wealth_collected += w - (a.last_sugar[s] + a.last_spice[s]);
        r.sum_burn_rate += time_to_live_x(c, s, false);
        r.sum_ttl_age_limited += time_to_live_x(c, s, true);
        r.sum_neighbors += a.n_neighborhood[s];
        r.sum_valid_moves += a.n_valid_moves[s];
        if (a.trade_volume[s] > 0) { r.n_traders++; r.trade_volume += a.trade_volume[s]; r.sum_trade_price += a.trade_price[s]; }
        r.sum_happiness += a.happiness[s];
        r.sum_wealth_happiness += a.wealth_happiness[s];
        r.sum_family_happiness += a.family_happiness[s];
        const int tr = a.tribe[s];
        if (tr >= 0 && tr < 32) { if (r.tribe_count[tr]++ == 0) r.tribe_first[tr] = i; }
    }
    r.population = seen;
    for (long long i = 0; i < n_dead; i++) {
        const int s = a.dead_list[i];
        if (!in_pass(c, s, pass)) continue;
        r.n_dead++;
        if (a.last_moved[s] == t) r.n_dead_moved++;
        r.sum_age_at_death += a.age[s];
        r.dead_wealth_collected += (a.sugar[s] + a.spice[s]) - (a.last_sugar[s] + a.last_spice[s]);
        r.dead_starvation += a.cod[s] == SSB_STARVATION;
        r.dead_aging += a.cod[s] == SSB_AGING;
        r.dead_combat += a.cod[s] == SSB_COMBAT;
    }
    // bornAgents of the pass (sugarscape.py:1359-1362): this step's newborn, alive or not
    for (long long i = 0; i < n; i++) { const int s = a.order[i]; if (a.born[s] == t && t > 0 && in_pass(c, s, pass)) r.n_born++; }
    for (long long i = 0; i < n_dead; i++) { const int s = a.dead_list[i]; if (a.born[s] == t && t > 0 && in_pass(c, s, pass)) r.n_born++; }
    r.n_replaced = a.counters[SSB_C_N_REPLACED];
    *out = r;
}

}  // namespace ssb
