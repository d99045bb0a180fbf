"""Host half of updateRuntimeStats: raw device reductions -> the reference's 59 log keys.

The device (K6, csrc/ssb_kernels.cuh) produces sums / extrema / counts; this module applies the
reference's arithmetic on top (means, Python ``round()``, the carrying-capacity recurrence, the
closed-form ``*LastProduced`` totals) so the resulting dict equals ``Sugarscape.runtimeStats``
(reference sugarscape.py:81-92 key order, 1005-1426 semantics).
"""
import math

# reference sugarscape.py:81-92
LOG_KEYS = (
    "timestep", "population", "meanMetabolism", "meanMovement", "meanVision", "meanWealth", "meanAge",
    "giniCoefficient", "meanTradePrice", "tradeVolume", "maxWealth", "minWealth", "meanHappiness",
    "meanWealthHappiness", "meanHealthHappiness", "meanSocialHappiness", "meanFamilyHappiness",
    "meanConflictHappiness", "meanAgeAtDeath", "seed", "agentsReplaced", "agentsBorn",
    "agentStarvationDeaths", "agentDiseaseDeaths", "environmentWealthCreated", "agentWealthTotal",
    "environmentWealthTotal", "agentWealthCollected", "agentWealthBurnRate", "agentMeanTimeToLive",
    "agentTotalMetabolism", "agentCombatDeaths", "agentAgingDeaths", "agentDeaths", "largestRace",
    "largestTribe", "largestRaceSize", "largestTribeSize", "remainingRaces", "remainingTribes",
    "sickAgents", "carryingCapacity", "meanDeathsPercentage", "sickAgentsPercentage", "meanSelfishness",
    "diseaseEffectiveReproductionRate", "diseaseIncidence", "diseasePrevalence",
    "agentLastMoveOptimalityPercentage", "meanNeighbors", "meanMoveRank", "meanMoveDifferenceFromOptimal",
    "meanValidMoves", "totalHappiness", "loanVolume", "meanAgeismFactor", "meanRacismFactor",
    "meanSexismFactor", "moveSpace",
)

# meanConflictHappiness is reported as 0: conflict happiness needs live combat (aggression > 0),
# which none of the supported shipped examples has.  Family happiness is kept in exact mode only.
UNVERIFIED_KEYS = ()


INTERACTIONS = ("combat", "disease", "lending", "reproduction", "trade")      # SSB_GI_* order
GROUP_NEIGHBOR_KEYS = ("meanControlNeighbors", "meanExperimentalNeighbors")
# keys a group pass does not compute: the log keeps their initial 0 (sugarscape.py:1403-1417)
GROUP_TEMPLATE_ZERO = ("timestep", "giniCoefficient", "seed", "environmentWealthCreated", "environmentWealthTotal")


def _prefixed(prefix, key):
    return prefix + key[0].upper() + key[1:]


def initial_runtime_stats(c):
    """The all-zero entry the reference logs at t = 0 (sugarscape.py:81-92, 887-905); with an experimental group
    the statistics are repeated for the group and for its complement "control" (998-1003): 203 keys."""
    s = {k: 0 for k in LOG_KEYS}
    s["seed"] = c["seed"]
    s["remainingRaces"] = c["environmentMaxRaces"]
    s["remainingTribes"] = c["environmentMaxTribes"]
    group = c["experimentalGroup"]
    if group is not None:
        for k in GROUP_NEIGHBOR_KEYS:
            s[k] = 0
        for prefix in ("control", group):
            for k in tuple(LOG_KEYS) + GROUP_NEIGHBOR_KEYS:
                s[_prefixed(prefix, k)] = 0
        for kind in INTERACTIONS:
            for side in ("Control", "Experimental"):
                s[f"{kind}{side}GroupToControlGroup"] = 0
                s[f"{kind}{side}GroupToExperimentalGroup"] = 0
    return s


def _as_number(x):
    """Values that are integral by construction print as ints in the reference's JSON; the
    device carries them in fp64."""
    return int(x) if float(x).is_integer() and abs(x) < 2 ** 53 else x


def environment_wealth_created(c, equator, t, n_cells_north, n_cells_south, capacity_total):
    """sum(sugarLastProduced + spiceLastProduced) in closed form:
    a cell's *LastProduced equals the regrow rate once the cell has been allowed to grow at
    least once (it is never cleared afterwards), plus all capacities at t == 1
    (sugarscape.py:1048-1051)."""
    if t <= 0:
        return 0
    rate = 0
    for key in ("environmentSugarRegrowRate", "environmentSpiceRegrowRate"):
        rate += c[key] if c[key] != 0 else 0
    interval, delay = c["environmentSeasonInterval"], c["environmentSeasonalGrowbackDelay"]
    if interval <= 0:
        created = (n_cells_north + n_cells_south) * rate
    else:
        # north starts wet (grows from t = 1); south starts dry: first growth at the first
        # t with t % delay == 0, or when it turns wet (t = interval + 1)
        south_first = interval + 1
        if delay > 0:
            south_first = min(south_first, delay)
        created = n_cells_north * rate + (n_cells_south * rate if t >= south_first else 0)
    if t == 1:
        created += capacity_total
    return created


def agent_log_entry(r):
    """One entry of the per-agent log (agent.py:1609-1618) from a raw device record (layout.AL_FIELDS values)."""
    (t, ident, age, sugar, spice, sugar_gained, spice_gained, movement, ttl, depression, happiness, prey_killed, prey_wealth,
     trade_partners, diseases_spread, mates, neighbors, valid_moves, move_rank, lending_partners, pollution_diff, ttl_diff,
     in_tribe, same_race, experimental, control) = [float(v) for v in r]
    neighbors = int(neighbors)
    return {"timestep": int(t), "ID": int(ident), "age": int(age), "wealth": _as_number(round(sugar + spice, 2)),
            "sugar": _as_number(round(sugar, 2)), "spice": _as_number(round(spice, 2)),
            "sugarGained": _as_number(round(sugar_gained, 2)), "spiceGained": _as_number(round(spice_gained, 2)),
            "wealthGained": _as_number(round(spice_gained + sugar_gained, 2)), "movement": _as_number(movement),
            "timeToLive": _as_number(round(ttl, 1)), "depression": bool(depression),
            "compositeHappiness": _as_number(round(happiness, 1)), "preyKilled": bool(prey_killed),
            "preyWealth": _as_number(prey_wealth), "tradePartners": int(trade_partners), "diseasesSpread": int(diseases_spread),
            "mates": int(mates), "neighbors": neighbors, "validMoves": int(valid_moves), "moveRank": int(move_rank),
            "lendingPartners": int(lending_partners), "pollutionDifference": _as_number(pollution_diff),
            "timeToLiveDifference": _as_number(ttl_diff), "neighborsInTribe": int(in_tribe),
            "neighborsNotInTribe": neighbors - int(in_tribe), "sameRaceNeighbors": int(same_race),
            "differentRaceNeighbors": neighbors - int(same_race), "experimentalGroupNeighbors": int(experimental),
            "controlGroupNeighbors": int(control)}


class StatsBuilder:
    """Carries the cross-step state of the stats (carrying capacity, previous dict)."""

    def __init__(self, c, equator):
        self.c = c
        self.equator = equator
        self.runtime_stats = initial_runtime_stats(c)
        W, H = c["environmentWidth"], c["environmentHeight"]
        north_rows = min(max(equator, 0), H)
        self.n_north = W * north_rows
        self.n_south = W * (H - north_rows)
        self.diseased = c["startingDiseases"] > 0 and c["agentImmuneSystemLength"] > 0
        self.racial = c["agentRacialTagStringLength"] > 0 and c["environmentMaxRaces"] > 0
        self.misc = (self.racial or c["agentUniversalSugar"][1] != 0 or c["agentUniversalSpice"][1] != 0 or
                     c["agentDepressionPercentage"] > 0)
        self.constant = {}
        for key, option in (("meanSelfishness", "agentSelfishnessFactor"),
                            ("meanAgeismFactor", "agentDecisionModelAgeismFactor"),
                            ("meanRacismFactor", "agentDecisionModelRacismFactor"),
                            ("meanSexismFactor", "agentDecisionModelSexismFactor")):
            lo, hi = c[option]
            self.constant[key] = lo if lo == hi else None

    def update(self, st, t, n_replaced=0, ethics=None, ext=None):
        """st: layout.Stats filled by the device (or by the oracle in the tests); ethics: layout.EthicsStats
        when the run uses decision models (sugarscape.py:1141-1160, 1218-1222, 1304, 1318, 1323-1324); ext:
        layout.ExtStats when lending / friends are live (loanVolume, social and conflict happiness)."""
        c, rs = self.c, self.runtime_stats
        n = int(st.population)
        prev_cc = rs["carryingCapacity"]
        carrying = math.ceil(0.05 * n + 0.95 * prev_cc)
        if t == 0:
            carrying = n
        out = {}
        out["agentAgingDeaths"] = int(st.dead_aging)
        out["agentCombatDeaths"] = int(st.dead_combat)
        out["agentDeaths"] = int(st.n_dead)
        out["agentDiseaseDeaths"] = 0
        out["agentsBorn"] = int(st.n_born) if t > 0 else 0
        out["agentsReplaced"] = int(n_replaced)
        out["agentStarvationDeaths"] = int(st.dead_starvation)
        out["agentTotalMetabolism"] = int(st.sum_sugar_metab + st.sum_spice_metab)
        out["agentWealthCollected"] = _as_number(st.sum_wealth_collected + st.dead_wealth_collected)
        out["carryingCapacity"] = carrying
        out["meanAgeAtDeath"] = round(st.sum_age_at_death / st.n_dead, 2) if st.n_dead > 0 else 0
        out["sickAgents"] = 0
        out["diseaseIncidence"] = 0
        out["diseasePrevalence"] = 0
        out["moveSpace"] = int(st.sum_valid_moves)
        out["population"] = n
        moves_optimal = n + int(st.n_dead)
        if n > 0:
            out["agentMeanTimeToLive"] = round(st.sum_ttl_age_limited / n, 2)
            out["agentWealthBurnRate"] = round(st.sum_burn_rate / n, 2)
            out["agentWealthTotal"] = _as_number(round(st.sum_wealth, 2))
            out["loanVolume"] = 0
            counts = [(int(st.tribe_first[i]), i, int(st.tribe_count[i])) for i in range(32) if st.tribe_count[i] > 0]
            if counts:
                # max(tribes, key=tribes.get): the first-inserted tribe among the largest
                best = max(cnt for _, _, cnt in counts)
                first, tribe, _ = min((f, i, cnt) for f, i, cnt in counts if cnt == best)
                out["largestTribe"], out["largestTribeSize"], out["remainingTribes"] = tribe, best, len(counts)
            else:
                out["largestTribe"], out["largestTribeSize"], out["remainingTribes"] = None, n, 1
            out["largestRace"], out["largestRaceSize"], out["remainingRaces"] = None, n, 1
            out["maxWealth"] = _as_number(round(st.max_wealth, 2))
            out["minWealth"] = _as_number(round(st.min_wealth, 2))
            out["meanAge"] = round(st.sum_age / n, 2)
            out["meanConflictHappiness"] = round(0 / n, 2)
            out["meanFamilyHappiness"] = round(st.sum_family_happiness / n, 2)
            out["meanHappiness"] = round(st.sum_happiness / n, 2)
            out["meanHealthHappiness"] = round(st.n_acted / n, 2)
            combined = st.sum_sugar_metab + st.sum_spice_metab
            if st.sum_sugar_metab > 0 and st.sum_spice_metab > 0:
                combined = round(combined / 2, 2)
            out["meanMetabolism"] = round(combined / n, 2)
            out["meanMovement"] = round(st.sum_movement / n, 2)
            out["meanSocialHappiness"] = round(0 / n, 2)
            out["meanTradePrice"] = round(st.sum_trade_price / st.n_traders, 2) if st.n_traders > 0 else 0
            out["meanVision"] = round(st.sum_vision / n, 2)
            out["meanWealth"] = round(st.sum_wealth / n, 2)
            out["meanWealthHappiness"] = round(st.sum_wealth_happiness / n, 2)
            out["tradeVolume"] = int(st.trade_volume)
            out["meanDeathsPercentage"] = round((st.n_dead / n) * 100, 2)
            out["sickAgentsPercentage"] = round((0 / n) * 100, 2)
            out["diseaseEffectiveReproductionRate"] = 0
            out["agentLastMoveOptimalityPercentage"] = round((moves_optimal / moves_optimal) * 100, 2)
            out["meanNeighbors"] = round(st.sum_neighbors / n, 2)
            out["meanValidMoves"] = round(st.sum_valid_moves / n, 2)
            out["meanMoveRank"] = round(0 / n, 2)
            out["meanMoveDifferenceFromOptimal"] = round(0 / out["meanNeighbors"], 2) if out["meanNeighbors"] > 0 else 0
            out["totalHappiness"] = st.sum_happiness
            for key, value in self.constant.items():
                if key == "meanSelfishness" and ethics is not None:
                    continue
                if value is None:
                    raise NotImplementedError(key + " needs per-agent factors")
                out[key] = round(value * n / n, 2)
            if ext is not None:
                out["loanVolume"] = int(ext.loan_volume)
                out["meanSocialHappiness"] = round(ext.sum_social_happiness / n, 2)
                out["meanConflictHappiness"] = round(ext.sum_conflict_happiness / n, 2)
            if ext is not None and self.racial:
                # max(races, key=races.get): the first race, in list order, among the largest (sugarscape.py:1167-1170, 1287-1288)
                races = [(int(ext.race_first[r]), r, int(ext.race_count[r])) for r in range(4) if ext.race_count[r] > 0]
                if races:
                    best = max(cnt for _, _, cnt in races)
                    _, race, _ = min((f, r, cnt) for f, r, cnt in races if cnt == best)
                    out["largestRace"], out["largestRaceSize"], out["remainingRaces"] = race, best, len(races)
            if ext is not None and self.misc and not self.diseased:
                out["meanHealthHappiness"] = round(ext.sum_health_happiness / n, 2)
                out["agentMeanTimeToLive"] = round(ext.sum_ttl_age_limited / n, 2)
                out["agentWealthBurnRate"] = round(ext.sum_burn_rate / n, 2)
            if ext is not None and self.diseased:
                # sugarscape.py:1162, 1200-1211, 1228, 1254-1277, 1316-1317; time to live with the disease modifiers
                out["sickAgentsPercentage"] = round((ext.sick_agents / n) * 100, 2)
                out["diseaseEffectiveReproductionRate"] = (round(ext.disease_incidence / ext.distinct_infectors, 2)
                                                           if ext.distinct_infectors else 0)
                out["meanHealthHappiness"] = round(ext.sum_health_happiness / n, 2)
                out["agentMeanTimeToLive"] = round(ext.sum_ttl_age_limited / n, 2)
                out["agentWealthBurnRate"] = round(ext.sum_burn_rate / n, 2)
            if ethics is not None:
                out["meanSelfishness"] = round(ethics.sum_selfishness / n, 2)
                out["agentLastMoveOptimalityPercentage"] = round((ethics.optimal_moves / ethics.moves) * 100, 2)
                out["meanMoveRank"] = round(ethics.sum_move_rank / n, 2)
                out["meanMoveDifferenceFromOptimal"] = (round(ethics.sum_move_diff / out["meanNeighbors"], 2)
                                                        if out["meanNeighbors"] > 0 else 0)
        else:
            # zero population: the reference resets some keys and leaves others stale/raw
            # (sugarscape.py:1325-1351)
            for key in ("agentMeanTimeToLive", "agentWealthBurnRate", "largestRace", "largestTribe", "maxWealth",
                        "meanAge", "meanAgeismFactor", "meanConflictHappiness", "meanFamilyHappiness",
                        "meanHappiness", "meanHealthHappiness", "meanMetabolism", "meanMovement",
                        "meanRacismFactor", "meanSelfishness", "meanSexismFactor", "meanSocialHappiness",
                        "meanVision", "meanWealth", "meanWealthHappiness", "minWealth", "remainingRaces",
                        "remainingTribes", "totalHappiness", "tradeVolume", "diseaseEffectiveReproductionRate"):
                out[key] = 0
            out["agentWealthTotal"] = 0
            out["loanVolume"] = 0
            out["largestRaceSize"] = 0
            out["largestTribeSize"] = 0
            out["meanTradePrice"] = 0
            out["meanDeathsPercentage"] = 0
            out["sickAgentsPercentage"] = 0
            # the raw count of optimal moves (of the dead) stays in place (sugarscape.py:1325-1351)
            out["agentLastMoveOptimalityPercentage"] = int(ethics.optimal_moves) if ethics is not None else int(st.n_dead)
            out["meanNeighbors"] = 0
            out["meanValidMoves"] = 0
            out["meanMoveRank"] = 0
            out["meanMoveDifferenceFromOptimal"] = 0
        if ext is not None and self.diseased:
            out["sickAgents"] = int(ext.sick_agents)
            out["diseaseIncidence"] = int(ext.disease_incidence)
            out["diseasePrevalence"] = int(ext.disease_prevalence)
            out["agentDiseaseDeaths"] = int(ext.disease_deaths)
            out["moveSpace"] = int(ext.move_space)
        for key, value in out.items():
            rs[key] = value
        rs["environmentWealthCreated"] = environment_wealth_created(
            c, self.equator, t, self.n_north, self.n_south, int(st.env_capacity_total))
        rs["environmentWealthTotal"] = int(st.env_wealth_total)
        rs["giniCoefficient"] = self.gini(st, n)
        rs["timestep"] = t
        return rs

    def update_with_groups(self, st, t, n_replaced, ethics, ext, gstats, n_replaced_members=0):
        """The log entry with an experimental group (sugarscape.py:998-1003): the statistics of the group, of its
        complement and of everybody.  gstats: layout.GroupStats[3] (everybody's neighbour sums, group, control).
        A group pass starts from everybody's previous entry and its cross-step state is discarded."""
        import copy
        group = self.c["experimentalGroup"]
        keys = tuple(LOG_KEYS) + GROUP_NEIGHBOR_KEYS
        passes = {}
        for which, prefix in ((1, group), (2, "control")):
            g = gstats[which]
            # agentsReplaced of a pass: this step's replacements that are in the group / in its complement (sugarscape.py:1353-1356)
            n_rep = n_replaced_members if which == 1 else n_replaced - n_replaced_members
            rs = copy.deepcopy(self).update(g.stats, t, n_rep, g.ethics if ethics is not None else None, g.ext)
            rs["agentsBorn"] -= n_rep        # (the pass counts its members born at t: the replacements were placed at t, too)
            n = rs["population"]
            rs["meanControlNeighbors"] = round(g.control_neighbors / n, 2) if n else 0
            rs["meanExperimentalNeighbors"] = round(g.experimental_neighbors / n, 2) if n else 0
            passes[prefix] = (dict(rs), g)
        entry = self.update(st, t, n_replaced, ethics, ext)
        n = entry["population"]
        entry["meanControlNeighbors"] = round(gstats[0].control_neighbors / n, 2) if n else 0
        entry["meanExperimentalNeighbors"] = round(gstats[0].experimental_neighbors / n, 2) if n else 0
        for prefix, (rs, g) in passes.items():
            rs["carryingCapacity"] = entry["carryingCapacity"]          # from len(self.agents), whatever the pass
            for k in keys:
                entry[_prefixed(prefix, k)] = 0 if k in GROUP_TEMPLATE_ZERO else rs[k]
            side = "Control" if prefix == "control" else "Experimental"
            for i, kind in enumerate(INTERACTIONS):
                entry[f"{kind}{side}GroupToControlGroup"] = int(g.to_control[i])
                entry[f"{kind}{side}GroupToExperimentalGroup"] = int(g.to_experimental[i])
        return entry

    @staticmethod
    def gini(st, n):
        """updateGiniCoefficient (sugarscape.py:913-934) from the sorted-wealth reduction."""
        if n == 0:
            return 0
        if st.gini_area >= 0:          # exact mode: the area accumulated on the device in the reference's own order
            return round((0.5 - st.gini_area) / 0.5, 3)
        total = st.sum_wealth
        if total == 0:
            return 1
        area = (st.lorenz_weighted_sum / total) / n
        return round((0.5 - area) / 0.5, 3)
