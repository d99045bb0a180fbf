"""ctypes binding of the CPU oracle (oracle/libss_oracle.so).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import random
import subprocess

import numpy as np

from sugarscape_b200 import layout

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(ROOT, "oracle", "libss_oracle.so")
        src = os.path.join(ROOT, "oracle", "ss_oracle.c")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
            build()
        L = C.CDLL(path)
        L.orc_sizeof_rng.restype = C.c_int
        L.orc_step_key.restype = C.c_uint64
        L.orc_step_key.argtypes = [C.c_uint64, C.c_int32]
        L.orc_priority.restype = C.c_uint32
        L.orc_priority.argtypes = [C.c_uint64, C.c_uint32, C.c_int]
        P, G, A = C.POINTER(layout.Params), C.POINTER(layout.Grid), C.POINTER(layout.Agents)
        L.orc_rng_init.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32]
        L.orc_rng_get.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_int32)]
        L.orc_env_step.argtypes = [P, G, C.c_int32]
        L.orc_agent_step.argtypes = [P, G, A, C.c_void_p, C.c_int32, C.c_double, C.c_void_p, C.c_void_p]
        L.orc_compact.argtypes = [P, G, A, C.c_int32]
        L.orc_replace.argtypes = [P, G, A, C.POINTER(layout.Endowments), C.c_int32]
        L.orc_stats.argtypes = [P, G, A, C.c_int32, C.POINTER(layout.Stats)]
        _LIB = L
    return _LIB


class Oracle:
    """Steps a HostState in place with the CPU restatement."""

    def __init__(self, params, state, endow_bits=None, tag_bits=None):
        self.L = lib()
        self.params = params
        self.state = state
        self.rng = np.zeros(self.L.orc_sizeof_rng() // 4 + 2, dtype=np.uint32)
        self.endow_bits = endow_bits
        self.tag_bits = tag_bits
        self.L.orc_rng_init(self.rng.ctypes.data, None, 624, params.rng_mode)

    def set_mt_from_python(self):
        st = random.getstate()[1]
        arr = np.array(st[:624], dtype=np.uint32)
        self.L.orc_rng_init(self.rng.ctypes.data, arr.ctypes.data, st[624], self.params.rng_mode)

    def mt_state(self):
        arr = np.zeros(624, dtype=np.uint32)
        idx = C.c_int32(0)
        self.L.orc_rng_get(self.rng.ctypes.data, arr.ctypes.data, C.byref(idx))
        return arr, idx.value

    def push_mt_to_python(self):
        arr, idx = self.mt_state()
        random.setstate((3, tuple(int(v) for v in arr) + (idx,), None))

    def env_step(self, t):
        g, a = self.state.as_structs()
        self.L.orc_env_step(C.byref(self.params), C.byref(g), t)

    def agent_step(self, t, prev_mean_wealth=0.0):
        g, a = self.state.as_structs()
        eb = self.endow_bits.ctypes.data if self.endow_bits is not None else None
        tb = self.tag_bits.ctypes.data if self.tag_bits is not None else None
        self.L.orc_agent_step(C.byref(self.params), C.byref(g), C.byref(a), self.rng.ctypes.data,
                              t, prev_mean_wealth, eb, tb)

    def compact(self, t):
        g, a = self.state.as_structs()
        ext = getattr(self.state, "ext", None)
        self.L.orc_keep_children_lists(1 if ext is not None and "lending" in ext.families else 0)
        self.L.orc_compact(C.byref(self.params), C.byref(g), C.byref(a), t)

    def bind_endowments(self, endowments):
        """endowments: list of endowment dicts (init.Population.endowments) or struct-of-arrays."""
        self.endow_arrays = endowments if isinstance(endowments, dict) else layout.endowment_arrays(endowments)
        self.endow = layout.endowments_struct(self.endow_arrays, lambda arr: arr.ctypes.data)

    def bind_lending(self, endowments):
        """Oracle-private lending state (the engine refuses lending configurations): per-slot
        endowments of the initial agents from init.Population.endowments (agent id = index)."""
        cap = self.state.capacity
        f64 = lambda v: np.full(cap, v, dtype=np.float64)
        self.lend = {"lending_factor": f64(0), "base_interest_rate": f64(0), "loan_duration": np.zeros(cap, np.int32),
                     "sugar_mean_income": f64(1), "spice_mean_income": f64(1),
                     "last_lended": np.full(cap, -1, np.int32), "last_loans": np.zeros(cap, np.int32)}
        n = int(self.state.counters[layout.C_N_LIST])
        for s in self.state.a_order[:n]:
            self._set_lending(int(s), endowments[int(self.state.a_id[s])])
        L = self.L
        L.orc_bind_lending.argtypes = [C.c_void_p] * 7 + [C.c_int64]
        L.orc_bind_lending.restype = None
        L.orc_loan_volume.argtypes = [C.c_void_p, C.c_int32]
        L.orc_loan_volume.restype = C.c_int64
        L.orc_bind_lending(*[self.lend[k].ctypes.data for k in ("lending_factor", "base_interest_rate", "loan_duration",
                                                                "sugar_mean_income", "spice_mean_income", "last_lended",
                                                                "last_loans")], cap)

    def _set_lending(self, s, e):
        self.lend["lending_factor"][s] = e["lendingFactor"]
        self.lend["base_interest_rate"][s] = e["baseInterestRate"]
        self.lend["loan_duration"][s] = e["loanDuration"]

    def adopt_new_agents(self, new_agents, c):
        """Agents the HOST just placed (exact-mode replacement, init.Population.place + HostState.append_agents):
        give them fresh oracle-private state from their endowments, like the reference's configureAgents."""
        self.L.orc_private_reset.argtypes = [C.c_int32]
        self.L.orc_private_reset.restype = None
        g, a = self.state.as_structs()
        for agent in new_agents:
            s, e = int(self.state.occ[agent["cell"]]), agent["endowment"]
            self.L.orc_private_reset(s)
            if getattr(self, "lend", None) is not None:
                self._set_lending(s, e)
            if getattr(self, "social", None) is not None:
                self.social["max_friends"][s] = e["maxFriends"]
            if getattr(self, "eth", None) is not None:
                self._set_ethics(s, e, c)
            if getattr(self, "dis", None) is not None:
                self._set_disease(s, e)
            if getattr(self, "misc", None) is not None:
                self._set_misc(s, e, a)

    def unbind_lending(self):
        self.L.orc_bind_lending(*([None] * 7), 0)

    def bind_disease(self, c, endowments, disease_endowments, horizon):
        """Oracle-private disease state + configureDiseases' initial infections (reference
        sugarscape.py:306-337), driven from here with the stdlib main stream like the rest of the
        init path.  Call before set_mt_from_python()."""
        from sugarscape_b200 import steptables
        cap, n_d = self.state.capacity, len(disease_endowments)
        length = int(c["agentImmuneSystemLength"])
        assert 0 < length <= 64

        class Disease(C.Structure):
            _fields_ = [("id", C.c_int32), ("incubation", C.c_int32), ("start_timestep", C.c_int32), ("tag_len", C.c_int32),
                        ("tags", C.c_uint32), ("pad", C.c_uint32), ("transmission", C.c_double), ("penalty", C.c_double * 8),
                        ("infected", C.c_int64), ("listed", C.c_int64)]
        L = self.L
        L.orc_sizeof_disease.restype = C.c_int
        assert L.orc_sizeof_disease() == C.sizeof(Disease)
        self.disease_table = (Disease * max(n_d, 1))()
        order = ("aggressionPenalty", "fertilityPenalty", "friendlinessPenalty", "happinessPenalty", "movementPenalty",
                 "spiceMetabolismPenalty", "sugarMetabolismPenalty", "visionPenalty")
        for i, e in enumerate(disease_endowments):
            d = self.disease_table[i]
            assert len(e["tags"]) <= 32
            d.id, d.incubation, d.start_timestep, d.tag_len = i, int(e["incubationPeriod"]), int(e["startTimestep"]), len(e["tags"])
            d.tags = sum(b << k for k, b in enumerate(e["tags"]))
            d.transmission = float(e["transmissionChance"])
            for k, name in enumerate(order):
                d.penalty[k] = float(e[name])
        self.dis = {"immune": np.zeros(cap, np.uint64), "start_immune": np.zeros(cap, np.uint64),
                    "protection": np.zeros(cap, np.float64), "modifier": np.zeros(cap * 8, np.float64),
                    "disease_death": np.zeros(cap, np.int32), "last_spread": np.full(cap, -1, np.int32),
                    "n_spread": np.zeros(cap, np.int32), "health_happiness": np.zeros(cap, np.float64),
                    "last_valid_moves": np.zeros(cap, np.int32),
                    "immune_bits": steptables.immune_tables(1, horizon, length)}
        n = int(self.state.counters[layout.C_N_LIST])
        agents = [int(s) for s in self.state.a_order[:n]]
        for s in agents:
            self._set_disease(s, endowments[int(self.state.a_id[s])])
        L.orc_bind_disease.argtypes = [C.c_void_p, C.c_int32, C.c_int32] + [C.c_void_p] * 10 + [C.c_int64]
        L.orc_move_space.argtypes = [C.c_void_p]
        L.orc_move_space.restype = C.c_int64
        L.orc_health_happiness_sum.argtypes = [C.c_void_p]
        L.orc_health_happiness_sum.restype = C.c_double
        L.orc_bind_disease.restype = None
        L.orc_disease_count.argtypes = [C.c_int32]
        L.orc_disease_count.restype = C.c_int32
        L.orc_disease_nearest.argtypes = [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
        L.orc_disease_nearest.restype = C.c_int32
        L.orc_disease_catch.argtypes = [C.c_int32] * 6 + [C.c_void_p]
        L.orc_disease_catch.restype = None
        L.orc_disease_listed.argtypes = [C.c_int32]
        L.orc_disease_listed.restype = None
        L.orc_disease_stats.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_int64 * 5)]
        L.orc_disease_stats.restype = None
        L.orc_bind_disease(C.cast(self.disease_table, C.c_void_p), n_d, length,
                           *[self.dis[k].ctypes.data for k in ("immune", "start_immune", "protection", "modifier",
                                                               "disease_death", "last_spread", "n_spread", "health_happiness",
                                                               "last_valid_moves", "immune_bits")], cap)
        # configureDiseases, after the agent shuffle (sugarscape.py:317-337)
        initial = [i for i in range(n_d) if int(disease_endowments[i]["startTimestep"]) == 0]      # (the others wait: remainingDiseases)
        lo, hi = c["startingDiseasesPerAgent"]
        unlimited = [lo, hi] == [0, 0]
        curr = lo
        for s in agents:
            random.shuffle(initial)
            for d in initial:
                if L.orc_disease_count(s) >= curr and not unlimited:
                    curr += 1
                    break
                if L.orc_disease_nearest(s, d, None, None) == 0:
                    continue
                L.orc_disease_catch(s, d, -1, -1, 0, 1, None)
                L.orc_disease_listed(d)
                if unlimited:
                    initial.remove(d)
                    break
            if curr > hi:
                curr = lo

    def _set_disease(self, s, e):
        bits = sum(b << k for k, b in enumerate(e["immuneSystem"]))
        self.dis["immune"][s] = self.dis["start_immune"][s] = bits
        self.dis["protection"][s] = e["diseaseProtectionChance"]

    def unbind_disease(self):
        self.L.orc_bind_disease(None, 0, 0, *([None] * 10), 0)

    def disease_stats(self, t):
        """sickAgents, diseaseIncidence, distinct infectors, diseasePrevalence, agentDiseaseDeaths."""
        g, a = self.state.as_structs()
        out = (C.c_int64 * 5)()
        self.L.orc_disease_stats(C.byref(a), t, C.byref(out))
        return list(out)

    def disease_log_keys(self, t, population):
        """The six disease keys of the log + meanHealthHappiness (sugarscape.py:1316-1317, 1147-1160)."""
        sick, incidence, infectors, prevalence, deaths = self.disease_stats(t)
        g, a = self.state.as_structs()
        health = self.L.orc_health_happiness_sum(C.byref(a))
        n = population
        return {"sickAgents": sick, "sickAgentsPercentage": round((sick / n) * 100, 2) if n else 0,
                "diseaseEffectiveReproductionRate": (round(incidence / infectors, 2) if infectors else 0) if n else 0,
                "diseaseIncidence": incidence, "diseasePrevalence": prevalence, "agentDiseaseDeaths": deaths,
                "meanHealthHappiness": round(health / n, 2) if n else 0,
                "moveSpace": int(self.L.orc_move_space(C.byref(a)))}

    def bind_misc(self, c, endowments, horizon):
        """Oracle-private universal income, racial tags and depression.  Call after bind_social (a
        depressed agent's maxFriends shrinks) and before stepping."""
        from sugarscape_b200 import steptables
        cap = self.state.capacity
        rl = int(c["agentRacialTagStringLength"]) if c["environmentMaxRaces"] > 0 else 0
        assert rl <= 16 and c["environmentMaxRaces"] <= 4
        picks, draw = steptables.racial_tables(1, horizon, rl)
        self.misc = {"universal_sugar": np.zeros(cap), "universal_spice": np.zeros(cap),
                     "last_income_sugar": np.zeros(cap, np.int32), "last_income_spice": np.zeros(cap, np.int32),
                     "racial_tags": np.zeros(cap, np.uint32), "race": np.full(cap, -1, np.int32),
                     "depressed": np.zeros(cap, np.int32), "happiness_unit": np.ones(cap), "picks": picks, "draw": draw}

        class Misc(C.Structure):
            _fields_ = [("capacity", C.c_int64), ("universal_sugar", C.c_void_p), ("universal_spice", C.c_void_p),
                        ("last_income_sugar", C.c_void_p), ("last_income_spice", C.c_void_p),
                        ("income_interval_sugar", C.c_int32), ("income_interval_spice", C.c_int32),
                        ("racial_tags", C.c_void_p), ("race", C.c_void_p), ("racial_len", C.c_int32), ("n_races", C.c_int32),
                        ("depressed", C.c_void_p), ("happiness_unit", C.c_void_p), ("depression_percentage", C.c_double),
                        ("racial_picks", C.c_void_p), ("depression_draw", C.c_void_p)]
        L = self.L
        L.orc_sizeof_misc.restype = C.c_int
        assert L.orc_sizeof_misc() == C.sizeof(Misc)
        m = Misc()
        m.capacity = cap
        for k in ("universal_sugar", "universal_spice", "last_income_sugar", "last_income_spice", "racial_tags", "race",
                  "depressed", "happiness_unit"):
            setattr(m, k, self.misc[k].ctypes.data)
        m.income_interval_sugar = int(c["environmentUniversalSugarIncomeInterval"])
        m.income_interval_spice = int(c["environmentUniversalSpiceIncomeInterval"])
        m.racial_len, m.n_races = rl, int(c["environmentMaxRaces"])
        m.depression_percentage = float(c["agentDepressionPercentage"])
        m.racial_picks, m.depression_draw = picks.ctypes.data, draw.ctypes.data
        L.orc_bind_misc.argtypes = [C.c_void_p]
        L.orc_bind_misc.restype = None
        L.orc_depress.argtypes = [C.c_void_p, C.c_int32]
        L.orc_depress.restype = None
        L.orc_race_census.argtypes = [C.c_void_p, C.POINTER(C.c_int64 * 3)]
        L.orc_race_census.restype = None
        L.orc_bind_misc(C.byref(m))
        g, a = self.state.as_structs()
        n = int(self.state.counters[layout.C_N_LIST])
        for s in self.state.a_order[:n]:
            self._set_misc(int(s), endowments[int(self.state.a_id[s])], a)

    def _set_misc(self, s, e, a):
        self.misc["universal_sugar"][s] = e["universalSugar"]
        self.misc["universal_spice"][s] = e["universalSpice"]
        if e["racialTags"] is not None:
            rt = e["racialTags"]
            self.misc["racial_tags"][s] = sum(v << (2 * k) for k, v in enumerate(rt))
            self.misc["race"][s] = max(set(rt), key=rt.count)
        if e["depressionFactor"] == 1:
            self.L.orc_depress(C.byref(a), s)

    def unbind_misc(self):
        self.L.orc_bind_misc(None)

    def race_census(self):
        """(largestRace or None, largestRaceSize, remainingRaces)."""
        g, a = self.state.as_structs()
        out = (C.c_int64 * 3)()
        self.L.orc_race_census(C.byref(a), C.byref(out))
        return (None if out[0] < 0 else int(out[0])), int(out[1]), int(out[2])

    def bind_ethics(self, c, endowments):
        """Oracle-private ethics.Bentham decision models (spawn rules: reference sugarscape.py:219-233)."""
        cap = self.state.capacity
        for k in ("agentDecisionModelAgeismFactor", "agentDecisionModelRacismFactor", "agentDecisionModelSexismFactor",
                  "agentDecisionModelTribalFactor"):
            assert c[k][1] < 0, "in-group factors are not restated"
        assert c["agentDynamicSelfishnessFactor"][1] == 0
        self.eth = {"model": np.zeros(cap, np.int32), "selfishness": np.zeros(cap), "decision_factor": np.zeros(cap),
                    "lookahead_factor": np.zeros(cap), "lookahead_discount": np.zeros(cap),
                    "last_move_optimal": np.ones(cap, np.int32), "move_rank": np.zeros(cap, np.int32), "move_diff": np.zeros(cap)}
        n = int(self.state.counters[layout.C_N_LIST])
        for s in self.state.a_order[:n]:
            self._set_ethics(int(s), endowments[int(self.state.a_id[s])], c)

        class Ethics(C.Structure):
            _fields_ = [("capacity", C.c_int64), ("model", C.c_void_p), ("selfishness", C.c_void_p), ("decision_factor", C.c_void_p),
                        ("lookahead_factor", C.c_void_p), ("lookahead_discount", C.c_void_p), ("last_move_optimal", C.c_void_p),
                        ("move_rank", C.c_void_p), ("move_diff", C.c_void_p), ("global_max_wealth", C.c_double)]
        L = self.L
        L.orc_sizeof_ethics.restype = C.c_int
        assert L.orc_sizeof_ethics() == C.sizeof(Ethics)
        e = Ethics()
        e.capacity = cap
        for k in ("model", "selfishness", "decision_factor", "lookahead_factor", "lookahead_discount", "last_move_optimal",
                  "move_rank", "move_diff"):
            setattr(e, k, self.eth[k].ctypes.data)
        e.global_max_wealth = float(c["environmentMaxSugar"] + c["environmentMaxSpice"])
        L.orc_bind_ethics.argtypes = [C.c_void_p]
        L.orc_bind_ethics.restype = None
        L.orc_ethics_stats.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_double * 5)]
        L.orc_ethics_stats.restype = None
        L.orc_bind_ethics(C.byref(e))

    def _set_ethics(self, s, e, c):
        name = e["decisionModel"]
        selfish = e["selfishnessFactor"]
        bentham = True
        if "altruist" in name:
            selfish = 0
        elif "bentham" in name:
            selfish = 0.5 if selfish < 0 else selfish
        elif "egoist" in name:
            selfish = 1
        elif "negativeBentham" in name:
            selfish = -1
        else:
            assert name == "none", name
            bentham = False
        look = c["agentDecisionModelLookaheadFactor"]
        if isinstance(look, list):      # the default is the list [0]: handed to the agents as is, and `[0] != 0` holds
            look = 1.0
        if "NoLookahead" in name:
            look = 0
        elif "HalfLookahead" in name:
            look = 0.5
        self.eth["model"][s] = int(bentham)
        self.eth["selfishness"][s] = selfish
        self.eth["decision_factor"][s] = e["decisionModelFactor"]
        self.eth["lookahead_factor"][s] = look
        self.eth["lookahead_discount"][s] = e["decisionModelLookaheadDiscount"]

    def unbind_ethics(self):
        self.L.orc_bind_ethics(None)

    def ethics_log_keys(self, t, population, mean_neighbors):
        """meanSelfishness, agentLastMoveOptimalityPercentage, meanMoveRank, meanMoveDifferenceFromOptimal
        (reference sugarscape.py:1304, 1318, 1323-1324)."""
        g, a = self.state.as_structs()
        out = (C.c_double * 5)()
        self.L.orc_ethics_stats(C.byref(a), t, C.byref(out))
        n = population
        if n == 0:      # the reference leaves the raw count of optimal moves (of the dead) in place (sugarscape.py:1325-1351)
            return {"meanSelfishness": 0, "agentLastMoveOptimalityPercentage": int(out[1]), "meanMoveRank": 0,
                    "meanMoveDifferenceFromOptimal": 0}
        return {"meanSelfishness": round(out[0] / n, 2),
                "agentLastMoveOptimalityPercentage": round((out[1] / out[2]) * 100, 2),
                "meanMoveRank": round(out[3] / n, 2),
                "meanMoveDifferenceFromOptimal": round(out[4] / mean_neighbors, 2) if mean_neighbors > 0 else 0}

    def ethics_stats(self, t):
        """The oracle's decision-model reductions as a layout.EthicsStats (what StatsBuilder.update takes)."""
        g, a = self.state.as_structs()
        out = (C.c_double * 5)()
        self.L.orc_ethics_stats(C.byref(a), t, C.byref(out))
        es = layout.EthicsStats()
        es.sum_selfishness, es.optimal_moves, es.moves, es.sum_move_rank, es.sum_move_diff = (
            out[0], int(out[1]), int(out[2]), int(out[3]), out[4])
        return es

    def bind_group(self):
        """Oracle-private bookkeeping of the experimental-group split of the log (group "depressed")."""
        cap = self.state.capacity
        self.grp = {"with_control": np.zeros(cap * 5, np.int32), "with_experimental": np.zeros(cap * 5, np.int32),
                    "control_neighbors": np.zeros(cap, np.int32), "experimental_neighbors": np.zeros(cap, np.int32)}

        class Group(C.Structure):
            _fields_ = [("capacity", C.c_int64), ("with_control", C.c_void_p), ("with_experimental", C.c_void_p),
                        ("control_neighbors", C.c_void_p), ("experimental_neighbors", C.c_void_p)]
        L = self.L
        L.orc_sizeof_group.restype = C.c_int
        assert L.orc_sizeof_group() == C.sizeof(Group)
        g = Group()
        g.capacity = cap
        for k in ("with_control", "with_experimental", "control_neighbors", "experimental_neighbors"):
            setattr(g, k, self.grp[k].ctypes.data)
        L.orc_bind_group.argtypes = [C.c_void_p]
        L.orc_bind_group.restype = None
        L.orc_set_pass.argtypes = [C.c_int32]
        L.orc_set_pass.restype = None
        L.orc_group_counters.argtypes = [C.c_void_p, C.POINTER(C.c_int64 * 12)]
        L.orc_group_counters.restype = None
        L.orc_bind_group(C.byref(g))

    def unbind_group(self):
        self.L.orc_bind_group(None)

    def set_pass(self, which):
        """0 = statistics over everybody, 1 = the experimental group, 2 = the control group."""
        self.L.orc_set_pass(which)

    def group_counters(self):
        g, a = self.state.as_structs()
        out = (C.c_int64 * 12)()
        self.L.orc_group_counters(C.byref(a), C.byref(out))
        return list(out)

    def bind_social(self, endowments):
        """Oracle-private friends lists + conflict happiness (log keys only)."""
        cap = self.state.capacity
        self.social = {"max_friends": np.zeros(cap, np.int32), "last_combat": np.full(cap, -1, np.int32),
                       "social_happiness": np.zeros(cap, np.float64), "conflict_happiness": np.zeros(cap, np.float64)}
        n = int(self.state.counters[layout.C_N_LIST])
        for s in self.state.a_order[:n]:
            self.social["max_friends"][s] = endowments[int(self.state.a_id[s])]["maxFriends"]
        L = self.L
        L.orc_bind_social.argtypes = [C.c_void_p] * 4 + [C.c_int64]
        L.orc_bind_social.restype = None
        L.orc_social_sums.argtypes = [C.c_void_p, C.POINTER(C.c_double * 2)]
        L.orc_social_sums.restype = None
        L.orc_bind_social(*[self.social[k].ctypes.data for k in ("max_friends", "last_combat", "social_happiness",
                                                                 "conflict_happiness")], cap)

    def unbind_social(self):
        self.L.orc_bind_social(*([None] * 4), 0)

    def social_sums(self):
        """(sum of socialHappiness, sum of conflictHappiness) over the agent list."""
        g, a = self.state.as_structs()
        out = (C.c_double * 2)()
        self.L.orc_social_sums(C.byref(a), C.byref(out))
        return out[0], out[1]

    def loan_volume(self, t):
        g, a = self.state.as_structs()
        return int(self.L.orc_loan_volume(C.byref(a), t))

    def replace(self, t):
        g, a = self.state.as_structs()
        self.L.orc_replace(C.byref(self.params), C.byref(g), C.byref(a), C.byref(self.endow), t)

    def stats(self, t):
        g, a = self.state.as_structs()
        out = layout.Stats()
        self.L.orc_stats(C.byref(self.params), C.byref(g), C.byref(a), t, C.byref(out))
        return out
