"""Struct-of-arrays layout shared by the host code and the C ABI (mirror of include/ssb.h).

One table (``AGENT_FIELDS`` / ``GRID_FIELDS``) drives the numpy host arrays, the torch device
tensors and the ctypes structs, so the three cannot drift apart.  tests/test_abi.py checks the
ctypes struct sizes against ``sizeof`` values exported by the built library.
"""
import ctypes as C

import numpy as np

ABI_VERSION = 4
RNG_EXACT, RNG_SCALABLE = 0, 1
ALIVE, STARVATION, AGING, COMBAT, OTHER = 0, 1, 2, 3, 4
SEX_NONE, SEX_MALE, SEX_FEMALE = 0, 1, 2
SEX_CODE = {None: SEX_NONE, "male": SEX_MALE, "female": SEX_FEMALE}

F_SPICE, F_POLLUTION, F_REPRO, F_TAGS, F_TAGGING, F_COMBAT, F_TRADE, F_TAGPREF, F_NOWRAP, F_MOORE, F_RADIAL = (
    1 << 0, 1 << 1, 1 << 2, 1 << 3, 1 << 4, 1 << 5, 1 << 6, 1 << 7, 1 << 8, 1 << 9, 1 << 10)

(C_N_LIST, C_N_SLOTS, C_NEXT_ID, C_N_FREE, C_N_BORN, C_N_DEAD, C_ROUNDS, C_OVERFLOW, C_N_KIN,
 C_ENDOW_INDEX, C_N_REPLACED, C_N_MIGRATED, C_KIN_FREE) = range(13)
INHERITANCE_CODE = {"none": 0, "children": 1, "sons": 2, "daughters": 3}
C_COUNT = 16

# (name, numpy dtype) in the order of ssb_grid_t
GRID_FIELDS = (
    ("sugar", np.int32), ("sugar_cap", np.int32), ("spice", np.int32), ("spice_cap", np.int32),
    ("pollution", np.float64), ("pollution_flux", np.float64), ("occ", np.int32),
)
# (name, numpy dtype, fill) in the order of ssb_agents_t after `counters`
AGENT_FIELDS = (
    ("order", np.int32, -1), ("free_slots", np.int32, -1), ("dead_list", np.int32, -1),
    ("id", np.int32, -1), ("pos", np.int32, -1), ("born", np.int32, 0), ("cod", np.int32, 0),
    ("sugar", np.float64, 0), ("spice", np.float64, 0),
    ("sugar_metab", np.int32, 0), ("spice_metab", np.int32, 0),
    ("vision", np.int32, 0), ("movement", np.int32, 0),
    ("age", np.int32, 0), ("max_age", np.int32, -1),
    ("sex", np.int32, 0),
    ("fert_age", np.int32, 0), ("infert_age", np.int32, 0),
    ("fert_factor", np.float64, 0),
    ("start_sugar", np.float64, 0), ("start_spice", np.float64, 0),
    ("aggression", np.float64, 0), ("lookahead", np.float64, 0),
    ("trade_factor", np.float64, 0), ("mrs", np.float64, 1),
    ("tags", np.uint32, 0), ("tribe", np.int32, -1),
    ("last_moved", np.int32, -1),
    ("last_sugar", np.float64, 0), ("last_spice", np.float64, 0),
    ("n_neighborhood", np.int32, 0), ("n_valid_moves", np.int32, 0),
    ("happiness", np.float64, 0), ("wealth_happiness", np.float64, 0),
    ("family_happiness", np.float64, 0),
    ("trade_volume", np.int32, 0), ("trade_price", np.float64, 0),
    ("last_repro", np.int32, -1),
    ("children_head", np.int32, -1), ("mates_head", np.int32, -1),
    ("father", np.int32, -1), ("mother", np.int32, -1),
)

_CT = {np.int32: C.c_int32, np.uint32: C.c_uint32, np.float64: C.c_double, np.int64: C.c_int64}


class Params(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("width", C.c_int32), ("height", C.c_int32),
        ("rng_mode", C.c_int32), ("seed", C.c_uint64), ("flags", C.c_int32),
        ("max_agent_range", C.c_int32),
        ("sugar_regrow", C.c_int32), ("spice_regrow", C.c_int32),
        ("season_interval", C.c_int32), ("seasonal_growback_delay", C.c_int32),
        ("equator", C.c_int32),
        ("pollution_start", C.c_int32), ("pollution_end", C.c_int32),
        ("diffusion_start", C.c_int32), ("diffusion_end", C.c_int32),
        ("diffusion_delay", C.c_int32),
        ("sugar_prod_pollution", C.c_double), ("spice_prod_pollution", C.c_double),
        ("sugar_cons_pollution", C.c_double), ("spice_cons_pollution", C.c_double),
        ("max_combat_loot", C.c_double),
        ("tag_len", C.c_int32), ("max_tribes", C.c_int32),
        ("max_timestep", C.c_int32), ("id_bits", C.c_int32),
        ("max_sugar_capacity", C.c_int32), ("max_spice_capacity", C.c_int32),
        ("inheritance_policy", C.c_int32), ("agent_replacements", C.c_int32),
        ("quad_w", C.c_int32), ("quad_h", C.c_int32), ("quad_mask", C.c_int32),
        ("tribe_per_quadrant", C.c_int32), ("_pad", C.c_int32),
    ]


class Grid(C.Structure):
    _fields_ = [("n_cells", C.c_int64)] + [(n, C.c_void_p) for n, _ in GRID_FIELDS]


KIN_FIELDS = ("kin_slot", "kin_id", "kin_next")

# (name, numpy dtype) in the order of ssb_endowments_t after `count`
ENDOWMENT_FIELDS = (
    ("sugar", np.float64), ("spice", np.float64),
    ("sugar_metab", np.int32), ("spice_metab", np.int32), ("vision", np.int32), ("movement", np.int32),
    ("max_age", np.int32), ("sex", np.int32), ("fert_age", np.int32), ("infert_age", np.int32),
    ("fert_factor", np.float64), ("aggression", np.float64), ("lookahead", np.float64),
    ("trade_factor", np.float64), ("tags", np.uint32),
)


class Endowments(C.Structure):
    _fields_ = [("count", C.c_int64)] + [(n, C.c_void_p) for n, _ in ENDOWMENT_FIELDS]


def endowment_arrays(endowments):
    """The endowment dicts of init.randomize_endowments as struct-of-arrays (creation order)."""
    n = len(endowments)
    out = {name: np.zeros(n, dtype=dt) for name, dt in ENDOWMENT_FIELDS}
    keys = {"sugar": "sugar", "spice": "spice", "sugar_metab": "sugarMetabolism", "spice_metab": "spiceMetabolism",
            "vision": "vision", "movement": "movement", "max_age": "maxAge", "fert_age": "fertilityAge",
            "infert_age": "infertilityAge", "fert_factor": "fertilityFactor", "aggression": "aggressionFactor",
            "lookahead": "lookaheadFactor", "trade_factor": "tradeFactor"}
    for i, e in enumerate(endowments):
        for name, key in keys.items():
            out[name][i] = e[key]
        out["sex"][i] = SEX_CODE[e["sex"]]
        out["tags"][i] = sum((1 << k) for k, b in enumerate(e["tags"] or ()) if b)
    return out


def endowments_struct(arrays, pointer_of):
    e = Endowments()
    e.count = len(arrays["sugar"])
    for name, _ in ENDOWMENT_FIELDS:
        setattr(e, name, pointer_of(arrays[name]))
    return e


def default_mark_shift(params):
    """The reservation-mark coarsening ssb_bind picks (csrc/ssb_abi.cu): 2 x 2 cell blocks when the
    map allows it, 4 x 4 on large maps for rule sets whose footprint is the bare cross."""
    W, H = params.width, params.height
    maxd = max(min(params.max_agent_range, W // 2), min(params.max_agent_range, H // 2))
    shift = 0
    if W % 2 == 0 and H % 2 == 0 and ((maxd + 2) >> 1) + 2 < (W >> 1) and ((maxd + 2) >> 1) + 2 < (H >> 1):
        shift = 1
    if shift == 1 and not (params.flags & (F_TAGGING | F_TRADE | F_REPRO)) and W * H >= (1 << 24) and W % 4 == 0 and H % 4 == 0:
        shift = 2
    return shift


MAX_RANKS = 8


class Shard(C.Structure):       # ssb_shard_t
    _fields_ = [("rank", C.c_int32), ("world", C.c_int32),
                ("cell_end", C.c_int64 * MAX_RANKS),
                ("slot_lo", C.c_int64), ("slot_hi", C.c_int64), ("mark_lo", C.c_int64), ("mark_hi", C.c_int64),
                ("ctrl", C.c_void_p), ("ctrl_stride", C.c_uint64),
                ("marks", C.c_void_p),
                ("gather", C.c_void_p), ("gather_stride", C.c_uint64)]


class Agents(C.Structure):
    _fields_ = ([("capacity", C.c_int64), ("counters", C.c_void_p)]
                + [(n, C.c_void_p) for n, _, _ in AGENT_FIELDS]
                + [("kin_capacity", C.c_int64)] + [(n, C.c_void_p) for n in KIN_FIELDS])


class Stats(C.Structure):
    _fields_ = [
        ("timestep", C.c_int64), ("population", C.c_int64),
        ("sum_sugar_metab", C.c_int64), ("sum_spice_metab", C.c_int64),
        ("sum_movement", C.c_int64), ("sum_vision", C.c_int64), ("sum_age", C.c_int64),
        ("sum_wealth", C.c_double), ("max_wealth", C.c_double), ("min_wealth", C.c_double),
        ("sum_wealth_collected", C.c_double), ("sum_burn_rate", C.c_double),
        ("sum_ttl_age_limited", C.c_double),
        ("sum_neighbors", C.c_int64), ("sum_valid_moves", C.c_int64),
        ("n_traders", C.c_int64), ("trade_volume", C.c_int64), ("sum_trade_price", C.c_double),
        ("sum_happiness", C.c_double), ("sum_wealth_happiness", C.c_double),
        ("sum_family_happiness", C.c_double),
        ("tribe_count", C.c_int64 * 32), ("tribe_first", C.c_int64 * 32),
        ("n_dead", C.c_int64), ("n_dead_moved", C.c_int64), ("dead_starvation", C.c_int64),
        ("dead_aging", C.c_int64), ("dead_combat", C.c_int64), ("sum_age_at_death", C.c_int64),
        ("dead_wealth_collected", C.c_double),
        ("n_born", C.c_int64), ("n_replaced", C.c_int64), ("n_acted", C.c_int64),
        ("env_wealth_total", C.c_int64), ("env_wealth_created", C.c_int64),
        ("env_capacity_total", C.c_int64),
        ("lorenz_weighted_sum", C.c_double), ("rounds", C.c_int64), ("gini_area", C.c_double),
    ]


# ---- ABI 2: decision models (ssb_ethics_t; reference ethics.py, spawn rules sugarscape.py:219-233) ----
# (name, numpy dtype, fill) in the order of ssb_ethics_t after `capacity`
ETHICS_FIELDS = (
    ("model", np.int32, 0), ("top_ordering", np.int32, 0), ("selfishness", np.float64, 0), ("decision_factor", np.float64, 0),
    ("lookahead_factor", np.float64, 0), ("lookahead_discount", np.float64, 0), ("dynamic_decision_factor", np.float64, 0),
    ("pecs_rules", np.int32, 0), ("pecs_counts", np.int32, 0), ("social_pressure", np.float64, 0),
    ("last_delta_ttl", np.float64, 0), ("dynamic_social_pressure", np.float64, 0), ("total_metabolism", np.float64, 0),
    ("dynamic_selfishness", np.float64, 0),
    ("last_move_optimal", np.int32, 1), ("move_rank", np.int32, 0), ("move_diff", np.float64, 0),
)
# after global_max_wealth: the stored neighbourhood lists (robots next to the Bentham family), [slot][hood_stride]
ETHICS_HOOD_FIELDS = (("hood_list", np.int32, -1), ("hood_id", np.int32, -1), ("hood_n", np.int32, -1))
BENTHAM_MODELS = ("altruist", "bentham", "egoist", "negativeBentham")


class Ethics(C.Structure):
    _fields_ = ([("capacity", C.c_int64)] + [(n, C.c_void_p) for n, _, _ in ETHICS_FIELDS]
                + [("global_max_wealth", C.c_double)] + [(n, C.c_void_p) for n, _, _ in ETHICS_HOOD_FIELDS]
                + [("hood_stride", C.c_int64)])


class EthicsStats(C.Structure):
    _fields_ = [("sum_selfishness", C.c_double), ("optimal_moves", C.c_int64), ("moves", C.c_int64),
                ("sum_move_rank", C.c_int64), ("sum_move_diff", C.c_double)]


def decision_model_class(name):
    """ssb_ethics_t.model of a decisionModel string: 1 = ethics.Bentham, 2 = ethics.Temperance (simple), 3 = with PECS,
    4 = ethics.Asimov (sugarscape.py:245-246 replaces whatever class the earlier tests chose), 0 = plain Agent."""
    if "asimov" in name:
        return 4
    if decision_model_base(name) is not None:
        return 1
    if "temperance" in name:
        return 3 if "PECS" in name else 2
    return 0


def decision_model_base(name):
    """The Bentham-family member a decisionModel string selects, in the reference's matching order
    (sugarscape.py:219-233: substring tests, "altruist" first), or None for a plain Agent."""
    for base in BENTHAM_MODELS:
        if base in name:
            return base
    return None


class HostEthics:
    """Per-slot decision-model state (host mirror of ssb_ethics_t)."""

    def __init__(self, capacity, c):
        self.capacity = capacity
        self.global_max_wealth = float(c["environmentMaxSugar"] + c["environmentMaxSpice"])
        look = c["agentDecisionModelLookaheadFactor"]
        # the reference's default is the LIST [0], handed to the agents as is; only `!= 0` is ever asked of
        # it (ethics.py:160-190) and `[0] != 0` holds
        self.lookahead_factor = 1.0 if isinstance(look, list) else float(look)
        self.default_dynamic = list(c["agentDynamicSelfishnessFactor"]) == [0.0, 0.0]
        self.group_model = c["experimentalGroup"] if group_kind(c) == GROUP_MODEL else None
        # robots next to the Bentham family: every agent's neighbourhood as a stored list (ssb_ethics_t.hood_*)
        models = c["agentDecisionModels"]
        self.hood_stride = 161 if ("asimov" in models and any(decision_model_class(m) == 1 for m in models)) else 0      # SSB_GROUP_HOOD_MAX
        widths = {"pecs_rules": 5, "pecs_counts": 3, "hood_list": max(self.hood_stride, 1), "hood_id": max(self.hood_stride, 1)}   # values per slot
        self.arrays = {n: np.full(capacity * widths.get(n, 1), fill, dtype=dt) for n, dt, fill in ETHICS_FIELDS + ETHICS_HOOD_FIELDS}

    def set_slot(self, slot, endowment):
        """A freshly configured agent (sugarscape.py:219-233)."""
        name = endowment["decisionModel"]
        base = decision_model_base(name)
        selfish = endowment["selfishnessFactor"]
        if base == "altruist":
            selfish = 0
        elif base == "bentham":
            selfish = 0.5 if selfish < 0 else selfish
        elif base == "egoist":
            selfish = 1
        elif base == "negativeBentham":
            selfish = -1
        look = self.lookahead_factor
        if "NoLookahead" in name:
            look = 0
        elif "HalfLookahead" in name:
            look = 0.5
        A = self.arrays
        A["model"][slot] = decision_model_class(name)
        # bit 1: the agent's decisionModel is the experimental group (agent.py:1262; SSB_GROUP_MODEL)
        A["top_ordering"][slot] = int(base is not None and "Top" in name) | (2 if name == self.group_model else 0)
        A["dynamic_decision_factor"][slot] = endowment["dynamicDecisionModelFactor"]
        A["dynamic_social_pressure"][slot] = endowment["dynamicSocialPressureFactor"]
        A["pecs_rules"][slot * 5:slot * 5 + 5] = 0
        A["pecs_counts"][slot * 3:slot * 3 + 3] = 0
        A["social_pressure"][slot] = 0
        A["last_delta_ttl"][slot] = 0
        A["selfishness"][slot] = selfish
        A["decision_factor"][slot] = endowment["decisionModelFactor"]
        A["lookahead_factor"][slot] = look
        A["lookahead_discount"][slot] = endowment["decisionModelLookaheadDiscount"]
        # "Dynamic" without a configured factor: a small degree of dynamic selfishness (sugarscape.py:236-238)
        A["dynamic_selfishness"][slot] = 0.01 if ("Dynamic" in name and self.default_dynamic) else endowment["dynamicSelfishnessFactor"]
        A["last_move_optimal"][slot] = 1
        A["move_rank"][slot] = 0
        A["move_diff"][slot] = 0

    def as_struct(self, pointer_of=lambda arr: arr.ctypes.data):
        e = Ethics()
        e.capacity = self.capacity
        for n, _, _ in ETHICS_FIELDS + ETHICS_HOOD_FIELDS:
            setattr(e, n, pointer_of(self.arrays[n]))
        e.global_max_wealth = self.global_max_wealth
        e.hood_stride = self.hood_stride
        return e

    def copy(self):
        h = HostEthics.__new__(HostEthics)
        h.capacity, h.global_max_wealth, h.lookahead_factor = self.capacity, self.global_max_wealth, self.lookahead_factor
        h.default_dynamic, h.group_model, h.hood_stride = self.default_dynamic, self.group_model, self.hood_stride
        h.arrays = {k: v.copy() for k, v in self.arrays.items()}
        return h


# ---- ABI 2: optional rule families (ssb_ext_t): lending, friends ----
_I32, _F64, _I64 = np.int32, np.float64, np.int64
LC_POOL_USED, LC_FREE_HEAD, LC_N_DEAD, LC_OVERFLOW, LC_COUNT = 0, 1, 2, 3, 8
MAX_FRIENDS = 16
# (struct member, dtype, fill, length in units of: "cap" slots / "pool" loans / "dead" creditors / "friends" = cap * 16 / int)
LENDING_LAYOUT = (
    ("lending_factor", _F64, 0, "cap"), ("base_interest_rate", _F64, 0, "cap"), ("loan_duration", _I32, 0, "cap"),
    ("sugar_mean_income", _F64, 1, "cap"), ("spice_mean_income", _F64, 1, "cap"),
    ("last_lended", _I32, -1, "cap"), ("last_loans", _I32, 0, "cap"),
    ("creditors_head", _I32, -1, "cap"), ("debtors_head", _I32, -1, "cap"),
    ("pool_capacity", None, None, None),
    ("loan_other_slot", _I32, -1, "pool"), ("loan_other_id", _I32, -1, "pool"),
    ("loan_sugar", _F64, 0, "pool"), ("loan_spice", _F64, 0, "pool"),
    ("loan_duration_v", _I32, 0, "pool"), ("loan_origin", _I32, 0, "pool"), ("loan_next", _I32, -1, "pool"),
    ("dead_capacity", None, None, None), ("max_loan_duration", None, None, None),
    ("dead_slot", _I32, -1, "dead"), ("dead_id", _I32, -1, "dead"), ("dead_children_head", _I32, -1, "dead"),
    ("dead_time", _I32, 0, "dead"),
    ("counters", _I64, 0, LC_COUNT),
)
SOCIAL_LAYOUT = (
    ("max_friends", _I32, 0, "cap"), ("last_combat", _I32, -1, "cap"),
    ("social_happiness", _F64, 0, "cap"), ("conflict_happiness", _F64, 0, "cap"), ("n_friends", _I32, 0, "cap"),
    ("friend_slot", _I32, -1, "friends"), ("friend_id", _I32, -1, "friends"), ("friend_hd", _I32, 0, "friends"),
)
DM_AGGRESSION, DM_FERTILITY, DM_FRIENDLINESS, DM_HAPPINESS, DM_MOVEMENT, DM_SPICE_METAB, DM_SUGAR_METAB, DM_VISION, DM_COUNT = range(9)
DISEASE_PENALTIES = ("aggressionPenalty", "fertilityPenalty", "friendlinessPenalty", "happinessPenalty", "movementPenalty",
                     "spiceMetabolismPenalty", "sugarMetabolismPenalty", "visionPenalty")      # SSB_DM_* order
DISEASE_DEF = np.dtype([("id", "<i4"), ("incubation", "<i4"), ("start_timestep", "<i4"), ("tag_len", "<i4"), ("tags", "<u4"),
                        ("pad", "<u4"), ("transmission", "<f8"), ("penalty", "<f8", (DM_COUNT,)), ("infected", "<i8"),
                        ("listed", "<i8")])      # ssb_disease_def_t
_U64 = np.uint64
DISEASE_LAYOUT = (
    ("immune", _U64, 0, "cap"), ("start_immune", _U64, 0, "cap"), ("protection", _F64, 0, "cap"),
    ("modifier", _F64, 0, "mods"),
    ("disease_death", _I32, 0, "cap"), ("last_spread", _I32, -1, "cap"), ("n_spread", _I32, 0, "cap"),
    ("health_happiness", _F64, 0, "cap"), ("last_valid_moves", _I32, 0, "cap"), ("stat_mark", _I32, -1, "cap"),
    ("n_sick", _I32, 0, "cap"),
    ("sick_disease", _I32, -1, "sick"), ("sick_start", _I32, 0, "sick"), ("sick_end", _I32, 0, "sick"),
    ("sick_caught", _I32, 0, "sick"), ("sick_incubation", _I32, 0, "sick"), ("sick_infector_slot", _I32, -1, "sick"),
    ("sick_infector_id", _I32, -1, "sick"),
    ("defs", DISEASE_DEF, 0, "defs"), ("immune_bits", _U64, 0, "bits"),
)
MISC_LAYOUT = (
    ("universal_sugar", _F64, 0, "cap"), ("universal_spice", _F64, 0, "cap"),
    ("last_income_sugar", _I32, 0, "cap"), ("last_income_spice", _I32, 0, "cap"),
    ("racial_tags", np.uint32, 0, "cap"), ("race", _I32, -1, "cap"), ("depressed", _I32, 0, "cap"),
    ("happiness_unit", _F64, 1, "cap"), ("health_happiness", _F64, 0, "cap"),
    ("racial_picks", np.uint32, 0, "bits"), ("depression_draw", _F64, 1, "bits"),
)
GI_KINDS = 5
GROUP_LAYOUT = (
    ("with_control", _I32, 0, "gi"), ("with_experimental", _I32, 0, "gi"),
    ("control_neighbors", _I32, 0, "cap"), ("experimental_neighbors", _I32, 0, "cap"),
    ("kind", None, None, None),
    ("hood_list", _I32, -1, "hood"), ("hood_n", _I32, 0, "cap"),
)
GROUP_HOOD_MAX = 161        # SSB_GROUP_HOOD_MAX
GROUP_KINDS = {"depressed": 1, "female": 2, "male": 3}      # SSB_GROUP_* (memberships fixed at birth)
GROUP_RACE, GROUP_MODEL, GROUP_SICK = 4, 5, 6      # (from GROUP_SICK on: membership changes during the run)


def group_kind(c):
    """ssb_group_t.kind of configuration["experimentalGroup"] (0: none / not supported).  Agent.isInGroup (agent.py:1256-1284)
    tests, in this order: "ageRange" in the name, name == the agent's decisionModel, "depressed", "disease" in the name,
    "female", "male", "race" in the name (race<N>), ..."""
    import re
    g = c["experimentalGroup"]
    if g is None:
        return 0
    if "ageRange" in g:        # (isInGroup reads self.cell.environment: the reference raises on the first dead agent it asks)
        return 0
    models = ["none" if m == "rawSugarscape" else m for m in c["agentDecisionModels"]]      # sugarscape.py:720-721
    if g in models and not (g in GROUP_KINDS or "disease" in g or "race" in g):
        return GROUP_MODEL if g != "none" else 0      # ("none": no ssb_ethics_t to carry the bit when every agent is plain)
    if g in GROUP_KINDS:
        return GROUP_KINDS[g]
    if "disease" in g:
        return 0
    if g == "sick":
        return GROUP_SICK
    m = re.search(r"race(?P<ID>\d+)", g) if "race" in g else None
    if m is not None and c["environmentMaxRaces"] > 0 and c["agentRacialTagStringLength"] > 0:
        return GROUP_RACE | (int(m.group("ID")) << 8)
    return 0
AL_FIELDS = 26          # SSB_AL_* (include/ssb.h)
AGENTLOG_LAYOUT = (
    ("time_to_live", _F64, 0, "cap"), ("last_pollution", _F64, 0, "cap"), ("prey_wealth", _F64, 0, "cap"),
    ("last_combat", _I32, -1, "cap"), ("last_mates", _I32, 0, "cap"), ("valid_moves", _I32, 0, "cap"),
    ("move_rank", _I32, 0, "cap"), ("n_neighbors", _I32, 0, "cap"), ("neighbor_slot", _I32, -1, "nbr"),
    ("records", _F64, 0, "recs"), ("record_capacity", None, None, None), ("n_records", _I64, 0, 1),
    ("write_records", None, None, None),
)
EXT_FAMILIES = (("lending", LENDING_LAYOUT), ("social", SOCIAL_LAYOUT), ("diseases", DISEASE_LAYOUT), ("misc", MISC_LAYOUT),
                ("group", GROUP_LAYOUT), ("agentlog", AGENTLOG_LAYOUT))


def _family_struct(name, table):
    return type(name, (C.Structure,), {"_fields_": [(n, C.c_int64 if dt is None else C.c_void_p) for n, dt, _, _ in table]})


Lending = _family_struct("Lending", LENDING_LAYOUT)
Social = _family_struct("Social", SOCIAL_LAYOUT)


class Diseases(C.Structure):
    _fields_ = ([(n, C.c_void_p) for n, _, _, _ in DISEASE_LAYOUT]
                + [("max_sick", C.c_int32), ("immune_len", C.c_int32), ("n_diseases", C.c_int32), ("n_immune_bits", C.c_int32)])


class Misc(C.Structure):
    _fields_ = ([(n, C.c_void_p) for n, _, _, _ in MISC_LAYOUT]
                + [("depression_percentage", C.c_double), ("income_interval_sugar", C.c_int32), ("income_interval_spice", C.c_int32),
                   ("racial_len", C.c_int32), ("n_races", C.c_int32), ("n_tables", C.c_int32), ("_pad", C.c_int32)])


Group = _family_struct("Group", GROUP_LAYOUT)
AgentLog = _family_struct("AgentLog", AGENTLOG_LAYOUT)


class Ext(C.Structure):
    _fields_ = [("capacity", C.c_int64), ("lending", Lending), ("social", Social), ("diseases", Diseases), ("misc", Misc),
                ("group", Group), ("agentlog", AgentLog)]


class ExtStats(C.Structure):
    _fields_ = [("loan_volume", C.c_int64), ("sum_social_happiness", C.c_double), ("sum_conflict_happiness", C.c_double),
                ("sick_agents", C.c_int64), ("disease_incidence", C.c_int64), ("distinct_infectors", C.c_int64),
                ("disease_prevalence", C.c_int64), ("disease_deaths", C.c_int64), ("sum_health_happiness", C.c_double),
                ("move_space", C.c_int64), ("sum_burn_rate", C.c_double), ("sum_ttl_age_limited", C.c_double),
                ("race_count", C.c_int64 * 4), ("race_first", C.c_int64 * 4)]


class GroupStats(C.Structure):      # ssb_group_stats_t
    _fields_ = [("stats", Stats), ("ethics", EthicsStats), ("ext", ExtStats), ("to_control", C.c_int64 * GI_KINDS),
                ("to_experimental", C.c_int64 * GI_KINDS), ("control_neighbors", C.c_int64), ("experimental_neighbors", C.c_int64)]


class HostExt:
    """Host mirror of ssb_ext_t: the per-slot state of the optional rule families that are live in this
    configuration.  `arrays["<family>.<member>"]` are numpy arrays; a family that is off has none."""

    def __init__(self, capacity, c, n_diseases=0, horizon=0):
        self.capacity = capacity
        self.families = set()
        if c["agentLendingFactor"][1] > 0:
            self.families.add("lending")
        # friends; also the home of lastCombatTimestep: conflict happiness (agent.py:912-918) exists wherever agents fight
        if c["agentMaxFriends"][1] > 0 or self.combat_possible(c):
            self.families.add("social")
        if self.uses_diseases(c):
            self.families.add("diseases")
        if self.uses_misc(c):
            self.families.add("misc")
        if c["experimentalGroup"] is not None:
            self.families.add("group")
        if c["agentLogfile"] is not None or self.uses_dynamic_selfishness(c):
            self.families.add("agentlog")         # (agent.timeToLive is tracked there: the dynamic selfishness update reads it)
        self.write_records = int(c["agentLogfile"] is not None)
        self.max_loan_duration = int(max(c["agentLoanDuration"]))
        self.group_kind = group_kind(c)
        self.racial_len = int(c["agentRacialTagStringLength"]) if c["environmentMaxRaces"] > 0 else 0
        self.misc_scalars = (float(c["agentDepressionPercentage"]), int(c["environmentUniversalSugarIncomeInterval"]),
                             int(c["environmentUniversalSpiceIncomeInterval"]), self.racial_len, int(c["environmentMaxRaces"]))
        self.n_diseases, self.max_sick = n_diseases, max(n_diseases, 1)
        self.immune_len = int(c["agentImmuneSystemLength"])
        self.sizes = {"cap": capacity, "pool": 8 * capacity, "dead": 4 * capacity, "friends": capacity * MAX_FRIENDS,
                      "hood": capacity * GROUP_HOOD_MAX if (self.group_kind & 0xff) >= GROUP_SICK else 1, "nbr": capacity * 4, "recs": capacity * AL_FIELDS, "record": capacity, "gi": capacity * GI_KINDS, "mods": capacity * DM_COUNT, "sick": capacity * self.max_sick, "defs": self.max_sick, "bits": horizon + 1}
        self.arrays = {}
        for fam, table in EXT_FAMILIES:
            if fam not in self.families:
                continue
            for n, dt, fill, length in table:
                if dt is not None:
                    count = self.sizes.get(length, length)
                    self.arrays[fam + "." + n] = np.zeros(count, dtype=dt) if fill == 0 else np.full(count, fill, dtype=dt)

    @staticmethod
    def uses_dynamic_selfishness(c):
        """Rules that read agent.timeToLive (tracked in the agentlog family): dynamic selfishness, Temperance."""
        bentham = [m for m in c["agentDecisionModels"] if decision_model_base(m) is not None]
        return ((bool(bentham) and (c["agentDynamicSelfishnessFactor"][1] != 0 or any("Dynamic" in m for m in bentham))) or
                any(decision_model_class(m) in (2, 3) for m in c["agentDecisionModels"]))

    @staticmethod
    def combat_possible(c):
        """Some agent may have findAggression() > 0: by endowment, or through a disease's aggression penalty."""
        return c["agentAggressionFactor"][1] > 0 or (HostExt.uses_diseases(c) and c["diseaseAggressionPenalty"][1] > 0)

    @staticmethod
    def uses_diseases(c):
        return c["startingDiseases"] > 0 and c["agentImmuneSystemLength"] > 0

    @staticmethod
    def uses_misc(c):
        return (c["agentUniversalSugar"][1] != 0 or c["agentUniversalSpice"][1] != 0 or c["agentDepressionPercentage"] > 0 or
                (c["agentRacialTagStringLength"] > 0 and c["environmentMaxRaces"] > 0) or c["experimentalGroup"] is not None)

    @staticmethod
    def needed(c, exact=True):
        """exact=False: what the scalable RNG mode would have to refuse (it computes no conflict happiness; the other
        families do not exist there)."""
        return (c["agentLendingFactor"][1] > 0 or c["agentMaxFriends"][1] > 0 or HostExt.uses_diseases(c) or
                HostExt.uses_misc(c) or c["agentLogfile"] is not None or HostExt.uses_dynamic_selfishness(c) or
                (exact and HostExt.combat_possible(c)))

    def _free_chain(self, head):
        """Give the loan records of a list back to the pool (a new agent object starts with empty lists)."""
        A = self.arrays
        nxt, cnt = A["lending.loan_next"], A["lending.counters"]
        while head >= 0:
            following = int(nxt[head])
            nxt[head] = int(cnt[LC_FREE_HEAD]) - 1
            cnt[LC_FREE_HEAD] = head + 1
            head = following

    def set_slot(self, slot, endowment):
        """A freshly configured agent (initial population or replacement) takes `slot`."""
        A = self.arrays
        if "lending" in self.families:
            for name in ("creditors_head", "debtors_head"):
                self._free_chain(int(A["lending." + name][slot]))
                A["lending." + name][slot] = -1
            A["lending.lending_factor"][slot] = endowment["lendingFactor"]
            A["lending.base_interest_rate"][slot] = endowment["baseInterestRate"]
            A["lending.loan_duration"][slot] = endowment["loanDuration"]
            A["lending.sugar_mean_income"][slot] = 1
            A["lending.spice_mean_income"][slot] = 1
            A["lending.last_lended"][slot] = -1
            A["lending.last_loans"][slot] = 0
        if "social" in self.families:
            A["social.max_friends"][slot] = endowment["maxFriends"]
            A["social.last_combat"][slot] = -1
            A["social.social_happiness"][slot] = 0
            A["social.conflict_happiness"][slot] = 0
            A["social.n_friends"][slot] = 0
        if "diseases" in self.families:
            A["diseases.n_sick"][slot] = 0
            A["diseases.modifier"][slot * DM_COUNT:(slot + 1) * DM_COUNT] = 0
            for name, fill in (("disease_death", 0), ("last_spread", -1), ("n_spread", 0), ("health_happiness", 0),
                               ("last_valid_moves", 0)):
                A["diseases." + name][slot] = fill
            bits = sum(b << k for k, b in enumerate(endowment["immuneSystem"]))
            A["diseases.immune"][slot] = A["diseases.start_immune"][slot] = bits
            A["diseases.protection"][slot] = endowment["diseaseProtectionChance"]

        if "misc" in self.families:
            A["misc.universal_sugar"][slot] = endowment["universalSugar"]
            A["misc.universal_spice"][slot] = endowment["universalSpice"]
            for name, fill in (("last_income_sugar", 0), ("last_income_spice", 0), ("depressed", 0), ("happiness_unit", 1),
                               ("health_happiness", 0), ("race", -1), ("racial_tags", 0)):
                A["misc." + name][slot] = fill
            rt = endowment["racialTags"]
            if rt is not None:
                A["misc.racial_tags"][slot] = sum(v << (2 * k) for k, v in enumerate(rt))
                A["misc.race"][slot] = max(set(rt), key=rt.count)

        if "group" in self.families:
            A["group.with_control"][slot * GI_KINDS:(slot + 1) * GI_KINDS] = 0
            A["group.with_experimental"][slot * GI_KINDS:(slot + 1) * GI_KINDS] = 0
            A["group.control_neighbors"][slot] = 0
            A["group.experimental_neighbors"][slot] = 0
            A["group.hood_n"][slot] = 0

        if "agentlog" in self.families:
            for name, fill in (("time_to_live", 0), ("last_pollution", 0), ("prey_wealth", 0), ("last_combat", -1),
                               ("last_mates", 0), ("valid_moves", 0), ("move_rank", 0), ("n_neighbors", 0)):
                A["agentlog." + name][slot] = fill

    # -- diseases: the init-time part of the rule (configureDiseases, sugarscape.py:286-337) ----------------
    def define_diseases(self, disease_endowments):
        defs = self.arrays["diseases.defs"]
        for i, e in enumerate(disease_endowments):
            if len(e["tags"]) > 32:
                raise NotImplementedError("disease tags longer than 32")
            d = defs[i]       # (start_timestep > 0: the disease waits in remainingDiseases, sugarscape.py:311-314)
            d["id"], d["incubation"], d["start_timestep"], d["tag_len"] = i, int(e["incubationPeriod"]), int(e["startTimestep"]), len(e["tags"])
            d["tags"] = sum(b << k for k, b in enumerate(e["tags"]))
            d["transmission"] = float(e["transmissionChance"])
            d["penalty"] = [float(e[name]) for name in DISEASE_PENALTIES]

    def disease_count(self, slot):
        return int(self.arrays["diseases.n_sick"][slot])

    def nearest_hamming(self, slot, di):
        """findNearestHammingDistanceInDisease (agent.py:1034-1050): (distance, start, end)."""
        d = self.arrays["diseases.defs"][di]
        length, tags, immune = int(d["tag_len"]), int(d["tags"]), int(self.arrays["diseases.immune"][slot])
        best, b0, b1 = length, 0, length - 1
        mask = (1 << length) - 1
        for i in range(self.immune_len - length):
            hd = bin(((immune >> i) & mask) ^ tags).count("1")
            if hd < best:
                best, b0, b1 = hd, i, i + (length - 1)
        return best, b0, b1

    def catch_initial(self, slot, di):
        """catchDisease(disease, initial=True) (agent.py:227-247, 184-188): no infection attempt, no infector."""
        A = self.arrays
        base, n = slot * self.max_sick, int(A["diseases.n_sick"][slot])
        if any(int(A["diseases.sick_disease"][base + k]) == di for k in range(n)):
            return
        hd, start, end = self.nearest_hamming(slot, di)
        if hd == 0:
            return
        d = A["diseases.defs"][di]
        for name, v in (("disease", di), ("start", start), ("end", end), ("caught", 0), ("incubation", int(d["incubation"])),
                        ("infector_slot", -1), ("infector_id", -1)):
            A["diseases.sick_" + name][base + n] = v
        A["diseases.n_sick"][slot] = n + 1
        d["infected"] += 1
        if int(d["incubation"]) == 0:
            A["diseases.modifier"][slot * DM_COUNT:(slot + 1) * DM_COUNT] += d["penalty"]

    def as_struct(self, pointer_of=lambda arr: arr.ctypes.data):
        x = Ext()
        x.capacity = self.capacity
        for fam, table in EXT_FAMILIES:
            if fam not in self.families:
                continue
            sub = getattr(x, fam)
            if fam == "misc":
                (sub.depression_percentage, sub.income_interval_sugar, sub.income_interval_spice, sub.racial_len,
                 sub.n_races) = self.misc_scalars
                sub.n_tables = len(self.arrays["misc.racial_picks"])
            if fam == "diseases":
                sub.max_sick, sub.immune_len, sub.n_diseases = self.max_sick, self.immune_len, self.n_diseases
                sub.n_immune_bits = len(self.arrays["diseases.immune_bits"])
            for n, dt, _, _ in table:
                if dt is None:          # pool_capacity / dead_capacity / max_loan_duration / the group's kind
                    setattr(sub, n, self.max_loan_duration if n == "max_loan_duration" else
                            self.group_kind if n == "kind" else self.write_records if n == "write_records" else
                            self.sizes[n.split("_")[0]])
                else:
                    setattr(sub, n, pointer_of(self.arrays[fam + "." + n]))
        return x

    def check_overflow(self):
        if "lending" in self.families and self.arrays["lending.counters"][LC_OVERFLOW] != 0:
            raise MemoryError("loan pool exhausted")

    def copy(self):
        h = HostExt.__new__(HostExt)
        h.capacity, h.families, h.sizes = self.capacity, set(self.families), dict(self.sizes)
        h.n_diseases, h.max_sick, h.immune_len = self.n_diseases, self.max_sick, self.immune_len
        h.racial_len, h.misc_scalars, h.max_loan_duration = self.racial_len, self.misc_scalars, self.max_loan_duration
        h.group_kind, h.write_records = self.group_kind, self.write_records
        h.arrays = {k: v.copy() for k, v in self.arrays.items()}
        return h


def default_capacity(c, n0):
    """Slots to allocate: births and replacements append, dead slots are recycled."""
    cells = c["environmentWidth"] * c["environmentHeight"]
    want = max(n0, int(c["agentReplacements"]), 1)
    if c["agentFertilityFactor"][1] > 0:
        want = max(want * 4, cells // 2)
    return int(min(max(2 * want, 1024), 2 * cells + 1024))


def make_params(c, rng_mode, flags, max_agent_range, equator):
    p = Params()
    p.abi_version = ABI_VERSION
    p.width, p.height = c["environmentWidth"], c["environmentHeight"]
    p.rng_mode = rng_mode
    p.seed = c["seed"] & 0xFFFFFFFFFFFFFFFF
    p.flags = flags
    p.max_agent_range = max_agent_range
    p.sugar_regrow = int(c["environmentSugarRegrowRate"])
    p.spice_regrow = int(c["environmentSpiceRegrowRate"])
    p.season_interval = int(c["environmentSeasonInterval"])
    p.seasonal_growback_delay = int(c["environmentSeasonalGrowbackDelay"])
    p.equator = equator
    p.pollution_start, p.pollution_end = c["environmentPollutionTimeframe"]
    p.diffusion_start, p.diffusion_end = c["environmentPollutionDiffusionTimeframe"]
    p.diffusion_delay = int(c["environmentPollutionDiffusionDelay"])
    p.sugar_prod_pollution = c["environmentSugarProductionPollutionFactor"]
    p.spice_prod_pollution = c["environmentSpiceProductionPollutionFactor"]
    p.sugar_cons_pollution = c["environmentSugarConsumptionPollutionFactor"]
    p.spice_cons_pollution = c["environmentSpiceConsumptionPollutionFactor"]
    p.max_combat_loot = c["environmentMaxCombatLoot"]
    p.tag_len = c["agentTagStringLength"]
    p.max_tribes = c["environmentMaxTribes"]
    p.max_timestep = int(min(c["timesteps"], 2 ** 31 - 1))
    p.id_bits = 26
    p.max_sugar_capacity = int(max(c["environmentMaxSugar"], 0))
    p.max_spice_capacity = int(max(c["environmentMaxSpice"], 0))
    p.inheritance_policy = INHERITANCE_CODE[c["agentInheritancePolicy"]]
    import math
    p.agent_replacements = int(c["agentReplacements"])
    p.quad_w = math.floor(c["environmentWidth"] / 2 * c["environmentQuadrantSizeFactor"])
    p.quad_h = math.floor(c["environmentHeight"] / 2 * c["environmentQuadrantSizeFactor"])
    p.quad_mask = sum(1 << (q - 1) for q in c["environmentStartingQuadrants"] if 1 <= q <= 4)
    p.tribe_per_quadrant = 1 if c["environmentTribePerQuadrant"] else 0
    return p


class HostState:
    """The whole simulation state as numpy arrays (host mirror of the device SoA)."""

    def __init__(self, width, height, capacity):
        self.width, self.height, self.capacity = width, height, capacity
        self.ethics = None          # HostEthics when the configuration uses decision models (ABI 2)
        self.ext = None             # HostExt when lending / friends ... are live (ABI 2)

    @classmethod
    def empty(cls, width, height, capacity, kin_capacity=None):
        s = cls(width, height, capacity)
        n = width * height
        for name, dt in GRID_FIELDS:
            fill = -1 if name == "occ" else 0
            setattr(s, name, np.full(n, fill, dtype=dt))
        s.counters = np.zeros(C_COUNT, dtype=np.int64)
        for name, dt, fill in AGENT_FIELDS:
            setattr(s, "a_" + name, np.full(capacity, fill, dtype=dt))
        s.kin_capacity = kin_capacity if kin_capacity is not None else 4 * capacity
        for name in KIN_FIELDS:
            setattr(s, name, np.full(s.kin_capacity, -1, dtype=np.int32))
        return s

    def copy(self):
        s = HostState(self.width, self.height, self.capacity)
        for name, _ in GRID_FIELDS:
            setattr(s, name, getattr(self, name).copy())
        s.counters = self.counters.copy()
        for name, _, _ in AGENT_FIELDS:
            setattr(s, "a_" + name, getattr(self, "a_" + name).copy())
        s.kin_capacity = self.kin_capacity
        for name in KIN_FIELDS:
            setattr(s, name, getattr(self, name).copy())
        s.ethics = self.ethics.copy() if self.ethics is not None else None
        s.ext = self.ext.copy() if self.ext is not None else None
        return s

    def as_structs(self):
        """ctypes views (HOST pointers) of this state, for the CPU oracle used by the tests."""
        g = Grid()
        g.n_cells = self.width * self.height
        for name, _ in GRID_FIELDS:
            setattr(g, name, getattr(self, name).ctypes.data)
        a = Agents()
        a.capacity = self.capacity
        a.counters = self.counters.ctypes.data
        for name, _, _ in AGENT_FIELDS:
            setattr(a, name, getattr(self, "a_" + name).ctypes.data)
        a.kin_capacity = self.kin_capacity
        for name in KIN_FIELDS:
            setattr(a, name, getattr(self, name).ctypes.data)
        return g, a

    # -- agents -------------------------------------------------------------------------------
    def append_agents(self, new_agents, protect_last=0):
        """Append freshly placed agents (init / replacement) at the end of the list.

        protect_last: number of entries on top of the free-slot stack that must not be reused
        yet (the slots of this step's dead, which the stats reduction still has to read)."""
        n_list = int(self.counters[C_N_LIST])
        n_slots = int(self.counters[C_N_SLOTS])
        first_placement = n_slots == 0
        n_free = int(self.counters[C_N_FREE])
        protected = self.a_free_slots[n_free - protect_last:n_free].copy() if protect_last else None
        n_free -= protect_last
        for a in new_agents:
            if n_free > 0:
                n_free -= 1
                slot = int(self.a_free_slots[n_free])
            else:
                slot = n_slots
                n_slots += 1
            if slot >= self.capacity or n_list >= self.capacity:
                raise MemoryError("agent capacity exceeded")
            e = a["endowment"]
            for name, _, fill in AGENT_FIELDS[3:]:
                getattr(self, "a_" + name)[slot] = fill
            self.a_id[slot] = a["id"]
            self.a_pos[slot] = a["cell"]
            self.a_born[slot] = a["born"]
            self.a_sugar[slot] = e["sugar"]
            self.a_spice[slot] = e["spice"]
            self.a_start_sugar[slot] = e["sugar"]
            self.a_start_spice[slot] = e["spice"]
            self.a_sugar_metab[slot] = e["sugarMetabolism"]
            self.a_spice_metab[slot] = e["spiceMetabolism"]
            self.a_vision[slot] = e["vision"]
            self.a_movement[slot] = e["movement"]
            self.a_max_age[slot] = e["maxAge"]
            self.a_sex[slot] = SEX_CODE[e["sex"]]
            self.a_fert_age[slot] = e["fertilityAge"]
            self.a_infert_age[slot] = e["infertilityAge"]
            self.a_fert_factor[slot] = e["fertilityFactor"]
            self.a_aggression[slot] = e["aggressionFactor"]
            self.a_lookahead[slot] = e["lookaheadFactor"]
            self.a_trade_factor[slot] = e["tradeFactor"]
            tags = a["tags"]
            self.a_tags[slot] = sum((1 << i) for i, b in enumerate(tags or ()) if b)
            self.a_tribe[slot] = a["tribe"]
            if self.ethics is not None:
                self.ethics.set_slot(slot, e)
            if self.ext is not None:
                self.ext.set_slot(slot, e)
                if "misc" in self.ext.families and e["depressionFactor"] == 1:
                    self.depress(slot)
            if self.ethics is not None:     # Temperance.totalMetabolism, taken at construction (after a depression trigger)
                self.ethics.arrays["total_metabolism"][slot] = max(int(self.a_sugar_metab[slot]), 0) + max(int(self.a_spice_metab[slot]), 0)
            self.a_order[n_list] = slot
            self.occ[a["cell"]] = slot
            n_list += 1
        if self.ethics is not None and self.ethics.hood_stride and len(new_agents):
            # configureAgents ends with findNeighborhood() for EVERY agent (sugarscape.py:265-267); the first time this happens
            # before the diseases are handed out (-2: ranges without their modifiers), at a replacement with them (-1)
            self.ethics.arrays["hood_n"][:] = -2 if first_placement else -1
        if protect_last:
            self.a_free_slots[n_free:n_free + protect_last] = protected
            n_free += protect_last
        self.counters[C_N_LIST] = n_list
        self.counters[C_N_SLOTS] = n_slots
        self.counters[C_N_FREE] = n_free

    def newest_members(self, k):
        """How many of the k agents at the end of the list (this step's replacements) are in the experimental group
        (Agent.isInGroup, agent.py:1256-1284, for the kinds ssb_group_t.kind holds): `<group>AgentsReplaced` of the log."""
        if self.ext is None or "group" not in self.ext.families or k <= 0:
            return 0
        n_list, kind = int(self.counters[C_N_LIST]), self.ext.group_kind
        slots, A = self.a_order[n_list - k:n_list], self.ext.arrays
        which = kind & 0xff
        if which == GROUP_KINDS["depressed"]:
            return int((A["misc.depressed"][slots] != 0).sum())
        if which in (GROUP_KINDS["female"], GROUP_KINDS["male"]):
            return int((self.a_sex[slots] == (SEX_FEMALE if which == GROUP_KINDS["female"] else SEX_MALE)).sum())
        if which == GROUP_RACE:
            return int((A["misc.race"][slots] == (kind >> 8)).sum())
        if which == GROUP_MODEL:
            return int(((self.ethics.arrays["top_ordering"][slots] & 2) != 0).sum())
        if which == GROUP_SICK:
            return int((A["diseases.n_sick"][slots] > 0).sum()) if "diseases" in self.ext.families else 0
        raise ValueError("unknown experimental group kind %d" % kind)

    def depress(self, slot):
        """Depression.trigger (condition.py:27-33) at the creation of a depressed agent (agent.py:137-141)."""
        import math
        A = self.ext.arrays
        A["misc.depressed"][slot] = 1
        self.a_aggression[slot] *= 1.145
        if "social" in self.ext.families:
            A["social.max_friends"][slot] = math.ceil(A["social.max_friends"][slot] * 0.6333)
        A["misc.happiness_unit"][slot] *= 0.5763
        # (v *= ceil(v * f), saturated at 2^30 like the device: csrc/ssb_ext.cuh:depress_scale)
        for arr, f in ((self.a_movement, 0.429), (self.a_spice_metab, 1.544), (self.a_sugar_metab, 1.544)):
            v = int(arr[slot])
            arr[slot] = min(v * math.ceil(v * f), 1 << 30)

    def list_view(self):
        """Arrays of the live agent list in list order (for parity checks and stats)."""
        n = int(self.counters[C_N_LIST])
        slots = self.a_order[:n]
        H = self.height
        pos = self.a_pos[slots]
        return {
            "a_id": self.a_id[slots].astype(np.int64), "a_x": (pos // H).astype(np.int64),
            "a_y": (pos % H).astype(np.int64), "a_age": self.a_age[slots].astype(np.int64),
            "a_sugar": self.a_sugar[slots].copy(), "a_spice": self.a_spice[slots].copy(),
            "a_tags": np.where(self.a_tribe[slots] >= 0, self.a_tags[slots].astype(np.int64), -1),
            "a_tribe": self.a_tribe[slots].astype(np.int64),
        }

    def canonical(self):
        """Canonical snapshot in the format of tests/golden/ref_harness.snapshot."""
        W, H = self.width, self.height
        snap = self.list_view()
        occ_id = np.where(self.occ >= 0, self.a_id[np.maximum(self.occ, 0)], -1).astype(np.int64)
        snap.update({
            "sugar": self.sugar.astype(np.float64).reshape(W, H),
            "spice": self.spice.astype(np.float64).reshape(W, H),
            "pollution": self.pollution.reshape(W, H).copy(),
            "occ": occ_id.reshape(W, H),
        })
        return snap
