"""Initial state of a simulation, as struct-of-arrays (host side, cold path).

Restates the reference's start-up sequence so that the main CPython ``random`` stream is at
exactly the same position when the first timestep begins:

  configureEnvironment / addResourcePeak   reference sugarscape.py:348-387,154-169
  findActiveQuadrants                      reference sugarscape.py:475-497
  randomizeAgentEndowments                 reference sugarscape.py:577-758
  configureAgents (initial + replacement)  reference sugarscape.py:171-267
  generateTribeTags                        reference sugarscape.py:548-559
  configureDiseases (the unconditional list shuffle)  reference sugarscape.py:305

Everything that draws random numbers uses the stdlib ``random`` module, i.e. the very generator
the reference uses; the resulting MT19937 state is then handed to the device
(``ssb_set_mt_state``).  The arrays built here become the device SoA of include/ssb.h.
"""
import hashlib
import math
import random

import numpy as np

from . import layout, steptables


class UnsupportedConfiguration(NotImplementedError):
    """Raised loudly for rule families this engine does not implement yet (no CPU fallback)."""


# reference sugarscape.py:625-657 -- name of the endowment, option holding its [min, max]
_RANGED_ENDOWMENTS = (
    ("aggressionFactor", "agentAggressionFactor"),
    ("baseInterestRate", "agentBaseInterestRate"),
    ("decisionModelAgeismFactor", "agentDecisionModelAgeismFactor"),
    ("decisionModelFactor", "agentDecisionModelFactor"),
    ("decisionModelLookaheadDiscount", "agentDecisionModelLookaheadDiscount"),
    ("decisionModelRacismFactor", "agentDecisionModelRacismFactor"),
    ("decisionModelSexismFactor", "agentDecisionModelSexismFactor"),
    ("decisionModelTribalFactor", "agentDecisionModelTribalFactor"),
    ("diseaseProtectionChance", "agentDiseaseProtectionChance"),
    ("dynamicDecisionModelFactor", "agentDynamicDecisionModelFactor"),
    ("dynamicSelfishnessFactor", "agentDynamicSelfishnessFactor"),
    ("dynamicSocialPressureFactor", "agentDynamicSocialPressureFactor"),
    ("femaleFertilityAge", "agentFemaleFertilityAge"),
    ("femaleInfertilityAge", "agentFemaleInfertilityAge"),
    ("fertilityFactor", "agentFertilityFactor"),
    ("lendingFactor", "agentLendingFactor"),
    ("loanDuration", "agentLoanDuration"),
    ("lookaheadFactor", "agentLookaheadFactor"),
    ("maleFertilityAge", "agentMaleFertilityAge"),
    ("maleInfertilityAge", "agentMaleInfertilityAge"),
    ("maxAge", "agentMaxAge"),
    ("maxFriends", "agentMaxFriends"),
    ("movement", "agentMovement"),
    ("selfishnessFactor", "agentSelfishnessFactor"),
    ("spice", "agentStartingSpice"),
    ("spiceMetabolism", "agentSpiceMetabolism"),
    ("sugar", "agentStartingSugar"),
    ("sugarMetabolism", "agentSugarMetabolism"),
    ("tradeFactor", "agentTradeFactor"),
    ("universalSpice", "agentUniversalSpice"),
    ("universalSugar", "agentUniversalSugar"),
    ("vision", "agentVision"),
)


def _md5_int(name):
    return int(hashlib.md5(name.encode()).hexdigest(), 16)


def _decimals(value):
    parts = str(value).split(".")
    return len(parts[1]) if len(parts) == 2 else None


# decisionModel strings with a CUDA implementation: the plain Agent and ethics.Bentham with its spawn
# variants (sugarscape.py:219-238, dynamic selfishness included); Temperance, the "Top" orderings and the Leader are refused
SUPPORTED_DECISION_MODELS = frozenset(("none", "rawSugarscape")) | frozenset(
    base + look + dyn + top for base in layout.BENTHAM_MODELS for look in ("", "NoLookahead", "HalfLookahead")
    for dyn in ("", "Dynamic") for top in ("", "Top") if not (top and base == "negativeBentham")      # (that one raises in the reference)
) | frozenset(("temperance", "temperancePECS", "asimov"))


def uses_decision_models(c):
    """True if some agent may be an ethics.Bentham instance (needs the ssb_ethics_t extension, mode E)."""
    return any(layout.decision_model_class(m) != 0 for m in c["agentDecisionModels"])


def check_supported(c):
    """Fail loudly on rule families without a CUDA implementation yet (DESIGN.md 'scope')."""
    problems = []
    if is_radial(c):
        # environment.py:146-173: discs of floor(distance) <= range instead of crosses (exact mode, torus; the engine's
        # candidate arrays hold RADIAL_MAX_CELLS cells -- a range of 6 on a map at least 13 cells wide)
        if c.get("b200RngMode", "exact") != "exact":
            problems.append("radial ranges in the scalable RNG mode")
        if radial_cells_in_range(c) > RADIAL_MAX_CELLS:
            problems.append("radial ranges of more than %d cells" % RADIAL_MAX_CELLS)
    if c["neighborhoodMode"] not in ("vonNeumann", "moore"):
        problems.append("neighborhoodMode " + str(c["neighborhoodMode"]))
    if c["neighborhoodMode"] == "moore":      # cell.py:77-89: eight neighbour cells (tagging, trade, births, lending, diseases, pollution flux)
        if c.get("b200RngMode", "exact") != "exact":
            problems.append("moore neighbourhood in the scalable RNG mode")
        if c["agentLogfile"] is not None:
            problems.append("moore neighbourhood with an agentLogfile")
    if not c["environmentWraparound"] and is_radial(c):
        if c.get("b200RngMode", "exact") != "exact":
            problems.append("environmentWraparound=false in the scalable RNG mode")
    elif not c["environmentWraparound"]:
        # cell.py:47-121: the neighbour lists lose the cells across the edge; the range tables (environment.py:111-122) do NOT
        # change as long as no range reaches half the map (they keep wrapping by `% width`, with the plain distance)
        # (an agent looks min(vision, movement) cells far; the memoised tables reach further without wraparound --
        # min(range, width - 1) instead of width // 2 -- but nobody reads those entries)
        # (disease bonuses can push an agent further: the engine flags that case at run time, overflow code 14)
        reach = min(c["agentVision"][1], c["agentMovement"][1])
        if reach > min(c["environmentWidth"], c["environmentHeight"]) // 2:
            problems.append("environmentWraparound=false with agents that see beyond half the map")
        if c.get("b200RngMode", "exact") != "exact":
            problems.append("environmentWraparound=false in the scalable RNG mode")
    if c["agentLeader"]:
        problems.append("agentLeader")
    if c["diseaseList"]:
        problems.append("diseaseList")
    if c["startingDiseases"] > 0 and (c["agentImmuneSystemLength"] > 64 or c["diseaseTagStringLength"][1] > 32 or
                                      c["startingDiseases"] > 64):
        problems.append("immune systems longer than 64 / disease tags longer than 32 / more than 64 diseases")
    if c["environmentMaxRaces"] > 0 and (c["agentRacialTagStringLength"] > 16 or c["environmentMaxRaces"] > 4):
        problems.append("racial tag strings longer than 16 / more than 4 races")
    if c["agentInheritancePolicy"] not in ("none", "children", "sons", "daughters"):
        problems.append("inheritance policy " + str(c["agentInheritancePolicy"]))
    if c["agentMaxFriends"][1] > layout.MAX_FRIENDS:
        problems.append("more than %d friends" % layout.MAX_FRIENDS)
    if "asimov" in c["agentDecisionModels"]:
        # ethics.Asimov (ethics.py:8-98) among plain agents only: its second law asks a neighbour with a decision model of its
        # own for THAT neighbour's findEthicalValueOfCell, which reads the neighbourhood list stored at the neighbour's last move
        # (negativeBentham's value is a dict: the reference raises when a robot compares it with 0, ethics.py:27)
        if any("negativeBentham" in m for m in c["agentDecisionModels"]):
            problems.append("asimov next to negativeBentham")
    bad_models = [m for m in c["agentDecisionModels"] if m not in SUPPORTED_DECISION_MODELS]
    if bad_models:
        problems.append("decision models other than rawSugarscape and the Bentham family: " + ", ".join(bad_models))
    # (agentDecisionModelFactor > 0 without a decision model: the plain Agent's findBestEthicalCell hands back the
    # greedy cell, agent.py:730-733 -- nothing to refuse)
    for k in ("agentDecisionModelAgeismFactor", "agentDecisionModelRacismFactor",
              "agentDecisionModelSexismFactor", "agentDecisionModelTribalFactor"):
        if c[k][1] >= 0:
            problems.append(k)
    if c["experimentalGroup"] is not None and layout.group_kind(c) == 0:
        problems.append("experimentalGroup other than " + " / ".join(sorted(layout.GROUP_KINDS)) + " / race<N> / a decision model's name")
    if layout.group_kind(c) == layout.GROUP_MODEL and not uses_decision_models(c):
        problems.append("a decision model as experimentalGroup without decision models")
    if c["agentTagStringLength"] > 32:
        problems.append("tag strings longer than 32")
    for k in ("environmentSugarRegrowRate", "environmentSpiceRegrowRate"):
        if c[k] != int(c[k]):
            problems.append("fractional " + k)
    for k in ("agentSugarMetabolism", "agentSpiceMetabolism", "agentVision", "agentMovement",
              "agentMaxAge"):
        if any(v != int(v) for v in c[k]):
            problems.append("fractional " + k)
    if c["environmentHeight"] != c["environmentWidth"] and c["environmentPollutionDiffusionDelay"] > 0:
        # reference quirk: diffusion indexes grid[i][j] with i over height
        problems.append("pollution diffusion on a non-square map")
    if problems:
        raise UnsupportedConfiguration(
            "this engine has no CUDA implementation yet for: " + ", ".join(problems))


# ---------------------------------------------------------------------------------------------
# environment
# ---------------------------------------------------------------------------------------------

def capacity_field(width, height, peaks):
    """addResourcePeak (sugarscape.py:154-169) for all peaks; x-major int32 [W, H]."""
    cap = np.zeros((width, height), dtype=np.int64)
    radius = math.ceil(math.sqrt(2 * (height + width)))
    xs = np.arange(width, dtype=np.float64)[:, None]
    ys = np.arange(height, dtype=np.float64)[None, :]
    for px, py, value in peaks:
        dispersion = math.sqrt(max(px, width - px) ** 2 + max(py, height - py) ** 2) * (radius / width)
        dist = np.sqrt((px - xs) ** 2 + (py - ys) ** 2)
        level = np.ceil(np.minimum(1 + value * (1 - dist / dispersion), value)).astype(np.int64)
        cap = np.maximum(cap, level)
    return cap.astype(np.int32)


def max_agent_range(c):
    """max(maxVision, maxMovement) incl. worst-case disease bonuses (environment.py:135-137);
    the per-axis caps (width // 2, height // 2) are applied by the engine."""
    max_vision = c["startingDiseases"] * max(c["diseaseVisionPenalty"][1], 0) + c["agentVision"][1]
    max_movement = c["startingDiseases"] * max(c["diseaseMovementPenalty"][1], 0) + c["agentMovement"][1]
    return int(min(max(max_vision, max_movement), 2 ** 30))


RADIAL_MAX_CELLS = 160      # MAXC_EXACT of csrc/ssb_exact.cuh


def is_radial(c):
    """environment.py:146, 152: radial ranges only when BOTH modes say so; anything else is the cardinal cross."""
    return c["agentVisionMode"] == "radial" and c["agentMovementMode"] == "radial"


def radial_cells_in_range(c):
    """len(cellsInRange) of the furthest-seeing agent the configured ranges allow (environment.py:132-173,
    agent.py:766-779; disease bonuses can push an agent further: the engine flags that at run time, overflow code 15)."""
    W, H, memo = c["environmentWidth"], c["environmentHeight"], max_agent_range(c)
    wrap = c["environmentWraparound"]
    far_x, far_y = (W // 2, H // 2) if wrap else (W - 1, H - 1)
    mdx, mdy = min(memo, far_x), min(memo, far_y)
    cap = min(c["agentVision"][1], c["agentMovement"][1], memo, math.isqrt(far_x ** 2 + far_y ** 2))
    if not wrap:        # the disc is clipped at the map's edge: the agent in the middle sees most (the per-axis caps never bind)
        return sum(0 < (x - W // 2) ** 2 + (y - H // 2) ** 2 < (cap + 1) ** 2 for x in range(W) for y in range(H))
    n = 0
    for ox in range(-mdx, mdx + 1):
        for oy in range(-mdy, mdy + 1):
            if (ox == -mdx and 2 * mdx == W) or (oy == -mdy and 2 * mdy == H):
                continue
            n += 0 < ox * ox + oy * oy < (cap + 1) ** 2
    return n


def equator(c):
    """environment.py:10."""
    return c["environmentEquator"] if c["environmentEquator"] >= 0 else math.ceil(c["environmentHeight"] / 2)


def active_quadrants(c):
    """findActiveQuadrants (sugarscape.py:475-497): lists of cell indices (x*H + y), built
    y-outer / x-inner like the reference's comprehensions."""
    W, H = c["environmentWidth"], c["environmentHeight"]
    qw = math.floor(W / 2 * c["environmentQuadrantSizeFactor"])
    qh = math.floor(H / 2 * c["environmentQuadrantSizeFactor"])
    spans = {1: (range(qh), range(qw)),
             2: (range(qh), range(W - qw, W)),
             3: (range(H - qh, H), range(W - qw, W)),
             4: (range(H - qh, H), range(qw))}
    quadrants = []
    for q in (1, 2, 3, 4):
        if q in c["environmentStartingQuadrants"]:
            ys, xs = spans[q]
            quadrants.append([x * H + y for y in ys for x in xs])
    return quadrants


# ---------------------------------------------------------------------------------------------
# endowments
# ---------------------------------------------------------------------------------------------

def generate_tribe_tags(c, tribe):
    """generateTribeTags (sugarscape.py:548-559): one randint + one shuffle."""
    length = c["agentTagStringLength"]
    tribe_size = (length + 1) / c["environmentMaxTribes"]
    min_zeroes = math.floor(tribe * tribe_size)
    max_zeroes = min(math.floor((tribe + 1) * tribe_size) - 1, length)
    zeroes = random.randint(min_zeroes, max_zeroes)
    tags = [0] * zeroes + [1] * (length - zeroes)
    random.shuffle(tags)
    return tags


def randomize_endowments(c, n, timestep):
    """randomizeAgentEndowments (sugarscape.py:577-758) -> list of endowment dicts."""
    n_depressed = int(math.ceil(n * c["agentDepressionPercentage"]))
    depression = [1] * n_depressed + [0] * (n - n_depressed)
    random.shuffle(depression)  # always drawn, even when all zero (Appendix C.6)

    ranged = {}
    for name, option in _RANGED_ENDOWMENTS:
        lo, hi = c[option][0], c[option][1]
        if name == "universalSpice":
            hi = c["agentUniversalSugar"][1]  # reference quirk (sugarscape.py:654)
        decs = [d for d in (_decimals(lo), _decimals(hi)) if d is not None]
        decimals = max(decs) if decs else 0
        ranged[name] = {"values": [], "curr": lo, "min": lo, "max": hi,
                        "inc": 10 ** (-1 * decimals), "decimals": decimals}

    # racial tags (generateAgentRacialTags / generateRacialTags, sugarscape.py:504-515, 535-546); the engine
    # refuses configurations with racial tags, the oracle-only tests use them
    n_races, racial_len = c["environmentMaxRaces"], c["agentRacialTagStringLength"]
    if racial_len == 0 or n_races == 0:
        racial_tags = [None] * n
    else:
        racial_tags = []
        for i in range(n):
            race = i % n_races
            if n_races == 1:
                racial_tags.append([race] * racial_len)
                continue
            assigned = random.randint(math.floor(racial_len / 2) + 1, racial_len)
            others = [r for r in range(n_races) if r != race]
            rt = [race] * assigned + [random.choice(others) for _ in range(racial_len - assigned)]
            random.shuffle(rt)
            racial_tags.append(rt)
        random.shuffle(racial_tags)
    # tribe tags (sugarscape.py:517-528)
    if c["agentTagStringLength"] == 0 or c["environmentMaxTribes"] == 0 or c["environmentTribePerQuadrant"]:
        tags = [None] * n
    else:
        tags = [generate_tribe_tags(c, i % c["environmentMaxTribes"]) for i in range(n)]
        random.shuffle(tags)

    ratio = c["agentMaleToFemaleRatio"]
    sexed = ratio is not None and ratio != 0
    males_left = math.floor(n / (ratio + 1)) * ratio if sexed else n
    sexes, models, immune = [], [], []
    model_names = c["agentDecisionModels"]
    for i in range(n):
        for r in ranged.values():
            r["values"].append(r["curr"])
            r["curr"] = round(r["curr"] + r["inc"], r["decimals"])
            if r["curr"] > r["max"]:
                r["curr"] = r["min"]
        if c["agentImmuneSystemLength"] > 0:
            immune.append([random.randrange(2) for _ in range(c["agentImmuneSystemLength"])])
        else:
            immune.append(None)
        if sexed:
            if males_left == 0:
                sexes.append("female")
            else:
                sexes.append("male")
                males_left -= 1
        else:
            sexes.append(None)
        model = model_names[i % len(model_names)]
        models.append("none" if model == "rawSugarscape" else model)

    # private hash-seeded shuffles: no main-stream consumption (sugarscape.py:725-729)
    saved = random.getstate()
    for name, r in ranged.items():
        random.seed(_md5_int(name) + timestep)
        random.shuffle(r["values"])
    random.setstate(saved)
    random.shuffle(sexes)
    random.shuffle(models)

    sex_specific = {"femaleFertilityAge": ("female", "fertilityAge"),
                    "femaleInfertilityAge": ("female", "infertilityAge"),
                    "maleFertilityAge": ("male", "fertilityAge"),
                    "maleInfertilityAge": ("male", "infertilityAge")}
    endowments = []
    for i in range(n):
        e = {"sex": sexes[i], "racialTags": racial_tags.pop(), "tags": tags.pop(),
             "immuneSystem": immune.pop(), "decisionModel": models.pop(),
             "depressionFactor": depression[i]}
        for name, r in ranged.items():
            if name in sex_specific:
                # a sexed agent pops ALL four lists: its own sex's values become
                # fertilityAge / infertilityAge, the other sex's land in unused keys
                # (the `else` branch of sugarscape.py:738-752); unsexed agents pop none
                if sexes[i] is None:
                    continue
                sex, target = sex_specific[name]
                value = r["values"].pop()
                if sexes[i] == sex:
                    e[target] = value
                continue
            e[name] = r["values"].pop()
        if sexes[i] is None:
            e["fertilityAge"] = 0
            e["infertilityAge"] = 0
        endowments.append(e)
    return endowments


def find_tribe(c, tags):
    """agent.py:1142-1151."""
    if tags is None:
        return -1
    zeroes = tags.count(0)
    tribe_size = (c["agentTagStringLength"] + 1) / c["environmentMaxTribes"]
    return min(math.ceil((zeroes + 1) / tribe_size) - 1, c["environmentMaxTribes"] - 1)


def tags_mask(tags):
    m = 0
    for i, b in enumerate(tags or ()):
        if b:
            m |= 1 << i
    return m


# ---------------------------------------------------------------------------------------------
# placement (initial population and replacements)
# ---------------------------------------------------------------------------------------------

class Population:
    """Host-side bookkeeping that outlives init: the endowment list reused cyclically by
    replacements (sugarscape.py:207-208) and the ID counter."""

    def __init__(self, configuration):
        self.c = configuration
        self.endowments = []
        self.endowment_index = 0
        self.quadrants = active_quadrants(configuration)

    def place(self, num_agents, occupied, timestep, next_id):
        """configureAgents (sugarscape.py:171-267) without the object graph.

        occupied: boolean array over cells (True = cell.agent is not None).
        Returns a list of new-agent dicts (cell, id, endowment fields, tags, tribe) in list
        order; consumes the main random stream exactly like the reference."""
        c = self.c
        empty = [[cell for cell in quadrant if not occupied[cell]] for quadrant in self.quadrants]
        total = sum(len(q) for q in empty)
        if total == 0:
            return []
        num_agents = int(min(num_agents, total))
        if not self.endowments:
            self.endowments = randomize_endowments(c, num_agents, timestep)
        for quadrant in empty:
            random.shuffle(quadrant)
        indices = list(range(len(empty)))
        random.shuffle(indices)
        new_agents = []
        for i in range(num_agents):
            q = indices[i % len(empty)]
            cell = empty[q].pop()
            e = self.endowments[self.endowment_index % len(self.endowments)]
            self.endowment_index += 1
            tags = e["tags"]
            if c["environmentTribePerQuadrant"]:
                tags = generate_tribe_tags(c, q)
            new_agents.append({"cell": cell, "id": next_id + i, "born": timestep,
                               "endowment": e, "tags": tags, "tribe": find_tribe(c, tags)})
        return new_agents


def rule_flags(c):
    f = 0
    if (c["environmentMaxSpice"] > 0 and c["environmentSpicePeaks"]) or c["environmentFile"] is not None \
            or c["agentSpiceMetabolism"][1] > 0 \
            or c["agentStartingSpice"][1] > 0 or c["environmentSpiceRegrowRate"] > 0:
        f |= layout.F_SPICE
    if c["environmentPollutionTimeframe"] != [0, 0] or c["environmentPollutionDiffusionDelay"] > 0:
        f |= layout.F_POLLUTION
    if c["agentFertilityFactor"][1] > 0:
        f |= layout.F_REPRO
    if c["agentTagStringLength"] > 0 and c["environmentMaxTribes"] > 0:
        f |= layout.F_TAGS
    if c["agentTagging"]:
        f |= layout.F_TAGGING
    if c["agentAggressionFactor"][1] > 0:
        f |= layout.F_COMBAT
    if c["agentTradeFactor"][1] != 0:
        f |= layout.F_TRADE
    if c["agentTagPreferences"]:
        f |= layout.F_TAGPREF
    if not c["environmentWraparound"]:
        f |= layout.F_NOWRAP
    if c["neighborhoodMode"] == "moore":
        f |= layout.F_MOORE
    if is_radial(c):
        f |= layout.F_RADIAL
    return f


_DISEASE_ENDOWMENTS = (          # key order of the reference's dict (sugarscape.py:775-787)
    ("aggressionPenalty", "diseaseAggressionPenalty"), ("fertilityPenalty", "diseaseFertilityPenalty"),
    ("friendlinessPenalty", "diseaseFriendlinessPenalty"), ("happinessPenalty", "diseaseHappinessPenalty"),
    ("incubationPeriod", "diseaseIncubationPeriod"), ("movementPenalty", "diseaseMovementPenalty"),
    ("spiceMetabolismPenalty", "diseaseSpiceMetabolismPenalty"), ("startTimestep", "diseaseTimeframe"),
    ("sugarMetabolismPenalty", "diseaseSugarMetabolismPenalty"), ("tagLength", "diseaseTagStringLength"),
    ("transmissionChance", "diseaseTransmissionChance"), ("visionPenalty", "diseaseVisionPenalty"),
)


def randomize_disease_endowments(c, n, timestep):
    """randomizeDiseaseEndowments (sugarscape.py:760-840): list of disease configurations.  The tag
    strings consume the main stream, the shuffles of the endowment lists private hash-seeded ones."""
    ranged = {}
    for name, option in _DISEASE_ENDOWMENTS:
        lo, hi = c[option][0], c[option][1]
        decs = [d for d in (_decimals(lo), _decimals(hi)) if d is not None]
        decimals = max(decs) if decs else 0
        ranged[name] = {"values": [], "curr": lo, "min": lo, "max": hi, "inc": 10 ** (-1 * decimals), "decimals": decimals}
    tags = []
    for _ in range(n):
        for name, r in ranged.items():
            if name == "tagLength":
                tags.append([random.randrange(2) for _ in range(r["curr"])])
            r["values"].append(r["curr"])
            r["curr"] += r["inc"]
            r["curr"] = round(r["curr"], r["decimals"])
            if r["curr"] > r["max"]:
                r["curr"] = r["min"]
    saved = random.getstate()
    for name, r in ranged.items():
        random.seed(_md5_int(name) + timestep)
        random.shuffle(r["values"])
    random.setstate(saved)
    random.shuffle(tags)
    endowments = []
    for _ in range(n):
        e = {"tags": tags.pop()}
        for name, r in ranged.items():
            e[name] = r["values"].pop()
        endowments.append(e)
    return endowments


def load_environment_file(c):
    """configureEnvironment with an environmentFile (sugarscape.py:370-384): a JSON list of lists of
    {"sugar", "spice"}.  The reference passes (spice, sugar) into Cell(x, y, env, maxSugar, maxSpice), so the
    file's "spice" becomes the SUGAR capacity and vice versa -- restated literally.  Cells start full."""
    import json
    with open(c["environmentFile"]) as f:
        data = json.load(f)
    W, H = c["environmentWidth"], c["environmentHeight"]
    if len(data) != H or len(data[0]) != W or W != H:
        # the reference sizes the grid from the options but indexes the file [x][y] with height = len(file)
        raise UnsupportedConfiguration("environmentFile must be square and match environmentWidth / environmentHeight")
    sugar_cap = np.zeros((W, H), dtype=np.int64)
    spice_cap = np.zeros((W, H), dtype=np.int64)
    for x in range(W):
        for y in range(H):
            sugar_cap[x, y] = data[x][y]["spice"]
            spice_cap[x, y] = data[x][y]["sugar"]
    if sugar_cap.max(initial=0) > max(c["environmentMaxSugar"], 0) or spice_cap.max(initial=0) > max(c["environmentMaxSpice"], 0) \
            or sugar_cap.min(initial=0) < 0 or spice_cap.min(initial=0) < 0:
        raise UnsupportedConfiguration("environmentFile capacities must lie within [0, environmentMaxSugar / environmentMaxSpice] "
                                       "(note the reference swaps the file's sugar and spice)")
    return sugar_cap, spice_cap


def configure_diseases(c, state, disease_endowments):
    """configureDiseases after the agent shuffle (sugarscape.py:317-337): the initial infections, drawn from
    the main stream like the rest of the init path."""
    ext = state.ext
    ext.define_diseases(disease_endowments)
    n = int(state.counters[layout.C_N_LIST])
    agents = [int(s) for s in state.a_order[:n]]
    initial = [i for i, e in enumerate(disease_endowments) if int(e["startTimestep"]) == 0]      # (the others wait: remainingDiseases)
    lo, hi = c["startingDiseasesPerAgent"]
    unlimited = [lo, hi] == [0, 0]
    curr = lo
    listed = ext.arrays["diseases.defs"]["listed"]
    for s in agents:
        random.shuffle(initial)
        for d in initial:
            if ext.disease_count(s) >= curr and not unlimited:
                curr += 1
                break
            if ext.nearest_hamming(s, d)[0] == 0:
                continue
            ext.catch_initial(s, d)
            listed[d] += 1          # sugarscape.diseases gains one entry per initial infection
            if unlimited:
                initial.remove(d)
                break
        if curr > hi:
            curr = lo


def build_initial_state(c, capacity=None, check=True, exact=True):
    """Environment + initial population + the init-time list shuffle; returns (state, population).

    Must be called right after verify_random_seed/verify_configuration so the main stream is
    positioned as in the reference's Sugarscape.__init__ (sugarscape.py:19-75).
    check=False skips the support check and the ABI 2 extension state (oracle-only tests that keep the rule
    families in the oracle binding); exact=False (scalable RNG mode) builds the ABI 1 state only -- the caller
    refuses configurations whose families exist in the exact mode only."""
    if check:
        check_supported(c)
    W, H = c["environmentWidth"], c["environmentHeight"]
    if c["environmentFile"] is None:
        sugar_cap = capacity_field(W, H, c["environmentSugarPeaks"])
        spice_cap = capacity_field(W, H, c["environmentSpicePeaks"])
    else:
        sugar_cap, spice_cap = load_environment_file(c)
    pop = Population(c)
    n0 = int(c["startingAgents"])
    if capacity is None:
        capacity = layout.default_capacity(c, n0)
    state = layout.HostState.empty(W, H, capacity)
    state.sugar_cap[:] = sugar_cap.reshape(-1)
    state.spice_cap[:] = spice_cap.reshape(-1)
    state.sugar[:] = state.sugar_cap
    state.spice[:] = state.spice_cap

    # (oracle-only tests pass check=False: the bare ABI 1 state, the rule families live in the oracle binding)
    if check and exact and uses_decision_models(c):
        state.ethics = layout.HostEthics(capacity, c)
    diseased = layout.HostExt.uses_diseases(c)
    horizon = int(min(c["timesteps"], 1 << 20)) + 8

    occupied = np.zeros(W * H, dtype=bool)
    new_agents = pop.place(n0, occupied, 0, 0) if n0 > 0 else []
    # configureDiseases (sugarscape.py:286-305): the disease endowments are drawn first, then the agent
    # list is shuffled once -- always, when any agent exists
    pop.disease_endowments = []
    if new_agents:
        pop.disease_endowments = randomize_disease_endowments(c, min(len(new_agents), int(c["startingDiseases"])), 0)
        random.shuffle(new_agents)
    if check and exact and layout.HostExt.needed(c):
        state.ext = layout.HostExt(capacity, c, len(pop.disease_endowments) if diseased else 0, horizon)
        if diseased:
            state.ext.arrays["diseases.immune_bits"][:] = steptables.immune_tables(1, horizon, int(c["agentImmuneSystemLength"]))
        if "misc" in state.ext.families:
            picks, draw = steptables.racial_tables(1, horizon, state.ext.racial_len)
            state.ext.arrays["misc.racial_picks"][:] = picks
            state.ext.arrays["misc.depression_draw"][:] = draw
    state.append_agents(new_agents)
    if diseased and state.ext is not None and new_agents:
        configure_diseases(c, state, pop.disease_endowments)
    state.counters[layout.C_NEXT_ID] = len(new_agents)
    state.counters[layout.C_ENDOW_INDEX] = pop.endowment_index
    return state, pop
