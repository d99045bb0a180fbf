"""Configuration schema of the drop-in entry point (``sugarscape.py --conf X``).

Keeps the reference's ``--conf`` JSON schema: the 111 option names and default values of
reference sugarscape.py:1843-1954, the overlay rules of ``parseConfiguration``
(sugarscape.py:1463-1480: flat file or ``"sugarscapeOptions"`` object, unknown keys ignored,
legacy ``agentEthicalTheory`` / ``agentEthicalFactor`` aliases) and the normalisation of
``verifyConfiguration`` (sugarscape.py:1532-1834) including the draws it may take from the
already-seeded main random stream.  Host-side, cold path.
"""
import json
import random
import re
import sys

# --- option groups (name -> default) ---------------------------------------------------------
_AGENT_RANGES = {
    "agentAggressionFactor": [0, 0], "agentBaseInterestRate": [0.0, 0.0],
    "agentDecisionModelAgeismFactor": [-1, -1], "agentDecisionModelFactor": [0, 0],
    "agentDecisionModelLookaheadDiscount": [0, 0], "agentDecisionModelLookaheadFactor": [0],
    "agentDecisionModelRacismFactor": [-1, -1], "agentDecisionModelSexismFactor": [-1, -1],
    "agentDecisionModelTribalFactor": [-1, -1], "agentDiseaseProtectionChance": [0.0, 0.0],
    "agentDynamicDecisionModelFactor": [0.0, 0.0], "agentDynamicSelfishnessFactor": [0.0, 0.0],
    "agentDynamicSocialPressureFactor": [0, 1.0], "agentFemaleInfertilityAge": [0, 0],
    "agentFemaleFertilityAge": [0, 0], "agentFertilityFactor": [0, 0],
    "agentLendingFactor": [0, 0], "agentLoanDuration": [0, 0], "agentLookaheadFactor": [0, 0],
    "agentMaleInfertilityAge": [0, 0], "agentMaleFertilityAge": [0, 0], "agentMaxAge": [-1, -1],
    "agentMaxFriends": [0, 0], "agentMovement": [1, 6], "agentSelfishnessFactor": [-1, -1],
    "agentSpiceMetabolism": [0, 0], "agentStartingSpice": [0, 0], "agentStartingSugar": [10, 40],
    "agentSugarMetabolism": [1, 4], "agentTemperanceFactor": [0, 0], "agentTradeFactor": [0, 0],
    "agentUniversalSpice": [0, 0], "agentUniversalSugar": [0, 0], "agentVision": [1, 6],
}
_AGENT_SCALARS = {
    "agentDecisionModels": ["none"], "agentDecisionModel": None, "agentDepressionPercentage": 0,
    "agentImmuneSystemLength": 0, "agentInheritancePolicy": "none", "agentLeader": False,
    "agentLogfile": None, "agentMaleToFemaleRatio": 1.0, "agentMovementMode": "cardinal",
    "agentRacialTagStringLength": 0, "agentReplacements": 0, "agentTagging": False,
    "agentTagPreferences": False, "agentTagStringLength": 0, "agentVisionMode": "cardinal",
}
_DISEASE = {
    "diseaseAggressionPenalty": [0, 0], "diseaseFertilityPenalty": [0, 0],
    "diseaseFriendlinessPenalty": [0, 0], "diseaseHappinessPenalty": [0, 0],
    "diseaseIncubationPeriod": [0, 0], "diseaseList": [], "diseaseMovementPenalty": [0, 0],
    "diseaseSpiceMetabolismPenalty": [0, 0], "diseaseSugarMetabolismPenalty": [0, 0],
    "diseaseTagStringLength": [0, 0], "diseaseTimeframe": [0, 0],
    "diseaseTransmissionChance": [1.0, 1.0], "diseaseVisionPenalty": [0, 0],
    "startingDiseases": 0, "startingDiseasesPerAgent": [0, 0],
}
_ENVIRONMENT = {
    "environmentAgeistAbsoluteRanges": [], "environmentAgeistRelativeRange": -1,
    "environmentEquator": -1, "environmentFile": None, "environmentHeight": 50,
    "environmentInGroupRaces": [], "environmentMaxCombatLoot": 0, "environmentMaxRaces": 0,
    "environmentMaxSpice": 0, "environmentMaxSugar": 4, "environmentMaxTribes": 0,
    "environmentPollutionDiffusionDelay": 0, "environmentPollutionDiffusionTimeframe": [0, 0],
    "environmentPollutionTimeframe": [0, 0], "environmentQuadrantSizeFactor": 1,
    "environmentSeasonalGrowbackDelay": 0, "environmentSeasonInterval": 0,
    "environmentSexistGroups": [], "environmentSpiceConsumptionPollutionFactor": 0,
    "environmentSpicePeaks": [[35, 35, 4], [15, 15, 4]],
    "environmentSpiceProductionPollutionFactor": 0, "environmentSpiceRegrowRate": 0,
    "environmentStartingQuadrants": [1, 2, 3, 4],
    "environmentSugarConsumptionPollutionFactor": 0,
    "environmentSugarPeaks": [[35, 15, 4], [15, 35, 4]],
    "environmentSugarProductionPollutionFactor": 0, "environmentSugarRegrowRate": 1,
    "environmentTribePerQuadrant": False, "environmentUniversalSpiceIncomeInterval": 0,
    "environmentUniversalSugarIncomeInterval": 0, "environmentWidth": 50,
    "environmentWraparound": True,
}
_RUN = {
    "debugMode": ["none"], "experimentalGroup": None, "headlessMode": False,
    "interfaceHeight": 1000, "interfaceWidth": 900, "keepAlivePostExtinction": False,
    "keepAliveAtEnd": False, "logfile": None, "logfileFormat": "json",
    "neighborhoodMode": "vonNeumann", "profileMode": False, "screenshots": False, "seed": -1,
    "startingAgents": 250, "timesteps": 200,
}
# Options of this engine only (never present in reference configs; ignored by the reference
# because it only honours keys of its own defaults dict, sugarscape.py:1477-1479).
ENGINE_OPTIONS = {
    "b200RngMode": "exact",      # "exact" (mode E) | "scalable" (mode S)
    "b200Device": 0,
    "b200StatsEvery": 1,         # reduce + log every k-th step (1 = reference behaviour)
    # headless stand-in for the reference's window (sugarscape_b200/frames.py): one PPM image of the map per timestep
    "b200FrameDump": None,       # directory (None: no frames)
    "b200FrameDumpEvery": 1,     # every k-th timestep
    "b200FrameColor": "default",             # agents: "default" (the GUI's red) | "tribes" | "sex"
    "b200FrameEnvironment": "sugarAndSpice", # empty cells: "sugarAndSpice" | "pollution"
    "b200FrameScale": 1,         # pixels per cell side
}


def default_configuration():
    """A fresh deep copy of the reference defaults plus this engine's own options."""
    import copy
    merged = {}
    for group in (_AGENT_RANGES, _AGENT_SCALARS, _DISEASE, _ENVIRONMENT, _RUN):
        merged.update(group)
    conf = {k: copy.deepcopy(merged[k]) for k in sorted(merged)}
    assert len(conf) == 111, len(conf)
    return conf


def parse_configuration(config_file, configuration):
    """sugarscape.py:1463-1480."""
    with open(config_file) as f:
        options = json.loads(f.read())
    return overlay_options(options, configuration)


def overlay_options(options, configuration):
    """The overlay half of parseConfiguration (sugarscape.py:1466-1480) on a parsed dict."""
    if "sugarscapeOptions" in options:
        options = options["sugarscapeOptions"]
    if "agentEthicalTheory" in options:
        options["agentDecisionModel"] = options["agentEthicalTheory"]
    if "agentEthicalFactor" in options:
        options["agentDecisionModelFactor"] = options["agentEthicalFactor"]
    for opt in configuration:
        if opt in options:
            configuration[opt] = options[opt]
    for opt, default in ENGINE_OPTIONS.items():
        configuration[opt] = options.get(opt, configuration.get(opt, default))
    return configuration


def verify_random_seed(configuration):
    """sugarscape.py:1836-1839 -- the ONE seeding of the main stream."""
    if type(configuration["seed"]) != int or configuration["seed"] == -1:
        configuration["seed"] = random.randrange(sys.maxsize)
    random.seed(configuration["seed"])


class ConfigurationError(SystemExit):
    pass


def _normalise_timeframe(configuration, name):
    """sugarscape.py:1507-1530."""
    start, end = configuration[name][0], configuration[name][1]
    if start > end and end >= 0:
        start, end = end, start
    if start < 0:
        start = 0
    if end < 0:
        end = configuration["timesteps"]
    return [start, end]


_NEGATIVES_ALLOWED = {
    "agentDecisionModelAgeismFactor", "agentDecisionModelRacismFactor",
    "agentDecisionModelSexismFactor", "agentDecisionModelTribalFactor", "agentMaxAge",
    "agentSelfishnessFactor", "diseaseAggressionPenalty", "diseaseFertilityPenalty",
    "diseaseFriendlinessPenalty", "diseaseHappinessPenalty", "diseaseMovementPenalty",
    "diseaseSpiceMetabolismPenalty", "diseaseSugarMetabolismPenalty", "diseaseTimeframe",
    "diseaseVisionPenalty", "environmentAgeistAbsoluteRanges", "environmentAgeistRelativeRange",
    "environmentEquator", "environmentPollutionDiffusionTimeframe",
    "environmentPollutionTimeframe", "environmentMaxSpice", "environmentMaxSugar",
    "interfaceHeight", "interfaceWidth", "seed", "timesteps",
}
_TIMEFRAMES = ("diseaseTimeframe", "environmentPollutionDiffusionTimeframe",
               "environmentPollutionTimeframe")
_FACTOR_RANGES = ("agentDecisionModelAgeismFactor", "agentDecisionModelRacismFactor",
                  "agentDecisionModelSexismFactor", "agentDecisionModelTribalFactor")
_DEBUG_MODES = ("agent", "all", "cell", "disease", "environment", "ethics", "none", "sugarscape")
HELP_TEXT = ("Usage:\n\tpython sugarscape.py --conf config.json\n\nOptions:\n\t-c,--conf\tUse "
             "specified config file for simulation settings.\n\t-h,--help\tDisplay this message.")


def verify_configuration(c):
    """Normalise the configuration exactly as the reference does (sugarscape.py:1532-1834),
    in the same order, so that any draw from the main random stream lands on the same word."""
    # 1541-1553: every list option is sorted in place; negatives clamped unless allowed
    for name, value in c.items():
        if not isinstance(value, list) or len(value) == 0:
            continue
        first_type = type(value[0])
        if name in _TIMEFRAMES:
            c[name] = _normalise_timeframe(c, name)
            value = c[name]
        else:
            value.sort()
        if name not in _NEGATIVES_ALLOWED and first_type in (int, float):
            for i in range(len(value)):
                if value[i] < 0:
                    value[i] = 0

    # 1563-1579: generic experimental groups
    group = c["experimentalGroup"]
    if group is not None:
        for word, generic in (("ageRange", "ageInGroup"), ("disease", "sick"), ("race", "raceInGroup")):
            group = c["experimentalGroup"]
            if group is not None and word in group and re.search(word + r"(?P<ID>\d+)", group) is None:
                c["experimentalGroup"] = generic

    # 1581-1606: random maxima / peaks (main-stream draws)
    if c["environmentMaxSpice"] < 0:
        c["environmentMaxSpice"] = random.randint(1, 10)
    if c["environmentMaxSugar"] < 0:
        c["environmentMaxSugar"] = random.randint(1, 10)
    for peaks, gmax, spice in ((c["environmentSpicePeaks"], c["environmentMaxSpice"], True),
                               (c["environmentSugarPeaks"], c["environmentMaxSugar"], False)):
        for peak in peaks:
            appended = len(peak) < 3
            if appended:
                peak.append(gmax)
            if peak[0] < 0 or peak[0] > c["environmentWidth"]:
                peak[0] = random.randint(0, c["environmentWidth"] - 1)
            if peak[1] < 0 or peak[1] > c["environmentHeight"]:
                peak[1] = random.randint(0, c["environmentHeight"] - 1)
            if peak[2] < 0:
                peak[2] = random.randint(1, gmax)
            elif peak[2] > gmax:
                peak[2] = gmax

    # 1608-1625
    if c["environmentQuadrantSizeFactor"] > 1 or c["environmentQuadrantSizeFactor"] < 0:
        c["environmentQuadrantSizeFactor"] = 1
    if len(c["environmentStartingQuadrants"]) == 0:
        c["environmentStartingQuadrants"] = [1, 2, 3, 4]
    if c["environmentTribePerQuadrant"] == True:
        c["environmentMaxTribes"] = len(c["environmentStartingQuadrants"])
    total_cells = c["environmentHeight"] * c["environmentWidth"]
    total_cells = total_cells * (c["environmentQuadrantSizeFactor"] ** 2) * len(c["environmentStartingQuadrants"]) / 4
    if c["startingAgents"] > total_cells:
        c["startingAgents"] = total_cells

    # 1628-1631
    if c["timesteps"] < 0:
        c["timesteps"] = sys.maxsize

    # 1634-1708: bias / selfishness / dynamic factors
    for name in _FACTOR_RANGES:
        if c[name][0] < 0:
            c[name] = [-1, -1]
        elif c[name][1] > 1:
            c[name][1] = 1
    if c["agentMaxAge"][0] < 0:
        c["agentMaxAge"] = [-1, -1]
    if c["agentSelfishnessFactor"][0] < 0:
        c["agentSelfishnessFactor"] = [-1, -1]
    elif c["agentSelfishnessFactor"][1] > 1:
        c["agentSelfishnessFactor"][1] = 1
    for name in ("agentDynamicDecisionModelFactor", "agentDynamicSocialPressureFactor"):
        if c[name][0] < 0:
            c[name] = [-1, -1]
        elif c[name][1] > 1:
            c[name][1] = 1.0

    # 1710-1754: tag lengths, tribes, races, colour cap
    for name in ("agentRacialTagStringLength", "agentTagStringLength", "environmentMaxRaces",
                 "environmentMaxTribes"):
        if c[name] < 0:
            c[name] = 0
    if c["agentTagStringLength"] > 0 and c["environmentMaxTribes"] > c["agentTagStringLength"]:
        c["environmentMaxTribes"] = c["agentTagStringLength"]
    max_colors = 25
    if c["environmentMaxRaces"] > max_colors:
        c["environmentMaxRaces"] = max_colors
    if c["environmentMaxTribes"] > max_colors:
        c["environmentMaxTribes"] = max_colors

    # 1756-1779: ageism ranges, in-group races
    for age_range in c["environmentAgeistAbsoluteRanges"][:]:
        lo, hi = age_range
        if lo < -1 or hi < -1 or (lo > hi and hi != -1):
            c["environmentAgeistAbsoluteRanges"].remove(age_range)
    if c["environmentAgeistRelativeRange"] < 0 and c["environmentAgeistRelativeRange"] != -1:
        c["environmentAgeistRelativeRange"] = -1
    if (c["environmentAgeistAbsoluteRanges"] == [] and c["environmentAgeistRelativeRange"] == -1
            and c["agentDecisionModelAgeismFactor"] != [-1, -1]):
        c["agentDecisionModelAgeismFactor"] = [-1, -1]
    if any(r >= c["environmentMaxRaces"] for r in c["environmentInGroupRaces"]):
        c["environmentInGroupRaces"] = [r for r in c["environmentInGroupRaces"] if r < c["environmentMaxRaces"]]

    # 1781-1789
    if c["startingDiseasesPerAgent"] != [0, 0]:
        per_agent = sorted(c["startingDiseasesPerAgent"])
        per_agent = [max(0, n) for n in per_agent]
        per_agent = [min(c["startingDiseases"], n) for n in per_agent]
        c["startingDiseasesPerAgent"] = per_agent

    # 1791-1812
    if c["logfile"] == "":
        c["logfile"] = None
    if c["agentLogfile"] == "":
        c["agentLogfile"] = None
    bad = [m for m in c["debugMode"] if m not in _DEBUG_MODES]
    for m in bad:
        print(f"Debug mode {m} not recognized")
    if bad:
        print(HELP_TEXT)
        raise ConfigurationError(0)
    if "all" in c["debugMode"] and "none" in c["debugMode"]:
        print("Cannot have \"all\" and \"none\" debug modes enabled at the same time")
        print(HELP_TEXT)
        raise ConfigurationError(0)
    elif "all" in c["debugMode"] and len(c["debugMode"]) > 1:
        c["debugMode"] = "all"
    elif "none" in c["debugMode"] and len(c["debugMode"]) > 1:
        c["debugMode"] = "none"

    # 1814-1820: legacy single decision model
    if c["agentDecisionModel"] is not None and type(c["agentDecisionModel"]) == str:
        c["agentDecisionModels"] = [c["agentDecisionModel"]]
    elif c["agentDecisionModel"] is not None and type(c["agentDecisionModel"]) == list:
        c["agentDecisionModels"] = c["agentDecisionModel"]
    if type(c["agentDecisionModels"]) == str:
        c["agentDecisionModels"] = [c["agentDecisionModels"]]

    # 1822-1833
    if c["experimentalGroup"] == "":
        c["experimentalGroup"] = None
    groups = ["ageInGroup", "depressed", "female", "male", "raceInGroup", "sick"] + list(c["agentDecisionModels"])
    g = c["experimentalGroup"]
    if g is not None and g not in groups and "ageRange" not in g and "disease" not in g and "race" not in g:
        c["experimentalGroup"] = None
    return c
