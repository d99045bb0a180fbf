"""Shared helpers of the parity tests: build a simulation from a shipped example config, step it
with the CPU oracle (or the CUDA engine) and compare with the golden fixtures."""
import copy
import gzip
import json
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
with open(os.path.join(GOLDEN, "example_configs.json")) as _f:
    EXAMPLE_OPTIONS = json.load(_f)
sys.path.insert(0, ROOT)
sys.path.insert(0, GOLDEN)

from sugarscape_b200 import config, init, layout, steptables  # noqa: E402
import ref_harness  # noqa: E402  (digest helpers only; the reference itself is not imported)


def load_golden(name):
    with gzip.open(os.path.join(GOLDEN, name + ".json.gz"), "rt") as f:
        meta = json.load(f)
    snaps = np.load(os.path.join(GOLDEN, name + ".npz"))
    return meta, snaps


def load_variant(name):
    """A fixture of tests/golden/make_golden_variants.py: (base example, overrides, steps, log)."""
    with gzip.open(os.path.join(GOLDEN, name + ".json.gz"), "rt") as f:
        meta = json.load(f)
    overrides = dict(meta["overrides"], timesteps=meta["meta"]["timesteps"])
    if overrides.get("environmentFile"):
        overrides["environmentFile"] = os.path.join(GOLDEN, overrides["environmentFile"])
    return meta["base"], overrides, meta["steps"], meta["log"]


def example_configuration(name, overrides=None):
    """tests/test.py protocol of the reference (headless, no debug, 200 steps)."""
    # deep copies: verify_configuration rewrites nested lists in place (out-of-range peaks, sorted ranges)
    c = config.overlay_options(copy.deepcopy(EXAMPLE_OPTIONS[name]), config.default_configuration())
    c["headlessMode"] = True
    c["debugMode"] = ["none"]
    c["logfile"] = None
    c["timesteps"] = 200
    for k, v in (overrides or {}).items():
        c[k] = copy.deepcopy(v)
    config.verify_random_seed(c)
    return config.verify_configuration(c)


def build(name, rng_mode=layout.RNG_EXACT, overrides=None, check=True):
    c = example_configuration(name, overrides)
    state, pop = init.build_initial_state(c, check=check, exact=rng_mode == layout.RNG_EXACT)
    flags = init.rule_flags(c)
    params = layout.make_params(c, rng_mode, flags, init.max_agent_range(c), init.equator(c))
    endow, tags = steptables.step_tables(1, min(c["timesteps"], 100000), c["agentTagStringLength"])
    return c, state, pop, params, endow, tags


def state_digests(state, mt_words=None, mt_index=None):
    snap = state.canonical()
    snap["a_alive"] = np.ones_like(snap["a_id"])
    if mt_words is None:
        mt_words, mt_index = np.zeros(624, np.uint32), 0
    snap["mt"] = np.concatenate([np.asarray(mt_words, dtype=np.uint32), np.array([mt_index], dtype=np.uint32)])
    return ref_harness.digests(snap), snap


def replace_dead_agents(c, state, pop, t):
    """replaceDeadAgents (reference sugarscape.py:855-859) on a HostState; uses the stdlib
    main stream, which the caller has synchronised with the stepping engine."""
    n = int(state.counters[layout.C_N_LIST])
    if n >= c["agentReplacements"]:
        return 0
    new_agents = pop.place(c["agentReplacements"] - n, state.occ >= 0, t, int(state.counters[layout.C_NEXT_ID]))
    state.append_agents(new_agents, protect_last=int(state.counters[layout.C_N_DEAD]))
    state.counters[layout.C_NEXT_ID] += len(new_agents)
    return len(new_agents)


def state_digests_by_id(state):
    """Digests of the state with the agents sorted by ID (list order is not part of the
    scalable-mode semantics)."""
    snap = state.canonical()
    order = np.argsort(snap["a_id"], kind="stable")
    for k in [k for k in snap if k.startswith("a_")]:
        snap[k] = snap[k][order]
    snap["a_alive"] = np.ones_like(snap["a_id"])
    snap["mt"] = np.zeros(625, np.uint32)
    return ref_harness.digests(snap), snap


def load_scalable_golden():
    with gzip.open(os.path.join(GOLDEN, "scalable_mode.json.gz"), "rt") as f:
        return json.load(f)
