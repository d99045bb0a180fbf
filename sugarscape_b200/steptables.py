"""Per-timestep tables for child endowments (host side, cold path).

The reference draws a child's inherited endowments from private, hash-seeded streams that are
a pure function of (endowment name, timestep): ``random.seed(md5(name) + timestep)`` followed by
``randrange(2)`` (reference agent.py:827-851), one shared draw for the twelve decision-model
keys (845-851), and a run of ``randrange(2)`` draws from ``md5("tags") + timestep`` for the tag
positions where the parents differ (859-875).  The main stream is saved and restored around
them (825, 909), so the host can precompute one small table per timestep and hand it to the
device (``ssb_set_step_tables``) -- SURVEY.md Appendix A.9.
"""
import hashlib
import random

import numpy as np

# bit positions inside endow_bits[t]; must match enum EB_* in csrc/ssb_rules.cuh and the oracle
# (bits 12-19 are read by the oracle-only restatements -- lending, friends, diseases, universal income, ethics -- the
# engine has none of them yet)
ENDOWMENT_BITS = ("aggressionFactor", "fertilityAge", "fertilityFactor", "infertilityAge",
                  "lookaheadFactor", "maxAge", "movement", "spiceMetabolism", "sugarMetabolism",
                  "sex", "tradeFactor", "vision", "baseInterestRate", "lendingFactor", "loanDuration", "maxFriends",
                  "diseaseProtectionChance", "universalSpice", "universalSugar", "decisionModel")


def _md5_int(name):
    return int(hashlib.md5(name.encode()).hexdigest(), 16)


_HASH = {name: _md5_int(name) for name in ENDOWMENT_BITS + ("tags", "immuneSystem", "racialTags")}


def racial_tables(t_first, t_last, racial_len):
    """(picks[t], depression[t]): the private stream md5("racialTags") + timestep first picks, per
    racial tag position, the parent the child's tag comes from (random.choice of two, reference
    agent.py:879-887; bit i of picks = index picked), then yields the random.random() the child's
    depression is decided with (889-895).  Without racial tags only the depression draw is taken."""
    n = t_last + 1
    picks = np.zeros(n, dtype=np.uint32)
    depression = np.ones(n, dtype=np.float64)
    rng = random.Random()
    for t in range(max(t_first, 0), n):
        rng.seed(_HASH["racialTags"] + t)
        bits = 0
        for i in range(racial_len):
            bits |= rng.choice([0, 1]) << i
        picks[t] = bits
        depression[t] = rng.random()
    return picks, depression


def immune_tables(t_first, t_last, immune_len):
    """immune_bits[t]: the randrange(2) draws for the positions where the parents' immune systems
    differ (reference agent.py:897-907, private stream md5("immuneSystem") + timestep)."""
    n = t_last + 1
    out = np.zeros(n, dtype=np.uint64)
    rng = random.Random()
    for t in range(max(t_first, 0), n):
        rng.seed(_HASH["immuneSystem"] + t)
        bits = 0
        for i in range(min(immune_len, 64)):
            bits |= rng.randrange(2) << i
        out[t] = bits
    return out


def step_tables(t_first, t_last, tag_len):
    """endow_bits[t], tag_bits[t] for t in [0, t_last]; entries below t_first are zero."""
    n = t_last + 1
    endow = np.zeros(n, dtype=np.uint32)
    tags = np.zeros(n, dtype=np.uint32)
    rng = random.Random()
    for t in range(max(t_first, 0), n):
        bits = 0
        for i, name in enumerate(ENDOWMENT_BITS):
            rng.seed(_HASH[name] + t)
            bits |= rng.randrange(2) << i
        endow[t] = bits
        if tag_len > 0:
            rng.seed(_HASH["tags"] + t)
            tb = 0
            for i in range(tag_len):
                tb |= rng.randrange(2) << i
            tags[t] = tb
    return endow, tags
