#!/usr/bin/env python3
"""Fuzz the engine's mode E DEVICE SOURCE (host-compiled, tests/hostemu) against the live, UNMODIFIED reference
(build container only): random rule / size / seed combinations over every ABI 2 family -- decision models, lending,
friends, diseases, universal income, racial tags, depression -- and the book rules.

    python tests/golden/fuzz_hostemu.py [cases] [seed]

Same case generator and digests as fuzz_oracle.py, plus the 59 log keys of every step; configurations the engine
refuses are skipped."""
import os
import random
import sys
import traceback

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import ref_harness as rh  # noqa: E402
import parity_util as pu  # noqa: E402
from fuzz_oracle import BASES, draw_overrides  # noqa: E402
from hostemu_binding import HostEmu  # noqa: E402
from sugarscape_b200 import init, layout, stats  # noqa: E402


def _close(a, b):
    import math
    if a is None or b is None:
        return a is b
    if isinstance(a, float) or isinstance(b, float):
        return math.isclose(a, b, rel_tol=1e-9, abs_tol=1e-9)
    return a == b


def run_case(base, ov, steps):
    path = os.path.join(rh.REFERENCE_DIR, "examples", base + ".json")
    try:
        S = rh.build_reference(path, ov)
        S.updateRuntimeStats()
        ref_digests, ref_logs = [rh.digests(rh.snapshot(S))], [dict(S.runtimeStats)]
        for _ in range(steps):
            if len(S.agents) == 0 and not S.keepAlive:
                break
            S.doTimestep()
            ref_digests.append(rh.digests(rh.snapshot(S)))
            ref_logs.append(dict(S.runtimeStats))
    except Exception as e:
        return "skipped (the reference itself raised %r)" % (e,)
    try:
        c, state, pop, params, endow, tags = pu.build(base, overrides=ov)
    except init.UnsupportedConfiguration as e:
        return "skipped (refused: %s)" % e
    if c["agentReplacements"] > 0 and ((c["agentTagging"] and c["agentTagStringLength"] > 0 and not c["environmentTribePerQuadrant"])
                                       or (c["agentImmuneSystemLength"] > 0 and c["startingDiseases"] > 0)):
        return "skipped (reference aliases list-valued endowments between replacement twins)"
    emu = HostEmu(params, state, endow, tags)
    emu.set_mt_from_python()
    d, _ = pu.state_digests(state, *emu.mt_state())
    if d != ref_digests[0]:
        return "t=0 differs: " + str({k: d[k] == ref_digests[0][k] for k in d})
    sb = stats.StatsBuilder(c, init.equator(c))
    grouped = c["experimentalGroup"] is not None

    def entry(t, replaced):
        if grouped:
            return sb.update_with_groups(emu.stats(t), t, replaced, emu.ethics_stats(), emu.ext_stats(t), emu.group_stats(t),
                                         state.newest_members(replaced))
        return sb.update(emu.stats(t), t, replaced, emu.ethics_stats(), emu.ext_stats(t))

    rs = entry(0, 0)
    for t in range(1, len(ref_digests)):
        emu.env_step(t)
        emu.agent_step(t, rs["meanWealth"])
        emu.compact(t)
        if state.counters[layout.C_OVERFLOW] != 0:       # the engine stops such a run (ssb_stats fails with the code); nothing to compare
            return "skipped (the engine stops this run at t=%d: overflow code %d)" % (t, state.counters[layout.C_OVERFLOW])
        replaced = 0
        if c["agentReplacements"] > 0:
            emu.push_mt_to_python()
            replaced = pu.replace_dead_agents(c, state, pop, t)
            emu.set_mt_from_python()
        d, _ = pu.state_digests(state, *emu.mt_state())
        if d != ref_digests[t]:
            return f"t={t} differs: " + str({k: d[k] == ref_digests[t][k] for k in d})
        rs = entry(t, replaced)
        for k in (ref_logs[t] if grouped else stats.LOG_KEYS):          # the log keys of the step, too (203 with a group)
            if k.endswith(("eanMovement", "eanMetabolism", "gentTotalMetabolism")) and isinstance(ref_logs[t][k], (int, float)) \
                    and ref_logs[t][k] > 1e6:
                continue          # depressed lineages: the reference's endowments outgrow i32, the engine saturates
            if not _close(rs[k], ref_logs[t][k]):
                return f"t={t} log key {k} differs: {rs[k]} vs {ref_logs[t][k]}"
    families = sorted(state.ext.families) if state.ext is not None else []
    return "ok %d steps, final population %d, families %s%s" % (len(ref_digests) - 1, rs["population"], families,
                                                                " + ethics" if state.ethics is not None else "")


def main():
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    bad = 0
    for i in range(cases):
        base = rng.choice(BASES)
        ov = draw_overrides(rng, base)
        if "agentDecisionModels" in ov and rng.random() < 0.4:          # dynamic selfishness
            ov["agentDecisionModels"] = [m + rng.choice(["", "Dynamic"]) + (rng.choice(["", "Top"]) if "negative" not in m else "")
                                         if m != "none" else m for m in ov["agentDecisionModels"]]
            ov["agentDynamicSelfishnessFactor"] = rng.choice([[0.0, 0.0], [0.02, 0.05]])
        if "agentDecisionModels" in ov and rng.random() < 0.3:          # Temperance agents among them
            ov["agentDecisionModels"] = ov["agentDecisionModels"] + [rng.choice(["temperance", "temperancePECS"])]
            ov["agentDynamicSocialPressureFactor"] = rng.choice([[0.0, 0.0], [0.02, 0.1]])
            ov["agentDecisionModelFactor"] = rng.choice([[1, 1], [0.2, 0.8]])
            ov["agentDynamicDecisionModelFactor"] = rng.choice([[0.0, 0.0], [0.05, 0.15]])
        if rng.random() < 0.3 or os.environ.get("SSB_FUZZ_GROUPS") == "1":      # SSB_FUZZ_GROUPS=1: every case with a group
            ov["experimentalGroup"] = rng.choice(["depressed", "female", "male", "sick", "sick", "race0", "race1", "race2"] +
                                                 [m for m in ov.get("agentDecisionModels", []) if m != "none"])
            if rng.random() < 0.5:
                ov["agentReplacements"] = 0
        try:
            res = run_case(base, ov, 60)
        except Exception as e:
            res = "EXCEPTION " + repr(e) + "\n" + traceback.format_exc(limit=6)
        flag = "" if res.startswith("ok") or res.startswith("skipped") else "   <<<<<<"
        print(f"case {i} {base}: {res}{flag}", flush=True)
        if flag:
            bad += 1
            print("   overrides:", ov, flush=True)
    print("mismatching cases:", bad)


if __name__ == "__main__":
    main()
