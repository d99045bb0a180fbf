"""Host-side mirror of the reference's simulation driver (class Sugarscape, reference
sugarscape.py:18-1461) for the per-timestep path: same constructor argument (the verified
configuration dict), same public methods (``runSimulation``, ``doTimestep``, ``runtimeStats``,
log writers) and the same log.json format; the timestep itself runs on the GPU through libssb.

  doTimestep          sugarscape.py:389-421
  replaceDeadAgents   sugarscape.py:855-859   (host: it is the init path, Population.place)
  runSimulation       sugarscape.py:861-885
  startLog/writeToLog/endLog   sugarscape.py:887-905, 1428-1457, 423-466
"""
import json
import random

import numpy as np

from . import frames, init, layout, stats, steptables
from .engine import Engine


class Sugarscape:
    def __init__(self, configuration, rng_mode=None, device=None):
        c = self.configuration = configuration
        mode_name = rng_mode or c.get("b200RngMode", "exact")
        if mode_name not in ("exact", "scalable"):
            raise ValueError("b200RngMode must be 'exact' or 'scalable'")
        self.rng_mode = layout.RNG_EXACT if mode_name == "exact" else layout.RNG_SCALABLE
        self.maxTimestep = c["timesteps"]
        self.timestep = 0
        self.seed = c["seed"]
        self.keepAlive = c["keepAlivePostExtinction"]
        self.end = False
        if not c["headlessMode"]:
            raise init.UnsupportedConfiguration("the GUI is out of scope: set headlessMode to true")
        if self.rng_mode != layout.RNG_EXACT and (init.uses_decision_models(c) or layout.HostExt.needed(c, exact=False)):
            raise init.UnsupportedConfiguration("decision models, lending and friends need b200RngMode 'exact'")
        if self.rng_mode != layout.RNG_EXACT and c["agentInheritancePolicy"] != "none":
            # (the scalable transaction has no inheritance step: running on would silently change the wealth dynamics)
            raise init.UnsupportedConfiguration("agentInheritancePolicy needs b200RngMode 'exact'")
        state, self.population = init.build_initial_state(c, exact=self.rng_mode == layout.RNG_EXACT)
        self.params = layout.make_params(c, self.rng_mode, init.rule_flags(c), init.max_agent_range(c),
                                         init.equator(c))
        horizon = int(min(c["timesteps"], 1 << 20))
        endow, tags = steptables.step_tables(1, horizon, c["agentTagStringLength"])
        self.engine = Engine(self.params, state, device if device is not None else c.get("b200Device", 0),
                             endow, tags)
        if self.rng_mode == layout.RNG_EXACT:
            self.engine.set_mt_from_python()
        elif c["agentReplacements"] > 0:
            self.engine.bind_endowments(self.population.endowments)
        self.stats_builder = stats.StatsBuilder(c, init.equator(c))
        self.runtimeStats = self.stats_builder.runtime_stats
        self.n_agents = int(state.counters[layout.C_N_LIST])
        self.agent_steps = 0
        self.log = open(c["logfile"], "a") if c["logfile"] is not None else None
        self.logFormat = c["logfileFormat"]
        self.agentLog = open(c["agentLogfile"], "a") if c["agentLogfile"] is not None else None
        self.agentRuntimeStats = []
        # headless stand-in for the reference's window (gui.py): one image of the map per timestep (frames.py)
        self.frames = frames.FrameDump(c)
        if self.frames.wanted(0):
            self.frames.write(0, self.engine.download(frames.FrameDump.FIELDS))

    # -- one timestep --------------------------------------------------------------------------
    def doTimestep(self):
        if self.timestep >= self.maxTimestep:
            self.end = True
            return
        self.timestep += 1
        if self.end or (self.n_agents == 0 and not self.keepAlive):
            self.end = True
            return
        t, eng = self.timestep, self.engine
        self.agent_steps += self.n_agents
        eng.env_step(t)
        eng.agent_step(t, self.runtimeStats["meanWealth"])
        eng.compact(t)
        replaced = self.replaceDeadAgents()
        eng.reduce_stats(t)
        st = eng.stats()
        self.updateRuntimeStats(st, replaced)
        if self.agentLog is not None:
            self.agentRuntimeStats += [stats.agent_log_entry(r) for r in eng.agent_records()]
        if self.frames.wanted(t):
            self.frames.write(t, eng.download(frames.FrameDump.FIELDS))
        if self.timestep != self.maxTimestep and self.n_agents > 0:
            self.writeToLog()
            # the per-agent log starts once agents have acted (sugarscape.py:417-421): startLog writes the records of
            # step 1 and the regular write repeats them, like the reference
            if self.timestep == 1:
                self.startAgentLog()
            self.writeToAgentLog()
            self.agentRuntimeStats = []

    def replaceDeadAgents(self):
        """sugarscape.py:855-859.  Exact mode: placement of new agents is the init path (host,
        stdlib random on the main stream); the device hands over occupancy + MT state and receives
        the agents.  Scalable mode: on the device (ssb_replace), counted in the step's stats."""
        c = self.configuration
        if c["agentReplacements"] <= 0:
            return 0
        eng = self.engine
        if self.rng_mode == layout.RNG_SCALABLE:
            eng.replace(self.timestep)
            return None
        counters = eng.counters()
        n = int(counters[layout.C_N_LIST])
        if n >= c["agentReplacements"]:
            return 0
        hs = eng.download()
        eng.push_mt_to_python()
        new_agents = self.population.place(c["agentReplacements"] - n, hs.occ >= 0, self.timestep,
                                           int(hs.counters[layout.C_NEXT_ID]))
        hs.append_agents(new_agents, protect_last=int(hs.counters[layout.C_N_DEAD]))
        hs.counters[layout.C_NEXT_ID] += len(new_agents)
        eng.upload(hs, [k for k in eng.tensors if k not in ("pollution_flux",)])
        eng.set_mt_from_python()
        self.replaced_members = hs.newest_members(len(new_agents))       # (for the experimental-group split of the log)
        return len(new_agents)

    def updateRuntimeStats(self, st=None, replaced=0):
        if st is None:
            self.engine.reduce_stats(self.timestep)
            st = self.engine.stats()
        eng = self.engine
        replaced = int(st.n_replaced) if replaced is None else replaced
        groups = eng.group_stats()
        if groups is not None:      # the log repeats every statistic for the experimental group and for its complement
            self.stats_builder.update_with_groups(st, self.timestep, replaced, eng.ethics_stats(), eng.ext_stats(), groups,
                                                  getattr(self, "replaced_members", 0) if replaced else 0)
        else:
            self.stats_builder.update(st, self.timestep, replaced, eng.ethics_stats(), eng.ext_stats())
        self.n_agents = int(st.population)

    # -- the run loop and the log (format: SURVEY.md Appendix F) -----------------------------------
    def runSimulation(self, timesteps=5):
        self.startLog()
        self.updateRuntimeStats()
        t = 1
        timesteps = timesteps - self.timestep
        while t <= timesteps:
            if self.n_agents == 0 and not self.keepAlive:
                break
            self.doTimestep()
            t += 1
        self.endSimulation()

    def endSimulation(self):
        self.endLog()
        self.endAgentLog()
        self.engine.close()

    # -- the per-agent log (sugarscape.py:887-905, 423-466, 1428-1457 with log == self.agentLog) ---------------
    def _agent_rows(self, closing=False):
        if self.logFormat == "csv":
            return "".join(",".join(str(v) for v in entry.values()) + "\n" for entry in self.agentRuntimeStats)
        rows = [f"\t{json.dumps(entry)},\n" for entry in self.agentRuntimeStats]
        if closing and rows:          # (the reference closes the array after the last record: entries equal to it, too)
            last = self.agentRuntimeStats[-1]
            rows = [f"\t{json.dumps(e)}\n]" if e == last else f"\t{json.dumps(e)},\n" for e in self.agentRuntimeStats]
        return "".join(rows)

    def startAgentLog(self):
        if self.agentLog is None:
            return
        if self.logFormat == "csv":
            self.agentLog.write(",".join(self.agentRuntimeStats[0]) + "\n")
        else:
            self.agentLog.write("[\n")
        self.writeToAgentLog()

    def writeToAgentLog(self):
        if self.agentLog is not None:
            self.agentLog.write(self._agent_rows())

    def endAgentLog(self):
        if self.agentLog is None:
            return
        self.agentLog.write(self._agent_rows(closing=True))
        self.agentLog.flush()
        self.agentLog.close()
        self.agentLog = None

    def _csv_row(self):
        return ",".join(str(self.runtimeStats[k]) for k in sorted(self.runtimeStats)) + "\n"

    def startLog(self):
        if self.log is None:
            return
        if self.logFormat == "csv":
            self.log.write(",".join(sorted(self.runtimeStats)) + "\n")
        else:
            self.log.write("[\n")
        self.writeToLog()

    def writeToLog(self):
        if self.log is None:
            return
        if self.logFormat == "csv":
            self.log.write(self._csv_row())
        else:
            self.log.write(f"\t{json.dumps(self.runtimeStats)},\n")

    def endLog(self):
        if self.log is None:
            return
        # the reference recomputes the environment totals here (sugarscape.py:437-444): the
        # totals are current already; "created" loses the one-off capacity sum of t == 1
        sb = self.stats_builder
        self.runtimeStats["environmentWealthCreated"] = stats.environment_wealth_created(
            self.configuration, sb.equator, self.timestep, sb.n_north, sb.n_south, 0)
        if self.logFormat == "csv":
            self.log.write(self._csv_row())
        else:
            self.log.write(f"\t{json.dumps(self.runtimeStats)}\n]")
        self.log.flush()
        self.log.close()
        self.log = None
