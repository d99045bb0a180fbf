#!/usr/bin/env python3
"""K1 roofline: growback and pollution diffusion at the BASELINE map shapes, CUDA events around
ssb_env_step, algorithmic bytes of SURVEY.md section 8d (12 B/cell growback, +16 B/cell diffusion)."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from sugarscape_b200 import synthetic, layout
from sugarscape_b200.engine import Engine

peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
for W, H, label in ((4096, 4096, "S1 4096x4096"), (2048, 16384, "S2/8 slab 2048x16384"), (16384, 16384, "S2 16384x16384")):
    for diffusion in (False, True):
        ov = {"environmentWidth": W, "environmentHeight": H, "startingAgents": 0, "timesteps": 1 << 20}
        if diffusion:
            ov.update({"environmentPollutionDiffusionDelay": 1, "environmentPollutionDiffusionTimeframe": [0, 1 << 20],
                       "environmentPollutionTimeframe": [0, 1 << 20]})
        c = synthetic.configuration(synthetic.S1_OPTIONS, ov)
        state, params = synthetic.build_state(c, capacity=1024)
        state.sugar[:] = 0                                  # every cell grows: no skipped stores
        state.pollution[:] = np.random.default_rng(0).random(W * H)
        eng = Engine(params, state, 0)
        for t in range(1, 4):
            eng.env_step(t)
        torch.cuda.synchronize()
        times = []
        for t in range(4, 4 + 3):                            # levels reach capacity after 4 steps: time steps with stores
            state2 = None
        eng.tensors["sugar"].zero_()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        n = 3
        ev[0].record()
        for t in range(10, 10 + n):
            eng.env_step(t)
        ev[1].record()
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / n
        bytes_ = W * H * (12 + (16 if diffusion else 0))
        gbs = bytes_ / (ms / 1e3) / 1e9
        print(json.dumps({"map": label, "kernel": "growback" + ("+diffusion" if diffusion else ""), "ms": round(ms, 4),
                          "algorithmic_bytes": bytes_, "achieved_gbs": round(gbs, 1), "peak_gbs": peak, "frac": round(gbs / peak, 3)}))
        eng.close(); del eng; torch.cuda.empty_cache()
