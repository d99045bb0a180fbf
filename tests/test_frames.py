"""sugarscape_b200/frames.py: the headless frame dump (stand-in for the reference's tkinter window, gui.py:462-477, 600-620,
653-682) -- host code on a downloaded state, so it is tested on the initial state without a GPU."""
import numpy as np
import pytest

import parity_util as pu
from sugarscape_b200 import frames, layout


def _hex(rgb):
    return "#%02X%02X%02X" % tuple(int(v) for v in rgb)


def test_sugar_and_spice_table_follows_the_gui_formula():
    table = frames.sugar_and_spice_colors(4, 4)
    assert _hex(table[0, 0]) == "#FFFFFF"                 # barren cell: white
    assert _hex(table[4, 0]) == "#F2FA00"                 # full sugar, no spice: the sugar yellow
    assert _hex(table[0, 4]) == "#9B4722"                 # full spice, no sugar: the spice brown
    # int((1 - f) * a + f * b) per channel, twice (gui.py:613-616): spot value computed by hand for (2, 0)
    assert tuple(table[2, 0]) == (int(0.5 * 255 + 0.5 * 0xF2), int(0.5 * 255 + 0.5 * 0xFA), int(0.5 * 255))
    assert frames.sugar_and_spice_colors(4, 0).shape == (5, 1, 3)
    assert _hex(frames.pollution_colors()[0]) == "#FFFFFF" and _hex(frames.pollution_colors()[20]) == "#803280"


@pytest.mark.parametrize("coloring", frames.AGENT_COLORINGS)
def test_render_marks_every_agent_and_colours_every_empty_cell(coloring):
    c, state, pop, params, endow, tags = pu.build("combat_limited")
    img = frames.render(state, c["environmentMaxSugar"], c["environmentMaxSpice"], agents=coloring)
    W, H = state.width, state.height
    assert img.shape == (H, W, 3) and img.dtype == np.uint8
    table = frames.sugar_and_spice_colors(c["environmentMaxSugar"], c["environmentMaxSpice"])
    n = int(state.counters[layout.C_N_LIST])
    for s in state.a_order[:min(n, 50)]:
        x, y = divmod(int(state.a_pos[s]), H)
        want = {"default": frames.NO_SEX, "tribes": frames.PALETTE[int(state.a_tribe[s])] if state.a_tribe[s] >= 0 else frames.NO_SEX,
                "sex": frames.FEMALE if state.a_sex[s] == layout.SEX_FEMALE else frames.MALE if state.a_sex[s] == layout.SEX_MALE else frames.NO_SEX}[coloring]
        assert tuple(img[y, x]) == tuple(want)
    empty = np.nonzero(state.occ < 0)[0][:200]
    for cell in empty:
        x, y = divmod(int(cell), H)
        assert tuple(img[y, x]) == tuple(table[state.sugar[cell], state.spice[cell]])
    assert int((state.occ >= 0).sum()) == n


def test_frame_dump_writes_ppm_files(tmp_path):
    c, state, pop, params, endow, tags = pu.build("pollution_formation")
    c.update({"b200FrameDump": str(tmp_path / "frames"), "b200FrameDumpEvery": 5, "b200FrameScale": 3, "b200FrameEnvironment": "pollution",
              "b200FrameColor": "tribes"})
    dump = frames.FrameDump(c)
    assert dump.wanted(0) and dump.wanted(10) and not dump.wanted(7)
    state.pollution[:] = np.linspace(0, 30, state.pollution.size)
    dump.write(10, state)
    data = (tmp_path / "frames" / "frame_000010.ppm").read_bytes()
    header = b"P6\n%d %d\n255\n" % (3 * state.width, 3 * state.height)
    assert data.startswith(header) and len(data) == len(header) + 9 * 3 * state.width * state.height
    assert not frames.FrameDump(pu.example_configuration("constant_growback")).wanted(0)      # off by default
    with pytest.raises(ValueError):
        frames.render(state, 4, 4, agents="vision")


def test_drop_in_driver_dumps_frames(tmp_path, monkeypatch):
    """The driver's hook (Sugarscape.doTimestep): frames at t = 0, 4, 8, 12 of a 12-step run.  No GPU here: the engine behind
    the driver is replaced by the test adapter over the host-compiled device source (tools/run_gpu_tests_on_hostemu.py)."""
    import os
    import sys
    sys.path.insert(0, os.path.join(pu.ROOT, "tools"))
    import sugarscape_b200.simulation as sim
    from run_gpu_tests_on_hostemu import HostEmuEngine
    monkeypatch.setattr(sim, "Engine", HostEmuEngine)
    c = pu.example_configuration("cultural_tagging", {"timesteps": 12})
    c.update({"logfile": None, "b200FrameDump": str(tmp_path), "b200FrameDumpEvery": 4, "b200FrameColor": "tribes", "b200FrameScale": 2})
    S = sim.Sugarscape(c)
    S.runSimulation(c["timesteps"])
    names = sorted(p.name for p in tmp_path.iterdir())
    assert names == ["frame_000000.ppm", "frame_000004.ppm", "frame_000008.ppm", "frame_000012.ppm"] and S.frames.written == 4
    first, last = (tmp_path / names[0]).read_bytes(), (tmp_path / names[-1]).read_bytes()
    assert len(first) == len(last) and first != last          # the agents moved
