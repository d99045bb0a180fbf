"""Headless frame dump: what the reference's window shows (gui.py), written as one image per timestep.

The reference draws every cell of the map (gui.py:462-477, lookupFillColor 653-682): an empty cell in the colour of its
sugar / spice levels (findSugarAndSpiceColors 600-620: white -> yellow with sugar, -> brown with spice) or of its pollution,
an occupied one in the colour of the chosen agent attribute -- by default (no colour option picked) the "noSex" red, with the
"Tribes" option the palette entry of the agent's tribe, with "Sex" magenta / blue.  The GUI itself is out of scope (tkinter
needs a display); this module turns a downloaded state into the same picture, one pixel per cell times
`scale`, as binary PPM (no imaging library in the image).

    "b200FrameDump": "frames/", "b200FrameDumpEvery": 10, "b200FrameColor": "tribes"      (config keys of the drop-in driver)
"""
import os

import numpy as np

SUGAR, SPICE, WHITE = (0xF2, 0xFA, 0x00), (0x9B, 0x47, 0x22), (255, 255, 255)
POLLUTION_LOW, POLLUTION_HIGH = (0xFF, 0xFF, 0xFF), (0x80, 0x32, 0x80)
NO_SEX, FEMALE, MALE = (0xFA, 0x32, 0x32), (0xFA, 0x32, 0xFA), (0x32, 0x32, 0xFA)
PALETTE = [(0xFA, 0x32, 0x32), (0x32, 0x32, 0xFA), (0x32, 0xFA, 0x32), (0x32, 0xFA, 0xFA), (0xFA, 0x32, 0xFA), (0xAA, 0x32, 0x32),
           (0x32, 0x32, 0xAA), (0x32, 0xAA, 0x32), (0x32, 0xAA, 0xAA), (0xAA, 0x32, 0xAA), (0xFA, 0x88, 0x00), (0x00, 0xFA, 0x88),
           (0x88, 0x00, 0xFA), (0xFA, 0x88, 0x88), (0x88, 0x88, 0xFA), (0x88, 0xFA, 0x88), (0xFA, 0x32, 0x88), (0x32, 0x88, 0xFA),
           (0x88, 0xFA, 0x32), (0xAA, 0x66, 0xAA), (0x66, 0xAA, 0xAA), (0x3E, 0xD0, 0x6E), (0x6E, 0x3E, 0xD0), (0xD0, 0x6E, 0x3E),
           (0x00, 0x00, 0x00)]
AGENT_COLORINGS = ("default", "tribes", "sex")
ENVIRONMENT_COLORINGS = ("sugarAndSpice", "pollution")


def _mix(a, b, factor):
    """gui.py:642-644: int((1 - f) * start + f * end) per channel."""
    return tuple(int((1 - factor) * x + factor * y) for x, y in zip(a, b))


def sugar_and_spice_colors(max_sugar, max_spice):
    """gui.py:600-620 as a [max_sugar + 1, max_spice + 1, 3] table."""
    both = _mix(SUGAR, SPICE, 0.5)
    table = np.zeros((max_sugar + 1, max_spice + 1, 3), dtype=np.uint8)
    for sugar in range(max_sugar + 1):
        sf = sugar / max_sugar if max_sugar > 0 else 0
        for spice in range(max_spice + 1):
            pf = spice / max_spice if max_spice > 0 else 0
            table[sugar, spice] = _mix(_mix(WHITE, SPICE, pf), _mix(SUGAR, both, pf), sf)
    return table


def pollution_colors():
    """gui.py:16: 21 steps from white to purple, indexed by min(round(pollution), 20)."""
    return np.array([_mix(POLLUTION_LOW, POLLUTION_HIGH, i / 20) for i in range(21)], dtype=np.uint8)


def render(state, max_sugar, max_spice, agents="default", environment="sugarAndSpice"):
    """The map as an [height, width, 3] uint8 image (row = y, column = x, like the reference's canvas)."""
    if agents not in AGENT_COLORINGS or environment not in ENVIRONMENT_COLORINGS:
        raise ValueError(f"colourings: agents {AGENT_COLORINGS}, environment {ENVIRONMENT_COLORINGS}")
    W, H = state.width, state.height
    if environment == "pollution":
        level = np.minimum(np.round(np.asarray(state.pollution)), 20).astype(np.int64)       # (Python's round: half to even, like numpy's)
        img = pollution_colors()[level]
    else:
        table = sugar_and_spice_colors(max_sugar, max_spice)
        img = table[np.clip(state.sugar, 0, max_sugar), np.clip(state.spice, 0, max_spice)]
    img = img.copy()
    cells = np.nonzero(np.asarray(state.occ) >= 0)[0]
    slots = np.asarray(state.occ)[cells]
    if agents == "tribes":
        tribe = np.asarray(state.a_tribe)[slots]
        colors = np.array(PALETTE, dtype=np.uint8)[np.clip(tribe, 0, len(PALETTE) - 1)]
        colors[tribe < 0] = NO_SEX
    elif agents == "sex":
        from . import layout
        sex = np.asarray(state.a_sex)[slots]
        colors = np.where((sex == layout.SEX_FEMALE)[:, None], np.array(FEMALE, dtype=np.uint8),
                          np.where((sex == layout.SEX_MALE)[:, None], np.array(MALE, dtype=np.uint8), np.array(NO_SEX, dtype=np.uint8)))
    else:
        colors = np.broadcast_to(np.array(NO_SEX, dtype=np.uint8), (len(cells), 3))
    img[cells] = colors
    return img.reshape(W, H, 3).transpose(1, 0, 2)          # state arrays are x-major


def write_ppm(path, image, scale=1):
    """Binary PPM (P6): every image viewer and ffmpeg read it; no imaging library needed."""
    if scale > 1:
        image = np.repeat(np.repeat(image, scale, axis=0), scale, axis=1)
    h, w, _ = image.shape
    with open(path, "wb") as f:
        f.write(b"P6\n%d %d\n255\n" % (w, h))
        f.write(np.ascontiguousarray(image, dtype=np.uint8).tobytes())


class FrameDump:
    """Writes <directory>/frame_<timestep>.ppm every `every` timesteps from a state the driver downloads for it."""
    FIELDS = ["sugar", "spice", "pollution", "occ", "a_tribe", "a_sex"]

    def __init__(self, configuration):
        self.directory = configuration.get("b200FrameDump")
        self.every = max(int(configuration.get("b200FrameDumpEvery", 1)), 1)
        self.agents = configuration.get("b200FrameColor", "default")
        self.environment = configuration.get("b200FrameEnvironment", "sugarAndSpice")
        self.scale = max(int(configuration.get("b200FrameScale", 1)), 1)
        self.max_sugar = int(configuration["environmentMaxSugar"])
        self.max_spice = int(configuration["environmentMaxSpice"])
        self.written = 0
        if self.directory:
            os.makedirs(self.directory, exist_ok=True)

    def wanted(self, timestep):
        return bool(self.directory) and timestep % self.every == 0

    def write(self, timestep, state):
        image = render(state, self.max_sugar, self.max_spice, self.agents, self.environment)
        write_ppm(os.path.join(self.directory, "frame_%06d.ppm" % timestep), image, self.scale)
        self.written += 1
