"""Import harness around the UNMODIFIED reference (nkremerh/sugarscape, /root/reference).

TEST INFRASTRUCTURE ONLY.  Used by ``make_golden.py`` (run in the build container, where
``/root/reference`` exists) to produce the committed fixtures under ``tests/golden/``.
Nothing in the product (``sugarscape_b200/``), ``bench.py`` or the ``-m gpu`` tests imports it:
the reference does not exist on the GPU box.

Recipe follows SURVEY.md Appendix E:
  * ``sugarscape.py`` is importable because its driver lives under ``__main__``
    (reference sugarscape.py:1841); the defaults dict is a literal inside that block
    (sugarscape.py:1843-1954) and is recovered with ``ast.literal_eval``.
  * build = parseConfiguration -> verifyRandomSeed -> verifyConfiguration -> Sugarscape(conf)
    -> updateRuntimeStats()  (mirrors runSimulation, sugarscape.py:861-863).
"""
import ast
import hashlib
import os
import random
import struct
import sys

import numpy as np

REFERENCE_DIR = os.environ.get("SUGARSCAPE_REFERENCE", "/root/reference")


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_DIR, "sugarscape.py"))


def import_reference():
    if REFERENCE_DIR not in sys.path:
        sys.path.insert(0, REFERENCE_DIR)
    import sugarscape as ref  # noqa: the reference's module, not this repo's CLI
    if not hasattr(ref, "Sugarscape"):
        raise ImportError("imported a sugarscape module that is not the reference")
    return ref


def reference_defaults():
    src = open(os.path.join(REFERENCE_DIR, "sugarscape.py")).read()
    for node in ast.parse(src).body:
        if isinstance(node, ast.If):
            for st in node.body:
                if isinstance(st, ast.Assign) and getattr(st.targets[0], "id", "") == "configuration":
                    return ast.literal_eval(st.value)
    raise RuntimeError("defaults dict not found in reference")


def build_reference(config_path, overrides=None, logfile=None):
    """Return a reference Sugarscape object positioned at t=0 (tests/test.py protocol overrides)."""
    ref = import_reference()
    conf = ref.parseConfiguration(config_path, reference_defaults())
    conf["headlessMode"] = True
    conf["debugMode"] = ["none"]
    conf["logfile"] = logfile
    for k, v in (overrides or {}).items():
        conf[k] = v
    ref.verifyRandomSeed(conf)
    conf = ref.verifyConfiguration(conf)
    S = ref.Sugarscape(conf)
    return S


# ---------------------------------------------------------------------------------------------
# canonical state extraction (value-level, not repr-level: the reference mixes int and float)
# ---------------------------------------------------------------------------------------------

def tags_to_mask(tags):
    if tags is None:
        return -1
    m = 0
    for i, b in enumerate(tags):
        if b:
            m |= (1 << i)
    return m


def snapshot(S):
    """Canonical arrays of the reference state.  Grid arrays are x-major (idx = x*H + y)."""
    env = S.environment
    W, H = env.width, env.height
    sugar = np.zeros((W, H), np.float64)
    spice = np.zeros((W, H), np.float64)
    poll = np.zeros((W, H), np.float64)
    occ = np.full((W, H), -1, np.int64)
    for x in range(W):
        col = env.grid[x]
        for y in range(H):
            c = col[y]
            sugar[x, y] = c.sugar
            spice[x, y] = c.spice
            poll[x, y] = c.pollution
            if c.agent is not None:
                occ[x, y] = c.agent.ID
    n = len(S.agents)
    a_id = np.zeros(n, np.int64)
    a_x = np.zeros(n, np.int64)
    a_y = np.zeros(n, np.int64)
    a_age = np.zeros(n, np.int64)
    a_sugar = np.zeros(n, np.float64)
    a_spice = np.zeros(n, np.float64)
    a_tags = np.zeros(n, np.int64)
    a_tribe = np.zeros(n, np.int64)
    a_alive = np.zeros(n, np.int64)
    for i, a in enumerate(S.agents):
        a_id[i] = a.ID
        a_x[i] = a.cell.x if a.cell is not None else -1
        a_y[i] = a.cell.y if a.cell is not None else -1
        a_age[i] = a.age
        a_sugar[i] = a.sugar
        a_spice[i] = a.spice
        a_tags[i] = tags_to_mask(a.tags)
        a_tribe[i] = -1 if a.tribe is None else a.tribe
        a_alive[i] = 1 if a.alive else 0
    st = random.getstate()
    mt = np.array(st[1], dtype=np.uint32)  # 624 words + index
    return {
        "sugar": sugar, "spice": spice, "pollution": poll, "occ": occ,
        "a_id": a_id, "a_x": a_x, "a_y": a_y, "a_age": a_age,
        "a_sugar": a_sugar, "a_spice": a_spice, "a_tags": a_tags, "a_tribe": a_tribe,
        "a_alive": a_alive, "mt": mt,
    }


def _h(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()[:16]


def digests(snap):
    """Short digests of the canonical state; the same function is applied to the oracle's
    and the CUDA engine's state in the parity tests (tests/parity_util.py re-implements it
    on the repo's own SoA)."""
    return {
        "agents_int": _h(snap["a_id"].astype("<i8"), snap["a_x"].astype("<i8"),
                         snap["a_y"].astype("<i8"), snap["a_age"].astype("<i8")),
        "agents_f64": _h(snap["a_sugar"].astype("<f8"), snap["a_spice"].astype("<f8")),
        "agents_tags": _h(snap["a_tags"].astype("<i8"), snap["a_tribe"].astype("<i8")),
        "grid_int": _h(snap["sugar"].astype("<f8"), snap["spice"].astype("<f8"),
                       snap["occ"].astype("<i8")),
        "pollution": _h(snap["pollution"].astype("<f8")),
        "mt": _h(snap["mt"].astype("<u4")),
    }
