"""The C-ABI library: builds for sm_100a, loads without a GPU and exports every symbol that
include/ssb.h declares; the ctypes mirrors of the ABI structs have the library's sizes."""
import ctypes
import os
import re

from sugarscape_b200 import build, engine, layout

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "ssb.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ssb_[a-z_0-9]+)\s*\(", text)))


def test_library_builds_and_exports_the_declared_abi():
    path = build.build()
    lib = ctypes.CDLL(path)
    names = declared_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/ssb.h but not exported"
    assert sorted(engine.EXPORTED_SYMBOLS) == names


def test_struct_mirrors_match_library():
    lib = engine.load_library()
    assert lib.ssb_abi_version() == layout.ABI_VERSION
    assert lib.ssb_sizeof_params() == ctypes.sizeof(layout.Params)
    assert lib.ssb_sizeof_grid() == ctypes.sizeof(layout.Grid)
    assert lib.ssb_sizeof_agents() == ctypes.sizeof(layout.Agents)
    assert lib.ssb_sizeof_stats() == ctypes.sizeof(layout.Stats)


def test_engine_refuses_to_run_without_cuda():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    import parity_util as pu
    c, state, pop, params, endow, tags = pu.build("constant_growback")
    with pytest.raises(engine.EngineError):
        engine.Engine(params, state, 0, endow, tags)
