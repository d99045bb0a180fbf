"""The engine's pow (csrc/ssb_pow.cuh) must equal the host libm's pow -- which is what CPython's
float ** calls in the reference (agent.py:1189,1200) -- in every bit."""
import math
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_fma():
    try:
        return " fma " in open("/proc/cpuinfo").read()
    except OSError:
        return False


@pytest.mark.skipif(not _has_fma(), reason="host CPU without FMA3: glibc selects another pow variant")
def test_restated_pow_matches_host_libm(tmp_path):
    exe = tmp_path / "pow_check"
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-mfma", "-o", str(exe),
                           os.path.join(ROOT, "tests", "pow_check.c"), "-lm"])
    out = subprocess.run([str(exe), "2000000"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout
    assert "bad=0" in out.stdout


@pytest.mark.gpu
@pytest.mark.skipif(not _has_fma(), reason="host CPU without FMA3")
def test_device_pow_matches_cpython_pow():
    import torch
    from sugarscape_b200 import engine
    L = engine.load_library()
    rng = np.random.default_rng(7)
    n = 200000
    x = np.concatenate([rng.integers(1, 4000, n // 2) / 2.0, rng.random(n // 2) * 1000 + 1e-3])
    m = rng.integers(1, 9, n)
    y = m / (m + rng.integers(0, 9, n))
    want = np.array([math.pow(a, b) for a, b in zip(x, y)])
    dx, dy = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
    out = torch.empty(n, dtype=torch.float64, device="cuda")
    exact = torch.empty(n, dtype=torch.int32, device="cuda")
    assert L.ssb_test_pow(dx.data_ptr(), dy.data_ptr(), out.data_ptr(), exact.data_ptr(), n) == 0
    got = out.cpu().numpy()
    assert int(exact.sum()) == n, "an argument of the welfare range fell outside the restated path"
    assert np.array_equal(got.view(np.uint64), want.view(np.uint64))
