"""include/pgd_math.h (the float32 sin / cos / atan2 / exp shared by the CUDA step and the CPU oracle) against double
libm over the ranges the step uses.  The functions are explicit fmaf chains, so every build gives the same bits; what
has to be pinned is that those bits are accurate (the reference computes in float64: numpy cos / sin / arctan2,
utils/math_utils.py:32-33)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE = os.path.join(os.path.dirname(HERE), "oracle")


@pytest.fixture(scope="module")
def probe():
    subprocess.check_call(["make", "-C", ORACLE, "_build/libpgd_math_probe.so"], stdout=subprocess.DEVNULL)
    L = C.CDLL(os.path.join(ORACLE, "_build", "libpgd_math_probe.so"))
    for f, n in (("probe_sincos", 3), ("probe_atan2", 3), ("probe_wrap", 2), ("probe_exp", 2), ("probe_pow10", 2),
                 ("probe_tan", 2), ("probe_asin", 2)):
        getattr(L, f).argtypes = [C.c_void_p] * n + [C.c_int]
    return L


def _call(fn, *arrays):
    out = [np.empty_like(arrays[0]) for _ in range(2 if fn.__name__ == "probe_sincos" else 1)]
    fn(*[a.ctypes.data for a in arrays], *[o.ctypes.data for o in out], len(arrays[0]))
    return out if len(out) > 1 else out[0]


def _ulp(got, want):
    sp = np.spacing(np.abs(want.astype(np.float32))).astype(np.float64)
    return np.abs(got.astype(np.float64) - want) / sp


def test_sincos(probe):
    rs = np.random.RandomState(0)
    a = np.concatenate([rs.uniform(-64, 64, 1000000), rs.uniform(-7, 7, 1000000),
                        np.linspace(-np.pi, np.pi, 100001), np.arange(-40, 41) * (np.pi / 2)]).astype(np.float32)
    s, c = _call(probe.probe_sincos, a)
    a64 = a.astype(np.float64)
    # absolute error: a heading's unit vector is used as such (1 ulp of 1.0 is 1.2e-7)
    assert np.abs(s - np.sin(a64)).max() < 1.2e-7
    assert np.abs(c - np.cos(a64)).max() < 1.2e-7
    assert np.abs(s * s + c * c - 1.0).max() < 4e-7
    z = np.zeros(1, np.float32)
    s0, c0 = _call(probe.probe_sincos, z)
    assert s0[0] == 0.0 and c0[0] == 1.0


def test_atan2(probe):
    rs = np.random.RandomState(1)
    y = rs.uniform(-400, 400, 2000000).astype(np.float32)
    x = rs.uniform(-400, 400, 2000000).astype(np.float32)
    r = _call(probe.probe_atan2, y, x)
    assert _ulp(r, np.arctan2(y.astype(np.float64), x.astype(np.float64))).max() < 2.5
    ax = np.array([0, 0, 1, -1, 0, 0, 1, -1, -1], np.float32)  # x
    ay = np.array([0, 1, 0, 0, -1, 0, 1, 1, -1], np.float32)   # y
    want = np.arctan2(ay.astype(np.float64), ax.astype(np.float64))
    assert np.abs(_call(probe.probe_atan2, ay, ax) - want).max() < 3e-7


def test_wrap_exp_tan_pow_asin(probe):
    rs = np.random.RandomState(2)
    w = rs.uniform(-30, 30, 1000000).astype(np.float32)
    r = _call(probe.probe_wrap, w)
    ref = (w.astype(np.float64) + np.pi) % (2 * np.pi) - np.pi
    d = np.abs(r - ref)
    assert np.minimum(d, np.abs(d - 2 * np.pi)).max() < 4e-7  # an exact odd multiple of pi may land on either end
    assert r.min() >= -np.pi - 1e-6 and r.max() <= np.pi + 1e-6
    e = rs.uniform(-2, 3, 1000000).astype(np.float32)
    assert _ulp(_call(probe.probe_exp, e), np.exp(e.astype(np.float64))).max() < 1.5
    t = rs.uniform(-1.4, 1.4, 1000000).astype(np.float32)
    assert _ulp(_call(probe.probe_tan, t), np.tan(t.astype(np.float64))).max() < 4.0
    p = rs.uniform(0, 3, 1000000).astype(np.float32)
    assert _ulp(_call(probe.probe_pow10, p), p.astype(np.float64) ** 10).max() < 8.0  # four roundings, each amplified by the remaining power
    q = rs.uniform(0, 1, 1000000).astype(np.float32)
    assert np.abs(_call(probe.probe_asin, q) - np.arcsin(q.astype(np.float64))).max() < 1e-6
