"""csrc/fastmath.cuh -- the loop's own log / exp / sincos.  CPU: the same source compiled for the host against 50-digit
mpmath (<= 0.8 ulp).  GPU: the device code against glibc through numpy (both are within 1 ulp of the truth, so within
2 ulp of each other, and equal in the large majority of cases)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = C.POINTER(C.c_double)


def _samples(seed=7, n=4000):
    rng = np.random.default_rng(seed)
    x_exp = -np.concatenate([rng.random(n) * 6, rng.random(n // 4) * 699, 10.0 ** rng.uniform(-18, 0, n // 2),
                             [0.0, 1e-300, 37.4, 37.6, 699.999]])
    u = np.concatenate([rng.random(n), 1 - 10.0 ** rng.uniform(-16, -1, n // 2), 10.0 ** rng.uniform(-16, 0, n // 2),
                        [2.0 ** -53, 1 - 2.0 ** -53, 0.5, 0.70710678118654746, 0.70710678118654757]])
    u = u[(u > 0) & (u < 1)]
    phi = np.concatenate([rng.random(n) * 2 * np.pi, np.arange(0, 9) * (np.pi / 4), np.arange(0, 9) * (np.pi / 4) +
                          rng.normal(0, 1e-9, 9), [2.0 ** -53 * 2 * np.pi, 6.283185307179586]])
    return x_exp, u, phi


def _ulps(got, want_mp):
    import mpmath as mp
    worst = 0.0
    for g, t in zip(got, want_mp):
        if t == 0:
            assert g == 0
            continue
        worst = max(worst, float(abs(mp.mpf(float(g)) - t) / mp.mpf(float(np.spacing(abs(float(t)))))))
    return worst


@pytest.fixture(scope="module")
def hostlib(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("fm") / "fastmath_host.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-DBG_FASTMATH_HOST",
                           "-o", so, os.path.join(ROOT, "tests", "fastmath_host.cc")])
    return C.CDLL(so)


def _call(fn, x):
    x = np.ascontiguousarray(x, np.float64)
    y = np.empty_like(x)
    fn(x.ctypes.data_as(P), y.ctypes.data_as(P), C.c_long(x.size))
    return y


def test_host_build_against_mpmath(hostlib):
    import mpmath as mp
    mp.mp.dps = 50
    x_exp, u, phi = _samples()
    assert _ulps(_call(hostlib.fmh_exp, x_exp), [mp.exp(mp.mpf(float(v))) for v in x_exp]) < 0.70
    assert _ulps(_call(hostlib.fmh_log, u), [mp.log(mp.mpf(float(v))) for v in u]) < 0.60
    wide = 10.0 ** np.random.default_rng(3).uniform(-300, 300, 2000)
    assert _ulps(_call(hostlib.fmh_log, wide), [mp.log(mp.mpf(float(v))) for v in wide]) < 0.60
    s, c = np.empty_like(phi), np.empty_like(phi)
    hostlib.fmh_sincos(phi.ctypes.data_as(P), s.ctypes.data_as(P), c.ctypes.data_as(P), C.c_long(phi.size))
    assert _ulps(s, [mp.sin(mp.mpf(float(v))) for v in phi]) < 0.80
    assert _ulps(c, [mp.cos(mp.mpf(float(v))) for v in phi]) < 0.80


def test_host_build_special_values(hostlib):
    e = _call(hostlib.fmh_exp, np.array([0.0, -0.0, -700.0, -1e5, -np.inf, 700.0, np.inf]))
    assert list(e[:5]) == [1.0, 1.0, 0.0, 0.0, 0.0] and np.isinf(e[5]) and np.isinf(e[6])
    assert np.isnan(_call(hostlib.fmh_exp, np.array([np.nan]))[0])
    # the flush never changes the consumer: 1 - exp(x) is exactly 1 long before x = -700
    assert 1.0 - np.exp(-37.5) == 1.0
    # log of the draw, straight from the Threefry word: log(((w >> 11) | 1) * 2^-53)
    w = np.random.default_rng(5).integers(0, 2 ** 64, 2000, dtype=np.uint64)
    y = np.empty(w.size)
    hostlib.fmh_log_u01_bits(w.ctypes.data_as(C.c_void_p), y.ctypes.data_as(P), C.c_long(w.size))
    u = ((w >> np.uint64(11)) | np.uint64(1)).astype(np.float64) * 2.0 ** -53
    assert np.array_equal(y, _call(hostlib.fmh_log, u))


def _ulp_diff(a, b):
    return np.abs(a - b) / np.spacing(np.maximum(np.abs(a), np.abs(b)))


@pytest.mark.gpu
def test_device_build_against_glibc():
    from branson_b200 import gpu
    rng = np.random.default_rng(11)
    n = 2_000_000
    x = -np.concatenate([rng.random(n) * 8, rng.random(n // 8) * 699, 10.0 ** rng.uniform(-18, 0, n // 4)])
    got, want = gpu.fastmath("exp", x), np.exp(x)
    assert _ulp_diff(got, want).max() <= 1.0 and np.mean(got == want) > 0.90
    w = rng.integers(0, 2 ** 64, n, dtype=np.uint64)
    u = ((w >> np.uint64(11)) | np.uint64(1)).astype(np.float64) * 2.0 ** -53
    u = np.concatenate([u, 1 - 10.0 ** rng.uniform(-16, -1, n // 8)])
    got, want = gpu.fastmath("log", u), np.log(u)
    assert _ulp_diff(got, want).max() <= 1.0 and np.mean(got == want) > 0.98
    phi = rng.random(n) * 2.0 * 3.1415926535897932
    (s, c), ws, wc = gpu.fastmath("sincos", phi), np.sin(phi), np.cos(phi)
    assert _ulp_diff(s, ws).max() <= 2.0 and _ulp_diff(c, wc).max() <= 2.0
    assert np.mean(s == ws) > 0.95 and np.mean(c == wc) > 0.95
    # no worse than libdevice's sincos on the same arguments
    cs, cc = gpu.fastmath("cuda_sincos", phi)
    assert np.mean(s == ws) >= np.mean(cs == ws) - 0.02 and np.mean(c == wc) >= np.mean(cc == wc) - 0.02


@pytest.mark.gpu
def test_device_division_is_ieee_on_the_loop_ranges():
    """fm_div (the compiler's division fast path without its range test) against IEEE division: bit for bit."""
    from branson_b200 import gpu
    rng = np.random.default_rng(23)
    n = 4_000_000
    x = rng.standard_normal(n) * 10.0 ** rng.uniform(-100, 100, n)
    x[: n // 4] = rng.uniform(-1, 1, n // 4)            # direction cosines, face distances
    x[n // 4: n // 2: 2] = 0.0                          # zero numerators (photon on a face, absorbed == 0)
    got = gpu.fastmath("div", x)
    den = x.reshape(-1, 2)[:, ::-1].reshape(-1)
    ok = den != 0.0
    want = x[ok] / den[ok]
    bad = got[ok].view(np.uint64) != want.view(np.uint64)
    assert not bad.any(), (int(bad.sum()), x[ok][bad][:5], den[ok][bad][:5])
    # a zero divisor gives NaN where IEEE gives +-inf (or NaN): every consumer in the loop is a `d < d_min` test
    assert not np.isfinite(got[~ok]).any()
    # fm_sqrt: the compiler's square-root fast path; the loop's arguments are 1 - mu^2 in [2^-51, 1]
    a = np.concatenate([1.0 - (2.0 * rng.random(n) - 1.0) ** 2, 2.0 ** rng.uniform(-52, 1, n // 4),
                        10.0 ** rng.uniform(-100, 100, n // 4)])
    a = a[a > 0]
    assert np.array_equal(gpu.fastmath("sqrt", a).view(np.uint64), np.sqrt(a).view(np.uint64))
