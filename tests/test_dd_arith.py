"""CPU unit tests of the product's double-double arithmetic (csrc/fts_dd.cuh is __host__ __device__):
compiled here with g++ -ffp-contract=off and compared with mpmath.  Double64 target: 1e-28."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def shim():
    out = os.path.join(ROOT, "tests", "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libdd_shim.so")
    subprocess.run(["/usr/bin/g++", "-O2", "-std=gnu++17", "-ffp-contract=off", "-fPIC", "-shared", "-o", so,
                    os.path.join(ROOT, "tests", "dd_host_shim.cpp")], check=True)
    return C.CDLL(so)


def _dd(v):
    import mpmath
    m = mpmath.mpf(v)
    hi = float(m)
    return np.array([hi, float(m - mpmath.mpf(hi))])


def _val(r):
    import mpmath
    return mpmath.mpf(float(r[0])) + mpmath.mpf(float(r[1]))


def _un(lib, name, a):
    r = np.zeros(2)
    getattr(lib, name)(a.ctypes.data_as(C.c_void_p), r.ctypes.data_as(C.c_void_p))
    return r


def _bi(lib, name, a, b):
    r = np.zeros(2)
    getattr(lib, name)(a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), r.ctypes.data_as(C.c_void_p))
    return r


def test_dd_accuracy_against_mpmath(shim):
    import mpmath
    mpmath.mp.prec = 250
    rng = np.random.default_rng(0)
    worst = {}

    def upd(k, e):
        worst[k] = max(worst.get(k, 0.0), float(e))

    for _ in range(1500):
        x = mpmath.mpf(rng.random()) * mpmath.mpf(10) ** int(rng.integers(-5, 5)) + mpmath.mpf(rng.random()) * mpmath.mpf(2) ** -60
        y = mpmath.mpf(rng.random() + 0.1) * mpmath.mpf(10) ** int(rng.integers(-5, 5)) + mpmath.mpf(rng.random()) * mpmath.mpf(2) ** -70
        a, b = _dd(x), _dd(y)
        xa, yb = _val(a), _val(b)
        upd("add", abs(_val(_bi(shim, "t_dd_add", a, b)) - (xa + yb)) / abs(xa + yb))
        upd("mul", abs(_val(_bi(shim, "t_dd_mul", a, b)) - (xa * yb)) / abs(xa * yb))
        upd("div", abs(_val(_bi(shim, "t_dd_div", a, b)) - (xa / yb)) / abs(xa / yb))
        upd("sqrt", abs(_val(_un(shim, "t_dd_sqrt", b)) - mpmath.sqrt(yb)) / mpmath.sqrt(yb))
        upd("rsqrt", abs(_val(_un(shim, "t_dd_rsqrt", b)) * mpmath.sqrt(yb) - 1))
        upd("log", abs(_val(_un(shim, "t_dd_log", b)) - mpmath.log(yb)) / max(abs(mpmath.log(yb)), mpmath.mpf(1e-3)))
        e = _dd(mpmath.mpf(rng.uniform(-300, 300)) + mpmath.mpf(rng.random()) * mpmath.mpf(2) ** -60)
        upd("exp", abs(_val(_un(shim, "t_dd_exp", e)) - mpmath.exp(_val(e))) / mpmath.exp(_val(e)))
        s = _dd(mpmath.mpf(rng.uniform(-3000, 3000)))
        upd("sin", abs(_val(_un(shim, "t_dd_sin", s)) - mpmath.sin(_val(s))))
        upd("cos", abs(_val(_un(shim, "t_dd_cos", s)) - mpmath.cos(_val(s))))
    assert worst["add"] < 1e-31 and worst["mul"] < 1e-31 and worst["div"] < 1e-31 and worst["sqrt"] < 2e-31
    assert worst["rsqrt"] < 1e-31
    assert worst["log"] < 1e-31 and worst["exp"] < 1e-31 and worst["sin"] < 2e-28 and worst["cos"] < 2e-28


def test_dd_edge_cases_of_the_complement_tables(shim):
    """subnormal complement distances (c(tm) ~ 3e-311) and non-finite values"""
    import mpmath
    mpmath.mp.prec = 250
    tiny = _dd("1e-310")
    assert abs(_un(shim, "t_dd_sqrt", tiny)[0] - 1e-155) < 1e-165
    assert abs(_val(_un(shim, "t_dd_rsqrt", tiny)) * mpmath.sqrt(_val(tiny)) - 1) < 1e-30
    assert abs(_val(_un(shim, "t_dd_rsqrt", _dd("1e305"))) * mpmath.sqrt(_val(_dd("1e305"))) - 1) < 1e-31
    assert _un(shim, "t_dd_rsqrt", np.array([0.0, 0.0]))[0] == np.inf and np.isnan(_un(shim, "t_dd_rsqrt", np.array([-1.0, 0.0]))[0])
    assert _un(shim, "t_dd_rsqrt", np.array([np.inf, 0.0]))[0] == 0.0
    lg = _val(_un(shim, "t_dd_log", tiny))
    assert abs(lg - mpmath.log(_val(tiny))) < mpmath.mpf("1e-27")
    inv = _bi(shim, "t_dd_div", _dd(-1), tiny)
    assert inv[0] == -np.inf and inv[1] == 0.0                       # like Float64: -1/1e-310 = -Inf
    assert np.array_equal(_un(shim, "t_dd_exp", inv), [0.0, 0.0])    # exp(-1/bmx) -> 0 at the endpoint
    inf = np.array([np.inf, 0.0])
    assert _bi(shim, "t_dd_mul", inf, _dd(2))[0] == np.inf and _bi(shim, "t_dd_add", inf, _dd(1))[0] == np.inf
    assert np.isnan(_bi(shim, "t_dd_mul", _dd(0), inf)[0])           # 0 * Inf = NaN as in the reference
    assert _un(shim, "t_dd_log", _dd(0))[0] == -np.inf and np.isnan(_un(shim, "t_dd_sqrt", _dd(-1))[0])
