"""GPU tests of the library's own elementary functions (pytest -m gpu): the table-driven log and exp
of fts_math.cuh and the lock-step rsqrt, evaluated pointwise through fts_eval_batch and held to their
ulp bounds against mpmath.  The oracle's libm (glibc) is <= 1 ulp as well, so integrals agree to the
1e-14 parity tolerance; these tests pin the functions themselves."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu(fts):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    fts.init(0)
    return fts


def _ulps(got, exact_mpf):
    """error of each got[i] in units of the last place of the correctly rounded result"""
    import mpmath
    out = np.zeros(len(got))
    for i, (g, e) in enumerate(zip(got, exact_mpf)):
        ef = float(e)
        if ef == 0.0 or not math.isfinite(ef):
            out[i] = 0.0 if (g == ef or (math.isnan(g) and math.isnan(ef))) else np.inf
            continue
        ulp = math.ulp(ef)
        out[i] = float(abs(mpmath.mpf(float(g)) - e) / mpmath.mpf(ulp))
    return out


def test_table_log_ulp(gpu):
    """log_n (128-entry table in shared memory, 15 FP64 instructions): <= 0.8 ulp (0.72 measured, in the buckets
    next to 1 where the rounding of r = z*invc - 1 weighs most; CUDA's log is 1 ulp) over (0, inf), including
    the neighbourhood of 1 where log -> 0 (the table's bucket around 1 is exact, no special case),
    bucket boundaries, subnormal results of 1 - x, and the special inputs that take CUDA's log."""
    import mpmath
    mpmath.mp.prec = 200
    rng = np.random.default_rng(7)
    # log(1 - x): arguments in (0, 2)
    x = np.concatenate([rng.uniform(-1, 1, 6000), 1 - 10.0 ** rng.uniform(-16, 0, 3000), 10.0 ** rng.uniform(-18, -1, 2000) * rng.choice([-1, 1], 2000),
                        np.array([0.0, 0.5, -0.5, 2.0 ** -60, -2.0 ** -60, 1 - 2.0 ** -53, -1.0, np.nextafter(-1.0, 0)])])
    arg = 1.0 - x
    got = gpu.evaluate(gpu.builtin("log1mx"), x)
    u = _ulps(got, [mpmath.log(mpmath.mpf(float(a))) for a in arg])
    assert u.max() <= 0.8, (u.max(), x[np.argmax(u)])
    # log(1 + x): arguments over the whole exponent range, and every bucket boundary of the table
    big = 10.0 ** rng.uniform(-300, 300, 6000)
    off = 0x3FE5F00000000000
    edges = np.array([off + (i << 45) + d for i in range(129) for d in (-1, 0, 1)], dtype=np.uint64).view(np.float64)
    xs = np.concatenate([big, edges * 2.0 ** rng.integers(-40, 40, edges.size)]) - 1.0
    arg = 1.0 + xs
    got = gpu.evaluate(gpu.builtin("log1px"), xs)
    u = _ulps(got, [mpmath.log(mpmath.mpf(float(a))) for a in arg])
    assert u.max() <= 0.8, (u.max(), xs[np.argmax(u)])
    assert np.median(u) < 0.3
    # special inputs go to CUDA's log: log(0) = -inf, log(negative) = nan, log(inf) = inf, subnormals are finite
    sp = gpu.evaluate(gpu.builtin("log1px"), np.array([-1.0, -2.0, np.inf, np.nan]))
    assert sp[0] == -np.inf and np.isnan(sp[1]) and sp[2] == np.inf and np.isnan(sp[3])
    # Float32 keeps CUDA's logf
    g32 = gpu.evaluate(gpu.builtin("log1mx"), np.float32([0.0, 0.5, -0.75]), T=np.float32)
    assert np.allclose(g32, np.log(1 - np.float32([0.0, 0.5, -0.75])), rtol=3e-7, atol=1e-7)


def test_table_exp_and_lockstep_rsqrt_ulp(gpu):
    import mpmath
    mpmath.mp.prec = 200
    rng = np.random.default_rng(8)
    x = np.concatenate([rng.uniform(-700, 700, 6000), rng.uniform(-1, 1, 3000), np.array([0.0, 708.0, -708.0, 709.7, -745.0, -746.0, 710.0])])
    got = gpu.evaluate(gpu.builtin("exp"), x)
    u = _ulps(got, [mpmath.exp(mpmath.mpf(float(v))) for v in x])
    assert u.max() <= 0.75, (u.max(), x[np.argmax(u)])           # 0.51 measured on the fast path; the tails use CUDA's exp (<= 1 ulp)
    # gauss: exp(-a x^2) with the argument formed as the kernel forms it
    a = rng.uniform(0.5, 50, 4000)
    xs = rng.uniform(-1, 1, 4000)
    got = gpu.evaluate(gpu.builtin("gauss"), xs, params=a.reshape(-1, 1))
    u = _ulps(got, [mpmath.exp(mpmath.mpf(float(-ai * (xi * xi)))) for ai, xi in zip(a, xs)])
    assert u.max() <= 0.75
    # 1/sqrt(1 - x^2): fma(-x, x, 1) is exact-rounded, then the lock-step rsqrt (MUFU seed + one third-order step, <= 2 ulp like CUDA's)
    xr = np.concatenate([rng.uniform(-1, 1, 6000), 1 - 10.0 ** rng.uniform(-16, -1, 2000)])
    got = gpu.evaluate(gpu.builtin("invsqrt1mx2"), xr)
    exact = []
    for v in xr:
        m = mpmath.mpf(float(v))
        arg = mpmath.mpf(float(1 - m * m))        # one rounding: what fma(-x, x, 1) returns
        exact.append(1 / mpmath.sqrt(arg))
    u = _ulps(got, exact)
    assert u.max() <= 2.0, (u.max(), xr[np.argmax(u)])
    sp = gpu.evaluate(gpu.builtin("invsqrt1mx2"), np.array([1.0, 2.0, 0.0]))
    assert sp[0] == np.inf and np.isnan(sp[1]) and sp[2] == 1.0
