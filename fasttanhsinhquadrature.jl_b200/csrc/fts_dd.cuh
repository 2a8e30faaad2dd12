// fts_dd.cuh — double-double arithmetic (the device twin of DoubleFloats.Double64: an unevaluated
// sum hi+lo).  Error-free transformations (Knuth TwoSum, FMA TwoProd) + QD-style exp/log/sqrt.
// Everything is __host__ __device__ so the same code is unit-tested on the CPU
// (tests/test_dd_arith.py compiles it with g++ -ffp-contract=off) and used by the sm_100a kernels
// (compiled with --fmad=false: no implicit contraction may break the transformations).
#pragma once
#if defined(__CUDACC__) || defined(__CUDACC_RTC__)
#define FTS_HD __host__ __device__ __forceinline__
#else
#include <cmath>
#define FTS_HD inline
#endif

namespace fts {

FTS_HD double dd_inf() { return 1.0 / 0.0; }
FTS_HD double dd_nan() { return 0.0 / 0.0; }

struct dd {
    double hi, lo;
    FTS_HD dd() {}
    FTS_HD constexpr dd(double h) : hi(h), lo(0.0) {}
    FTS_HD constexpr dd(double h, double l) : hi(h), lo(l) {}
};

FTS_HD dd quick_two_sum(double a, double b)
{
    double s = a + b;
    return dd(s, b - (s - a));
}
FTS_HD dd two_sum(double a, double b)
{
    double s = a + b;
    double bb = s - a;
    return dd(s, (a - (s - bb)) + (b - bb));
}
FTS_HD dd two_prod(double a, double b)
{
    double p = a * b;
    return dd(p, ::fma(a, b, -p));
}

FTS_HD dd operator-(dd a) { return dd(-a.hi, -a.lo); }

// (non-finite results collapse to (x, 0): TwoSum/TwoProd would turn an Inf into NaN)
FTS_HD bool dd_finite(double v) { return v - v == 0.0; }

FTS_HD dd operator+(dd a, dd b)
{
    dd s = two_sum(a.hi, b.hi);
    if (!dd_finite(s.hi)) return dd(a.hi + b.hi);
    dd t = two_sum(a.lo, b.lo);
    s.lo += t.hi;
    s = quick_two_sum(s.hi, s.lo);
    s.lo += t.lo;
    return quick_two_sum(s.hi, s.lo);
}
FTS_HD dd operator-(dd a, dd b) { return a + (-b); }
FTS_HD dd operator+(dd a, double b)
{
    dd s = two_sum(a.hi, b);
    if (!dd_finite(s.hi)) return dd(a.hi + b);
    s.lo += a.lo;
    return quick_two_sum(s.hi, s.lo);
}
FTS_HD dd operator*(dd a, dd b)
{
    dd p = two_prod(a.hi, b.hi);
    if (!dd_finite(p.hi)) return dd(p.hi);
    p.lo += a.hi * b.lo + a.lo * b.hi;
    return quick_two_sum(p.hi, p.lo);
}
FTS_HD dd operator*(dd a, double b)
{
    dd p = two_prod(a.hi, b);
    if (!dd_finite(p.hi)) return dd(p.hi);
    p.lo += a.lo * b;
    return quick_two_sum(p.hi, p.lo);
}
FTS_HD dd operator/(dd a, dd b)
{
    double q1 = a.hi / b.hi;
    if (!dd_finite(q1) || !dd_finite(b.hi)) return dd(q1);
    dd r = a - b * q1;
    double q2 = r.hi / b.hi;
    r = r - b * q2;
    double q3 = r.hi / b.hi;
    dd q = quick_two_sum(q1, q2);
    return q + q3;
}
FTS_HD dd& operator+=(dd& a, dd b) { a = a + b; return a; }
FTS_HD dd& operator*=(dd& a, dd b) { a = a * b; return a; }
FTS_HD bool operator<(dd a, dd b) { return a.hi < b.hi || (a.hi == b.hi && a.lo < b.lo); }
FTS_HD bool operator>(dd a, dd b) { return b < a; }
FTS_HD bool operator<=(dd a, dd b) { return a.hi < b.hi || (a.hi == b.hi && a.lo <= b.lo); }
FTS_HD bool operator>=(dd a, dd b) { return b <= a; }
FTS_HD bool operator==(dd a, dd b) { return a.hi == b.hi && a.lo == b.lo; }
FTS_HD bool operator!=(dd a, dd b) { return !(a == b); }

FTS_HD dd dd_ldexp(dd a, int e) { return dd(::ldexp(a.hi, e), ::ldexp(a.lo, e)); }
FTS_HD dd dd_mul_pwr2(dd a, double p) { return dd(a.hi * p, a.lo * p); }

FTS_HD dd dd_sqrt(dd a)
{
    if (a.hi <= 0.0) return dd(a.hi == 0.0 ? 0.0 : dd_nan());
    if (!dd_finite(a.hi)) return dd(a.hi);
    // Karp & Markstein: sqrt(a) = a*x + [a - (a*x)^2] * x / 2, x ~ 1/sqrt(a)
    double x = 1.0 / ::sqrt(a.hi);
    double ax = a.hi * x;
    dd axx = two_prod(ax, ax);
    dd r = a - axx;
    return quick_two_sum(ax, 0.0) + r.hi * (x * 0.5);
}

// exp(dd): k = round(x/ln2), r = (x - k ln2)/512, Taylor on r, square 9 times, scale by 2^k.
FTS_HD dd dd_exp(dd a)
{
    const dd LN2(6.931471805599452862e-01, 2.319046813846299558e-17);
    if (a.hi != a.hi) return dd(a.hi);
    if (a.hi < -745.2) return dd(0.0);  // down to the smallest subnormal (ldexp below rounds once)
    if (a.hi >= 709.0) return dd(dd_inf());
    if (a.hi == 0.0 && a.lo == 0.0) return dd(1.0);
    double kf = ::floor(a.hi / LN2.hi + 0.5);
    dd r = dd_mul_pwr2(a - LN2 * kf, 1.0 / 512.0);
    // Taylor series sum_{i>=1} r^i/i!  (|r| <= ln2/1024: 9 terms reach 2^-110)
    const dd inv_fact[8] = {dd(1.66666666666666657e-01, 9.25185853854297066e-18), dd(4.16666666666666644e-02, 2.31296463463574266e-18),
                            dd(8.33333333333333322e-03, 1.15648231731787138e-19), dd(1.38888888888888894e-03, -5.30054395437357706e-20),
                            dd(1.98412698412698413e-04, 1.72095582934207053e-22), dd(2.48015873015873016e-05, 2.15119478667758816e-23),
                            dd(2.75573192239858925e-06, -1.85839327404647208e-22), dd(2.75573192239858883e-07, 2.37677146222502973e-23)};
    dd p = r * r;
    dd s = r + dd_mul_pwr2(p, 0.5);
    p = p * r;
    dd t = p * inv_fact[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) {
        s = s + t;
        p = p * r;
        t = p * inv_fact[i];
    }
    s = s + t;
    // (1+s)^512 - 1 by nine doublings: s <- 2s + s^2
#pragma unroll
    for (int i = 0; i < 9; ++i) s = dd_mul_pwr2(s, 2.0) + s * s;
    s = s + 1.0;
    return dd_ldexp(s, (int)kf);
}

// log(dd): one Newton step on exp from the double estimate: x1 = x0 + a*exp(-x0) - 1
FTS_HD dd dd_log(dd a)
{
    if (a.hi <= 0.0)
        return dd(a.hi == 0.0 ? (-dd_inf()) : dd_nan());
    if (a.hi - a.hi != 0.0) return a;  // +Inf
    // a = m * 2^e with m in [0.5, 1): Newton on m keeps exp(-x) of order one for every finite a
    // (subnormal complement distances down to 1e-310 included), then log a = log m + e ln2
    const dd LN2(6.931471805599452862e-01, 2.319046813846299558e-17);
    int e = 0;
    const double mh = ::frexp(a.hi, &e);
    const dd m(mh, ::ldexp(a.lo, -e));
    dd x(::log(mh));
    x = x + (m * dd_exp(-x) - dd(1.0));
    x = x + (m * dd_exp(-x) - dd(1.0));
    return x + LN2 * (double)e;
}

// sin/cos(dd): k = round(x/(pi/2)), r = x - k pi/2 (|r| <= pi/4), Taylor for sin(r/8),
// cos = sqrt(1 - sin^2), three angle doublings, quadrant fix-up.  pi/2 is carried to 106 bits, so
// the absolute error grows like |x| * 1e-32 (fine for the oscillatory test families, |x| < 1e5).
FTS_HD void dd_sincos(dd a, dd& s, dd& c)
{
    const dd PIO2(1.570796326794896558e+00, 6.123233995736766036e-17);
    if (!(a.hi - a.hi == 0.0)) {  // NaN or Inf
        s = c = dd(dd_nan());
        return;
    }
    const double kf = ::floor(a.hi / PIO2.hi + 0.5);
    const dd r = a - PIO2 * kf;
    const dd t = dd_mul_pwr2(r, 0.125);
    const dd t2 = t * t;
    dd term = t, sum = t;
#pragma unroll 1
    for (int i = 1; i <= 10; ++i) {
        term = -(term * t2) / dd((double)((2 * i) * (2 * i + 1)));
        sum = sum + term;
    }
    dd st = sum, ct = dd_sqrt(dd(1.0) - st * st);
#pragma unroll 1
    for (int i = 0; i < 3; ++i) {
        const dd s2 = dd_mul_pwr2(st * ct, 2.0);
        const dd c2 = dd(1.0) - dd_mul_pwr2(st * st, 2.0);
        st = s2;
        ct = c2;
    }
    switch (((long long)kf) & 3) {
    case 0: s = st; c = ct; break;
    case 1: s = ct; c = -st; break;
    case 2: s = -st; c = -ct; break;
    default: s = -ct; c = st; break;
    }
}
FTS_HD dd dd_sin(dd a) { dd s, c; dd_sincos(a, s, c); return s; }
FTS_HD dd dd_cos(dd a) { dd s, c; dd_sincos(a, s, c); return c; }
FTS_HD dd sin(dd a) { return dd_sin(a); }
FTS_HD dd cos(dd a) { return dd_cos(a); }

// unqualified exp/log/sqrt/fabs on a dd argument resolve here by ADL, so a user integrand body
// such as `return exp(-p[0]*x*x);` compiles unchanged for float, double and Double64
FTS_HD dd exp(dd a) { return dd_exp(a); }
FTS_HD dd log(dd a) { return dd_log(a); }
FTS_HD dd sqrt(dd a) { return dd_sqrt(a); }
FTS_HD dd fabs(dd a) { return (a.hi < 0.0 || (a.hi == 0.0 && a.lo < 0.0)) ? -a : a; }
FTS_HD dd abs(dd a) { return fabs(a); }
FTS_HD dd operator+(double a, dd b) { return b + a; }
FTS_HD dd operator*(double a, dd b) { return b * a; }
FTS_HD dd operator-(dd a, double b) { return a + (-b); }
FTS_HD dd operator-(double a, dd b) { return (-b) + a; }
FTS_HD dd operator/(dd a, double b) { return a / dd(b); }
FTS_HD dd operator/(double a, dd b) { return dd(a) / b; }


}  // namespace fts
