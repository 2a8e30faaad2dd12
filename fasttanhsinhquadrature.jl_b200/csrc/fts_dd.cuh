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
    // the later quotient digits come off one reciprocal (a double division is ~25 instructions on the GPU;
    // the digits only have to be good to a few ulps, each remainder corrects the one before)
    const double rcp = 1.0 / b.hi;
    dd r = a - b * q1;
    double q2 = r.hi * rcp;
    r = r - b * q2;
    double q3 = r.hi * rcp;
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

// 1/sqrt(a): double estimate y, then one Newton step in double-double, y + y (1 - a y^2)/2 (error 1.5 delta^2 of the
// estimate's delta ~ 1e-16); ~45 FP64 instructions where 1/dd_sqrt(a) spends three divisions and a square root
FTS_HD dd dd_rsqrt(dd a)
{
    if (a.hi <= 0.0) return dd(a.hi == 0.0 ? dd_inf() : dd_nan());
    if (!dd_finite(a.hi)) return dd(a.hi != a.hi ? a.hi : 0.0);
    // subnormal and near-overflow arguments: scale by an even power of two (exact)
    double sc = 1.0;
    if (a.hi < 0x1p-900) { a = dd_mul_pwr2(a, 0x1p200); sc = 0x1p100; }
    else if (a.hi > 0x1p900) { a = dd_mul_pwr2(a, 0x1p-200); sc = 0x1p-100; }
#ifdef __CUDA_ARCH__
    const double y = ::rsqrt(a.hi);
#else
    const double y = 1.0 / ::sqrt(a.hi);
#endif
    const dd t = a * two_prod(y, y);
    const double e = (1.0 - t.hi) - t.lo;
    const dd r = quick_two_sum(y, (0.5 * y) * e);
    return sc == 1.0 ? r : dd_mul_pwr2(r, sc);
}

// ---- exp / log: table-driven (tools/gen_dd_tables.py), ~180 / ~160 FP64 instructions per call ----------------
// (the QD-style forms they replace — Taylor + nine squarings, two Newton steps on that exp — cost ~700 / ~1500)
#include "fts_dd_tables.inc"
static const double dd_exp2_tab_host[64][2] = FTS_DD_EXP2_TAB_VALUES;
static const double dd_log_tab_host[97][2] = FTS_DD_LOG_TAB_VALUES;
#if defined(__CUDACC__) || defined(__CUDACC_RTC__)
static __device__ const double dd_exp2_tab_dev[64][2] = FTS_DD_EXP2_TAB_VALUES;
static __device__ const double dd_log_tab_dev[97][2] = FTS_DD_LOG_TAB_VALUES;
#endif
FTS_HD dd dd_exp2_tab(int j)
{
#ifdef __CUDA_ARCH__
    const double2 v = __ldg(reinterpret_cast<const double2*>(&dd_exp2_tab_dev[j][0]));
    return dd(v.x, v.y);
#else
    return dd(dd_exp2_tab_host[j][0], dd_exp2_tab_host[j][1]);
#endif
}
FTS_HD dd dd_log_tab(int j)
{
#ifdef __CUDA_ARCH__
    const double2 v = __ldg(reinterpret_cast<const double2*>(&dd_log_tab_dev[j][0]));
    return dd(v.x, v.y);
#else
    return dd(dd_log_tab_host[j][0], dd_log_tab_host[j][1]);
#endif
}

// a + b for |a.hi| >= |b.hi| (or a.hi == 0): 8 flops instead of 20
FTS_HD dd dd_add_big(dd a, dd b)
{
    dd s = quick_two_sum(a.hi, b.hi);
    s.lo += a.lo + b.lo;
    return quick_two_sum(s.hi, s.lo);
}

// exp(dd): x = n ln2/64 + r, |r| <= ln2/128; exp(r) = 1 + r + r^2 p(r) with the terms r^2/2! ... r^6/6! in
// double-double and r^7/7! ... r^13/13! in double; exp(x) = 2^(n >> 6) * 2^((n & 63)/64) * exp(r).
FTS_HD dd dd_exp(dd a)
{
    // ln2/64 = H + M: H carries 30 bits, so n H is exact for |n| < 2^23 and the reduction loses nothing at |x| ~ 700
    const double LN2_64_H = 0x1.62e42fe800000p-7;
    const dd LN2_64_M(0x1.e8e7bcd5e4f1ep-37, -0x1.8cff81a12a17ep-91);
    if (a.hi != a.hi) return dd(a.hi);
    if (a.hi < -745.2) return dd(0.0);  // down to the smallest subnormal (ldexp below rounds once)
    if (a.hi >= 709.0) return dd(dd_inf());
    if (a.hi == 0.0 && a.lo == 0.0) return dd(1.0);
    const double nf = ::floor(a.hi * 0x1.71547652b82fep+6 + 0.5);
    const dd r = (a + (-(LN2_64_H * nf))) - LN2_64_M * nf;
    const int n = (int)nf;
    const double rh = r.hi;
    double q = 0x1.6124613a86d09p-33;
    q = ::fma(q, rh, 0x1.1eed8eff8d898p-29);
    q = ::fma(q, rh, 0x1.ae64567f544e4p-26);
    q = ::fma(q, rh, 0x1.27e4fb7789f5cp-22);
    q = ::fma(q, rh, 0x1.71de3a556c734p-19);
    q = ::fma(q, rh, 0x1.a01a01a01a01ap-16);
    q = ::fma(q, rh, 0x1.a01a01a01a01ap-13);
    dd p = dd_add_big(dd(0x1.6c16c16c16c17p-10, -0x1.f49f49f49f49fp-65), r * q);
    p = dd_add_big(dd(0x1.1111111111111p-7, 0x1.1111111111111p-63), p * r);
    p = dd_add_big(dd(0x1.5555555555555p-5, 0x1.5555555555555p-59), p * r);
    p = dd_add_big(dd(0x1.5555555555555p-3, 0x1.5555555555555p-57), p * r);
    p = dd_add_big(dd(0.5), p * r);
    const dd s = dd_add_big(r, (r * r) * p);
    const dd e = dd_add_big(dd(1.0), s) * dd_exp2_tab(n & 63);
    return dd_ldexp(e, n >> 6);
}

// 1/d to ~2^-50 for d in [1, 4) (the quotient digits of dd_log's division need no more)
FTS_HD double dd_rcp_approx(double d)
{
#ifdef __CUDA_ARCH__
    double x;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
    double e = ::fma(-d, x, 1.0);
    x = ::fma(x, e, x);
    e = ::fma(-d, x, 1.0);
    return ::fma(x, e, x);
#else
    return 1.0 / d;
#endif
}

// log(dd): a = m 2^e with m in [0.75, 1.5); c = the nearest of the 97 points 0.75 + j/128 (c = 1 for m near 1, so
// that log keeps its relative accuracy there); s = (m - c)/(m + c), |s| < 1/383;
// log m = log c + 2 atanh(s) = log c + 2 (s + s^3/3 + s^5/5 + ... + s^15/15), the terms from s^7 on in double.
FTS_HD dd dd_log(dd a)
{
    if (a.hi <= 0.0)
        return dd(a.hi == 0.0 ? (-dd_inf()) : dd_nan());
    if (a.hi - a.hi != 0.0) return a;  // +Inf
    const dd LN2(6.931471805599452862e-01, 2.319046813846299558e-17);
    int e = 0;
    double mh = ::frexp(a.hi, &e);     // [0.5, 1); subnormal complement distances down to 1e-310 included
    if (mh < 0.75) {
        mh *= 2.0;
        e -= 1;
    }
    const double ml = ::ldexp(a.lo, -e);
    const int j = (int)::floor((mh - 0.75) * 128.0 + 0.5);
    const double c = 0.75 + (double)j * 0.0078125;
    const dd num = two_sum(mh - c, ml);          // mh - c is exact (Sterbenz)
    dd den = two_sum(mh, c);
    den.lo += ml;
    den = quick_two_sum(den.hi, den.lo);
    // s = num / den, three quotient digits off one approximate reciprocal
    const double rcp = dd_rcp_approx(den.hi);
    const double q1 = num.hi * rcp;
    dd rem = num - den * q1;
    const double q2 = rem.hi * rcp;
    rem = rem - den * q2;
    const double q3 = rem.hi * rcp;
    const dd s = dd_add_big(quick_two_sum(q1, q2), dd(q3));
    const dd t = s * s;
    const double th = t.hi;
    double q = 0x1.1111111111111p-4;
    q = ::fma(q, th, 0x1.3b13b13b13b14p-4);
    q = ::fma(q, th, 0x1.745d1745d1746p-4);
    q = ::fma(q, th, 0x1.c71c71c71c71cp-4);
    q = ::fma(q, th, 0x1.2492492492492p-3);
    dd u = dd_add_big(dd(0x1.999999999999ap-3, -0x1.999999999999ap-57), dd(th * q));
    u = dd_add_big(dd(0x1.5555555555555p-2, 0x1.5555555555555p-56), t * u);
    const dd lg = dd_mul_pwr2(dd_add_big(s, s * (t * u)), 2.0);
    if (e == 0) return j == 32 ? lg : dd_log_tab(j) + lg;
    return (LN2 * (double)e + dd_log_tab(j)) + lg;
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
