/*
 * fts_oracle_kernels.h — TEST INFRASTRUCTURE ONLY (see fts_oracle_prims.h).
 * Type-generic quadrature kernels of the CPU oracle, included once per (T, ACC) pair:
 *   ACC == T          K(name) = name_<t>      the reference's own sequential rounding
 *   ACC == __float128 K(name) = name_<t>x     exact sum of the same T-rounded terms
 */
/* ---------------------------------------------------------------- fixed grids -------------- */
/* src/integrate1D.jl:90-102 integrate1D(f, low, up, x, w, h)  (== :74-82 on [-1,1]) */
T K(fts_or_integrate1d)(int fam, const T* p, void* cb, T low, T up, const T* x, const T* w, int n, T h)
{
    S(fts_or_fn) f = {fam, p, cb};
    T half = (T)1 / ((T)1 + (T)1);
    T dx = half * (up - low), x0 = half * (up + low); /* core.jl:27-30 */
    ACC s = (ACC)(S(half_pi)() * S(F1)(&f, x0));
    for (int i = 0; i < n; ++i) {
        T xp = x0 + dx * x[i];
        T xm = x0 - dx * x[i];
#if ACC_IS_T
        s += w[i] * (S(F1)(&f, xm) + S(F1)(&f, xp));
#else
        s += (ACC)w[i] * ((ACC)S(F1)(&f, xm) + (ACC)S(F1)(&f, xp));
#endif
    }
#if ACC_IS_T
    return dx * h * s;
#else
    return (T)((ACC)dx * (ACC)h * s);
#endif
}

/* src/integrate1D.jl:49-65 integrate1D_cmpl(T, f, N): f(x, 1-x, 1+x) over [-1,1] on the grid
 * (x, w, h) = tanhsinh(T, N); the complements are formed by subtraction */
T K(fts_or_integrate1d_cmpl)(int fam, const T* p, void* cb, const T* x, const T* w, int n, T h)
{
    S(fts_or_fn) f = {fam, p, cb};
    T one = (T)1;
    ACC s = (ACC)(S(half_pi)() * S(FC1)(&f, (T)0, one, one));
    for (int i = 0; i < n; ++i) {
        T xi = x[i];
        T one_minus_x = one - xi, one_plus_x = one + xi;
#if ACC_IS_T
        s += w[i] * (S(FC1)(&f, xi, one_minus_x, one_plus_x) + S(FC1)(&f, -xi, one_plus_x, one_minus_x));
#else
        s += (ACC)w[i] * ((ACC)S(FC1)(&f, xi, one_minus_x, one_plus_x) + (ACC)S(FC1)(&f, -xi, one_plus_x, one_minus_x));
#endif
    }
#if ACC_IS_T
    return h * s;
#else
    return (T)((ACC)h * s);
#endif
}

/* src/integrate2D.jl:91-124 */
T K(fts_or_integrate2d)(int fam, const T* p, void* cb, const T* low, const T* up, const T* x,
                        const T* w, int n, T h)
{
    S(fts_or_fn) f = {fam, p, cb};
    T half = (T)1 / ((T)1 + (T)1);
    T dx = half * (up[0] - low[0]), x0 = half * (up[0] + low[0]);
    T dy = half * (up[1] - low[1]), y0 = half * (up[1] + low[1]);
    T w0 = S(half_pi)();
    T w02 = w0 * w0;
    ACC s = (ACC)(w02 * S(F2)(&f, x0, y0));
    for (int k = 0; k < n; ++k) {
        T wk = w[k];
        T xkp = dx * x[k] + x0, xkm = x0 - dx * x[k];
        T ykp = dy * x[k] + y0, ykm = y0 - dy * x[k];
#if ACC_IS_T
        s += wk * w0 * (S(F2)(&f, xkp, y0) + S(F2)(&f, xkm, y0) + S(F2)(&f, x0, ykp) + S(F2)(&f, x0, ykm));
#else
        s += (ACC)wk * (ACC)w0 * ((ACC)S(F2)(&f, xkp, y0) + (ACC)S(F2)(&f, xkm, y0) +
                                  (ACC)S(F2)(&f, x0, ykp) + (ACC)S(F2)(&f, x0, ykm));
#endif
    }
    for (int i = 0; i < n; ++i) {
        T wi = w[i];
        T xip = dx * x[i] + x0, xim = x0 - dx * x[i];
        ACC inner = 0;
        for (int j = 0; j < n; ++j) {
            T wj = w[j];
            T yjp = dy * x[j] + y0, yjm = y0 - dy * x[j];
#if ACC_IS_T
            inner += wj * (S(F2)(&f, xim, yjm) + S(F2)(&f, xip, yjm) + S(F2)(&f, xim, yjp) + S(F2)(&f, xip, yjp));
#else
            inner += (ACC)wj * ((ACC)S(F2)(&f, xim, yjm) + (ACC)S(F2)(&f, xip, yjm) +
                                (ACC)S(F2)(&f, xim, yjp) + (ACC)S(F2)(&f, xip, yjp));
#endif
        }
        s += (ACC)wi * inner;
    }
#if ACC_IS_T
    return dx * dy * (h * h) * s;
#else
    return (T)((ACC)dx * (ACC)dy * ((ACC)h * (ACC)h) * s);
#endif
}

/* src/integrate3D.jl:150-207 */
T K(fts_or_integrate3d)(int fam, const T* p, void* cb, const T* low, const T* up, const T* x,
                        const T* w, int n, T h)
{
    S(fts_or_fn) f = {fam, p, cb};
    T half = (T)1 / ((T)1 + (T)1);
    T dx = half * (up[0] - low[0]), x0 = half * (up[0] + low[0]);
    T dy = half * (up[1] - low[1]), y0 = half * (up[1] + low[1]);
    T dz = half * (up[2] - low[2]), z0 = half * (up[2] + low[2]);
    T w0 = S(half_pi)();
    T w02 = w0 * w0;
    T w03 = w02 * w0;
    ACC total = (ACC)(w03 * S(F3)(&f, x0, y0, z0));
    for (int i = 0; i < n; ++i) {
        T wi = w[i];
#if ACC_IS_T
        T axis = (S(F3)(&f, dx * x[i] + x0, y0, z0) + S(F3)(&f, x0 - dx * x[i], y0, z0)) +
                 (S(F3)(&f, x0, dy * x[i] + y0, z0) + S(F3)(&f, x0, y0 - dy * x[i], z0)) +
                 (S(F3)(&f, x0, y0, dz * x[i] + z0) + S(F3)(&f, x0, y0, z0 - dz * x[i]));
        total += wi * w02 * axis;
#else
        ACC axis = ((ACC)S(F3)(&f, dx * x[i] + x0, y0, z0) + (ACC)S(F3)(&f, x0 - dx * x[i], y0, z0)) +
                   ((ACC)S(F3)(&f, x0, dy * x[i] + y0, z0) + (ACC)S(F3)(&f, x0, y0 - dy * x[i], z0)) +
                   ((ACC)S(F3)(&f, x0, y0, dz * x[i] + z0) + (ACC)S(F3)(&f, x0, y0, z0 - dz * x[i]));
        total += (ACC)wi * (ACC)w02 * axis;
#endif
    }
    for (int i = 0; i < n; ++i) {
        T wi = w[i];
        T xp = dx * x[i] + x0, xm = x0 - dx * x[i];
        T ypi = dy * x[i] + y0, ymi = y0 - dy * x[i];
        for (int j = 0; j < n; ++j) {
            T wj = w[j];
            T yp = dy * x[j] + y0, ym = y0 - dy * x[j];
            T zpj = dz * x[j] + z0, zmj = z0 - dz * x[j];
#if ACC_IS_T
            T wiwj = wi * wj;
            T plane = (S(F3)(&f, xp, yp, z0) + S(F3)(&f, xm, yp, z0) + S(F3)(&f, xp, ym, z0) + S(F3)(&f, xm, ym, z0)) +
                      (S(F3)(&f, xp, y0, zpj) + S(F3)(&f, xm, y0, zpj) + S(F3)(&f, xp, y0, zmj) + S(F3)(&f, xm, y0, zmj)) +
                      (S(F3)(&f, x0, ypi, zpj) + S(F3)(&f, x0, ymi, zpj) + S(F3)(&f, x0, ypi, zmj) + S(F3)(&f, x0, ymi, zmj));
            total += wiwj * w0 * plane;
            T oct = 0;
            for (int k = 0; k < n; ++k) {
                T wk = w[k];
                T zp = dz * x[k] + z0, zm = z0 - dz * x[k];
                oct += wk * ((S(F3)(&f, xp, yp, zp) + S(F3)(&f, xm, yp, zp) + S(F3)(&f, xp, ym, zp) + S(F3)(&f, xm, ym, zp)) +
                             (S(F3)(&f, xp, yp, zm) + S(F3)(&f, xm, yp, zm) + S(F3)(&f, xp, ym, zm) + S(F3)(&f, xm, ym, zm)));
            }
            total += wiwj * oct;
#else
            ACC wiwj = (ACC)wi * (ACC)wj;
            ACC plane = ((ACC)S(F3)(&f, xp, yp, z0) + (ACC)S(F3)(&f, xm, yp, z0) + (ACC)S(F3)(&f, xp, ym, z0) + (ACC)S(F3)(&f, xm, ym, z0)) +
                        ((ACC)S(F3)(&f, xp, y0, zpj) + (ACC)S(F3)(&f, xm, y0, zpj) + (ACC)S(F3)(&f, xp, y0, zmj) + (ACC)S(F3)(&f, xm, y0, zmj)) +
                        ((ACC)S(F3)(&f, x0, ypi, zpj) + (ACC)S(F3)(&f, x0, ymi, zpj) + (ACC)S(F3)(&f, x0, ypi, zmj) + (ACC)S(F3)(&f, x0, ymi, zmj));
            total += wiwj * (ACC)w0 * plane;
            ACC oct = 0;
            for (int k = 0; k < n; ++k) {
                T zp = dz * x[k] + z0, zm = z0 - dz * x[k];
                oct += (ACC)w[k] * (((ACC)S(F3)(&f, xp, yp, zp) + (ACC)S(F3)(&f, xm, yp, zp) + (ACC)S(F3)(&f, xp, ym, zp) + (ACC)S(F3)(&f, xm, ym, zp)) +
                                    ((ACC)S(F3)(&f, xp, yp, zm) + (ACC)S(F3)(&f, xm, yp, zm) + (ACC)S(F3)(&f, xp, ym, zm) + (ACC)S(F3)(&f, xm, ym, zm)));
            }
            total += wiwj * oct;
#endif
        }
    }
#if ACC_IS_T
    return (dx * dy * dz * (h * h * h)) * total;
#else
    return (T)(((ACC)dx * (ACC)dy * (ACC)dz * ((ACC)h * (ACC)h * (ACC)h)) * total);
#endif
}

/* ---------------------------------------------------------------- adaptive ----------------- */
typedef struct {
    T value;
    T err;
    int level;     /* level at which the returned value was computed */
    int converged; /* stop rule met */
    long evals;    /* integrand evaluations performed */
} K(fts_or_result);

/* core.jl:49 */
static inline T K(err_target)(T I, T rtol, T atol)
{
    T r = rtol * M(fabs)(I);
    /* Julia's max propagates NaN */
    if (r != r) return r;
    return atol > r ? atol : r;
}

/* ACC-typed scale/stop helper shared by all adaptive kernels */
#if ACC_IS_T
#define OR_SCALE1(dx, h, s) ((dx) * (h) * (s))
#define OR_SCALE2(dx, dy, h, s) ((dx) * (dy) * ((h) * (h)) * (s))
#define OR_SCALE3(dx, dy, dz, h, s) ((dx) * (dy) * (dz) * ((h) * (h) * (h)) * (s))
#else
#define OR_SCALE1(dx, h, s) ((T)((ACC)(dx) * (ACC)(h) * (s)))
#define OR_SCALE2(dx, dy, h, s) ((T)((ACC)(dx) * (ACC)(dy) * ((ACC)(h) * (ACC)(h)) * (s)))
#define OR_SCALE3(dx, dy, dz, h, s) ((T)((ACC)(dx) * (ACC)(dy) * (ACC)(dz) * ((ACC)(h) * (ACC)(h) * (ACC)(h)) * (s)))
#endif

/* src/adaptive.jl:28-75  adaptive_integrate_1D (typed core) */
void K(fts_or_adaptive1d)(int fam, const T* p, void* cb, T a, T b, T rtol, T atol, int max_levels,
                          T tm, const T* ix, const T* iw, const T* xs, const T* ws,
                          K(fts_or_result)* out)
{
    S(fts_or_fn) f = {fam, p, cb};
    T half = (T)1 / ((T)1 + (T)1);
    T dx = half * (b - a), x0 = half * (b + a);
    T h = tm * half;
    T w0 = S(pi)() * half;
    long evals = 1;
    ACC s_total = (ACC)(w0 * S(F1)(&f, x0));
    for (int i = 0; i < 2; ++i) {
        T xk = ix[i];
#if ACC_IS_T
        s_total += iw[i] * (S(F1)(&f, dx * xk + x0) + S(F1)(&f, x0 - dx * xk));
#else
        s_total += (ACC)iw[i] * ((ACC)S(F1)(&f, dx * xk + x0) + (ACC)S(F1)(&f, x0 - dx * xk));
#endif
        evals += 2;
    }
    T old_res = OR_SCALE1(dx, h, s_total);
    T err = 0;
    out->level = 0;
    out->converged = 0;
    for (int level = 1; level <= max_levels; ++level) {
        h *= half;
        ACC s_new = 0;
        long n = 1L << level, off = n - 2;
        for (long i = 0; i < n; ++i) {
            T xk = xs[off + i];
            T xp = dx * xk + x0;
            T xm = x0 - dx * xk;
#if ACC_IS_T
            s_new += ws[off + i] * (S(F1)(&f, xp) + S(F1)(&f, xm));
#else
            s_new += (ACC)ws[off + i] * ((ACC)S(F1)(&f, xp) + (ACC)S(F1)(&f, xm));
#endif
        }
        evals += 2 * n;
        s_total += s_new;
        T new_res = OR_SCALE1(dx, h, s_total);
        err = M(fabs)(new_res - old_res);
        out->level = level;
        if (err <= K(err_target)(new_res, rtol, atol)) {
            out->value = new_res;
            out->err = err;
            out->converged = 1;
            out->evals = evals;
            return;
        }
        old_res = new_res;
    }
    out->value = old_res;
    out->err = err;
    out->evals = evals;
}

/* src/adaptive.jl:727-785  adaptive_integrate_1D_cmpl (typed core) */
void K(fts_or_adaptive1d_cmpl)(int fam, const T* p, void* cb, T a, T b, T rtol, T atol,
                               int max_levels, T tm, const T* ix, const T* iw, const T* ic,
                               const T* xs, const T* ws, const T* cs, K(fts_or_result)* out)
{
    S(fts_or_fn) f = {fam, p, cb};
    T half = (T)1 / ((T)1 + (T)1);
    T one = (T)1;
    T dx = half * (b - a), x0 = half * (b + a);
    T h = tm * half;
    T w0 = S(half_pi)();
    long evals = 1;
    ACC s_total = (ACC)(w0 * S(FC1)(&f, x0, dx, dx));
    for (int i = 0; i < 2; ++i) {
        T xk = ix[i], ck = ic[i], wk = iw[i];
        T dxck = dx * ck;
        T dx1pxk = dx * (one + xk);
#if ACC_IS_T
        s_total += wk * (S(FC1)(&f, dx * xk + x0, dxck, dx1pxk) + S(FC1)(&f, x0 - dx * xk, dx1pxk, dxck));
#else
        s_total += (ACC)wk * ((ACC)S(FC1)(&f, dx * xk + x0, dxck, dx1pxk) + (ACC)S(FC1)(&f, x0 - dx * xk, dx1pxk, dxck));
#endif
        evals += 2;
    }
    T old_res = OR_SCALE1(dx, h, s_total);
    T err = 0;
    out->level = 0;
    out->converged = 0;
    for (int level = 1; level <= max_levels; ++level) {
        h *= half;
        ACC s_new = 0;
        long n = 1L << level, off = n - 2;
        for (long i = 0; i < n; ++i) {
            T xk = xs[off + i], ck = cs[off + i];
            T dxck = dx * ck;
            T dx1pxk = dx * (one + xk);
#if ACC_IS_T
            s_new += ws[off + i] * (S(FC1)(&f, dx * xk + x0, dxck, dx1pxk) + S(FC1)(&f, x0 - dx * xk, dx1pxk, dxck));
#else
            s_new += (ACC)ws[off + i] * ((ACC)S(FC1)(&f, dx * xk + x0, dxck, dx1pxk) + (ACC)S(FC1)(&f, x0 - dx * xk, dx1pxk, dxck));
#endif
        }
        evals += 2 * n;
        s_total += s_new;
        T new_res = OR_SCALE1(dx, h, s_total);
        err = M(fabs)(new_res - old_res);
        out->level = level;
        if (err <= K(err_target)(new_res, rtol, atol)) {
            out->value = new_res;
            out->err = err;
            out->converged = 1;
            out->evals = evals;
            return;
        }
        old_res = new_res;
    }
    out->value = old_res;
    out->err = err;
    out->evals = evals;
}

/* helpers of adaptive_integrate_2D (adaptive.jl:165-176) */
static inline ACC K(quadrants)(const S(fts_or_fn)* f, T dx, T x0, T dy, T y0, T xi, T yi, T wi, T wj)
{
    T xp = dx * xi + x0, xm = x0 - dx * xi;
    T yp = dy * yi + y0, ym = y0 - dy * yi;
#if ACC_IS_T
    return wi * wj * (S(F2)(f, xp, yp) + S(F2)(f, xm, yp) + S(F2)(f, xp, ym) + S(F2)(f, xm, ym));
#else
    return (ACC)wi * (ACC)wj * ((ACC)S(F2)(f, xp, yp) + (ACC)S(F2)(f, xm, yp) + (ACC)S(F2)(f, xp, ym) + (ACC)S(F2)(f, xm, ym));
#endif
}
static inline ACC K(axes2)(const S(fts_or_fn)* f, T dx, T x0, T dy, T y0, T w0, T val, T wk)
{
    T xp = dx * val + x0, xm = x0 - dx * val;
    T yp = dy * val + y0, ym = y0 - dy * val;
#if ACC_IS_T
    return wk * w0 * (S(F2)(f, xp, y0) + S(F2)(f, xm, y0) + S(F2)(f, x0, yp) + S(F2)(f, x0, ym));
#else
    return (ACC)wk * (ACC)w0 * ((ACC)S(F2)(f, xp, y0) + (ACC)S(F2)(f, xm, y0) + (ACC)S(F2)(f, x0, yp) + (ACC)S(F2)(f, x0, ym));
#endif
}

/* src/adaptive.jl:154-224  adaptive_integrate_2D (typed core) */
void K(fts_or_adaptive2d)(int fam, const T* p, void* cb, const T* low, const T* up, T rtol, T atol,
                          int max_levels, T tm, const T* ix, const T* iw, const T* xs, const T* ws,
                          K(fts_or_result)* out)
{
    S(fts_or_fn) f = {fam, p, cb};
    T half = (T)1 / ((T)1 + (T)1);
    T dx = half * (up[0] - low[0]), x0 = half * (up[0] + low[0]);
    T dy = half * (up[1] - low[1]), y0 = half * (up[1] + low[1]);
    T h = tm * half;
    T w0 = S(half_pi)();
    T w02 = w0 * w0;
    long evals = 1;
    ACC s_total = (ACC)(w02 * S(F2)(&f, x0, y0));
    for (int i = 0; i < 2; ++i) {
        for (int j = 0; j < 2; ++j) s_total += K(quadrants)(&f, dx, x0, dy, y0, ix[i], ix[j], iw[i], iw[j]);
        s_total += K(axes2)(&f, dx, x0, dy, y0, w0, ix[i], iw[i]);
        evals += 12;
    }
    T old_res = OR_SCALE2(dx, dy, h, s_total);
    T err = 0;
    out->level = 0;
    out->converged = 0;
    for (int level = 1; level <= max_levels; ++level) {
        h *= half;
        ACC s_new = 0;
        long n = 2L << level, off = n - 4;
        const T* xl = xs + off;
        const T* wl = ws + off;
        for (long i = 1; i <= n; ++i) {
            T xi = xl[i - 1], wi = wl[i - 1];
            for (long j = 1; j <= n; ++j) {
                if (!(i & 1) && !(j & 1)) continue;
                s_new += K(quadrants)(&f, dx, x0, dy, y0, xi, xl[j - 1], wi, wl[j - 1]);
                evals += 4;
            }
            if (i & 1) {
                s_new += K(axes2)(&f, dx, x0, dy, y0, w0, xi, wi);
                evals += 4;
            }
        }
        s_total += s_new;
        T new_res = OR_SCALE2(dx, dy, h, s_total);
        err = M(fabs)(new_res - old_res);
        out->level = level;
        if (err <= K(err_target)(new_res, rtol, atol)) {
            out->value = new_res;
            out->err = err;
            out->converged = 1;
            out->evals = evals;
            return;
        }
        old_res = new_res;
    }
    out->value = old_res;
    out->err = err;
    out->evals = evals;
}

/* Tensor-complement adaptive 2D: NOT in the reference (quad_cmpl is 1D only, quad.jl:298-318).
 * Defined as the tensor product of the reference's 1D complement rule (adaptive.jl:764-769
 * coordinates applied per axis) on a tensor level table built with the complement window
 * t_w_max(T,1); refinement bookkeeping identical to adaptive_integrate_2D.  cs holds the
 * complements of xs. */
static inline ACC K(quadrants_c)(const S(fts_or_fn)* f, T dx, T x0, T dy, T y0, T xi, T ci, T yj, T cj, T wi, T wj)
{
    T one = (T)1;
    T xp = dx * xi + x0, xm = x0 - dx * xi;
    T yp = dy * yj + y0, ym = y0 - dy * yj;
    T bx_p = dx * ci, ax_p = dx * (one + xi); /* at xp: b-x, x-a ; swapped at xm */
    T dy_p = dy * cj, cy_p = dy * (one + yj);
    T f1 = S(evalc2)(f->fam, xp, yp, bx_p, ax_p, dy_p, cy_p, f->p);
    T f2 = S(evalc2)(f->fam, xm, yp, ax_p, bx_p, dy_p, cy_p, f->p);
    T f3 = S(evalc2)(f->fam, xp, ym, bx_p, ax_p, cy_p, dy_p, f->p);
    T f4 = S(evalc2)(f->fam, xm, ym, ax_p, bx_p, cy_p, dy_p, f->p);
#if ACC_IS_T
    return wi * wj * (f1 + f2 + f3 + f4);
#else
    return (ACC)wi * (ACC)wj * ((ACC)f1 + (ACC)f2 + (ACC)f3 + (ACC)f4);
#endif
}
static inline ACC K(axes2_c)(const S(fts_or_fn)* f, T dx, T x0, T dy, T y0, T w0, T v, T c, T wk)
{
    T one = (T)1;
    T xp = dx * v + x0, xm = x0 - dx * v;
    T yp = dy * v + y0, ym = y0 - dy * v;
    T bx = dx * c, ax = dx * (one + v);
    T dyc = dy * c, cy = dy * (one + v);
    T f1 = S(evalc2)(f->fam, xp, y0, bx, ax, dy, dy, f->p);
    T f2 = S(evalc2)(f->fam, xm, y0, ax, bx, dy, dy, f->p);
    T f3 = S(evalc2)(f->fam, x0, yp, dx, dx, dyc, cy, f->p);
    T f4 = S(evalc2)(f->fam, x0, ym, dx, dx, cy, dyc, f->p);
#if ACC_IS_T
    return wk * w0 * (f1 + f2 + f3 + f4);
#else
    return (ACC)wk * (ACC)w0 * ((ACC)f1 + (ACC)f2 + (ACC)f3 + (ACC)f4);
#endif
}

void K(fts_or_adaptive2d_cmpl)(int fam, const T* p, const T* low, const T* up, T rtol, T atol,
                               int max_levels, T tm, const T* ix, const T* iw, const T* ic,
                               const T* xs, const T* ws, const T* cs, K(fts_or_result)* out)
{
    S(fts_or_fn) f = {fam, p, NULL};
    T half = (T)1 / ((T)1 + (T)1);
    T dx = half * (up[0] - low[0]), x0 = half * (up[0] + low[0]);
    T dy = half * (up[1] - low[1]), y0 = half * (up[1] + low[1]);
    T h = tm * half;
    T w0 = S(half_pi)();
    T w02 = w0 * w0;
    long evals = 1;
    ACC s_total = (ACC)(w02 * S(evalc2)(fam, x0, y0, dx, dx, dy, dy, p));
    for (int i = 0; i < 2; ++i) {
        for (int j = 0; j < 2; ++j)
            s_total += K(quadrants_c)(&f, dx, x0, dy, y0, ix[i], ic[i], ix[j], ic[j], iw[i], iw[j]);
        s_total += K(axes2_c)(&f, dx, x0, dy, y0, w0, ix[i], ic[i], iw[i]);
        evals += 12;
    }
    T old_res = OR_SCALE2(dx, dy, h, s_total);
    T err = 0;
    out->level = 0;
    out->converged = 0;
    for (int level = 1; level <= max_levels; ++level) {
        h *= half;
        ACC s_new = 0;
        long n = 2L << level, off = n - 4;
        const T *xl = xs + off, *wl = ws + off, *cl = cs + off;
        for (long i = 1; i <= n; ++i) {
            for (long j = 1; j <= n; ++j) {
                if (!(i & 1) && !(j & 1)) continue;
                s_new += K(quadrants_c)(&f, dx, x0, dy, y0, xl[i - 1], cl[i - 1], xl[j - 1], cl[j - 1], wl[i - 1], wl[j - 1]);
                evals += 4;
            }
            if (i & 1) {
                s_new += K(axes2_c)(&f, dx, x0, dy, y0, w0, xl[i - 1], cl[i - 1], wl[i - 1]);
                evals += 4;
            }
        }
        s_total += s_new;
        T new_res = OR_SCALE2(dx, dy, h, s_total);
        err = M(fabs)(new_res - old_res);
        out->level = level;
        if (err <= K(err_target)(new_res, rtol, atol)) {
            out->value = new_res;
            out->err = err;
            out->converged = 1;
            out->evals = evals;
            return;
        }
        old_res = new_res;
    }
    out->value = old_res;
    out->err = err;
    out->evals = evals;
}

/* helpers of adaptive_integrate_3D (adaptive.jl:365-403) */
typedef struct {
    T dx, x0, dy, y0, dz, z0, w0, w02;
} K(box3);

static inline ACC K(octant)(const S(fts_or_fn)* f, const K(box3)* b, T vi, T vj, T vk, T wi, T wj, T wk)
{
    T xp = b->dx * vi + b->x0, xm = b->x0 - b->dx * vi;
    T yp = b->dy * vj + b->y0, ym = b->y0 - b->dy * vj;
    T zp = b->dz * vk + b->z0, zm = b->z0 - b->dz * vk;
#if ACC_IS_T
    T w = wi * wj * wk;
    return w * ((S(F3)(f, xp, yp, zp) + S(F3)(f, xm, yp, zp) + S(F3)(f, xp, ym, zp) + S(F3)(f, xm, ym, zp)) +
                (S(F3)(f, xp, yp, zm) + S(F3)(f, xm, yp, zm) + S(F3)(f, xp, ym, zm) + S(F3)(f, xm, ym, zm)));
#else
    ACC w = (ACC)wi * (ACC)wj * (ACC)wk;
    return w * (((ACC)S(F3)(f, xp, yp, zp) + (ACC)S(F3)(f, xm, yp, zp) + (ACC)S(F3)(f, xp, ym, zp) + (ACC)S(F3)(f, xm, ym, zp)) +
                ((ACC)S(F3)(f, xp, yp, zm) + (ACC)S(F3)(f, xm, yp, zm) + (ACC)S(F3)(f, xp, ym, zm) + (ACC)S(F3)(f, xm, ym, zm)));
#endif
}

static inline ACC K(planes)(const S(fts_or_fn)* f, const K(box3)* b, T vi, T vj, T wi, T wj)
{
    T x0 = b->x0, y0 = b->y0, z0 = b->z0;
    T xp = b->dx * vi + x0, xm = x0 - b->dx * vi;
    T ypi = b->dy * vi + y0, ymi = y0 - b->dy * vi;
    T ypj = b->dy * vj + y0, ymj = y0 - b->dy * vj;
    T zpj = b->dz * vj + z0, zmj = z0 - b->dz * vj;
#if ACC_IS_T
    T w = wi * wj * b->w0;
    return w * ((S(F3)(f, xp, ypj, z0) + S(F3)(f, xm, ypj, z0) + S(F3)(f, xp, ymj, z0) + S(F3)(f, xm, ymj, z0)) +
                (S(F3)(f, xp, y0, zpj) + S(F3)(f, xm, y0, zpj) + S(F3)(f, xp, y0, zmj) + S(F3)(f, xm, y0, zmj)) +
                (S(F3)(f, x0, ypi, zpj) + S(F3)(f, x0, ymi, zpj) + S(F3)(f, x0, ypi, zmj) + S(F3)(f, x0, ymi, zmj)));
#else
    ACC w = (ACC)wi * (ACC)wj * (ACC)b->w0;
    return w * (((ACC)S(F3)(f, xp, ypj, z0) + (ACC)S(F3)(f, xm, ypj, z0) + (ACC)S(F3)(f, xp, ymj, z0) + (ACC)S(F3)(f, xm, ymj, z0)) +
                ((ACC)S(F3)(f, xp, y0, zpj) + (ACC)S(F3)(f, xm, y0, zpj) + (ACC)S(F3)(f, xp, y0, zmj) + (ACC)S(F3)(f, xm, y0, zmj)) +
                ((ACC)S(F3)(f, x0, ypi, zpj) + (ACC)S(F3)(f, x0, ymi, zpj) + (ACC)S(F3)(f, x0, ypi, zmj) + (ACC)S(F3)(f, x0, ymi, zmj)));
#endif
}

static inline ACC K(axes3)(const S(fts_or_fn)* f, const K(box3)* b, T vi, T wi)
{
    T x0 = b->x0, y0 = b->y0, z0 = b->z0;
    T xp = b->dx * vi + x0, xm = x0 - b->dx * vi;
    T yp = b->dy * vi + y0, ym = y0 - b->dy * vi;
    T zp = b->dz * vi + z0, zm = z0 - b->dz * vi;
#if ACC_IS_T
    T w = wi * b->w02;
    return w * ((S(F3)(f, xp, y0, z0) + S(F3)(f, xm, y0, z0)) + (S(F3)(f, x0, yp, z0) + S(F3)(f, x0, ym, z0)) +
                (S(F3)(f, x0, y0, zp) + S(F3)(f, x0, y0, zm)));
#else
    ACC w = (ACC)wi * (ACC)b->w02;
    return w * (((ACC)S(F3)(f, xp, y0, z0) + (ACC)S(F3)(f, xm, y0, z0)) + ((ACC)S(F3)(f, x0, yp, z0) + (ACC)S(F3)(f, x0, ym, z0)) +
                ((ACC)S(F3)(f, x0, y0, zp) + (ACC)S(F3)(f, x0, y0, zm)));
#endif
}

/* level-0 sum of adaptive_integrate_3D (adaptive.jl:406-418), exposed for the sharded tests */
ACC K(fts_or_adaptive3d_level0)(const S(fts_or_fn)* f, const K(box3)* bx, const T* ix, const T* iw)
{
    T w03 = bx->w02 * bx->w0;
    ACC s_total = (ACC)(w03 * S(F3)(f, bx->x0, bx->y0, bx->z0));
    for (int i = 0; i < 2; ++i) {
        for (int j = 0; j < 2; ++j) {
            for (int k = 0; k < 2; ++k) s_total += K(octant)(f, bx, ix[i], ix[j], ix[k], iw[i], iw[j], iw[k]);
            s_total += K(planes)(f, bx, ix[i], ix[j], iw[i], iw[j]);
        }
        s_total += K(axes3)(f, bx, ix[i], iw[i]);
    }
    return s_total;
}

/* new-point sum of one level (adaptive.jl:430-456); i restricted to i ≡ rank+1 (mod nranks)
 * for the outer-axis sharding tests (rank=0,nranks=1 -> the reference loop) */
ACC K(fts_or_adaptive3d_level_sum)(const S(fts_or_fn)* f, const K(box3)* bx, const T* xl, const T* wl, long n,
                                   int rank, int nranks)
{
    ACC s_new = 0;
    for (long i = 1 + rank; i <= n; i += nranks) {
        T xi = xl[i - 1], wi = wl[i - 1];
        if (i & 1) {
            for (long j = 1; j <= n; ++j) {
                T xj = xl[j - 1], wj = wl[j - 1];
                for (long k = 1; k <= n; ++k) s_new += K(octant)(f, bx, xi, xj, xl[k - 1], wi, wj, wl[k - 1]);
                s_new += K(planes)(f, bx, xi, xj, wi, wj);
            }
            s_new += K(axes3)(f, bx, xi, wi);
        } else {
            for (long j = 1; j <= n; ++j) {
                T xj = xl[j - 1], wj = wl[j - 1];
                if (j & 1) {
                    for (long k = 1; k <= n; ++k) s_new += K(octant)(f, bx, xi, xj, xl[k - 1], wi, wj, wl[k - 1]);
                    s_new += K(planes)(f, bx, xi, xj, wi, wj);
                } else {
                    for (long k = 1; k <= n; k += 2) s_new += K(octant)(f, bx, xi, xj, xl[k - 1], wi, wj, wl[k - 1]);
                }
            }
        }
    }
    return s_new;
}

/* src/adaptive.jl:352-471  adaptive_integrate_3D (typed core) */
void K(fts_or_adaptive3d)(int fam, const T* p, void* cb, const T* low, const T* up, T rtol, T atol,
                          int max_levels, T tm, const T* ix, const T* iw, const T* xs, const T* ws,
                          K(fts_or_result)* out)
{
    S(fts_or_fn) f = {fam, p, cb};
    T half = (T)1 / ((T)1 + (T)1);
    K(box3) bx;
    bx.dx = half * (up[0] - low[0]); bx.x0 = half * (up[0] + low[0]);
    bx.dy = half * (up[1] - low[1]); bx.y0 = half * (up[1] + low[1]);
    bx.dz = half * (up[2] - low[2]); bx.z0 = half * (up[2] + low[2]);
    bx.w0 = S(half_pi)();
    bx.w02 = bx.w0 * bx.w0;
    T h = tm * half;
    ACC s_total = K(fts_or_adaptive3d_level0)(&f, &bx, ix, iw);
    long evals = 125;
    T old_res = OR_SCALE3(bx.dx, bx.dy, bx.dz, h, s_total);
    T err = 0;
    out->level = 0;
    out->converged = 0;
    for (int level = 1; level <= max_levels; ++level) {
        h *= half;
        long n = 2L << level, off = n - 4;
        ACC s_new = K(fts_or_adaptive3d_level_sum)(&f, &bx, xs + off, ws + off, n, 0, 1);
        evals += 7 * n * n * n + 9 * n * n + 3 * n;
        s_total += s_new;
        T new_res = OR_SCALE3(bx.dx, bx.dy, bx.dz, h, s_total);
        err = M(fabs)(new_res - old_res);
        out->level = level;
        if (err <= K(err_target)(new_res, rtol, atol)) {
            out->value = new_res;
            out->err = err;
            out->converged = 1;
            out->evals = evals;
            return;
        }
        old_res = new_res;
    }
    out->value = old_res;
    out->err = err;
    out->evals = evals;
}

#undef OR_SCALE1
#undef OR_SCALE2
#undef OR_SCALE3
