// Row f1 of the scope table: the gamma-fit tail of DErand (pySpade/DErand.py:250-292) for all
// genes at once.  Per gene the reference calls scipy.stats.gamma.fit(-log p over the draws,
// finite values only): rv_continuous.fit -> start values from gamma_gen._fitstart (skewness ->
// shape) and rv_continuous._fit_loc_scale_support (moments -> loc, scale, pushed outside the
// data if needed), then scipy.optimize.fmin (Nelder-Mead, xtol = ftol = 1e-4, 600 iterations /
// 600 evaluations) on rv_continuous._penalized_nnlf.  This file restates exactly that
// procedure (same start values, same simplex moves, same penalty, same stopping rule) for one
// gene, so that the parameters agree with scipy's up to the sensitivity of the simplex path to
// last-bit differences of log / lgamma and of the summation order.  On the device one warp
// owns one gene: the lanes split the draws of every likelihood evaluation and all lanes carry
// the same simplex (WarpReduce); the same source compiles for the host with one "lane"
// (SerialReduce, tests/gamma_host_harness.cpp) to be checked against scipy without a GPU.
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define SPADE_HD __host__ __device__
#else
#define SPADE_HD
#endif

namespace spade {

constexpr double kLogXMax100 = 709.782712893384 * 100.0;      // scipy.stats._constants._LOGXMAX * 100

struct GammaColumn {                 // x_i = sign * tab[i * ld], non-finite values are skipped
    const double *tab;
    int64_t ld;
    int S;
    double sign;
    int lane = 0, nlanes = 1;        // this caller walks i = lane, lane + nlanes, ...
};

struct SerialReduce {                // one caller sees every value
    SPADE_HD double sum(double v) const { return v; }
    SPADE_HD int sum(int v) const { return v; }
    SPADE_HD double min(double v) const { return v; }
    SPADE_HD double max(double v) const { return v; }
};

#ifdef __CUDACC__
struct WarpReduce {                  // butterfly: every lane ends with the same bits
    __device__ double sum(double v) const
    {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        return v;
    }
    __device__ int sum(int v) const
    {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        return v;
    }
    __device__ double min(double v) const
    {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
        return v;
    }
    __device__ double max(double v) const
    {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
        return v;
    }
};
#endif

SPADE_HD inline bool gf_finite(double v) { return v - v == 0.0; }

// rv_continuous._penalized_nnlf for the gamma distribution (theta = shape a, loc, scale).
template <class Reduce>
SPADE_HD inline double gamma_penalized_nnlf(const GammaColumn &c, int n, double a, double loc, double scale,
                                            const Reduce &red)
{
    if (!(a > 0.0) || !(scale > 0.0)) return INFINITY;            // _argcheck, scale <= 0
    const double am1 = a - 1.0, lga = lgamma(a);
    double total = 0.0;
    int n_bad = 0;
    for (int i = c.lane; i < c.S; i += c.nlanes) {
        const double x = c.sign * c.tab[(int64_t)i * c.ld];
        if (!gf_finite(x)) continue;
        const double y = (x - loc) / scale;
        if (!(0.0 <= y && y <= INFINITY)) { ++n_bad; continue; }  // outside the (closed) support
        const double xl = (am1 == 0.0) ? 0.0 : am1 * log(y);      // sc.xlogy(a - 1, y)
        const double lp = xl - y - lga;
        if (gf_finite(lp)) total += lp; else ++n_bad;
    }
    total = red.sum(total);
    n_bad = red.sum(n_bad);
    return -total + (double)n_bad * kLogXMax100 + (double)n * log(scale);
}

// Status: 0 = fitted, 1 = skipped (<= 50 finite values: outputs untouched zeros),
//         2 = all finite values equal -> (nan, 0, 1), 3 = scipy would raise FitError -> nan.
// *evals (optional): likelihood evaluations the simplex used; 600 = scipy's budget exhausted (its
// fmin then stops with "maximum number of function evaluations exceeded" and returns the best
// vertex - the fits that are least reproducible across libm / summation-order differences).
template <class Reduce>
SPADE_HD inline int gamma_fit_column(const GammaColumn &c, double out[3], const Reduce &red, int *evals = nullptr)
{
    if (evals) *evals = 0;
    // ---- census: count, mean, min, max, all-equal (DErand.py:268-272) ----
    int n = 0;
    double sum = 0.0, lo = INFINITY, hi = -INFINITY;
    for (int i = c.lane; i < c.S; i += c.nlanes) {
        const double x = c.sign * c.tab[(int64_t)i * c.ld];
        if (!gf_finite(x)) continue;
        sum += x;
        lo = fmin(lo, x);
        hi = fmax(hi, x);
        ++n;
    }
    n = red.sum(n);
    sum = red.sum(sum);
    lo = red.min(lo);
    hi = red.max(hi);
    out[0] = out[1] = out[2] = 0.0;
    if (n <= 50) return 1;
    if (lo == hi) { out[0] = NAN; out[1] = 0.0; out[2] = 1.0; return 2; }     // all finite values equal

    // ---- start values: _skew, gamma_gen._fitstart, fit_loc_scale, _fit_loc_scale_support ----
    const double mu = sum / n;
    double m2 = 0.0, m3 = 0.0;
    for (int i = c.lane; i < c.S; i += c.nlanes) {
        const double x = c.sign * c.tab[(int64_t)i * c.ld];
        if (!gf_finite(x)) continue;
        const double d = x - mu;
        m2 += d * d;
        m3 += d * d * d;
    }
    m2 = red.sum(m2) / n;
    m3 = red.sum(m3) / n;
    const double sk = m3 / pow(m2, 1.5);
    const double a0 = 4.0 / (1e-8 + sk * sk);
    double scale0 = sqrt(m2 / a0);                               // Shat = sqrt(var / mu2), mu2 = a
    double loc0 = mu - scale0 * a0;                              // Lhat = mean - Shat * mu,  mu = a
    if (!gf_finite(loc0)) loc0 = 0.0;
    if (!(gf_finite(scale0) && 0.0 < scale0)) scale0 = 1.0;
    if (!(loc0 < lo && hi < INFINITY)) {                         // support [loc, inf) must hold the data
        loc0 = lo - (hi - lo) * 0.1;
        scale0 = 1.0;
    }

    // ---- scipy.optimize.fmin: Nelder-Mead, rho 1, chi 2, psi 0.5, sigma 0.5 ----
    const int N = 3, maxiter = 600, maxfun = 600;
    double sim[4][3], fsim[4];
    const double x0[3] = {a0, loc0, scale0};
    for (int k = 0; k < 4; ++k)
        for (int j = 0; j < 3; ++j) sim[k][j] = x0[j];
    for (int k = 0; k < 3; ++k) sim[k + 1][k] = (x0[k] != 0.0) ? 1.05 * x0[k] : 0.00025;
    int fcalls = 0;
    for (int k = 0; k < 4; ++k) { fsim[k] = gamma_penalized_nnlf(c, n, sim[k][0], sim[k][1], sim[k][2], red); ++fcalls; }

    auto sort_simplex = [&]() {                                  // np.argsort on 4 values: insertion sort
        for (int i = 1; i < 4; ++i) {
            const double f = fsim[i], p0 = sim[i][0], p1 = sim[i][1], p2 = sim[i][2];
            int j = i - 1;
            while (j >= 0 && (fsim[j] > f || (fsim[j] != fsim[j] && f == f))) {   // nan sorts last
                fsim[j + 1] = fsim[j];
                sim[j + 1][0] = sim[j][0]; sim[j + 1][1] = sim[j][1]; sim[j + 1][2] = sim[j][2];
                --j;
            }
            fsim[j + 1] = f;
            sim[j + 1][0] = p0; sim[j + 1][1] = p1; sim[j + 1][2] = p2;
        }
    };
    sort_simplex();
    int iterations = 1;
    while (fcalls < maxfun && iterations < maxiter) {
        double dx = 0.0, df = 0.0;
        for (int k = 1; k < 4; ++k) {
            for (int j = 0; j < 3; ++j) dx = fmax(dx, fabs(sim[k][j] - sim[0][j]));
            df = fmax(df, fabs(fsim[0] - fsim[k]));
        }
        if (dx <= 1e-4 && df <= 1e-4) break;

        double xbar[3], xt[3];
        for (int j = 0; j < 3; ++j) xbar[j] = ((sim[0][j] + sim[1][j]) + sim[2][j]) / N;
        // a function evaluation past maxfun raises in scipy and ends the iteration without a move
        auto eval = [&](const double *p, double &f) -> bool {
            if (fcalls >= maxfun) return false;
            f = gamma_penalized_nnlf(c, n, p[0], p[1], p[2], red);
            ++fcalls;
            return true;
        };
        bool ok = true;
        double fxr = 0.0;
        for (int j = 0; j < 3; ++j) xt[j] = 2.0 * xbar[j] - 1.0 * sim[3][j];
        ok = eval(xt, fxr);
        if (ok) {
            bool doshrink = false;
            if (fxr < fsim[0]) {
                double xe[3], fxe = 0.0;
                for (int j = 0; j < 3; ++j) xe[j] = 3.0 * xbar[j] - 2.0 * sim[3][j];
                ok = eval(xe, fxe);
                if (ok) {
                    if (fxe < fxr) { for (int j = 0; j < 3; ++j) sim[3][j] = xe[j]; fsim[3] = fxe; }
                    else { for (int j = 0; j < 3; ++j) sim[3][j] = xt[j]; fsim[3] = fxr; }
                }
            } else if (fxr < fsim[2]) {
                for (int j = 0; j < 3; ++j) sim[3][j] = xt[j];
                fsim[3] = fxr;
            } else if (fxr < fsim[3]) {                          // outside contraction
                double xc[3], fxc = 0.0;
                for (int j = 0; j < 3; ++j) xc[j] = 1.5 * xbar[j] - 0.5 * sim[3][j];
                ok = eval(xc, fxc);
                if (ok) {
                    if (fxc <= fxr) { for (int j = 0; j < 3; ++j) sim[3][j] = xc[j]; fsim[3] = fxc; }
                    else doshrink = true;
                }
            } else {                                             // inside contraction
                double xcc[3], fxcc = 0.0;
                for (int j = 0; j < 3; ++j) xcc[j] = 0.5 * xbar[j] + 0.5 * sim[3][j];
                ok = eval(xcc, fxcc);
                if (ok) {
                    if (fxcc < fsim[3]) { for (int j = 0; j < 3; ++j) sim[3][j] = xcc[j]; fsim[3] = fxcc; }
                    else doshrink = true;
                }
            }
            if (ok && doshrink) {
                for (int k = 1; k < 4 && ok; ++k) {
                    for (int j = 0; j < 3; ++j) sim[k][j] = sim[0][j] + 0.5 * (sim[k][j] - sim[0][j]);
                    ok = eval(sim[k], fsim[k]);
                }
            }
            if (ok) ++iterations;
        }
        sort_simplex();
    }
    if (evals) *evals = fcalls;
    const double a = sim[0][0], loc = sim[0][1], scale = sim[0][2];
    const double obj = gamma_penalized_nnlf(c, n, a, loc, scale, red);
    if (!(a > 0.0 && scale > 0.0) || !gf_finite(obj)) {          // scipy raises FitError here
        out[0] = out[1] = out[2] = NAN;
        return 3;
    }
    out[0] = a;
    out[1] = loc;
    out[2] = scale;
    return 0;
}

}  // namespace spade
