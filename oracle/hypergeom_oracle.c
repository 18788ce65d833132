/*
 * Plain-C restatement of scipy.stats.hypergeom.logcdf / logsf (scipy 1.18.1).
 *
 * TEST INFRASTRUCTURE ONLY.  Never linked into or called from the product
 * (pyspade_b200/); it is the bulk checker behind tests/ and bench.py's CPU leg.
 *
 * Algorithm restated (see oracle/hypergeom_port.py for the line references):
 *   support clamps        scipy/stats/_distn_infrastructure.py:3625-3663, 3703-3741
 *   branch rule + sums    scipy/stats/_discrete_distns.py:708-731
 *   log pmf               scipy/stats/_discrete_distns.py:669-675 (betaln form;
 *                         here a table of lgamma(i+1), the same quantity)
 * Call sites in the reference: pySpade/utils.py:212-213 and :263-264, including
 * the all([k, M, n, N]) guard that turns both tails into nan.
 *
 * The direct tail is a two-pass logsumexp over every pmf term of that tail, as
 * scipy does (no truncation), so that this file stays an independent statement
 * of the answer rather than a copy of the CUDA kernel's recurrence.
 *
 * Build:  gcc -O2 -fPIC -shared -fopenmp -o oracle/_build/libspade_oracle.so \
 *             oracle/hypergeom_oracle.c -lm        (see oracle/build.py)
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

/* ln(i!) for i = 0..lf_len-1, each entry an independent lgamma(i+1) call (no
 * running sum), rebuilt when a larger population size M arrives. */
static double *lf_tab = NULL;
static int64_t lf_len = 0;

static void ensure_table(int64_t M)
{
    if (M + 2 <= lf_len) return;
    int64_t len = M + 2;
    double *t = (double *)malloc((size_t)len * sizeof(double));
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < len; ++i) {
        int sg;
        t[i] = lgamma_r((double)i + 1.0, &sg);
    }
    free(lf_tab);
    lf_tab = t;
    lf_len = len;
}

static double lnchoose(int64_t a, int64_t b)
{
    if (b < 0 || b > a) return -INFINITY;
    return lf_tab[a] - lf_tab[b] - lf_tab[a - b];
}

static double logpmf(int64_t k, int64_t M, int64_t n, int64_t N)
{
    return lnchoose(n, k) + lnchoose(M - n, N - k) - lnchoose(M, N);
}

/* logsumexp of logpmf(j) for j in [lo, hi] */
static double tail_logsumexp(int64_t lo, int64_t hi, int64_t M, int64_t n, int64_t N)
{
    if (hi < lo) return -INFINITY;
    double mx = -INFINITY;
    for (int64_t j = lo; j <= hi; ++j) {
        double v = logpmf(j, M, n, N);
        if (v > mx) mx = v;
    }
    if (!isfinite(mx)) return mx;
    double acc = 0.0;
    for (int64_t j = lo; j <= hi; ++j) acc += exp(logpmf(j, M, n, N) - mx);
    return mx + log(acc);
}

static double oracle_logsf(int64_t k, int64_t M, int64_t n, int64_t N);

static double oracle_logcdf(int64_t k, int64_t M, int64_t n, int64_t N)
{
    if (!(M > 0 && n >= 0 && N >= 0 && n <= M && N <= M)) return NAN;
    int64_t a = N - (M - n); if (a < 0) a = 0;
    int64_t b = n < N ? n : N;
    if (k >= b) return 0.0;
    if (k < a) return -INFINITY;
    if (((double)k + 0.5) * ((double)M + 0.5) > ((double)n - 0.5) * ((double)N - 0.5))
        return log1p(-exp(oracle_logsf(k, M, n, N)));
    return tail_logsumexp(0, k, M, n, N);
}

static double oracle_logsf(int64_t k, int64_t M, int64_t n, int64_t N)
{
    if (!(M > 0 && n >= 0 && N >= 0 && n <= M && N <= M)) return NAN;
    int64_t a = N - (M - n); if (a < 0) a = 0;
    int64_t b = n < N ? n : N;
    if (k < a) return 0.0;
    if (k >= b) return -INFINITY;
    if (((double)k + 0.5) * ((double)M + 0.5) < ((double)n - 0.5) * ((double)N - 0.5))
        return log1p(-exp(oracle_logcdf(k, M, n, N)));
    return tail_logsumexp(k + 1, N, M, n, N);
}

/* Vector entry: plain scipy semantics (no reference guard). */
void spade_oracle_hypergeom(const int64_t *k, int64_t M, const int64_t *n, const int64_t *N,
                            int64_t count, double *logcdf_out, double *logsf_out)
{
    ensure_table(M > 0 ? M : 0);
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t i = 0; i < count; ++i) {
        logcdf_out[i] = oracle_logcdf(k[i], M, n[i], N[i]);
        logsf_out[i] = oracle_logsf(k[i], M, n[i], N[i]);
    }
}

/* Vector entry with the reference's all([k, M, n, N]) guard (utils.py:212-213):
 * up = logcdf, down = logsf, both nan when any argument is zero. */
void spade_oracle_pvalues(const int64_t *k, int64_t M, const int64_t *n, const int64_t *N,
                          int64_t count, double *up_out, double *down_out)
{
    ensure_table(M > 0 ? M : 0);
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t i = 0; i < count; ++i) {
        if (k[i] == 0 || M == 0 || n[i] == 0 || N[i] == 0) {
            up_out[i] = NAN;
            down_out[i] = NAN;
        } else {
            up_out[i] = oracle_logcdf(k[i], M, n[i], N[i]);
            down_out[i] = oracle_logsf(k[i], M, n[i], N[i]);
        }
    }
}
