// Row f4 of the scope table: the significance scoring that `pySpade global` / `pySpade local`
// run on every observed region against the DErand background (pySpade/global.py:107-187,
// local.py:104-177) -
//   p        = the region's raw log p-value with 0 -> 1, +-inf -> 0, nan -> 0 (global.py:118-123)
//   score    = scipy.stats.gamma.logsf(-p, A_g, B_g, C_g)                   (global.py:151-171)
//   emp      = #{draws s : rand[s, g] < p} / S                               (global.py:181,187)
// for ALL genes of a batch of regions at once (the reference evaluates the same expressions
// for the genes it keeps; picking those out is index bookkeeping on the host).
// scipy's gamma.logsf is rv_continuous.logsf (support / argument handling) over the generic
// _logsf: log(gammaincc(a, y)) above the median, log1p(-gammainc(a, y)) below, y = (x - loc) /
// scale.  The incomplete gamma functions (cephes igam / igamc) are restated here as the
// textbook pair: power series of P for y < a + 1, modified-Lentz continued fraction of Q
// otherwise, both with the prefactor exp(a ln y - y - lgamma(a)).
#pragma once
#include "common.cuh"

namespace spade {

// (P(a, x), Q(a, x)) for a > 0, 0 < x < inf: the side computed directly keeps its relative accuracy.
__device__ inline void incomplete_gamma_pair(double a, double x, double &P, double &Q)
{
    const double lfac = a * log(x) - x - lgamma(a);
    if (x < a + 1.0) {
        // P(a, x) = x^a e^-x / Gamma(a + 1) * sum_n x^n / ((a+1)...(a+n))
        double term = 1.0 / a, sum = term, ap = a;
        for (int n = 0; n < 100000; ++n) {
            ap += 1.0;
            term *= x / ap;
            sum += term;
            if (term < sum * 1e-17) break;
        }
        P = sum * exp(lfac);
        Q = 1.0 - P;
        return;
    }
    if (lfac < -7.09782712893383996843e2) {   // cephes igam_fac: the prefactor underflows, Q = 0
        P = 1.0;
        Q = 0.0;
        return;
    }
    // Q(a, x) = x^a e^-x / Gamma(a) * 1 / (x + 1 - a - 1 (1 - a) / (x + 3 - a - ...)), modified Lentz
    const double tiny = 1e-300;
    double b = x + 1.0 - a, c = 1.0 / tiny, d = 1.0 / b, h = d;
    for (int i = 1; i < 100000; ++i) {
        const double an = -(double)i * ((double)i - a);
        b += 2.0;
        d = an * d + b;
        if (fabs(d) < tiny) d = tiny;
        c = b + an / c;
        if (fabs(c) < tiny) c = tiny;
        d = 1.0 / d;
        const double del = d * c;
        h *= del;
        if (fabs(del - 1.0) < 1e-16) break;
    }
    Q = exp(lfac) * h;
    P = 1.0 - Q;
}

// scipy.stats.gamma.logsf(x, a, loc, scale) (rv_continuous.logsf): nan for invalid shape /
// scale or nan input, 0 at and below the lower end of the support, -inf at +inf; inside the
// support scipy 1.18's generic _logsf - log(sf) above the median (cdf > 1/2), log1p(-cdf) below.
__device__ inline double gamma_logsf_device(double x, double a, double loc, double scale)
{
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    if (!(a > 0.0) || !(scale > 0.0) || isinf(a) || isnan(loc) || isinf(loc) || isinf(scale)) return nan;
    const double y = (x - loc) / scale;
    if (isnan(y)) return nan;
    if (y <= 0.0) return 0.0;
    if (isinf(y)) return -INFINITY;
    double P, Q;
    incomplete_gamma_pair(a, y, P, Q);
    return P > 0.5 ? log(Q) : log1p(-P);
}

__device__ __forceinline__ double score_preprocess(double p)     // global.py:118-123 / local.py:106-114
{
    if (p == 0.0) return 1.0;
    if (isinf(p) || isnan(p)) return 0.0;
    return p;
}

// Ascending copy of every gene's column of the background table, nan last: one CTA per gene,
// bitonic sort in shared memory (S <= kScoreMaxDraws).  sorted[g][S]; valid[g] = #non-nan.
constexpr int kScoreMaxDraws = 16384;

__global__ void __launch_bounds__(256)
k_sort_background(const double *__restrict__ rand, int S, int64_t ld, double *__restrict__ sorted,
                  int32_t *__restrict__ valid)
{
    extern __shared__ double col[];
    __shared__ int n_valid;
    const int64_t g = blockIdx.x;
    int m = 32;
    while (m < S) m <<= 1;
    if (threadIdx.x == 0) n_valid = 0;
    __syncthreads();
    int mine = 0;
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
        double v = INFINITY;
        if (i < S) {
            v = rand[(int64_t)i * ld + g];
            if (isnan(v)) v = INFINITY; else ++mine;     // nan < p is never true: park them at the end
        }
        col[i] = v;
    }
    atomicAdd(&n_valid, mine);
    __syncthreads();
    for (int k = 2; k <= m; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < m; i += blockDim.x) {
                const int l = i ^ j;
                if (l > i) {
                    const double a = col[i], b = col[l];
                    if ((a > b) == ((i & k) == 0)) { col[i] = b; col[l] = a; }
                }
            }
            __syncthreads();
        }
    for (int i = threadIdx.x; i < S; i += blockDim.x) sorted[g * (int64_t)S + i] = col[i];
    if (threadIdx.x == 0) valid[g] = n_valid;
}

struct ScoreArgs {
    const double *obs;        // [R][obs_ld] raw log p-values of the observed regions
    int64_t obs_ld;
    const double *sorted;     // [G][S] from k_sort_background (may be null: no empirical p)
    const int32_t *valid;
    const double *A, *B, *C;  // [G] gamma parameters (may be null: no score)
    int64_t R, G;
    int S;
    double *score, *emp;      // [R][out_ld]; either may be null
    int64_t out_ld;
};

__global__ void __launch_bounds__(256) k_score(const ScoreArgs a)
{
    const int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (g >= a.G) return;
    double A = 0, B = 0, C = 0;
    if (a.A) { A = a.A[g]; B = a.B[g]; C = a.C[g]; }
    const double *__restrict__ col = a.sorted ? a.sorted + g * (int64_t)a.S : nullptr;
    const int nv = a.sorted ? a.valid[g] : 0;
    for (int64_t r = blockIdx.y; r < a.R; r += gridDim.y) {
        const double p = score_preprocess(a.obs[r * a.obs_ld + g]);
        if (a.score && a.A) a.score[r * a.out_ld + g] = gamma_logsf_device(-p, A, B, C);
        if (a.emp && col) {
            int lo = 0, hi = nv;                         // first index with col[i] >= p
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (col[mid] < p) lo = mid + 1; else hi = mid;
            }
            a.emp[r * a.out_ld + g] = (double)lo / (double)a.S;
        }
    }
}

}  // namespace spade
