// K3: log-hypergeometric tails with scipy's definition (scipy/stats/_discrete_distns.py,
// hypergeom_gen._logcdf/_logsf, restated in oracle/hypergeom_port.py), and the per-run
// tail table used when every cell set of a batch has the same size N (DErand).
#pragma once
#include "common.cuh"

namespace spade {

// Terms below this fraction of the running sum end a tail series.  The terms decay at least
// geometrically there, so the sum is short by a few times this fraction: ~1e-11 relative, five
// orders below the 1e-6 contract on the log p-values (1e-15 cost 11 % more terms in the FP64-bound
// direct tails for nothing the contract or the comparison with scipy - 6e-9, set by lgamma - can see).
constexpr double kTailEps = 1e-12;

__device__ __forceinline__ double ln_choose(const double *__restrict__ lf, int a, int b)
{
    return lf[a] - lf[b] - lf[a - b];
}

// Sum of a tail relative to its edge term: 1 + r1 + r1 r2 + ... with r_i = (fa fb)/(fc fd),
// the four factors stepping by one per term (fa, fb down; fc, fd up), until a factor reaches
// zero (end of the support) or a term drops below kTailEps of the sum.  Division-free: with
// P_i = prod(fa fb), D_i = prod(fc fd) the sum is N_i / D_i, N_i = N_{i-1} fc fd + P_i; the
// three running values are rescaled by a power of two every 8 terms (factors are < 2^31, so
// eight steps grow them by less than 2^500).  The two factor products are stepped by their
// differences ((fa-1)(fb-1) = fa fb - (fa + fb - 1), (fc+1)(fd+1) = fc fd + (fc + fd + 1): exact
// integers in double up to 2^53, i.e. for every population below 9.4e7 cells) - 10 instead of
// 13 float64 operations per term of this FP64-bound loop.
__device__ __forceinline__ double tail_series(double fa, double fb, double fc, double fd)
{
    double P = 1.0, D = 1.0, Nn = 1.0;
    double pab = fa * fb, sab = fa + fb - 1.0, pcd = fc * fd, tcd = fc + fd + 1.0;
    bool more = fa > 0.0 && fb > 0.0;
    while (more) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (more) {
                P *= pab;
                D *= pcd;
                Nn = fma(Nn, pcd, P);
                pab -= sab; sab -= 2.0; pcd += tcd; tcd += 2.0;
                more = pab > 0.0 && !(P < Nn * kTailEps);
            }
        }
        // exact power-of-two rescale that brings D into [1, 2)
        const int e = (int)((__double2hiint(D) >> 20) & 0x7ff) - 1023;
        const double sc = __hiloint2double((1023 - e) << 20, 0);
        P *= sc;
        D *= sc;
        Nn *= sc;
    }
    return Nn / D;
}

// logcdf = ln P(X <= k), logsf = ln P(X > k), X ~ Hypergeom(M, n, N).
// The tail on the far side of the mean is summed directly as pmf(edge) * tail_series, anchored
// at a log-factorial-table log pmf; the near side is log1p(-exp(far side)) - the branch rule
// of scipy, so values next to 0 keep their precision.
__device__ void hyper_tails(int k, int M, int n, int N, const double *__restrict__ lf,
                            double &logcdf, double &logsf)
{
    const int a = max(N - (M - n), 0), b = min(n, N);
    if (k >= b) { logcdf = 0.0; logsf = -INFINITY; return; }
    if (k < a)  { logcdf = -INFINITY; logsf = 0.0; return; }
    const double ln_total = ln_choose(lf, M, N);
    const bool upper_small = ((double)k + 0.5) * ((double)M + 0.5) > ((double)n - 0.5) * ((double)N - 0.5);
    if (upper_small) {
        const int j = k + 1;
        const double lp = ln_choose(lf, n, j) + ln_choose(lf, M - n, N - j) - ln_total;
        const double sum = tail_series((double)(n - j), (double)(N - j), (double)(j + 1),
                                       (double)(M - n - N + j + 1));
        logsf = lp + log(sum);
        logcdf = log1p(-exp(logsf));
    } else {
        const int j = k;
        const double lp = ln_choose(lf, n, j) + ln_choose(lf, M - n, N - j) - ln_total;
        const double sum = tail_series((double)j, (double)(M - n - N + j), (double)(n - j + 1),
                                       (double)(N - j + 1));
        logcdf = lp + log(sum);
        logsf = log1p(-exp(logcdf));
    }
}

// Tail table: for gene g and k in [lo_g, lo_g + len_g) the pair (logcdf, logsf) of
// Hypergeom(M, n_g, N).  Entries are hyper_tails() itself, so a looked-up value is
// bit-identical to a directly evaluated one.  One warp per gene.
__global__ void __launch_bounds__(256)
k_tail_table(const TabMeta *__restrict__ tmeta, const GeneMeta *__restrict__ meta, int64_t G,
             int M, int N, const double *__restrict__ lf, double2 *__restrict__ tab)
{
    const int64_t g = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (g >= G) return;
    const TabMeta t = tmeta[g];
    const int n = meta[g].n;
    for (int i = lane; i < t.len; i += 32) {
        double lc, ls;
        hyper_tails(t.lo + i, M, n, N, lf, lc, ls);
        tab[t.off + i] = make_double2(lc, ls);
    }
}

// Direct tails of finished counts, for batches whose sets differ in size (DEobs: no tail table).
// The fused kernel leaves k[row][gene]; here hyper_tails runs at full occupancy (the fused kernel
// keeps 8 KB of shared memory per warp and 24 warps per SM, which starves this FP64-bound step).
// A warp takes ONE gene and 32 sets that are neighbours in the order by set size: its lanes then
// share n and have similar N and k, so their tail series have nearly the same length (with
// 32 neighbouring genes per warp the dense genes' long series would stall the sparse genes'
// short ones).  Rows [0, S), genes [g0, g1); sizes[row] = N of the row's set (or offsets[row + 1]
// - offsets[row] when sizes is null); order[i] = row of rank i by size (null: identity).
__global__ void __launch_bounds__(256)
k_tails_direct(const int32_t *__restrict__ k, int64_t k_ld, const GeneMeta *__restrict__ meta,
               const int64_t *__restrict__ sizes, const int64_t *__restrict__ offsets,
               const int32_t *__restrict__ order, int M, const double *__restrict__ lf,
               double *__restrict__ up, double *__restrict__ down, int64_t ld, int64_t S, int64_t g0,
               int64_t g1, int64_t col0)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t g = g0 + blockIdx.x * (int64_t)(blockDim.x >> 5) + warp;
    if (g >= g1) return;
    const int n = meta[g].n;
    const int64_t col = g - col0;                          // column of gene g in k / up / down
    for (int64_t i = blockIdx.y * 32 + lane; i < S; i += (int64_t)gridDim.y * 32) {
        const int64_t row = order ? order[i] : i;
        const long long N = sizes ? sizes[row] : offsets[row + 1] - offsets[row];
        const int kk = k[row * k_ld + col];
        double lc, ls;
        if (kk == 0 || M == 0 || n == 0 || N == 0) lc = ls = __longlong_as_double(0x7ff8000000000000ll);
        else hyper_tails(kk, M, n, (int)N, lf, lc, ls);
        if (up) up[row * ld + col] = lc;
        if (down) down[row * ld + col] = ls;
    }
}

__global__ void k_hyper_probe(const int64_t *__restrict__ k, const int64_t *__restrict__ M,
                              const int64_t *__restrict__ n, const int64_t *__restrict__ N,
                              int64_t count, const double *__restrict__ lf,
                              double *__restrict__ logcdf, double *__restrict__ logsf)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= count) return;
    const int64_t kk = k[i], MM = M[i], nn = n[i], NN = N[i];
    double lc, ls;
    if (!(MM > 0 && nn >= 0 && NN >= 0 && nn <= MM && NN <= MM)) {
        lc = ls = __longlong_as_double(0x7ff8000000000000ll);
    } else {
        // clamp k into int range around the support; the clamps in hyper_tails decide
        const int64_t kc = kk < -1 ? -1 : (kk > MM ? MM : kk);
        hyper_tails((int)kc, (int)MM, (int)nn, (int)NN, lf, lc, ls);
    }
    logcdf[i] = lc;
    logsf[i] = ls;
}

}  // namespace spade
