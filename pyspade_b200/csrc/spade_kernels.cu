// libspade_b200: sm_100a kernels + C ABI for pySpade's hypergeometric DE hot path.
// See include/spade.h for the contract and the reference lines each entry replaces.
//
// Data layout in HBM (per engine, one engine per GPU)
//   CSR copy of the caller's matrix      indptr int64[G+1], indices int32[nnz], values f64[nnz]
//   cell-major (CSC) entries             colptr int64[C+1], ent uint2[nnz]:
//                                          .x = gene | above_cutoff << 31, .y = fp32 bits of the value
//   per gene                             median f64, n int32, negmode u8, rowsum int64 (fixed point),
//                                        bgmean f64 (NC background only)
//   log-factorial table                  lf f64[C+1], lf[i] = ln(i!)
//   per batch of sets                    acc_sum int64[S*G], acc_cnt int32[S*G]
//
// Sums are carried as signed 64-bit fixed point (value * 2^fix_shift): integer
// adds commute, so results are bit-identical for any launch geometry, any order
// of atomics and any sharding of sets over GPUs.
#include <cuda_runtime.h>
#include <cub/cub.cuh>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "spade.h"

namespace {

thread_local std::string g_create_error = "";

constexpr int kStageThreads = 256;
constexpr int kAccWarpsPerBlock = 8;
constexpr int kFinThreads = 256;
constexpr int64_t kMaxTestsPerChunk = int64_t(1) << 24;      // device outputs: 16 Mi tests per chunk
constexpr int64_t kMaxTestsPerChunkHost = int64_t(1) << 22;  // host outputs: 4 Mi, copy/compute overlap
constexpr double kTailEps = 1e-18;

#define CUDA_TRY(expr)                                                                         \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));                     \
            return SPADE_ERR_CUDA;                                                             \
        }                                                                                      \
    } while (0)

// ------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long order_key(double v)
{
    if (v == 0.0) v = 0.0;   // -0.0 and +0.0 share a key
    unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

__device__ __forceinline__ double key_value(unsigned long long k)
{
    unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}

__device__ __forceinline__ long long to_fixed(float v, float scale)
{
    return __float2ll_rn(v * scale);
}

template <typename T>
__device__ __forceinline__ T warp_sum(T v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

// Sum over the block; result valid in every thread.  `scratch` holds >= 33 elements.
template <typename T>
__device__ T block_sum(T v, T *scratch)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    if (warp == 0) {
        T t = lane < nwarp ? scratch[lane] : T(0);
        t = warp_sum(t);
        if (lane == 0) scratch[32] = t;
    }
    __syncthreads();
    return scratch[32];
}

// ------------------------------------------------------------------------------------------
// K0: CPM normalisation (utils.py:150-155).  value = fl(fl(x * 1e6) * fl(1 / colsum)).
// ------------------------------------------------------------------------------------------
__global__ void k_col_sums(const int32_t *__restrict__ indices, const int32_t *__restrict__ counts,
                           int64_t nnz, int64_t C, unsigned long long *__restrict__ colsum,
                           int *__restrict__ err)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nnz;
         i += (int64_t)gridDim.x * blockDim.x) {
        int32_t c = indices[i];
        if (c < 0 || c >= C) { atomicExch(err, 1); continue; }
        atomicAdd(&colsum[c], (unsigned long long)(long long)counts[i]);
    }
}

__global__ void k_cpm(const int32_t *__restrict__ indices, const int32_t *__restrict__ counts,
                      int64_t nnz, const unsigned long long *__restrict__ colsum,
                      double *__restrict__ values)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nnz;
         i += (int64_t)gridDim.x * blockDim.x) {
        double inv = __drcp_rn((double)(long long)colsum[indices[i]]);    // fl(1/colsum)
        values[i] = __dmul_rn(__dmul_rn((double)counts[i], 1e6), inv);    // no FMA contraction
    }
}

// ------------------------------------------------------------------------------------------
// CSR -> CSC plumbing
// ------------------------------------------------------------------------------------------
__global__ void k_iota(uint32_t *p, int64_t n)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x)
        p[i] = (uint32_t)i;
}

__global__ void k_col_hist(const int32_t *__restrict__ indices, int64_t nnz, int64_t C,
                           unsigned long long *__restrict__ hist, int *__restrict__ err)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nnz;
         i += (int64_t)gridDim.x * blockDim.x) {
        int32_t c = indices[i];
        if (c < 0 || c >= C) { atomicExch(err, 1); continue; }
        atomicAdd(&hist[c], 1ull);
    }
}

__global__ void k_invert_perm(const uint32_t *__restrict__ sorted_pos, int64_t nnz,
                              uint32_t *__restrict__ cscpos)
{
    for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < nnz;
         j += (int64_t)gridDim.x * blockDim.x)
        cscpos[sorted_pos[j]] = (uint32_t)j;
}

__global__ void k_row_abs(const int64_t *__restrict__ indptr, const double *__restrict__ values,
                          double *__restrict__ rowabs)
{
    __shared__ double scratch[33];
    const int64_t g = blockIdx.x;
    double acc = 0.0;
    for (int64_t i = indptr[g] + threadIdx.x; i < indptr[g + 1]; i += blockDim.x)
        acc += fabs((double)(float)values[i]);
    acc = block_sum(acc, scratch);
    if (threadIdx.x == 0) rowabs[g] = acc;
}

// ------------------------------------------------------------------------------------------
// K1: per-gene cutoff (utils.py:199-200, NC variant :253-254), flags, row sums.
// One CTA per gene.  The cutoff is np.median over the scope (all C cells, or the NC
// cells), zeros included: order statistics of the multiset {stored values in scope}
// + {implicit zeros}, found with an 8-bit-per-pass radix select on order-preserving
// keys of the float64 values.
// ------------------------------------------------------------------------------------------
struct StageArgs {
    const int64_t *indptr;
    const int32_t *indices;
    const double *values;
    const uint32_t *cscpos;
    const uint8_t *ncmask;    // null -> scope = all cells
    int64_t C;
    int64_t scope_len;        // C or |NC|
    float fix_scale;
    double fix_inv;           // 2^-fix_shift
    uint2 *ent;
    double *median;
    int32_t *n;
    uint8_t *negmode;
    long long *rowsum;
    double *bgmean;
};

__device__ double block_select(const StageArgs &a, int64_t beg, int64_t end, int64_t implicit_zeros,
                               int64_t rank, unsigned long long *hist /*256*/,
                               unsigned long long *bcast /*2*/)
{
    const unsigned long long zkey = order_key(0.0);
    unsigned long long prefix = 0;
    for (int pass = 7; pass >= 0; --pass) {
        const int shift = pass * 8;
        for (int b = threadIdx.x; b < 256; b += blockDim.x) hist[b] = 0;
        __syncthreads();
        for (int64_t i = beg + threadIdx.x; i < end; i += blockDim.x) {
            if (a.ncmask && !a.ncmask[a.indices[i]]) continue;
            unsigned long long key = order_key(a.values[i]);
            bool match = (pass == 7) || ((key >> (shift + 8)) == (prefix >> (shift + 8)));
            if (match) atomicAdd(&hist[(key >> shift) & 255], 1ull);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            bool zmatch = (pass == 7) || ((zkey >> (shift + 8)) == (prefix >> (shift + 8)));
            if (zmatch) hist[(zkey >> shift) & 255] += (unsigned long long)implicit_zeros;
            unsigned long long r = (unsigned long long)rank, cum = 0;
            int bin = 255;
            for (int b = 0; b < 256; ++b) {
                if (cum + hist[b] > r) { bin = b; break; }
                cum += hist[b];
            }
            bcast[0] = prefix | ((unsigned long long)bin << shift);
            bcast[1] = r - cum;
        }
        __syncthreads();
        prefix = bcast[0];
        rank = (int64_t)bcast[1];
        __syncthreads();
    }
    return key_value(prefix);
}

__global__ void __launch_bounds__(kStageThreads) k_gene_stage(StageArgs a)
{
    __shared__ unsigned long long hist[256];
    __shared__ unsigned long long bcast[2];
    __shared__ long long scratch_ll[33];

    const int64_t g = blockIdx.x;
    const int64_t beg = a.indptr[g], end = a.indptr[g + 1];

    // pass 0: census of the scope
    long long in_scope = 0, neg = 0, pos = 0;
    for (int64_t i = beg + threadIdx.x; i < end; i += blockDim.x) {
        if (a.ncmask && !a.ncmask[a.indices[i]]) continue;
        double v = a.values[i];
        ++in_scope;
        neg += v < 0.0;
        pos += v > 0.0;
    }
    in_scope = block_sum(in_scope, scratch_ll);
    neg = block_sum(neg, scratch_ll);
    pos = block_sum(pos, scratch_ll);
    const int64_t L = a.scope_len;
    const int64_t implicit_zeros = L - in_scope;
    const int64_t zeros = L - neg - pos;            // implicit + explicitly stored zeros
    const int64_t lo_i = (L - 1) / 2, hi_i = L / 2;

    double m = 0.0;
    if (L > 0) {
        const bool lo_zero = lo_i >= neg && lo_i < neg + zeros;
        const bool hi_zero = hi_i >= neg && hi_i < neg + zeros;
        double va = lo_zero ? 0.0 : block_select(a, beg, end, implicit_zeros, lo_i, hist, bcast);
        double vb = (hi_i == lo_i) ? va
                    : (hi_zero ? 0.0 : block_select(a, beg, end, implicit_zeros, hi_i, hist, bcast));
        m = (hi_i == lo_i) ? va : __dmul_rn(__dadd_rn(va, vb), 0.5);   // np.median: fl(a+b)/2
    }
    const bool negmode = m < 0.0;

    long long n_low = 0, rowsum = 0, bgsum = 0;
    for (int64_t i = beg + threadIdx.x; i < end; i += blockDim.x) {
        const double v = a.values[i];
        const int32_t c = a.indices[i];
        const float v32 = (float)v;
        const long long fx = to_fixed(v32, a.fix_scale);
        const bool low = v <= m;
        n_low += low;
        rowsum += fx;
        if (a.ncmask && a.ncmask[c]) bgsum += fx;
        const uint32_t flag = negmode ? (low ? 1u : 0u) : (low ? 0u : 1u);
        a.ent[a.cscpos[i]] = make_uint2((uint32_t)g | (flag << 31), __float_as_uint(v32));
    }
    n_low = block_sum(n_low, scratch_ll);
    rowsum = block_sum(rowsum, scratch_ll);
    bgsum = block_sum(bgsum, scratch_ll);
    if (threadIdx.x == 0) {
        a.median[g] = m;
        a.n[g] = (int32_t)(n_low + ((0.0 <= m) ? (a.C - (end - beg)) : 0));
        a.negmode[g] = negmode ? 1 : 0;
        a.rowsum[g] = rowsum;
        a.bgmean[g] = a.ncmask ? ((double)bgsum * a.fix_inv) * (1.0 / (double)L) : 0.0;
    }
}

// ------------------------------------------------------------------------------------------
// K2 + K4 (v1): one warp per (set, member cell) walks the cell's CSC column and adds
// (fixed-point value, above-cutoff flag) into the set's per-gene accumulators.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kAccWarpsPerBlock * 32)
k_accumulate(const int64_t *__restrict__ colptr, const uint2 *__restrict__ ent,
             const int32_t *__restrict__ members, const int64_t *__restrict__ offsets,
             int64_t s0, int64_t s1, int64_t G, int64_t C, float fix_scale,
             unsigned long long *__restrict__ acc_sum, int *__restrict__ acc_cnt,
             int *__restrict__ err)
{
    const int lane = threadIdx.x & 31;
    const int64_t p0 = offsets[s0], p1 = offsets[s1];
    const int64_t p = p0 + (int64_t)blockIdx.x * kAccWarpsPerBlock + (threadIdx.x >> 5);
    if (p >= p1) return;
    // set of this member: last s in [s0, s1) with offsets[s] <= p
    int64_t lo = s0, hi = s1 - 1;
    while (lo < hi) {
        int64_t mid = (lo + hi + 1) >> 1;
        if (offsets[mid] <= p) lo = mid; else hi = mid - 1;
    }
    const int64_t base = (lo - s0) * G;
    const int32_t c = members[p];
    if (c < 0 || c >= C) { if (lane == 0) atomicExch(err, 2); return; }
    const int64_t beg = colptr[c], end = colptr[c + 1];
    for (int64_t i = beg + lane; i < end; i += 32) {
        const uint2 e = ent[i];
        const uint32_t gene = e.x & 0x7fffffffu;
        const long long fx = to_fixed(__uint_as_float(e.y), fix_scale);
        atomicAdd(&acc_sum[base + gene], (unsigned long long)fx);
        if (e.x >> 31) atomicAdd(&acc_cnt[base + gene], 1);
    }
}

// ------------------------------------------------------------------------------------------
// K3: log-hypergeometric tails, scipy's definition (see oracle/hypergeom_port.py):
// the tail on the far side of the mean is summed directly as pmf(edge) * (1 + r + r r' + ...),
// with the ratio recurrence in float64 anchored at a log-factorial-table log pmf; the near
// side is log1p(-exp(far side)).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double ln_choose(const double *__restrict__ lf, int a, int b)
{
    return lf[a] - lf[b] - lf[a - b];
}

__device__ void hyper_tails(int k, int M, int n, int N, const double *__restrict__ lf,
                            double &logcdf, double &logsf)
{
    const int a = max(N - (M - n), 0), b = min(n, N);
    if (k >= b) { logcdf = 0.0; logsf = -INFINITY; return; }
    if (k < a)  { logcdf = -INFINITY; logsf = 0.0; return; }
    const double ln_total = ln_choose(lf, M, N);
    const bool upper_small = ((double)k + 0.5) * ((double)M + 0.5) > ((double)n - 0.5) * ((double)N - 0.5);
    double t = 1.0, sum = 1.0;
    if (upper_small) {
        const int j = k + 1;
        const double lp = ln_choose(lf, n, j) + ln_choose(lf, M - n, N - j) - ln_total;
        double fa = (double)(n - j), fb = (double)(N - j), fc = (double)(j + 1),
               fd = (double)(M - n - N + j + 1);
        while (fa > 0.0 && fb > 0.0) {
            t *= (fa * fb) / (fc * fd);
            sum += t;
            if (t < sum * kTailEps) break;
            fa -= 1.0; fb -= 1.0; fc += 1.0; fd += 1.0;
        }
        logsf = lp + log(sum);
        logcdf = log1p(-exp(logsf));
    } else {
        const int j = k;
        const double lp = ln_choose(lf, n, j) + ln_choose(lf, M - n, N - j) - ln_total;
        double fa = (double)j, fb = (double)(M - n - N + j), fc = (double)(n - j + 1),
               fd = (double)(N - j + 1);
        while (fa > 0.0 && fb > 0.0) {
            t *= (fa * fb) / (fc * fd);
            sum += t;
            if (t < sum * kTailEps) break;
            fa -= 1.0; fb -= 1.0; fc += 1.0; fd += 1.0;
        }
        logcdf = lp + log(sum);
        logsf = log1p(-exp(logcdf));
    }
}

struct FinalArgs {
    const unsigned long long *acc_sum;
    const int *acc_cnt;
    const int64_t *offsets;
    int64_t s0, s1, G, C;
    const int32_t *n;
    const uint8_t *negmode;
    const long long *rowsum;
    const double *bgmean;
    const double *lf;
    int nc_mode;
    double fix_inv;
    double *up, *down, *fc, *cpm;   // already offset to set s0; may be null
    int32_t *k;
};

__global__ void __launch_bounds__(kFinThreads) k_finalize(FinalArgs a)
{
    const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t total = (a.s1 - a.s0) * a.G;
    if (idx >= total) return;
    const int64_t sl = idx / a.G, g = idx - sl * a.G;
    const int64_t s = a.s0 + sl;
    const long long N = (long long)(a.offsets[s + 1] - a.offsets[s]);
    const long long sum = (long long)a.acc_sum[idx];
    const int cnt = a.acc_cnt[idx];
    const int kk = a.negmode[g] ? cnt : (int)(N - cnt);
    const int n = a.n[g];
    const int M = (int)a.C;

    const double mean_set = ((double)sum * a.fix_inv) * (1.0 / (double)N);
    double mean_bg;
    if (a.nc_mode) {
        mean_bg = a.bgmean[g];
    } else {
        const long long rest = a.rowsum[g] - sum;      // exact: same fixed-point values
        mean_bg = ((double)rest * a.fix_inv) * (1.0 / (double)((long long)a.C - N));
    }
    if (a.cpm) a.cpm[idx] = mean_set;
    if (a.fc) a.fc[idx] = (mean_set + 0.01) / (mean_bg + 0.01);
    if (a.k) a.k[idx] = kk;
    if (a.up || a.down) {
        double lc, ls;
        if (kk == 0 || M == 0 || n == 0 || N == 0) {       // all([k, M, n, N]) guard, utils.py:212-213
            lc = ls = __longlong_as_double(0x7ff8000000000000ll);
        } else {
            hyper_tails(kk, M, n, (int)N, a.lf, lc, ls);
        }
        if (a.up) a.up[idx] = lc;
        if (a.down) a.down[idx] = ls;
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

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t count)
    {
        if (count <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMalloc((void **)&p, std::max<size_t>(count, 1) * sizeof(T));
        if (e == cudaSuccess) cap = count;
        return e;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

inline int grid_for(int64_t n, int threads, int cap = 148 * 32)
{
    int64_t b = (n + threads - 1) / threads;
    return (int)std::max<int64_t>(1, std::min<int64_t>(b, cap));
}

}  // namespace

struct spade_engine {
    int device = 0;
    int64_t G = 0, C = 0, nnz = 0;
    int fix_shift = 0;
    float fix_scale = 1.f;
    double fix_inv = 1.0;
    int64_t n_nc = 0;

    DevBuf<int64_t> indptr, colptr;
    DevBuf<int32_t> indices;
    DevBuf<double> values;
    DevBuf<uint32_t> cscpos;
    DevBuf<uint2> ent;
    DevBuf<double> median, bgmean, lf;
    DevBuf<int32_t> n;
    DevBuf<uint8_t> negmode, ncmask;
    DevBuf<long long> rowsum;
    DevBuf<int> err;

    // run workspace
    DevBuf<unsigned long long> acc_sum;
    DevBuf<int> acc_cnt;
    DevBuf<int32_t> members;
    DevBuf<int64_t> offsets;
    DevBuf<double> out_f64[2];  // host-output path: 4 tables of one chunk, double-buffered
    DevBuf<int32_t> out_k[2];
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t buf_ready[2] = {nullptr, nullptr}, buf_free[2] = {nullptr, nullptr};

    std::vector<cudaEvent_t> ev_pool;
    size_t ev_used = 0;
    double acc_ms = 0.0, fin_ms = 0.0;
    int64_t launches = 0;

    std::string error = "";

    void set_error(const std::string &m) { error = m; }

    cudaEvent_t next_event()
    {
        if (ev_used == ev_pool.size()) {
            cudaEvent_t e;
            if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
            ev_pool.push_back(e);
        }
        return ev_pool[ev_used++];
    }

    int collect_timing()
    {
        // events are recorded in triples (start, after accumulate, after finalize)
        for (size_t i = 0; i + 3 <= ev_used; i += 3) {
            float a = 0.f, f = 0.f;
            CUDA_TRY(cudaEventSynchronize(ev_pool[i + 2]));
            CUDA_TRY(cudaEventElapsedTime(&a, ev_pool[i], ev_pool[i + 1]));
            CUDA_TRY(cudaEventElapsedTime(&f, ev_pool[i + 1], ev_pool[i + 2]));
            acc_ms += a;
            fin_ms += f;
        }
        ev_used = 0;
        return SPADE_OK;
    }

    int check_device_error(const char *what)
    {
        int h = 0;
        CUDA_TRY(cudaMemcpy(&h, err.p, sizeof(int), cudaMemcpyDeviceToHost));
        if (h != 0) {
            CUDA_TRY(cudaMemset(err.p, 0, sizeof(int)));
            set_error(std::string(what) + (h == 2 ? ": member cell index out of range"
                                                  : ": cell index out of range in the matrix"));
            return SPADE_ERR_ARG;
        }
        return SPADE_OK;
    }

    int stage_genes(cudaStream_t st)
    {
        StageArgs a;
        a.indptr = indptr.p;
        a.indices = indices.p;
        a.values = values.p;
        a.cscpos = cscpos.p;
        a.ncmask = n_nc > 0 ? ncmask.p : nullptr;
        a.C = C;
        a.scope_len = n_nc > 0 ? n_nc : C;
        a.fix_scale = fix_scale;
        a.fix_inv = fix_inv;
        a.ent = ent.p;
        a.median = median.p;
        a.n = n.p;
        a.negmode = negmode.p;
        a.rowsum = rowsum.p;
        a.bgmean = bgmean.p;
        if (G > 0) k_gene_stage<<<(unsigned)G, kStageThreads, 0, st>>>(a);
        CUDA_TRY(cudaGetLastError());
        return SPADE_OK;
    }

    // Everything after the CSR (indptr/indices/values) is resident on the device.
    int build(cudaStream_t st)
    {
        CUDA_TRY(err.reserve(1));
        CUDA_TRY(cudaMemsetAsync(err.p, 0, sizeof(int), st));
        CUDA_TRY(colptr.reserve(C + 1));
        CUDA_TRY(cscpos.reserve(nnz));
        CUDA_TRY(ent.reserve(nnz));
        CUDA_TRY(median.reserve(G));
        CUDA_TRY(bgmean.reserve(G));
        CUDA_TRY(n.reserve(G));
        CUDA_TRY(negmode.reserve(G));
        CUDA_TRY(rowsum.reserve(G));
        CUDA_TRY(ncmask.reserve(C));

        // log-factorial table, host lgamma (<= 1 ulp), one entry per population size
        {
            std::vector<double> h(C + 1);
            for (int64_t i = 0; i <= C; ++i) {
                int sg;
                h[i] = lgamma_r((double)i + 1.0, &sg);
            }
            CUDA_TRY(lf.reserve(C + 1));
            CUDA_TRY(cudaMemcpyAsync(lf.p, h.data(), (C + 1) * sizeof(double), cudaMemcpyHostToDevice, st));
            CUDA_TRY(cudaStreamSynchronize(st));
        }

        // fixed-point shift from the largest row |sum| (of the fp32-rounded values)
        {
            DevBuf<double> rowabs;
            CUDA_TRY(rowabs.reserve(G));
            if (G > 0) k_row_abs<<<(unsigned)G, kStageThreads, 0, st>>>(indptr.p, values.p, rowabs.p);
            CUDA_TRY(cudaGetLastError());
            std::vector<double> h(G);
            CUDA_TRY(cudaMemcpyAsync(h.data(), rowabs.p, G * sizeof(double), cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaStreamSynchronize(st));
            rowabs.release();
            double mx = 0.0;
            for (double v : h) {
                if (!(v == v) || std::isinf(v)) {
                    set_error("matrix holds non-finite values");
                    return SPADE_ERR_ARG;
                }
                mx = std::max(mx, v);
            }
            int need = 1;
            while (need < 62 && std::ldexp(1.0, need) <= mx + 1.0) ++need;
            fix_shift = std::max(0, std::min(30, 61 - need));
            fix_scale = std::ldexp(1.0f, fix_shift);
            fix_inv = std::ldexp(1.0, -fix_shift);
        }

        // CSR -> CSC: stable radix sort of entry positions by cell index
        if (nnz > 0) {
            DevBuf<uint32_t> pos_in, pos_out;
            DevBuf<int32_t> keys_out;
            DevBuf<unsigned long long> hist;
            CUDA_TRY(pos_in.reserve(nnz));
            CUDA_TRY(pos_out.reserve(nnz));
            CUDA_TRY(keys_out.reserve(nnz));
            CUDA_TRY(hist.reserve(C + 1));
            k_iota<<<grid_for(nnz, 256), 256, 0, st>>>(pos_in.p, nnz);
            CUDA_TRY(cudaGetLastError());
            int end_bit = 1;
            while ((int64_t(1) << end_bit) < C && end_bit < 31) ++end_bit;
            size_t temp_bytes = 0;
            CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, indices.p, keys_out.p, pos_in.p,
                                                     pos_out.p, nnz, 0, end_bit, st));
            DevBuf<uint8_t> temp;
            CUDA_TRY(temp.reserve(temp_bytes));
            CUDA_TRY(cub::DeviceRadixSort::SortPairs(temp.p, temp_bytes, indices.p, keys_out.p, pos_in.p,
                                                     pos_out.p, nnz, 0, end_bit, st));
            k_invert_perm<<<grid_for(nnz, 256), 256, 0, st>>>(pos_out.p, nnz, cscpos.p);
            CUDA_TRY(cudaGetLastError());
            CUDA_TRY(cudaMemsetAsync(hist.p, 0, (C + 1) * sizeof(unsigned long long), st));
            k_col_hist<<<grid_for(nnz, 256), 256, 0, st>>>(indices.p, nnz, C, hist.p, err.p);
            CUDA_TRY(cudaGetLastError());
            size_t scan_bytes = 0;
            CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (long long *)hist.p,
                                                   (long long *)colptr.p, C + 1, st));
            CUDA_TRY(temp.reserve(scan_bytes));
            CUDA_TRY(cub::DeviceScan::ExclusiveSum(temp.p, scan_bytes, (long long *)hist.p,
                                                   (long long *)colptr.p, C + 1, st));
            CUDA_TRY(cudaStreamSynchronize(st));
            pos_in.release(); pos_out.release(); keys_out.release(); hist.release(); temp.release();
            int rc = check_device_error("spade_create");
            if (rc != SPADE_OK) return rc;
        } else {
            CUDA_TRY(cudaMemsetAsync(colptr.p, 0, (C + 1) * sizeof(int64_t), st));
        }
        n_nc = 0;
        int rc = stage_genes(st);
        if (rc != SPADE_OK) return rc;
        CUDA_TRY(cudaStreamSynchronize(st));
        return SPADE_OK;
    }

    void free_all()
    {
        indptr.release(); colptr.release(); indices.release(); values.release(); cscpos.release();
        ent.release(); median.release(); bgmean.release(); lf.release(); n.release();
        negmode.release(); ncmask.release(); rowsum.release(); err.release(); acc_sum.release();
        acc_cnt.release(); members.release(); offsets.release(); 
        for (int b = 0; b < 2; ++b) {
            out_f64[b].release();
            out_k[b].release();
            if (buf_ready[b]) cudaEventDestroy(buf_ready[b]);
            if (buf_free[b]) cudaEventDestroy(buf_free[b]);
            buf_ready[b] = buf_free[b] = nullptr;
        }
        if (copy_stream) cudaStreamDestroy(copy_stream);
        copy_stream = nullptr;
        for (cudaEvent_t e : ev_pool) cudaEventDestroy(e);
        ev_pool.clear();
    }
};

namespace {

int validate_csr(int64_t G, int64_t C, const int64_t *indptr, const void *indices, const void *vals,
                 int location, std::string &msg)
{
    if (G < 0 || C < 0 || C > 0x7fffffff || G > 0x7fffffff) { msg = "G and C must be in [0, 2^31)"; return SPADE_ERR_ARG; }
    if (!indptr) { msg = "indptr is NULL"; return SPADE_ERR_ARG; }
    if (location != SPADE_HOST && location != SPADE_DEVICE) { msg = "bad location"; return SPADE_ERR_ARG; }
    (void)indices; (void)vals;
    return SPADE_OK;
}

int create_common(int device, int64_t G, int64_t C, const int64_t *indptr, const int32_t *indices,
                  const double *values, const int32_t *counts, int location, spade_engine **out)
{
    auto fail = [&](int code, const std::string &m) { g_create_error = m; return code; };
    if (!out) return fail(SPADE_ERR_ARG, "engine_out is NULL");
    *out = nullptr;
    std::string msg;
    int rc = validate_csr(G, C, indptr, indices, values, location, msg);
    if (rc != SPADE_OK) return fail(rc, msg);
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0)
        return fail(SPADE_ERR_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(ce));
    if (device < 0 || device >= ndev) return fail(SPADE_ERR_ARG, "device index out of range");
    ce = cudaSetDevice(device);
    if (ce != cudaSuccess) return fail(SPADE_ERR_CUDA, cudaGetErrorString(ce));

    spade_engine *e = new spade_engine();
    e->device = device;
    e->G = G;
    e->C = C;
    auto bail = [&](int code) {
        g_create_error = e->error;
        e->free_all();
        delete e;
        return code;
    };
#define E_TRY(expr)                                                                            \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            e->set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));                  \
            return bail(SPADE_ERR_CUDA);                                                       \
        }                                                                                      \
    } while (0)
    const cudaMemcpyKind kind = location == SPADE_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
    cudaStream_t st = 0;
    E_TRY(e->indptr.reserve(G + 1));
    E_TRY(cudaMemcpy(e->indptr.p, indptr, (G + 1) * sizeof(int64_t), kind));
    int64_t first = 0, nnz = 0;
    E_TRY(cudaMemcpy(&first, e->indptr.p, sizeof(int64_t), cudaMemcpyDeviceToHost));
    E_TRY(cudaMemcpy(&nnz, e->indptr.p + G, sizeof(int64_t), cudaMemcpyDeviceToHost));
    if (first != 0 || nnz < 0 || nnz > 0xfffffff0ll) {
        e->set_error("indptr must start at 0 and hold fewer than 2^32 entries");
        return bail(SPADE_ERR_ARG);
    }
    if (nnz > 0 && (!indices || (!values && !counts))) {
        e->set_error("indices/values is NULL");
        return bail(SPADE_ERR_ARG);
    }
    e->nnz = nnz;
    E_TRY(e->indices.reserve(nnz));
    E_TRY(e->values.reserve(nnz));
    E_TRY(e->err.reserve(1));
    E_TRY(cudaMemset(e->err.p, 0, sizeof(int)));
    if (nnz > 0) {
        E_TRY(cudaMemcpy(e->indices.p, indices, nnz * sizeof(int32_t), kind));
        if (counts) {
            DevBuf<int32_t> dcounts;
            DevBuf<unsigned long long> colsum;
            E_TRY(dcounts.reserve(nnz));
            E_TRY(colsum.reserve(C));
            E_TRY(cudaMemcpy(dcounts.p, counts, nnz * sizeof(int32_t), kind));
            E_TRY(cudaMemsetAsync(colsum.p, 0, C * sizeof(unsigned long long), st));
            k_col_sums<<<grid_for(nnz, 256), 256, 0, st>>>(e->indices.p, dcounts.p, nnz, C, colsum.p, e->err.p);
            E_TRY(cudaGetLastError());
            k_cpm<<<grid_for(nnz, 256), 256, 0, st>>>(e->indices.p, dcounts.p, nnz, colsum.p, e->values.p);
            E_TRY(cudaGetLastError());
            E_TRY(cudaStreamSynchronize(st));
            dcounts.release();
            colsum.release();
            if (e->check_device_error("spade_create_from_counts") != SPADE_OK) return bail(SPADE_ERR_ARG);
        } else {
            E_TRY(cudaMemcpy(e->values.p, values, nnz * sizeof(double), kind));
        }
    }
#undef E_TRY
    rc = e->build(st);
    if (rc != SPADE_OK) return bail(rc);
    *out = e;
    return SPADE_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
extern "C" {

int spade_abi_version(void) { return SPADE_ABI_VERSION; }

const char *spade_last_error(const spade_engine *engine)
{
    return engine ? engine->error.c_str() : g_create_error.c_str();
}

int spade_create(int device, int64_t G, int64_t C, const int64_t *indptr, const int32_t *indices,
                 const double *values, int location, spade_engine **engine_out)
{
    return create_common(device, G, C, indptr, indices, values, nullptr, location, engine_out);
}

int spade_create_from_counts(int device, int64_t G, int64_t C, const int64_t *indptr,
                             const int32_t *indices, const int32_t *counts, int location,
                             spade_engine **engine_out)
{
    return create_common(device, G, C, indptr, indices, nullptr, counts, location, engine_out);
}

void spade_destroy(spade_engine *engine)
{
    if (!engine) return;
    cudaSetDevice(engine->device);
    engine->free_all();
    delete engine;
}

#define set_error engine->set_error

int spade_set_background(spade_engine *engine, const int32_t *nc_idx, int64_t n_nc)
{
    if (!engine) return SPADE_ERR_ARG;
    CUDA_TRY(cudaSetDevice(engine->device));
    if (n_nc < 0 || (n_nc > 0 && !nc_idx)) { set_error("bad nc_idx / n_nc"); return SPADE_ERR_ARG; }
    if (n_nc > 0) {
        std::vector<uint8_t> mask(engine->C, 0);
        for (int64_t i = 0; i < n_nc; ++i) {
            int32_t c = nc_idx[i];
            if (c < 0 || c >= engine->C) { set_error("NC cell index out of range"); return SPADE_ERR_ARG; }
            if (mask[c]) { set_error("duplicate NC cell index"); return SPADE_ERR_ARG; }
            mask[c] = 1;
        }
        CUDA_TRY(cudaMemcpy(engine->ncmask.p, mask.data(), engine->C, cudaMemcpyHostToDevice));
    }
    engine->n_nc = n_nc;
    int rc = engine->stage_genes(0);
    if (rc != SPADE_OK) return rc;
    CUDA_TRY(cudaStreamSynchronize(0));
    return SPADE_OK;
}

int spade_dims(const spade_engine *engine, int64_t *G, int64_t *C, int64_t *nnz, int32_t *fix_shift)
{
    if (!engine) return SPADE_ERR_ARG;
    if (G) *G = engine->G;
    if (C) *C = engine->C;
    if (nnz) *nnz = engine->nnz;
    if (fix_shift) *fix_shift = engine->fix_shift;
    return SPADE_OK;
}

int spade_gene_cutoffs(spade_engine *engine, double *median, int32_t *n)
{
    if (!engine) return SPADE_ERR_ARG;
    CUDA_TRY(cudaSetDevice(engine->device));
    if (median) CUDA_TRY(cudaMemcpy(median, engine->median.p, engine->G * sizeof(double), cudaMemcpyDeviceToHost));
    if (n) CUDA_TRY(cudaMemcpy(n, engine->n.p, engine->G * sizeof(int32_t), cudaMemcpyDeviceToHost));
    return SPADE_OK;
}

int spade_cpm_values(spade_engine *engine, double *values)
{
    if (!engine || !values) return SPADE_ERR_ARG;
    CUDA_TRY(cudaSetDevice(engine->device));
    CUDA_TRY(cudaMemcpy(values, engine->values.p, engine->nnz * sizeof(double), cudaMemcpyDeviceToHost));
    return SPADE_OK;
}

int spade_run(spade_engine *engine, const int32_t *members, int members_location,
              const int64_t *offsets, int64_t S, double *up, double *down, double *fc, double *cpm,
              int32_t *k, int out_location, void *stream)
{
    if (!engine) return SPADE_ERR_ARG;
    CUDA_TRY(cudaSetDevice(engine->device));
    if (S < 0 || !offsets) { set_error("offsets is NULL or S < 0"); return SPADE_ERR_ARG; }
    if ((members_location != SPADE_HOST && members_location != SPADE_DEVICE) ||
        (out_location != SPADE_HOST && out_location != SPADE_DEVICE)) {
        set_error("bad location argument");
        return SPADE_ERR_ARG;
    }
    if (offsets[0] != 0) { set_error("offsets[0] must be 0"); return SPADE_ERR_ARG; }
    for (int64_t s = 0; s < S; ++s)
        if (offsets[s + 1] < offsets[s]) { set_error("offsets must be non-decreasing"); return SPADE_ERR_ARG; }
    const int64_t total = offsets[S];
    if (total > 0 && !members) { set_error("members is NULL"); return SPADE_ERR_ARG; }
    if (S == 0 || engine->G == 0) return SPADE_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t G = engine->G;

    const int32_t *d_members = members;
    if (members_location == SPADE_HOST) {
        CUDA_TRY(engine->members.reserve(total));
        if (total > 0)
            CUDA_TRY(cudaMemcpyAsync(engine->members.p, members, total * sizeof(int32_t),
                                     cudaMemcpyHostToDevice, st));
        d_members = engine->members.p;
    }
    CUDA_TRY(engine->offsets.reserve(S + 1));
    CUDA_TRY(cudaMemcpyAsync(engine->offsets.p, offsets, (S + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, st));

    const bool host_out = out_location == SPADE_HOST;
    const int64_t chunk_cap = host_out ? kMaxTestsPerChunkHost : kMaxTestsPerChunk;
    const int64_t chunk_sets = std::max<int64_t>(1, std::min<int64_t>(S, chunk_cap / G));
    CUDA_TRY(engine->acc_sum.reserve(chunk_sets * G));
    CUDA_TRY(engine->acc_cnt.reserve(chunk_sets * G));
    if (host_out) {
        // results leave through two device buffers: chunk i is copied to the host on
        // copy_stream while chunk i+1 is computed on the caller's stream
        if (!engine->copy_stream)
            CUDA_TRY(cudaStreamCreateWithFlags(&engine->copy_stream, cudaStreamNonBlocking));
        for (int b = 0; b < 2; ++b) {
            CUDA_TRY(engine->out_f64[b].reserve(4 * chunk_sets * G));
            if (k) CUDA_TRY(engine->out_k[b].reserve(chunk_sets * G));
            if (!engine->buf_ready[b]) CUDA_TRY(cudaEventCreateWithFlags(&engine->buf_ready[b], cudaEventDisableTiming));
            if (!engine->buf_free[b]) CUDA_TRY(cudaEventCreateWithFlags(&engine->buf_free[b], cudaEventDisableTiming));
        }
    }
    int64_t chunk_no = 0;

    for (int64_t s0 = 0; s0 < S; s0 += chunk_sets) {
        const int64_t s1 = std::min(S, s0 + chunk_sets);
        const int64_t tests = (s1 - s0) * G;
        const int64_t nmem = offsets[s1] - offsets[s0];
        const int buf = (int)(chunk_no & 1);
        if (host_out && chunk_no >= 2) CUDA_TRY(cudaStreamWaitEvent(st, engine->buf_free[buf], 0));
        ++chunk_no;
        cudaEvent_t e0 = engine->next_event(), e1 = engine->next_event(), e2 = engine->next_event();
        if (!e0 || !e1 || !e2) { set_error("cudaEventCreate failed"); return SPADE_ERR_CUDA; }
        CUDA_TRY(cudaEventRecord(e0, st));
        CUDA_TRY(cudaMemsetAsync(engine->acc_sum.p, 0, tests * sizeof(unsigned long long), st));
        CUDA_TRY(cudaMemsetAsync(engine->acc_cnt.p, 0, tests * sizeof(int), st));
        if (nmem > 0) {
            const int64_t blocks = (nmem + kAccWarpsPerBlock - 1) / kAccWarpsPerBlock;
            k_accumulate<<<(unsigned)blocks, kAccWarpsPerBlock * 32, 0, st>>>(
                engine->colptr.p, engine->ent.p, d_members, engine->offsets.p, s0, s1, G, engine->C,
                engine->fix_scale, engine->acc_sum.p, engine->acc_cnt.p, engine->err.p);
            CUDA_TRY(cudaGetLastError());
            engine->launches += 1;
        }
        CUDA_TRY(cudaEventRecord(e1, st));
        FinalArgs a;
        a.acc_sum = engine->acc_sum.p;
        a.acc_cnt = engine->acc_cnt.p;
        a.offsets = engine->offsets.p;
        a.s0 = s0; a.s1 = s1; a.G = G; a.C = engine->C;
        a.n = engine->n.p;
        a.negmode = engine->negmode.p;
        a.rowsum = engine->rowsum.p;
        a.bgmean = engine->bgmean.p;
        a.lf = engine->lf.p;
        a.nc_mode = engine->n_nc > 0;
        a.fix_inv = engine->fix_inv;
        const int64_t cs = chunk_sets * G;
        if (host_out) {
            a.up = up ? engine->out_f64[buf].p : nullptr;
            a.down = down ? engine->out_f64[buf].p + cs : nullptr;
            a.fc = fc ? engine->out_f64[buf].p + 2 * cs : nullptr;
            a.cpm = cpm ? engine->out_f64[buf].p + 3 * cs : nullptr;
            a.k = k ? engine->out_k[buf].p : nullptr;
        } else {
            a.up = up ? up + s0 * G : nullptr;
            a.down = down ? down + s0 * G : nullptr;
            a.fc = fc ? fc + s0 * G : nullptr;
            a.cpm = cpm ? cpm + s0 * G : nullptr;
            a.k = k ? k + s0 * G : nullptr;
        }
        k_finalize<<<(unsigned)((tests + kFinThreads - 1) / kFinThreads), kFinThreads, 0, st>>>(a);
        CUDA_TRY(cudaGetLastError());
        engine->launches += 1;
        CUDA_TRY(cudaEventRecord(e2, st));
        if (host_out) {
            const size_t fb = tests * sizeof(double);
            cudaStream_t cs_ = engine->copy_stream;
            CUDA_TRY(cudaEventRecord(engine->buf_ready[buf], st));
            CUDA_TRY(cudaStreamWaitEvent(cs_, engine->buf_ready[buf], 0));
            if (up) CUDA_TRY(cudaMemcpyAsync(up + s0 * G, a.up, fb, cudaMemcpyDeviceToHost, cs_));
            if (down) CUDA_TRY(cudaMemcpyAsync(down + s0 * G, a.down, fb, cudaMemcpyDeviceToHost, cs_));
            if (fc) CUDA_TRY(cudaMemcpyAsync(fc + s0 * G, a.fc, fb, cudaMemcpyDeviceToHost, cs_));
            if (cpm) CUDA_TRY(cudaMemcpyAsync(cpm + s0 * G, a.cpm, fb, cudaMemcpyDeviceToHost, cs_));
            if (k) CUDA_TRY(cudaMemcpyAsync(k + s0 * G, a.k, tests * sizeof(int32_t), cudaMemcpyDeviceToHost, cs_));
            CUDA_TRY(cudaEventRecord(engine->buf_free[buf], cs_));
        }
        if (engine->ev_used > 3000) {            // bound the event pool on very long runs
            int rc = engine->collect_timing();
            if (rc != SPADE_OK) return rc;
        }
    }
    if (host_out) {
        CUDA_TRY(cudaStreamSynchronize(engine->copy_stream));
        CUDA_TRY(cudaStreamSynchronize(st));
        return engine->check_device_error("spade_run");
    }
    return SPADE_OK;
}

int spade_timing(spade_engine *engine, double *accumulate_ms, double *finalize_ms, int64_t *launches,
                 int reset)
{
    if (!engine) return SPADE_ERR_ARG;
    CUDA_TRY(cudaSetDevice(engine->device));
    CUDA_TRY(cudaDeviceSynchronize());
    int rc = engine->collect_timing();
    if (rc != SPADE_OK) return rc;
    rc = engine->check_device_error("spade_run");
    if (accumulate_ms) *accumulate_ms = engine->acc_ms;
    if (finalize_ms) *finalize_ms = engine->fin_ms;
    if (launches) *launches = engine->launches;
    if (reset) {
        engine->acc_ms = engine->fin_ms = 0.0;
        engine->launches = 0;
    }
    return rc;
}

#undef set_error

int spade_hypergeom(int device, const int64_t *k, const int64_t *M, const int64_t *n,
                    const int64_t *N, int64_t count, double *logcdf, double *logsf)
{
    auto fail = [&](int code, const std::string &m) { g_create_error = m; return code; };
    if (count < 0 || (count > 0 && (!k || !M || !n || !N || !logcdf || !logsf)))
        return fail(SPADE_ERR_ARG, "NULL argument");
    if (count == 0) return SPADE_OK;
    cudaError_t ce = cudaSetDevice(device);
    if (ce != cudaSuccess) return fail(SPADE_ERR_CUDA, cudaGetErrorString(ce));
    int64_t maxM = 0;
    for (int64_t i = 0; i < count; ++i) {
        if (M[i] > 0x7ffffff0ll) return fail(SPADE_ERR_ARG, "M too large");
        maxM = std::max(maxM, M[i]);
    }
    std::vector<double> h(maxM + 1);
    for (int64_t i = 0; i <= maxM; ++i) {
        int sg;
        h[i] = lgamma_r((double)i + 1.0, &sg);
    }
    DevBuf<int64_t> dk, dM, dn, dN;
    DevBuf<double> dlf, dc, ds;
#define H_TRY(expr)                                                                            \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            dk.release(); dM.release(); dn.release(); dN.release();                            \
            dlf.release(); dc.release(); ds.release();                                         \
            return fail(SPADE_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));   \
        }                                                                                      \
    } while (0)
    H_TRY(dk.reserve(count)); H_TRY(dM.reserve(count)); H_TRY(dn.reserve(count)); H_TRY(dN.reserve(count));
    H_TRY(dlf.reserve(maxM + 1)); H_TRY(dc.reserve(count)); H_TRY(ds.reserve(count));
    H_TRY(cudaMemcpy(dk.p, k, count * sizeof(int64_t), cudaMemcpyHostToDevice));
    H_TRY(cudaMemcpy(dM.p, M, count * sizeof(int64_t), cudaMemcpyHostToDevice));
    H_TRY(cudaMemcpy(dn.p, n, count * sizeof(int64_t), cudaMemcpyHostToDevice));
    H_TRY(cudaMemcpy(dN.p, N, count * sizeof(int64_t), cudaMemcpyHostToDevice));
    H_TRY(cudaMemcpy(dlf.p, h.data(), (maxM + 1) * sizeof(double), cudaMemcpyHostToDevice));
    k_hyper_probe<<<(unsigned)((count + 127) / 128), 128>>>(dk.p, dM.p, dn.p, dN.p, count, dlf.p, dc.p, ds.p);
    H_TRY(cudaGetLastError());
    H_TRY(cudaMemcpy(logcdf, dc.p, count * sizeof(double), cudaMemcpyDeviceToHost));
    H_TRY(cudaMemcpy(logsf, ds.p, count * sizeof(double), cudaMemcpyDeviceToHost));
#undef H_TRY
    dk.release(); dM.release(); dn.release(); dN.release(); dlf.release(); dc.release(); ds.release();
    return SPADE_OK;
}

int spade_host_alloc(void **ptr, int64_t bytes)
{
    if (!ptr || bytes < 0) return SPADE_ERR_ARG;
    cudaError_t e = cudaHostAlloc(ptr, (size_t)std::max<int64_t>(bytes, 1), cudaHostAllocDefault);
    if (e != cudaSuccess) {
        g_create_error = std::string("cudaHostAlloc: ") + cudaGetErrorString(e);
        return SPADE_ERR_CUDA;
    }
    return SPADE_OK;
}

int spade_host_free(void *ptr)
{
    if (!ptr) return SPADE_OK;
    return cudaFreeHost(ptr) == cudaSuccess ? SPADE_OK : SPADE_ERR_CUDA;
}

}  // extern "C"
