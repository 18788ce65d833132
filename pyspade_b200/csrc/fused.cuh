// K2 + K4 + K3 fused: the batched test.  One warp owns one (gene partition, cell set) pair:
// it walks the partition's cell-major segments of the set's member cells, adds the packed
// entries into its PRIVATE shared-memory accumulators (plain 64-bit read-modify-write, no
// atomics: genes are unique inside a cell's segment and a warp handles one cell at a time),
// then turns every accumulator into (k, mean CPM, fold change, log-hypergeometric tails)
// and writes the result tables.  Replaces G*S calls of hypergeo_test_sparse /
// hypergeo_test_NC_sparse (pySpade/utils.py:194-215, 248-266).
#pragma once
#include "common.cuh"
#include "tails.cuh"

namespace spade {

constexpr int kWarpsPerCta = 4;
constexpr int kMemberBatch = 4;   // member cells whose segments are in flight per warp
constexpr int kIters = 3;         // unrolled 32-entry steps per member segment

struct FinCtx {
    int64_t G;
    int M;                        // population size = number of cells
    int nc_mode;
    const GeneMeta *meta;
    const TabMeta *tmeta;         // null -> evaluate every tail directly
    const double2 *tab;
    const double *lf;
    double *up, *down, *fc, *cpm; // row 0 = first set of the launch; any may be null
    int32_t *k;
    int64_t ld;                   // row stride (elements)
};

// One (gene, set) test from its member sum / member count (utils.py:203-215, 257-266).
__device__ __forceinline__ void finalize_test(const FinCtx &f, int64_t row, int64_t g, long long sumfx,
                                              int cnt, long long N, double invN, double invRest)
{
    const GeneMeta mt = f.meta[g];
    const int kk = mt.negmode ? cnt : (int)(N - cnt);
    const double inv = pow2_double(-mt.shift);
    const double mean_set = ((double)sumfx * inv) * invN;
    const int64_t o = row * f.ld + g;
    if (f.cpm) f.cpm[o] = mean_set;
    if (f.fc) {
        const double mean_bg = f.nc_mode ? mt.bgmean : ((double)(mt.rowsum - sumfx) * inv) * invRest;
        f.fc[o] = (mean_set + 0.01) / (mean_bg + 0.01);
    }
    if (f.k) f.k[o] = kk;
    if (f.up || f.down) {
        double lc, ls;
        if (kk == 0 || f.M == 0 || mt.n == 0 || N == 0) {       // all([k, M, n, N]), utils.py:212-213
            lc = ls = __longlong_as_double(0x7ff8000000000000ll);
        } else {
            bool hit = false;
            if (f.tmeta) {
                const TabMeta tm = f.tmeta[g];
                const unsigned d = (unsigned)(kk - tm.lo);
                if (d < (unsigned)tm.len) {
                    const double2 v = f.tab[tm.off + d];
                    lc = v.x;
                    ls = v.y;
                    hit = true;
                }
            }
            if (!hit) hyper_tails(kk, f.M, mt.n, (int)N, f.lf, lc, ls);
        }
        if (f.up) f.up[o] = lc;
        if (f.down) f.down[o] = ls;
    }
}

struct RunArgs {
    const unsigned long long *ent;
    const uint32_t *pptr;
    const int32_t *members;
    const int64_t *offsets;       // device copy of the caller's offsets
    int64_t s0;                   // first set of this launch
    int64_t S;                    // sets of this launch
    // split mode (some set larger than the member block): work units are member blocks
    const int32_t *blk_row;       // set index relative to s0
    const int64_t *blk_off;       // absolute offset into members
    const int32_t *blk_cnt;
    int64_t nblk;
    unsigned long long *gsum;     // [S][G] global accumulators of split mode
    int *gcnt;
    int p0, p1;                   // gene partitions of this launch
    int Gp;
    int64_t C;
    FinCtx fin;
    int *err;
};

__device__ __forceinline__ void decode_entry(unsigned long long e, uint32_t &slot, unsigned long long &add)
{
    const uint32_t hi = (uint32_t)(e >> 32);
    slot = hi >> (kFieldBits - 32);
    const int32_t hs = ((int32_t)(hi << kSlotBits)) >> kSlotBits;      // sign-extend the 51-bit field
    add = ((unsigned long long)(uint32_t)hs << 32) | (uint32_t)e;
}

template <bool kSplit>
__global__ void __launch_bounds__(kWarpsPerCta * 32) k_de_fused(const RunArgs a)
{
    extern __shared__ unsigned long long smem_acc[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long *acc = smem_acc + (size_t)warp * a.Gp;

    const int64_t units = kSplit ? a.nblk : a.S;
    const int64_t t = blockIdx.x * (int64_t)kWarpsPerCta + warp;
    const int64_t pi = t / units;
    if (pi >= a.p1 - a.p0) return;
    const int p = a.p0 + (int)pi;
    const int64_t u = t - pi * units;

    int64_t row, mbeg;
    int cntm;
    long long N = 0;
    if (kSplit) {
        row = a.blk_row[u];
        mbeg = a.blk_off[u];
        cntm = a.blk_cnt[u];
    } else {
        row = u;
        mbeg = a.offsets[a.s0 + u];
        N = a.offsets[a.s0 + u + 1] - mbeg;
        cntm = (int)N;
    }
    const int64_t g0 = (int64_t)p * a.Gp;
    const int ng = (int)min((int64_t)a.Gp, a.fin.G - g0);
    for (int j = lane; j < ng; j += 32) acc[j] = 0ull;
    __syncwarp();

    const uint32_t *__restrict__ pp = a.pptr + (int64_t)p * a.C;
    const int32_t *__restrict__ mem = a.members + mbeg;
    const unsigned long long *__restrict__ ent = a.ent;

    for (int m0 = 0; m0 < cntm; m0 += 32) {
        uint32_t b = 0, e = 0;
        if (m0 + lane < cntm) {
            const int32_t c = mem[m0 + lane];
            if ((uint32_t)c < (uint32_t)a.C) {
                b = __ldg(pp + c);
                e = __ldg(pp + c + 1);
            } else {
                atomicExch(a.err, 2);
            }
        }
        const int lim = min(32, cntm - m0);
        for (int j = 0; j < lim; j += kMemberBatch) {
            uint32_t bb[kMemberBatch], ee[kMemberBatch];
            unsigned long long r[kMemberBatch][kIters];
#pragma unroll
            for (int q = 0; q < kMemberBatch; ++q) {
                bb[q] = __shfl_sync(0xffffffffu, b, j + q) + lane;
                ee[q] = __shfl_sync(0xffffffffu, e, j + q);
            }
#pragma unroll
            for (int q = 0; q < kMemberBatch; ++q)
#pragma unroll
                for (int it = 0; it < kIters; ++it) {
                    const uint32_t i = bb[q] + 32 * it;
                    r[q][it] = i < ee[q] ? __ldg(ent + i) : 0ull;
                }
#pragma unroll
            for (int q = 0; q < kMemberBatch; ++q) {
                uint32_t slot[kIters];
                unsigned long long add[kIters], cur[kIters];
                bool ok[kIters];
#pragma unroll
                for (int it = 0; it < kIters; ++it) {
                    ok[it] = bb[q] + 32 * it < ee[q];
                    decode_entry(r[q][it], slot[it], add[it]);
                }
                // the genes of one cell are distinct: the loads of all steps may precede the stores
#pragma unroll
                for (int it = 0; it < kIters; ++it)
                    if (ok[it]) cur[it] = acc[slot[it]];
#pragma unroll
                for (int it = 0; it < kIters; ++it)
                    if (ok[it]) acc[slot[it]] = cur[it] + add[it];
                for (uint32_t i = bb[q] + 32 * kIters; i < ee[q]; i += 32) {     // long segments
                    uint32_t sl;
                    unsigned long long ad;
                    decode_entry(__ldg(ent + i), sl, ad);
                    acc[sl] += ad;
                }
                __syncwarp();      // next cell may hit the same genes from other lanes
            }
        }
    }

    if (!kSplit) {
        const double invN = 1.0 / (double)N;
        const double invRest = 1.0 / (double)((long long)a.C - N);
        for (int j = lane; j < ng; j += 32) {
            const long long v = (long long)acc[j];
            const long long sumfx = (v << 16) >> 16;
            const int cnt = (int)((unsigned long long)(v - sumfx) >> kCountShift);
            finalize_test(a.fin, row, g0 + j, sumfx, cnt, N, invN, invRest);
        }
    } else {
        for (int j = lane; j < ng; j += 32) {
            const long long v = (long long)acc[j];
            if (v == 0) continue;
            const long long sumfx = (v << 16) >> 16;
            const int cnt = (int)((unsigned long long)(v - sumfx) >> kCountShift);
            const int64_t o = row * a.fin.G + g0 + j;
            if (sumfx) atomicAdd(&a.gsum[o], (unsigned long long)sumfx);
            if (cnt) atomicAdd(&a.gcnt[o], cnt);
        }
    }
}

// Split mode only: the tests of a batch from its global accumulators.
__global__ void __launch_bounds__(256)
k_finalize_split(const FinCtx f, const unsigned long long *__restrict__ gsum,
                 const int *__restrict__ gcnt, const int64_t *__restrict__ offsets, int64_t s0,
                 int64_t S)
{
    const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (idx >= S * f.G) return;
    const int64_t row = idx / f.G, g = idx - row * f.G;
    const long long N = offsets[s0 + row + 1] - offsets[s0 + row];
    finalize_test(f, row, g, (long long)gsum[idx], gcnt[idx], N, 1.0 / (double)N,
                  1.0 / (double)((long long)f.M - N));
}

}  // namespace spade
