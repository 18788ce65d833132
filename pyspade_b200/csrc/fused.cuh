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

#ifndef SPADE_MEMBER_BATCH
#define SPADE_MEMBER_BATCH 4
#endif
#ifndef SPADE_ITERS
#define SPADE_ITERS 3
#endif
#ifndef SPADE_FIN_BATCH
#define SPADE_FIN_BATCH 2
#endif
#ifndef SPADE_MIN_CTAS
#define SPADE_MIN_CTAS 6
#endif
// Streaming 64-bit load of layout data.  The kernel keeps almost all of the SM's 228 KB as
// shared memory, so L1 has only a few hundred lines; ld.global.cg goes through L2 only.
__device__ __forceinline__ unsigned long long stream_load(const unsigned long long *p)
{
    return __ldcg(p);
}

constexpr int kWarpsPerCta = 4;
constexpr int kFinBatch = SPADE_FIN_BATCH;           // genes per lane whose finalize loads are in flight together
constexpr int kMemberBatch = SPADE_MEMBER_BATCH;   // member cells whose segments are in flight per warp
constexpr int kIters = SPADE_ITERS;                // unrolled 32-entry steps per member segment
static_assert(32 % kMemberBatch == 0, "a batch must not straddle two 32-member groups");
// (SPADE_* macros: tuning variants built next to the product library, see _build.build(defines=...))

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
    int64_t k_ld;                 // row stride of k; 0 = ld
};

// One (gene, set) test from its member sum / member count (utils.py:203-215, 257-266), in two
// steps so that a lane can have the metadata and tail-table loads of several genes in flight:
// fin_load issues the loads (gene metadata, table window, the table entry when k is inside the
// window), fin_store does the arithmetic and the stores.
// The float arithmetic of a finished test, with explicit roundings (no FMA contraction): the
// same test gives the same bits from every kernel that finishes it (fused, split, expand).
__device__ __forceinline__ double fixed_mean(long long sumfx, int shift, double inv_count)
{
    return __dmul_rn(__dmul_rn((double)sumfx, pow2_double(-shift)), inv_count);
}

__device__ __forceinline__ double fold_change(double mean_set, double mean_bg)    // utils.py:207 / :258
{
    return __ddiv_rn(__dadd_rn(mean_set, 0.01), __dadd_rn(mean_bg, 0.01));
}

struct FinItem {
    GeneMeta mt;
    double2 tv;
    long long sumfx;
    int kk;
    int how;          // 0 nan, 1 upper support edge, 2 table entry in tv, 3 evaluate directly
};

__device__ __forceinline__ void fin_load(const FinCtx &f, int64_t g, long long sumfx, int cnt, long long N,
                                         FinItem &it)
{
    it.mt = f.meta[g];
    it.sumfx = sumfx;
    it.kk = it.mt.negmode ? cnt : (int)(N - cnt);
    it.how = 3;
    if (!(f.up || f.down)) return;
    if (it.kk == 0 || f.M == 0 || it.mt.n == 0 || N == 0) {     // all([k, M, n, N]), utils.py:212-213
        it.how = 0;
    } else if (it.kk >= min(it.mt.n, (int)N)) {                 // upper support edge: no lookup needed
        it.how = 1;                                              // (what hyper_tails returns first)
    } else if (f.tmeta) {
        const TabMeta tm = f.tmeta[g];
        const unsigned d = (unsigned)(it.kk - tm.lo);
        if (d < (unsigned)tm.len) {
            it.tv = f.tab[tm.off + d];
            it.how = 2;
        }
    }
}

__device__ __forceinline__ void fin_store(const FinCtx &f, int64_t row, int64_t g, const FinItem &it, long long N,
                                          double invN, double invRest)
{
    const double mean_set = fixed_mean(it.sumfx, it.mt.shift, invN);
    const int64_t o = row * f.ld + g;
    if (f.cpm) f.cpm[o] = mean_set;
    if (f.fc) {
        const double mean_bg = f.nc_mode ? it.mt.bgmean : fixed_mean(it.mt.rowsum - it.sumfx, it.mt.shift, invRest);
        f.fc[o] = fold_change(mean_set, mean_bg);
    }
    if (f.k) f.k[f.k_ld ? row * f.k_ld + g : o] = it.kk;
    if (f.up || f.down) {
        double lc, ls;
        if (it.how == 0) {
            lc = ls = __longlong_as_double(0x7ff8000000000000ll);
        } else if (it.how == 1) {
            lc = 0.0;
            ls = -INFINITY;
        } else if (it.how == 2) {
            lc = it.tv.x;
            ls = it.tv.y;
        } else {
            hyper_tails(it.kk, f.M, it.mt.n, (int)N, f.lf, lc, ls);
        }
        if (f.up) f.up[o] = lc;
        if (f.down) f.down[o] = ls;
    }
}

__device__ __forceinline__ void finalize_test(const FinCtx &f, int64_t row, int64_t g, long long sumfx,
                                              int cnt, long long N, double invN, double invRest)
{
    FinItem it;
    fin_load(f, g, sumfx, cnt, N, it);
    fin_store(f, row, g, it, N, invN, invRest);
}

constexpr int kMaxGeneBlocks = 16;   // destinations of the routed raw form (GPUs of one node)

struct RunArgs {
    const unsigned long long *ent;
    const uint2 *seg;             // [P*C] (begin, end) of every (partition, cell) segment
    const int32_t *dgene;         // [P*32] gene held by each lane of the dense head, -1 = none
    const uint8_t *dense_on;      // [P] partition uses its dense block
    const int32_t *members;
    const int64_t *offsets;       // device copy of the caller's offsets
    int64_t s0;                   // first set of this launch
    int64_t S;                    // sets of this launch
    const int32_t *order;         // ragged batches: sets (relative to s0) by decreasing size, so that the
                                  // long tasks of a partition start first; null = as given
    // split mode (some set larger than the member block): work units are member blocks
    const int32_t *blk_row;       // set index relative to s0
    const int64_t *blk_off;       // absolute offset into members
    const int32_t *blk_cnt;
    int64_t nblk;
    unsigned long long *gsum;     // [S][G] global accumulators of split mode
    int *gcnt;
    unsigned long long *raw;      // non-null: store the accumulator words instead of finished tests
    int64_t raw_ld;
    // raw form routed by gene block (multi-GPU exchange): partitions [dst_pbegin[b], dst_pbegin[b+1])
    // go to table dst_base[b] (row stride dst_ld[b], column 0 = first gene of the block), set u of
    // the launch to row raw_row0 + u * raw_row_step.  ndst == 0: the single table `raw`.
    int ndst;
    unsigned long long *dst_base[kMaxGeneBlocks];
    int64_t dst_ld[kMaxGeneBlocks];
    int dst_pbegin[kMaxGeneBlocks + 1];
    int64_t raw_row0, raw_row_step;
    int p0, p1;                   // gene partitions of this launch
    int coop;                     // warps of a CTA that share one (partition, set) task: 1, 2 or 4 (non-split form).
                                  // Each takes a contiguous part of the member list; the accumulators meet in shared memory.
    int Gp;
    int two;                      // hot genes per partition (common.cuh), 0 = none: a test is finished from A + B
    const int16_t *hot_slot;      // [P * two] slot of the gene whose B accumulator is Gp + b, negative = unused
    int acc_len;                  // accumulators per warp: Gp + two
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

// The register accumulators of the dense head join the shared-memory ones, and the second
// accumulators of the hot genes are added into the first.
__device__ __forceinline__ void collect_accumulators(const RunArgs &a, unsigned long long *acc,
                                                const unsigned long long (&dacc)[kDenseRegs], bool dense, int lane,
                                                int p, int64_t g0)
{
    if (a.two) {
        __syncwarp();
        const int16_t *__restrict__ hs = a.hot_slot + (int64_t)p * a.two;
        for (int b = lane; b < a.two; b += 32) {
            const int sl = hs[b];
            if (sl >= 0) acc[sl] += acc[a.Gp + b];
        }
        __syncwarp();
    }
    if (!dense) return;
#pragma unroll
    for (int w2 = 0; w2 < kDenseRegs; ++w2) {
        const int32_t dg = a.dgene[p * kDenseLanes + lane + 32 * w2];
        if (dg >= 0) acc[dg - g0] = dacc[w2];
    }
    __syncwarp();
}

// What a warp does with the finished accumulators of its (partition p, set `row`) task: the raw
// words to the owner of the gene block (routed), the raw words to one table, or the finished tests.
// kRawOnly: instantiation without the finalize code (raw forms only; fewer registers).
template <bool kRawOnly = false, int stride = 32>
__device__ __forceinline__ void finish_task(const RunArgs &a, const unsigned long long *acc, int lane,
                                            int p, int64_t g0, int ng, int64_t row, long long N)
{   // `lane` = index of the thread among the `stride` threads that finish this task together
    // the accumulator word of gene j of the partition (count << 48 | sum)
    auto word = [&](int j) -> unsigned long long { return acc[j]; };
    if (a.ndst > 0) {
        // raw form, routed: this partition's words go to the owner of its gene block (a table in
        // this GPU's or a peer GPU's memory; coalesced 256-byte stores over NVLink)
        int b = 0;
        while (b + 1 < a.ndst && p >= a.dst_pbegin[b + 1]) ++b;
        unsigned long long *dst = a.dst_base[b] + (a.raw_row0 + row * a.raw_row_step) * a.dst_ld[b] +
                                  (g0 - (int64_t)a.dst_pbegin[b] * a.Gp);
        for (int j = lane; j < ng; j += stride) dst[j] = word(j);
    } else if (kRawOnly || a.raw) {
        // raw form: the accumulator words leave as they are (8 bytes per test instead of 24-36);
        // whoever receives them finishes the tests with k_expand
        for (int j = lane; j < ng; j += stride) a.raw[row * a.raw_ld + g0 + j] = word(j);
    } else if (!kRawOnly) {
        const double invN = 1.0 / (double)N;
        const double invRest = 1.0 / (double)((long long)a.C - N);
        // kFinBatch genes per lane at a time: their metadata / table loads overlap
        for (int j0 = lane; j0 < ng; j0 += stride * kFinBatch) {
            FinItem it[kFinBatch];
#pragma unroll
            for (int u = 0; u < kFinBatch; ++u) {
                const int j = j0 + stride * u;
                if (j < ng) {
                    const long long v = (long long)word(j);
                    const long long sumfx = (v << 16) >> 16;
                    const int cnt = (int)((unsigned long long)(v - sumfx) >> kCountShift);
                    fin_load(a.fin, g0 + j, sumfx, cnt, N, it[u]);
                }
            }
#pragma unroll
            for (int u = 0; u < kFinBatch; ++u) {
                const int j = j0 + stride * u;
                if (j < ng) fin_store(a.fin, row, g0 + j, it[u], N, invN, invRest);
            }
        }
    }
}

template <bool kSplit, bool kRawOnly = false, int kCoop = 1>
__global__ void __launch_bounds__(kWarpsPerCta * 32, SPADE_MIN_CTAS) k_de_fused(const RunArgs a)
{
    extern __shared__ unsigned long long smem_acc[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long *acc = smem_acc + (size_t)warp * a.acc_len;
    uint2 *mbuf = reinterpret_cast<uint2 *>(smem_acc + (size_t)kWarpsPerCta * a.acc_len) + warp * 32;

    // Few sets per partition (DEobs: hundreds of regions): `coop` warps of the CTA share a task, so
    // that the resident warps work on two or three gene partitions instead of seven and the cell-major
    // layout of those partitions stays in L2 (and a long region is cut into shorter member lists).
    constexpr int coop = kSplit ? 1 : kCoop;   // (= a.coop)
    const int part = warp % coop;              // which part of the member list
    const int64_t units = kSplit ? a.nblk : a.S;
    const int64_t t = blockIdx.x * (int64_t)(kWarpsPerCta / coop) + warp / coop;
    const int64_t pi = t / units;
    if (pi >= a.p1 - a.p0) return;             // (all warps of a task group leave together)
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
        row = a.order ? a.order[u] : u;
        mbeg = a.offsets[a.s0 + row];
        N = a.offsets[a.s0 + row + 1] - mbeg;
        cntm = (int)N;
        if (coop > 1) {
            const int chunk = (cntm + coop - 1) / coop;
            const int first = min(cntm, part * chunk);
            cntm = min(cntm, first + chunk) - first;
            mbeg += first;
        }
    }
    const int64_t g0 = (int64_t)p * a.Gp;
    const int ng = (int)min((int64_t)a.Gp, a.fin.G - g0);
    for (int j = lane; j < a.acc_len; j += 32) acc[j] = 0ull;
    __syncwarp();
    const uint2 *__restrict__ seg = a.seg + (int64_t)p * a.C;
    const int32_t *__restrict__ mem = a.members + mbeg;
    const unsigned long long *__restrict__ ent = a.ent;
    // dense head of this partition's segments: lane l accumulates the partition's l-th densest
    // gene in a register, one aligned 256-byte read per member cell, no shared-memory traffic
    const bool dense = a.dense_on[p] != 0;
    const uint32_t head = dense ? kDenseLanes : 0;
    unsigned long long dacc[kDenseRegs];
#pragma unroll
    for (int r = 0; r < kDenseRegs; ++r) dacc[r] = 0ull;

    // Batches of kMemberBatch member cells: all segment loads of a batch are issued, then the
    // batch is added into the accumulators cell by cell.  Member ids and their segment bounds
    // are fetched 32 members at a time, one group ahead (two register sets by parity).
    // (Keeping a second batch in flight in registers was measured slower: the SM is bound by
    // the LSU data pipe, not by latency.)
    uint32_t mb[2] = {0, 0}, me[2] = {0, 0};
    auto fetch_meta = [&](int grp, uint32_t &b, uint32_t &e) {
        b = 0; e = 0;                                  // b == e == 0: no member in this lane
        const int m = grp * 32 + lane;
        if (m < cntm) {
            const int32_t c = mem[m];
            if ((uint32_t)c < (uint32_t)a.C) {
                const uint2 se = __ldg(seg + c);
                b = se.x;
                e = se.y;
            } else {
                flag_error(a.err, 2);
            }
        }
    };
    struct Batch {
        uint32_t bb[kMemberBatch], ee[kMemberBatch];
        unsigned long long r[kMemberBatch][kIters], dv[kMemberBatch][kDenseRegs];
    };
    constexpr int kBatchesPerGroup = 32 / kMemberBatch;
    auto load = [&](Batch &t, int ib) {
        const int grp = ib / kBatchesPerGroup, w = (ib % kBatchesPerGroup) * kMemberBatch;
        if (w == 0) {
            // entering a group: park its 32 (begin, end) pairs in shared memory - one broadcast
            // read per member costs a quarter of the two shuffles it replaces on the LSU data
            // pipe - and prefetch the next group's pairs into the other register set
            __syncwarp();
            mbuf[lane] = (grp & 1) ? make_uint2(mb[1], me[1]) : make_uint2(mb[0], me[0]);
            __syncwarp();
            if (grp & 1) fetch_meta(grp + 1, mb[0], me[0]);
            else fetch_meta(grp + 1, mb[1], me[1]);
        }
#pragma unroll
        for (int q = 0; q < kMemberBatch; ++q) {
            const uint2 se = mbuf[w + q];
            t.ee[q] = se.y;
#pragma unroll
            for (int w2 = 0; w2 < kDenseRegs; ++w2)
                t.dv[q][w2] = (dense && se.y != 0) ? stream_load(ent + se.x + lane + 32 * w2) : 0ull;
            t.bb[q] = se.x + head + lane;
        }
#pragma unroll
        for (int q = 0; q < kMemberBatch; ++q)
#pragma unroll
            for (int it = 0; it < kIters; ++it) {
                const uint32_t i = t.bb[q] + 32 * it;
                // (the zero is never used, but without it ptxas schedules the batch 8 % slower)
                t.r[q][it] = i < t.ee[q] ? stream_load(ent + i) : 0ull;
            }
    };
    auto process = [&](const Batch &t) {
#pragma unroll
        for (int q = 0; q < kMemberBatch; ++q) {
            uint32_t slot[kIters];
            unsigned long long add[kIters], cur[kIters];
            bool ok[kIters];
#pragma unroll
            for (int it = 0; it < kIters; ++it) {
                ok[it] = t.bb[q] + 32 * it < t.ee[q];
                decode_entry(t.r[q][it], slot[it], add[it]);
            }
            // the genes of one cell are distinct: the loads of all steps may precede the stores
#pragma unroll
            for (int it = 0; it < kIters; ++it)
                if (ok[it]) cur[it] = acc[slot[it]];
#pragma unroll
            for (int it = 0; it < kIters; ++it)
                if (ok[it]) acc[slot[it]] = cur[it] + add[it];
            for (uint32_t i = t.bb[q] + 32 * kIters; i < t.ee[q]; i += 32) {     // long segments
                uint32_t sl;
                unsigned long long ad;
                decode_entry(stream_load(ent + i), sl, ad);
                acc[sl] += ad;
            }
#pragma unroll
            for (int w2 = 0; w2 < kDenseRegs; ++w2) dacc[w2] += t.dv[q][w2];
            __syncwarp();          // next cell may hit the same genes from other lanes
        }
    };
    const int nb = (cntm + kMemberBatch - 1) / kMemberBatch;
    if (nb > 0) {
        fetch_meta(0, mb[0], me[0]);
        Batch A;
        for (int ib = 0; ib < nb; ++ib) {
            load(A, ib);
            process(A);
        }
    }
    collect_accumulators(a, acc, dacc, dense, lane, p, g0);
    if (!kSplit) {
        if (coop > 1) {
            // the parts meet in the accumulators of the group's first warp (named barrier per group)
            const int group = warp / coop, gthreads = coop * 32, gtid = part * 32 + lane;
            unsigned long long *acc0 = smem_acc + (size_t)(group * coop) * a.acc_len;
            asm volatile("bar.sync %0, %1;" ::"r"(1 + group), "r"(gthreads) : "memory");
            for (int j = gtid; j < ng; j += gthreads) {
                unsigned long long v = acc0[j];
                for (int w = 1; w < coop; ++w) v += acc0[(size_t)w * a.acc_len + j];
                acc0[j] = v;
            }
            asm volatile("bar.sync %0, %1;" ::"r"(1 + group), "r"(gthreads) : "memory");
            finish_task<kRawOnly, coop * 32>(a, acc0, gtid, p, g0, ng, row, N);
        } else {
            finish_task<kRawOnly>(a, acc, lane, p, g0, ng, row, N);
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

// Member cells of every set in ascending order (copy): one CTA per set, bitonic sort in shared
// memory, sets of up to kSortMax cells.  Order inside a set does not change any sum; sorted
// walks make the warps that work on one gene partition sweep the cells in step (L2 reuse).
constexpr int kSortMax = 8192;
constexpr int kSortThreads = 256;

// A duplicated cell index inside the set sits next to its twin after the sort: flagged here (err).
__global__ void __launch_bounds__(kSortThreads)
k_sort_members(const int32_t *__restrict__ members, const int64_t *__restrict__ offsets,
               int32_t *__restrict__ sorted, int *__restrict__ err)
{
    extern __shared__ int32_t keys[];
    const int64_t beg = offsets[blockIdx.x];
    const int n = (int)(offsets[blockIdx.x + 1] - beg);
    if (n <= 0) return;
    int m = 32;
    while (m < n) m <<= 1;
    for (int i = threadIdx.x; i < m; i += blockDim.x) keys[i] = i < n ? members[beg + i] : 0x7fffffff;
    __syncthreads();
    for (int k = 2; k <= m; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < m; i += blockDim.x) {
                const int l = i ^ j;
                if (l > i) {
                    const int32_t a = keys[i], b = keys[l];
                    if ((a > b) == ((i & k) == 0)) { keys[i] = b; keys[l] = a; }
                }
            }
            __syncthreads();
        }
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        sorted[beg + i] = keys[i];
        if (i + 1 < n && keys[i] == keys[i + 1]) flag_error(err, 4);
    }
}

// The same check after CUB's segmented sort (sets above kSortMax cells).
__global__ void k_check_duplicates(const int32_t *__restrict__ sorted, const int64_t *__restrict__ offsets,
                                   int64_t S, int64_t total, int *__restrict__ err)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i + 1 < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        if (sorted[i] != sorted[i + 1]) continue;
        // same set?  i + 1 must not be the first member of a set
        int64_t lo = 0, hi = S;                      // first s with offsets[s] > i
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (offsets[mid] > i) hi = mid; else lo = mid + 1;
        }
        if (lo > S || offsets[lo] != i + 1) flag_error(err, 4);
    }
}

// Finish tests from raw accumulator words (k_de_fused raw form), e.g. words another GPU stored
// into this GPU's memory.  sizes[row] = N of the row's cell set (sizes null: offsets[row + 1] -
// offsets[row], the offsets of the batch the words came from).  A thread owns one gene and
// walks kExpandRows rows: the gene's metadata is read once, raw reads and table writes are
// coalesced across the genes of a warp.
#ifndef SPADE_EXPAND_ROWS
#define SPADE_EXPAND_ROWS 8
#endif
#ifndef SPADE_EXPAND_MIN_CTAS
#define SPADE_EXPAND_MIN_CTAS 4
#endif
constexpr int kExpandRows = SPADE_EXPAND_ROWS;

// (Measured per 33 M tests x 4 tables: 16 rows / 80 registers / 3 CTAs per SM 0.30 ms; 8 rows / 64
// registers / 4 CTAs 0.254 ms = 5.2 TB/s; the rare direct evaluation as a non-inlined call 0.29 ms.)

// Gene block [g_begin, g_begin + g_count): column j of raw and of the tables = gene g_begin + j.
__global__ void __launch_bounds__(256, SPADE_EXPAND_MIN_CTAS)
k_expand(const FinCtx f, const unsigned long long *__restrict__ raw, int64_t raw_ld,
         const int64_t *__restrict__ sizes, const int64_t *__restrict__ offsets, int64_t rows, int64_t g_begin,
         int64_t g_count)
{
    const int64_t col = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (col >= g_count) return;
    const int64_t g = g_begin + col;
    const int64_t r0 = (int64_t)blockIdx.y * kExpandRows;
    const int64_t r1 = min(rows, r0 + kExpandRows);
    const GeneMeta mt = f.meta[g];
    TabMeta tm;
    tm.off = 0; tm.lo = 0; tm.len = 0;
    if (f.tmeta) tm = f.tmeta[g];
    unsigned long long w[kExpandRows];
#pragma unroll
    for (int i = 0; i < kExpandRows; ++i)
        w[i] = r0 + i < r1 ? __ldcs(raw + (r0 + i) * raw_ld + col) : 0ull;
#pragma unroll
    for (int i = 0; i < kExpandRows; ++i) {
        const int64_t row = r0 + i;
        if (row >= r1) break;
        const long long v = (long long)w[i];
        const long long sumfx = (v << 16) >> 16;
        const int cnt = (int)((unsigned long long)(v - sumfx) >> kCountShift);
        const long long N = sizes ? sizes[row] : offsets[row + 1] - offsets[row];
        const int kk = mt.negmode ? cnt : (int)(N - cnt);
        const double mean_set = fixed_mean(sumfx, mt.shift, 1.0 / (double)N);
        const int64_t o = row * f.ld + col;
        if (f.cpm) __stcs(f.cpm + o, mean_set);
        if (f.fc) {
            const double mean_bg = f.nc_mode ? mt.bgmean
                                             : fixed_mean(mt.rowsum - sumfx, mt.shift, 1.0 / (double)((long long)f.M - N));
            __stcs(f.fc + o, fold_change(mean_set, mean_bg));
        }
        if (f.k) f.k[f.k_ld ? row * f.k_ld + col : o] = kk;
        if (f.up || f.down) {
            double lc, ls;
            if (kk == 0 || f.M == 0 || mt.n == 0 || N == 0) {
                lc = ls = __longlong_as_double(0x7ff8000000000000ll);
            } else {
                const unsigned d = (unsigned)(kk - tm.lo);
                if (d < (unsigned)tm.len) {
                    const double2 t = f.tab[tm.off + d];
                    lc = t.x;
                    ls = t.y;
                } else {
                    hyper_tails(kk, f.M, mt.n, (int)N, f.lf, lc, ls);
                }
            }
            if (f.up) __stcs(f.up + o, lc);
            if (f.down) __stcs(f.down + o, ls);
        }
    }
}

}  // namespace spade
