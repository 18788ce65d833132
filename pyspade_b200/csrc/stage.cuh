// One-time staging kernels: K0 CPM normalisation, the gene-partitioned cell-major layout, and
// K1 per-gene cutoffs / flags / fixed-point sums.  Reference lines: pySpade/utils.py:150-155
// (cpm_normalization), :199-200 and :253-254 (median cutoff and low set).
#pragma once
#include "common.cuh"

namespace spade {

constexpr int kStageThreads = 256;

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
        if (c < 0 || c >= C) { flag_error(err, 1); continue; }
        atomicAdd(&colsum[c], (unsigned long long)(long long)counts[i]);
    }
}

__global__ void k_cpm(const int32_t *__restrict__ indices, const int32_t *__restrict__ counts,
                      int64_t nnz, int64_t C, const unsigned long long *__restrict__ colsum,
                      double *__restrict__ values)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nnz;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int32_t c = indices[i];
        if (c < 0 || c >= C) continue;                                    // flagged by k_col_sums
        double inv = __drcp_rn((double)(long long)colsum[c]);             // fl(1/colsum)
        values[i] = __dmul_rn(__dmul_rn((double)counts[i], 1e6), inv);    // no FMA contraction
    }
}

// ------------------------------------------------------------------------------------------
// Layout: genes are cut into partitions of Gp consecutive genes; the stored values of
// partition p are kept cell-major, segment (p, c) = ent[pptr[p*C + c] .. pptr[p*C + c + 1]).
// One CTA per gene counts its entries per (partition, cell); an exclusive scan gives pptr.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kStageThreads)
k_part_hist(const int64_t *__restrict__ indptr, const int32_t *__restrict__ indices, int64_t C,
            int Gp, const int8_t *__restrict__ dslot, uint32_t *__restrict__ hist,
            int *__restrict__ err)
{
    const int64_t g = blockIdx.x;
    const int64_t base = (g / Gp) * C;
    const bool dense = dslot[g] >= 0;             // lives in the dense block, not in a segment
    for (int64_t i = indptr[g] + threadIdx.x; i < indptr[g + 1]; i += blockDim.x) {
        const int32_t c = indices[i];
        if (c < 0 || c >= C) { flag_error(err, 1); continue; }
        if (!dense) atomicAdd(&hist[base + c], 1u);
    }
}

// Segments start on 128-byte boundaries (16 entries): a warp's 256-byte read of 32 entries
// then touches two lines instead of three.  A partition with a dense block keeps it at the
// head of every segment: kDenseLanes words, word l = the partition's l-th densest gene in this
// cell (zero = not expressed), then the packed entries of the other genes.
// padded[i] = (dense head + count[i]) rounded up to 16.
// `spare` extra entries per non-empty segment (the two-choice layout may need one more row).
__global__ void k_pad_counts(const uint32_t *__restrict__ count, int64_t n, int64_t C,
                             const uint8_t *__restrict__ dense_on, uint32_t spare, uint32_t *__restrict__ padded)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t head = dense_on[i / C] ? kDenseLanes : 0;
        padded[i] = ((head + count[i] + (kSegAlign - 1)) & ~(uint32_t)(kSegAlign - 1)) + (count[i] ? spare : 0);
    }
}

// seg[i] = (first word of the segment, one past its last packed entry)
__global__ void k_make_segments(const uint32_t *__restrict__ begin, const uint32_t *__restrict__ count,
                                int64_t n, int64_t C, const uint8_t *__restrict__ dense_on,
                                uint2 *__restrict__ seg)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t head = dense_on[i / C] ? kDenseLanes : 0;
        seg[i] = make_uint2(begin[i], begin[i] + head + count[i]);
    }
}

// cursor[i] = first packed entry of the segment (after the dense head)
__global__ void k_init_cursor(const uint32_t *__restrict__ begin, int64_t n, int64_t C,
                              const uint8_t *__restrict__ dense_on, uint32_t *__restrict__ cursor)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x)
        cursor[i] = begin[i] + (dense_on[i / C] ? kDenseLanes : 0);
}

// ------------------------------------------------------------------------------------------
// K1: per-gene cutoff (utils.py:199-200, NC variant :253-254), flags, fixed-point sums.
// One CTA per gene.  The cutoff is np.median over the scope (all C cells, or the NC cells),
// zeros included: order statistics of the multiset {stored values in scope} + {implicit
// zeros}, found with an 8-bit-per-pass radix select on order-preserving keys of the float64
// values.  The same CTA then writes the gene's packed entries into the partitioned layout.
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

struct StageArgs {
    const int64_t *indptr;
    const int32_t *indices;
    const double *values;
    const uint8_t *ncmask;    // null -> scope = all cells
    int64_t C;
    int64_t scope_len;        // C or |NC|
    int Gp;
    uint32_t *cursor;         // running write position of every (partition, cell) segment
    unsigned long long *ent;
    const int8_t *dslot;      // per gene: lane of the dense block, or -1
    const uint32_t *begin;    // first word of every (partition, cell) segment
    double *median;
    int32_t *n;
    GeneMeta *meta;
    uint8_t *wide_flag;       // per gene: 1 = float64 side path (see the fixed-point comment below)
    int *err;
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
    __shared__ double scratch_d[33];

    const int64_t g = blockIdx.x;
    const int64_t beg = a.indptr[g], end = a.indptr[g + 1];

    // pass 0: census of the scope, magnitude of the row
    long long in_scope = 0, neg = 0, pos = 0;
    double rowabs = 0.0, maxabs = 0.0, minabs = INFINITY, bgd = 0.0;
    bool bad = false;
    for (int64_t i = beg + threadIdx.x; i < end; i += blockDim.x) {
        const double v = a.values[i];
        const double av = fabs(v);
        bad |= !(av < 8.0e270);                   // nan, inf, or too large for the fixed point
        rowabs += av;
        maxabs = fmax(maxabs, av);
        if (av > 0.0) minabs = fmin(minabs, av);
        if (a.ncmask && !a.ncmask[a.indices[i]]) continue;
        ++in_scope;
        neg += v < 0.0;
        pos += v > 0.0;
        bgd += v;                                 // float64 NC sum (used by wide genes only)
    }
    if (bad) flag_error(a.err, 3);
    in_scope = block_sum(in_scope, scratch_ll);
    neg = block_sum(neg, scratch_ll);
    pos = block_sum(pos, scratch_ll);
    rowabs = block_sum(rowabs, scratch_d);
    maxabs = block_max(maxabs, scratch_d);
    minabs = block_min(minabs, scratch_d);
    bgd = block_sum(bgd, scratch_d);
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

    // Fixed point of this gene: any sum over <= kMemberBlock member cells stays below 2^46.
    //   bound = min(sum |v|, kMemberBlock * max |v|) < 2^e   ->   shift = 45 - e
    // Every stored value is rounded to a multiple of 2^-shift, so a member mean is off by at most
    // 2^-(shift+1); relative to a mean of same-signed values that is at most
    // 2^-(shift+1) / min|nonzero value|.  Where this bound exceeds kWideRelBound (a few huge cells
    // next to tiny ones) the gene is marked WIDE: its member and background means are summed in
    // float64 by k_wide_means instead (the counts k stay with the packed flags, they are exact).
    int shift = 0;
    bool wide = false;
    {
        const double bound = fmin(rowabs, (double)kMemberBlock * maxabs);
        if (bound > 0.0 && !bad) {
            int e;
            frexp(bound, &e);
            shift = min(900, max(-900, 45 - e));
            wide = !(scalbn(minabs, shift + 1) * kWideRelBound >= 1.0);
        }
    }

    const int64_t part = g / a.Gp;
    const unsigned long long slot = (unsigned long long)(g - part * a.Gp) << kFieldBits;
    uint32_t *cursor = a.cursor + part * a.C;
    const int ds = a.dslot[g];
    long long n_low = 0, rowsum = 0, bgsum = 0;
    for (int64_t i = beg + threadIdx.x; i < end; i += blockDim.x) {
        const double v = a.values[i];
        const int32_t c = a.indices[i];
        const long long fx = bad ? 0ll : __double2ll_rn(scalbn(v, shift));
        const bool low = v <= m;
        n_low += low;
        rowsum += fx;
        if (a.ncmask && a.ncmask[c]) bgsum += fx;
        const long long flag = negmode ? (low ? 1 : 0) : (low ? 0 : 1);
        const unsigned long long field =
            (unsigned long long)((flag << kCountShift) + fx) & ((1ull << kFieldBits) - 1);
        if (ds >= 0) {
            // dense gene: the sign-extended field itself, one 64-bit word per (cell, lane)
            a.ent[a.begin[part * a.C + c] + ds] = (unsigned long long)((flag << kCountShift) + fx);
        } else {
            const uint32_t at = atomicAdd(&cursor[c], 1u);
            a.ent[at] = slot | field;
        }
    }
    n_low = block_sum(n_low, scratch_ll);
    rowsum = block_sum(rowsum, scratch_ll);
    bgsum = block_sum(bgsum, scratch_ll);
    if (threadIdx.x == 0) {
        const int32_t n = (int32_t)(n_low + ((0.0 <= m) ? (a.C - (end - beg)) : 0));
        a.median[g] = m;
        a.n[g] = n;
        GeneMeta mt;
        mt.n = n;
        mt.shift = (int16_t)shift;
        mt.negmode = negmode ? 1 : 0;
        mt.wide = wide ? 1 : 0;
        if (a.ncmask) mt.bgmean = wide ? bgd * (1.0 / (double)L)
                                       : ((double)bgsum * pow2_double(-shift)) * (1.0 / (double)L);
        else mt.rowsum = rowsum;
        a.meta[g] = mt;
        a.wide_flag[g] = wide ? 1 : 0;
    }
}

// ------------------------------------------------------------------------------------------
// Bank-aware order inside every (partition, cell) segment.  The test kernel reads a segment
// 32 entries per step and each lane does a 64-bit read-modify-write of its gene's
// accumulator: the hardware serves that as two half-warps of 16 lanes, and a half-warp costs
// as many shared-memory wavefronts as its most-requested 8-byte bank pair (accumulator index
// mod 16) is asked for.  Entry order inside a segment is free (integer adds commute), so the
// segment is composed as rows of 16 entries (one row = one half-warp of one step):
//
//   1. every bank pair c with l_c entries puts one entry into each of the rows 0 .. l_c - 1
//      (largest l_c first, rows that are full are skipped): a row then asks for each pair once;
//   2. what did not fit (pairs with more entries than there are rows) fills the free places of
//      the later rows, at most one more entry per pair and row at a time.
//
//   The segment stays compact (no holes) and costs about max_c l_c wavefronts per access (the
//   minimum for the given l_c; a partial last row adds to it).  Simulated on bench segments (59
//   entries, 4.2 rows): 8.15 wavefronts at max_c l_c = 7.4 with one accumulator per gene, 5.7 at
//   4.9 with the hot genes below.
//
// Hot genes (SPADE_TWO_CHOICE = H > 0, common.cuh): the H genes of a partition with the most
// stored values after the dense head own a second accumulator in another bank pair; per
// segment each of their entries is given to A or B so that max_c l_c comes down (greedy on the
// emptier pair, then moves out of overfull pairs) before the rows are composed.
// One warp per segment, chunks of kOrderChunk entries staged in shared memory; lane 0 does the
// serial part (a segment holds ~50 entries).
// ------------------------------------------------------------------------------------------
constexpr int kOrderChunk = 512;
constexpr int kOrderWarps = 5;       // 8.4 KB of scratch per warp

struct OrderScratch {
    unsigned long long buf[kOrderChunk];   // the chunk's entries
    uint16_t idx[kOrderChunk];             // accumulator index an entry adds into (A or B)
    uint16_t alt[kOrderChunk];             // the other accumulator of a hot gene's entry, 0xffff = none
    uint16_t by_pair[kOrderChunk];         // entries grouped by bank pair
    uint16_t pos[kOrderChunk];             // final place of an entry inside the chunk
    int cnt[16], start[16], fill[16];
    uint8_t start2[16];                    // place of a bank pair in the order by decreasing count
    uint8_t occ[kOrderChunk / 16];
};

// entries [beg, end) of one segment, chunk by chunk.  hotb: per gene of the partition (index =
// slot) the B accumulator of a hot gene (Gp + hotb[slot]) or a negative value; null = no hot genes.
__device__ void order_segment(unsigned long long *__restrict__ ent, uint32_t beg, uint32_t end, int Gp,
                              const int16_t *__restrict__ hotb, OrderScratch &w, int lane)
{
    for (uint32_t c0 = beg; c0 < end; c0 += kOrderChunk) {
        const int m = (int)min((uint32_t)kOrderChunk, end - c0);
        if (m <= 16 && !hotb) break;                         // a single row of distinct genes: nothing to gain
        const int rows = (m + 15) >> 4;
        for (int i = lane; i < m; i += 32) {
            const unsigned long long e = ent[c0 + i];
            const int sl = (int)(e >> kFieldBits);
            w.buf[i] = e;
            w.idx[i] = (uint16_t)sl;
            const int hb = hotb ? hotb[sl] : -1;
            w.alt[i] = hb >= 0 ? (uint16_t)(Gp + hb) : (uint16_t)0xffff;
        }
        if (lane < 16) w.cnt[lane] = 0;
        __syncwarp();
        if (hotb && lane == 0) {
            // A or B for the entries of hot genes: fixed entries first, then greedy, then moves
            const int L = rows;
            for (int i = 0; i < m; ++i)
                if (w.alt[i] == 0xffff) ++w.cnt[w.idx[i] & 15];
            for (int i = 0; i < m; ++i) {
                if (w.alt[i] == 0xffff) continue;
                const int a = w.idx[i] & 15, o = w.alt[i] & 15;
                if (w.cnt[a] <= w.cnt[o]) { ++w.cnt[a]; }
                else { const uint16_t t = w.idx[i]; w.idx[i] = w.alt[i]; w.alt[i] = t; ++w.cnt[o]; }
            }
            for (int pass = 0; pass < 4; ++pass) {
                bool moved = false;
                for (int i = 0; i < m; ++i) {
                    if (w.alt[i] == 0xffff) continue;
                    const int cur = w.idx[i] & 15, o = w.alt[i] & 15;
                    if (w.cnt[cur] > L && w.cnt[o] + 1 < w.cnt[cur]) {
                        const uint16_t t = w.idx[i]; w.idx[i] = w.alt[i]; w.alt[i] = t;
                        --w.cnt[cur]; ++w.cnt[o];
                        moved = true;
                    }
                }
                if (!moved) break;
            }
        } else if (!hotb) {
            for (int i = lane; i < m; i += 32) atomicAdd(&w.cnt[w.idx[i] & 15], 1);
        }
        __syncwarp();
        {   // group the entries by bank pair: exclusive scan of the 16 counts, then a counting sort
            int v = lane < 16 ? w.cnt[lane] : 0, inc = v;
#pragma unroll
            for (int o = 1; o < 16; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += t;
            }
            if (lane < 16) { w.start[lane] = inc - v; w.fill[lane] = 0; }
        }
        __syncwarp();
        for (int i = lane; i < m; i += 32) {
            const int c = w.idx[i] & 15;
            const int k = atomicAdd(&w.fill[c], 1);          // rank of the entry inside its bank pair
            w.by_pair[w.start[c] + k] = (uint16_t)i;
            w.pos[i] = (uint16_t)k;
        }
        __syncwarp();
        // 1. one entry per row, in closed form.  With the pairs ordered by decreasing count (q = place of
        // a pair in that order), row r < rows - 1 holds one entry of every pair with more than r entries -
        // the first pairs of the order - so the k-th entry of pair q sits at 16 k + q.  The last row has
        // last_cap places: the first last_cap pairs that reach it get one, what is left goes to step 2.
        const int last_cap = m - 16 * (rows - 1);
        int q = 0;                                           // lanes 0..15: place of pair `lane` in the order
        if (lane < 16) {
            const int mine = w.cnt[lane];
            for (int c = 0; c < 16; ++c) {
                const int other = w.cnt[c];
                q += (other > mine || (other == mine && c < lane)) ? 1 : 0;
            }
            w.start2[lane] = (uint8_t)q;
            const int placed = min(mine, rows - 1) + ((mine >= rows && q < last_cap) ? 1 : 0);
            w.fill[lane] = placed;
        }
        {   // places taken in row `lane` after step 1
            int present = 0;
            if (lane < rows)
                for (int c = 0; c < 16; ++c) present += w.cnt[c] > lane ? 1 : 0;
            if (lane < rows) w.occ[lane] = (uint8_t)(lane == rows - 1 ? min(present, last_cap) : present);
        }
        __syncwarp();
        for (int i = lane; i < m; i += 32) {
            const int k = w.pos[i], qq = w.start2[w.idx[i] & 15];
            if (k < rows - 1 || (k == rows - 1 && qq < last_cap)) w.pos[i] = (uint16_t)(16 * k + qq);
        }
        __syncwarp();
        if (lane == 0) {
            // 2. the rest (pairs with more entries than rows, or that missed the last row) fills the free
            // places of the rows in order, one more entry per pair and row at a time
            int order[16];
            for (int c = 0; c < 16; ++c) order[w.start2[c]] = c;
            int r = 0;
            bool left = true;
            while (left) {
                left = false;
                for (int o = 0; o < 16; ++o) {
                    const int c = order[o];
                    if (w.fill[c] >= w.cnt[c]) continue;
                    while (r < rows && w.occ[r] >= (r == rows - 1 ? last_cap : 16)) ++r;
                    if (r >= rows) break;                    // cannot happen: places = entries
                    w.pos[w.by_pair[w.start[c] + w.fill[c]]] = (uint16_t)(16 * r + w.occ[r]);
                    ++w.occ[r];
                    ++w.fill[c];
                    left = true;
                }
            }
        }
        __syncwarp();
        for (int i = lane; i < m; i += 32)
            ent[c0 + w.pos[i]] = ((unsigned long long)w.idx[i] << kFieldBits) | (w.buf[i] & ((1ull << kFieldBits) - 1));
        __syncwarp();
    }
}

__global__ void __launch_bounds__(kOrderWarps * 32)
k_order_segments(const uint2 *__restrict__ segs, int64_t nseg, int64_t C, int Gp, int hot,
                 const int16_t *__restrict__ hotb, const uint8_t *__restrict__ dense_on,
                 unsigned long long *__restrict__ ent)
{
    __shared__ OrderScratch scratch[kOrderWarps];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t seg = blockIdx.x * (int64_t)kOrderWarps + warp;
    if (seg >= nseg) return;
    const int64_t part = seg / C;
    const uint32_t beg = segs[seg].x + (dense_on[part] ? kDenseLanes : 0), end = segs[seg].y;
    order_segment(ent, beg, end, Gp, hot > 0 ? hotb + part * Gp : nullptr, scratch[warp], lane);
}

}  // namespace spade
