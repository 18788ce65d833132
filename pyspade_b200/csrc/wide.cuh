// Float64 side path for WIDE genes: genes whose stored values span so many binary orders of
// magnitude that the per-gene fixed point of the main kernel cannot keep a member mean within
// kWideRelBound (stage.cuh).  perform_DE accepts any float64 matrix (pySpade/utils.py:296), and
// the reference sums in float64 (scipy sparse mean, utils.py:206-208 / :258-259), so these
// genes are summed in float64 here: one warp per (wide gene, set) walks the gene's CSR row and
// tests every cell against the set's membership bitmask.  The order of the additions is fixed
// by the CSR order and the lane count, so the result does not depend on launch geometry or
// sharding.  Count data normalised to CPM never produces a wide gene; this path exists for the
// drop-in contract, not for speed.
#pragma once
#include "common.cuh"

namespace spade {

// bits[s][w]: membership of cells 32 w .. 32 w + 31 in set s (zeroed by the caller)
__global__ void k_set_bitmask(const int32_t *__restrict__ members, const int64_t *__restrict__ offsets,
                              int64_t s0, int64_t S, int64_t C, int64_t words, uint32_t *__restrict__ bits)
{
    const int64_t s = blockIdx.x;
    if (s >= S) return;
    const int64_t beg = offsets[s0 + s], end = offsets[s0 + s + 1];
    uint32_t *row = bits + s * words;
    for (int64_t i = beg + threadIdx.x; i < end; i += blockDim.x) {
        const int32_t c = members[i];
        if ((uint32_t)c < (uint32_t)C) atomicOr(&row[c >> 5], 1u << (c & 31));   // range errors are flagged by the test kernel
    }
}

constexpr int kWideWarps = 8;

struct WideArgs {
    const int32_t *genes;        // wide gene ids of this launch
    int nw;
    const int64_t *indptr;
    const int32_t *indices;
    const double *values;
    const GeneMeta *meta;
    const uint32_t *bits;        // [S][words]
    int64_t words;
    const int64_t *offsets;
    int64_t s0, S, C;
    int nc_mode;
    double *fc, *cpm;            // row 0 = set s0; either may be null
    int64_t ld;
};

__global__ void __launch_bounds__(kWideWarps * 32) k_wide_means(const WideArgs a)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t s = (int64_t)blockIdx.y * kWideWarps + warp;
    if (s >= a.S) return;
    const int64_t g = a.genes[blockIdx.x];
    const int64_t beg = a.indptr[g], end = a.indptr[g + 1];
    const uint32_t *__restrict__ row = a.bits + s * a.words;
    double in = 0.0, out = 0.0;
    for (int64_t i = beg + lane; i < end; i += 32) {
        const int32_t c = a.indices[i];
        const double v = a.values[i];
        if ((row[c >> 5] >> (c & 31)) & 1u) in += v; else out += v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        in += __shfl_xor_sync(0xffffffffu, in, o);
        out += __shfl_xor_sync(0xffffffffu, out, o);
    }
    if (lane != 0) return;
    const long long N = a.offsets[a.s0 + s + 1] - a.offsets[a.s0 + s];
    const double mean_set = __dmul_rn(in, 1.0 / (double)N);
    const int64_t o = s * a.ld + g;
    if (a.cpm) a.cpm[o] = mean_set;
    if (a.fc) {
        const double mean_bg = a.nc_mode ? a.meta[g].bgmean : __dmul_rn(out, 1.0 / (double)((long long)a.C - N));
        a.fc[o] = __ddiv_rn(__dadd_rn(mean_set, 0.01), __dadd_rn(mean_bg, 0.01));
    }
}

}  // namespace spade
