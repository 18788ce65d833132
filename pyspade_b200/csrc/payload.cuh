// Column-compressed payload of a DErand result table, built on the GPU (row f3: the wire format
// of DErand.py:230-248).  The reference saves every table as scipy.io.savemat of
// csr_matrix(dense): a MAT-v5 sparse array, i.e. column-compressed storage with column = gene -
// jc[G+1] column pointers, ir[nnz] row (draw) indices ascending inside a column, pr[nnz] values;
// stored entries are those != 0 (nan and -inf are stored, +-0.0 are not).  One warp per gene
// walks the gene's column of the [S][ld] table (all draws of a gene: exactly what a rank holds
// after the gene-sharded exchange), counts, and in a second pass writes the entries in order.
#pragma once
#include "common.cuh"

namespace spade {

__global__ void __launch_bounds__(256)
k_csc_count(const double *__restrict__ table, int64_t S, int64_t G, int64_t ld, int32_t *__restrict__ count)
{
    const int64_t g = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (g >= G) return;
    int c = 0;
    for (int64_t s = lane; s < S; s += 32) c += table[s * ld + g] != 0.0;      // nan != 0: stored
    c = warp_sum(c);
    if (lane == 0) count[g] = c;
}

__global__ void __launch_bounds__(256)
k_csc_fill(const double *__restrict__ table, int64_t S, int64_t G, int64_t ld,
           const int64_t *__restrict__ begin, int32_t *__restrict__ ir, double *__restrict__ pr)
{
    const int64_t g = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (g >= G) return;
    int64_t at = begin[g];
    for (int64_t s0 = 0; s0 < S; s0 += 32) {
        const int64_t s = s0 + lane;
        const double v = s < S ? table[s * ld + g] : 0.0;
        const bool keep = s < S && v != 0.0;
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (keep) {
            const int64_t o = at + __popc(m & ((1u << lane) - 1));
            ir[o] = (int32_t)s;
            pr[o] = v;
        }
        at += __popc(m);
    }
}

// Mean over the rows of every column (Cpm_mean, DErand.py:252): one warp per gene, lane-strided
// partial sums and a fixed shuffle tree, so the value does not depend on the launch geometry.
__global__ void __launch_bounds__(256)
k_column_mean(const double *__restrict__ table, int64_t S, int64_t G, int64_t ld, double *__restrict__ mean)
{
    const int64_t g = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (g >= G) return;
    double t = 0.0;
    for (int64_t s = lane; s < S; s += 32) t += table[s * ld + g];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane == 0) mean[g] = t / (double)S;
}

}  // namespace spade
