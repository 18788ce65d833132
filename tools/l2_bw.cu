// Microbenchmark (not part of the product): random-chunk gather bandwidth as a function of the
// footprint - the access pattern of the fused test kernel (a warp reads a contiguous, 128-byte
// aligned chunk at a random position) with 8-byte or 16-byte loads per lane, everything in
// registers, BATCH chunks in flight per warp.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

template <int W, int LOADS, int BATCH>
__global__ void __launch_bounds__(128) gather(const char *__restrict__ buf, uint32_t nchunks, int iters,
                                              unsigned long long *out)
{
    const int lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    constexpr int chunk_bytes = 32 * W * LOADS;
    unsigned long long acc = 0;
    uint32_t s = warp * 2654435761u + 12345u;
    for (int it = 0; it < iters; ++it) {
        ulonglong2 v[BATCH][LOADS];
#pragma unroll
        for (int b = 0; b < BATCH; ++b) {
            s = s * 1664525u + 1013904223u;
            const uint32_t c = (uint32_t)(((unsigned long long)s * nchunks) >> 32);
            const char *p = buf + (size_t)c * chunk_bytes + lane * W;
#pragma unroll
            for (int k = 0; k < LOADS; ++k) {
                if (W == 16) v[b][k] = __ldcg((const ulonglong2 *)(p + 32 * W * k));
                else { v[b][k].x = __ldcg((const unsigned long long *)(p + 32 * W * k)); v[b][k].y = 0; }
            }
        }
#pragma unroll
        for (int b = 0; b < BATCH; ++b)
#pragma unroll
            for (int k = 0; k < LOADS; ++k) acc += v[b][k].x + v[b][k].y;
    }
    if (acc == 0x1234567ull) out[0] = acc;
}

template <int W, int LOADS, int BATCH>
void run(const char *buf, unsigned long long *out, cudaEvent_t e0, cudaEvent_t e1)
{
    constexpr int chunk = 32 * W * LOADS;
    for (int warps_per_sm : {24, 48})
        for (size_t mb : {32, 96, 2048}) {
            const uint32_t nchunks = (uint32_t)((mb << 20) / chunk);
            const int iters = 4096 / BATCH;
            const int blocks = 148 * warps_per_sm / 4;
            float ms = 0;
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord(e0);
                gather<W, LOADS, BATCH><<<blocks, 128>>>(buf, nchunks, iters, out);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                cudaEventElapsedTime(&ms, e0, e1);
            }
            const double gb = (double)blocks * 4 * iters * BATCH * chunk / 1e9;
            printf("load %2d B/lane x%d (chunk %4d B) batch %d warps/SM %2d footprint %5zu MB: %7.1f GB/s\n", W, LOADS,
                   chunk, BATCH, warps_per_sm, mb, gb / (ms * 1e-3));
        }
}

int main()
{
    const size_t maxbytes = (size_t)2 << 30;
    char *buf;
    unsigned long long *out;
    cudaMalloc(&buf, maxbytes);
    cudaMalloc(&out, 8);
    cudaMemset(buf, 1, maxbytes);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    run<8, 2, 4>(buf, out, e0, e1);
    run<8, 3, 4>(buf, out, e0, e1);
    run<16, 1, 4>(buf, out, e0, e1);
    run<16, 1, 8>(buf, out, e0, e1);
    run<16, 2, 4>(buf, out, e0, e1);
    run<8, 3, 1>(buf, out, e0, e1);
    return 0;
}
