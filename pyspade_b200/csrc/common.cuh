// Shared helpers of libspade_b200 (device utilities, device buffers).
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cstddef>
#include <cstdint>

namespace spade {

// Packed cell-major entry (one stored value of the genes x cells matrix), 64 bits:
//   bits 63..51  gene index inside its gene partition (13 bits, partitions hold <= 8192 genes)
//   bits 50..0   two's-complement 51-bit field  flag * 2^48 + fx
//                fx   = value in the gene's fixed point (round(value * 2^shift_g)), |sum| < 2^47
//                flag = 1 when the cell is on the counted side of the gene's cutoff
// Adding the sign-extended field into a 64-bit accumulator therefore carries the member count
// in bits 63..48 and the member sum in bits 47..0 with ONE integer add per stored value.
constexpr int kSlotBits = 13;
constexpr int kFieldBits = 64 - kSlotBits;            // 51
constexpr int kCountShift = 48;
constexpr int kMaxGenesPerPart = 1 << kSlotBits;      // 8192
// A warp accumulates at most this many member cells before its accumulators are drained:
// the 16-bit count field and the 47-bit sum budget (see stage.cuh) are sized for it.
constexpr int kMemberBlock = 32768;
// Guaranteed relative accuracy of a fixed-point member mean of same-signed values; genes that
// cannot meet it take the float64 side path (stage.cuh, wide.cuh).  Fold changes: twice this.
constexpr double kWideRelBound = 0x1p-22;             // 2.4e-7
// Hot genes (stage.cuh order_segment): the H genes of a partition with the most stored values
// after the dense head own TWO accumulators in different shared-memory bank pairs, A at index
// slot and B at index Gp + b (b < H; b mod 16 != slot mod 16); an entry carries the index of the
// one it adds into, chosen per segment at staging time so that the bank pairs of a segment are
// asked for as evenly as possible.  The test is finished from A + B.  H = Gp: every gene, B at
// Gp + alt_slot(slot).
__host__ __device__ __forceinline__ int alt_slot(int s)    // same 16-group, bank pair shifted by 1..15
{
    return (s & ~15) | ((s + 1 + ((s >> 4) % 15)) & 15);
}
constexpr int kSegAlign = 16;                         // entries: segments start on 128-byte lines
#ifndef SPADE_DENSE_REGS
#define SPADE_DENSE_REGS 1
#endif
constexpr int kDenseRegs = SPADE_DENSE_REGS;          // register accumulators per lane for the dense head
constexpr int kDenseLanes = 32 * kDenseRegs;          // genes per partition kept in the dense head

struct __align__(16) GeneMeta {      // one per gene, fixed by the matrix and the background
    int32_t n;                       // #{cells : x <= cutoff}             (utils.py:200,211)
    int16_t shift;                   // fixed point: fx = round(value * 2^shift)
    uint8_t negmode;                 // cutoff < 0: flag marks the LOW side, k = count
    uint8_t wide;                    // dynamic range too wide for the fixed point: the member
                                     // means of this gene come from the float64 side path (wide.cuh)
    union {
        long long rowsum;            // complement background: sum of fx over all cells
        double bgmean;               // NC background: mean over the NC cells (utils.py:258)
    };
};

struct __align__(16) TabMeta {       // one per gene, per run: window of the tail table
    unsigned long long off;          // first entry in the table
    int32_t lo;                      // k of that entry
    int32_t len;                     // entries; 0 = no table for this gene
};

// Device-side error word: lives in page-locked host memory mapped into the device, so the host
// can read it at any time without synchronising (1 bad matrix index, 2 member out of range,
// 3 non-finite matrix value, 4 duplicate member).  A plain store: any code is as good as another.
__device__ __forceinline__ void flag_error(int *err, int code)
{
    *(volatile int *)err = code;
}

__device__ __forceinline__ double pow2_double(int e)   // 2^e for -1022 <= e <= 1023
{
    return __longlong_as_double((long long)(1023 + e) << 52);
}

template <typename T>
__device__ __forceinline__ T warp_sum(T v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ double warp_max(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, o));
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

__device__ inline double block_max(double v, double *scratch)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
    v = warp_max(v);
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    if (warp == 0) {
        double t = lane < nwarp ? scratch[lane] : 0.0;
        t = warp_max(t);
        if (lane == 0) scratch[32] = t;
    }
    __syncthreads();
    return scratch[32];
}

// Minimum over the block (identity +inf); result valid in every thread.
__device__ inline double block_min(double v, double *scratch)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    if (warp == 0) {
        double t = lane < nwarp ? scratch[lane] : INFINITY;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t = fmin(t, __shfl_xor_sync(0xffffffffu, t, o));
        if (lane == 0) scratch[32] = t;
    }
    __syncthreads();
    return scratch[32];
}

template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t cap = 0;
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { release(); }                 // temporaries on error paths do not leak
    cudaError_t reserve(size_t count)
    {
        if (count <= cap && p) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMalloc((void **)&p, std::max<size_t>(count, 1) * sizeof(T));
        if (e == cudaSuccess) cap = std::max<size_t>(count, 1);
        return e;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

// Page-locked staging buffer for small host->device uploads issued on a stream: a copy from
// pageable memory would synchronise the stream first.  `wait()` before refilling.
template <typename T>
struct PinnedBuf {
    T *p = nullptr;
    size_t cap = 0;
    cudaEvent_t done = nullptr;
    bool pending = false;
    cudaError_t reserve(size_t count)
    {
        cudaError_t e = wait();
        if (e != cudaSuccess) return e;
        if (count <= cap && p) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        e = cudaHostAlloc((void **)&p, std::max<size_t>(count, 1) * sizeof(T), cudaHostAllocDefault);
        if (e == cudaSuccess) cap = std::max<size_t>(count, 1);
        if (e == cudaSuccess && !done) e = cudaEventCreateWithFlags(&done, cudaEventDisableTiming);
        return e;
    }
    cudaError_t wait()
    {
        if (!pending) return cudaSuccess;
        pending = false;
        return cudaEventSynchronize(done);
    }
    cudaError_t upload(T *dst, size_t count, cudaStream_t st)
    {
        cudaError_t e = cudaMemcpyAsync(dst, p, count * sizeof(T), cudaMemcpyHostToDevice, st);
        if (e != cudaSuccess) return e;
        pending = true;
        return cudaEventRecord(done, st);
    }
    void release()
    {
        if (pending) cudaEventSynchronize(done);
        pending = false;
        if (p) cudaFreeHost(p);
        if (done) cudaEventDestroy(done);
        p = nullptr;
        done = nullptr;
        cap = 0;
    }
};

inline int grid_for(int64_t n, int threads, int cap = 148 * 32)
{
    int64_t b = (n + threads - 1) / threads;
    return (int)std::max<int64_t>(1, std::min<int64_t>(b, cap));
}

}  // namespace spade
