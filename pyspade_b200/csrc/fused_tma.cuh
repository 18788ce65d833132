// EXPERIMENT (SPADE_TMA=1 / 2, off by default: measured slower).
// The batched test with segments staged through shared memory (sm_100a): same task, layout and
// results as k_de_fused<false> (fused.cuh), different data movement.  k_de_fused pulls every packed
// entry through the LSU into registers (ld.global, 8 bytes per lane).  Here a warp keeps a private
// ring of kTmaStages x kStageEnt entries in shared memory and fills it asynchronously, either with
// ONE cp.async.bulk per (partition, cell) segment issued by an elected lane to the TMA unit
// (completion counted in bytes on an mbarrier; k_de_tma<true>) or with 16-byte cp.async copies by
// all lanes (one commit group per segment; k_de_tma<false>).  The lanes then read the entries from
// shared memory (conflict-free 64-bit reads) and add them into the warp's accumulators as before.
// The idea: asynchronous copies cover the latency of L2 with the ring instead of with resident
// warps, which would make a second accumulator for EVERY gene (16 KB per warp) affordable.
// Measured per 1 000 draws of 500 cells: TMA 3.55 ms (the ~650-byte copies are bound by the TMA
// unit's per-copy service time), cp.async 2.96 ms, k_de_fused 2.01 ms (profiles/r02_tma_bulk_sweep.log,
// profiles/r02_ldgsts_ring_sweep.log).
//
// Per warp, in shared memory: accumulators | ring | 32 segment bounds of the current member group |
// kTmaStages mbarriers | kTmaStages chunk descriptors.  A segment longer than a stage is taken in
// chunks of kStageEnt entries.
#pragma once
#include "fused.cuh"

namespace spade {

#ifndef SPADE_TMA_STAGES
#define SPADE_TMA_STAGES 4
#endif
#ifndef SPADE_TMA_WARPS
#define SPADE_TMA_WARPS 4
#endif
#ifndef SPADE_TMA_STAGE_ENT
#define SPADE_TMA_STAGE_ENT 128
#endif
#ifndef SPADE_TMA_MIN_CTAS
#define SPADE_TMA_MIN_CTAS 1
#endif
constexpr int kTmaStages = SPADE_TMA_STAGES;
constexpr int kTmaWarps = SPADE_TMA_WARPS;
constexpr int kStageEnt = SPADE_TMA_STAGE_ENT;        // entries per stage
constexpr int kStageIters = kStageEnt / 32;
static_assert(kStageEnt % 64 == 0 && kStageEnt >= kDenseLanes + 32, "a stage holds the dense head and one more step");
static_assert(kStageEnt <= 0xffff, "chunk length is kept in 16 bits");
static_assert(kTmaStages >= 2, "the descriptor of a refilled stage is written while another one is read");

// 64-bit words of shared memory per warp (a multiple of 16 words: every region stays 128-byte aligned)
__host__ __device__ constexpr int tma_warp_words(int acc_len)
{
    return (acc_len + kTmaStages * kStageEnt + 32 + 2 * kTmaStages + 15) / 16 * 16;
}

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

// global -> shared bulk copy by the TMA unit; `bytes` a multiple of 16, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok)
                     : "r"(bar), "r"(parity)
                     : "memory");
    } while (!ok);
}

// 16-byte asynchronous copy global -> shared (LDGSTS, L2 only), grouped per thread
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int kPending>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(kPending) : "memory"); }

// kBulk: one cp.async.bulk per chunk by the TMA unit (completion on an mbarrier); otherwise every
// lane copies 16-byte pieces of the chunk with cp.async (LDGSTS), one commit group per chunk.
template <bool kBulk>
__global__ void __launch_bounds__(kTmaWarps * 32, SPADE_TMA_MIN_CTAS) k_de_tma(const RunArgs a)
{
    extern __shared__ __align__(128) unsigned long long smem_acc[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long *acc = smem_acc + (size_t)warp * tma_warp_words(a.acc_len);
    unsigned long long *ring = acc + a.acc_len;
    uint2 *mbuf = reinterpret_cast<uint2 *>(ring + kTmaStages * kStageEnt);
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(mbuf + 32);
    uint32_t *sinfo = reinterpret_cast<uint32_t *>(bars + kTmaStages);

    const int64_t t = blockIdx.x * (int64_t)kTmaWarps + warp;
    const int64_t pi = t / a.S;
    if (pi >= a.p1 - a.p0) return;
    const int p = a.p0 + (int)pi;
    const int64_t u = t - pi * a.S;
    const int64_t row = a.order ? a.order[u] : u;
    const int64_t mbeg = a.offsets[a.s0 + row];
    const long long N = a.offsets[a.s0 + row + 1] - mbeg;
    const int cntm = (int)N;
    const int64_t g0 = (int64_t)p * a.Gp;
    const int ng = (int)min((int64_t)a.Gp, a.fin.G - g0);

    if (kBulk && lane < kTmaStages) mbar_init(smem_addr(bars + lane), 1);
    for (int j = lane; j < a.acc_len; j += 32) acc[j] = 0ull;
    if (kBulk) {
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();

    const uint2 *__restrict__ seg = a.seg + (int64_t)p * a.C;
    const int32_t *__restrict__ mem = a.members + mbeg;
    const unsigned long long *__restrict__ ent = a.ent;
    const bool dense = a.dense_on[p] != 0;
    unsigned long long dacc[kDenseRegs];
#pragma unroll
    for (int r = 0; r < kDenseRegs; ++r) dacc[r] = 0ull;
    const uint32_t bar0 = smem_addr(bars), ring0 = smem_addr(ring);

    // segment bounds of 32 members at a time, fetched one group ahead (two register sets by parity)
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

    // producer side (every lane keeps the state, lane 0 talks to the TMA unit)
    int pm = 0;                      // next member whose segment has not been started
    uint32_t pcur = 0, pend = 0;     // what is left of the segment being issued
    bool pfirst = false;             // the next chunk is the first of its segment (holds the dense head)
    int issued = 0, istage = 0;      // chunks issued so far; stage of the next one
    auto issue_one = [&]() -> bool {
        while (pcur >= pend) {
            if (pm >= cntm) return false;
            if ((pm & 31) == 0) {    // entering a member group: park its bounds, prefetch the next group's
                const int grp = pm >> 5;
                __syncwarp();
                mbuf[lane] = (grp & 1) ? make_uint2(mb[1], me[1]) : make_uint2(mb[0], me[0]);
                __syncwarp();
                if (grp & 1) fetch_meta(grp + 1, mb[0], me[0]);
                else fetch_meta(grp + 1, mb[1], me[1]);
            }
            const uint2 se = mbuf[pm & 31];
            pcur = se.x;
            pend = se.y;
            pfirst = true;
            ++pm;
        }
        const uint32_t n = min((uint32_t)kStageEnt, pend - pcur);
        if (lane == 0) sinfo[istage] = n | (pfirst ? 0x10000u : 0u);
        if (kBulk) {
            if (lane == 0) {
                const uint32_t bytes = ((n + 1) & ~1u) * 8;  // segments are padded to 128-byte lines
                mbar_arrive_expect_tx(bar0 + 8 * istage, bytes);
                bulk_copy_g2s(ring0 + istage * (kStageEnt * 8), ent + pcur, bytes, bar0 + 8 * istage);
            }
        } else {
            const uint32_t pieces = (n + 1) >> 1;            // 16-byte pieces
#pragma unroll
            for (int q = 0; q < kStageEnt / 64; ++q) {
                const uint32_t i = lane + 32 * q;
                if (i < pieces) cp_async16(ring0 + istage * (kStageEnt * 8) + 16 * i, ent + pcur + 2 * i);
            }
        }
        pcur += n;
        pfirst = false;
        ++issued;
        istage = istage + 1 == kTmaStages ? 0 : istage + 1;
        return true;
    };

    // consumer side
    int done = 0, cstage = 0;
    uint32_t cparity = 0;
    auto consume = [&]() {
        if (kBulk) {
            mbar_wait(bar0 + 8 * cstage, cparity);
        } else {
            cp_async_wait<kTmaStages - 1>();     // the oldest of the kTmaStages groups has landed ...
            __syncwarp();                        // ... in every lane
        }
        const uint32_t info = sinfo[cstage];
        const int n = (int)(info & 0xffffu);
        const unsigned long long *sb = ring + cstage * kStageEnt;
        int off = lane;
        if (dense && (info >> 16)) {
#pragma unroll
            for (int w2 = 0; w2 < kDenseRegs; ++w2) dacc[w2] += sb[lane + 32 * w2];
            off += kDenseLanes;
        }
        uint32_t slot[kStageIters];
        unsigned long long add[kStageIters], cur[kStageIters];
        bool ok[kStageIters];
#pragma unroll
        for (int it = 0; it < kStageIters; ++it) {
            ok[it] = off + 32 * it < n;
            const unsigned long long e = ok[it] ? sb[off + 32 * it] : 0ull;
            decode_entry(e, slot[it], add[it]);
        }
        // the genes of one cell are distinct: the loads of all steps may precede the stores
#pragma unroll
        for (int it = 0; it < kStageIters; ++it)
            if (ok[it]) cur[it] = acc[slot[it]];
#pragma unroll
        for (int it = 0; it < kStageIters; ++it)
            if (ok[it]) acc[slot[it]] = cur[it] + add[it];
        __syncwarp();            // the next cell may hit the same genes from other lanes; the stage is free
        ++done;
        if (++cstage == kTmaStages) { cstage = 0; cparity ^= 1u; }
    };

    if (cntm > 0) {
        fetch_meta(0, mb[0], me[0]);
        // (LDGSTS form: exactly one commit group per stage slot and per loop turn, empty when there
        // is nothing left to issue, so the oldest of kTmaStages groups is always the chunk to consume)
        for (int s = 0; s < kTmaStages; ++s) {
            issue_one();
            if (!kBulk) cp_async_commit();
        }
        __syncwarp();
        while (done < issued) {
            consume();
            issue_one();         // refills the stage just consumed
            if (!kBulk) cp_async_commit();
        }
    }
    collect_accumulators(a, acc, dacc, dense, lane, p, g0);
    finish_task(a, acc, lane, p, g0, ng, row, N);
}

}  // namespace spade
