// libspade_b200: sm_100a kernels + C ABI for pySpade's hypergeometric DE hot path.
// See include/spade.h for the contract and the reference lines each entry replaces.
//
// Data layout in HBM (per engine, one engine per GPU)
//   CSR copy of the caller's matrix      indptr int64[G+1], indices int32[nnz], values f64[nnz]
//   gene-partitioned cell-major entries  seg uint2[P*C] (begin, end), ent uint64[padded nnz]
//                                          segment (p, c) = the stored values of cell c whose
//                                          gene lies in partition p (Gp consecutive genes);
//                                          segments start on 128-byte lines (common.cuh)
//                                          a partition whose 32 densest genes cover enough
//                                          entries keeps them as a dense 32-word head of each
//                                          segment (one word per gene, zero = not expressed)
//   per gene                             median f64, n int32, GeneMeta {n, shift, negmode,
//                                          rowsum | bgmean}
//   log-factorial table                  lf f64[C+1], lf[i] = ln(i!)
//   per run                              members int32, offsets int64, TabMeta[G] + tail table
//
// Sums are carried as signed fixed point (value * 2^shift_g, per gene): integer adds commute,
// so results are bit-identical for any launch geometry, any entry order inside a segment and
// any sharding of sets over GPUs.
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

#include "common.cuh"
#include "fused.cuh"
#include "fused_tma.cuh"
#include "gamma_fit.cuh"
#include "payload.cuh"
#include "score.cuh"
#include "stage.cuh"
#include "tails.cuh"
#include "wide.cuh"

using namespace spade;

namespace {

thread_local std::string g_create_error = "";

constexpr int64_t kMaxTestsHostChunk = int64_t(1) << 26;    // host outputs: device tables of 64 Mi tests
constexpr int64_t kMaxTestsSplitChunk = int64_t(1) << 24;   // split mode: global accumulators
constexpr int64_t kMaxTestsTwoPhase = int64_t(1) << 28;     // two-phase mode: 2 GB of accumulator words
constexpr int64_t kMaxTableEntries = int64_t(1) << 26;      // tail table: 1 GiB of double2
constexpr int kDefaultGenesPerPart = 1024;
constexpr int kTargetCopyGroups = 12;                       // host outputs: strips per batch
constexpr double kDenseMinEntries = 6.0;                    // dense block must remove >= this many entries per cell
constexpr int64_t kMaxBitmaskBytes = int64_t(1) << 30;      // wide-gene side path: membership bitmasks per chunk of sets

#define CUDA_TRY(expr)                                                                         \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));                     \
            return SPADE_ERR_CUDA;                                                             \
        }                                                                                      \
    } while (0)

struct TimedSpan {
    cudaEvent_t a, b;
    int kind;     // 0 = fused test kernel, 1 = tail table, 2 = split finalize
};

}  // namespace

struct spade_engine {
    int device = 0;
    int64_t G = 0, C = 0, nnz = 0;
    int Gp = kDefaultGenesPerPart, P = 1;
    int64_t n_nc = 0;

    // options
    int tail_mode = 1;                 // 0 never, 1 auto, 2 whenever the batch has one set size
    int member_block = kMemberBlock;
    int gp_request = 0;                // 0 = default
    int bank_order = 1;                // lay segments out for conflict-free accumulator access
    int dense_blocks = 1;              // register-accumulated dense block per partition
    int sort_members = 1;              // walk every set in ascending cell order (L2 locality)
    int two_choice = 112;              // hot genes per partition with a second accumulator (common.cuh), 0 = none;
                                       // 112: what fits next to 1 024 first accumulators at 6 CTAs per SM
    int coop_request = 0;              // warps per (partition, set) task: 0 = by the number of sets, else 1 / 2 / 4
    int use_tma = 0;                   // test kernel with TMA-staged segments (fused_tma.cuh) instead of k_de_fused
    int hot() const       // at most Gp, and what the 227 KB of shared memory leave for 4 warps
    {
        return std::max(0, std::min(std::min(two_choice, Gp), 227 * 1024 / (8 * kWarpsPerCta) - 32 - Gp)) / 16 * 16;
    }
    int acc_len() const { return Gp + hot(); }
    DevBuf<int16_t> hotb, hot_slot;    // [G] B accumulator of a hot gene (or -1); [P * hot()] its slot

    DevBuf<int64_t> indptr;
    DevBuf<int32_t> indices;
    DevBuf<double> values;
    DevBuf<uint32_t> pptr, cursor;     // pptr = first entry of every segment
    DevBuf<uint2> seg;
    DevBuf<unsigned long long> ent;
    DevBuf<int8_t> dslot;
    DevBuf<int32_t> dgene;
    DevBuf<uint8_t> dense_on;
    bool any_dense = false;
    int64_t ent_len = 0;
    DevBuf<double> median, lf;
    DevBuf<int32_t> n;
    DevBuf<GeneMeta> meta;
    DevBuf<uint8_t> ncmask;
    int *h_err = nullptr, *d_err = nullptr;   // error word: mapped page-locked host memory (common.cuh flag_error)
    std::vector<int32_t> h_n;
    // wide genes (float64 side path, wide.cuh): ascending gene ids, host and device copies
    DevBuf<uint8_t> wide_flag;
    DevBuf<int32_t> wide_genes;
    std::vector<int32_t> h_wide;
    DevBuf<uint32_t> bits_buf[2];

    // run workspace
    DevBuf<int32_t> members, members_sorted;
    DevBuf<uint8_t> sort_temp;
    DevBuf<int64_t> offsets;
    // Tail table, two copies used in turn: a table stays untouched while the NEXT one is built,
    // so work still reading it on another stream (spade_expand of the previous batch) is safe.
    DevBuf<TabMeta> tmeta_buf[2];
    DevBuf<double2> tab_buf[2];
    int tab_cur = 0;
    DevBuf<int32_t> blk_row, blk_cnt;
    DevBuf<int64_t> blk_off;
    DevBuf<unsigned long long> gsum;
    DevBuf<int> gcnt;
    DevBuf<double> out_f64[4];         // host-output path: device tables of one batch
    DevBuf<int32_t> out_k;
    DevBuf<int32_t> kbuf;              // counts of a ragged batch when the caller does not ask for k
    DevBuf<unsigned long long> rawbuf; // accumulator words between the two phases of a uniform batch (two_phase)
    int two_phase = 1;                 // uniform batches with device outputs: k_de_fused leaves raw words,
                                       // k_expand finishes them gene-major (metadata and tail-table window
                                       // of a gene reused over 16 sets, full occupancy)
    cudaStream_t work[2] = {nullptr, nullptr}, copy_stream = nullptr;
    cudaEvent_t ev_ready = nullptr;
    std::vector<cudaEvent_t> sync_events;
    int64_t tab_N = -1;                // set size the tail table on the device was built for
    int64_t tab_epoch = -1, bg_epoch = 0;
    PinnedBuf<TabMeta> h_tmeta;
    PinnedBuf<int64_t> h_offsets, h_sizes;
    PinnedBuf<int32_t> h_order, h_order_x;
    DevBuf<int32_t> order, order_x;    // ragged batches: sets by decreasing size (_x: spade_expand's own copy)
    DevBuf<int64_t> sizes;             // spade_expand's own upload buffers (it may run on another stream)

    std::vector<TimedSpan> spans;
    std::vector<cudaEvent_t> ev_pool;
    size_t ev_used = 0;
    double test_ms = 0.0, table_ms = 0.0, finish_ms = 0.0;
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

    int span_begin(cudaStream_t st, int kind)
    {
        cudaEvent_t a = next_event(), b = next_event();
        if (!a || !b) { set_error("cudaEventCreate failed"); return SPADE_ERR_CUDA; }
        spans.push_back({a, b, kind});
        CUDA_TRY(cudaEventRecord(a, st));
        return SPADE_OK;
    }

    int span_end(cudaStream_t st)
    {
        CUDA_TRY(cudaEventRecord(spans.back().b, st));
        return SPADE_OK;
    }

    int collect_timing()
    {
        for (const TimedSpan &s : spans) {
            float ms = 0.f;
            CUDA_TRY(cudaEventSynchronize(s.b));
            CUDA_TRY(cudaEventElapsedTime(&ms, s.a, s.b));
            (s.kind == 0 || s.kind == 3 ? test_ms : table_ms) += ms;
            if (s.kind == 3) finish_ms += ms;
        }
        spans.clear();
        ev_used = 0;
        return SPADE_OK;
    }

    int ensure_error_word()
    {
        if (h_err) return SPADE_OK;
        CUDA_TRY(cudaHostAlloc((void **)&h_err, sizeof(int), cudaHostAllocMapped | cudaHostAllocPortable));
        *h_err = 0;
        CUDA_TRY(cudaHostGetDevicePointer((void **)&d_err, h_err, 0));
        return SPADE_OK;
    }

    // What the kernels that have FINISHED so far flagged; never waits for the device.  Callers
    // that need the verdict on everything they launched synchronise first.
    int check_device_error(const char *what)
    {
        const int h = h_err ? *(volatile int *)h_err : 0;
        if (h != 0) {
            *(volatile int *)h_err = 0;
            const char *m = h == 2   ? ": member cell index out of range"
                            : h == 3 ? ": matrix holds non-finite (or > 8e270) values"
                            : h == 4 ? ": duplicate cell index inside a cell set"
                                     : ": cell index out of range in the matrix";
            set_error(std::string(what) + m);
            return SPADE_ERR_ARG;
        }
        return SPADE_OK;
    }

    // K1 + entry scatter; re-run whenever the background changes.
    int stage_genes(cudaStream_t st)
    {
        ++bg_epoch;                    // cutoffs change: any tail table on the device is stale
        const size_t np = (size_t)P * (size_t)C;
        if (np > 0) {
            k_init_cursor<<<grid_for(np, 256), 256, 0, st>>>(pptr.p, (int64_t)np, C, dense_on.p, cursor.p);
            CUDA_TRY(cudaGetLastError());
        }
        if (any_dense && ent_len > 0)      // dense heads: words of genes a cell does not express stay zero
            CUDA_TRY(cudaMemsetAsync(ent.p, 0, ent_len * sizeof(unsigned long long), st));
        StageArgs a;
        a.indptr = indptr.p;
        a.indices = indices.p;
        a.values = values.p;
        a.ncmask = n_nc > 0 ? ncmask.p : nullptr;
        a.C = C;
        a.scope_len = n_nc > 0 ? n_nc : C;
        a.Gp = Gp;
        a.cursor = cursor.p;
        a.ent = ent.p;
        a.dslot = dslot.p;
        a.begin = pptr.p;
        a.median = median.p;
        a.n = n.p;
        a.meta = meta.p;
        a.err = d_err;
        a.wide_flag = wide_flag.p;
        if (G > 0) k_gene_stage<<<(unsigned)G, kStageThreads, 0, st>>>(a);
        CUDA_TRY(cudaGetLastError());
        if (nnz > 0 && (bank_order || hot() > 0)) {
            const int64_t nseg = (int64_t)P * C;
            k_order_segments<<<(unsigned)((nseg + kOrderWarps - 1) / kOrderWarps), kOrderWarps * 32, 0, st>>>(
                seg.p, nseg, C, Gp, hot(), hotb.p, dense_on.p, ent.p);
            CUDA_TRY(cudaGetLastError());
        }
        h_n.resize((size_t)G);
        std::vector<uint8_t> hw((size_t)G);
        if (G > 0) {
            CUDA_TRY(cudaMemcpyAsync(h_n.data(), n.p, G * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaMemcpyAsync(hw.data(), wide_flag.p, G * sizeof(uint8_t), cudaMemcpyDeviceToHost, st));
        }
        CUDA_TRY(cudaStreamSynchronize(st));
        h_wide.clear();
        for (int64_t g = 0; g < G; ++g)
            if (hw[g]) h_wide.push_back((int32_t)g);
        if (!h_wide.empty()) {
            CUDA_TRY(wide_genes.reserve(h_wide.size()));
            CUDA_TRY(cudaMemcpy(wide_genes.p, h_wide.data(), h_wide.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
        }
        return check_device_error("staging");
    }

    // Everything after the CSR (indptr/indices/values) is resident on the device.
    int build(cudaStream_t st)
    {
        int erc = ensure_error_word();
        if (erc != SPADE_OK) return erc;

        Gp = gp_request > 0 ? gp_request : kDefaultGenesPerPart;
        Gp = std::min<int64_t>(Gp, std::max<int64_t>(32, (G + 31) / 32 * 32));
        Gp = std::max(32, std::min(Gp, 4096)) / 32 * 32;
        P = (int)std::max<int64_t>(1, (G + Gp - 1) / Gp);
        if ((double)P * (double)C + 1 > 4.0e9) {
            set_error("partition index does not fit: (G / genes_per_part) * C must stay below 4e9");
            return SPADE_ERR_ARG;
        }
        const size_t np = (size_t)P * (size_t)C;

        CUDA_TRY(pptr.reserve(np));
        CUDA_TRY(cursor.reserve(np));
        CUDA_TRY(seg.reserve(np));
        CUDA_TRY(median.reserve(G));
        CUDA_TRY(n.reserve(G));
        CUDA_TRY(meta.reserve(G));
        CUDA_TRY(wide_flag.reserve(G));
        CUDA_TRY(ncmask.reserve(C));
        CUDA_TRY(dslot.reserve(G));
        CUDA_TRY(dgene.reserve((size_t)P * kDenseLanes));
        CUDA_TRY(dense_on.reserve(P));

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

        // Dense block: the kDenseLanes genes of each partition with the most stored values, used
        // when they take enough entries out of the segments to pay for the extra 256-byte read.
        {
            std::vector<int64_t> hp(G + 1);
            CUDA_TRY(cudaMemcpy(hp.data(), indptr.p, (G + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost));
            std::vector<int8_t> hslot(G, (int8_t)-1);
            std::vector<int32_t> hgene((size_t)P * kDenseLanes, -1);
            std::vector<uint8_t> hon(P, 0);
            any_dense = false;
            std::vector<int64_t> order;
            for (int p = 0; p < P && dense_blocks; ++p) {
                const int64_t g0 = (int64_t)p * Gp, g1 = std::min<int64_t>(G, g0 + Gp);
                order.resize(g1 - g0);
                for (int64_t g = g0; g < g1; ++g) order[g - g0] = g;
                const size_t top = std::min<size_t>(kDenseLanes, order.size());
                std::partial_sort(order.begin(), order.begin() + top, order.end(), [&](int64_t x, int64_t y) {
                    const int64_t nx = hp[x + 1] - hp[x], ny = hp[y + 1] - hp[y];
                    return nx != ny ? nx > ny : x < y;
                });
                int64_t covered = 0;
                for (size_t i = 0; i < top; ++i) covered += hp[order[i] + 1] - hp[order[i]];
                if (C > 0 && (double)covered / (double)C >= kDenseMinEntries) {
                    hon[p] = 1;
                    any_dense = true;
                    for (size_t i = 0; i < top; ++i) {
                        hslot[order[i]] = (int8_t)i;
                        hgene[(size_t)p * kDenseLanes + i] = (int32_t)order[i];
                    }
                }
            }
            CUDA_TRY(cudaMemcpy(dslot.p, hslot.data(), G * sizeof(int8_t), cudaMemcpyHostToDevice));
            CUDA_TRY(cudaMemcpy(dgene.p, hgene.data(), hgene.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
            CUDA_TRY(cudaMemcpy(dense_on.p, hon.data(), P * sizeof(uint8_t), cudaMemcpyHostToDevice));

            // Hot genes: per partition the H genes with the most stored values after the dense head
            // get a second accumulator Gp + b whose bank pair (b mod 16) differs from the first's.
            const int H = hot();
            if (H > 0) {
                std::vector<int16_t> hb(G, (int16_t)-1), hs((size_t)P * H, (int16_t)-1);
                std::vector<int64_t> cand;
                for (int p = 0; p < P; ++p) {
                    const int64_t g0 = (int64_t)p * Gp, g1 = std::min<int64_t>(G, g0 + Gp);
                    int16_t *slot_of = hs.data() + (size_t)p * H;
                    if (H >= Gp) {                           // every gene: B at Gp + alt_slot(slot)
                        for (int64_t g = g0; g < g1; ++g) {
                            if (hslot[g] >= 0) continue;
                            const int b = alt_slot((int)(g - g0));
                            hb[g] = (int16_t)b;
                            slot_of[b] = (int16_t)(g - g0);
                        }
                        continue;
                    }
                    cand.clear();
                    for (int64_t g = g0; g < g1; ++g)
                        if (hslot[g] < 0 && hp[g + 1] > hp[g]) cand.push_back(g);
                    const size_t top = std::min<size_t>((size_t)H, cand.size());
                    std::partial_sort(cand.begin(), cand.begin() + top, cand.end(), [&](int64_t x, int64_t y) {
                        const int64_t nx = hp[x + 1] - hp[x], ny = hp[y + 1] - hp[y];
                        return nx != ny ? nx > ny : x < y;
                    });
                    for (size_t i = 0; i < top; ++i) {
                        const int sl = (int)(cand[i] - g0);
                        const int want = alt_slot(sl) & 15;
                        int b = -1;
                        for (int d = 0; d < 16 && b < 0; ++d) {          // preferred pair first, never the gene's own
                            const int pair = (want + d) & 15;
                            if (pair == (sl & 15)) continue;
                            for (int q = pair; q < H; q += 16)
                                if (slot_of[q] < 0) { b = q; break; }
                        }
                        if (b < 0) continue;                             // only its own pair is left: not hot
                        hb[cand[i]] = (int16_t)b;
                        slot_of[b] = (int16_t)sl;
                    }
                }
                CUDA_TRY(hotb.reserve(G));
                CUDA_TRY(hot_slot.reserve(hs.size()));
                CUDA_TRY(cudaMemcpy(hotb.p, hb.data(), G * sizeof(int16_t), cudaMemcpyHostToDevice));
                CUDA_TRY(cudaMemcpy(hot_slot.p, hs.data(), hs.size() * sizeof(int16_t), cudaMemcpyHostToDevice));
            }
        }

        // segment sizes per (partition, cell) -> padded to 128-byte lines -> exclusive scan
        ent_len = 0;
        if (np > 0) {
            DevBuf<uint32_t> count;
            CUDA_TRY(count.reserve(np));
            CUDA_TRY(cudaMemsetAsync(count.p, 0, np * sizeof(uint32_t), st));
            if (G > 0 && nnz > 0) {
                k_part_hist<<<(unsigned)G, kStageThreads, 0, st>>>(indptr.p, indices.p, C, Gp, dslot.p, count.p, d_err);
                CUDA_TRY(cudaGetLastError());
            }
            k_pad_counts<<<grid_for(np, 256), 256, 0, st>>>(count.p, (int64_t)np, C, dense_on.p,
                                                           0u, cursor.p);
            CUDA_TRY(cudaGetLastError());
            size_t scan_bytes = 0;
            CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, cursor.p, pptr.p, (int64_t)np, st));
            DevBuf<uint8_t> temp;
            CUDA_TRY(temp.reserve(scan_bytes));
            // padded total <= nnz + (dense head + 15) per segment: must fit the 32-bit entry index
            if ((unsigned long long)nnz + (unsigned long long)(kSegAlign - 1 + (any_dense ? kDenseLanes : 0)) * np >
                0xffff0000ull) {
                set_error("matrix too large: padded entry count must stay below 2^32 - 65536");
                return SPADE_ERR_ARG;
            }
            CUDA_TRY(cub::DeviceScan::ExclusiveSum(temp.p, scan_bytes, cursor.p, pptr.p, (int64_t)np, st));
            k_make_segments<<<grid_for(np, 256), 256, 0, st>>>(pptr.p, count.p, (int64_t)np, C, dense_on.p, seg.p);
            CUDA_TRY(cudaGetLastError());
            uint32_t last_begin = 0, last_pad = 0;
            CUDA_TRY(cudaMemcpyAsync(&last_begin, pptr.p + (np - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaMemcpyAsync(&last_pad, cursor.p + (np - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaStreamSynchronize(st));
            ent_len = (int64_t)last_begin + (int64_t)last_pad;
            temp.release();
            count.release();
        }
        CUDA_TRY(ent.reserve(ent_len));
        int rc = check_device_error("spade_create");
        if (rc != SPADE_OK) return rc;

        {
            const int max_smem = std::min(227 * 1024, kWarpsPerCta * (2 * 4096 + 32) * 8);
            const void *kernels[] = {(const void *)k_de_fused<true>,
                                     (const void *)k_de_fused<false, false, 1>, (const void *)k_de_fused<false, false, 2>,
                                     (const void *)k_de_fused<false, false, 4>, (const void *)k_de_fused<false, true, 1>,
                                     (const void *)k_de_fused<false, true, 2>, (const void *)k_de_fused<false, true, 4>};
            for (const void *k : kernels) {
                CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
                CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
            }
            const int tma_smem = std::min(227 * 1024, kTmaWarps * tma_warp_words(2 * 4096) * 8);
            CUDA_TRY(cudaFuncSetAttribute(k_de_tma<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tma_smem));
            CUDA_TRY(cudaFuncSetAttribute(k_de_tma<true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
            CUDA_TRY(cudaFuncSetAttribute(k_de_tma<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tma_smem));
            CUDA_TRY(cudaFuncSetAttribute(k_de_tma<false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        }
        return stage_genes(st);
    }

    int ensure_streams()
    {
        for (int i = 0; i < 2; ++i)
            if (!work[i]) CUDA_TRY(cudaStreamCreateWithFlags(&work[i], cudaStreamNonBlocking));
        if (!copy_stream) CUDA_TRY(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
        if (!ev_ready) CUDA_TRY(cudaEventCreateWithFlags(&ev_ready, cudaEventDisableTiming));
        return SPADE_OK;
    }

    int sync_event(size_t i, cudaEvent_t *out)
    {
        while (sync_events.size() <= i) {
            cudaEvent_t e;
            CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            sync_events.push_back(e);
        }
        *out = sync_events[i];
        return SPADE_OK;
    }

    // Tail table for a batch whose sets all have N members: window of plausible k per gene
    // (mean +- 6 sd + 6, clipped to the support; a k outside it is evaluated directly) computed on the host from n_g, entries filled
    // on the device by hyper_tails itself.  *built = false when the table would be too large.
    int build_tail_table(int64_t N, cudaStream_t st, bool *built)
    {
        *built = false;
        tab_N = -1;
        CUDA_TRY(h_tmeta.reserve((size_t)G));
        const double M = (double)C;
        unsigned long long total = 0;
        for (int64_t g = 0; g < G; ++g) {
            TabMeta t;
            t.off = total;
            t.lo = 0;
            t.len = 0;
            const int64_t ng = h_n[g];
            if (ng > 0 && N > 0) {
                const double pr = (double)ng / M;
                const double mu = (double)N * pr;
                const double var = C > 1 ? (double)N * pr * (1.0 - pr) * (M - (double)N) / (M - 1.0) : 0.0;
                const double w = 6.0 * std::sqrt(std::max(var, 0.0)) + 6.0;
                const int64_t a = std::max<int64_t>(N - (C - ng), 0), b = std::min<int64_t>(ng, N);
                const int64_t lo = std::max<int64_t>(a, (int64_t)std::floor(mu - w));
                const int64_t hi = std::min<int64_t>(b, (int64_t)std::ceil(mu + w));
                if (hi >= lo) {
                    t.lo = (int32_t)lo;
                    t.len = (int32_t)(hi - lo + 1);
                }
            }
            h_tmeta.p[g] = t;
            total += (unsigned long long)t.len;
        }
        if (total > (unsigned long long)kMaxTableEntries) return SPADE_OK;
        tab_cur ^= 1;
        DevBuf<TabMeta> &tmeta = tmeta_buf[tab_cur];
        DevBuf<double2> &tab = tab_buf[tab_cur];
        CUDA_TRY(tmeta.reserve(G));
        CUDA_TRY(tab.reserve(total));
        CUDA_TRY(h_tmeta.upload(tmeta.p, (size_t)G, st));
        int rc = span_begin(st, 1);
        if (rc != SPADE_OK) return rc;
        const int64_t threads = G * 32;
        k_tail_table<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(tmeta.p, meta.p, G, (int)C, (int)N,
                                                                       lf.p, tab.p);
        CUDA_TRY(cudaGetLastError());
        rc = span_end(st);
        if (rc != SPADE_OK) return rc;
        launches += 1;
        *built = true;
        tab_N = N;
        tab_epoch = bg_epoch;
        return SPADE_OK;
    }

    // mean_n: mean number of member cells per set of the launch (decides the warps per task)
    int launch_fused(const RunArgs &a, bool split, cudaStream_t st, int64_t mean_n = 0)
    {
        const int64_t units = split ? a.nblk : a.S;
        const int64_t tasks = units * (a.p1 - a.p0);
        if (tasks <= 0) return SPADE_OK;
        const size_t tma_smem = (size_t)kTmaWarps * tma_warp_words(acc_len()) * sizeof(unsigned long long);
        const bool tma = use_tma && !split && tma_smem <= 227 * 1024;
        // Warps per task (RunArgs::coop): enough warp tasks per partition that the ~3 500 resident
        // warps of the GPU cover two or three partitions, whose layout then stays in L2.
        int coop = 1;
        if (!split && !tma) {
            // Ragged batches only (a.order set): a long region is cut into four shorter member lists and the
            // resident warps cover fewer partitions.  Measured on 50 000 x 33 000: 500 ragged regions 1.58 /
            // 1.44 / 1.15 ms with 1 / 2 / 4 warps per task; uniform batches lose (400 draws of 500 cells
            // 0.78 -> 0.87 ms with 4, 1 000 draws 1.48 -> 1.56 ms with 2).
            const bool ragged = a.order != nullptr;
            coop = coop_request > 0 ? coop_request
                   : !ragged || mean_n < 128 ? 1 : units <= 600 ? 4 : units <= 1200 ? 2 : 1;
            if (coop != 2 && coop != 4) coop = 1;
        }
        RunArgs b = a;
        b.coop = coop;
        const int warps = tma ? kTmaWarps : kWarpsPerCta / coop;      // tasks per CTA
        const int64_t blocks = (tasks + warps - 1) / warps;
        if (blocks > 0x7fffffffll) { set_error("batch too large for one launch"); return SPADE_ERR_ARG; }
        const size_t smem = (size_t)kWarpsPerCta * (acc_len() + 32) * sizeof(unsigned long long);   // accumulators + pair buffers
        int rc = span_begin(st, 0);
        if (rc != SPADE_OK) return rc;
        if (tma && use_tma == 1) k_de_tma<true><<<(unsigned)blocks, kTmaWarps * 32, tma_smem, st>>>(b);
        else if (tma) k_de_tma<false><<<(unsigned)blocks, kTmaWarps * 32, tma_smem, st>>>(b);
        else if (split) k_de_fused<true><<<(unsigned)blocks, kWarpsPerCta * 32, smem, st>>>(b);
        else {
            const bool raw_only = a.raw || a.ndst > 0;
            auto kernel = raw_only ? (coop == 4 ? k_de_fused<false, true, 4> : coop == 2 ? k_de_fused<false, true, 2>
                                                                                         : k_de_fused<false, true, 1>)
                                   : (coop == 4 ? k_de_fused<false, false, 4> : coop == 2 ? k_de_fused<false, false, 2>
                                                                                          : k_de_fused<false, false, 1>);
            kernel<<<(unsigned)blocks, kWarpsPerCta * 32, smem, st>>>(b);
        }
        CUDA_TRY(cudaGetLastError());
        rc = span_end(st);
        if (rc != SPADE_OK) return rc;
        launches += 1;
        return SPADE_OK;
    }

    // up / down of rows [0, Sb) x genes [g0, g1) from the counts the test kernel left in k
    // (ragged batches, see k_tails_direct).  sizes: device N per row, or null -> offsets of set s0.
    int launch_tails(cudaStream_t st, const int32_t *k, int64_t k_ld, const int64_t *sizes, const int32_t *by_size,
                     int64_t s0, int64_t Sb, int64_t g0, int64_t g1, int64_t col0, double *up, double *down, int64_t ld)
    {
        if (Sb <= 0 || g1 <= g0 || (!up && !down)) return SPADE_OK;
        int rc = span_begin(st, 2);
        if (rc != SPADE_OK) return rc;
        const dim3 grid((unsigned)((g1 - g0 + 7) / 8), (unsigned)std::min<int64_t>((Sb + 31) / 32, 64));
        k_tails_direct<<<grid, 256, 0, st>>>(k, k_ld, meta.p, sizes, sizes ? nullptr : offsets.p + s0, by_size, (int)C,
                                             lf.p, up, down, ld, Sb, g0, g1, col0);
        CUDA_TRY(cudaGetLastError());
        rc = span_end(st);
        if (rc != SPADE_OK) return rc;
        launches += 1;
        return SPADE_OK;
    }

    // Float64 means of the wide genes in [g0, g1) for sets [s0, s0 + Sb) of the uploaded batch,
    // written over what the fixed-point path left in fc / cpm (row 0 = set s0, row stride ld).
    // Ordered on `st` after the kernels that wrote those tables.  `which`: bitmask buffer to use
    // (two streams may run this at the same time in the host-output pipeline).
    int fix_wide(cudaStream_t st, int which, const int32_t *d_members, int64_t s0, int64_t Sb, int64_t g0,
                 int64_t g1, double *fc, double *cpm, int64_t ld)
    {
        if (h_wide.empty() || (!fc && !cpm) || Sb <= 0) return SPADE_OK;
        const size_t lo = std::lower_bound(h_wide.begin(), h_wide.end(), (int32_t)g0) - h_wide.begin();
        const size_t hi = std::lower_bound(h_wide.begin(), h_wide.end(), (int32_t)std::min<int64_t>(g1, 0x7fffffff)) - h_wide.begin();
        if (hi <= lo) return SPADE_OK;
        const int64_t words = (C + 31) / 32;
        const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(Sb, kMaxBitmaskBytes / (words * 4)));
        DevBuf<uint32_t> &bits = bits_buf[which];
        CUDA_TRY(bits.reserve((size_t)(chunk * words)));
        for (int64_t c0 = 0; c0 < Sb; c0 += chunk) {
            const int64_t nc = std::min(chunk, Sb - c0);
            CUDA_TRY(cudaMemsetAsync(bits.p, 0, (size_t)(nc * words) * sizeof(uint32_t), st));
            for (int64_t b0 = 0; b0 < nc; b0 += 0x7fffffff) {
                const int64_t nb = std::min<int64_t>(nc - b0, 0x7fffffff);
                k_set_bitmask<<<(unsigned)nb, 256, 0, st>>>(d_members, offsets.p, s0 + c0 + b0, nb, C, words,
                                                           bits.p + b0 * words);
                CUDA_TRY(cudaGetLastError());
            }
            WideArgs w;
            w.indptr = indptr.p; w.indices = indices.p; w.values = values.p; w.meta = meta.p;
            w.bits = bits.p; w.words = words; w.offsets = offsets.p; w.C = C; w.nc_mode = n_nc > 0;
            w.ld = ld;
            for (int64_t r0 = 0; r0 < nc; r0 += (int64_t)65535 * kWideWarps) {           // gridDim.y limit
                const int64_t nr = std::min<int64_t>(nc - r0, (int64_t)65535 * kWideWarps);
                w.bits = bits.p + r0 * words;
                w.s0 = s0 + c0 + r0;
                w.S = nr;
                w.fc = fc ? fc + (c0 + r0) * ld : nullptr;
                w.cpm = cpm ? cpm + (c0 + r0) * ld : nullptr;
                for (size_t i0 = lo; i0 < hi; i0 += 0x7fffffff) {
                    w.genes = wide_genes.p + i0;
                    w.nw = (int)std::min<size_t>(hi - i0, 0x7fffffff);
                    const dim3 grid((unsigned)w.nw, (unsigned)((nr + kWideWarps - 1) / kWideWarps));
                    k_wide_means<<<grid, kWideWarps * 32, 0, st>>>(w);
                    CUDA_TRY(cudaGetLastError());
                }
            }
            launches += 2;
        }
        return SPADE_OK;
    }

    void free_all()
    {
        indptr.release(); indices.release(); values.release(); pptr.release(); cursor.release();
        seg.release(); dslot.release(); dgene.release(); dense_on.release(); hotb.release(); hot_slot.release();
        ent.release(); median.release(); lf.release(); n.release(); meta.release(); ncmask.release();
        wide_flag.release(); wide_genes.release(); bits_buf[0].release(); bits_buf[1].release();
        if (h_err) cudaFreeHost(h_err);
        h_err = d_err = nullptr;
        members.release(); members_sorted.release(); sort_temp.release();
        offsets.release();
        for (int i = 0; i < 2; ++i) { tmeta_buf[i].release(); tab_buf[i].release(); }
        blk_row.release(); blk_cnt.release(); blk_off.release(); gsum.release(); gcnt.release();
        for (int b = 0; b < 4; ++b) out_f64[b].release();
        out_k.release();
        kbuf.release();
        rawbuf.release();
        h_tmeta.release();
        h_offsets.release();
        h_sizes.release();
        h_order.release();
        h_order_x.release();
        order.release();
        order_x.release();
        sizes.release();
        for (int i = 0; i < 2; ++i) {
            if (work[i]) cudaStreamDestroy(work[i]);
            work[i] = nullptr;
        }
        if (copy_stream) cudaStreamDestroy(copy_stream);
        copy_stream = nullptr;
        if (ev_ready) cudaEventDestroy(ev_ready);
        ev_ready = nullptr;
        for (cudaEvent_t e : sync_events) cudaEventDestroy(e);
        sync_events.clear();
        for (cudaEvent_t e : ev_pool) cudaEventDestroy(e);
        ev_pool.clear();
        spans.clear();
    }
};

namespace {

int create_common(int device, int64_t G, int64_t C, const int64_t *indptr, const int32_t *indices,
                  const double *values, const int32_t *counts, int location, spade_engine **out)
{
    auto fail = [&](int code, const std::string &m) { g_create_error = m; return code; };
    if (!out) return fail(SPADE_ERR_ARG, "engine_out is NULL");
    *out = nullptr;
    if (G < 0 || C < 0 || C > 0x7fffffff || G > 0x7fffffff) return fail(SPADE_ERR_ARG, "G and C must be in [0, 2^31)");
    if (!indptr) return fail(SPADE_ERR_ARG, "indptr is NULL");
    if (location != SPADE_HOST && location != SPADE_DEVICE) return fail(SPADE_ERR_ARG, "bad location");
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
    if (const char *gp = getenv("SPADE_GENES_PER_PART")) e->gp_request = atoi(gp);
    if (const char *bo = getenv("SPADE_BANK_ORDER")) e->bank_order = atoi(bo) != 0;
    if (const char *db = getenv("SPADE_DENSE_BLOCKS")) e->dense_blocks = atoi(db) != 0;
    if (const char *tc = getenv("SPADE_TWO_CHOICE")) e->two_choice = atoi(tc) == 1 ? 8192 : atoi(tc);   // 1 = every gene
    if (const char *tm = getenv("SPADE_TMA")) e->use_tma = atoi(tm);   // 1 = cp.async.bulk (TMA), 2 = cp.async (LDGSTS)
    if (const char *tp = getenv("SPADE_TWO_PHASE")) e->two_phase = atoi(tp) != 0;
    if (const char *co = getenv("SPADE_COOP")) e->coop_request = atoi(co);
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
    if (first != 0 || nnz < 0 || nnz > 0xffff0000ll) {
        e->set_error("indptr must start at 0 and hold fewer than 2^32 - 65536 entries");
        return bail(SPADE_ERR_ARG);
    }
    if (nnz > 0 && (!indices || (!values && !counts))) {
        e->set_error("indices/values is NULL");
        return bail(SPADE_ERR_ARG);
    }
    e->nnz = nnz;
    E_TRY(e->indices.reserve(nnz));
    E_TRY(e->values.reserve(nnz));
    if (e->ensure_error_word() != SPADE_OK) return bail(SPADE_ERR_CUDA);
    if (nnz > 0) {
        E_TRY(cudaMemcpy(e->indices.p, indices, nnz * sizeof(int32_t), kind));
        if (counts) {
            DevBuf<int32_t> dcounts;
            DevBuf<unsigned long long> colsum;
            E_TRY(dcounts.reserve(nnz));
            E_TRY(colsum.reserve(C));
            E_TRY(cudaMemcpy(dcounts.p, counts, nnz * sizeof(int32_t), kind));
            E_TRY(cudaMemsetAsync(colsum.p, 0, C * sizeof(unsigned long long), st));
            k_col_sums<<<grid_for(nnz, 256), 256, 0, st>>>(e->indices.p, dcounts.p, nnz, C, colsum.p, e->d_err);
            E_TRY(cudaGetLastError());
            k_cpm<<<grid_for(nnz, 256), 256, 0, st>>>(e->indices.p, dcounts.p, nnz, C, colsum.p, e->values.p);
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
    int rc = e->build(st);
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
    cudaDeviceSynchronize();
    engine->free_all();
    delete engine;
}

#define set_error engine->set_error

int spade_set_background(spade_engine *engine, const int32_t *nc_idx, int64_t n_nc)
{
    if (!engine) return SPADE_ERR_ARG;
    CUDA_TRY(cudaSetDevice(engine->device));
    if (n_nc < 0 || (n_nc > 0 && !nc_idx)) { set_error("bad nc_idx / n_nc"); return SPADE_ERR_ARG; }
    CUDA_TRY(cudaDeviceSynchronize());
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
    return engine->stage_genes(0);
}

int spade_set_option(spade_engine *engine, int option, int64_t value)
{
    if (!engine) return SPADE_ERR_ARG;
    CUDA_TRY(cudaSetDevice(engine->device));
    switch (option) {
    case SPADE_OPT_TAIL_TABLE:
        if (value < 0 || value > 2) { set_error("tail table mode must be 0, 1 or 2"); return SPADE_ERR_ARG; }
        engine->tail_mode = (int)value;
        return SPADE_OK;
    case SPADE_OPT_MEMBER_BLOCK:
        if (value < 1 || value > kMemberBlock) { set_error("member block must be in [1, 32768]"); return SPADE_ERR_ARG; }
        engine->member_block = (int)value;
        return SPADE_OK;
    case SPADE_OPT_SORT_MEMBERS:
        engine->sort_members = value != 0;
        return SPADE_OK;
    case SPADE_OPT_GENES_PER_PART:
        if (value < 0 || value > 4096) { set_error("genes per partition must be in [0, 4096]"); return SPADE_ERR_ARG; }
        CUDA_TRY(cudaDeviceSynchronize());
        engine->gp_request = (int)value;
        return engine->build(0);
    default:
        set_error("unknown option");
        return SPADE_ERR_ARG;
    }
}

int spade_dims(const spade_engine *engine, int64_t *G, int64_t *C, int64_t *nnz, int32_t *genes_per_part)
{
    if (!engine) return SPADE_ERR_ARG;
    if (G) *G = engine->G;
    if (C) *C = engine->C;
    if (nnz) *nnz = engine->nnz;
    if (genes_per_part) *genes_per_part = engine->Gp;
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

static int run_impl(spade_engine *engine, const int32_t *members, int members_location,
                    const int64_t *offsets, int64_t S, double *up, double *down, double *fc, double *cpm,
                    int32_t *k, int out_location, int64_t out_row_stride, void *stream,
                    unsigned long long *raw = nullptr, const spade_gene_block *blocks = nullptr, int nblocks = 0,
                    int64_t first_row = 0, int64_t row_step = 1)
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
    {   // member errors of earlier device-output runs that have finished by now (see spade_check)
        int rc = engine->check_device_error("an earlier spade_run");
        if (rc != SPADE_OK) return rc;
    }
    if (raw && !engine->h_wide.empty()) {
        set_error("the raw form is not available: the matrix holds genes that need the float64 side path "
                  "(spade_wide_genes)");
        return SPADE_ERR_STATE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t G = engine->G, C = engine->C;
    const int P = engine->P;

    const int32_t *d_members = members;
    if (members_location == SPADE_HOST) {
        CUDA_TRY(engine->members.reserve(total));
        if (total > 0)
            CUDA_TRY(cudaMemcpyAsync(engine->members.p, members, total * sizeof(int32_t),
                                     cudaMemcpyHostToDevice, st));
        d_members = engine->members.p;
    }
    CUDA_TRY(engine->offsets.reserve(S + 1));
    CUDA_TRY(engine->h_offsets.reserve(S + 1));
    memcpy(engine->h_offsets.p, offsets, (S + 1) * sizeof(int64_t));
    CUDA_TRY(engine->h_offsets.upload(engine->offsets.p, (size_t)(S + 1), st));

    // Member cells of every set in ascending order: the warps working on one gene partition
    // then sweep the cells in step, so a segment fetched for one set is still in L2 when the
    // other sets that contain the cell ask for it.  Order inside a set does not change any sum.
    if (engine->sort_members && total > 1) {
        CUDA_TRY(engine->members_sorted.reserve(total));
        int64_t maxN = 0;
        for (int64_t s = 0; s < S; ++s) maxN = std::max(maxN, offsets[s + 1] - offsets[s]);
        if (maxN <= kSortMax) {
            int m = 32;
            while (m < maxN) m <<= 1;
            k_sort_members<<<(unsigned)S, kSortThreads, m * sizeof(int32_t), st>>>(d_members, engine->offsets.p,
                                                                                engine->members_sorted.p, engine->d_err);
            CUDA_TRY(cudaGetLastError());
            engine->launches += 1;
        } else {
            // very large sets: CUB's segmented sort (it synchronises the stream internally)
            size_t tb = 0;
            CUDA_TRY(cub::DeviceSegmentedSort::SortKeys(nullptr, tb, d_members, engine->members_sorted.p, total, S,
                                                        engine->offsets.p, engine->offsets.p + 1, st));
            CUDA_TRY(engine->sort_temp.reserve(tb));
            CUDA_TRY(cub::DeviceSegmentedSort::SortKeys(engine->sort_temp.p, tb, d_members, engine->members_sorted.p,
                                                        total, S, engine->offsets.p, engine->offsets.p + 1, st));
            k_check_duplicates<<<grid_for(total, 256), 256, 0, st>>>(engine->members_sorted.p, engine->offsets.p, S, total,
                                                                     engine->d_err);
            CUDA_TRY(cudaGetLastError());
            engine->launches += 2;
        }
        d_members = engine->members_sorted.p;
    }

    const bool host_out = out_location == SPADE_HOST;
    double *outs[4] = {up, down, fc, cpm};
    if (host_out) {
        int rc = engine->ensure_streams();
        if (rc != SPADE_OK) return rc;
    }

    int64_t s0 = 0;
    while (s0 < S) {
        // ---- batch [s0, s1): as many sets as the workspace allows ----
        int64_t s1 = S;
        if (host_out) s1 = std::min(S, s0 + std::max<int64_t>(1, kMaxTestsHostChunk / G));
        else if (engine->two_phase && !raw) s1 = std::min(S, s0 + std::max<int64_t>(1, kMaxTestsTwoPhase / G));
        int64_t maxN = 0;
        bool uniform = true;
        const int64_t N0 = offsets[s0 + 1] - offsets[s0];
        for (int64_t s = s0; s < s1; ++s) {
            const int64_t N = offsets[s + 1] - offsets[s];
            maxN = std::max(maxN, N);
            uniform &= (N == N0);
        }
        const bool split = maxN > engine->member_block;
        if (split && raw) {
            set_error("the raw form takes sets of at most SPADE_OPT_MEMBER_BLOCK cells");
            return SPADE_ERR_ARG;
        }
        if (split && (s1 - s0) * G > kMaxTestsSplitChunk) {
            s1 = s0 + std::max<int64_t>(1, kMaxTestsSplitChunk / G);
            uniform = true;
            for (int64_t s = s0; s < s1; ++s) uniform &= (offsets[s + 1] - offsets[s] == N0);
        }
        const int64_t Sb = s1 - s0;

        bool table = false;
        if (!raw && uniform && N0 > 0 && (engine->tail_mode == 2 || (engine->tail_mode == 1 && Sb >= 32))) {
            if (engine->tab_N == N0 && engine->tab_epoch == engine->bg_epoch) {
                table = true;      // a pure function of (matrix, background, N): kept across calls
            } else {
                int rc = engine->build_tail_table(N0, st, &table);
                if (rc != SPADE_OK) return rc;
            }
        }

        RunArgs a;
        memset(&a, 0, sizeof(a));
        a.ent = engine->ent.p;
        a.seg = engine->seg.p;
        a.dgene = engine->dgene.p;
        a.dense_on = engine->dense_on.p;
        a.members = d_members;
        a.offsets = engine->offsets.p;
        a.s0 = s0;
        a.S = Sb;
        a.Gp = engine->Gp;
        a.two = engine->hot();
        a.hot_slot = engine->hot_slot.p;
        a.acc_len = engine->acc_len();
        a.C = C;
        a.err = engine->d_err;
        a.raw = (raw && !blocks) ? raw + s0 * out_row_stride : nullptr;
        a.raw_ld = out_row_stride;
        if (blocks) {
            a.ndst = nblocks;
            for (int b = 0; b < nblocks; ++b) {
                a.dst_base[b] = (unsigned long long *)blocks[b].base;
                a.dst_ld[b] = blocks[b].ld;
                a.dst_pbegin[b] = (int)(blocks[b].gene_begin / engine->Gp);
            }
            a.dst_pbegin[nblocks] = P;
            a.raw_row0 = first_row + s0 * row_step;
            a.raw_row_step = row_step;
        }
        a.fin.G = G;
        a.fin.M = (int)C;
        a.fin.nc_mode = engine->n_nc > 0;
        a.fin.meta = engine->meta.p;
        a.fin.tmeta = table ? engine->tmeta_buf[engine->tab_cur].p : nullptr;
        a.fin.tab = table ? engine->tab_buf[engine->tab_cur].p : nullptr;
        a.fin.lf = engine->lf.p;
        a.fin.ld = host_out ? G : out_row_stride;
        if (host_out) {
            for (int t = 0; t < 4; ++t)
                if (outs[t]) CUDA_TRY(engine->out_f64[t].reserve(Sb * G));
            if (k) CUDA_TRY(engine->out_k.reserve(Sb * G));
            a.fin.up = up ? engine->out_f64[0].p : nullptr;
            a.fin.down = down ? engine->out_f64[1].p : nullptr;
            a.fin.fc = fc ? engine->out_f64[2].p : nullptr;
            a.fin.cpm = cpm ? engine->out_f64[3].p : nullptr;
            a.fin.k = k ? engine->out_k.p : nullptr;
        } else {
            a.fin.up = up ? up + s0 * out_row_stride : nullptr;
            a.fin.down = down ? down + s0 * out_row_stride : nullptr;
            a.fin.fc = fc ? fc + s0 * out_row_stride : nullptr;
            a.fin.cpm = cpm ? cpm + s0 * out_row_stride : nullptr;
            a.fin.k = k ? k + s0 * out_row_stride : nullptr;
        }

        // Ragged batch: longest sets first (the cost of a warp's task is proportional to its set size;
        // in the given order a few large regions scheduled late leave most of the GPU idle at the end)
        if (!uniform && !split && Sb > 1) {
            CUDA_TRY(engine->h_order.reserve((size_t)Sb));
            CUDA_TRY(engine->order.reserve((size_t)Sb));
            int32_t *o = engine->h_order.p;
            for (int64_t i = 0; i < Sb; ++i) o[i] = (int32_t)i;
            std::stable_sort(o, o + Sb, [&](int32_t x, int32_t y) {
                return offsets[s0 + x + 1] - offsets[s0 + x] > offsets[s0 + y + 1] - offsets[s0 + y];
            });
            CUDA_TRY(engine->h_order.upload(engine->order.p, (size_t)Sb, st));
            a.order = engine->order.p;
        }

        // Ragged batch (no tail table): the test kernel leaves the counts, k_tails_direct turns them
        // into up / down afterwards at full occupancy.
        const bool defer = !table && !raw && !split && (a.fin.up || a.fin.down);
        double *t_up = a.fin.up, *t_down = a.fin.down;
        const int32_t *t_k = a.fin.k;
        int64_t t_kld = a.fin.ld;
        if (defer) {
            if (!a.fin.k) {
                CUDA_TRY(engine->kbuf.reserve(Sb * G));
                a.fin.k = engine->kbuf.p;
                a.fin.k_ld = G;
                t_k = engine->kbuf.p;
                t_kld = G;
            }
            a.fin.up = a.fin.down = nullptr;
        }

        if (split) {
            // member blocks of at most member_block cells; partial sums meet in global accumulators
            std::vector<int32_t> brow, bcnt;
            std::vector<int64_t> boff;
            for (int64_t s = s0; s < s1; ++s)
                for (int64_t o = offsets[s]; o < offsets[s + 1]; o += engine->member_block) {
                    brow.push_back((int32_t)(s - s0));
                    boff.push_back(o);
                    bcnt.push_back((int32_t)std::min<int64_t>(engine->member_block, offsets[s + 1] - o));
                }
            const int64_t nb = (int64_t)brow.size();
            CUDA_TRY(engine->blk_row.reserve(nb));
            CUDA_TRY(engine->blk_off.reserve(nb));
            CUDA_TRY(engine->blk_cnt.reserve(nb));
            CUDA_TRY(engine->gsum.reserve(Sb * G));
            CUDA_TRY(engine->gcnt.reserve(Sb * G));
            CUDA_TRY(cudaMemcpyAsync(engine->blk_row.p, brow.data(), nb * sizeof(int32_t), cudaMemcpyHostToDevice, st));
            CUDA_TRY(cudaMemcpyAsync(engine->blk_off.p, boff.data(), nb * sizeof(int64_t), cudaMemcpyHostToDevice, st));
            CUDA_TRY(cudaMemcpyAsync(engine->blk_cnt.p, bcnt.data(), nb * sizeof(int32_t), cudaMemcpyHostToDevice, st));
            CUDA_TRY(cudaMemsetAsync(engine->gsum.p, 0, Sb * G * sizeof(unsigned long long), st));
            CUDA_TRY(cudaMemsetAsync(engine->gcnt.p, 0, Sb * G * sizeof(int), st));
            a.blk_row = engine->blk_row.p;
            a.blk_off = engine->blk_off.p;
            a.blk_cnt = engine->blk_cnt.p;
            a.nblk = nb;
            a.gsum = engine->gsum.p;
            a.gcnt = engine->gcnt.p;
            a.p0 = 0;
            a.p1 = P;
            int rc = engine->launch_fused(a, true, st);
            if (rc != SPADE_OK) return rc;
            rc = engine->span_begin(st, 2);
            if (rc != SPADE_OK) return rc;
            k_finalize_split<<<(unsigned)((Sb * G + 255) / 256), 256, 0, st>>>(a.fin, a.gsum, a.gcnt,
                                                                              a.offsets, s0, Sb);
            CUDA_TRY(cudaGetLastError());
            rc = engine->span_end(st);
            if (rc != SPADE_OK) return rc;
            engine->launches += 1;
            rc = engine->fix_wide(st, 0, d_members, s0, Sb, 0, G, a.fin.fc, a.fin.cpm, a.fin.ld);
            if (rc != SPADE_OK) return rc;
            CUDA_TRY(cudaStreamSynchronize(st));     // the host vectors above go out of scope
            if (host_out) {
                for (int t = 0; t < 4; ++t)
                    if (outs[t])
                        CUDA_TRY(cudaMemcpy2DAsync(outs[t] + s0 * out_row_stride, out_row_stride * sizeof(double),
                                                   engine->out_f64[t].p, G * sizeof(double), G * sizeof(double),
                                                   Sb, cudaMemcpyDeviceToHost, st));
                if (k)
                    CUDA_TRY(cudaMemcpy2DAsync(k + s0 * out_row_stride, out_row_stride * sizeof(int32_t),
                                               engine->out_k.p, G * sizeof(int32_t), G * sizeof(int32_t), Sb,
                                               cudaMemcpyDeviceToHost, st));
                CUDA_TRY(cudaStreamSynchronize(st));
            }
        } else if (!host_out) {
            a.p0 = 0;
            a.p1 = P;
            int rc;
            if (engine->two_phase && table && !raw) {
                // Uniform batch (DErand), device tables, two phases: the test kernel leaves its 8-byte accumulator words,
                // k_expand finishes them with a thread per gene walking 8 sets (gene metadata and
                // tail-table window reused across the sets, full occupancy): 0.90 -> 0.64 ms per 1 000
                // draws of 50 cells, 2.23 -> 2.08 ms at 500 cells.  Ragged batches keep the fused
                // finalize (their tails are deferred anyway; measured 1.58 vs 1.64 ms per 500 regions).
                CUDA_TRY(engine->rawbuf.reserve((size_t)Sb * (size_t)G));
                RunArgs a2 = a;
                a2.raw = engine->rawbuf.p;
                a2.raw_ld = G;
                rc = engine->launch_fused(a2, false, st, (offsets[s1] - offsets[s0]) / Sb);
                if (rc != SPADE_OK) return rc;
                rc = engine->span_begin(st, 3);
                if (rc != SPADE_OK) return rc;
                const dim3 grid((unsigned)((G + 255) / 256), (unsigned)((Sb + kExpandRows - 1) / kExpandRows));
                k_expand<<<grid, 256, 0, st>>>(a.fin, engine->rawbuf.p, G, nullptr, engine->offsets.p + s0, Sb, 0, G);
                CUDA_TRY(cudaGetLastError());
                rc = engine->span_end(st);
                if (rc != SPADE_OK) return rc;
                engine->launches += 1;
            } else {
                rc = engine->launch_fused(a, false, st, (offsets[s1] - offsets[s0]) / Sb);
                if (rc != SPADE_OK) return rc;
            }
            if (defer) {
                rc = engine->launch_tails(st, t_k, t_kld, nullptr, a.order, s0, Sb, 0, G, 0, t_up, t_down, a.fin.ld);
                if (rc != SPADE_OK) return rc;
            }
            if (!raw) {
                rc = engine->fix_wide(st, 0, d_members, s0, Sb, 0, G, a.fin.fc, a.fin.cpm, a.fin.ld);
                if (rc != SPADE_OK) return rc;
            }
        } else {
            // Host outputs: strips of gene partitions, kernels alternating over two streams,
            // each finished strip copied out (2-D copy into the caller's [S][G] tables) while
            // the next strips are computed.
            CUDA_TRY(cudaEventRecord(engine->ev_ready, st));
            for (int i = 0; i < 2; ++i) CUDA_TRY(cudaStreamWaitEvent(engine->work[i], engine->ev_ready, 0));
            const int pg = std::max(1, (P + kTargetCopyGroups - 1) / kTargetCopyGroups);
            size_t gi = 0;
            for (int p0 = 0; p0 < P; p0 += pg, ++gi) {
                cudaStream_t ws = engine->work[gi & 1];
                a.p0 = p0;
                a.p1 = std::min(P, p0 + pg);
                int rc = engine->launch_fused(a, false, ws, (offsets[s1] - offsets[s0]) / Sb);
                if (rc != SPADE_OK) return rc;
                if (defer) {
                    rc = engine->launch_tails(ws, t_k, t_kld, nullptr, a.order, s0, Sb, (int64_t)a.p0 * engine->Gp,
                                              std::min<int64_t>(G, (int64_t)a.p1 * engine->Gp), 0, t_up, t_down, a.fin.ld);
                    if (rc != SPADE_OK) return rc;
                }
                rc = engine->fix_wide(ws, (int)(gi & 1), d_members, s0, Sb, (int64_t)a.p0 * engine->Gp,
                                      std::min<int64_t>(G, (int64_t)a.p1 * engine->Gp), a.fin.fc, a.fin.cpm, a.fin.ld);
                if (rc != SPADE_OK) return rc;
                cudaEvent_t done;
                rc = engine->sync_event(gi, &done);
                if (rc != SPADE_OK) return rc;
                CUDA_TRY(cudaEventRecord(done, ws));
                CUDA_TRY(cudaStreamWaitEvent(engine->copy_stream, done, 0));
                const int64_t g0 = (int64_t)a.p0 * engine->Gp;
                const int64_t g1 = std::min<int64_t>(G, (int64_t)a.p1 * engine->Gp);
                for (int t = 0; t < 4; ++t)
                    if (outs[t])
                        CUDA_TRY(cudaMemcpy2DAsync(outs[t] + s0 * out_row_stride + g0, out_row_stride * sizeof(double),
                                                   engine->out_f64[t].p + g0, G * sizeof(double),
                                                   (g1 - g0) * sizeof(double), Sb, cudaMemcpyDeviceToHost,
                                                   engine->copy_stream));
                if (k)
                    CUDA_TRY(cudaMemcpy2DAsync(k + s0 * out_row_stride + g0, out_row_stride * sizeof(int32_t), engine->out_k.p + g0,
                                               G * sizeof(int32_t), (g1 - g0) * sizeof(int32_t), Sb,
                                               cudaMemcpyDeviceToHost, engine->copy_stream));
            }
            CUDA_TRY(cudaStreamSynchronize(engine->copy_stream));
            for (int i = 0; i < 2; ++i) CUDA_TRY(cudaStreamSynchronize(engine->work[i]));
        }
        if (engine->spans.size() > 2000) {            // bound the event pool on very long runs
            int rc = engine->collect_timing();
            if (rc != SPADE_OK) return rc;
        }
        s0 = s1;
    }
    if (host_out) {
        CUDA_TRY(cudaStreamSynchronize(st));
        return engine->check_device_error("spade_run");
    }
    return SPADE_OK;
}

int spade_run(spade_engine *engine, const int32_t *members, int members_location,
              const int64_t *offsets, int64_t S, double *up, double *down, double *fc, double *cpm,
              int32_t *k, int out_location, void *stream)
{
    return run_impl(engine, members, members_location, offsets, S, up, down, fc, cpm, k, out_location,
                    engine ? engine->G : 0, stream);
}

int spade_run_strided(spade_engine *engine, const int32_t *members, int members_location,
                      const int64_t *offsets, int64_t S, double *up, double *down, double *fc,
                      double *cpm, int32_t *k, int out_location, int64_t out_row_stride, void *stream)
{
    if (!engine) return SPADE_ERR_ARG;
    if (out_row_stride < engine->G) { set_error("out_row_stride must be >= G"); return SPADE_ERR_ARG; }
    return run_impl(engine, members, members_location, offsets, S, up, down, fc, cpm, k, out_location,
                    out_row_stride, stream);
}

int spade_run_raw(spade_engine *engine, const int32_t *members, int members_location,
                  const int64_t *offsets, int64_t S, uint64_t *raw, int64_t out_row_stride, void *stream)
{
    if (!engine) return SPADE_ERR_ARG;
    if (!raw) { set_error("raw is NULL"); return SPADE_ERR_ARG; }
    if (out_row_stride < engine->G) { set_error("out_row_stride must be >= G"); return SPADE_ERR_ARG; }
    return run_impl(engine, members, members_location, offsets, S, nullptr, nullptr, nullptr, nullptr, nullptr,
                    SPADE_DEVICE, out_row_stride, stream, (unsigned long long *)raw);
}

int spade_run_raw_blocks(spade_engine *engine, const int32_t *members, int members_location,
                         const int64_t *offsets, int64_t S, const spade_gene_block *blocks, int nblocks,
                         int64_t first_row, int64_t row_step, void *stream)
{
    if (!engine) return SPADE_ERR_ARG;
    if (!blocks || nblocks < 1 || nblocks > kMaxGeneBlocks) { set_error("nblocks must be in [1, 16]"); return SPADE_ERR_ARG; }
    if (first_row < 0 || row_step < 1) { set_error("first_row must be >= 0 and row_step >= 1"); return SPADE_ERR_ARG; }
    int64_t at = 0;
    for (int b = 0; b < nblocks; ++b) {
        const spade_gene_block &gb = blocks[b];
        if (gb.gene_begin != at || gb.gene_end < gb.gene_begin || gb.gene_end > engine->G ||
            (gb.gene_begin % engine->Gp) != 0 || (gb.gene_end > gb.gene_begin && (!gb.base || gb.ld < gb.gene_end - gb.gene_begin))) {
            set_error("gene blocks must tile [0, G) in order, start on multiples of genes_per_part and have ld >= their width");
            return SPADE_ERR_ARG;
        }
        at = gb.gene_end;
    }
    if (at != engine->G) { set_error("gene blocks must cover all G genes"); return SPADE_ERR_ARG; }
    return run_impl(engine, members, members_location, offsets, S, nullptr, nullptr, nullptr, nullptr, nullptr,
                    SPADE_DEVICE, engine->G, stream, (unsigned long long *)blocks[0].base /* marks the raw form */,
                    blocks, nblocks, first_row, row_step);
}

int spade_expand(spade_engine *engine, const uint64_t *raw, int64_t rows, int64_t raw_row_stride,
                 const int64_t *set_sizes, double *up, double *down, double *fc, double *cpm, int32_t *k,
                 int64_t out_row_stride, void *stream)
{
    if (!engine) return SPADE_ERR_ARG;
    return spade_expand_block(engine, raw, rows, raw_row_stride, 0, engine->G, set_sizes, up, down, fc, cpm, k,
                              out_row_stride, stream);
}

int spade_expand_block(spade_engine *engine, const uint64_t *raw, int64_t rows, int64_t raw_row_stride,
                       int64_t gene_begin, int64_t gene_count, const int64_t *set_sizes, double *up,
                       double *down, double *fc, double *cpm, int32_t *k, int64_t out_row_stride, void *stream)
{
    if (!engine) return SPADE_ERR_ARG;
    CUDA_TRY(cudaSetDevice(engine->device));
    if (rows < 0 || !set_sizes || (rows > 0 && gene_count > 0 && !raw)) { set_error("bad raw / rows / set_sizes"); return SPADE_ERR_ARG; }
    if (gene_begin < 0 || gene_count < 0 || gene_begin + gene_count > engine->G) { set_error("gene block out of range"); return SPADE_ERR_ARG; }
    if (raw_row_stride < gene_count || out_row_stride < gene_count) { set_error("row strides must be >= the gene count of the block"); return SPADE_ERR_ARG; }
    if (rows == 0 || gene_count == 0) return SPADE_OK;
    if (!engine->h_wide.empty()) {
        set_error("spade_expand is not available: the matrix holds genes that need the float64 side path");
        return SPADE_ERR_STATE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t G = engine->G;
    bool uniform = true;
    for (int64_t r = 0; r < rows; ++r) {
        if (set_sizes[r] < 0 || set_sizes[r] > engine->C) { set_error("set size out of range"); return SPADE_ERR_ARG; }
        uniform &= set_sizes[r] == set_sizes[0];
    }
    CUDA_TRY(engine->sizes.reserve(rows));
    CUDA_TRY(engine->h_sizes.reserve(rows));
    memcpy(engine->h_sizes.p, set_sizes, rows * sizeof(int64_t));
    CUDA_TRY(engine->h_sizes.upload(engine->sizes.p, (size_t)rows, st));
    bool table = false;
    if (uniform && set_sizes[0] > 0 && (engine->tail_mode == 2 || (engine->tail_mode == 1 && rows >= 32))) {
        if (engine->tab_N == set_sizes[0] && engine->tab_epoch == engine->bg_epoch) {
            table = true;      // the table the last spade_run / spade_expand built for this N is still valid
        } else {
            int rc = engine->build_tail_table(set_sizes[0], st, &table);
            if (rc != SPADE_OK) return rc;
        }
    }
    FinCtx f;
    f.G = G;
    f.M = (int)engine->C;
    f.nc_mode = engine->n_nc > 0;
    f.meta = engine->meta.p;
    f.tmeta = table ? engine->tmeta_buf[engine->tab_cur].p : nullptr;
    f.tab = table ? engine->tab_buf[engine->tab_cur].p : nullptr;
    f.lf = engine->lf.p;
    f.up = up; f.down = down; f.fc = fc; f.cpm = cpm; f.k = k;
    f.ld = out_row_stride;
    f.k_ld = 0;
    // ragged rows (no tail table): leave the counts, evaluate the tails at full occupancy afterwards
    const bool defer = !table && (up || down);
    const int32_t *t_k = k;
    int64_t t_kld = out_row_stride;
    if (defer) {
        if (!k) {
            CUDA_TRY(engine->kbuf.reserve((size_t)rows * (size_t)gene_count));
            f.k = engine->kbuf.p;
            f.k_ld = gene_count;
            t_k = engine->kbuf.p;
            t_kld = gene_count;
        }
        f.up = f.down = nullptr;
    }
    int rc = engine->span_begin(st, 2);
    if (rc != SPADE_OK) return rc;
    const int64_t rows_per_launch = (int64_t)65535 * kExpandRows;        // gridDim.y limit
    for (int64_t r0 = 0; r0 < rows; r0 += rows_per_launch) {
        const int64_t nr = std::min(rows_per_launch, rows - r0);
        FinCtx fr = f;
        if (fr.up) fr.up += r0 * out_row_stride;
        if (fr.down) fr.down += r0 * out_row_stride;
        if (fr.fc) fr.fc += r0 * out_row_stride;
        if (fr.cpm) fr.cpm += r0 * out_row_stride;
        if (fr.k) fr.k += r0 * out_row_stride;
        const dim3 grid((unsigned)((gene_count + 255) / 256), (unsigned)((nr + kExpandRows - 1) / kExpandRows));
        k_expand<<<grid, 256, 0, st>>>(fr, (const unsigned long long *)raw + r0 * raw_row_stride, raw_row_stride,
                                       engine->sizes.p + r0, nullptr, nr, gene_begin, gene_count);
        CUDA_TRY(cudaGetLastError());
    }
    rc = engine->span_end(st);
    if (rc != SPADE_OK) return rc;
    engine->launches += 1;
    if (defer) {
        const int32_t *by_size = nullptr;
        if (rows <= 0x7fffffff) {                          // rows by decreasing set size (see k_tails_direct)
            CUDA_TRY(engine->h_order_x.reserve((size_t)rows));
            CUDA_TRY(engine->order_x.reserve((size_t)rows));
            int32_t *o = engine->h_order_x.p;
            for (int64_t i = 0; i < rows; ++i) o[i] = (int32_t)i;
            std::stable_sort(o, o + rows, [&](int32_t x, int32_t y) { return set_sizes[x] > set_sizes[y]; });
            CUDA_TRY(engine->h_order_x.upload(engine->order_x.p, (size_t)rows, st));
            by_size = engine->order_x.p;
        }
        rc = engine->launch_tails(st, t_k, t_kld, engine->sizes.p, by_size, 0, rows, gene_begin,
                                  gene_begin + gene_count, gene_begin, up, down, out_row_stride);
        if (rc != SPADE_OK) return rc;
    }
    return SPADE_OK;
}

int spade_check(spade_engine *engine)
{
    if (!engine) return SPADE_ERR_ARG;
    CUDA_TRY(cudaSetDevice(engine->device));
    CUDA_TRY(cudaDeviceSynchronize());
    return engine->check_device_error("spade_run");
}

int spade_wide_genes(const spade_engine *engine, int64_t *count)
{
    if (!engine || !count) return SPADE_ERR_ARG;
    *count = (int64_t)engine->h_wide.size();
    return SPADE_OK;
}

int spade_timing(spade_engine *engine, double *test_ms, double *table_ms, int64_t *launches, int reset)
{
    if (!engine) return SPADE_ERR_ARG;
    CUDA_TRY(cudaSetDevice(engine->device));
    CUDA_TRY(cudaDeviceSynchronize());
    int rc = engine->collect_timing();
    if (rc != SPADE_OK) return rc;
    rc = engine->check_device_error("spade_run");
    if (test_ms) *test_ms = engine->test_ms;
    if (table_ms) *table_ms = engine->table_ms;
    if (launches) *launches = engine->launches;
    if (reset) {
        engine->test_ms = engine->table_ms = engine->finish_ms = 0.0;
        engine->launches = 0;
    }
    return rc;
}

int spade_timing_finish(spade_engine *engine, double *finish_ms)
{
    if (!engine || !finish_ms) return SPADE_ERR_ARG;
    CUDA_TRY(cudaSetDevice(engine->device));
    CUDA_TRY(cudaDeviceSynchronize());
    int rc = engine->collect_timing();
    if (rc != SPADE_OK) return rc;
    *finish_ms = engine->finish_ms;
    return SPADE_OK;
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
        if (M[i] > (int64_t(1) << 27)) return fail(SPADE_ERR_ARG, "M too large for the probe (limit 2^27)");
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

#define IPC_TRY(expr)                                                                          \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            g_create_error = std::string(#expr) + ": " + cudaGetErrorString(_e);               \
            return SPADE_ERR_CUDA;                                                             \
        }                                                                                      \
    } while (0)

int spade_device_alloc(int device, void **dptr, int64_t bytes)
{
    if (!dptr || bytes < 0) return SPADE_ERR_ARG;
    IPC_TRY(cudaSetDevice(device));
    IPC_TRY(cudaMalloc(dptr, (size_t)std::max<int64_t>(bytes, 1)));
    return SPADE_OK;
}

int spade_device_free(int device, void *dptr)
{
    if (!dptr) return SPADE_OK;
    IPC_TRY(cudaSetDevice(device));
    IPC_TRY(cudaFree(dptr));
    return SPADE_OK;
}

int spade_ipc_export(int device, void *dptr, unsigned char handle[64])
{
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    if (!dptr || !handle) return SPADE_ERR_ARG;
    IPC_TRY(cudaSetDevice(device));
    cudaIpcMemHandle_t h;
    IPC_TRY(cudaIpcGetMemHandle(&h, dptr));
    memcpy(handle, &h, 64);
    return SPADE_OK;
}

int spade_ipc_open(int device, const unsigned char handle[64], void **dptr)
{
    if (!dptr || !handle) return SPADE_ERR_ARG;
    IPC_TRY(cudaSetDevice(device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    IPC_TRY(cudaIpcOpenMemHandle(dptr, h, cudaIpcMemLazyEnablePeerAccess));
    return SPADE_OK;
}

int spade_ipc_close(int device, void *dptr)
{
    if (!dptr) return SPADE_OK;
    IPC_TRY(cudaSetDevice(device));
    IPC_TRY(cudaIpcCloseMemHandle(dptr));
    return SPADE_OK;
}
#undef IPC_TRY

namespace {
// One warp per gene, kGammaGenes genes per CTA.  The CTA first copies its [S][kGammaGenes] tile
// of the table into shared memory (32-byte runs per row: whole sectors), so that the ~300
// likelihood evaluations of a gene read its column from shared memory, lane-strided and
// conflict-free; tables with too many rows for that are read from global memory directly.
constexpr int kGammaGenes = 4;

__global__ void __launch_bounds__(kGammaGenes * 32)
k_gamma_fit(const double *__restrict__ tab, int S, int64_t G, int64_t ld, double sign, double *__restrict__ A,
            double *__restrict__ B, double *__restrict__ C, int32_t *__restrict__ status,
            int32_t *__restrict__ evals, int use_smem)
{
    extern __shared__ double cols[];                       // [kGammaGenes][S]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t g0 = (int64_t)blockIdx.x * kGammaGenes;
    if (use_smem) {
        for (int idx = threadIdx.x; idx < S * kGammaGenes; idx += blockDim.x) {
            const int r = idx / kGammaGenes, j = idx % kGammaGenes;
            cols[(size_t)j * S + r] = g0 + j < G ? tab[(int64_t)r * ld + g0 + j] : NAN;
        }
        __syncthreads();
    }
    const int64_t g = g0 + warp;
    if (g >= G) return;
    GammaColumn col;
    col.tab = use_smem ? cols + (size_t)warp * S : tab + g;
    col.ld = use_smem ? 1 : ld;
    col.S = S;
    col.sign = sign;
    col.lane = lane;
    col.nlanes = 32;
    double out[3];
    int ne = 0;
    const int st = gamma_fit_column(col, out, WarpReduce(), &ne);
    if (lane == 0) {
        status[g] = st;
        evals[g] = ne;
        A[g] = out[0];
        B[g] = out[1];
        C[g] = out[2];
    }
}
}  // namespace

int spade_gamma_fit(int device, const double *table, int table_location, int64_t S, int64_t G, int64_t ld,
                    double sign, double *A, double *B, double *C, int32_t *status, int32_t *evals)
{
    auto fail = [&](int code, const std::string &m) { g_create_error = m; return code; };
    if (S < 0 || G < 0 || ld < G || S > 0x7fffffff || (S > 0 && G > 0 && !table) || (G > 0 && (!A || !B || !C)))
        return fail(SPADE_ERR_ARG, "bad table / output arguments");
    if (table_location != SPADE_HOST && table_location != SPADE_DEVICE) return fail(SPADE_ERR_ARG, "bad location");
    if (G == 0) return SPADE_OK;
    cudaError_t ce = cudaSetDevice(device);
    if (ce != cudaSuccess) return fail(SPADE_ERR_CUDA, cudaGetErrorString(ce));
    DevBuf<double> dtab, dout;
    DevBuf<int32_t> dst;
#define G_TRY(expr)                                                                            \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) return fail(SPADE_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    } while (0)
    const double *dt = table;
    if (table_location == SPADE_HOST && S > 0) {
        G_TRY(dtab.reserve((size_t)S * (size_t)G));
        G_TRY(cudaMemcpy2D(dtab.p, G * sizeof(double), table, ld * sizeof(double), G * sizeof(double), S,
                           cudaMemcpyHostToDevice));
        dt = dtab.p;
        ld = G;
    }
    G_TRY(dout.reserve(3 * (size_t)G));
    G_TRY(dst.reserve(2 * (size_t)G));
    const size_t tile = (size_t)kGammaGenes * (size_t)S * sizeof(double);
    const int use_smem = tile <= 96 * 1024;
    if (use_smem) G_TRY(cudaFuncSetAttribute(k_gamma_fit, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    k_gamma_fit<<<(unsigned)((G + kGammaGenes - 1) / kGammaGenes), kGammaGenes * 32, use_smem ? tile : 0>>>(
        dt, (int)S, G, ld, sign, dout.p, dout.p + G, dout.p + 2 * G, dst.p, dst.p + G, use_smem);
    G_TRY(cudaGetLastError());
    G_TRY(cudaMemcpy(A, dout.p, G * sizeof(double), cudaMemcpyDeviceToHost));
    G_TRY(cudaMemcpy(B, dout.p + G, G * sizeof(double), cudaMemcpyDeviceToHost));
    G_TRY(cudaMemcpy(C, dout.p + 2 * G, G * sizeof(double), cudaMemcpyDeviceToHost));
    if (status) G_TRY(cudaMemcpy(status, dst.p, G * sizeof(int32_t), cudaMemcpyDeviceToHost));
    if (evals) G_TRY(cudaMemcpy(evals, dst.p + G, G * sizeof(int32_t), cudaMemcpyDeviceToHost));
#undef G_TRY
    return SPADE_OK;
}

int spade_score(int device, const double *rand, int rand_location, int64_t S, int64_t G, int64_t rand_ld,
                const double *obs, int obs_location, int64_t R, int64_t obs_ld, const double *A, const double *B,
                const double *C, double *score, double *emp, int out_location, int64_t out_ld)
{
    auto fail = [&](int code, const std::string &m) { g_create_error = m; return code; };
    const bool want_score = score && A && B && C, want_emp = emp && rand;
    if (S < 0 || G < 0 || R < 0 || obs_ld < G || out_ld < G || (want_emp && rand_ld < G))
        return fail(SPADE_ERR_ARG, "bad sizes / strides");
    if ((score && !(A && B && C)) || (emp && !rand)) return fail(SPADE_ERR_ARG, "score needs A, B, C; emp needs the rand table");
    if (want_emp && (S < 1 || S > kScoreMaxDraws)) return fail(SPADE_ERR_ARG, "the rand table must hold 1 .. 16384 draws");
    for (int loc : {rand_location, obs_location, out_location})
        if (loc != SPADE_HOST && loc != SPADE_DEVICE) return fail(SPADE_ERR_ARG, "bad location");
    if (R == 0 || G == 0 || (!want_score && !want_emp)) return SPADE_OK;
    if (!obs) return fail(SPADE_ERR_ARG, "obs is NULL");
    cudaError_t ce = cudaSetDevice(device);
    if (ce != cudaSuccess) return fail(SPADE_ERR_CUDA, cudaGetErrorString(ce));
#define S_TRY(expr)                                                                            \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) return fail(SPADE_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    } while (0)
    DevBuf<double> drand, dobs, dabc, dsorted, dscore, demp;
    DevBuf<int32_t> dvalid;
    ScoreArgs a;
    memset(&a, 0, sizeof(a));
    a.R = R; a.G = G; a.S = (int)S;
    if (want_emp) {
        const double *dr = rand;
        int64_t ld = rand_ld;
        if (rand_location == SPADE_HOST) {
            S_TRY(drand.reserve((size_t)S * (size_t)G));
            S_TRY(cudaMemcpy2D(drand.p, G * sizeof(double), rand, rand_ld * sizeof(double), G * sizeof(double), S,
                               cudaMemcpyHostToDevice));
            dr = drand.p;
            ld = G;
        }
        S_TRY(dsorted.reserve((size_t)S * (size_t)G));
        S_TRY(dvalid.reserve((size_t)G));
        int m = 32;
        while (m < S) m <<= 1;
        const size_t smem = (size_t)m * sizeof(double);
        S_TRY(cudaFuncSetAttribute(k_sort_background, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_sort_background<<<(unsigned)G, 256, smem>>>(dr, (int)S, ld, dsorted.p, dvalid.p);
        S_TRY(cudaGetLastError());
        a.sorted = dsorted.p;
        a.valid = dvalid.p;
    }
    a.obs = obs;
    a.obs_ld = obs_ld;
    if (obs_location == SPADE_HOST) {
        S_TRY(dobs.reserve((size_t)R * (size_t)G));
        S_TRY(cudaMemcpy2D(dobs.p, G * sizeof(double), obs, obs_ld * sizeof(double), G * sizeof(double), R,
                           cudaMemcpyHostToDevice));
        a.obs = dobs.p;
        a.obs_ld = G;
    }
    if (want_score) {                                      // A, B, C: host vectors
        S_TRY(dabc.reserve(3 * (size_t)G));
        S_TRY(cudaMemcpy(dabc.p, A, G * sizeof(double), cudaMemcpyHostToDevice));
        S_TRY(cudaMemcpy(dabc.p + G, B, G * sizeof(double), cudaMemcpyHostToDevice));
        S_TRY(cudaMemcpy(dabc.p + 2 * G, C, G * sizeof(double), cudaMemcpyHostToDevice));
        a.A = dabc.p; a.B = dabc.p + G; a.C = dabc.p + 2 * G;
    }
    a.score = want_score ? score : nullptr;
    a.emp = want_emp ? emp : nullptr;
    a.out_ld = out_ld;
    if (out_location == SPADE_HOST) {
        if (want_score) { S_TRY(dscore.reserve((size_t)R * (size_t)G)); a.score = dscore.p; }
        if (want_emp) { S_TRY(demp.reserve((size_t)R * (size_t)G)); a.emp = demp.p; }
        a.out_ld = G;
    }
    const dim3 grid((unsigned)((G + 255) / 256), (unsigned)std::min<int64_t>(R, 4096));
    k_score<<<grid, 256>>>(a);
    S_TRY(cudaGetLastError());
    if (out_location == SPADE_HOST) {
        if (want_score)
            S_TRY(cudaMemcpy2D(score, out_ld * sizeof(double), dscore.p, G * sizeof(double), G * sizeof(double), R,
                               cudaMemcpyDeviceToHost));
        if (want_emp)
            S_TRY(cudaMemcpy2D(emp, out_ld * sizeof(double), demp.p, G * sizeof(double), G * sizeof(double), R,
                               cudaMemcpyDeviceToHost));
    }
    S_TRY(cudaDeviceSynchronize());
#undef S_TRY
    return SPADE_OK;
}

int spade_csc_count(int device, const double *table, int64_t S, int64_t G, int64_t ld, int32_t *jc)
{
    auto fail = [&](int code, const std::string &m) { g_create_error = m; return code; };
    if (S < 0 || G < 0 || ld < G || !jc || (S > 0 && G > 0 && !table) || S > 0x7fffffff)
        return fail(SPADE_ERR_ARG, "bad table / jc arguments");
    jc[0] = 0;
    if (G == 0) return SPADE_OK;
    cudaError_t ce = cudaSetDevice(device);
    if (ce != cudaSuccess) return fail(SPADE_ERR_CUDA, cudaGetErrorString(ce));
    DevBuf<int32_t> cnt;
    ce = cnt.reserve((size_t)G);
    if (ce != cudaSuccess) return fail(SPADE_ERR_CUDA, cudaGetErrorString(ce));
    k_csc_count<<<(unsigned)((G * 32 + 255) / 256), 256>>>(table, S, G, ld, cnt.p);
    ce = cudaMemcpy(jc + 1, cnt.p, G * sizeof(int32_t), cudaMemcpyDeviceToHost);
    if (ce != cudaSuccess) return fail(SPADE_ERR_CUDA, cudaGetErrorString(ce));
    int64_t run = 0;
    for (int64_t g = 1; g <= G; ++g) {
        run += jc[g];
        if (run > 0x7fffffff) return fail(SPADE_ERR_ARG, "more than 2^31 - 1 stored entries: too large for a MAT-v5 sparse array");
        jc[g] = (int32_t)run;
    }
    return SPADE_OK;
}

int spade_csc_fill(int device, const double *table, int64_t S, int64_t G, int64_t ld, const int32_t *jc,
                   int32_t *ir, double *pr, int out_location)
{
    auto fail = [&](int code, const std::string &m) { g_create_error = m; return code; };
    if (S < 0 || G < 0 || ld < G || !jc || (S > 0 && G > 0 && !table))
        return fail(SPADE_ERR_ARG, "bad table / jc arguments");
    if (out_location != SPADE_HOST && out_location != SPADE_DEVICE) return fail(SPADE_ERR_ARG, "bad location");
    const int64_t nnz = G > 0 ? jc[G] : 0;
    if (nnz == 0) return SPADE_OK;
    if (!ir || !pr) return fail(SPADE_ERR_ARG, "ir / pr is NULL");
    cudaError_t ce = cudaSetDevice(device);
    if (ce != cudaSuccess) return fail(SPADE_ERR_CUDA, cudaGetErrorString(ce));
#define P_TRY(expr)                                                                            \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) return fail(SPADE_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    } while (0)
    DevBuf<int64_t> begin;
    DevBuf<int32_t> dir;
    DevBuf<double> dpr;
    std::vector<int64_t> hb((size_t)G);
    for (int64_t g = 0; g < G; ++g) hb[g] = jc[g] - jc[0];
    P_TRY(begin.reserve((size_t)G));
    P_TRY(cudaMemcpy(begin.p, hb.data(), G * sizeof(int64_t), cudaMemcpyHostToDevice));
    int32_t *oir = ir;
    double *opr = pr;
    if (out_location == SPADE_HOST) {
        P_TRY(dir.reserve((size_t)nnz));
        P_TRY(dpr.reserve((size_t)nnz));
        oir = dir.p;
        opr = dpr.p;
    }
    k_csc_fill<<<(unsigned)((G * 32 + 255) / 256), 256>>>(table, S, G, ld, begin.p, oir, opr);
    P_TRY(cudaGetLastError());
    if (out_location == SPADE_HOST) {
        P_TRY(cudaMemcpy(ir, dir.p, nnz * sizeof(int32_t), cudaMemcpyDeviceToHost));
        P_TRY(cudaMemcpy(pr, dpr.p, nnz * sizeof(double), cudaMemcpyDeviceToHost));
    }
    P_TRY(cudaDeviceSynchronize());
    return SPADE_OK;
}

int spade_column_mean(int device, const double *table, int64_t S, int64_t G, int64_t ld, double *mean)
{
    auto fail = [&](int code, const std::string &m) { g_create_error = m; return code; };
    if (S < 0 || G < 0 || ld < G || (G > 0 && !mean) || (S > 0 && G > 0 && !table))
        return fail(SPADE_ERR_ARG, "bad table / mean arguments");
    if (G == 0) return SPADE_OK;
    cudaError_t ce = cudaSetDevice(device);
    if (ce != cudaSuccess) return fail(SPADE_ERR_CUDA, cudaGetErrorString(ce));
    DevBuf<double> dm;
    P_TRY(dm.reserve((size_t)G));
    k_column_mean<<<(unsigned)((G * 32 + 255) / 256), 256>>>(table, S, G, ld, dm.p);
    P_TRY(cudaGetLastError());
    P_TRY(cudaMemcpy(mean, dm.p, G * sizeof(double), cudaMemcpyDeviceToHost));
#undef P_TRY
    return SPADE_OK;
}

int spade_memcpy2d(void *dst, int64_t dst_pitch, const void *src, int64_t src_pitch, int64_t width_bytes,
                   int64_t rows, int device_to_host, void *stream)
{
    if (!dst || !src || width_bytes < 0 || rows < 0 || dst_pitch < width_bytes || src_pitch < width_bytes)
        return SPADE_ERR_ARG;
    if (width_bytes == 0 || rows == 0) return SPADE_OK;
    cudaError_t e = cudaMemcpy2DAsync(dst, (size_t)dst_pitch, src, (size_t)src_pitch, (size_t)width_bytes, (size_t)rows,
                                      device_to_host ? cudaMemcpyDeviceToHost : cudaMemcpyHostToDevice,
                                      (cudaStream_t)stream);
    if (e != cudaSuccess) {
        g_create_error = std::string("cudaMemcpy2DAsync: ") + cudaGetErrorString(e);
        return SPADE_ERR_CUDA;
    }
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

int spade_host_register(void *ptr, int64_t bytes)
{
    if (!ptr || bytes <= 0) return SPADE_ERR_ARG;
    cudaError_t e = cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterPortable);
    if (e != cudaSuccess) {
        g_create_error = std::string("cudaHostRegister: ") + cudaGetErrorString(e);
        return SPADE_ERR_CUDA;
    }
    return SPADE_OK;
}

int spade_host_unregister(void *ptr)
{
    if (!ptr) return SPADE_OK;
    return cudaHostUnregister(ptr) == cudaSuccess ? SPADE_OK : SPADE_ERR_CUDA;
}

}  // extern "C"
