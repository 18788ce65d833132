/*
 * libspade_b200 - C ABI of the B200 engine for pySpade's hypergeometric
 * differential-expression hot path.
 *
 * The reference (Hon-lab/pySpade v0.1.7) is pure Python and has no FFI; the seam
 * this library sits behind is the pair of functions
 *     perform_DE(sgrna_idx, input_array, idx, num_processes, down, up, fc, cpm)
 *         pySpade/utils.py:296-317
 *     perform_DE_NC(sgrna_idx, NC_idx, input_array, idx, num_processes, ...)
 *         pySpade/utils.py:319-341
 * called once per perturbation region by DEobs.py:137,197 and once per random
 * draw by DErand.py:143,207.  Each entry point below names the reference lines
 * it replaces.  INTEGRATION.md shows the ctypes binding a maintainer would add.
 *
 * Conventions
 *   - every function returns an int status, 0 = SPADE_OK; nothing throws across
 *     the ABI; spade_last_error() gives the message of the last failure;
 *   - input buffers are borrowed for the duration of the call only;
 *   - outputs are caller-allocated; "location" arguments say whether a pointer
 *     is host (SPADE_HOST) or device (SPADE_DEVICE) memory;
 *   - one engine per GPU; an engine is not thread-safe; engines are independent;
 *   - no CPU fallback: without a CUDA device every call fails with SPADE_ERR_CUDA.
 *
 * Vocabulary: G genes, C cells, a "set" is the list of member cells of one
 * perturbation region (DEobs) or one random draw (DErand); a "test" is one
 * (gene, set) pair.
 */
#ifndef SPADE_B200_H
#define SPADE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPADE_OK 0
#define SPADE_ERR_ARG 1      /* bad argument (null pointer, index out of range, ...) */
#define SPADE_ERR_CUDA 2     /* CUDA runtime failure, no device, out of memory */
#define SPADE_ERR_STATE 3    /* call not valid in the engine's current state */

#define SPADE_HOST 0
#define SPADE_DEVICE 1

#define SPADE_ABI_VERSION 5

typedef struct spade_engine spade_engine;

/* ABI version of the loaded library (compare with SPADE_ABI_VERSION). */
int spade_abi_version(void);

/* Message of the last failure on this engine (or of the last failed
 * spade_create* when engine is NULL).  Never NULL; valid until the next call. */
const char *spade_last_error(const spade_engine *engine);

/*
 * Stage a genes x cells matrix of already-normalised float64 expression values
 * (CSR: indptr[G+1], indices[nnz] = cell index, values[nnz]) on GPU `device`.
 * Replaces what the reference redoes on every call: input_array.tocsr() and the
 * per-row pickling (utils.py:303-307, 325-331), np.median over the dense row and
 * the low-set selection (utils.py:199-200).  One-time work per matrix:
 * per-gene median cutoff, n_g = #{cells <= median}, per-entry "above cutoff"
 * flags, row sums, and the gene-partitioned cell-major copy the batched test walks.
 * Column indices inside a row must be unique; they need not be sorted.
 */
int spade_create(int device, int64_t G, int64_t C, const int64_t *indptr, const int32_t *indices,
                 const double *values, int location, spade_engine **engine_out);

/*
 * Same, from raw integer UMI counts: runs the reference's CPM normalisation
 * (cpm_normalization, utils.py:150-155; value = fl(fl(x*1e6) * fl(1/colsum)),
 * bit-identical to the reference) on the device first.
 */
int spade_create_from_counts(int device, int64_t G, int64_t C, const int64_t *indptr,
                             const int32_t *indices, const int32_t *counts, int location,
                             spade_engine **engine_out);

void spade_destroy(spade_engine *engine);

/*
 * Choose the background.  n_nc == 0: complement background (hypergeo_test_sparse,
 * utils.py:194-215).  n_nc > 0: negative-control background
 * (hypergeo_test_NC_sparse, utils.py:248-266): median over the NC cells only,
 * low set over all cells, fold-change denominator = mean over the NC cells.
 * nc_idx is a host array of unique cell indices.  Re-stages cutoffs and flags.
 */
int spade_set_background(spade_engine *engine, const int32_t *nc_idx, int64_t n_nc);

/*
 * Tuning / test knobs (defaults are right for production use):
 *   SPADE_OPT_TAIL_TABLE      0 = evaluate every log-hypergeometric tail directly,
 *                             1 = (default) per-run tail table when a batch of >= 32 sets
 *                                 shares one set size N (DErand), 2 = table whenever a batch
 *                                 shares one N.  Table entries are the direct evaluation
 *                                 itself, so all three settings give bit-identical tables.
 *   SPADE_OPT_MEMBER_BLOCK    member cells one warp accumulates before draining (1..32768,
 *                             default 32768); sets larger than this take the split path.
 *   SPADE_OPT_GENES_PER_PART  genes per partition of the cell-major layout (multiple of 32,
 *                             <= 4096, 0 = default 1024); re-stages the matrix.
 *   SPADE_OPT_SORT_MEMBERS    1 = (default) visit the cells of every set in ascending order
 *                             (device-side segmented sort of a copy of the member lists;
 *                             better L2 locality), 0 = keep the caller's order.  Same tables.
 */
#define SPADE_OPT_TAIL_TABLE 1
#define SPADE_OPT_MEMBER_BLOCK 2
#define SPADE_OPT_GENES_PER_PART 3
#define SPADE_OPT_SORT_MEMBERS 4
int spade_set_option(spade_engine *engine, int option, int64_t value);

/* Matrix facts: G, C, stored non-zeros, genes per partition of the staged layout. */
int spade_dims(const spade_engine *engine, int64_t *G, int64_t *C, int64_t *nnz, int32_t *genes_per_part);

/* Per-gene cutoffs for integer parity checks (host outputs, either may be NULL):
 * median[G] (float64) and n[G] = #{cells : x <= median} (int32). */
int spade_gene_cutoffs(spade_engine *engine, double *median, int32_t *n);

/* Copy the staged CPM values back in the caller's CSR order (host output,
 * float64[nnz]); parity check of spade_create_from_counts against utils.py:150-155. */
int spade_cpm_values(spade_engine *engine, double *values);

/*
 * The batched test: every gene against S cell sets.  Replaces S calls of
 * perform_DE / perform_DE_NC (utils.py:296-341), i.e. G*S calls of
 * hypergeo_test[_NC]_sparse (utils.py:194-215 / 248-266).
 *
 *   members   int32[offsets[S]]  concatenated member cell indices (no duplicates
 *                                inside a set), host or device per members_location
 *   offsets   int64[S+1]         HOST array, offsets[0] == 0, non-decreasing
 *   up, down, fc, cpm            float64[S*G], row-major [set][gene]; any may be NULL
 *                                up   = hypergeom.logcdf(k, M, n, N)   (utils.py:212)
 *                                down = hypergeom.logsf (k, M, n, N)   (utils.py:213)
 *                                both nan when any of k, M, n, N is 0
 *                                fc   = (mean_set+.01)/(mean_background+.01) (utils.py:207 / :258)
 *                                cpm  = mean over the set's cells     (utils.py:208 / :259)
 *   k         int32[S*G]         optional: |{cells <= cutoff} ∩ set|  (utils.py:203,211)
 *   out_location                 SPADE_HOST or SPADE_DEVICE for all five outputs
 *   stream                       cudaStream_t (as void*) the work is ordered on; NULL =
 *                                the legacy default stream.  With device outputs the call
 *                                returns without waiting for the kernels; with host outputs
 *                                it returns when the tables are complete.
 * Sets of any size are accepted (sets above SPADE_OPT_MEMBER_BLOCK cells take a slower
 * split path).  Device workspace kept by the engine between calls: the per-run tail table and,
 * for batches of one set size with device outputs, 8 bytes per test of the largest chunk
 * (at most 2^28 tests = 2 GB) for the accumulator words between the two launches of the test.
 */
int spade_run(spade_engine *engine, const int32_t *members, int members_location,
              const int64_t *offsets, int64_t S, double *up, double *down, double *fc, double *cpm,
              int32_t *k, int out_location, void *stream);

/*
 * Verdict on the member lists of every spade_run / spade_run_raw issued so far.  With host
 * outputs spade_run reports a member cell index out of range or a duplicated cell inside a set
 * itself; with device outputs it returns before its kernels have run, so such an error surfaces
 * here (the call synchronises the device) or, once the kernels have finished, at the next
 * spade_run / spade_timing of the engine.  SPADE_ERR_ARG + spade_last_error on a bad list.
 */
int spade_check(spade_engine *engine);

/*
 * Accuracy contract of fc and cpm.  Member sums are carried in a per-gene fixed point
 * (2^-shift_g, chosen so that any sum over 32 768 cells fits 47 bits): the mean of same-signed
 * values is within 2^-22 (2.4e-7) relative of the exact mean, fold changes within 2^-21, for
 * ANY float64 matrix.  A gene whose values span too many binary orders of magnitude for that
 * bound ("wide": a few huge cells next to tiny ones) is summed in float64 by a side kernel
 * instead, like the reference does (scipy sparse mean, utils.py:206-208).  Count data
 * normalised to CPM has no wide genes.  count = number of wide genes of the staged matrix with
 * the current background; spade_run_raw / spade_expand refuse (SPADE_ERR_STATE) when it is > 0.
 * Rows that mix signs can cancel: there the bound is absolute, 2^-22 * min |nonzero value|.
 * Values must be finite and below 8e270 in magnitude (else SPADE_ERR_ARG at staging).
 */
int spade_wide_genes(const spade_engine *engine, int64_t *count);

/*
 * spade_run with an explicit row stride of the result tables (elements between the rows of
 * consecutive sets; spade_run uses G).  With out_row_stride = P*G and table pointers offset
 * by rank*G, P ranks interleave their rows (set i of rank r = row i*P + r) in ONE table:
 *   - device outputs: a table that lives in another GPU's memory and was opened with
 *     spade_ipc_open - the test kernel then stores its results straight into the destination
 *     GPU over NVLink while it computes (gather fused into the kernel, no separate collective);
 *   - host outputs: a table in host memory shared by the ranks of the node (one process per
 *     GPU, e.g. a POSIX shared-memory mapping registered with spade_host_register) - every
 *     GPU delivers its rows over its own PCIe link.
 */
int spade_run_strided(spade_engine *engine, const int32_t *members, int members_location,
                      const int64_t *offsets, int64_t S, double *up, double *down, double *fc,
                      double *cpm, int32_t *k, int out_location, int64_t out_row_stride,
                      void *stream);

/*
 * Raw form of the batched test, for a result that has to cross a link: instead of the
 * finished tables, store the 64-bit accumulator word of every test (member count << 48 | member
 * sum in the gene's fixed point; 8 bytes per test instead of 24-36) into raw[S][out_row_stride]
 * (device memory, possibly a peer GPU's).  spade_expand, on an engine staged from the same
 * matrix with the same background, turns rows of such words into the tables - bit-identical to
 * what spade_run would have written.  set_sizes: host int64[rows], N of every row's set.
 * spade_expand may run on another stream than the spade_run calls of the same engine: it
 * reuses the tail table the last spade_run built when that fits its set size, and that
 * table is not overwritten by the next spade_run (two tables are used in turn) - it must only
 * have finished before the spade_run after the next one starts.
 * Sets above SPADE_OPT_MEMBER_BLOCK cells are refused by the raw form.
 */
int spade_run_raw(spade_engine *engine, const int32_t *members, int members_location,
                  const int64_t *offsets, int64_t S, uint64_t *raw, int64_t out_row_stride,
                  void *stream);
int spade_expand(spade_engine *engine, const uint64_t *raw, int64_t rows, int64_t raw_row_stride,
                 const int64_t *set_sizes, double *up, double *down, double *fc, double *cpm,
                 int32_t *k, int64_t out_row_stride, void *stream);

/*
 * Raw form routed by GENE BLOCK - the exchange step of the multi-GPU path (one process per
 * GPU, sets sharded over the GPUs, matrix replicated).  The genes are cut into `nblocks`
 * consecutive blocks (boundaries on multiples of genes_per_part, spade_dims), block b owned by
 * GPU b; a GPU that has tested its share of the sets stores the word of (set i, gene g) straight
 * into the owner's table, blocks[b].base[(first_row + i * row_step) * blocks[b].ld + (g -
 * gene_begin)] - its own memory for its own block, a peer's (spade_ipc_open) over NVLink /
 * NVSwitch for the others, from inside the test kernel while it computes.  After all GPUs have
 * finished (any barrier ordered after the launch stream), every GPU holds ALL sets x ITS genes
 * and finishes them with spade_expand_block: the all-to-all is fused into the kernel, every GPU
 * receives only (P-1)/P of 1/P of the words, and the layout - all draws of a gene on one GPU -
 * is what the per-gene consumers of the tables need next (gamma fit over the draws
 * DErand.py:250-292, empirical p-values global.py:177-187, the column-compressed MAT tables).
 * first_row / row_step place this GPU's sets in the global set order (DErand: draw i of rank r
 * is row i * P + r; DEobs: a contiguous run of rows per rank).
 * spade_expand_block: rows x gene_count words (column j = gene gene_begin + j) into tables with
 * the same column meaning; set_sizes[rows] on the host.  spade_expand = the block [0, G).
 */
typedef struct spade_gene_block {
    uint64_t *base;      /* device pointer of the owner's word table (own or peer memory) */
    int64_t ld;          /* elements between its rows, >= gene_end - gene_begin */
    int64_t gene_begin;  /* first gene of the block, multiple of genes_per_part */
    int64_t gene_end;    /* one past its last gene */
} spade_gene_block;
int spade_run_raw_blocks(spade_engine *engine, const int32_t *members, int members_location,
                         const int64_t *offsets, int64_t S, const spade_gene_block *blocks, int nblocks,
                         int64_t first_row, int64_t row_step, void *stream);
int spade_expand_block(spade_engine *engine, const uint64_t *raw, int64_t rows, int64_t raw_row_stride,
                       int64_t gene_begin, int64_t gene_count, const int64_t *set_sizes, double *up,
                       double *down, double *fc, double *cpm, int32_t *k, int64_t out_row_stride, void *stream);

/*
 * Wire format of the DErand tables, built on the GPU (pySpade/DErand.py:230-248 saves each
 * table with scipy.io.savemat(csr_matrix(dense)): a MAT-v5 sparse array, column-compressed with
 * column = gene; stored entries are those != 0 - nan and -inf are stored, +-0.0 are not).
 * table: DEVICE float64 [S][ld], first G columns (all draws x a block of genes).
 *   spade_csc_count  jc[G+1] (HOST int32): jc[0] = 0, jc[g+1] = stored entries of columns 0..g
 *   spade_csc_fill   ir[nnz] row (draw) indices, ascending inside a column, and pr[nnz] values,
 *                    nnz = jc[G] - jc[0]; host or device per out_location
 * so that the host only writes headers and these arrays (pyspade_b200/io.py), and the dense
 * tables never cross PCIe.  spade_column_mean: mean over the S rows of every column (host
 * float64[G]), DErand.py:252's Cpm_mean.  All three synchronise.
 */
int spade_csc_count(int device, const double *table, int64_t S, int64_t G, int64_t ld, int32_t *jc);
int spade_csc_fill(int device, const double *table, int64_t S, int64_t G, int64_t ld, const int32_t *jc,
                   int32_t *ir, double *pr, int out_location);
int spade_column_mean(int device, const double *table, int64_t S, int64_t G, int64_t ld, double *mean);

/* Strided copy between device and host memory on a stream (cudaMemcpy2DAsync): a gene block
 * [rows][width] of a device table into / out of a wider host table. */
int spade_memcpy2d(void *dst, int64_t dst_pitch, const void *src, int64_t src_pitch, int64_t width_bytes,
                   int64_t rows, int device_to_host, void *stream);

/*
 * Plain device memory that can be shared between the processes of one node (one process per
 * GPU): the owner allocates and exports a 64-byte handle, peers open it and get a device
 * pointer valid in their own process (peer access over NVLink / NVSwitch).  Thin wrappers of
 * cudaMalloc / cudaIpcGetMemHandle / cudaIpcOpenMemHandle / cudaIpcCloseMemHandle.
 */
int spade_device_alloc(int device, void **dptr, int64_t bytes);
int spade_device_free(int device, void *dptr);
int spade_ipc_export(int device, void *dptr, unsigned char handle[64]);
int spade_ipc_open(int device, const unsigned char handle[64], void **dptr);
int spade_ipc_close(int device, void *dptr);

/* Kernel-time accounting (CUDA events on the launch streams, accumulated since the last
 * reset).  test_ms = the test itself: k_de_fused (member accumulation, and tails + table writes
 * when it finishes its tests itself) plus, for uniform batches with device tables, the k_expand
 * launch that finishes its accumulator words (two-phase form);
 * table_ms = everything else spade_run / spade_expand launch (tail-table builds, direct tails of
 * ragged batches, split-path finalize, expansion of raw words received from other GPUs).
 * Synchronises.  launches = kernels launched by spade_run.
 * spade_timing_finish: the part of test_ms spent in the finishing launch of two-phase runs
 * (same accumulation period; reset by spade_timing). */
int spade_timing(spade_engine *engine, double *test_ms, double *table_ms, int64_t *launches,
                 int reset);
int spade_timing_finish(spade_engine *engine, double *finish_ms);

/*
 * K3 in isolation (parity probe): logcdf / logsf of scipy.stats.hypergeom for `count`
 * tuples (k, M, n, N) given as host int64 arrays, results to host float64 arrays.  Plain
 * scipy semantics (support clamps, nan for invalid shapes); the reference's
 * all([k, M, n, N]) guard of utils.py:212-213 is NOT applied here.  M <= 2^27 (the probe
 * builds a log-factorial table of M + 1 entries).
 */
int spade_hypergeom(int device, const int64_t *k, const int64_t *M, const int64_t *n,
                    const int64_t *N, int64_t count, double *logcdf, double *logsf);

/*
 * Gamma-fit tail of DErand for all genes at once (pySpade/DErand.py:250-292): per gene
 * scipy.stats.gamma.fit of sign * table[:, g] over its finite values - the same start values
 * (gamma_gen._fitstart, rv_continuous._fit_loc_scale_support), the same penalised negative
 * log-likelihood and the same Nelder-Mead moves and stopping rule as scipy.optimize.fmin,
 * restated in csrc/gamma_fit.cuh, one gene per warp.  table: float64 [S][ld] (host or
 * device per table_location), the first G columns are used; sign = -1 for log p tables.
 * Host outputs A, B, C (shape, loc, scale), float64[G], and optionally status int32[G]:
 * 0 fitted, 1 skipped (<= 50 finite values; A = B = C = 0 as the reference leaves them),
 * 2 all values equal (nan, 0, 1), 3 scipy would raise FitError (nan, nan, nan); and optionally
 * evals int32[G]: likelihood evaluations used (600 = the budget of scipy's fmin exhausted).
 * Parameters agree with scipy's to the optimiser's own tolerance (1e-4) wherever its optimum
 * is stable; see tests/test_gamma_fit.py and DESIGN.md section 7 for the measured agreement.
 */
int spade_gamma_fit(int device, const double *table, int table_location, int64_t S, int64_t G,
                    int64_t ld, double sign, double *A, double *B, double *C, int32_t *status,
                    int32_t *evals);

/*
 * Significance scoring of observed regions against the DErand background - the arithmetic of
 * `pySpade global` / `pySpade local` that consumes the tables of this path
 * (pySpade/global.py:107-187, local.py:104-177), for all genes of R regions at once:
 *   p            = obs[r][g] with 0 -> 1, +-inf -> 0, nan -> 0          (global.py:118-123)
 *   score[r][g]  = scipy.stats.gamma.logsf(-p, A[g], B[g], C[g])        (global.py:151-171;
 *                  "Significance_score"; nan where the parameters are invalid, e.g. the
 *                  (0, 0, 0) of a gene DErand skipped or the (nan, 0, 1) of a constant gene)
 *   emp[r][g]    = #{draws s : rand[s][g] < p} / S                       (global.py:181,187;
 *                  "pval-empirical"; nan entries of rand never count, -inf entries always do)
 * rand: the DErand table [S][rand_ld] (up or down log p, as saved), obs: the DEobs rows
 * [R][obs_ld] of the same direction, host or device per *_location; A, B, C: host float64[G]
 * (the Up_/Down_dist_gamma npy vectors).  score / emp: [R][out_ld], host or device per
 * out_location; either may be NULL (then its inputs may be NULL too).  Synchronous.
 * gamma.logsf = log(gammaincc(a, (x - loc) / scale)) with rv_continuous.logsf's support and
 * argument rules; agrees with scipy within 1e-6 relative (+ 1e-15 absolute next to 0, where
 * scipy's own double-rounded sf decides), see tests/test_score_gpu.py.  S <= 16384 draws.
 */
int spade_score(int device, const double *rand, int rand_location, int64_t S, int64_t G, int64_t rand_ld,
                const double *obs, int obs_location, int64_t R, int64_t obs_ld, const double *A, const double *B,
                const double *C, double *score, double *emp, int out_location, int64_t out_ld);

/* Page-locked host memory for result tables (fast device->host copies). */
int spade_host_alloc(void **ptr, int64_t bytes);
int spade_host_free(void *ptr);
/* Page-lock an existing host range (e.g. a shared-memory mapping) for fast copies. */
int spade_host_register(void *ptr, int64_t bytes);
int spade_host_unregister(void *ptr);

#ifdef __cplusplus
}
#endif
#endif /* SPADE_B200_H */
