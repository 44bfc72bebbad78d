/* xpore_b200.h — C ABI of the B200 diffmod hot path (libxpore_b200.so).
 *
 * The reference (GoekeLab/xpore) is pure Python and has no FFI; its seam for this path is
 *   xpore/scripts/diffmod.py:15  execute(idx, data_dict, ...)            one gene
 *   xpore/diffmod/io.py:21       load_data(...) -> {(idx,pos,kmer): {y,x,r,...}}
 *   xpore/diffmod/statstest.py:4 StatsTest(data).fit('t-test')           prefilter
 *   xpore/diffmod/gmm.py:11,93   GMM(method, data, priors, kmer_signal).fit()
 *   xpore/diffmod/io.py:230      generate_result_table(models, data_info)
 * i.e. "positions in, table values out".  This library is that seam for a whole batch of
 * positions at once: plain pointers and sizes, no Python or torch types.  INTEGRATION.md shows
 * the ctypes stub a maintainer of the reference would add inside execute().
 *
 * Conventions
 *  - Ownership: the caller allocates every buffer; the library never frees caller memory.
 *  - Errors: every entry point returns 0 on success, a negative XPB200_E* code otherwise;
 *    xpb200_last_error() (thread-local) describes the failure.  Numerical conditions of single
 *    positions (ELBO decreased, iteration cap) are reported in `status`, never abort the batch
 *    (the reference prints and continues, gmm.py:111-115).
 *  - Threading: calls on different contexts/streams are independent; one batch in flight per
 *    context.
 *  - There is no CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef XPORE_B200_H
#define XPORE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XPB200_OK 0
#define XPB200_EINVAL (-1)   /* bad argument / unsupported shape */
#define XPB200_ECUDA (-2)    /* CUDA runtime error */
#define XPB200_ENOMEM (-3)   /* workspace too small / allocation failed */

/* status bits (per position) */
#define XPB200_SELECTED 1u        /* passed the prefilter gate (diffmod.py:54): position was fitted */
#define XPB200_CONVERGED 2u       /* gmm.py:120-122 */
#define XPB200_ELBO_DECREASED 4u  /* gmm.py:110-115 */
#define XPB200_HIT_MAX 8u         /* ran max_iters-1 sweeps (gmm.py:103) */
#define XPB200_KEEP_ROW 16u       /* survives the row filter io.py:363-367 */
#define XPB200_MOD_IS_1 32u       /* modified cluster = component 1 (io.py:277-282) */
#define XPB200_HIGHER 64u         /* mod_assignment == "higher" (io.py:288) */
#define XPB200_TTEST_VALID 128u   /* exactly two conditions present (statstest.py:20-21) */
#define XPB200_NOT_FITTED 256u    /* selected but NOT fitted (n_iter = -1, record NaN): only when a device-pointer caller
                                     understated batch.max_reads; the host entry points compute max_reads themselves */
#define XPB200_NEAR_THRESHOLD 512u /* t-test p-value within 1e-9 (relative) of the prefilter threshold: the gate
                                     `p < threshold` (diffmod.py:54) could fall either way under a different rounding of
                                     the incomplete beta function; the decision taken is the one of THIS p-value, the bit
                                     lets a caller re-check such positions in higher precision (none seen in 10^7 tests) */

/* A CSR-packed ragged batch of positions (what io.load_data yields for many genes).
 * Reads of a position are contiguous, and within the position contiguous per group slot, in
 * slot order.  A group slot is a run (default) or a condition (method.pooling); a slot with
 * grp_len == 0 is absent at that position (run filtered by readcount, io.py:51-52). */
typedef struct xpb200_batch {
    int32_t n_positions;      /* P */
    int32_t n_groups;         /* G  (1..64) */
    int32_t n_conditions;     /* C  (1..32) */
    int32_t max_reads;        /* max over positions of reads per position */
    int64_t n_reads;          /* R = read_off[P] */
    const void* y;            /* [R + padding]  event means (data.json values), see y_format  */
    const int64_t* read_off;  /* [P+1]                                                       */
    const int32_t* grp_len;   /* [P*G]                                                       */
    const double* kmer_mean;  /* [P] model_kmer.csv model_mean of the position's k-mer       */
    const double* kmer_std;   /* [P] model_kmer.csv model_stdv                               */
    const int32_t* grp_cond;  /* [G] condition index (rank among sorted condition names)     */
    int32_t y_format;         /* XPB200_Y_F64 | XPB200_Y_I32 | XPB200_Y_U16                   */
    int32_t reserved;
    double y_scale;           /* scaled-integer formats: value = (double)q / y_scale, exactly */
} xpb200_batch;

/* Encodings of `y`.  dataprep rounds event means to 2 decimals (dataprep.py:638; 1 decimal in
 * gene mode, :133), so a read is q/100 for an integer q: the packer may ship q as u16 or i32
 * when (double)q / y_scale reproduces every value bit for bit, which quarters / halves the
 * PCIe and HBM bytes.  The buffer must be padded to a multiple of 16 bytes (2 / 4 / 8 elements). */
#define XPB200_Y_F64 0
#define XPB200_Y_I32 1
#define XPB200_Y_U16 2

/* method section of the YAML (configurator.py:46-60) + the fixed priors (:62-77). */
typedef struct xpb200_method {
    int32_t max_iters;           /* default 500; at most max_iters-1 sweeps are run */
    int32_t prefilter;           /* 0 = off, 1 = t-test (statstest.py); other methods => NaN => fit all */
    double stopping_criteria;    /* default 1e-5 */
    double prefilter_threshold;  /* fit iff isnan(p) || p < threshold (diffmod.py:54) */
    double dirichlet_conc;       /* 0.001 (configurator.py:72) */
} xpb200_method;

/* Result arrays, all [P, width] row-major; rows of positions that were not selected are left
 * untouched except status and pval. */
typedef struct xpb200_result {
    uint32_t* status;  /* [P] */
    int32_t* n_iter;   /* [P] sweeps run (GMM.info['n_iterations']) */
    double* pval;      /* [P] prefilter p-value (NaN when not applicable / prefilter off) */
    double* fit;       /* [P, xpb200_fit_width(G)]  mu0 mu1 lam0 lam1 alpha0 alpha1 beta0 beta1 elbo N[g][k] */
    double* stats;     /* [P, xpb200_stats_width(G,C)] mu_unmod mu_mod s2_unmod s2_mod conf_unmod conf_mod
                          p_overlap cdf[4] mod_rate[G] coverage[G] pairwise[C(C,2)][diff,pval,z]
                          one_vs_all[C][diff,pval,z] (only when C > 2) */
    double* coef;      /* optional (may be NULL), device paths only: [P, G+2] coefficients of the LAST E-step of
                          every fitted position, d_n = A_g + y'(B + C y') with y' = y - kmer_mean, stored as
                          A_0..A_{G-1}, B, C - what xpb200_posterior_device needs to restate the z node */
} xpb200_result;

int32_t xpb200_fit_width(int32_t n_groups);
int32_t xpb200_stats_width(int32_t n_groups, int32_t n_conditions);

/* Bytes of device scratch xpb200_fit_device needs for P positions. */
size_t xpb200_workspace_bytes(int32_t n_positions);

/* Whole path on DEVICE pointers: prefilter t-test -> worklist -> VB-EM fit -> statistics,
 * enqueued on `cuda_stream` (a cudaStream_t, may be NULL); does not synchronise.  Positions are fitted by depth class:
 * one warp per position below 1024 reads, a 4-warp team up to what one CTA's shared memory holds (~23 000 reads), a
 * streaming CTA that re-reads the position from global memory in every sweep beyond that - there is no cap on the
 * reads of a position (the reference has none, gmm.py:93-127).  With more than one class the launches run on internal
 * streams forked from and joined back into `cuda_stream`: everything the call enqueues is ordered before later work
 * on `cuda_stream`, as with one stream.  batch->max_reads must not be understated (a position deeper than stated is
 * reported with XPB200_NOT_FITTED instead of being fitted).
 * `n_fitted_dev` (optional, device uint32) receives the number of selected positions. */
int32_t xpb200_fit_device(const xpb200_batch* batch, const xpb200_method* method, const xpb200_result* result,
                          void* workspace, size_t workspace_bytes, uint32_t* n_fitted_dev, void* cuda_stream);

/* Same path on HOST pointers (the drop-in call): copies the batch to the device, runs, copies
 * the results back, synchronises.  `device` is the CUDA ordinal.  Pinned host buffers make the
 * copies asynchronous; pageable ones work too; the result arrays may also be device memory
 * (copies use cudaMemcpyDefault).  `n_fitted` (optional) gets the selected count. */
int32_t xpb200_fit_host(int32_t device, const xpb200_batch* batch, const xpb200_method* method,
                        const xpb200_result* result, uint32_t* n_fitted);

/* Same path on HOST pointers when only diffmod.table matters: nothing but the records of the table's rows
 * comes back - the positions that were fitted AND pass the reference's row filter (io.py:363-367) - as
 * rows[n_rows][4+FW+SW] f64 = [position, status, n_iter, pval, fit[FW], stats[SW]] (the record of
 * xpb200_pack_records), ascending by position.  This is what `execute()` (scripts/diffmod.py:15-77) hands to
 * the table writer.  `rows` holds `capacity_rows` rows (XPB200_ENOMEM and *n_rows = rows needed if more);
 * copies the batch up in chunks like xpb200_fit_host, synchronises.
 * A u16 batch is fitted by ONE persistent launch that follows the chunks as they arrive ("chained
 * form"); it is sized by batch->max_reads, so an understated max_reads costs a restart (never a wrong result), and it
 * keeps all but six SMs of the device busy for the duration of the call. */
int32_t xpb200_fit_host_rows(int32_t device, const xpb200_batch* batch, const xpb200_method* method, double* rows,
                             int64_t capacity_rows, int64_t* n_rows, uint32_t* n_fitted);

/* The same rows left on the DEVICE, without any host synchronisation - the per-rank step of a multi-GPU run, whose
 * rows go on to the writer rank over NCCL (the reference appends rows from many processes under a file lock,
 * scripts/diffmod.py:68-69).  Host buffers in (pinned: the copies are asynchronous; they must stay valid until the
 * work completes); rows_dev[capacity_rows][4+FW+SW] and counts_dev[2] = {rows needed, positions fitted} are device
 * memory.  Record column 0 is index_offset + position.  Rows beyond capacity_rows are dropped (counts_dev[0] still
 * reports the number needed; 0xffffffff there: the persistent launch of the chained form gave up waiting for a chunk -
 * the device was not this call's alone - and the rows are incomplete).  Everything is ordered after the work already enqueued on `cuda_stream`, and later work on
 * `cuda_stream` is ordered after it; the call returns as soon as the work is enqueued. */
int32_t xpb200_fit_host_rows_async(int32_t device, const xpb200_batch* batch, const xpb200_method* method,
                                   double* rows_dev, int64_t capacity_rows, int64_t index_offset, uint32_t* counts_dev,
                                   void* cuda_stream);

/* Page-locked host memory for the buffers of the host entry points (copies from / to it are asynchronous and run at
 * full PCIe rate); NULL when no CUDA device is usable.  Free with xpb200_host_free. */
void* xpb200_host_alloc(size_t bytes);
void xpb200_host_free(void* p);
/* The CUDA device of the calling host thread (cudaSetDevice): xpb200_host_alloc and the first call of a thread create
 * their context there.  A driver calls it once per worker thread with its rank's device, e.g. from a thread started
 * at process entry, so that the CUDA context (0.3 - 1 s) comes up while the inputs are being parsed. */
int32_t xpb200_set_device(int32_t device);

/* Release the cached device buffers of xpb200_fit_host (all devices). */
void xpb200_release(void);

/* Individual stages on device pointers (used by the tests and the profiler runs). */
int32_t xpb200_prefilter_device(const xpb200_batch* batch, const xpb200_method* method, uint32_t* status,
                                double* pval, void* cuda_stream);
int32_t xpb200_stats_device(const xpb200_batch* batch, const xpb200_method* method, const xpb200_result* result,
                            const int32_t* worklist, const uint32_t* n_work_dev, void* cuda_stream);

/* `--save_models` support (io.save_models_to_hdf5, io.py:99-139): the z node of every fitted position, i.e. the
 * responsibilities of its last E-step (gmm.py:236-243), recomputed from result->coef (which the fit must have
 * been given).  prob0[R] = q(z_n = 0) (q(z_n = 1) = 1 - prob0, gmm.py:242); ln_prob[R, 2] = ln q(z_n = k)
 * (gmm.py:243, not clipped).  Reads of positions that were not fitted are left untouched.  Device pointers. */
int32_t xpb200_posterior_device(const xpb200_batch* batch, const xpb200_result* result, double* prob0,
                                double* ln_prob, void* cuda_stream);

/* Compact records of the fitted positions of the last xpb200_fit_device call that used `workspace`
 * (worklist order): row = [index_offset + position, status, n_iter, pval, fit[FW], stats[SW]] as
 * f64, written to out[capacity_rows, 4+FW+SW] (device).  This is the payload of the final gather to
 * the writer rank (the reference appends rows under a file lock instead, scripts/diffmod.py:68-69). */
int32_t xpb200_pack_records(const xpb200_result* result, int32_t n_positions, int32_t n_groups,
                            int32_t n_conditions, const void* workspace, int64_t index_offset, double* out,
                            int64_t capacity_rows, void* cuda_stream);

/* Records (same layout) of the table's rows only - fitted AND past the row filter (io.py:363-367) - of the last
 * xpb200_fit_device call whose results are in `result`, compacted in position order into out[capacity_rows, 4+FW+SW]
 * (device); *n_rows_dev (device) = rows needed.  `workspace` is the fit's workspace (scratch).  Deterministic, no atomics. */
int32_t xpb200_pack_rows(const xpb200_result* result, int32_t n_positions, int32_t n_groups, int32_t n_conditions,
                         void* workspace, int64_t index_offset, double* out, int64_t capacity_rows,
                         uint32_t* n_rows_dev, void* cuda_stream);

/* Launch geometry of the class that fits the positions below the team split of such a batch (for reporting).
 * positions_per_warp and tmem_columns are 0 for the shared-memory kernels (f64 / i32 reads); u16 batches run the
 * tensor-memory kernel: slab_reads = reads a position may have there, tmem_columns = columns per position. */
typedef struct xpb200_launch_info {
    int32_t sm_count, ctas_per_sm, warps_per_cta, smem_per_cta, slab_reads, regs_per_thread;
    int32_t positions_per_warp, tmem_columns;
} xpb200_launch_info;
int32_t xpb200_fit_launch_info(int32_t device, int32_t n_groups, int32_t max_reads, int32_t y_format,
                               xpb200_launch_info* out);

/* Positions with at least this many reads are fitted by the streaming class (reads re-read from global memory in
 * every sweep) instead of being held in shared memory; depends on the number of group slots only. */
int32_t xpb200_deep_split(int32_t n_groups);

/* How the host entry points cut a batch into chunks of positions (no GPU needed): cuts[0] = 0 < cuts[1] < ... <
 * cuts[n] = n_positions, n returned (XPB200_EINVAL when `capacity` < n + 1 entries).  elem_bytes = 8 / 4 / 2 for the f64 /
 * i32 / u16 read formats; chained != 0: the plan of the chained form. */
int32_t xpb200_host_chunk_plan(const int64_t* read_off, int32_t n_positions, int32_t elem_bytes, int32_t chained,
                               int32_t* cuts, int32_t capacity);

/* Kernel launches issued by this library in the calling thread since load (bench.py's gpu_launches). */
uint64_t xpb200_launch_count(void);

/* Per-stage device timing of the calling thread's xpb200_fit_device calls, with CUDA events on
 * the launch stream: after xpb200_set_stage_timing(1), xpb200_get_stage_ms returns the duration of
 * the last call's stages {prefilter, worklist, fit, stats} in milliseconds (it synchronises). */
int32_t xpb200_set_stage_timing(int32_t on);
int32_t xpb200_get_stage_ms(float* ms4);

/* FP64 roofline denominator: dependent-free DFMA streams on every SM; returns TFLOP/s
 * (2 flop per FMA) measured with CUDA events over `repeats` launches. */
int32_t xpb200_measure_fp64_peak(int32_t device, int32_t repeats, double* tflops);

/* Special functions exactly as the kernels evaluate them, run on the device over n values
 * (device pointers); for the accuracy tests.  which: 0 psi, 1 lgamma, 2 log, 3 exp,
 * 4 student_t two-sided p (x = t, aux = df), 5 E-step r1 (x = d), 6 E-step entropy term. */
int32_t xpb200_eval_special(int32_t which, const double* x, const double* aux, double* out, int64_t n,
                            void* cuda_stream);

const char* xpb200_last_error(void);
const char* xpb200_version(void);

/* ---- native ingest of dataprep output (host code, no GPU involved) ---------------------------
 * Replaces, for many genes per call and on several threads, the per-gene Python of the reference:
 *   xpore/scripts/diffmod.py:156-169  seek / read / ujson.loads of one gene from every run's data.json
 *   xpore/diffmod/io.py:21-90         load_data(): position selection and read concatenation
 * and yields the arrays of an xpb200_batch directly.
 *
 * open:  one data.json per run (config order).  run_slot[r] = group slot of run r when not pooling,
 *        run_cond[r] = condition index (rank among the sorted condition names); readcount_min/max are
 *        the YAML `criteria`; `kmers` are the rows of model_kmer.csv (the output refers to them by row).
 * genes: starts/ends are [n_genes, n_runs] byte spans from data.index, start = -1 where the run does
 *        not have the gene.  Parses and selects on `n_threads` threads, keeps the result inside the
 *        handle and reports its sizes.
 * copy:  writes the result: y[n_reads] f64, read_off[n_positions+1], grp_len[n_positions*G],
 *        kmer_row[n_positions], gene_of_position[n_positions] (index into the call's gene list),
 *        strings = "<position>\0<kmer>\0" per position.  Any pointer but read_off may be NULL.
 * Positions come out sorted by (int(position), kmer) inside a gene, genes in call order. */
void* xpb200_ingest_open(int32_t n_runs, const char* const* json_paths, const int32_t* run_slot,
                         const int32_t* run_cond, int32_t n_groups, int32_t n_conditions, int32_t pooling,
                         int32_t readcount_min, int32_t readcount_max, int32_t n_kmers, const char* const* kmers);
int32_t xpb200_ingest_genes(void* handle, int32_t n_genes, const int64_t* starts, const int64_t* ends,
                            int32_t n_threads, int64_t* n_positions, int64_t* n_reads, int64_t* string_bytes);
int32_t xpb200_ingest_copy(void* handle, double* y, int64_t* read_off, int32_t* grp_len, int32_t* kmer_row,
                           int32_t* gene_of_position, char* strings);
/* Encoding the reads of the last xpb200_ingest_genes call can travel in: XPB200_Y_U16 when every literal was a
 * non-negative decimal with at most two fraction digits below 655.36 (what dataprep writes, dataprep.py:638) - the
 * codes q = 100 y are taken from the text itself, and q / 100.0 is bit for bit the double xpb200_ingest_copy yields -
 * else XPB200_Y_F64.  *max_reads (optional) = reads of the deepest position.  xpb200_ingest_copy_u16 writes the codes
 * (yq[n_reads]; pad the buffer to a multiple of 8 elements for xpb200_batch). */
int32_t xpb200_ingest_encoding(void* handle, int32_t* max_reads);
int32_t xpb200_ingest_copy_u16(void* handle, uint16_t* yq);
void xpb200_ingest_close(void* handle);
const char* xpb200_ingest_last_error(void);

/* Test utility: fill the shared memory of every SM with a 64-bit pattern, so that a kernel which reads
 * shared memory before writing it shows up in tests (tests/test_gpu_parity.py uses a NaN pattern). */
int32_t xpb200_debug_fill_smem(uint64_t pattern, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* XPORE_B200_H */
