/*
 * similaripy_b200.h -- C ABI of the B200-native sparse-KNN similarity hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch / scipy types.
 * Each entry point names the reference interface it replaces (paths relative to
 * the upstream repository bogliosimone/similaripy @ a16d939, v0.6.0).
 *
 * Conventions
 *   - every function returns 0 on success, a negative spy_status on failure;
 *     spy_last_error() returns a thread-local, NUL-terminated description.
 *   - "_dev" functions take DEVICE pointers and a cudaStream_t passed as void*;
 *     they are asynchronous on that stream and never synchronise the device.
 *   - "_host" functions take HOST pointers, allocate/copy/free what they need on
 *     the current device and return when the outputs are complete.
 *   - CSR index arrays are int32 (the reference kernel is instantiated for
 *     <int,float> only, s_plus.pyx:360); offsets into the output slab are 64-bit.
 *   - library is re-entrant per (device, stream); no global mutable state except
 *     the cached SM count per device.
 */
#ifndef SIMILARIPY_B200_H_
#define SIMILARIPY_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPY_ABI_VERSION 8

typedef enum {
    SPY_OK = 0,
    SPY_ERR_CUDA = -1,       /* a CUDA runtime call failed (see spy_last_error) */
    SPY_ERR_INVALID = -2,    /* bad argument / enum                              */
    SPY_ERR_NOMEM = -3,      /* device allocation failed                          */
    SPY_ERR_UNSUPPORTED = -4 /* configuration outside what the kernels cover      */
} spy_status;

/* column selection modes == s_plus::SelectionMode (s_plus.h:24-28) */
#define SPY_SEL_NONE 0
#define SPY_SEL_ARRAY 1
#define SPY_SEL_MATRIX 2

/* TF / IDF modes == normalization.pyx:12-24 */
#define SPY_TF_BINARY 0
#define SPY_TF_RAW 1
#define SPY_TF_SQRT 2
#define SPY_TF_FREQ 3
#define SPY_TF_LOG 4
#define SPY_IDF_UNARY 0
#define SPY_IDF_BASE 1
#define SPY_IDF_SMOOTH 2
#define SPY_IDF_PROB 3
#define SPY_IDF_BM25 4

/* value / index dtypes of the normalizer entry points (Cython fused types
 * floating x integral, normalization.pyx:97-102) */
#define SPY_F32 0
#define SPY_F64 1
#define SPY_I32 0
#define SPY_I64 1
/* extra source dtypes accepted by spy_cast_values_dev (matrix.data.astype(float32), s_plus_utils.pyx:307-308) */
#define SPY_VAL_I32 2
#define SPY_VAL_I64 3

/* ---- library / device ---------------------------------------------------- */
int spy_abi_version(void);
/* replaces utils.get_num_threads (utils.pyx:18-25): number of visible B200s */
int spy_device_count(void);
const char *spy_last_error(void);

/*
 * Arguments of the similarity kernel.  Field for field this is the parameter
 * list of s_plus::compute_similarities_parallel<int,float> (s_plus.h:265-304,
 * call site s_plus.pyx:359-384) minus the progress bar / thread count, plus the
 * launch plan the GPU needs (panel split points) and an explicit per-row count.
 */
typedef struct spy_knn_args {
    /* target rows: n_targets row ids of A (s_plus.h:267-268) */
    int32_t n_targets;
    const int32_t *targets;
    /* A = matrix1 in CSR (s_plus.h:269-271) */
    int32_t a_rows;
    const int32_t *a_indptr;
    const int32_t *a_indices;
    const float *a_data;
    /* B = matrix2 in CSR, n_output_cols columns (s_plus.h:272-274, 291).
     * When n_panels > 1 the column indices inside every row must be ascending
     * (same requirement as the reference's blocked path, s_plus.h:381-394). */
    int32_t b_rows;
    int32_t n_cols;
    const int32_t *b_indptr;
    const int32_t *b_indices;
    const float *b_data;
    /* similarity terms (s_plus.h:275-289); a vector may be NULL when its weight is 0 */
    const float *Xtversky, *Ytversky;
    const float *Xcosine, *Ycosine;
    const float *Xdepop, *Ydepop;
    float a1, l1, l2, l3, t1, t2;
    float stabilized_shrink, bayesian_shrink, threshold;
    int32_t k;
    /* per-row column selectors (s_plus.h:292-297): CSR index arrays with sorted rows,
     * indexed by the TARGET ROW ID.  SPY_SEL_ARRAY is treated like NONE (the caller
     * pre-filters B, s_plus.pyx:290-295). */
    int32_t filter_mode;
    const int32_t *filter_indptr, *filter_indices;
    int32_t target_mode;
    const int32_t *target_indptr, *target_indices;
    /* outputs: slab of n_targets*k entries, row i owns [k*i, k*i+k) (s_plus.h:298-300,443-450).
     * Entries are written best-first (value descending, ties by ascending column);
     * the tail of a short row is zero-filled, so the slab never needs a memset.
     * out_rows may be NULL (rows are implied by the slab position); out_counts may be NULL. */
    int32_t *out_rows;
    int32_t *out_cols;
    float *out_values;
    int32_t *out_counts;
    /* launch plan -- zero means "let the library choose" */
    int32_t panel_width;        /* accumulator columns per pass (multiple of 128)          */
    const int32_t *b_split;     /* [b_rows * split_stride] from spy_knn_build_split_dev     */
    int32_t split_stride;
    int32_t n_panels;
    int32_t threads;            /* 512 or 1024 threads per CTA (1024 / threads CTAs per SM)    */
    const void *b_pairs;        /* B's (column, value) entries packed 8 bytes each by
                                 * spy_knn_pack_pairs_dev: the layout the kernel streams       */
    const int32_t *row_order;   /* optional permutation of [0,n_targets): processing order  */
    int64_t b_nnz;              /* stored entries of B (b_indptr[b_rows]); 0 = unknown: only used to
                                 * size `group` from the mean segment length                      */
    int32_t group;              /* flat engine: lanes streaming one B-row segment together (4, 8, 16 or 32);
                                 * stream engine: drain warps of the CTA -- 8, or 16 for target rows with few
                                 * products per panel; 0 = let spy_knn_plan choose (it returns the choice here) */
    /* kernel generation: 0 = let spy_knn_plan choose, SPY_ENGINE_FLAT = knn_flat_kernel (expansion and drain
     * alternate; every configuration), SPY_ENGINE_STREAM = knn_stream_kernel (cp.async ring, the panel is
     * snapshotted into tensor memory and drained concurrently; needs the three tables below) */
    int32_t engine;
    const int32_t *b_chunk_indptr; /* [b_rows * n_panels + 1] first 16-byte chunk of every (row, panel) segment of B  */
    const void *b_chunks;       /* B as chunks of two (column, value) pairs, segments padded to whole chunks
                                 * (spy_knn_pad_chunks_dev)                                                  */
    const int64_t *toff;        /* [n_targets + 1] exclusive scan of the target rows' lengths               */
    int64_t n_entries;          /* toff[n_targets]: stored entries of A in the target rows                    */
    const void *aexp;           /* [n_panels][n_entries] (first chunk, end chunk) of every (entry, panel):
                                 * spy_knn_build_aexp_dev                                                     */
    int64_t a_nnz;              /* stored entries of A (0 = unknown): with b_nnz the planner estimates the scalar
                                 * products per (target row, panel) and keeps short rows on the flat engine       */
    int32_t unit_values;        /* non-zero: every stored value of A and B is 1.0 (binary=True, s_plus_utils.pyx:301-304).
                                 * The stream engine then counts the products with native integer adds; results are
                                 * bit-identical to the float path (sums of ones are exact).  0 is always correct.    */
} spy_knn_args;

#define SPY_ENGINE_AUTO 0
#define SPY_ENGINE_FLAT 1
#define SPY_ENGINE_STREAM 2
/* what SPY_ENGINE_AUTO resolves to when the configuration is covered by both */
#define SPY_ENGINE_DEFAULT SPY_ENGINE_STREAM

/* Choose panel_width / n_panels / split_stride / threads / group for a problem (device < 0: plan with B200
 * defaults without touching the CUDA runtime).  Fills the plan fields of *args. */
int spy_knn_plan(spy_knn_args *args, int device);

/* Bytes of device scratch spy_knn_topk_dev needs for *args (after spy_knn_plan). */
int64_t spy_knn_scratch_bytes(const spy_knn_args *args, int device);

/* Panel split points of B: split[u*stride + p] = first position q in row u with
 * b_indices[q] >= p*panel_width (p = 0..n_panels).  Replaces the per-(row,block)
 * std::lower_bound calls of the blocked path (s_plus.h:381-394) by one pass over B. */
int spy_knn_build_split_dev(int32_t b_rows, const int32_t *b_indptr, const int32_t *b_indices,
                            int32_t panel_width, int32_t n_panels, int32_t split_stride,
                            int32_t *split_out, void *stream);

/* pairs_out[q] = (b_indices[q], bits of b_data[q]) as one 8-byte word per stored entry of B: the layout
 * the hot kernel streams (16-byte gathers of two pairs).  pairs_out holds (nnz + 1) * 8 bytes, 16-byte aligned
 * (the kernel may read one pair past the end; its content is ignored). */
int spy_knn_pack_pairs_dev(int64_t nnz, const int32_t *b_indices, const float *b_data, void *pairs_out,
                           void *stream);

/* ---- tables of the stream engine (SPY_ENGINE_STREAM), built on the device ------------------------------------
 * Every (row u of B, panel p) segment -- the entries of B[u,:] with columns in panel p -- is stored as whole 16-byte
 * chunks of two (byte offset of the column's accumulator slot inside panel p = 4 * (column - p * panel_width), value
 * bits) pairs; inside a segment the pairs are ordered by shared-memory bank of their slot (column mod 32), an odd
 * segment ends with the filler pair (4 * panel_width, 0) -- the spare slot behind the panel.
 *   counts[u * n_panels + p] = chunks of the segment; the caller scans them into chunk_indptr
 *   (spy_exclusive_scan_i32_dev, b_rows * n_panels + 1 entries).  b_split may be NULL when n_panels == 1. */
int spy_knn_chunk_counts_dev(int32_t b_rows, const int32_t *b_indptr, const int32_t *b_split, int32_t split_stride,
                             int32_t n_panels, int32_t *counts, void *stream);
/* chunks_out holds chunk_indptr[b_rows * n_panels] * 16 bytes; panel_width = the plan's (the pairs are stored relative to
 * their panel, see above) */
int spy_knn_pad_chunks_dev(int32_t b_rows, const int32_t *b_indptr, const int32_t *b_indices, const float *b_data,
                           const int32_t *b_split, int32_t split_stride, int32_t n_panels, const int32_t *chunk_indptr,
                           void *chunks_out, void *stream, int32_t panel_width);
/* len[i] = stored entries of target row i; the caller scans them into toff (spy_exclusive_scan_i64_dev) */
int spy_knn_row_lengths_dev(int32_t n_targets, const int32_t *targets, const int32_t *a_indptr, int32_t *len, void *stream);
/* aexp[p * n_entries + toff[i] + j] = chunk range of the part of B[u,:] (u = j-th entry of target row i) that falls
 * into panel p.  Replaces the per-(target row, block, B row) std::lower_bound of the blocked path
 * (s_plus.h:381-394) by one coalesced table per call.  Uses targets, a_*, b_chunk_indptr, n_panels, toff, n_entries of
 * *args; writes args->aexp. */
int spy_knn_build_aexp_dev(const spy_knn_args *args, void *stream);

/* The hot path.  Replaces s_plus::compute_similarities_parallel<int,float>
 * (s_plus.h:265-453).  Device pointers; scratch is spy_knn_scratch_bytes() bytes. */
int spy_knn_topk_dev(const spy_knn_args *args, void *scratch, int64_t scratch_bytes, void *stream);

/* Same contract with HOST pointers (exactly what the reference's Cython call site holds,
 * s_plus.pyx:359-384): uploads, plans, runs, downloads.  out_rows/out_cols/out_values are
 * host arrays of n_targets*k entries. */
int spy_knn_topk_host(const spy_knn_args *host_args, int device);

/* The same result in the REFERENCE'S OWN ORDER (opt-in, slower): float sums built entry by entry of the target row
 * (s_plus.h:418-438), candidates enlisted like SparseMatrixMultiplier::add (s_plus.h:112-117), ties at the k-th
 * value kept like TopK's heap keeps them (s_plus.h:45-59; closed form in csrc/knn_reforder.cu).  Takes plain CSR
 * operands (no launch plan, no stream layouts).  block_size > 0 and < n_cols: the blocked path (s_plus.h:350-410) --
 * the caller passes matrix2, the Y vectors and the selector matrices with columns permuted by popularity
 * (s_plus_utils.pyx:493-618) and out_col_map = the back permutation; ties then compare permuted ids, like the
 * reference's heap.  out_col_map may be NULL (identity). */
int64_t spy_knn_reforder_scratch_bytes(const spy_knn_args *args);
int spy_knn_topk_reforder_dev(const spy_knn_args *args, int32_t block_size, const int32_t *out_col_map, void *scratch,
                              int64_t scratch_bytes, void *stream);

/* The same call over several GPUs of one box from one process: the target rows are cut into
 * n_devices contiguous ranges of equal work (stored entries of A), every range runs on its device from its own host
 * thread, B is replicated.  The caller's slab is the complete result in target order (assemble is accepted for
 * symmetry with the torch.distributed path, similaripy_b200/sharded.py, where gathering is optional); range_bounds
 * (n_devices + 1 entries, may be NULL) receives the cut positions.  The reference has one OpenMP loop (s_plus.h:313-338). */
int spy_knn_topk_multi_host(const spy_knn_args *host_args, const int32_t *devices, int32_t n_devices, int32_t assemble,
                            int32_t *range_bounds);

/* work[i] = number of scalar products the expansion of target row i performs,
 * sum over u in A[targets[i],:] of nnz(B[u,:]).  The reference balances rows over threads with
 * `omp for schedule(dynamic)` (s_plus.h:337); across GPUs the target rows are cut into contiguous ranges
 * of equal work instead (similaripy_b200/sharded.py). */
int spy_knn_row_work_dev(int32_t n_targets, const int32_t *targets, const int32_t *a_indptr,
                         const int32_t *a_indices, const int32_t *b_indptr, int64_t *work, void *stream);

/* Kernel launches performed by the calling thread since the last call (diagnostics / bench). */
int64_t spy_launch_count(int reset);

/* ---- pre-processing on CSR (device pointers) ------------------------------ */
/* csr_sum axis=1 (s_plus_utils.pyx:151-159): out[r] = sum data (or data^2) of row r, fp32 */
int spy_csr_row_sum_dev(int32_t n_rows, const int32_t *indptr, const float *data, int square,
                        float *out, void *stream);
/* csr_sum axis=0 (s_plus_utils.pyx:160-164): fp64 accumulation, cast to fp32.
 * acc64 is n_cols doubles of scratch. */
int spy_csr_col_sum_dev(int64_t nnz, const int32_t *indices, const float *data, int square,
                        int32_t n_cols, double *acc64, float *out, void *stream);
/* np.power(x + shift, p, dtype=float32) (s_plus_utils.pyx:226-227, 258-274); x may be f32 or f64 */
int spy_pow_shift_dev(int64_t n, const void *x, int x_dtype, float shift, float p, float *out, void *stream);
/* column histogram of a CSR (used by transpose and by array-mode filtering) */
int spy_csr_col_count_dev(int64_t nnz, const int32_t *indices, int32_t n_cols, int32_t *counts, void *stream);
/* exclusive scan of int32 counts into int32 offsets of length n+1 (n <= 2^31-2); tmp = spy_scan_tmp_bytes(n) */
int64_t spy_scan_tmp_bytes(int64_t n);
int spy_exclusive_scan_i32_dev(int64_t n, const int32_t *counts, int32_t *offsets, void *tmp, void *stream);
int spy_exclusive_scan_i64_dev(int64_t n, const int32_t *counts, int64_t *offsets, void *tmp, void *stream);
/* CSR -> CSC (== transpose in CSR).  Replaces scipy's csr_tocsc behind matrix1.T.tocsr()
 * (s_plus.pyx:170,205-206).  With sort_rows != 0 every output row has ascending indices like scipy's; with 0
 * the entries of a row are in arrival order, which is all the LEFT operand of the similarity kernel needs
 * (only B's rows are searched for panel boundaries).  t_indptr must already hold the exclusive scan of
 * the column counts (n_cols+1); cursor is n_cols int32 of scratch. */
int spy_csr_transpose_dev(int32_t n_rows, int32_t n_cols, const int32_t *indptr, const int32_t *indices,
                          const float *data, const int32_t *t_indptr, int32_t *t_indices, float *t_data,
                          int32_t *cursor, int sort_rows, void *stream);
/* sort_indices (s_plus_utils.pyx:344,573): ascending column ids inside every row, in place */
int spy_csr_sort_rows_dev(int32_t n_rows, const int32_t *indptr, int32_t *indices, float *data, void *stream);
/* keep[i] != 0 entries only (eliminate_zeros s_plus.pyx:210-211 with keep = data != 0;
 * _filter_matrix_columns s_plus_utils.pyx:424-490 with keep = mask[indices]).
 * Two calls: count (fills row_counts), then -- after the caller scanned them -- compact. */
int spy_csr_filter_count_dev(int32_t n_rows, const int32_t *indptr, const int32_t *indices, const float *data,
                             const uint8_t *col_mask, int drop_zeros, int32_t *row_counts, void *stream);
int spy_csr_filter_compact_dev(int32_t n_rows, const int32_t *indptr, const int32_t *indices, const float *data,
                               const uint8_t *col_mask, int drop_zeros, const int32_t *new_indptr,
                               int32_t *new_indices, float *new_data, void *stream);
/* ---- matrices with more than 2^31-1 stored entries (64-bit indptr) -----------------------------
 * The reference narrows indptr / indices to int32 (s_plus.pyx:241-244) and its kernel is <int,float>
 * (s_plus.pyx:360), so such a matrix overflows there.  Here it is cut into int32-indexed BLOCKS that keep the
 * shape of the whole matrix: block [lo, hi) of the stored entries keeps those entries and leaves every other
 * row empty; its indices / values are the views indices + lo, data + lo of the 64-bit matrix and
 *   block_indptr[r] = clamp(indptr[r], lo, hi) - lo           (r = 0..n_rows, so n = n_rows + 1)
 * comes from spy_csr_wide_block_indptr_dev.  Pieces that are populated on disjoint row ranges and stored one
 * after the other are stacked by adding their indptr arrays (spy_csr_indptr_add_dev: acc[r] += piece[r]).
 * Target rows are computed block by block of matrix1's rows; the columns of matrix2 block by block, each
 * giving the best k of its columns, merged by spy_slab_merge_dev. */
int spy_csr_wide_block_indptr_dev(int64_t n, const int64_t *indptr, int64_t lo, int64_t hi, int32_t *out, void *stream);
int spy_csr_indptr_add_dev(int64_t n, const int32_t *piece, int32_t *acc, void *stream);
/* Best k of two slabs over DISJOINT column sets, both with their rows best-first (value descending, then
 * column ascending -- the order every kernel of this library writes): the k largest values of the union
 * (TopK, s_plus.h:45-59), rows padded with (0, 0.0) like the reference's slab (s_plus.pyx:351-353).
 * out_* must not alias the inputs. */
int spy_slab_merge_dev(int32_t n_targets, int32_t k, const int32_t *cols_a, const float *vals_a, const int32_t *counts_a,
                       const int32_t *cols_b, const float *vals_b, const int32_t *counts_b, int32_t *out_cols,
                       float *out_vals, int32_t *out_counts, void *stream);
/* int64 -> int32 index arrays on the device (the reference's host-side astype(int32), s_plus.pyx:241-244);
 * the caller has checked that every value fits (matrix dimensions and stored entries within int32) */
int spy_narrow_index_dev(int64_t n, const int64_t *src, int32_t *dst, void *stream);
/* dtype conversion of values: binary => ones (s_plus_utils.pyx:281-308) */
int spy_cast_values_dev(int64_t n, const void *src, int src_dtype, int binary, float *dst, void *stream);

/* ---- output assembly (device pointers) ------------------------------------- */
/* Slab -> CSR with explicit zeros removed: replaces build_csr_matrix + coo_to_csr +
 * eliminate_zeros (utils.pyx:67-173, coo_to_csr.h:28-71, s_plus.pyx:424).
 *   1. spy_slab_row_nnz_dev: row_nnz[targets[i]] = #entries with value != 0 among the first
 *      counts[i] of slab row i (row_nnz has one int per OUTPUT row and must be zeroed first;
 *      target rows must be unique -- duplicates are assembled by the host wrapper; a slab row
 *      with targets[i] < 0 is padding (gathered slabs of sharded runs) and is skipped);
 *   2. the caller scans row_nnz into csr_indptr (spy_exclusive_scan_i64_dev);
 *   3. spy_slab_compact_dev writes the kept (col, value) pairs at csr_indptr[targets[i]].
 * idx_dtype selects int32 / int64 csr_indices (get_index_dtype, utils.pyx:28-40). */
int spy_slab_row_nnz_dev(int32_t n_targets, int32_t k, const float *values, const int32_t *counts,
                         const int32_t *targets, int32_t *row_nnz, void *stream);
int spy_slab_compact_dev(int32_t n_targets, int32_t k, const int32_t *cols, const float *values,
                         const int32_t *counts, const int32_t *targets, const int64_t *csr_indptr,
                         void *csr_indices, int idx_dtype, float *csr_data, void *stream);
/* COO rows for the slab exactly as the reference leaves them (s_plus.h:443-450,
 * s_plus.pyx:351-353): rows[i*k+j] = targets[i] for j < counts[i], else 0. */
int spy_slab_fill_rows_dev(int32_t n_targets, int32_t k, const int32_t *targets, const int32_t *counts,
                           int32_t *rows, void *stream);

/* ---- in-place CSR normalizers (device pointers) --------------------------- */
/* inplace_normalize_csr_{l1,l2,max} (normalization.pyx:97-197); norm: 0=l1 1=l2 2=max */
int spy_normalize_rows_dev(int norm, int64_t n_rows, void *data, int val_dtype, const void *indptr,
                           int idx_dtype, void *stream);
/* inplace_normalize_csr_tfidf (normalization.pyx:200-257).
 * scratch: spy_tfidf_scratch_bytes(n_rows, n_cols, val_dtype) bytes. */
int64_t spy_tfidf_scratch_bytes(int64_t n_rows, int64_t n_cols, int val_dtype);
int spy_tfidf_dev(int64_t n_rows, int64_t n_cols, void *data, int val_dtype, const void *indices,
                  const void *indptr, int idx_dtype, int tf_mode, int idf_mode, double logbase,
                  void *scratch, void *stream);
/* inplace_normalize_csr_bm25plus (normalization.pyx:260-334); bm25 == delta 0 */
int spy_bm25plus_dev(int64_t n_rows, int64_t n_cols, void *data, int val_dtype, const void *indices,
                     const void *indptr, int idx_dtype, double k1, double b, double delta,
                     int tf_mode, int idf_mode, double logbase, void *scratch, void *stream);

/* Host -> device copy of PAGEABLE memory (what a scipy matrix normally lives in) through two pinned 16 MB staging buffers:
 * several host threads fill one buffer while the DMA of the other runs on `stream`.  Returns once src_host has been read
 * (the DMA of the last chunks may still be in flight on the stream). */
int spy_h2d_staged(void *dst_dev, const void *src_host, int64_t bytes, int device, void *stream);

/* The same three with HOST pointers: exactly the arguments of the reference's Cython functions
 * inplace_normalize_csr_{l1,l2,max} / _tfidf / _bm25plus (normalization.pyx:97-102, 200-208, 260-271) -- the data,
 * indices and indptr arrays of a scipy CSR matrix; `data` is overwritten in place. */
int spy_normalize_rows_host(int norm, int64_t n_rows, void *data, int val_dtype, const void *indptr, int idx_dtype, int device);
int spy_tfidf_host(int64_t n_rows, int64_t n_cols, void *data, int val_dtype, const void *indices, const void *indptr,
                   int idx_dtype, int tf_mode, int idf_mode, double logbase, int device);
int spy_bm25plus_host(int64_t n_rows, int64_t n_cols, void *data, int val_dtype, const void *indices, const void *indptr,
                      int idx_dtype, double k1, double b, double delta, int tf_mode, int idf_mode, double logbase, int device);

#ifdef __cplusplus
}
#endif
#endif /* SIMILARIPY_B200_H_ */
