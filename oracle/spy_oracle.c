/*
 * spy_oracle.c -- CPU restatement of the similaripy sparse-KNN hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity oracle: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load it.  The product (similaripy_b200/) never links or calls it.
 *
 * Every function cites the reference code it restates (paths relative to
 * /root/reference).  It is written from the algorithm's description, not copied:
 * plain C99, explicit binary heap, explicit touched-list accumulator.
 *
 * Parity pinning: checked against oracle/_ref (the reference itself, compiled
 * from /root/reference by oracle/build_ref.py) in tests/test_oracle_vs_ref.py
 * and against the committed golden fixtures tests/golden/ that were generated
 * from oracle/_ref by tests/golden/make_golden.py.
 *
 * Floating point: built with -O2 and WITHOUT -ffast-math, so sums are taken in
 * exactly the order written here (the reference is built with -ffast-math and
 * may reassociate; agreement is to ~1e-6 relative, exact for integer data).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define SPY_SEL_NONE 0
#define SPY_SEL_ARRAY 1
#define SPY_SEL_MATRIX 2

/* ------------------------------------------------------------------------- */
/* Top-K: size-K min-heap on (score, index), lexicographic order.            */
/* Restates TopK::operator() (similaripy/cython_code/s_plus.h:39-64):        */
/*   - while fewer than K held: insert;                                      */
/*   - else replace the minimum iff score > min.score (strict).              */
/* The minimum under std::greater<pair<Value,Index>> is the smallest score   */
/* and, among equal scores, the smallest index.                              */
/* ------------------------------------------------------------------------- */
typedef struct { float score; int32_t index; } hitem_t;

static inline int item_less(hitem_t a, hitem_t b) {
    return (a.score < b.score) || (a.score == b.score && a.index < b.index);
}

typedef struct { hitem_t *h; int size; int cap; } heap_t;

/* The slab keeps the heap's ARRAY order (s_plus.h:443-450), so the oracle reproduces the array layout of
 * std::push_heap / std::pop_heap as libstdc++ implements them (bottom-up variant: pop_heap walks the hole
 * down to a leaf along the preferred children, then sifts the displaced last element up from there).
 * comp = std::greater<pair>: comp(a, b) == item_less(b, a); the root is the minimum. */
static void heap_push_up(hitem_t *h, int hole, int top, hitem_t value) {
    int parent = (hole - 1) / 2;
    while (hole > top && item_less(value, h[parent])) {   /* comp(h[parent], value) */
        h[hole] = h[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    h[hole] = value;
}

static void heap_adjust(hitem_t *h, int hole, int len, hitem_t value) {
    const int top = hole;
    int child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (item_less(h[child - 1], h[child])) child--;    /* comp(h[child], h[child-1]) */
        h[hole] = h[child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        h[hole] = h[child - 1];
        hole = child - 1;
    }
    heap_push_up(h, hole, top, value);
}

static inline void topk_offer(heap_t *hp, int32_t index, float score) {
    hitem_t x; x.score = score; x.index = index;
    if (hp->size < hp->cap) {                               /* push_back + push_heap */
        hp->size++;
        heap_push_up(hp->h, hp->size - 1, 0, x);
    } else if (score > hp->h[0].score) {                    /* pop_heap; back() = x; push_heap */
        int n = hp->size;
        if (n > 1) {
            hitem_t last = hp->h[n - 1];
            hp->h[n - 1] = hp->h[0];
            heap_adjust(hp->h, 0, n - 1, last);
        }
        heap_push_up(hp->h, n - 1, 0, x);
    }
}

/* ------------------------------------------------------------------------- */
/* Per-thread accumulator state.                                             */
/* Restates SparseMatrixMultiplier (s_plus.h:71-240).                        */
/* ------------------------------------------------------------------------- */
typedef struct {
    float *sums;        /* dense, min(block, n_cols) entries, all zero between rows   */
    int32_t *touched;   /* first-touch list ("nonzero_cols", s_plus.h:112-117)        */
    int n_touched, cap_touched;
} accum_t;

static inline void accum_add(accum_t *a, int32_t local_col, float v) {
    /* s_plus.h:112-117: push when the slot is exactly zero BEFORE the add */
    if (a->sums[local_col] == 0.0f) {
        if (a->n_touched == a->cap_touched) {
            a->cap_touched = a->cap_touched * 2;
            a->touched = (int32_t *)realloc(a->touched, sizeof(int32_t) * (size_t)a->cap_touched);
        }
        a->touched[a->n_touched++] = local_col;
    }
    a->sums[local_col] += v;
}

typedef struct {
    const float *Xt, *Yt, *Xc, *Yc, *Xd, *Yd;
    float a1, l1, l2, l3, t1, t2, stab, bayes, thr;
    int filter_mode; const int32_t *f_indptr, *f_indices;
    int target_mode; const int32_t *t_indptr, *t_indices;
} simpar_t;

/* s_plus.h:129-156 computeSimilarity */
static inline float similarity_value(const simpar_t *p, int32_t row, int32_t col, float xy) {
    float vT = 0.f, vC = 0.f, vD = 0.f, val = xy;
    if (p->l1 != 0.f) vT = p->l1 * (p->t1 * (p->Xt[row] - xy) + p->t2 * (p->Yt[col] - xy) + xy);
    if (p->l2 != 0.f) vC = p->l2 * (p->Xc[row] * p->Yc[col]);
    if (p->l3 != 0.f) vD = p->l3 * (p->Xd[row] * p->Yd[col]);
    if (p->a1 != 1.f) xy = powf(xy, p->a1);          /* NB: Tversky term used the un-powered xy */
    if (p->l1 != 0.f || p->l2 != 0.f || p->l3 != 0.f || p->stab != 0.f || p->bayes != 0.f) {
        float den = vT + vC + vD + p->stab;
        val = (den != 0.f) ? xy / den : 0.f;
        if (p->bayes != 0.f) val = val * (xy / (xy + p->bayes));
    }
    return val;
}

/* std::binary_search on a sorted CSR row (s_plus.h:165-169, 181-185) */
static inline int row_has(const int32_t *indptr, const int32_t *indices, int32_t row, int32_t x) {
    int32_t lo = indptr[row], hi = indptr[row + 1];
    while (lo < hi) {
        int32_t mid = lo + ((hi - lo) >> 1);
        if (indices[mid] < x) lo = mid + 1; else hi = mid;
    }
    return lo < indptr[row + 1] && indices[lo] == x;
}

/* s_plus.h:193-215 foreach: drain in first-touch order, clear as we go */
static void accum_drain(accum_t *a, const simpar_t *p, int32_t row, int32_t block_offset, heap_t *hp) {
    for (int i = 0; i < a->n_touched; i++) {
        int32_t lc = a->touched[i];
        float xy = a->sums[lc];
        int32_t col = block_offset + lc;
        int filtered = (p->filter_mode == SPY_SEL_MATRIX) && row_has(p->f_indptr, p->f_indices, row, col);
        int targeted = (p->target_mode != SPY_SEL_MATRIX) || row_has(p->t_indptr, p->t_indices, row, col);
        if (!filtered && targeted) {
            float val = similarity_value(p, row, col, xy);
            if (val >= p->thr) topk_offer(hp, col, val);
        }
        a->sums[lc] = 0.0f;
    }
    a->n_touched = 0;
}

static int32_t lower_bound_i32(const int32_t *a, int32_t lo, int32_t hi, int32_t x) {
    while (lo < hi) {
        int32_t mid = lo + ((hi - lo) >> 1);
        if (a[mid] < x) lo = mid + 1; else hi = mid;
    }
    return lo;
}

/*
 * spy_oracle_knn: restates compute_similarities_parallel<int,float>
 * (s_plus.h:265-453; call site s_plus.pyx:359-384).
 * rows/cols/values must be caller-allocated and ZEROED, n_targets*k each
 * (s_plus.pyx:351-353).  Row i writes <= k triples starting at k*i in heap
 * array order (s_plus.h:443-450).  block_size == 0 disables column blocking.
 */
void spy_oracle_knn(
    int32_t n_targets, const int32_t *targets,
    const float *a_data, const int32_t *a_indices, const int32_t *a_indptr,
    const float *b_data, const int32_t *b_indices, const int32_t *b_indptr,
    const float *Xt, const float *Yt, const float *Xc, const float *Yc, const float *Xd, const float *Yd,
    float a1, float l1, float l2, float l3, float t1, float t2,
    float stab, float bayes, float thr,
    int32_t k, int32_t n_output_cols,
    int32_t filter_mode, const int32_t *f_indptr, const int32_t *f_indices,
    int32_t target_mode, const int32_t *t_indptr, const int32_t *t_indices,
    int32_t *rows, int32_t *cols, float *values,
    int32_t num_threads, int32_t block_size)
{
    const int use_blocking = (block_size > 0) && (n_output_cols > block_size);      /* s_plus.h:311 */
    const int32_t n_blocks = (block_size > 0) ? (n_output_cols + block_size - 1) / block_size : 1;
    const int32_t width = use_blocking ? block_size : n_output_cols;
    simpar_t par = { Xt, Yt, Xc, Yc, Xd, Yd, a1, l1, l2, l3, t1, t2, stab, bayes, thr,
                     filter_mode, f_indptr, f_indices, target_mode, t_indptr, t_indices };
#ifdef _OPENMP
    if (num_threads <= 0) num_threads = omp_get_max_threads();
#else
    num_threads = 1;
#endif
#pragma omp parallel num_threads(num_threads)
    {
        accum_t acc;
        acc.sums = (float *)calloc((size_t)(width > 0 ? width : 1), sizeof(float));
        acc.cap_touched = 1024; acc.n_touched = 0;
        acc.touched = (int32_t *)malloc(sizeof(int32_t) * (size_t)acc.cap_touched);
        heap_t hp; hp.cap = k; hp.size = 0;
        hp.h = (hitem_t *)malloc(sizeof(hitem_t) * (size_t)(k > 0 ? k : 1));
#pragma omp for schedule(dynamic)
        for (int32_t i = 0; i < n_targets; i++) {
            const int32_t t = targets[i];
            const int32_t s1 = a_indptr[t], e1 = a_indptr[t + 1];
            hp.size = 0;
            if (use_blocking) {                                                       /* s_plus.h:350-410 */
                for (int32_t blk = 0; blk < n_blocks; blk++) {
                    const int32_t c0 = blk * block_size;
                    const int32_t c1 = (c0 + block_size < n_output_cols) ? c0 + block_size : n_output_cols;
                    for (int32_t ia = s1; ia < e1; ia++) {
                        const int32_t u = a_indices[ia];
                        const float v1 = a_data[ia];
                        const int32_t s2 = b_indptr[u], e2 = b_indptr[u + 1];
                        if (s2 == e2) continue;
                        if (b_indices[e2 - 1] < c0 || b_indices[s2] >= c1) continue;
                        const int32_t lo = lower_bound_i32(b_indices, s2, e2, c0);
                        const int32_t hi = lower_bound_i32(b_indices, lo, e2, c1);
                        for (int32_t ib = lo; ib < hi; ib++)
                            accum_add(&acc, b_indices[ib] - c0, v1 * b_data[ib]);
                    }
                    if (acc.n_touched > 0) accum_drain(&acc, &par, t, c0, &hp);
                }
            } else {                                                                  /* s_plus.h:411-441 */
                for (int32_t ia = s1; ia < e1; ia++) {
                    const int32_t u = a_indices[ia];
                    const float v1 = a_data[ia];
                    for (int32_t ib = b_indptr[u]; ib < b_indptr[u + 1]; ib++)
                        accum_add(&acc, b_indices[ib], b_data[ib] * v1);
                }
                accum_drain(&acc, &par, t, 0, &hp);
            }
            int64_t o = (int64_t)k * i;                                               /* s_plus.h:443-450 */
            for (int j = 0; j < hp.size; j++) {
                rows[o + j] = t;
                cols[o + j] = hp.h[j].index;
                values[o + j] = hp.h[j].score;
            }
        }
        free(acc.sums); free(acc.touched); free(hp.h);
    }
}

/* ------------------------------------------------------------------------- */
/* coo_to_csr (similaripy/cython_code/coo_to_csr.h:28-71): stable counting   */
/* sort of the slab by row.  64-bit offsets.                                 */
/* ------------------------------------------------------------------------- */
void spy_oracle_coo_to_csr(int32_t n_row, int64_t nnz, const int32_t *Ai, const int32_t *Aj, const float *Ax,
                           int64_t *Bp, int64_t *Bj, float *Bx)
{
    for (int32_t r = 0; r <= n_row; r++) Bp[r] = 0;
    for (int64_t n = 0; n < nnz; n++) Bp[Ai[n] + 1]++;
    for (int32_t r = 0; r < n_row; r++) Bp[r + 1] += Bp[r];
    int64_t *cursor = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n_row > 0 ? n_row : 1));
    for (int32_t r = 0; r < n_row; r++) cursor[r] = Bp[r];
    for (int64_t n = 0; n < nnz; n++) {
        int64_t d = cursor[Ai[n]]++;
        Bj[d] = Aj[n];
        Bx[d] = Ax[n];
    }
    free(cursor);
}

/* ------------------------------------------------------------------------- */
/* Normalizers (similaripy/cython_code/normalization.pyx).  Generated for    */
/* float/double data x int32/int64 index types like the Cython fused types.  */
/* Intermediate precision follows the C that Cython emits for the float      */
/* specialisation: any expression with a 1.0/0.5 literal or log/sqrt is      */
/* evaluated in double and truncated on store.                               */
/* ------------------------------------------------------------------------- */
enum { TF_BINARY = 0, TF_RAW = 1, TF_SQRT = 2, TF_FREQ = 3, TF_LOG = 4 };
enum { IDF_UNARY = 0, IDF_BASE = 1, IDF_SMOOTH = 2, IDF_PROB = 3, IDF_BM25 = 4 };

#define DEFINE_NORMALIZERS(SUF, FT, IT)                                                              \
/* normalization.pyx:47-69 */                                                                        \
static inline FT tf_##SUF(FT freq, FT doc_len, int mode, FT log_logbase) {                           \
    switch (mode) {                                                                                  \
    case TF_BINARY: return (freq != 0) ? (FT)1.0 : (FT)0.0;                                          \
    case TF_RAW:    return freq;                                                                     \
    case TF_SQRT:   return (FT)sqrt((double)freq);                                                   \
    case TF_FREQ:   return freq / doc_len;                                                           \
    default:        return (FT)(log(1.0 + (double)freq) / (double)log_logbase);                      \
    }                                                                                                \
}                                                                                                    \
/* normalization.pyx:72-94 */                                                                        \
static inline FT idf_##SUF(FT df, FT n_docs, int mode, FT log_logbase) {                             \
    switch (mode) {                                                                                  \
    case IDF_UNARY:  return (FT)1.0;                                                                 \
    case IDF_BASE:   return (FT)(log((double)(n_docs / df)) / (double)log_logbase);                  \
    case IDF_SMOOTH: return (FT)(log((double)n_docs / (1.0 + (double)df)) / (double)log_logbase);    \
    case IDF_PROB:   return (FT)(log((double)((n_docs - df) / df)) / (double)log_logbase);           \
    default:         return (FT)(log((((double)(n_docs - df)) + 0.5) / ((double)df + 0.5))           \
                                 / (double)log_logbase);                                             \
    }                                                                                                \
}                                                                                                    \
/* normalization.pyx:97-128 */                                                                       \
void spy_oracle_l2_##SUF(int64_t n_rows, FT *data, const IT *indptr) {                               \
    for (int64_t i = 0; i < n_rows; i++) {                                                           \
        FT s = 0;                                                                                    \
        for (IT j = indptr[i]; j < indptr[i + 1]; j++) s += data[j] * data[j];                       \
        if (s == 0) continue;                                                                        \
        s = (FT)sqrt((double)s);                                                                     \
        for (IT j = indptr[i]; j < indptr[i + 1]; j++) data[j] /= s;                                 \
    }                                                                                                \
}                                                                                                    \
/* normalization.pyx:131-161 */                                                                      \
void spy_oracle_l1_##SUF(int64_t n_rows, FT *data, const IT *indptr) {                               \
    for (int64_t i = 0; i < n_rows; i++) {                                                           \
        FT s = 0;                                                                                    \
        for (IT j = indptr[i]; j < indptr[i + 1]; j++) s += (FT)fabs((double)data[j]);               \
        if (s == 0) continue;                                                                        \
        for (IT j = indptr[i]; j < indptr[i + 1]; j++) data[j] /= s;                                 \
    }                                                                                                \
}                                                                                                    \
/* normalization.pyx:164-197 */                                                                      \
void spy_oracle_max_##SUF(int64_t n_rows, FT *data, const IT *indptr) {                              \
    for (int64_t i = 0; i < n_rows; i++) {                                                           \
        if (indptr[i] == indptr[i + 1]) continue;                                                    \
        FT m = data[indptr[i]];                                                                      \
        for (IT j = indptr[i] + 1; j < indptr[i + 1]; j++) if (data[j] > m) m = data[j];             \
        if (m <= 0) continue;                                                                        \
        for (IT j = indptr[i]; j < indptr[i + 1]; j++) data[j] /= m;                                 \
    }                                                                                                \
}                                                                                                    \
/* normalization.pyx:200-257 */                                                                      \
void spy_oracle_tfidf_##SUF(int64_t n_docs, int64_t n_words, FT *data, const IT *indices,            \
                            const IT *indptr, int tf_mode, int idf_mode, FT logbase) {               \
    FT log_logbase = (FT)log((double)logbase);                                                       \
    FT *idf_ = (FT *)calloc((size_t)(n_words > 0 ? n_words : 1), sizeof(FT));                        \
    FT *doc_len = (FT *)calloc((size_t)(n_docs > 0 ? n_docs : 1), sizeof(FT));                       \
    for (int64_t i = 0; i < n_docs; i++)                                                             \
        for (IT j = indptr[i]; j < indptr[i + 1]; j++) {                                             \
            doc_len[i] += data[j];                                                                   \
            if (data[j] > 0) idf_[indices[j]] += 1;                                                  \
        }                                                                                            \
    for (int64_t w = 0; w < n_words; w++)                                                            \
        if (idf_[w] != 0) idf_[w] = idf_##SUF(idf_[w], (FT)n_docs, idf_mode, log_logbase);           \
    for (int64_t i = 0; i < n_docs; i++)                                                             \
        for (IT j = indptr[i]; j < indptr[i + 1]; j++) {                                             \
            FT t = tf_##SUF(data[j], doc_len[i], tf_mode, log_logbase);                              \
            data[j] = t * idf_[indices[j]];                                                          \
        }                                                                                            \
    free(idf_); free(doc_len);                                                                       \
}                                                                                                    \
/* normalization.pyx:260-334 */                                                                      \
void spy_oracle_bm25plus_##SUF(int64_t n_docs, int64_t n_words, FT *data, const IT *indices,         \
                               const IT *indptr, FT k1, FT b, FT delta, int tf_mode, int idf_mode,   \
                               FT logbase) {                                                         \
    FT log_logbase = (FT)log((double)logbase);                                                       \
    FT avg = 0;                                                                                      \
    FT *idf_ = (FT *)calloc((size_t)(n_words > 0 ? n_words : 1), sizeof(FT));                        \
    FT *doc_len = (FT *)calloc((size_t)(n_docs > 0 ? n_docs : 1), sizeof(FT));                       \
    for (int64_t i = 0; i < n_docs; i++) {                                                           \
        for (IT j = indptr[i]; j < indptr[i + 1]; j++) {                                             \
            doc_len[i] += data[j];                                                                   \
            if (data[j] > 0) idf_[indices[j]] += 1;                                                  \
        }                                                                                            \
        avg += doc_len[i];                                                                           \
    }                                                                                                \
    for (int64_t w = 0; w < n_words; w++)                                                            \
        if (idf_[w] != 0) idf_[w] = idf_##SUF(idf_[w], (FT)n_docs, idf_mode, log_logbase);           \
    if (n_docs == 0) { free(idf_); free(doc_len); return; }                                          \
    avg = avg / (FT)n_docs;                                                                          \
    for (int64_t i = 0; i < n_docs; i++) {                                                           \
        FT ndl = (FT)((1.0 - (double)b) + (double)((b * doc_len[i]) / avg));                         \
        for (IT j = indptr[i]; j < indptr[i + 1]; j++) {                                             \
            FT t = tf_##SUF(data[j], doc_len[i], tf_mode, log_logbase);                              \
            data[j] = (FT)((double)idf_[indices[j]] *                                                \
                           ((((double)t * ((double)k1 + 1.0)) / (double)(t + k1 * ndl)) + (double)delta)); \
        }                                                                                            \
    }                                                                                                \
    free(idf_); free(doc_len);                                                                       \
}

DEFINE_NORMALIZERS(f32_i32, float, int32_t)
DEFINE_NORMALIZERS(f32_i64, float, int64_t)
DEFINE_NORMALIZERS(f64_i32, double, int32_t)
DEFINE_NORMALIZERS(f64_i64, double, int64_t)

int spy_oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
