// normalize.cu -- in-place CSR normalizers for sm_100a: l1 / l2 / max, tf-idf, BM25 / BM25+.
//
// Replaces the serial Cython loops of similaripy/cython_code/normalization.pyx:
//   inplace_normalize_csr_l2 :97-128, _l1 :131-161, _max :164-197,
//   inplace_normalize_csr_tfidf :200-257, inplace_normalize_csr_bm25plus :260-334,
//   tf :47-69, idf :72-94
// for the fused types floating in {float,double} x integral in {int32,int64}.
//
// All kernels are HBM-bound streams: one warp per CSR row (coalesced segments), grid-stride over
// rows with the grid a multiple of the SM count.  Intermediate precision mirrors the C that
// Cython generates for the float specialisation: every expression holding a 1.0/0.5 literal or a
// libm call is evaluated in double and truncated on store.  The BM25 average document length is
// accumulated sequentially in storage precision, in row order, like normalization.pyx:310-323.
#include "common.cuh"
#include <mutex>

namespace spy {

constexpr int kNT = 256;
constexpr int kNW = kNT / 32;

static inline int rows_grid(long long n_rows) {
    long long b = (n_rows + kNW - 1) / kNW;
    const long long cap = (long long)kB200SmCount * 16;
    if (b > cap) b = cap;
    return (int)(b < 1 ? 1 : b);
}

template <typename T>
__device__ __forceinline__ T warp_sum_t(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
template <typename T>
__device__ __forceinline__ T warp_max_t(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        T w = __shfl_xor_sync(0xffffffffu, v, o);
        v = (w > v) ? w : v;
    }
    return v;
}

// norm: 0 = l1, 1 = l2, 2 = max
template <typename T, typename I, int NORM>
__global__ void normalize_rows_kernel(long long n_rows, T *__restrict__ data, const I *__restrict__ indptr) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = (long long)blockIdx.x * kNW + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * kNW;
    for (long long r = warp0; r < n_rows; r += nwarps) {
        const long long s = (long long)indptr[r], e = (long long)indptr[r + 1];
        if (s == e) continue;
        T red;
        if (NORM == 2) {
            T m = data[s];  // normalization.pyx:189 starts from the first stored value
            for (long long q = s + lane; q < e; q += 32) { T v = data[q]; m = (v > m) ? v : m; }
            red = warp_max_t(m);
            if (red <= (T)0) continue;  // normalization.pyx:194
        } else {
            T acc = (T)0;
            for (long long q = s + lane; q < e; q += 32) {
                T v = data[q];
                acc += (NORM == 1) ? v * v : (T)fabs((double)v);
            }
            red = warp_sum_t(acc);
            if (red == (T)0) continue;  // normalization.pyx:124, :158
            if (NORM == 1) red = (T)sqrt((double)red);  // normalization.pyx:126
        }
        for (long long q = s + lane; q < e; q += 32) data[q] = data[q] / red;
    }
}

// pass 1 of tf-idf / bm25: doc_len[row] = sum of the row (storage precision), df[col] += (value > 0).  bm25 runs the two
// halves as two launches (DOCLEN, then DF) so that the serial mean of doc_len overlaps the histogram; the row sums are
// built in the same order either way.
template <typename T, typename I, bool DOCLEN, bool DF>
__global__ void doclen_df_kernel(long long n_rows, const T *__restrict__ data, const I *__restrict__ indices,
                                 const I *__restrict__ indptr, T *__restrict__ doc_len, int *__restrict__ df) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = (long long)blockIdx.x * kNW + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * kNW;
    for (long long r = warp0; r < n_rows; r += nwarps) {
        const long long s = (long long)indptr[r], e = (long long)indptr[r + 1];
        T acc = (T)0;
        for (long long q = s + lane; q < e; q += 32) {
            const T v = data[q];
            if (DOCLEN) acc += v;
            if (DF && v > (T)0) atomicAdd(df + (long long)indices[q], 1);
        }
        if (DOCLEN) {
            acc = warp_sum_t(acc);
            if (lane == 0) doc_len[r] = acc;
        }
    }
}

// idf (normalization.pyx:72-94) applied where df != 0 (normalization.pyx:248-250, 317-319)
template <typename T>
__device__ __forceinline__ T idf_value(T df, T n_docs, int mode, T log_logbase) {
    switch (mode) {
    case SPY_IDF_UNARY: return (T)1.0;
    case SPY_IDF_BASE: return (T)(log((double)(n_docs / df)) / (double)log_logbase);
    case SPY_IDF_SMOOTH: return (T)(log((double)n_docs / (1.0 + (double)df)) / (double)log_logbase);
    case SPY_IDF_PROB: return (T)(log((double)((n_docs - df) / df)) / (double)log_logbase);
    default: return (T)(log((((double)(n_docs - df)) + 0.5) / ((double)df + 0.5)) / (double)log_logbase);
    }
}
template <typename T>
__global__ void idf_kernel(long long n_cols, const int *__restrict__ df, T n_docs, int mode, T log_logbase,
                           T *__restrict__ idf) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < n_cols; c += stride) {
        const int d = df[c];
        idf[c] = (d != 0) ? idf_value<T>((T)d, n_docs, mode, log_logbase) : (T)0;
    }
}

// tf (normalization.pyx:47-69)
template <typename T>
__device__ __forceinline__ T tf_value(T freq, T doc_len, int mode, T log_logbase) {
    switch (mode) {
    case SPY_TF_BINARY: return (freq != (T)0) ? (T)1.0 : (T)0.0;
    case SPY_TF_RAW: return freq;
    case SPY_TF_SQRT: return (T)sqrt((double)freq);
    case SPY_TF_FREQ: return freq / doc_len;
    default: return (T)(log(1.0 + (double)freq) / (double)log_logbase);
    }
}

template <typename T, typename I>
__global__ void tfidf_apply_kernel(long long n_rows, T *__restrict__ data, const I *__restrict__ indices,
                                   const I *__restrict__ indptr, const T *__restrict__ doc_len,
                                   const T *__restrict__ idf, int tf_mode, T log_logbase) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = (long long)blockIdx.x * kNW + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * kNW;
    for (long long r = warp0; r < n_rows; r += nwarps) {
        const long long s = (long long)indptr[r], e = (long long)indptr[r + 1];
        const T dl = doc_len[r];
        for (long long q = s + lane; q < e; q += 32) {
            const T t = tf_value<T>(data[q], dl, tf_mode, log_logbase);
            data[q] = t * idf[(long long)indices[q]];  // normalization.pyx:257
        }
    }
}

// avg_doc_len: sequential sum in row order and storage precision (normalization.pyx:297,315,323).  The add chain is
// inherently serial (one dependent add per row: ~4 cycles each); everything else is taken off it: the whole CTA
// streams the next tile into shared memory (coalesced, deep memory parallelism) while thread 0 adds the current one.
constexpr int kMeanTile = 2048;  // 2 x 2048 doubles = 32 KB of static shared memory
template <typename T>
__global__ void __launch_bounds__(1024) sequential_mean_kernel(long long n, const T *__restrict__ x, T *__restrict__ out) {
    __shared__ T buf[2][kMeanTile];
    const int tid = threadIdx.x;
    T acc = (T)0;
    for (int i = tid; i < kMeanTile; i += 1024) buf[0][i] = (i < n) ? x[i] : (T)0;
    __syncthreads();
    int cur = 0;
    for (long long c0 = 0; c0 < n; c0 += kMeanTile) {
        const long long nxt = c0 + kMeanTile;
        if (nxt < n)
            for (int i = tid; i < kMeanTile; i += 1024) buf[cur ^ 1][i] = (nxt + i < n) ? x[nxt + i] : (T)0;
        if (tid == 0) {
            const int m = (int)((n - c0 < kMeanTile) ? (n - c0) : kMeanTile);
            const T *b = buf[cur];
            int i = 0;
            for (; i + 8 <= m; i += 8) {  // loads first, then the dependent adds in row order (a version that loads the next
                                          // eight values while the current eight are added, and one that feeds the chain by
                                          // warp shuffles, were measured slower: 6.0 and 6.5 vs 5.1 ms per bm25 call)
                const T v0 = b[i], v1 = b[i + 1], v2 = b[i + 2], v3 = b[i + 3], v4 = b[i + 4], v5 = b[i + 5], v6 = b[i + 6], v7 = b[i + 7];
                acc += v0; acc += v1; acc += v2; acc += v3; acc += v4; acc += v5; acc += v6; acc += v7;
            }
            for (; i < m; i++) acc += b[i];
        }
        __syncthreads();
        cur ^= 1;
    }
    if (tid == 0) out[0] = acc / (T)n;
}

template <typename T, typename I>
__global__ void bm25_apply_kernel(long long n_rows, T *__restrict__ data, const I *__restrict__ indices,
                                  const I *__restrict__ indptr, const T *__restrict__ doc_len,
                                  const T *__restrict__ idf, const T *__restrict__ avg_ptr, T k1, T b, T delta,
                                  int tf_mode, T log_logbase) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = (long long)blockIdx.x * kNW + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * kNW;
    const T avg = avg_ptr[0];
    for (long long r = warp0; r < n_rows; r += nwarps) {
        const long long s = (long long)indptr[r], e = (long long)indptr[r + 1];
        const T dl = doc_len[r];
        // normalization.pyx:327   norm_doc_len = (1.0 - b) + b * doc_len / avg
        const T ndl = (T)((1.0 - (double)b) + (double)((b * dl) / avg));
        for (long long q = s + lane; q < e; q += 32) {
            const T t = tf_value<T>(data[q], dl, tf_mode, log_logbase);
            // normalization.pyx:334   idf * (tf * (k1 + 1.0) / (tf + k1 * ndl) + delta)
            const double w = (((double)t * ((double)k1 + 1.0)) / (double)(t + k1 * ndl)) + (double)delta;
            data[q] = (T)((double)idf[(long long)indices[q]] * w);
        }
    }
}

// one non-blocking side stream per device, created on first use (NULL when the runtime refuses: the caller then stays on
// its own stream)
static cudaStream_t side_stream() {
    static cudaStream_t streams[64] = {nullptr};
    static std::mutex mu;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    std::lock_guard<std::mutex> lock(mu);
    if (!streams[dev]) {
        // highest priority: when the row sums are done, the one CTA of the serial mean must get its SM before the histogram's
        // grid has filled every SM (it would otherwise start when the first of those CTAs retires, i.e. at the end)
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        if (cudaStreamCreateWithPriority(&streams[dev], cudaStreamNonBlocking, hi) != cudaSuccess) streams[dev] = nullptr;
    }
    return streams[dev];
}

struct Scratch {
    void *doc_len, *idf, *avg;
    int *df;
};
static Scratch carve(void *scratch, long long n_rows, long long n_cols) {
    unsigned char *p = (unsigned char *)scratch;
    Scratch s;
    s.avg = p; p += 256;
    s.doc_len = p; p += ((n_rows * 8 + 255) / 256) * 256;
    s.idf = p; p += ((n_cols * 8 + 255) / 256) * 256;
    s.df = (int *)p;
    return s;
}

template <typename T, typename I>
static int run_tfidf_bm25(bool bm25, long long n_rows, long long n_cols, T *data, const I *indices, const I *indptr,
                          double k1, double b, double delta, int tf_mode, int idf_mode, double logbase,
                          void *scratch, cudaStream_t st) {
    if (n_rows <= 0) return SPY_OK;  // normalization.pyx:321-322 (nothing to weight)
    Scratch sc = carve(scratch, n_rows, n_cols);
    const T log_logbase = (T)log((double)(T)logbase);  // `floating logbase`, normalization.pyx:227,293
    if (n_cols > 0) SPY_CUDA_OK(cudaMemsetAsync(sc.df, 0, (size_t)n_cols * sizeof(int), st));
    const int grid = rows_grid(n_rows);
    cudaEvent_t ev_mean = nullptr;
    if (!bm25) {
        doclen_df_kernel<T, I, true, true><<<grid, kNT, 0, st>>>(n_rows, data, indices, indptr, (T *)sc.doc_len, sc.df);
        SPY_LAUNCH_OK();
    } else {
        // avg_doc_len is one dependent add per row on ONE SM (about 3 ms for 10^6 rows); it only needs doc_len, so it runs on
        // a side stream next to the document-frequency histogram and the idf table, which keep the other SMs busy
        doclen_df_kernel<T, I, true, false><<<grid, kNT, 0, st>>>(n_rows, data, indices, indptr, (T *)sc.doc_len, sc.df);
        SPY_LAUNCH_OK();
        cudaStream_t side = side_stream();
        cudaEvent_t ev_len = nullptr;
        if (side) {
            SPY_CUDA_OK(cudaEventCreateWithFlags(&ev_len, cudaEventDisableTiming));
            SPY_CUDA_OK(cudaEventCreateWithFlags(&ev_mean, cudaEventDisableTiming));
            SPY_CUDA_OK(cudaEventRecord(ev_len, st));
            SPY_CUDA_OK(cudaStreamWaitEvent(side, ev_len, 0));
        }
        sequential_mean_kernel<T><<<1, 1024, 0, side ? side : st>>>(n_rows, (const T *)sc.doc_len, (T *)sc.avg);
        SPY_LAUNCH_OK();
        if (side) {
            SPY_CUDA_OK(cudaEventRecord(ev_mean, side));
            SPY_CUDA_OK(cudaEventDestroy(ev_len));  // (released when the work queued on it has completed)
        }
        doclen_df_kernel<T, I, false, true><<<grid, kNT, 0, st>>>(n_rows, data, indices, indptr, (T *)sc.doc_len, sc.df);
        SPY_LAUNCH_OK();
    }
    if (n_cols > 0) {
        long long cb = (n_cols + kNT - 1) / kNT;
        if (cb > kB200SmCount * 16) cb = kB200SmCount * 16;
        idf_kernel<T><<<(int)cb, kNT, 0, st>>>(n_cols, sc.df, (T)n_rows, idf_mode, log_logbase, (T *)sc.idf);
        SPY_LAUNCH_OK();
    }
    if (!bm25) {
        tfidf_apply_kernel<T, I><<<grid, kNT, 0, st>>>(n_rows, data, indices, indptr, (const T *)sc.doc_len,
                                                       (const T *)sc.idf, tf_mode, log_logbase);
        SPY_LAUNCH_OK();
    } else {
        if (ev_mean) {
            SPY_CUDA_OK(cudaStreamWaitEvent(st, ev_mean, 0));
            SPY_CUDA_OK(cudaEventDestroy(ev_mean));
        }
        bm25_apply_kernel<T, I><<<grid, kNT, 0, st>>>(n_rows, data, indices, indptr, (const T *)sc.doc_len,
                                                      (const T *)sc.idf, (const T *)sc.avg, (T)k1, (T)b, (T)delta,
                                                      tf_mode, log_logbase);
        SPY_LAUNCH_OK();
    }
    return SPY_OK;
}

template <typename T, typename I>
static int run_normalize(int norm, long long n_rows, T *data, const I *indptr, cudaStream_t st) {
    if (n_rows <= 0) return SPY_OK;
    const int grid = rows_grid(n_rows);
    switch (norm) {
    case 0: normalize_rows_kernel<T, I, 0><<<grid, kNT, 0, st>>>(n_rows, data, indptr); break;
    case 1: normalize_rows_kernel<T, I, 1><<<grid, kNT, 0, st>>>(n_rows, data, indptr); break;
    case 2: normalize_rows_kernel<T, I, 2><<<grid, kNT, 0, st>>>(n_rows, data, indptr); break;
    default: set_error("normalize: norm must be 0 (l1), 1 (l2) or 2 (max), got %d", norm); return SPY_ERR_INVALID;
    }
    SPY_LAUNCH_OK();
    return SPY_OK;
}

}  // namespace spy

using namespace spy;

#define SPY_DISPATCH_VI(val_dtype, idx_dtype, CALL)                                               \
    do {                                                                                          \
        if (val_dtype == SPY_F32 && idx_dtype == SPY_I32) { typedef float T; typedef int I; CALL; }             \
        else if (val_dtype == SPY_F32 && idx_dtype == SPY_I64) { typedef float T; typedef long long I; CALL; }  \
        else if (val_dtype == SPY_F64 && idx_dtype == SPY_I32) { typedef double T; typedef int I; CALL; }       \
        else if (val_dtype == SPY_F64 && idx_dtype == SPY_I64) { typedef double T; typedef long long I; CALL; } \
        else { set_error("unsupported dtype combination val=%d idx=%d", val_dtype, idx_dtype); return SPY_ERR_INVALID; } \
    } while (0)

extern "C" {

int spy_normalize_rows_dev(int norm, int64_t n_rows, void *data, int val_dtype, const void *indptr, int idx_dtype,
                           void *stream) {
    SPY_DISPATCH_VI(val_dtype, idx_dtype, return (run_normalize<T, I>(norm, n_rows, (T *)data, (const I *)indptr, as_stream(stream))));
    return SPY_OK;
}

int64_t spy_tfidf_scratch_bytes(int64_t n_rows, int64_t n_cols, int val_dtype) {
    (void)val_dtype;
    return 256 + ((n_rows * 8 + 255) / 256) * 256 + ((n_cols * 8 + 255) / 256) * 256 + n_cols * 4 + 256;
}

int spy_tfidf_dev(int64_t n_rows, int64_t n_cols, void *data, int val_dtype, const void *indices, const void *indptr,
                  int idx_dtype, int tf_mode, int idf_mode, double logbase, void *scratch, void *stream) {
    SPY_REQUIRE(tf_mode >= 0 && tf_mode <= 4 && idf_mode >= 0 && idf_mode <= 4, "bad tf/idf mode");
    SPY_DISPATCH_VI(val_dtype, idx_dtype,
                    return (run_tfidf_bm25<T, I>(false, n_rows, n_cols, (T *)data, (const I *)indices, (const I *)indptr, 0, 0, 0,
                                                 tf_mode, idf_mode, logbase, scratch, as_stream(stream))));
    return SPY_OK;
}

int spy_bm25plus_dev(int64_t n_rows, int64_t n_cols, void *data, int val_dtype, const void *indices, const void *indptr,
                     int idx_dtype, double k1, double b, double delta, int tf_mode, int idf_mode, double logbase,
                     void *scratch, void *stream) {
    SPY_REQUIRE(tf_mode >= 0 && tf_mode <= 4 && idf_mode >= 0 && idf_mode <= 4, "bad tf/idf mode");
    SPY_DISPATCH_VI(val_dtype, idx_dtype,
                    return (run_tfidf_bm25<T, I>(true, n_rows, n_cols, (T *)data, (const I *)indices, (const I *)indptr, k1, b, delta,
                                                 tf_mode, idf_mode, logbase, scratch, as_stream(stream))));
    return SPY_OK;
}

}  // extern "C"
