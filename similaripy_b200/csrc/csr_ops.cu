// csr_ops.cu -- O(nnz) CSR kernels either side of the hot path: norm vectors, transposition,
// zero / column filtering, value casts, scans and output-slab assembly.  sm_100a, HBM-bound:
// coalesced warp-per-row segments, grids sized in multiples of the SM count via grid-stride loops.
//
// Reference code replaced (bogliosimone/similaripy @ a16d939):
//   csr_sum / _build_squared_norms   similaripy/cython_code/s_plus_utils.pyx:128-201
//   _build_cosine_normalization      s_plus_utils.pyx:204-228   (np.power(x + h, c))
//   _build_depop_normalization       s_plus_utils.pyx:231-278
//   _build_matrix_data               s_plus_utils.pyx:281-308   (binary / astype(float32))
//   _filter_matrix_columns           s_plus_utils.pyx:424-490
//   matrix.T.tocsr(), eliminate_zeros  s_plus.pyx:170,205-211   (scipy csr_tocsc / csr_eliminate_zeros)
//   build_coo_matrix / build_csr_matrix / coo_to_csr   utils.pyx:43-173, coo_to_csr.h:28-71
#include "common.cuh"

namespace spy {

constexpr int kThreads = 256;
constexpr int kWarpsPerBlock = kThreads / 32;

static inline int grid_for(long long work_items, int per_block) {
    long long b = (work_items + per_block - 1) / per_block;
    const long long cap = (long long)kB200SmCount * 16;  // 16 resident 256-thread CTAs cover an SM's 2048 threads twice
    if (b > cap) b = cap;
    return (int)(b < 1 ? 1 : b);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- row sums ---------------------------------------------------------------------------
template <bool SQUARE>
__global__ void row_sum_kernel(int n_rows, const int *__restrict__ indptr, const float *__restrict__ data,
                               float *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int warp0 = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const int nwarps = gridDim.x * kWarpsPerBlock;
    for (int r = warp0; r < n_rows; r += nwarps) {
        const int s = indptr[r], e = indptr[r + 1];
        float acc = 0.f;
        for (int q = s + lane; q < e; q += 32) {
            const float v = data[q];
            acc += SQUARE ? __fmul_rn(v, v) : v;
        }
        acc = warp_sum(acc);
        if (lane == 0) out[r] = acc;
    }
}

// ---- work per target row: number of scalar products its expansion performs --------------
// w(t) = sum over u in A[t,:] of nnz(B[u,:]); the unit the row partition across GPUs is balanced by.
__global__ void row_work_kernel(int n_targets, const int *__restrict__ targets, const int *__restrict__ a_indptr,
                                const int *__restrict__ a_indices, const int *__restrict__ b_indptr,
                                long long *__restrict__ work) {
    const int lane = threadIdx.x & 31;
    const int warp0 = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const int nwarps = gridDim.x * kWarpsPerBlock;
    for (int i = warp0; i < n_targets; i += nwarps) {
        const int t = targets[i];
        const int s = a_indptr[t], e = a_indptr[t + 1];
        int acc = 0;  // a row's products fit int32 whenever nnz(B) does
        for (int q = s + lane; q < e; q += 32) {
            const int u = a_indices[q];
            acc += b_indptr[u + 1] - b_indptr[u];
        }
        long long w = acc;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
        if (lane == 0) work[i] = w;
    }
}

// ---- column sums (fp64 accumulation like np.bincount(weights=...)) -----------------------
template <bool SQUARE>
__global__ void col_sum_kernel(long long nnz, const int *__restrict__ indices, const float *__restrict__ data,
                               double *__restrict__ acc) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < nnz; q += stride) {
        const float v = data[q];
        // the reference squares in fp32 (np.square(dtype=float32)) and sums in fp64
        const double w = SQUARE ? (double)__fmul_rn(v, v) : (double)v;
        atomicAdd(acc + indices[q], w);
    }
}
__global__ void f64_to_f32_kernel(long long n, const double *__restrict__ in, float *__restrict__ out) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = (float)in[i];
}

// ---- np.power(float32(x) + shift, p, dtype=float32) -------------------------------------
template <typename T>
__global__ void pow_shift_kernel(long long n, const T *__restrict__ x, float shift, float p, float *__restrict__ out) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float base = __fadd_rn((float)x[i], shift);
        // evaluated in fp64 and rounded once: the closest fp32 to the true power
        out[i] = (float)pow((double)base, (double)p);
    }
}

// ---- column histogram --------------------------------------------------------------------
__global__ void col_count_kernel(long long nnz, const int *__restrict__ indices, int *__restrict__ counts) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < nnz; q += stride)
        atomicAdd(counts + indices[q], 1);
}

// ---- exclusive scan (tiles of 2048, recursive on tile sums) -------------------------------
constexpr int kScanThreads = 512;
constexpr int kScanItems = 4;
constexpr int kScanTile = kScanThreads * kScanItems;

template <typename OUT>
__global__ void scan_tile_kernel(long long n, const int *__restrict__ in, OUT *__restrict__ out,
                                 long long *__restrict__ tile_sums) {
    __shared__ long long warp_tot[kScanThreads / 32];
    const long long base = (long long)blockIdx.x * kScanTile + (long long)threadIdx.x * kScanItems;
    long long v[kScanItems];
    long long local = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; i++) {
        v[i] = (base + i < n) ? (long long)in[base + i] : 0;
        local += v[i];
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    long long incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        long long t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        long long w = (lane < kScanThreads / 32) ? warp_tot[lane] : 0;
        long long wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            long long t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += t;
        }
        if (lane < kScanThreads / 32) warp_tot[lane] = wi - w;  // exclusive warp offsets
        if (lane == kScanThreads / 32 - 1 && tile_sums) tile_sums[blockIdx.x] = wi;
    }
    __syncthreads();
    long long run = warp_tot[warp] + (incl - local);
#pragma unroll
    for (int i = 0; i < kScanItems; i++) {
        if (base + i < n) out[base + i] = (OUT)run;
        run += v[i];
    }
}
// scan of 64-bit tile sums in place (exclusive), single block, sequential over chunks
__global__ void scan_sums_kernel(long long n, long long *__restrict__ sums, long long *__restrict__ total) {
    __shared__ long long warp_tot[32];
    __shared__ long long carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (long long c0 = 0; c0 < n; c0 += blockDim.x) {
        const long long i = c0 + threadIdx.x;
        const long long v = (i < n) ? sums[i] : 0;
        long long incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            long long t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            long long w = warp_tot[lane];
            long long wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                long long t = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += t;
            }
            warp_tot[lane] = wi - w;
        }
        __syncthreads();
        const long long carry = carry_s;
        if (i < n) sums[i] = carry + warp_tot[warp] + incl - v;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry_s = carry + warp_tot[warp] + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0 && total) *total = carry_s;
}
template <typename OUT>
__global__ void scan_add_kernel(long long n, OUT *__restrict__ out, const long long *__restrict__ tile_offsets,
                                const long long *__restrict__ total) {
    const long long base = (long long)blockIdx.x * kScanTile;
    const long long off = tile_offsets[blockIdx.x];
    for (int i = threadIdx.x; i < kScanTile; i += blockDim.x)
        if (base + i < n) out[base + i] = (OUT)((long long)out[base + i] + off);
    if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = (OUT)(*total);
}

template <typename OUT>
static int exclusive_scan(long long n, const int *counts, OUT *offsets, void *tmp, cudaStream_t st) {
    if (n < 0) { set_error("scan length < 0"); return SPY_ERR_INVALID; }
    const long long tiles = (n + kScanTile - 1) / kScanTile;
    long long *tile_sums = reinterpret_cast<long long *>(tmp);
    long long *total = tile_sums + (tiles > 0 ? tiles : 1);
    if (n == 0) {
        SPY_CUDA_OK(cudaMemsetAsync(offsets, 0, sizeof(OUT), st));
        return SPY_OK;
    }
    scan_tile_kernel<OUT><<<(unsigned)tiles, kScanThreads, 0, st>>>(n, counts, offsets, tile_sums);
    SPY_LAUNCH_OK();
    scan_sums_kernel<<<1, 1024, 0, st>>>(tiles, tile_sums, total);
    SPY_LAUNCH_OK();
    scan_add_kernel<OUT><<<(unsigned)tiles, 256, 0, st>>>(n, offsets, tile_sums, total);
    SPY_LAUNCH_OK();
    return SPY_OK;
}

// ---- transpose: scatter + per-row sort ----------------------------------------------------
__global__ void copy_i32_kernel(long long n, const int *__restrict__ in, int *__restrict__ out) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = in[i];
}
__global__ void transpose_scatter_kernel(int n_rows, const int *__restrict__ indptr, const int *__restrict__ indices,
                                         const float *__restrict__ data, int *__restrict__ cursor,
                                         int *__restrict__ t_indices, float *__restrict__ t_data) {
    const int lane = threadIdx.x & 31;
    const int warp0 = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const int nwarps = gridDim.x * kWarpsPerBlock;
    for (int r = warp0; r < n_rows; r += nwarps) {
        const int s = indptr[r], e = indptr[r + 1];
        for (int q = s + lane; q < e; q += 32) {
            const int pos = atomicAdd(cursor + indices[q], 1);
            t_indices[pos] = r;
            t_data[pos] = data[q];
        }
    }
}
// Sort every row by column index, in place.  One CTA per row (grid-stride).  The network is the
// all-ascending bitonic variant (first step of every merge compares i with i ^ (size-1)), so a
// virtual +inf padding beyond the row length never moves and rows of any length are sorted in
// place: in shared memory as packed (index,value) keys up to kSortSmem entries, directly on the
// global arrays beyond that.
constexpr int kSortThreads = 256;
constexpr int kSortSmem = 4096;
typedef unsigned long long u64;

__global__ void __launch_bounds__(kSortThreads)
sort_rows_kernel(int n_rows, const int *__restrict__ indptr, int *__restrict__ indices, float *__restrict__ data) {
    __shared__ u64 keys[kSortSmem];
    for (int r = blockIdx.x; r < n_rows; r += gridDim.x) {
        const int s = indptr[r], e = indptr[r + 1], len = e - s;
        if (len <= 1) continue;  // uniform per CTA
        int unsorted = 0;
        for (int q = s + 1 + threadIdx.x; q < e; q += kSortThreads) unsorted |= (indices[q - 1] > indices[q]);
        if (!__syncthreads_or(unsorted)) continue;  // already ascending
        int S = 2;
        while (S < len) S <<= 1;
        if (len <= kSortSmem) {
            for (int i = threadIdx.x; i < len; i += kSortThreads)
                keys[i] = ((u64)(unsigned)indices[s + i] << 32) | (u64)__float_as_uint(data[s + i]);
            __syncthreads();
            for (int size = 2; size <= S; size <<= 1) {
                // flip step
                for (int i = threadIdx.x; i < len; i += kSortThreads) {
                    const int l = i ^ (size - 1);
                    if (l > i && l < len) {
                        const u64 a = keys[i], b = keys[l];
                        if (a > b) { keys[i] = b; keys[l] = a; }
                    }
                }
                __syncthreads();
                for (int j = size >> 2; j > 0; j >>= 1) {
                    for (int i = threadIdx.x; i < len; i += kSortThreads) {
                        const int l = i ^ j;
                        if (l > i && l < len) {
                            const u64 a = keys[i], b = keys[l];
                            if (a > b) { keys[i] = b; keys[l] = a; }
                        }
                    }
                    __syncthreads();
                }
            }
            for (int i = threadIdx.x; i < len; i += kSortThreads) {
                const u64 kv = keys[i];
                indices[s + i] = (int)(kv >> 32);
                data[s + i] = __uint_as_float((unsigned)(kv & 0xffffffffull));
            }
            __syncthreads();
        } else {
            int *ix = indices + s;
            float *dv = data + s;
            for (int size = 2; size <= S; size <<= 1) {
                for (int i = threadIdx.x; i < len; i += kSortThreads) {
                    const int l = i ^ (size - 1);
                    if (l > i && l < len) {
                        const int a = ix[i], b = ix[l];
                        if (a > b) { ix[i] = b; ix[l] = a; const float t = dv[i]; dv[i] = dv[l]; dv[l] = t; }
                    }
                }
                __syncthreads();
                for (int j = size >> 2; j > 0; j >>= 1) {
                    for (int i = threadIdx.x; i < len; i += kSortThreads) {
                        const int l = i ^ j;
                        if (l > i && l < len) {
                            const int a = ix[i], b = ix[l];
                            if (a > b) { ix[i] = b; ix[l] = a; const float t = dv[i]; dv[i] = dv[l]; dv[l] = t; }
                        }
                    }
                    __syncthreads();
                }
            }
        }
    }
}

// ---- filter (zeros and/or column mask) -----------------------------------------------------
__device__ __forceinline__ bool keep_entry(const int *indices, const float *data, const uint8_t *mask, int drop_zeros, int q) {
    bool k = true;
    if (drop_zeros) k = k && (data[q] != 0.f);
    if (mask) k = k && (mask[indices[q]] != 0);
    return k;
}
__global__ void filter_count_kernel(int n_rows, const int *__restrict__ indptr, const int *__restrict__ indices,
                                    const float *__restrict__ data, const uint8_t *__restrict__ mask, int drop_zeros,
                                    int *__restrict__ row_counts) {
    const int lane = threadIdx.x & 31;
    const int warp0 = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const int nwarps = gridDim.x * kWarpsPerBlock;
    for (int r = warp0; r < n_rows; r += nwarps) {
        const int s = indptr[r], e = indptr[r + 1];
        int c = 0;
        for (int q = s + lane; q < e; q += 32) c += keep_entry(indices, data, mask, drop_zeros, q) ? 1 : 0;
        c = warp_sum_i(c);
        if (lane == 0) row_counts[r] = c;
    }
}
__global__ void filter_compact_kernel(int n_rows, const int *__restrict__ indptr, const int *__restrict__ indices,
                                      const float *__restrict__ data, const uint8_t *__restrict__ mask, int drop_zeros,
                                      const int *__restrict__ new_indptr, int *__restrict__ new_indices,
                                      float *__restrict__ new_data) {
    const int lane = threadIdx.x & 31;
    const int warp0 = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const int nwarps = gridDim.x * kWarpsPerBlock;
    for (int r = warp0; r < n_rows; r += nwarps) {
        const int s = indptr[r], e = indptr[r + 1];
        int dst = new_indptr[r];
        for (int q0 = s; q0 < e; q0 += 32) {
            const int q = q0 + lane;
            const bool k = (q < e) && keep_entry(indices, data, mask, drop_zeros, q);
            const unsigned m = __ballot_sync(0xffffffffu, k);
            if (k) {
                const int pos = dst + __popc(m & ((1u << lane) - 1u));
                new_indices[pos] = indices[q];
                new_data[pos] = data[q];
            }
            dst += __popc(m);
        }
    }
}

// ---- value casts ---------------------------------------------------------------------------
template <typename T>
__global__ void cast_values_kernel(long long n, const T *__restrict__ src, int binary, float *__restrict__ dst) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        dst[i] = binary ? 1.0f : (float)src[i];
}

// ---- output slab ---------------------------------------------------------------------------
__global__ void slab_row_nnz_kernel(int n_targets, int k, const float *__restrict__ values,
                                    const int *__restrict__ counts, const int *__restrict__ targets,
                                    int *__restrict__ row_nnz) {
    const int lane = threadIdx.x & 31;
    const int warp0 = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const int nwarps = gridDim.x * kWarpsPerBlock;
    for (int i = warp0; i < n_targets; i += nwarps) {
        if (targets[i] < 0) continue;  // padding row of a gathered slab (sharded runs)
        const int n = counts[i];
        const float *v = values + (size_t)i * k;
        int c = 0;
        for (int j = lane; j < n; j += 32) c += (v[j] != 0.f) ? 1 : 0;
        c = warp_sum_i(c);
        if (lane == 0) row_nnz[targets[i]] = c;
    }
}
template <typename IDX>
__global__ void slab_compact_kernel(int n_targets, int k, const int *__restrict__ cols, const float *__restrict__ values,
                                    const int *__restrict__ counts, const int *__restrict__ targets,
                                    const long long *__restrict__ csr_indptr, IDX *__restrict__ csr_indices,
                                    float *__restrict__ csr_data) {
    const int lane = threadIdx.x & 31;
    const int warp0 = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const int nwarps = gridDim.x * kWarpsPerBlock;
    for (int i = warp0; i < n_targets; i += nwarps) {
        if (targets[i] < 0) continue;  // padding row of a gathered slab (sharded runs)
        const int n = counts[i];
        const size_t o = (size_t)i * k;
        long long dst = csr_indptr[targets[i]];
        for (int j0 = 0; j0 < n; j0 += 32) {
            const int j = j0 + lane;
            const bool keep = (j < n) && (values[o + j] != 0.f);
            const unsigned m = __ballot_sync(0xffffffffu, keep);
            if (keep) {
                const long long pos = dst + __popc(m & ((1u << lane) - 1u));
                csr_indices[pos] = (IDX)cols[o + j];
                csr_data[pos] = values[o + j];
            }
            dst += __popc(m);
        }
    }
}
__global__ void slab_fill_rows_kernel(int n_targets, int k, const int *__restrict__ targets,
                                      const int *__restrict__ counts, int *__restrict__ rows) {
    const long long total = (long long)n_targets * k;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += stride) {
        const int i = (int)(q / k), j = (int)(q % k);
        rows[q] = (targets[i] >= 0 && j < counts[i]) ? targets[i] : 0;
    }
}


// int64 index arrays (scipy's choice for some large matrices) narrowed on the device: the reference narrows them on the host
// (s_plus.pyx:241-244: astype(int32)), which costs more than uploading the wider array
__global__ void narrow_index_kernel(long long n, const long long *__restrict__ src, int *__restrict__ dst) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = (int)src[i];
}

// ---- matrices beyond int32 stored entries (64-bit indptr): int32-indexed blocks ---------------
// indptr of the block that keeps the stored entries [lo, hi) of a 64-bit CSR and leaves every other row empty:
// out[r] = clamp(indptr[r], lo, hi) - lo.  The block has the shape of the whole matrix and shares its index / value
// arrays (a view at offset lo), so the int32 kernels either side of the hot path run on it unchanged.
__global__ void wide_block_indptr_kernel(long long n, const long long *__restrict__ indptr, long long lo, long long hi,
                                         int *__restrict__ out) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += stride) {
        const long long v = indptr[r];
        out[r] = (int)((v < lo ? lo : (v > hi ? hi : v)) - lo);
    }
}
// indptr of pieces stacked on top of each other (each the full shape, populated on disjoint row ranges, entries
// concatenated in piece order): the element-wise sum of the pieces' indptr arrays
__global__ void indptr_add_kernel(long long n, const int *__restrict__ piece, int *__restrict__ acc) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += stride) acc[r] += piece[r];
}

// ---- best k of two slabs (column blocks of matrix2 computed one after the other) ----------------
// Both slabs hold their rows best-first in the key order of the hot kernels (value descending, then column ascending;
// s_plus.h:45-59 keeps the k largest values), over DISJOINT column sets: an entry's place in the merged row is its own
// position plus the number of entries of the other row that beat it (binary search).
__device__ __forceinline__ unsigned long long merge_key(float v, int col) {
    unsigned u = __float_as_uint(v);
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    return ((unsigned long long)u << 32) | (unsigned long long)(0xffffffffu - (unsigned)col);
}
__global__ void slab_merge_kernel(int n_targets, int k, const int *__restrict__ cols_a, const float *__restrict__ vals_a,
                                  const int *__restrict__ counts_a, const int *__restrict__ cols_b,
                                  const float *__restrict__ vals_b, const int *__restrict__ counts_b,
                                  int *__restrict__ out_cols, float *__restrict__ out_vals, int *__restrict__ out_counts) {
    const long long total = (long long)n_targets * 2 * k;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += stride) {
        const int i = (int)(q / (2 * k)), e = (int)(q % (2 * k));
        const bool from_b = e >= k;
        const int j = from_b ? e - k : e;
        const size_t o = (size_t)i * k;
        const int na = counts_a[i], nb = counts_b[i];
        const int n_out = min(k, na + nb);
        if (j < (from_b ? nb : na)) {
            const int *oc = from_b ? cols_a : cols_b;
            const float *ov = from_b ? vals_a : vals_b;
            const int on = from_b ? na : nb;
            const int col = (from_b ? cols_b : cols_a)[o + j];
            const float val = (from_b ? vals_b : vals_a)[o + j];
            const unsigned long long key = merge_key(val, col);
            int lo = 0, hi = on;  // entries of the other row with a larger key: the first `lo` of them
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (merge_key(ov[o + mid], oc[o + mid]) > key) lo = mid + 1; else hi = mid;
            }
            const int rank = j + lo;
            if (rank < k) { out_cols[o + rank] = col; out_vals[o + rank] = val; }
        }
        if (e >= n_out && e < k) { out_cols[o + e] = 0; out_vals[o + e] = 0.f; }  // padding like the reference's slab (s_plus.pyx:351-353)
        if (e == 0) out_counts[i] = n_out;
    }
}

}  // namespace spy

using namespace spy;

extern "C" {

int spy_csr_row_sum_dev(int32_t n_rows, const int32_t *indptr, const float *data, int square, float *out, void *stream) {
    if (n_rows <= 0) return SPY_OK;
    const int grid = grid_for(n_rows, kWarpsPerBlock);
    if (square) row_sum_kernel<true><<<grid, kThreads, 0, as_stream(stream)>>>(n_rows, indptr, data, out);
    else row_sum_kernel<false><<<grid, kThreads, 0, as_stream(stream)>>>(n_rows, indptr, data, out);
    SPY_LAUNCH_OK();
    return SPY_OK;
}

int spy_knn_row_work_dev(int32_t n_targets, const int32_t *targets, const int32_t *a_indptr, const int32_t *a_indices,
                         const int32_t *b_indptr, int64_t *work, void *stream) {
    if (n_targets <= 0) return SPY_OK;
    SPY_REQUIRE(targets && a_indptr && b_indptr && work, "row_work: NULL pointer");
    row_work_kernel<<<grid_for(n_targets, kWarpsPerBlock), kThreads, 0, as_stream(stream)>>>(
        n_targets, targets, a_indptr, a_indices, b_indptr, reinterpret_cast<long long *>(work));
    SPY_LAUNCH_OK();
    return SPY_OK;
}

int spy_csr_col_sum_dev(int64_t nnz, const int32_t *indices, const float *data, int square, int32_t n_cols,
                        double *acc64, float *out, void *stream) {
    if (n_cols <= 0) return SPY_OK;
    cudaStream_t st = as_stream(stream);
    SPY_CUDA_OK(cudaMemsetAsync(acc64, 0, (size_t)n_cols * sizeof(double), st));
    if (nnz > 0) {
        const int grid = grid_for(nnz, kThreads * 4);
        if (square) col_sum_kernel<true><<<grid, kThreads, 0, st>>>(nnz, indices, data, acc64);
        else col_sum_kernel<false><<<grid, kThreads, 0, st>>>(nnz, indices, data, acc64);
        SPY_LAUNCH_OK();
    }
    f64_to_f32_kernel<<<grid_for(n_cols, kThreads), kThreads, 0, st>>>(n_cols, acc64, out);
    SPY_LAUNCH_OK();
    return SPY_OK;
}

int spy_pow_shift_dev(int64_t n, const void *x, int x_dtype, float shift, float p, float *out, void *stream) {
    if (n <= 0) return SPY_OK;
    const int grid = grid_for(n, kThreads);
    if (x_dtype == SPY_F32) pow_shift_kernel<float><<<grid, kThreads, 0, as_stream(stream)>>>(n, (const float *)x, shift, p, out);
    else if (x_dtype == SPY_F64) pow_shift_kernel<double><<<grid, kThreads, 0, as_stream(stream)>>>(n, (const double *)x, shift, p, out);
    else { set_error("pow_shift: unsupported dtype %d", x_dtype); return SPY_ERR_INVALID; }
    SPY_LAUNCH_OK();
    return SPY_OK;
}

int spy_csr_col_count_dev(int64_t nnz, const int32_t *indices, int32_t n_cols, int32_t *counts, void *stream) {
    cudaStream_t st = as_stream(stream);
    if (n_cols > 0) SPY_CUDA_OK(cudaMemsetAsync(counts, 0, (size_t)n_cols * sizeof(int), st));
    if (nnz <= 0) return SPY_OK;
    col_count_kernel<<<grid_for(nnz, kThreads * 4), kThreads, 0, st>>>(nnz, indices, counts);
    SPY_LAUNCH_OK();
    return SPY_OK;
}

int64_t spy_scan_tmp_bytes(int64_t n) {
    const int64_t tiles = (n + kScanTile - 1) / kScanTile;
    return ((tiles > 0 ? tiles : 1) + 1) * 8;
}
int spy_exclusive_scan_i32_dev(int64_t n, const int32_t *counts, int32_t *offsets, void *tmp, void *stream) {
    return exclusive_scan<int>(n, counts, offsets, tmp, as_stream(stream));
}
int spy_exclusive_scan_i64_dev(int64_t n, const int32_t *counts, int64_t *offsets, void *tmp, void *stream) {
    return exclusive_scan<long long>(n, counts, (long long *)offsets, tmp, as_stream(stream));
}

int spy_csr_transpose_dev(int32_t n_rows, int32_t n_cols, const int32_t *indptr, const int32_t *indices,
                          const float *data, const int32_t *t_indptr, int32_t *t_indices, float *t_data,
                          int32_t *cursor, int sort_rows, void *stream) {
    if (n_rows <= 0 || n_cols <= 0) return SPY_OK;
    cudaStream_t st = as_stream(stream);
    copy_i32_kernel<<<grid_for(n_cols, kThreads), kThreads, 0, st>>>(n_cols, t_indptr, cursor);
    SPY_LAUNCH_OK();
    transpose_scatter_kernel<<<grid_for(n_rows, kWarpsPerBlock), kThreads, 0, st>>>(n_rows, indptr, indices, data, cursor,
                                                                                   t_indices, t_data);
    SPY_LAUNCH_OK();
    if (sort_rows) {  // the scatter lands in arrival order
        int grid = n_cols < kB200SmCount * 8 ? n_cols : kB200SmCount * 8;
        sort_rows_kernel<<<grid, kSortThreads, 0, st>>>(n_cols, t_indptr, t_indices, t_data);
        SPY_LAUNCH_OK();
    }
    return SPY_OK;
}

int spy_csr_sort_rows_dev(int32_t n_rows, const int32_t *indptr, int32_t *indices, float *data, void *stream) {
    if (n_rows <= 0) return SPY_OK;
    int grid = n_rows < kB200SmCount * 8 ? n_rows : kB200SmCount * 8;
    sort_rows_kernel<<<grid, kSortThreads, 0, as_stream(stream)>>>(n_rows, indptr, indices, data);
    SPY_LAUNCH_OK();
    return SPY_OK;
}

int spy_csr_filter_count_dev(int32_t n_rows, const int32_t *indptr, const int32_t *indices, const float *data,
                             const uint8_t *col_mask, int drop_zeros, int32_t *row_counts, void *stream) {
    if (n_rows <= 0) return SPY_OK;
    filter_count_kernel<<<grid_for(n_rows, kWarpsPerBlock), kThreads, 0, as_stream(stream)>>>(n_rows, indptr, indices, data,
                                                                                             col_mask, drop_zeros, row_counts);
    SPY_LAUNCH_OK();
    return SPY_OK;
}
int spy_csr_filter_compact_dev(int32_t n_rows, const int32_t *indptr, const int32_t *indices, const float *data,
                               const uint8_t *col_mask, int drop_zeros, const int32_t *new_indptr,
                               int32_t *new_indices, float *new_data, void *stream) {
    if (n_rows <= 0) return SPY_OK;
    filter_compact_kernel<<<grid_for(n_rows, kWarpsPerBlock), kThreads, 0, as_stream(stream)>>>(
        n_rows, indptr, indices, data, col_mask, drop_zeros, new_indptr, new_indices, new_data);
    SPY_LAUNCH_OK();
    return SPY_OK;
}

int spy_cast_values_dev(int64_t n, const void *src, int src_dtype, int binary, float *dst, void *stream) {
    if (n <= 0) return SPY_OK;
    const int grid = grid_for(n, kThreads * 2);
    cudaStream_t st = as_stream(stream);
    switch (src_dtype) {
    case SPY_F32: cast_values_kernel<float><<<grid, kThreads, 0, st>>>(n, (const float *)src, binary, dst); break;
    case SPY_F64: cast_values_kernel<double><<<grid, kThreads, 0, st>>>(n, (const double *)src, binary, dst); break;
    case SPY_VAL_I32: cast_values_kernel<int><<<grid, kThreads, 0, st>>>(n, (const int *)src, binary, dst); break;
    case SPY_VAL_I64: cast_values_kernel<long long><<<grid, kThreads, 0, st>>>(n, (const long long *)src, binary, dst); break;
    default: set_error("cast_values: unsupported dtype %d", src_dtype); return SPY_ERR_INVALID;
    }
    SPY_LAUNCH_OK();
    return SPY_OK;
}

int spy_slab_row_nnz_dev(int32_t n_targets, int32_t k, const float *values, const int32_t *counts,
                         const int32_t *targets, int32_t *row_nnz, void *stream) {
    if (n_targets <= 0) return SPY_OK;
    slab_row_nnz_kernel<<<grid_for(n_targets, kWarpsPerBlock), kThreads, 0, as_stream(stream)>>>(n_targets, k, values, counts,
                                                                                                targets, row_nnz);
    SPY_LAUNCH_OK();
    return SPY_OK;
}

int spy_slab_compact_dev(int32_t n_targets, int32_t k, const int32_t *cols, const float *values, const int32_t *counts,
                         const int32_t *targets, const int64_t *csr_indptr, void *csr_indices, int idx_dtype,
                         float *csr_data, void *stream) {
    if (n_targets <= 0) return SPY_OK;
    const int grid = grid_for(n_targets, kWarpsPerBlock);
    if (idx_dtype == SPY_I32)
        slab_compact_kernel<int><<<grid, kThreads, 0, as_stream(stream)>>>(n_targets, k, cols, values, counts, targets,
                                                                          (const long long *)csr_indptr, (int *)csr_indices, csr_data);
    else
        slab_compact_kernel<long long><<<grid, kThreads, 0, as_stream(stream)>>>(n_targets, k, cols, values, counts, targets,
                                                                                (const long long *)csr_indptr,
                                                                                (long long *)csr_indices, csr_data);
    SPY_LAUNCH_OK();
    return SPY_OK;
}

int spy_slab_fill_rows_dev(int32_t n_targets, int32_t k, const int32_t *targets, const int32_t *counts, int32_t *rows,
                           void *stream) {
    if (n_targets <= 0) return SPY_OK;
    slab_fill_rows_kernel<<<grid_for((long long)n_targets * k, kThreads * 4), kThreads, 0, as_stream(stream)>>>(n_targets, k, targets,
                                                                                                               counts, rows);
    SPY_LAUNCH_OK();
    return SPY_OK;
}

int spy_narrow_index_dev(int64_t n, const int64_t *src, int32_t *dst, void *stream) {
    if (n <= 0) return SPY_OK;
    SPY_REQUIRE(src && dst, "narrow_index: NULL pointer");
    narrow_index_kernel<<<grid_for(n, kThreads * 4), kThreads, 0, as_stream(stream)>>>(n, reinterpret_cast<const long long *>(src), dst);
    SPY_LAUNCH_OK();
    return SPY_OK;
}

int spy_csr_wide_block_indptr_dev(int64_t n, const int64_t *indptr, int64_t lo, int64_t hi, int32_t *out, void *stream) {
    if (n <= 0) return SPY_OK;
    SPY_REQUIRE(indptr && out, "wide_block_indptr: NULL pointer");
    SPY_REQUIRE(lo >= 0 && hi >= lo && hi - lo <= 2147483647LL, "wide_block_indptr: a block holds at most 2^31-1 entries");
    wide_block_indptr_kernel<<<grid_for(n, kThreads), kThreads, 0, as_stream(stream)>>>(
        n, reinterpret_cast<const long long *>(indptr), lo, hi, out);
    SPY_LAUNCH_OK();
    return SPY_OK;
}

int spy_csr_indptr_add_dev(int64_t n, const int32_t *piece, int32_t *acc, void *stream) {
    if (n <= 0) return SPY_OK;
    SPY_REQUIRE(piece && acc, "indptr_add: NULL pointer");
    indptr_add_kernel<<<grid_for(n, kThreads), kThreads, 0, as_stream(stream)>>>(n, piece, acc);
    SPY_LAUNCH_OK();
    return SPY_OK;
}

int spy_slab_merge_dev(int32_t n_targets, int32_t k, const int32_t *cols_a, const float *vals_a, const int32_t *counts_a,
                       const int32_t *cols_b, const float *vals_b, const int32_t *counts_b, int32_t *out_cols,
                       float *out_vals, int32_t *out_counts, void *stream) {
    if (n_targets <= 0 || k <= 0) return SPY_OK;
    SPY_REQUIRE(cols_a && vals_a && counts_a && cols_b && vals_b && counts_b && out_cols && out_vals && out_counts,
                "slab_merge: NULL pointer");
    slab_merge_kernel<<<grid_for((long long)n_targets * 2 * k, kThreads), kThreads, 0, as_stream(stream)>>>(
        n_targets, k, cols_a, vals_a, counts_a, cols_b, vals_b, counts_b, out_cols, out_vals, out_counts);
    SPY_LAUNCH_OK();
    return SPY_OK;
}

}  // extern "C"
