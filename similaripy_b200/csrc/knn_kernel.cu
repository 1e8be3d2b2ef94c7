// knn_kernel.cu -- the hot path: CSR x CSR row-expansion SpGEMM with the similarity
// denominator fused in and per-row top-k selection, for sm_100a (B200).
//
// Replaces s_plus::compute_similarities_parallel<int,float> (reference
// similaripy/cython_code/s_plus.h:265-453) behind the C ABI in include/similaripy_b200.h.
//
// Design:
//   * one persistent CTA per resident slot; CTAs pull target rows from an atomic queue
//     (the reference's `omp for schedule(dynamic)`, s_plus.h:337);
//   * the output columns are cut into panels of `panel_width` columns; a panel of fp32
//     partial sums lives in shared memory (the reference's `sums` buffer, s_plus.h:91,
//     at shared-memory instead of L2-cache scale, s_plus.h:305-311);
//   * panel boundaries inside every sorted row of B are precomputed once (b_split) --
//     the reference does a std::lower_bound per (target row, block, B row), s_plus.h:381-394;
//   * sub-warp groups of G lanes stream one B-row segment each with coalesced loads and
//     accumulate with shared-memory float atomics (measured 538-605 Gproducts/s on B200,
//     profiles/microbench/accum_bench_r01.txt);
//   * "touched" is encoded in the accumulator itself: slots start at -0.0f, and
//     (-0.0f) + x == x, so a slot whose bits are still 0x80000000 was never written
//     (the reference keeps a touched-list, s_plus.h:112-117);
//   * the drain applies filter / target selectors, computeSimilarity (s_plus.h:129-156) with
//     the same operation order and no FMA contraction, the threshold test (s_plus.h:206), and
//     feeds a running-threshold candidate buffer; a bitonic sort compacts it to the best k
//     whenever it fills and once at the end of the row;
//   * ties are resolved deterministically: larger value first, then smaller column id.
#include "common.cuh"
#include <algorithm>
#include <vector>

namespace spy {

typedef unsigned long long u64;

constexpr unsigned kSentinelBits = 0x80000000u;  // -0.0f : "slot never written"
constexpr int kChunk = 1024;                     // A-row entries staged per pass

struct KnnDev {
    int n_targets;
    const int *targets;
    const int *row_order;
    const int *a_indptr, *a_indices;
    const float *a_data;
    const int *b_indptr, *b_indices;
    const float *b_data;
    const int *b_split;
    int split_stride, n_panels, W, n_cols;
    const float *Xt, *Yt, *Xc, *Yc, *Xd, *Yd;
    float a1, l1, l2, l3, t1, t2, stab, bayes, thr;
    int k, cap;
    int filter_mode;
    const int *f_indptr, *f_indices;
    int target_mode;
    const int *t_indptr, *t_indices;
    int *out_rows, *out_cols;
    float *out_vals;
    int *out_counts;
    int *work_counter;
    u64 *cand_global;
};

// ---- key packing: (value, column) -> 64-bit key whose unsigned order is
//      "value descending, then column ascending" when sorted descending. ----
__device__ __forceinline__ unsigned ordered_bits(float v) {
    unsigned u = __float_as_uint(v);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float unordered_bits(unsigned o) {
    unsigned u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
    return __uint_as_float(u);
}
__device__ __forceinline__ u64 make_key(float v, int col) {
    return ((u64)ordered_bits(v) << 32) | (u64)(0xffffffffu - (unsigned)col);
}

// computeSimilarity (s_plus.h:129-156): same expression order, explicit _rn intrinsics so that
// nvcc cannot contract mul+add into FMA (the reference's x86-64 build has no FMA).
struct SimRow {
    float Xt, Xc, Xd;
};
__device__ __forceinline__ float similarity_value(const KnnDev &p, const SimRow &r, int col, float xy) {
    float vT = 0.f, vC = 0.f, vD = 0.f, val = xy;
    if (p.l1 != 0.f) {
        float a = __fmul_rn(p.t1, __fsub_rn(r.Xt, xy));
        float b = __fmul_rn(p.t2, __fsub_rn(__ldg(p.Yt + col), xy));
        vT = __fmul_rn(p.l1, __fadd_rn(__fadd_rn(a, b), xy));
    }
    if (p.l2 != 0.f) vC = __fmul_rn(p.l2, __fmul_rn(r.Xc, __ldg(p.Yc + col)));
    if (p.l3 != 0.f) vD = __fmul_rn(p.l3, __fmul_rn(r.Xd, __ldg(p.Yd + col)));
    if (p.a1 != 1.f) xy = powf(xy, p.a1);
    if (p.l1 != 0.f || p.l2 != 0.f || p.l3 != 0.f || p.stab != 0.f || p.bayes != 0.f) {
        float den = __fadd_rn(__fadd_rn(__fadd_rn(vT, vC), vD), p.stab);
        val = (den != 0.f) ? __fdiv_rn(xy, den) : 0.f;
        if (p.bayes != 0.f) val = __fmul_rn(val, __fdiv_rn(xy, __fadd_rn(xy, p.bayes)));
    }
    return val;
}

__device__ __forceinline__ int lower_bound_dev(const int *a, int lo, int hi, int x) {
    while (lo < hi) {
        int mid = lo + ((hi - lo) >> 1);
        if (__ldg(a + mid) < x) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// Bitonic sort of S (power of two) keys, descending.  All NT threads participate.
template <int NT>
__device__ void bitonic_sort_desc(u64 *cand, int S) {
    const int tid = threadIdx.x;
    for (int size = 2; size <= S; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = tid; i < (S >> 1); i += NT) {
                int lo = ((i & ~(stride - 1)) << 1) | (i & (stride - 1));
                int hi = lo + stride;
                bool desc = ((lo & size) == 0) || (size == S);
                u64 a = cand[lo], b = cand[hi];
                if ((a < b) == desc) { cand[lo] = b; cand[hi] = a; }
            }
            __syncthreads();
        }
    }
}

// Keep the best k of the n = min(*s_cnt, cap) buffered keys (sorted, best first) and raise tau.
// Entered and left with all threads converged; ends with a barrier.
template <int NT>
__device__ void select_topk(u64 *cand, int *s_cnt, u64 *s_tau, int cap, int k) {
    const int tid = threadIdx.x;
    int n = min(*s_cnt, cap);
    int S = 2;
    while (S < n) S <<= 1;
    for (int i = n + tid; i < S; i += NT) cand[i] = 0ull;
    __syncthreads();
    bitonic_sort_desc<NT>(cand, S);
    if (tid == 0) {
        *s_cnt = min(n, k);
        if (n >= k) *s_tau = cand[k - 1];
    }
    __syncthreads();
}

// Warp-aggregated append of up to 4 keys per thread; returns true when a wanted key found the
// buffer full (it stays in keys[] for the retry after a select).
__device__ __forceinline__ bool append_keys(u64 (&keys)[4], u64 tau, u64 *cand, int cap, int *s_cnt) {
    const unsigned lane = threadIdx.x & 31u;
    bool fail = false;
#pragma unroll
    for (int r = 0; r < 4; r++) {
        bool want = keys[r] > tau;
        unsigned m = __ballot_sync(0xffffffffu, want);
        if (m) {
            int leader = __ffs(m) - 1;
            int base = 0;
            if ((int)lane == leader) base = atomicAdd(s_cnt, __popc(m));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (want) {
                int pos = base + __popc(m & ((1u << lane) - 1u));
                if (pos < cap) { cand[pos] = keys[r]; keys[r] = 0ull; }
                else fail = true;
            }
        }
        if (!want) keys[r] = 0ull;  // at or below the running threshold: dropped for good
    }
    return fail;
}

template <int NT, int G, bool CAND_SMEM>
__global__ void __launch_bounds__(NT, 1024 / NT)
knn_panel_kernel(const __grid_constant__ KnnDev p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *acc = reinterpret_cast<float *>(smem_raw);
    unsigned char *ptr = smem_raw + (size_t)p.W * sizeof(float);
    u64 *cand;
    if (CAND_SMEM) { cand = reinterpret_cast<u64 *>(ptr); ptr += (size_t)p.cap * sizeof(u64); }
    else cand = p.cand_global + (size_t)blockIdx.x * p.cap;
    int *seg_start = reinterpret_cast<int *>(ptr);
    int *seg_end = seg_start + kChunk;
    float *seg_v = reinterpret_cast<float *>(seg_end + kChunk);

    __shared__ int s_row, s_cnt, s_any;
    __shared__ u64 s_tau;

    const int tid = threadIdx.x;
    constexpr int NG = NT / G;
    const int lg = tid & (G - 1);
    const int gid = tid / G;
    const float sentinel = __uint_as_float(kSentinelBits);

    for (int i = tid; i < p.W; i += NT) acc[i] = sentinel;

    for (;;) {
        __syncthreads();
        if (tid == 0) { s_row = atomicAdd(p.work_counter, 1); s_cnt = 0; s_tau = 0ull; }
        __syncthreads();
        const int slot = s_row;
        if (slot >= p.n_targets) break;
        const int i_out = p.row_order ? __ldg(p.row_order + slot) : slot;
        const int t = __ldg(p.targets + i_out);
        const int a0 = __ldg(p.a_indptr + t), a1 = __ldg(p.a_indptr + t + 1);
        SimRow sr;
        sr.Xt = (p.l1 != 0.f) ? __ldg(p.Xt + t) : 0.f;
        sr.Xc = (p.l2 != 0.f) ? __ldg(p.Xc + t) : 0.f;
        sr.Xd = (p.l3 != 0.f) ? __ldg(p.Xd + t) : 0.f;
        u64 tau = 0ull;

        for (int pn = 0; pn < p.n_panels; pn++) {
            const int base = pn * p.W;
            const int width = min(p.W, p.n_cols - base);
            if (tid == 0) s_any = 0;
            // ---------------- expand + accumulate (s_plus.h:358-403 / 418-438) ----------------
            for (int c0 = a0; c0 < a1; c0 += kChunk) {
                const int n = min(kChunk, a1 - c0);
                __syncthreads();  // previous chunk fully consumed (and s_any reset visible)
                int any = 0;
                for (int j = tid; j < n; j += NT) {
                    const int u = __ldg(p.a_indices + c0 + j);
                    int s, e;
                    if (p.n_panels == 1) { s = __ldg(p.b_indptr + u); e = __ldg(p.b_indptr + u + 1); }
                    else {
                        const int *sp = p.b_split + (size_t)u * p.split_stride + pn;
                        s = __ldg(sp); e = __ldg(sp + 1);
                    }
                    seg_start[j] = s; seg_end[j] = e; seg_v[j] = __ldg(p.a_data + c0 + j);
                    any |= (e > s);
                }
                if (any) s_any = 1;
                __syncthreads();
                for (int j = gid; j < n; j += NG) {
                    const int s = seg_start[j], e = seg_end[j];
                    const float v1 = seg_v[j];
                    for (int q = s + lg; q < e; q += 4 * G) {
                        int c[4]; float w[4];
#pragma unroll
                        for (int r = 0; r < 4; r++) {
                            const int qq = q + r * G;
                            const bool ok = qq < e;
                            c[r] = ok ? __ldg(p.b_indices + qq) : -1;
                            w[r] = ok ? __ldg(p.b_data + qq) : 0.f;
                        }
#pragma unroll
                        for (int r = 0; r < 4; r++)
                            if (c[r] >= 0) atomicAdd(&acc[c[r] - base], __fmul_rn(w[r], v1));
                    }
                }
            }
            __syncthreads();
            if (!s_any) continue;  // nothing landed in this panel: accumulator still clean

            // ---------------- per-row filter matrix: erase filtered columns (s_plus.h:159-172) --
            if (p.filter_mode == SPY_SEL_MATRIX) {
                const int fs = __ldg(p.f_indptr + t), fe = __ldg(p.f_indptr + t + 1);
                const int lo = lower_bound_dev(p.f_indices, fs, fe, base);
                const int hi = lower_bound_dev(p.f_indices, lo, fe, base + width);
                for (int q = lo + tid; q < hi; q += NT) acc[__ldg(p.f_indices + q) - base] = sentinel;
                __syncthreads();
            }

            // ---------------- drain: similarity, threshold, top-k (s_plus.h:193-215) ----------
            if (p.target_mode == SPY_SEL_MATRIX) {
                // only columns listed in the target row can be candidates (s_plus.h:175-188)
                const int ts = __ldg(p.t_indptr + t), te = __ldg(p.t_indptr + t + 1);
                const int lo = lower_bound_dev(p.t_indices, ts, te, base);
                const int hi = lower_bound_dev(p.t_indices, lo, te, base + width);
                for (int it0 = lo; it0 < hi; it0 += NT * 4) {
                    u64 keys[4];
#pragma unroll
                    for (int r = 0; r < 4; r++) {
                        keys[r] = 0ull;
                        const int q = it0 + r * NT + tid;
                        if (q < hi) {
                            const int col = __ldg(p.t_indices + q);
                            const bool dup = (q > lo) && (__ldg(p.t_indices + q - 1) == col);
                            const float xy = acc[col - base];
                            if (!dup && __float_as_uint(xy) != kSentinelBits) {
                                const float val = similarity_value(p, sr, col, xy);
                                if (val >= p.thr) keys[r] = make_key(val, col);
                            }
                        }
                    }
                    bool fail = append_keys(keys, tau, cand, p.cap, &s_cnt);
                    while (__syncthreads_or(fail)) {
                        select_topk<NT>(cand, &s_cnt, &s_tau, p.cap, p.k);
                        tau = s_tau;
                        fail = append_keys(keys, tau, cand, p.cap, &s_cnt);
                    }
                }
                __syncthreads();
                for (int i = tid; i < width; i += NT) acc[i] = sentinel;
            } else {
                for (int it0 = 0; it0 < width; it0 += NT * 4) {
                    u64 keys[4];
#pragma unroll
                    for (int r = 0; r < 4; r++) {
                        keys[r] = 0ull;
                        const int idx = it0 + r * NT + tid;
                        if (idx < width) {
                            const float xy = acc[idx];
                            if (__float_as_uint(xy) != kSentinelBits) {
                                acc[idx] = sentinel;
                                const int col = base + idx;
                                const float val = similarity_value(p, sr, col, xy);
                                if (val >= p.thr) keys[r] = make_key(val, col);
                            }
                        }
                    }
                    bool fail = append_keys(keys, tau, cand, p.cap, &s_cnt);
                    while (__syncthreads_or(fail)) {
                        select_topk<NT>(cand, &s_cnt, &s_tau, p.cap, p.k);
                        tau = s_tau;
                        fail = append_keys(keys, tau, cand, p.cap, &s_cnt);
                    }
                }
            }
        }

        // ---------------- final selection and slab write (s_plus.h:443-450) ----------------
        __syncthreads();
        select_topk<NT>(cand, &s_cnt, &s_tau, p.cap, p.k);
        const int n_out = s_cnt;
        const size_t o = (size_t)i_out * (size_t)p.k;
        for (int j = tid; j < p.k; j += NT) {
            int col = 0; float val = 0.f; int row = 0;
            if (j < n_out) {
                const u64 key = cand[j];
                col = (int)(0xffffffffu - (unsigned)(key & 0xffffffffull));
                val = unordered_bits((unsigned)(key >> 32));
                row = t;
            }
            p.out_cols[o + j] = col;
            p.out_vals[o + j] = val;
            if (p.out_rows) p.out_rows[o + j] = row;
        }
        if (tid == 0 && p.out_counts) p.out_counts[i_out] = n_out;
    }
}

// split[u*stride + pn] = first q in row u with b_indices[q] >= pn*W
__global__ void build_split_kernel(int b_rows, const int *__restrict__ b_indptr, const int *__restrict__ b_indices,
                                   int W, int n_panels, int stride, int *__restrict__ split) {
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)b_rows * (n_panels + 1);
    if (gtid >= total) return;
    const int u = (int)(gtid / (n_panels + 1));
    const int pn = (int)(gtid % (n_panels + 1));
    const int s = b_indptr[u], e = b_indptr[u + 1];
    int r;
    if (pn == 0) r = s;
    else if (pn == n_panels) r = e;
    else r = lower_bound_dev(b_indices, s, e, (int)min((long long)pn * W, (long long)0x7fffffff));
    split[(size_t)u * stride + pn] = r;
}

// ------------------------------------------------------------------------------------------
// host side: planning and launch
// ------------------------------------------------------------------------------------------
static int next_pow2(int x) { int p = 1; while (p < x) p <<= 1; return p; }

struct Plan {
    int threads, ctas_per_sm, lanes, W, n_panels, split_stride, cap;
    bool cand_smem;
    size_t smem_bytes;
};

static int make_plan(const spy_knn_args &a, double avg_b_row_nnz, int device, Plan &pl) {
    DeviceInfo di = device_info(device);
    pl.threads = a.threads ? a.threads : 512;
    if (pl.threads != 256 && pl.threads != 512 && pl.threads != 1024) {
        set_error("threads must be 256, 512 or 1024 (got %d)", pl.threads);
        return SPY_ERR_INVALID;
    }
    pl.ctas_per_sm = 1024 / pl.threads;
    pl.cap = std::max(2048, next_pow2(2 * std::max(a.k, 1)));
    pl.cand_smem = (size_t)pl.cap * 8 <= 65536;
    const size_t fixed = (size_t)kChunk * 12 + (pl.cand_smem ? (size_t)pl.cap * 8 : 0);
    // 1 KB per CTA is reserved by the driver; keep a little slack for static shared memory
    const size_t budget = (size_t)di.max_smem_optin / pl.ctas_per_sm - 1024 - 64;
    if (budget <= fixed + 128 * 4) {
        set_error("k=%d leaves no shared memory for the accumulator", a.k);
        return SPY_ERR_UNSUPPORTED;
    }
    const int w_max = (int)((budget - fixed) / 4 / 128) * 128;
    const int n_cols = std::max(a.n_cols, 1);
    int W = a.panel_width;
    if (W <= 0) {
        int P = ceil_div(n_cols, w_max);
        W = ceil_div(ceil_div(n_cols, P), 128) * 128;
    } else {
        if (W % 128 != 0 || W > w_max) {
            set_error("panel_width must be a multiple of 128 and <= %d (got %d)", w_max, W);
            return SPY_ERR_INVALID;
        }
    }
    pl.W = W;
    pl.n_panels = ceil_div(n_cols, W);
    int stride = pl.n_panels + 1;
    if (stride <= 8) stride = next_pow2(stride);  // one 32-byte sector per B row
    pl.split_stride = (pl.n_panels > 1) ? stride : 0;
    pl.smem_bytes = (size_t)W * 4 + fixed;
    int lanes = a.lanes_per_segment;
    if (lanes <= 0) {
        const double seg = avg_b_row_nnz / pl.n_panels;
        lanes = seg <= 10 ? 4 : seg <= 40 ? 8 : seg <= 96 ? 16 : 32;
    }
    if (lanes != 4 && lanes != 8 && lanes != 16 && lanes != 32) {
        set_error("lanes_per_segment must be 4, 8, 16 or 32 (got %d)", lanes);
        return SPY_ERR_INVALID;
    }
    pl.lanes = lanes;
    return SPY_OK;
}

typedef void (*knn_kernel_t)(const KnnDev);

template <int NT, int G>
static knn_kernel_t pick_cand(bool cand_smem) {
    return cand_smem ? (knn_kernel_t)knn_panel_kernel<NT, G, true> : (knn_kernel_t)knn_panel_kernel<NT, G, false>;
}
template <int NT>
static knn_kernel_t pick_lanes(int lanes, bool cand_smem) {
    switch (lanes) {
    case 4: return pick_cand<NT, 4>(cand_smem);
    case 8: return pick_cand<NT, 8>(cand_smem);
    case 16: return pick_cand<NT, 16>(cand_smem);
    default: return pick_cand<NT, 32>(cand_smem);
    }
}
static knn_kernel_t pick_kernel(int threads, int lanes, bool cand_smem) {
    switch (threads) {
    case 256: return pick_lanes<256>(lanes, cand_smem);
    case 1024: return pick_lanes<1024>(lanes, cand_smem);
    default: return pick_lanes<512>(lanes, cand_smem);
    }
}

static int grid_size(const Plan &pl, int n_targets, int device) {
    DeviceInfo di = device_info(device);
    long long g = (long long)di.sm_count * pl.ctas_per_sm;
    if (g > n_targets) g = n_targets;
    return (int)std::max(1LL, g);
}

}  // namespace spy

using namespace spy;

extern "C" {

int spy_knn_plan(spy_knn_args *args, double avg_b_row_nnz, int device) {
    SPY_REQUIRE(args != nullptr, "args is NULL");
    SPY_REQUIRE(args->k >= 1, "k must be >= 1 (got %d)", args->k);
    SPY_REQUIRE(args->n_cols >= 0, "n_cols must be >= 0");
    Plan pl;
    int rc = make_plan(*args, avg_b_row_nnz, device, pl);
    if (rc != SPY_OK) return rc;
    args->threads = pl.threads;
    args->lanes_per_segment = pl.lanes;
    args->panel_width = pl.W;
    args->n_panels = pl.n_panels;
    args->split_stride = pl.split_stride;
    return SPY_OK;
}

int64_t spy_knn_scratch_bytes(const spy_knn_args *args, int device) {
    if (!args) return SPY_ERR_INVALID;
    Plan pl;
    if (make_plan(*args, 32.0, device, pl) != SPY_OK) return SPY_ERR_INVALID;
    int64_t bytes = 256;  // work counter
    if (!pl.cand_smem) bytes += (int64_t)grid_size(pl, std::max(args->n_targets, 1), device) * pl.cap * 8;
    return bytes;
}

int spy_knn_build_split_dev(int32_t b_rows, const int32_t *b_indptr, const int32_t *b_indices,
                            int32_t panel_width, int32_t n_panels, int32_t split_stride,
                            int32_t *split_out, void *stream) {
    SPY_REQUIRE(n_panels >= 1 && split_stride >= n_panels + 1, "bad split geometry");
    if (b_rows <= 0) return SPY_OK;
    const long long total = (long long)b_rows * (n_panels + 1);
    const int threads = 256;
    const long long blocks = (total + threads - 1) / threads;
    build_split_kernel<<<(unsigned)blocks, threads, 0, as_stream(stream)>>>(b_rows, b_indptr, b_indices, panel_width,
                                                                           n_panels, split_stride, split_out);
    SPY_LAUNCH_OK();
    return SPY_OK;
}

int spy_knn_topk_dev(const spy_knn_args *args, void *scratch, int64_t scratch_bytes, void *stream) {
    SPY_REQUIRE(args != nullptr, "args is NULL");
    const spy_knn_args &a = *args;
    SPY_REQUIRE(a.k >= 1, "k must be >= 1 (got %d)", a.k);
    SPY_REQUIRE(a.n_targets >= 0, "n_targets must be >= 0");
    if (a.n_targets == 0) return SPY_OK;
    SPY_REQUIRE(a.out_cols && a.out_values, "output slab pointers are NULL");
    SPY_REQUIRE(a.panel_width > 0 && a.n_panels > 0 && a.threads > 0 && a.lanes_per_segment > 0,
                "launch plan missing: call spy_knn_plan first");
    SPY_REQUIRE(a.n_panels == 1 || a.b_split != nullptr, "n_panels > 1 needs b_split (spy_knn_build_split_dev)");
    SPY_REQUIRE((long long)a.panel_width * a.n_panels >= a.n_cols, "panels do not cover n_cols");
    SPY_REQUIRE(a.filter_mode != SPY_SEL_MATRIX || (a.filter_indptr && a.filter_indices), "filter matrix is NULL");
    SPY_REQUIRE(a.target_mode != SPY_SEL_MATRIX || (a.target_indptr && a.target_indices), "target matrix is NULL");
    SPY_REQUIRE(a.l1 == 0.f || (a.Xtversky && a.Ytversky), "l1 != 0 needs Xtversky/Ytversky");
    SPY_REQUIRE(a.l2 == 0.f || (a.Xcosine && a.Ycosine), "l2 != 0 needs Xcosine/Ycosine");
    SPY_REQUIRE(a.l3 == 0.f || (a.Xdepop && a.Ydepop), "l3 != 0 needs Xdepop/Ydepop");

    int device = 0;
    SPY_CUDA_OK(cudaGetDevice(&device));
    Plan pl;
    int rc = make_plan(a, 32.0, device, pl);
    if (rc != SPY_OK) return rc;
    const int grid = grid_size(pl, a.n_targets, device);
    int64_t need = 256 + (pl.cand_smem ? 0 : (int64_t)grid * pl.cap * 8);
    SPY_REQUIRE(scratch != nullptr && scratch_bytes >= need, "scratch too small: need %lld bytes", (long long)need);

    KnnDev d;
    d.n_targets = a.n_targets; d.targets = a.targets; d.row_order = a.row_order;
    d.a_indptr = a.a_indptr; d.a_indices = a.a_indices; d.a_data = a.a_data;
    d.b_indptr = a.b_indptr; d.b_indices = a.b_indices; d.b_data = a.b_data;
    d.b_split = a.b_split; d.split_stride = a.split_stride; d.n_panels = pl.n_panels; d.W = pl.W;
    d.n_cols = a.n_cols;
    d.Xt = a.Xtversky; d.Yt = a.Ytversky; d.Xc = a.Xcosine; d.Yc = a.Ycosine; d.Xd = a.Xdepop; d.Yd = a.Ydepop;
    d.a1 = a.a1; d.l1 = a.l1; d.l2 = a.l2; d.l3 = a.l3; d.t1 = a.t1; d.t2 = a.t2;
    d.stab = a.stabilized_shrink; d.bayes = a.bayesian_shrink; d.thr = a.threshold;
    d.k = a.k; d.cap = pl.cap;
    d.filter_mode = a.filter_mode; d.f_indptr = a.filter_indptr; d.f_indices = a.filter_indices;
    d.target_mode = a.target_mode; d.t_indptr = a.target_indptr; d.t_indices = a.target_indices;
    d.out_rows = a.out_rows; d.out_cols = a.out_cols; d.out_vals = a.out_values; d.out_counts = a.out_counts;
    d.work_counter = reinterpret_cast<int *>(scratch);
    d.cand_global = reinterpret_cast<u64 *>(reinterpret_cast<unsigned char *>(scratch) + 256);

    cudaStream_t st = as_stream(stream);
    SPY_CUDA_OK(cudaMemsetAsync(scratch, 0, 256, st));
    knn_kernel_t kern = pick_kernel(pl.threads, pl.lanes, pl.cand_smem);
    SPY_CUDA_OK(cudaFuncSetAttribute((const void *)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_bytes));
    kern<<<grid, pl.threads, pl.smem_bytes, st>>>(d);
    SPY_LAUNCH_OK();
    return SPY_OK;
}

// Host-pointer variant: the exact data the reference's Cython call site holds (s_plus.pyx:359-384).
int spy_knn_topk_host(const spy_knn_args *host_args, int device) {
    SPY_REQUIRE(host_args != nullptr, "args is NULL");
    spy_knn_args a = *host_args;
    SPY_REQUIRE(a.k >= 1 && a.n_targets >= 0 && a.a_rows >= 0 && a.b_rows >= 0, "bad sizes");
    if (a.n_targets == 0) return SPY_OK;
    SPY_CUDA_OK(cudaSetDevice(device));
    std::vector<void *> owned;
    auto cleanup = [&]() { for (void *q : owned) cudaFree(q); };
    int rc = SPY_OK;
#define UP(dst, src, count, type)                                                                       \
    do {                                                                                                \
        dst = nullptr;                                                                                  \
        if ((src) != nullptr && (count) > 0) {                                                          \
            void *_d = nullptr;                                                                         \
            cudaError_t _e = cudaMalloc(&_d, (size_t)(count) * sizeof(type));                           \
            if (_e == cudaSuccess) { owned.push_back(_d);                                               \
                _e = cudaMemcpy(_d, (src), (size_t)(count) * sizeof(type), cudaMemcpyHostToDevice); }   \
            if (_e != cudaSuccess) { set_error("upload failed: %s", cudaGetErrorString(_e)); cleanup(); \
                return _e == cudaErrorMemoryAllocation ? SPY_ERR_NOMEM : SPY_ERR_CUDA; }                \
            dst = (const type *)_d;                                                                     \
        }                                                                                               \
    } while (0)
    int32_t a_nnz = host_args->a_indptr[a.a_rows];
    int32_t b_nnz = host_args->b_indptr[a.b_rows];
    UP(a.targets, host_args->targets, a.n_targets, int32_t);
    UP(a.a_indptr, host_args->a_indptr, a.a_rows + 1, int32_t);
    UP(a.a_indices, host_args->a_indices, a_nnz, int32_t);
    UP(a.a_data, host_args->a_data, a_nnz, float);
    UP(a.b_indptr, host_args->b_indptr, a.b_rows + 1, int32_t);
    UP(a.b_indices, host_args->b_indices, b_nnz, int32_t);
    UP(a.b_data, host_args->b_data, b_nnz, float);
    if (a.l1 != 0.f) { UP(a.Xtversky, host_args->Xtversky, a.a_rows, float); UP(a.Ytversky, host_args->Ytversky, a.n_cols, float); }
    if (a.l2 != 0.f) { UP(a.Xcosine, host_args->Xcosine, a.a_rows, float); UP(a.Ycosine, host_args->Ycosine, a.n_cols, float); }
    if (a.l3 != 0.f) { UP(a.Xdepop, host_args->Xdepop, a.a_rows, float); UP(a.Ydepop, host_args->Ydepop, a.n_cols, float); }
    if (a.filter_mode == SPY_SEL_MATRIX) {
        UP(a.filter_indptr, host_args->filter_indptr, a.a_rows + 1, int32_t);
        UP(a.filter_indices, host_args->filter_indices, host_args->filter_indptr[a.a_rows], int32_t);
    }
    if (a.target_mode == SPY_SEL_MATRIX) {
        UP(a.target_indptr, host_args->target_indptr, a.a_rows + 1, int32_t);
        UP(a.target_indices, host_args->target_indices, host_args->target_indptr[a.a_rows], int32_t);
    }
    a.row_order = nullptr;
#undef UP
    const size_t slab = (size_t)a.n_targets * a.k;
    void *d_cols = nullptr, *d_vals = nullptr, *d_rows = nullptr, *d_counts = nullptr, *d_split = nullptr, *d_scratch = nullptr;
    auto dev_alloc = [&](void **q, size_t bytes) -> bool {
        if (cudaMalloc(q, bytes) != cudaSuccess) { set_error("device allocation of %zu bytes failed", bytes); return false; }
        owned.push_back(*q);
        return true;
    };
    if (!dev_alloc(&d_cols, slab * 4) || !dev_alloc(&d_vals, slab * 4) || !dev_alloc(&d_counts, (size_t)a.n_targets * 4) ||
        (host_args->out_rows && !dev_alloc(&d_rows, slab * 4))) { cleanup(); return SPY_ERR_NOMEM; }
    a.out_cols = (int32_t *)d_cols; a.out_values = (float *)d_vals; a.out_rows = (int32_t *)d_rows; a.out_counts = (int32_t *)d_counts;
    a.panel_width = 0; a.n_panels = 0; a.split_stride = 0; a.b_split = nullptr;
    a.threads = host_args->threads; a.lanes_per_segment = host_args->lanes_per_segment;
    rc = spy_knn_plan(&a, a.b_rows > 0 ? (double)b_nnz / a.b_rows : 0.0, device);
    if (rc != SPY_OK) { cleanup(); return rc; }
    if (a.n_panels > 1) {
        if (!dev_alloc(&d_split, (size_t)a.b_rows * a.split_stride * 4)) { cleanup(); return SPY_ERR_NOMEM; }
        rc = spy_knn_build_split_dev(a.b_rows, a.b_indptr, a.b_indices, a.panel_width, a.n_panels, a.split_stride,
                                     (int32_t *)d_split, nullptr);
        if (rc != SPY_OK) { cleanup(); return rc; }
        a.b_split = (const int32_t *)d_split;
    }
    int64_t sb = spy_knn_scratch_bytes(&a, device);
    if (sb < 0 || !dev_alloc(&d_scratch, (size_t)sb)) { cleanup(); return SPY_ERR_NOMEM; }
    rc = spy_knn_topk_dev(&a, d_scratch, sb, nullptr);
    if (rc != SPY_OK) { cleanup(); return rc; }
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaMemcpy(host_args->out_cols, d_cols, slab * 4, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(host_args->out_values, d_vals, slab * 4, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && host_args->out_rows) e = cudaMemcpy(host_args->out_rows, d_rows, slab * 4, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && host_args->out_counts)
        e = cudaMemcpy(host_args->out_counts, d_counts, (size_t)a.n_targets * 4, cudaMemcpyDeviceToHost);
    cleanup();
    if (e != cudaSuccess) { set_error("knn host run failed: %s", cudaGetErrorString(e)); return SPY_ERR_CUDA; }
    return SPY_OK;
}

}  // extern "C"
