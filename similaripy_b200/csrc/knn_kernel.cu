// knn_kernel.cu -- the hot path: CSR x CSR row-expansion SpGEMM with the similarity
// denominator fused in and per-row top-k selection, for sm_100a (B200).
//
// Replaces s_plus::compute_similarities_parallel<int,float> (reference
// similaripy/cython_code/s_plus.h:265-453) behind the C ABI in include/similaripy_b200.h.
//
// Design:
//   * one persistent CTA per SM slot; CTAs pull target rows from an atomic queue
//     (the reference's `omp for schedule(dynamic)`, s_plus.h:337);
//   * the output columns are cut into panels of `panel_width` columns; a panel of fp32
//     partial sums lives in shared memory (the reference's `sums` buffer, s_plus.h:91,
//     at shared-memory instead of L2-cache scale, s_plus.h:305-311);
//   * panel boundaries inside every sorted row of B are precomputed once (b_split) --
//     the reference does a std::lower_bound per (target row, block, B row), s_plus.h:381-394;
//   * the products of (target row, panel) are flattened into one index space by a block scan of
//     the segment lengths, so every lane of every warp streams 8-byte (column, value) pairs
//     whatever the segment lengths are; accumulation is a shared-memory float atomic
//     (measured 538-605 Gproducts/s on B200, profiles/microbench/accum_bench_r01.txt);
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

struct KnnDev {
    int n_targets;
    const int *targets;
    const int *row_order;
    const int *a_indptr, *a_indices;
    const float *a_data;
    const int *b_indptr, *b_indices;
    const float *b_data;
    const uint2 *b_pairs;
    const int *b_split;
    int split_stride, n_panels, W, n_cols;
    const float *Xt, *Yt, *Xc, *Yc, *Xd, *Yd;
    float a1, l1, l2, l3, t1, t2, stab, bayes, thr;
    int has_den;     // any of l1, l2, l3, stab, bayes != 0 (s_plus.h:144)
    int exact_only;  // skip the fast pre-filter (a1 != 1: powf involved)
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

// ---- selection -------------------------------------------------------------------------------------
// The candidate buffer holds 64-bit keys; 0 marks a dead entry (a candidate that lost against the threshold
// or the running k-th best when its exact value was computed).  select_topk leaves the best
// m = min(k, #live) keys sorted best-first in cand[0, m), sets *s_cnt = m and, when m == k, *s_tau = the k-th.
//
// Sorting the whole buffer (2048 keys -> 66 compare-exchange steps over 1024 threads) every time it fills
// used to cost as much as the accumulation itself.  Instead: one warp sorts 64 strided samples, a pivot is
// picked a safe distance below the sample quantile of the k-th best, the keys above the pivot (a few
// hundred) are compacted into `tmp` and only those are sorted.  If the pivot turns out too high (fewer than
// k keys above it) or too low (tmp overflows) the full sort runs -- results never depend on the sampling.
template <int NT>
__device__ void sort_and_publish(u64 *buf, int n, int k, u64 *cand, int *s_cnt, u64 *s_tau, int *s_live) {
    // sort buf[0, n) descending (padded with zeros), copy the best min(k, live) to cand if buf != cand
    const int tid = threadIdx.x;
    int S = 2;
    while (S < n) S <<= 1;
    for (int i = n + tid; i < S; i += NT) buf[i] = 0ull;
    if (tid == 0) *s_live = 0;
    __syncthreads();
    bitonic_sort_desc<NT>(buf, S);
    for (int i = tid; i < S; i += NT)  // live keys are a prefix: find its end
        if (buf[i] != 0ull && (i == S - 1 || buf[i + 1] == 0ull)) *s_live = i + 1;
    __syncthreads();
    const int m = min(*s_live, k);
    if (buf != cand)
        for (int i = tid; i < m; i += NT) cand[i] = buf[i];
    __syncthreads();
    if (tid == 0) {
        *s_cnt = m;
        if (m == k) *s_tau = buf[k - 1];
    }
    __syncthreads();
}

template <int NT>
__device__ void select_topk(u64 *cand, int n, int k, u64 *tmp, int tmp_cap, int *s_cnt, u64 *s_tau, int *s_live,
                            u64 *s_pivot) {
    const int tid = threadIdx.x;
    // pivot rank among 64 sorted samples: mean + 3 sigma above the sample quantile of the k-th best
    const float q = 65.f * (float)k / (float)max(n, 1);
    const int j = (int)ceilf(q + 3.f * sqrtf(q) + 1.5f);
    if (n <= 512 || j > 40 || 2 * k > tmp_cap) {  // small buffer or k too close to n: sort it all
        sort_and_publish<NT>(cand, n, k, cand, s_cnt, s_tau, s_live);
        return;
    }
    if (tid < 64) tmp[tid] = cand[(int)(((long long)tid * n) >> 6)];
    if (tid == 0) *s_live = 0;
    __syncthreads();
    if (tid < 32) {  // 64-key bitonic sort by one warp, descending
        for (int size = 2; size <= 64; size <<= 1)
            for (int stride = size >> 1; stride > 0; stride >>= 1) {
                const int lo = ((tid & ~(stride - 1)) << 1) | (tid & (stride - 1));
                const int hi = lo + stride;
                const bool desc = ((lo & size) == 0) || (size == 64);
                const u64 a = tmp[lo], b = tmp[hi];
                if ((a < b) == desc) { tmp[lo] = b; tmp[hi] = a; }
                __syncwarp();
            }
        if (tid == 0) *s_pivot = tmp[j - 1];
    }
    __syncthreads();
    const u64 pivot = *s_pivot;
    for (int i = tid; i < n; i += NT) {
        const u64 key = cand[i];
        if (key > pivot) {
            const int pos = atomicAdd(s_live, 1);
            if (pos < tmp_cap) tmp[pos] = key;
        }
    }
    __syncthreads();
    const int c = *s_live;
    __syncthreads();
    if (c < k || c > tmp_cap) {  // unlucky pivot: exact fallback
        sort_and_publish<NT>(cand, n, k, cand, s_cnt, s_tau, s_live);
        return;
    }
    sort_and_publish<NT>(tmp, c, k, cand, s_cnt, s_tau, s_live);
}

// ---- fast pre-filter of the drain ---------------------------------------------------------------
// The exact similarity (similarity_value) costs an IEEE division per candidate, and after the first
// selection more than 95 % of the candidates lose against the running k-th best.  For a1 == 1 and no
// bayesian shrink the value is xy / den with
//     den = A0 + cT*Yt[c] + cX*xy + cC*Yc[c] + cD*Yd[c]
//     A0 = stab + l1*t1*Xt[r], cT = l1*t2, cX = l1*(1 - t1 - t2), cC = l2*Xc[r], cD = l3*Xd[r]
// (computeSimilarity, s_plus.h:129-156, regrouped), so "value < lo" is "xy < lo * den" for den > 0: two or
// three FMAs and a compare, no division.  The test only ever REJECTS, with lo sitting 1e-4 (relative)
// below what can still enter the result, and it abstains when den is not safely positive or suffers
// cancellation (den*64 < sum of |addends|); everything it lets through takes the exact path, so results
// do not depend on it.
constexpr int KIND_RAW = 0;  // value = xy (dot_product, p3alpha without shrink)
constexpr int KIND_T = 1;    // Tversky / Jaccard / Dice: Yt only
constexpr int KIND_C = 2;    // cosine family: Yc only
constexpr int KIND_D = 4;    // depop only (rp3beta): Yd only
constexpr int KIND_GEN = 7;  // anything else: terms selected at run time

struct FastRow {
    float A0, cT, cX, cC, cD;
};

__device__ __forceinline__ float4 load_y4(const float *v, int col0, int n_cols) {
    if (col0 + 3 < n_cols) return __ldg(reinterpret_cast<const float4 *>(v + col0));
    float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
    if (col0 < n_cols) r.x = __ldg(v + col0);
    if (col0 + 1 < n_cols) r.y = __ldg(v + col0 + 1);
    if (col0 + 2 < n_cols) r.z = __ldg(v + col0 + 2);
    return r;
}

// lower end of what can still enter the result: max(threshold, running k-th best) minus a 1e-4 relative
// safety margin that covers the rounding differences between the fast test and the exact value.
__device__ __forceinline__ float reject_bound(const KnnDev &p, u64 tau) {
    float bound = p.thr;
    if (tau != 0ull) bound = fmaxf(bound, unordered_bits((unsigned)(tau >> 32)));
    return bound - fabsf(bound) * 1e-4f - 1e-37f;
}

// Exclusive block scan of one int per thread (NT <= 1024).  Two barriers.  wtot has 33 ints.
template <int NT>
__device__ __forceinline__ int block_exclusive_scan(int v, int *wtot, int &total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) wtot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const int w = (lane < NT / 32) ? wtot[lane] : 0;
        int wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += t;
        }
        wtot[lane] = wi - w;
        if (lane == 31) wtot[32] = wi;
    }
    __syncthreads();
    total = wtot[32];
    return wtot[warp] + incl - v;
}

// shared-memory float add on a 32-bit shared address: LDS / FADD / ATOMS.CAST.SPIN loop, 543 Gadd/s on
// B200 (profiles/microbench); the plain atomicAdd(float*) form recomputes the shared window base per call.
__device__ __forceinline__ void smem_add_f32(unsigned addr, float x) {
    asm volatile("red.shared.add.f32 [%0], %1;" ::"r"(addr), "f"(x) : "memory");
}

#ifndef SPY_UNROLL
#define SPY_UNROLL 8
#endif
constexpr int kUnroll = SPY_UNROLL;  // independent 8-byte gathers in flight per lane
constexpr int kGroup = 8;   // lanes that walk one contiguous run of products together

// shared-memory vector accesses on 32-bit shared addresses (the generic-pointer forms make ptxas rebuild
// the shared window base inside the loops)
__device__ __forceinline__ float4 lds128(unsigned addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts128(unsigned addr, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// raw (not yet evaluated) candidate: accumulated dot product + column
__device__ __forceinline__ u64 make_raw(float xy, int col) { return ((u64)__float_as_uint(xy) << 32) | (u64)(unsigned)col; }

// Buffer one raw candidate; false when the buffer is full (the caller then leaves the slot as it is).
__device__ __forceinline__ bool push_raw(u64 *cand, int *s_cnt, int cap, float xy, int col) {
    if (*reinterpret_cast<volatile int *>(s_cnt) >= cap) return false;
    const int pos = atomicAdd(s_cnt, 1);
    if (pos >= cap) return false;
    cand[pos] = make_raw(xy, col);
    return true;
}

// One drain pass over the current panel.  No barriers inside.  A slot is reset to "untouched" once it has
// been rejected by the pre-filter or buffered as a raw candidate; when the buffer is full the slot keeps its
// value and *s_overflow is raised, so that the caller evaluates + selects (which raises tau) and runs
// another pass over what is left.
template <int NT, int KIND>
__device__ __forceinline__ void drain_pass(const KnnDev &p, const FastRow &fr, unsigned acc32, int base, int width,
                                           float lo, u64 *cand, int *s_cnt, int *s_overflow) {
    const float sent = __uint_as_float(kSentinelBits);
    bool overflow = false;
    for (int idx = threadIdx.x * 4; idx < width; idx += NT * 4) {  // W % 128 == 0: the quad stays inside the panel
        const unsigned a = acc32 + (unsigned)idx * 4u;
        const float4 a4 = lds128(a);
        const float xs[4] = {a4.x, a4.y, a4.z, a4.w};
        bool sv[4];
#pragma unroll
        for (int r = 0; r < 4; r++) sv[r] = __float_as_uint(xs[r]) != kSentinelBits;
        if (!(sv[0] | sv[1] | sv[2] | sv[3])) continue;
        const int col0 = base + idx;
        if (!p.exact_only) {
            if (KIND == KIND_RAW) {
#pragma unroll
                for (int r = 0; r < 4; r++) sv[r] = sv[r] && !(xs[r] < lo);
            } else {
                float den[4] = {fr.A0, fr.A0, fr.A0, fr.A0};
                float sab[4] = {0.f, 0.f, 0.f, 0.f};
                if (KIND == KIND_T || (KIND == KIND_GEN && p.l1 != 0.f)) {
                    const float4 y = load_y4(p.Yt, col0, p.n_cols);
                    const float ys[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
                    for (int r = 0; r < 4; r++) {
                        const float u = fr.cT * ys[r], w = fr.cX * xs[r];
                        den[r] += u + w;
                        sab[r] += fabsf(u) + fabsf(w);
                    }
                }
                if (KIND == KIND_C || (KIND == KIND_GEN && p.l2 != 0.f)) {
                    const float4 y = load_y4(p.Yc, col0, p.n_cols);
                    const float ys[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
                    for (int r = 0; r < 4; r++) {
                        const float u = fr.cC * ys[r];
                        den[r] += u;
                        if (KIND != KIND_C) sab[r] += fabsf(u);
                    }
                }
                if (KIND == KIND_D || (KIND == KIND_GEN && p.l3 != 0.f)) {
                    const float4 y = load_y4(p.Yd, col0, p.n_cols);
                    const float ys[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
                    for (int r = 0; r < 4; r++) {
                        const float u = fr.cD * ys[r];
                        den[r] += u;
                        sab[r] += fabsf(u);
                    }
                }
#pragma unroll
                for (int r = 0; r < 4; r++) {
                    // den huge / inf: lo * den is +-inf or NaN and the comparison does the right thing
                    bool rej = den[r] > 0.f && xs[r] < lo * den[r];
                    if (KIND != KIND_C) rej = rej && (den[r] * 64.f >= sab[r] + fabsf(fr.A0));
                    sv[r] = sv[r] && !rej;
                }
            }
        }
        float ws[4] = {sent, sent, sent, sent};
        if (sv[0] | sv[1] | sv[2] | sv[3]) {
#pragma unroll
            for (int r = 0; r < 4; r++)
                if (sv[r] && !push_raw(cand, s_cnt, p.cap, xs[r], col0 + r)) { ws[r] = xs[r]; overflow = true; }
        }
        sts128(a, make_float4(ws[0], ws[1], ws[2], ws[3]));
    }
    if (overflow) *s_overflow = 1;
}

// Same for MODE_MATRIX target columns (s_plus.h:175-188): only the columns listed in the target row's sorted
// list [tlo, thi) can be candidates.  A consumed slot is reset, which also makes duplicate list entries harmless.
template <int NT>
__device__ __forceinline__ void drain_pass_list(const KnnDev &p, float *acc, int base, int tlo, int thi, u64 *cand,
                                                int *s_cnt, int *s_overflow) {
    bool overflow = false;
    for (int q = tlo + threadIdx.x; q < thi; q += NT) {
        const int col = __ldg(p.t_indices + q);
        if (q > tlo && __ldg(p.t_indices + q - 1) == col) continue;  // one owner per column
        const float xy = acc[col - base];
        if (__float_as_uint(xy) == kSentinelBits) continue;
        if (push_raw(cand, s_cnt, p.cap, xy, col)) acc[col - base] = __uint_as_float(kSentinelBits);
        else overflow = true;
    }
    if (overflow) *s_overflow = 1;
}

// The kernel.  One persistent CTA per resident slot; each CTA owns one target row at a time.
//   stage   : one A-row entry per thread: the B-row segment [s, e) that falls into the current column
//             panel (from the precomputed split points), its length, A's value; block scan of the
//             lengths -> the row's products of this panel become one flat index space [0, T);
//   expand  : every warp takes a contiguous slice of [0, T) and every group of 8 lanes a contiguous quarter
//             of it; lane l of a group handles products g0 + l + 8 i and walks the segment boundaries as the
//             index grows, so all lanes are busy whatever the segment lengths are, a group reads 64 contiguous
//             bytes of 8-byte (column, value) pairs per step, and a boundary is crossed once per segment;
//   accumulate : shared-memory float adds (LDS / FADD / ATOMS.CAST.SPIN) into the panel, -0.0f = untouched;
//   drain   : 4 slots per thread per step (LDS.128), division-free pre-filter; survivors are buffered RAW
//             (dot product, column) and evaluated densely afterwards -- computeSimilarity with its IEEE
//             division runs with all lanes busy -- then the sampled selection keeps the best k.
template <int NT, int KIND, bool CAND_SMEM>
__global__ void __launch_bounds__(NT, 1024 / NT)
knn_flat_kernel(const __grid_constant__ KnnDev p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *acc = reinterpret_cast<float *>(smem_raw);
    unsigned char *ptr = smem_raw + (size_t)p.W * sizeof(float);
    u64 *cand;
    if (CAND_SMEM) { cand = reinterpret_cast<u64 *>(ptr); ptr += (size_t)p.cap * sizeof(u64); }
    else cand = p.cand_global + (size_t)blockIdx.x * p.cap;
    // staged segments of the current chunk (12 bytes per thread); the same bytes serve as the selection's
    // scratch (`tmp`) while no accumulation is running
    int *st_nb = reinterpret_cast<int *>(ptr);             // first flat index AFTER the segment
    int *st_delta = st_nb + NT;                            // position in b_pairs minus flat index
    float *st_v = reinterpret_cast<float *>(st_delta + NT);  // A's value
    u64 *tmp = reinterpret_cast<u64 *>(ptr);
    constexpr int kTmpCap = (NT * 12 / 8) >= 1024 ? 1024 : 512;

    __shared__ int s_row, s_cnt, s_overflow, s_live;
    __shared__ u64 s_tau, s_pivot;
    __shared__ int s_wtot[33];

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    constexpr int NW = NT / 32;
    const float sentinel = __uint_as_float(kSentinelBits);
    const float4 sentinel4 = make_float4(sentinel, sentinel, sentinel, sentinel);
    const unsigned acc32 = (unsigned)__cvta_generic_to_shared(acc);

    for (int i = tid * 4; i < p.W; i += NT * 4) *reinterpret_cast<float4 *>(acc + i) = sentinel4;

    for (;;) {
        __syncthreads();
        if (tid == 0) { s_row = atomicAdd(p.work_counter, 1); s_cnt = 0; s_tau = 0ull; s_overflow = 0; }
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
        FastRow fr;
        fr.A0 = p.stab + p.l1 * p.t1 * sr.Xt;
        fr.cT = p.l1 * p.t2;
        fr.cX = p.l1 * (1.f - p.t1 - p.t2);
        fr.cC = p.l2 * sr.Xc;
        fr.cD = p.l3 * sr.Xd;
        u64 tau = 0ull;
        float lo = reject_bound(p, tau);
        int n_eval = 0;  // cand[0, n_eval) are evaluated keys, cand[n_eval, s_cnt) raw candidates (uniform)

        for (int pn = 0; pn < p.n_panels; pn++) {
            const int base = pn * p.W;
            const int width = min(p.W, p.n_cols - base);
            const unsigned accb32 = acc32 - (unsigned)base * 4u;  // &acc[col - base] == accb32 + 4 * col
            bool panel_any = false;
            // ---------------- expand + accumulate (s_plus.h:358-403 / 418-438) ----------------
            for (int c0 = a0; c0 < a1; c0 += NT) {
                const int n = min(NT, a1 - c0);
                int s = 0, len = 0;
                float v1 = 0.f;
                if (tid < n) {
                    const int u = __ldg(p.a_indices + c0 + tid);
                    int e;
                    if (p.n_panels == 1) { s = __ldg(p.b_indptr + u); e = __ldg(p.b_indptr + u + 1); }
                    else {
                        const int *sp = p.b_split + (size_t)u * p.split_stride + pn;
                        s = __ldg(sp); e = __ldg(sp + 1);
                    }
                    len = e - s;
                    v1 = __ldg(p.a_data + c0 + tid);
                }
                int T;
                // the first barrier inside the scan also fences earlier readers of the staging area
                const int off = block_exclusive_scan<NT>(len, s_wtot, T);
                st_nb[tid] = off + len;
                st_delta[tid] = s - off;
                st_v[tid] = v1;
                __syncthreads();
                if (T == 0) continue;  // uniform
                panel_any = true;
                const int per_warp = (((T + NW - 1) / NW) + 31) & ~31;
                const int quarter = per_warp / (32 / kGroup);
                const int gbeg = warp * per_warp + (lane / kGroup) * quarter;
                const int gend = min(T, gbeg + quarter);
                const int j0 = gbeg + (lane & (kGroup - 1));
                // segment holding this lane's first product: smallest sg with st_nb[sg] > j
                int sg = 0;
                {
                    const int jj = min(j0, T - 1);
                    int hi = n - 1;
                    while (sg < hi) {
                        const int mid = (sg + hi) >> 1;
                        if (st_nb[mid] > jj) hi = mid; else sg = mid + 1;
                    }
                }
                int nb = st_nb[sg], delta = st_delta[sg];
                float v = st_v[sg];
                for (int i0 = 0; i0 < quarter; i0 += kGroup * kUnroll) {  // warp-uniform trip count
                    if (warp * per_warp + i0 >= T) break;                 // uniform: group 0 (the earliest) has run out
                    int addr[kUnroll];
                    float vv[kUnroll];
#pragma unroll
                    for (int r = 0; r < kUnroll; r++) {
                        const int j = j0 + i0 + kGroup * r;
                        addr[r] = -1;
                        if (j < gend) {
                            if (j >= nb) {
                                do { sg++; nb = st_nb[sg]; } while (j >= nb);
                                delta = st_delta[sg];
                                v = st_v[sg];
                            }
                            addr[r] = j + delta;
                        }
                        vv[r] = v;
                    }
                    uint2 pr[kUnroll];
#pragma unroll
                    for (int r = 0; r < kUnroll; r++)
                        if (addr[r] >= 0) pr[r] = __ldg(p.b_pairs + addr[r]);
#pragma unroll
                    for (int r = 0; r < kUnroll; r++)
                        if (addr[r] >= 0) smem_add_f32(accb32 + pr[r].x * 4u, __fmul_rn(__uint_as_float(pr[r].y), vv[r]));
                }
            }
            if (!panel_any) continue;  // uniform; nothing landed in this panel: accumulator still clean
            __syncthreads();

            // ---------------- per-row filter matrix: erase filtered columns (s_plus.h:159-172) --
            if (p.filter_mode == SPY_SEL_MATRIX) {
                const int fs = __ldg(p.f_indptr + t), fe = __ldg(p.f_indptr + t + 1);
                const int flo = lower_bound_dev(p.f_indices, fs, fe, base);
                const int fhi = lower_bound_dev(p.f_indices, flo, fe, base + width);
                for (int q = flo + tid; q < fhi; q += NT) acc[__ldg(p.f_indices + q) - base] = sentinel;
                __syncthreads();
            }

            // ---------------- drain: pre-filter, similarity, threshold, top-k (s_plus.h:193-215) ----------
            int tlo = 0, thi = 0;
            if (p.target_mode == SPY_SEL_MATRIX) {
                const int ts = __ldg(p.t_indptr + t), te = __ldg(p.t_indptr + t + 1);
                tlo = lower_bound_dev(p.t_indices, ts, te, base);
                thi = lower_bound_dev(p.t_indices, tlo, te, base + width);
            }
            for (;;) {
                if (p.target_mode == SPY_SEL_MATRIX) drain_pass_list<NT>(p, acc, base, tlo, thi, cand, &s_cnt, &s_overflow);
                else drain_pass<NT, KIND>(p, fr, acc32, base, width, lo, cand, &s_cnt, &s_overflow);
                __syncthreads();
                const bool again = s_overflow != 0;
                const int cnt = min(s_cnt, p.cap);
                // exact values of the raw candidates, all lanes busy (computeSimilarity, s_plus.h:129-156, 206)
                for (int i = n_eval + tid; i < cnt; i += NT) {
                    const u64 raw = cand[i];
                    const int col = (int)(unsigned)(raw & 0xffffffffull);
                    const float val = similarity_value(p, sr, col, __uint_as_float((unsigned)(raw >> 32)));
                    u64 key = 0ull;
                    if (val >= p.thr) key = make_key(val, col);
                    cand[i] = (key > tau) ? key : 0ull;
                }
                n_eval = cnt;
                __syncthreads();
                if (again || cnt > p.cap / 2) {  // tighten tau while the buffer is reasonably full
                    if (tid == 0) s_overflow = 0;
                    select_topk<NT>(cand, cnt, p.k, tmp, kTmpCap, &s_cnt, &s_tau, &s_live, &s_pivot);
                    tau = s_tau;
                    lo = reject_bound(p, tau);
                    n_eval = s_cnt;
                }
                if (!again) break;
            }
            if (p.target_mode == SPY_SEL_MATRIX) {  // touched slots outside the target list
                for (int i = tid * 4; i < width; i += NT * 4) *reinterpret_cast<float4 *>(acc + i) = sentinel4;
            }
        }

        // ---------------- final selection and slab write (s_plus.h:443-450) ----------------
        __syncthreads();
        select_topk<NT>(cand, n_eval, p.k, tmp, kTmpCap, &s_cnt, &s_tau, &s_live, &s_pivot);
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

// (column, value) pairs of B packed into one 8-byte word per stored entry: one LDG.64 per product
// instead of two LDG.32 from two arrays, and half the sector over-fetch on short segments.
__global__ void pack_pairs_kernel(long long nnz, const int *__restrict__ indices, const float *__restrict__ data,
                                  uint2 *__restrict__ pairs) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < nnz; q += stride)
        pairs[q] = make_uint2((unsigned)indices[q], __float_as_uint(data[q]));
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
    int threads, ctas_per_sm, W, n_panels, split_stride, cap;
    bool cand_smem;
    size_t smem_bytes;
};

static int make_plan(const spy_knn_args &a, int device, Plan &pl) {
    DeviceInfo di = device_info(device);
    pl.threads = a.threads ? a.threads : 1024;
    if (pl.threads != 512 && pl.threads != 1024) {
        set_error("threads must be 512 or 1024 (got %d)", pl.threads);
        return SPY_ERR_INVALID;
    }
    pl.ctas_per_sm = 1024 / pl.threads;
    pl.cap = std::max(2048, next_pow2(2 * std::max(a.k, 1)));
    pl.cand_smem = (size_t)pl.cap * 8 <= 65536;
    // staging: 12 bytes per thread (doubles as the selection's scratch)
    const size_t fixed = (size_t)pl.threads * 12 + (pl.cand_smem ? (size_t)pl.cap * 8 : 0);
    // 1 KB per CTA is reserved by the driver; keep a little slack for static shared memory
    const size_t budget = (size_t)di.max_smem_optin / pl.ctas_per_sm - 1024 - 256;
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
    return SPY_OK;
}

typedef void (*knn_kernel_t)(const KnnDev);

template <int NT, int KIND>
static knn_kernel_t pick_cand(bool cand_smem) {
    return cand_smem ? (knn_kernel_t)knn_flat_kernel<NT, KIND, true> : (knn_kernel_t)knn_flat_kernel<NT, KIND, false>;
}
template <int NT>
static knn_kernel_t pick_kind(int kind, bool cand_smem) {
    switch (kind) {
    case KIND_RAW: return pick_cand<NT, KIND_RAW>(cand_smem);
    case KIND_T: return pick_cand<NT, KIND_T>(cand_smem);
    case KIND_C: return pick_cand<NT, KIND_C>(cand_smem);
    case KIND_D: return pick_cand<NT, KIND_D>(cand_smem);
    default: return pick_cand<NT, KIND_GEN>(cand_smem);
    }
}
static knn_kernel_t pick_kernel(int threads, int kind, bool cand_smem) {
    return threads == 512 ? pick_kind<512>(kind, cand_smem) : pick_kind<1024>(kind, cand_smem);
}

// Which specialisation of the drain's pre-filter applies (see KIND_* above).
static int similarity_kind(const spy_knn_args &a, int &exact_only) {
    const bool has_den = a.l1 != 0.f || a.l2 != 0.f || a.l3 != 0.f || a.stabilized_shrink != 0.f || a.bayesian_shrink != 0.f;
    exact_only = (a.a1 != 1.f || a.bayesian_shrink != 0.f) ? 1 : 0;
    if (!has_den) return KIND_RAW;
    if (exact_only) return KIND_GEN;
    const bool plain = a.stabilized_shrink >= 0.f && a.l1 >= 0.f && a.l2 >= 0.f && a.l3 >= 0.f;
    if (plain && a.l1 != 0.f && a.l2 == 0.f && a.l3 == 0.f) return KIND_T;
    if (plain && a.l1 == 0.f && a.l2 != 0.f && a.l3 == 0.f) return KIND_C;
    if (plain && a.l1 == 0.f && a.l2 == 0.f && a.l3 != 0.f) return KIND_D;
    return KIND_GEN;
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

int spy_knn_plan(spy_knn_args *args, int device) {
    SPY_REQUIRE(args != nullptr, "args is NULL");
    SPY_REQUIRE(args->k >= 1, "k must be >= 1 (got %d)", args->k);
    SPY_REQUIRE(args->n_cols >= 0, "n_cols must be >= 0");
    Plan pl;
    int rc = make_plan(*args, device, pl);
    if (rc != SPY_OK) return rc;
    args->threads = pl.threads;
    args->panel_width = pl.W;
    args->n_panels = pl.n_panels;
    args->split_stride = pl.split_stride;
    return SPY_OK;
}

int64_t spy_knn_scratch_bytes(const spy_knn_args *args, int device) {
    if (!args) return SPY_ERR_INVALID;
    Plan pl;
    if (make_plan(*args, device, pl) != SPY_OK) return SPY_ERR_INVALID;
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

int spy_knn_pack_pairs_dev(int64_t nnz, const int32_t *b_indices, const float *b_data, void *pairs_out, void *stream) {
    if (nnz <= 0) return SPY_OK;
    SPY_REQUIRE(b_indices && b_data && pairs_out, "pack_pairs: NULL pointer");
    long long blocks = (nnz + 256 * 4 - 1) / (256 * 4);
    if (blocks > (long long)kB200SmCount * 16) blocks = (long long)kB200SmCount * 16;
    pack_pairs_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(nnz, b_indices, b_data,
                                                                      reinterpret_cast<uint2 *>(pairs_out));
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
    SPY_REQUIRE(a.panel_width > 0 && a.n_panels > 0 && a.threads > 0, "launch plan missing: call spy_knn_plan first");
    SPY_REQUIRE(a.b_pairs != nullptr, "b_pairs is NULL: pack B with spy_knn_pack_pairs_dev first");
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
    int rc = make_plan(a, device, pl);
    if (rc != SPY_OK) return rc;
    const int grid = grid_size(pl, a.n_targets, device);
    int64_t need = 256 + (pl.cand_smem ? 0 : (int64_t)grid * pl.cap * 8);
    SPY_REQUIRE(scratch != nullptr && scratch_bytes >= need, "scratch too small: need %lld bytes", (long long)need);

    KnnDev d;
    d.n_targets = a.n_targets; d.targets = a.targets; d.row_order = a.row_order;
    d.a_indptr = a.a_indptr; d.a_indices = a.a_indices; d.a_data = a.a_data;
    d.b_indptr = a.b_indptr; d.b_indices = a.b_indices; d.b_data = a.b_data;
    d.b_pairs = reinterpret_cast<const uint2 *>(a.b_pairs);
    d.has_den = (a.l1 != 0.f || a.l2 != 0.f || a.l3 != 0.f || a.stabilized_shrink != 0.f || a.bayesian_shrink != 0.f) ? 1 : 0;
    int exact_only = 0;
    const int kind = similarity_kind(a, exact_only);
    d.exact_only = exact_only;
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
    knn_kernel_t kern = pick_kernel(pl.threads, kind, pl.cand_smem);
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
    a.panel_width = 0; a.n_panels = 0; a.split_stride = 0; a.b_split = nullptr; a.b_pairs = nullptr;
    a.threads = host_args->threads;
    rc = spy_knn_plan(&a, device);
    if (rc != SPY_OK) { cleanup(); return rc; }
    {
        void *d_pairs = nullptr;
        if (!dev_alloc(&d_pairs, (size_t)std::max(b_nnz, 1) * 8)) { cleanup(); return SPY_ERR_NOMEM; }
        rc = spy_knn_pack_pairs_dev(b_nnz, a.b_indices, a.b_data, d_pairs, nullptr);
        if (rc != SPY_OK) { cleanup(); return rc; }
        a.b_pairs = d_pairs;
    }
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
