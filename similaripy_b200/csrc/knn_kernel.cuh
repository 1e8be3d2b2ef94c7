// knn_kernel.cuh -- the hot path: CSR x CSR row-expansion SpGEMM with the similarity
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
#pragma once
#include "common.cuh"

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
    int group;       // lanes that stream one B-row segment together (4, 8, 16 or 32)
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

// 16-byte gathers (two pairs each) a lane has in flight before its first add: 64 registers per thread at
// 1024 threads per CTA leave room for 4, 128 registers at 512 threads for 8
#ifndef SPY_UNROLL
#define SPY_UNROLL 4
#endif
#ifndef SPY_UNROLL_512
#define SPY_UNROLL_512 8
#endif
__host__ __device__ constexpr int unroll_for(int threads) { return threads == 512 ? SPY_UNROLL_512 : SPY_UNROLL; }

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

// One drain pass over the current panel, 4 slots per thread per step (LDS.128 / STS.128).  No barriers inside.
// Every slot visited is reset to "untouched" unless it survived the pre-filter and the candidate buffer was
// full; in that case the thread raises *s_overflow and stops, and the index it returns is where it resumes
// after the caller has evaluated + selected (which raises tau), so no slot is scanned twice.
// An untouched slot holds -0.0f: with lo > 0 it fails "x >= lo * den" like any small dot product, so the
// common path needs no separate "touched" test -- only survivors are checked against the sentinel.
template <int NT, int KIND>
__device__ __forceinline__ int drain_pass(const KnnDev &p, const FastRow &fr, unsigned acc32, int base, int width,
                                          float lo, u64 *cand, int *s_cnt, int *s_overflow, int resume) {
    const float sent = __uint_as_float(kSentinelBits);
    int idx = resume;
    for (; idx < width; idx += NT * 4) {  // W % 128 == 0: the quad stays inside the panel
        const unsigned a = acc32 + (unsigned)idx * 4u;
        const float4 a4 = lds128(a);
        const float xs[4] = {a4.x, a4.y, a4.z, a4.w};
        const int col0 = base + idx;
        bool sv[4] = {true, true, true, true};
        if (!p.exact_only) {
            if (KIND == KIND_RAW) {
#pragma unroll
                for (int r = 0; r < 4; r++) sv[r] = !(xs[r] < lo);
            } else if (KIND == KIND_C) {
                const float4 y = load_y4(p.Yc, col0, p.n_cols);
                const float ys[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
                for (int r = 0; r < 4; r++) {
                    const float den = fmaf(fr.cC, ys[r], fr.A0);
                    sv[r] = !(den > 0.f && xs[r] < lo * den);
                }
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
                if (KIND == KIND_GEN && p.l2 != 0.f) {
                    const float4 y = load_y4(p.Yc, col0, p.n_cols);
                    const float ys[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
                    for (int r = 0; r < 4; r++) {
                        const float u = fr.cC * ys[r];
                        den[r] += u;
                        sab[r] += fabsf(u);
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
                    const bool rej = den[r] > 0.f && xs[r] < lo * den[r] && (den[r] * 64.f >= sab[r] + fabsf(fr.A0));
                    sv[r] = !rej;
                }
            }
        }
        float ws[4] = {sent, sent, sent, sent};
        bool full = false;
        if (sv[0] | sv[1] | sv[2] | sv[3]) {
#pragma unroll
            for (int r = 0; r < 4; r++)
                if (sv[r] && __float_as_uint(xs[r]) != kSentinelBits && !push_raw(cand, s_cnt, p.cap, xs[r], col0 + r)) {
                    ws[r] = xs[r];
                    full = true;
                }
        }
        sts128(a, make_float4(ws[0], ws[1], ws[2], ws[3]));
        if (full) { *s_overflow = 1; break; }
    }
    return idx;
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

// Expand + accumulate one staged chunk of the target row into the current panel (s_plus.h:358-403 / 418-438).
// A group of G lanes owns one A entry (u, v) at a time: its lanes stream the entries of B[u,:] that fall into
// the panel -- the contiguous run [s, e) of 8-byte (column, value) pairs given by the precomputed split
// points -- G pairs per step, U steps' loads in flight before the first shared-memory add is issued.  No
// block-wide scan, no barriers, no per-product index arithmetic: a warp only synchronises with itself, and
// the split points of a group's NEXT entry are fetched while the current one is being accumulated.
struct ExpandArgs {  // what the expansion needs of KnnDev, by value: the routine is compiled out of line
    const int *b_indptr, *b_split;
    const uint2 *b_pairs;
    int split_stride, n_panels, pn;
};
#ifndef SPY_BATCH_CAS
#define SPY_BATCH_CAS 0
#endif
__device__ __forceinline__ float lds32(unsigned addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ unsigned cas_shared(unsigned addr, unsigned expect, unsigned desired) {
    unsigned got;
    asm volatile("atom.shared.cas.b32 %0, [%1], %2, %3;" : "=r"(got) : "r"(addr), "r"(expect), "r"(desired));
    return got;
}
__device__ __forceinline__ uint4 ldg128(const uint2 *ptr) {
    return __ldg(reinterpret_cast<const uint4 *>(ptr));
}

// N float adds to N (pairwise distinct per lane) shared-memory addresses with the compare-and-swap loops
// interleaved: one round issues every pending load, then every pending CAS, so a lane has N independent
// shared-memory operations in flight instead of one (the plain red.shared.add.f32 is an LDS / FADD /
// ATOMS.CAST.SPIN loop per add, each waiting for the previous one).  `pend` has bit r set for a live add.
template <int N>
__device__ __forceinline__ void smem_add_batch(const unsigned (&addr)[N], const float (&x)[N], unsigned pend) {
#if SPY_BATCH_CAS
    while (pend) {
        float old[N];
#pragma unroll
        for (int r = 0; r < N; r++)
            if (pend & (1u << r)) old[r] = lds32(addr[r]);
#pragma unroll
        for (int r = 0; r < N; r++)
            if (pend & (1u << r)) {
                const unsigned o = __float_as_uint(old[r]);
                if (cas_shared(addr[r], o, __float_as_uint(__fadd_rn(old[r], x[r]))) == o) pend &= ~(1u << r);
            }
    }
#else
#pragma unroll
    for (int r = 0; r < N; r++)
        if (pend & (1u << r)) smem_add_f32(addr[r], x[r]);
#endif
}

template <int NT, int G>
__device__ __forceinline__ bool accumulate_chunk(const ExpandArgs x, int n, const int *st_u, const float *st_v,
                                              unsigned accb32) {
    constexpr int GROUPS = NT / G;
    constexpr int U = unroll_for(NT);  // 16-byte loads (2 pairs each) in flight per lane
    const int tid = threadIdx.x;
    const int gl = tid & (G - 1);
    const int grp0 = (tid & ~31) / G;  // first group of this warp: the loop below is warp-uniform
    int idx = tid / G;
    bool any = false;
    int s = 0, e = 0, s2, e2;
    float v = 0.f, v2;
    auto fetch = [&](int i, int &s_, int &e_, float &v_) {
        s_ = 0; e_ = 0; v_ = 0.f;
        if (i < n) {
            const int u = st_u[i];
            v_ = st_v[i];
            if (x.n_panels == 1) { s_ = __ldg(x.b_indptr + u); e_ = __ldg(x.b_indptr + u + 1); }
            else {
                const int *sp = x.b_split + (size_t)u * x.split_stride + x.pn;
                s_ = __ldg(sp); e_ = __ldg(sp + 1);
            }
        }
    };
    fetch(idx, s, e, v);
    for (int i0 = grp0; i0 < n; i0 += GROUPS) {
        fetch(idx + GROUPS, s2, e2, v2);
        const int sa = s & ~1;  // pairs are 8 bytes: an even position is 16-byte aligned
        const int maxspan = __reduce_max_sync(0xffffffffu, e - sa);
        any |= e > s;
        for (int b = 0; b < maxspan; b += 2 * G * U) {
            const int q0 = sa + b + 2 * gl;
            uint4 pr[U];
#pragma unroll
            for (int r = 0; r < U; r++)
                if (q0 + 2 * G * r < e) pr[r] = ldg128(x.b_pairs + q0 + 2 * G * r);  // the array is padded by one pair
            unsigned addr[2 * U];
            float val[2 * U];
            unsigned pend = 0u;
#pragma unroll
            for (int r = 0; r < U; r++) {
                const int q = q0 + 2 * G * r;
                addr[2 * r] = accb32 + pr[r].x * 4u;
                val[2 * r] = __fmul_rn(__uint_as_float(pr[r].y), v);
                addr[2 * r + 1] = accb32 + pr[r].z * 4u;
                val[2 * r + 1] = __fmul_rn(__uint_as_float(pr[r].w), v);
                if (q >= s && q < e) pend |= 1u << (2 * r);
                if (q + 1 < e) pend |= 1u << (2 * r + 1);  // q + 1 >= s always (q >= sa >= s - 1)
            }
            smem_add_batch<2 * U>(addr, val, pend);
        }
        s = s2; e = e2; v = v2;
        idx += GROUPS;
    }
    return any;
}

// The kernel.  One persistent CTA per resident slot; each CTA owns one target row at a time.
//   stage   : the target row's (column, value) entries go to shared memory once (rows longer than NT entries:
//             chunk by chunk) and are reused by every column panel;
//   expand  : a group of G lanes per A entry streams the B-row segment that falls into the panel
//             (accumulate_chunk above);
//   accumulate : shared-memory float adds (LDS / FADD / ATOMS.CAST.SPIN) into the panel, -0.0f = untouched;
//   drain   : 4 slots per thread per step (LDS.128), division-free pre-filter; survivors are buffered RAW
//             (dot product, column) and evaluated densely afterwards -- computeSimilarity with its IEEE
//             division runs with all lanes busy -- then the sampled selection keeps the best k.
template <int NT, int KIND, bool CAND_SMEM, int G>
__global__ void __launch_bounds__(NT, 1)
knn_flat_kernel(const __grid_constant__ KnnDev p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *acc = reinterpret_cast<float *>(smem_raw);
    unsigned char *ptr = smem_raw + (size_t)p.W * sizeof(float);
    u64 *cand;
    if (CAND_SMEM) { cand = reinterpret_cast<u64 *>(ptr); ptr += (size_t)p.cap * sizeof(u64); }
    else cand = p.cand_global + (size_t)blockIdx.x * p.cap;
    // staged chunk of the target row (8 bytes per thread); the same bytes serve as the selection's scratch
    // (`tmp`), so a selection invalidates the staged row (staged_ok)
    int *st_u = reinterpret_cast<int *>(ptr);               // A's column ids = rows of B
    float *st_v = reinterpret_cast<float *>(st_u + NT);     // A's values
    u64 *tmp = reinterpret_cast<u64 *>(ptr);
    constexpr int kTmpCap = NT;

    __shared__ int s_row, s_cnt, s_overflow, s_live;
    __shared__ u64 s_tau, s_pivot;

    const int tid = threadIdx.x;
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
        const bool single_chunk = (a1 - a0) <= NT;
        bool staged_ok = false;
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
            bool any = false;
            // ---------------- expand + accumulate (s_plus.h:358-403 / 418-438) ----------------
            for (int c0 = a0; c0 < a1; c0 += NT) {
                const int n = min(NT, a1 - c0);
                if (!staged_ok) {
                    __syncthreads();  // earlier readers of the staging area (previous chunk / selection scratch)
                    if (tid < n) { st_u[tid] = __ldg(p.a_indices + c0 + tid); st_v[tid] = __ldg(p.a_data + c0 + tid); }
                    __syncthreads();
                    staged_ok = single_chunk;
                }
                const ExpandArgs x = {p.b_indptr, p.b_split, p.b_pairs, p.split_stride, p.n_panels, pn};
                any |= accumulate_chunk<NT, G>(x, n, st_u, st_v, accb32);
            }
            // barrier: all adds have landed; nothing landed in this panel => accumulator still clean
            if (!__syncthreads_or(any ? 1 : 0)) continue;

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
            int resume = tid * 4;
            for (;;) {
                if (p.target_mode == SPY_SEL_MATRIX) drain_pass_list<NT>(p, acc, base, tlo, thi, cand, &s_cnt, &s_overflow);
                else resume = drain_pass<NT, KIND>(p, fr, acc32, base, width, lo, cand, &s_cnt, &s_overflow, resume);
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
                    staged_ok = false;  // tmp overlays the staged row
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

typedef void (*knn_kernel_t)(const KnnDev);
// one translation unit per group width (knn_inst_g*.cu): the kernel for (threads, similarity kind, buffer placement)
knn_kernel_t pick_kernel_g4(int threads, int kind, bool cand_smem);
knn_kernel_t pick_kernel_g8(int threads, int kind, bool cand_smem);
knn_kernel_t pick_kernel_g16(int threads, int kind, bool cand_smem);
knn_kernel_t pick_kernel_g32(int threads, int kind, bool cand_smem);

}  // namespace spy
