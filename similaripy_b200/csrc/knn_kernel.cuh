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
//   * a group of G lanes owns one entry (u, v) of the target row at a time and streams the run of B-row u
//     that falls into the panel as 16-byte gathers of two (column, value) pairs; accumulation is a
//     shared-memory float add (measured 540 Gproducts/s on B200 on its own, profiles/microbench/);
//   * "touched" is encoded in the accumulator itself: slots start at -0.0f, and
//     (-0.0f) + x == x, so a slot whose bits are still 0x80000000 was never written
//     (the reference keeps a touched-list, s_plus.h:112-117);
//   * the drain applies filter / target selectors and a division-free pre-filter against the running
//     k-th best (per-block minima of Y first, per slot second), buffers the survivors, then computes
//     computeSimilarity (s_plus.h:129-156) with the same operation order and no FMA contraction and the
//     threshold test (s_plus.h:206) for the whole buffer; sampling rounds shrink the buffer when it
//     fills, a rank sort orders the best k at the end of the row;
//   * a row's first panel is drained with a speculative bound taken from a sample and validated
//     before anything it rejected is discarded;
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
    const float *y_block_min;  // min over every 128 consecutive columns of Yc (cosine family) or Yd (depop only)
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
    u64 *phase;  // SPY_PHASE_TIMING builds: 8 cycle counters summed over CTAs (in the scratch header)
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
// yt / yc / yd = Ytversky / Ycosine / Ydepop of the candidate's column (read only where the weight is non-zero)
__device__ __forceinline__ float similarity_value(const KnnDev &p, const SimRow &r, float xy, float yt, float yc, float yd) {
    float vT = 0.f, vC = 0.f, vD = 0.f, val = xy;
    if (p.l1 != 0.f) {
        float a = __fmul_rn(p.t1, __fsub_rn(r.Xt, xy));
        float b = __fmul_rn(p.t2, __fsub_rn(yt, xy));
        vT = __fmul_rn(p.l1, __fadd_rn(__fadd_rn(a, b), xy));
    }
    if (p.l2 != 0.f) vC = __fmul_rn(p.l2, __fmul_rn(r.Xc, yc));
    if (p.l3 != 0.f) vD = __fmul_rn(p.l3, __fmul_rn(r.Xd, yd));
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
__device__ __noinline__ void sort_and_publish(u64 *buf, int n, int k, u64 *cand, int *s_cnt, u64 *s_tau, int *s_live) {
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

// Rank sort: every live key of buf[0, n) counts the keys that beat it and lands at cand[rank] when rank < k
// (keys are distinct: they carry the column).  n^2 / NT comparisons per thread and four barriers, against
// 36-55 barrier-separated compare-exchange steps of the bitonic network: the cheaper one below ~600 keys.
// buf must not alias cand.  NT / P threads share one key (P = n rounded up to a power of two >= 32).
template <int NT>
__device__ __noinline__ void rank_and_publish(const u64 *buf, int n, int k, u64 *cand, int *s_cnt, u64 *s_tau, int *s_live) {
    const int tid = threadIdx.x;
    int P = 32;
    while (P < n) P <<= 1;
    if (tid == 0) *s_live = 0;
    __syncthreads();
    int live = 0;
    constexpr int NTP = NT >= 1024 ? 1024 : 512;  // largest power of two <= NT: the threads that rank
    if (P <= NTP) {
        if (tid < NTP) {
            const int tpk = NTP / P;  // power of two <= 32: the threads of a key are neighbouring lanes
            const int i = tid / tpk, part = tid & (tpk - 1);
            const u64 key = (i < n) ? buf[i] : 0ull;
            int cnt = 0;
            for (int j = part; j < n; j += tpk) cnt += (buf[j] > key) ? 1 : 0;
            for (int o = tpk >> 1; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
            if (part == 0 && key != 0ull) {
                live = 1;
                if (cnt < k) cand[cnt] = key;
            }
        }
    } else {
        for (int i = tid; i < n; i += NT) {
            const u64 key = buf[i];
            if (key == 0ull) continue;
            int cnt = 0;
            for (int j = 0; j < n; j++) cnt += (buf[j] > key) ? 1 : 0;
            live++;
            if (cnt < k) cand[cnt] = key;
        }
    }
    for (int o = 16; o > 0; o >>= 1) live += __shfl_xor_sync(0xffffffffu, live, o);
    if ((tid & 31) == 0 && live) atomicAdd(s_live, live);
    __syncthreads();
    const int m = min(*s_live, k);
    if (tid == 0) {
        *s_cnt = m;
        if (m == k) *s_tau = cand[k - 1];
    }
    __syncthreads();
}

// One sampling round of the selection: one warp sorts 64 strided samples of cand[0, n), a pivot is picked a safe
// distance (3 sigma) below the sample quantile of the k-th best, and the keys above the pivot are compacted
// into tmp.  Returns their number c (block-uniform); the caller checks k <= c <= tmp_cap.
template <int NT>
__device__ __noinline__ int pivot_compact(const u64 *cand, int n, int j, u64 *tmp, int tmp_cap, int *s_live, u64 *s_pivot) {
    const int tid = threadIdx.x;
    if (tid < 32) {
        // 64 strided samples, two per lane (element e = lane + 32 r), sorted descending by a bitonic network in
        // registers: strides below 32 exchange with lane ^ stride, stride 32 is the lane's own pair
        u64 k0 = cand[(int)(((long long)tid * n) >> 6)];
        u64 k1 = cand[(int)(((long long)(tid + 32) * n) >> 6)];
#pragma unroll
        for (int size = 2; size <= 64; size <<= 1) {
#pragma unroll
            for (int stride = size >> 1; stride > 0; stride >>= 1) {
                if (stride == 32) {
                    const u64 hi = k0 > k1 ? k0 : k1, lo = k0 > k1 ? k1 : k0;
                    k0 = hi; k1 = lo;
                } else {
                    const u64 o0 = __shfl_xor_sync(0xffffffffu, k0, stride), o1 = __shfl_xor_sync(0xffffffffu, k1, stride);
                    const bool lower = (tid & stride) == 0;
                    const bool desc0 = size == 64 || (size == 32 ? true : (tid & size) == 0);   // e = lane
                    const bool desc1 = size == 64 || (size == 32 ? false : (tid & size) == 0);  // e = lane + 32
                    k0 = ((k0 > o0) == (lower == desc0)) ? k0 : o0;
                    k1 = ((k1 > o1) == (lower == desc1)) ? k1 : o1;
                }
            }
        }
        const u64 pv = __shfl_sync(0xffffffffu, (j - 1) < 32 ? k0 : k1, (j - 1) & 31);
        if (tid == 0) { *s_pivot = pv; *s_live = 0; }
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
    return c;
}
// pivot rank among 64 sorted samples for "the k-th best of n": its expected rank q plus 3 sigma
__device__ __forceinline__ int pivot_rank(int n, int k) {
    const float q = 65.f * (float)k / (float)max(n, 1);
    return (int)ceilf(q + 3.f * sqrtf(q) + 1.5f);
}

// Exact selection: leaves the best m = min(k, #live) keys sorted best-first in cand[0, m), *s_cnt = m and,
// when m == k, *s_tau = the k-th.  Sampling rounds shrink the buffer to a few hundred keys, a rank sort
// orders them; an unlucky pivot (fewer than k keys above it, or tmp overflow) falls back to the full bitonic
// sort -- results never depend on the sampling.
template <int NT>
__device__ void select_topk(u64 *cand, int n, int k, u64 *tmp, int tmp_cap, int *s_cnt, u64 *s_tau, int *s_live,
                            u64 *s_pivot) {
    const int tid = threadIdx.x;
    for (int round = 0; round < 3; round++) {
        const int j = pivot_rank(n, k);
        if (n <= 2 * k || n <= 192 || j > 40) break;
        const int c = pivot_compact<NT>(cand, n, j, tmp, tmp_cap, s_live, s_pivot);
        if (c < k || c > tmp_cap) break;  // unlucky pivot: keep what we have
        for (int i = tid; i < c; i += NT) cand[i] = tmp[i];
        n = c;
        __syncthreads();
    }
    if (n <= tmp_cap && n <= 640) {
        for (int i = tid; i < n; i += NT) tmp[i] = cand[i];
        __syncthreads();
        rank_and_publish<NT>(tmp, n, k, cand, s_cnt, s_tau, s_live);
        return;
    }
    sort_and_publish<NT>(cand, n, k, cand, s_cnt, s_tau, s_live);
}

// Cheap selection between panels: sampling rounds only.  Leaves c >= k unsorted survivors in cand[0, c) with
// *s_cnt = c and *s_tau = the last pivot -- a valid lower bound of the k-th best, because at least k keys beat
// it -- or falls back to the exact selection when the first pivot fails.
template <int NT>
__device__ void tighten_topk(u64 *cand, int n, int k, u64 *tmp, int tmp_cap, int *s_cnt, u64 *s_tau, int *s_live,
                             u64 *s_pivot) {
    const int tid = threadIdx.x;
    bool done_one = false;
    u64 tau = 0ull;
    for (int round = 0; round < 3; round++) {
        const int j = pivot_rank(n, k);
        if (n <= 2 * k || j > 40) break;
        const int c = pivot_compact<NT>(cand, n, j, tmp, tmp_cap, s_live, s_pivot);
        if (c < k || c > tmp_cap) break;
        tau = *s_pivot;
        for (int i = tid; i < c; i += NT) cand[i] = tmp[i];
        n = c;
        done_one = true;
        __syncthreads();
    }
    if (!done_one) {
        select_topk<NT>(cand, n, k, tmp, tmp_cap, s_cnt, s_tau, s_live, s_pivot);
        return;
    }
    if (tid == 0) { *s_cnt = n; *s_tau = tau; }
    __syncthreads();
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

// shared-memory float add on a 32-bit shared address: LDS / FADD / ATOMS.CAST.SPIN loop, 543 Gadd/s on
// B200 (profiles/microbench); the plain atomicAdd(float*) form recomputes the shared window base per call.
__device__ __forceinline__ void smem_add_f32(unsigned addr, float x) {
    asm volatile("red.shared.add.f32 [%0], %1;" ::"r"(addr), "f"(x) : "memory");
}

// 16-byte gathers (two pairs each) a lane issues before its first add.  Measured on cfg2 (B200, 50-entry segments,
// 8 lanes per segment): 2 -> 204 ms, 1 -> 209 ms, 4 -> 220 ms, 3 -> 225 ms: with 64 registers per thread the deeper
// batches cost more in register pressure than they gain in memory-level parallelism (the L2 prefetch of the next
// segment provides that).
#ifndef SPY_UNROLL
#define SPY_UNROLL 2
#endif
__host__ __device__ constexpr int unroll_for(int threads) { return SPY_UNROLL; }
// Entries of a target row staged in shared memory at a time (8 bytes each): a quarter more than the CTA has
// threads, so that rows a little longer than the thread count (e.g. ~1000 +- 32 entries against 1024 threads)
// do not pay a second, nearly empty chunk in every panel.
__host__ __device__ constexpr int stage_entries(int threads) { return threads + threads / 4; }

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

// "this slot cannot enter the result": x < lo * den for den safely positive (see above)
template <int KIND>
__device__ __forceinline__ bool slot_rejected(const KnnDev &p, const FastRow &fr, float x, float lo, float yt, float yc,
                                              float yd) {
    if (KIND == KIND_RAW) return x < lo;
    if (KIND == KIND_C) {
        const float den = fmaf(fr.cC, yc, fr.A0);
        return den > 0.f && x < lo * den;
    }
    float den = fr.A0, sab = fabsf(fr.A0);
    if (KIND == KIND_T || (KIND == KIND_GEN && p.l1 != 0.f)) {
        const float u = fr.cT * yt, w = fr.cX * x;
        den += u + w;
        sab += fabsf(u) + fabsf(w);
    }
    if (KIND == KIND_GEN && p.l2 != 0.f) {
        const float u = fr.cC * yc;
        den += u;
        sab += fabsf(u);
    }
    if (KIND == KIND_D || (KIND == KIND_GEN && p.l3 != 0.f)) {
        const float u = fr.cD * yd;
        den += u;
        sab += fabsf(u);
    }
    // den huge / inf: lo * den is +-inf or NaN and the comparison does the right thing
    return den > 0.f && x < lo * den && den * 64.f >= sab;
}

// Survivors of a quad as a 4-bit mask: touched slots the pre-filter cannot reject.  KIND_C and KIND_RAW fold the
// bound into one FFMA + compare per slot:  x < lo * (cC * y + A0)  ==  x < (lo * cC) * y + lo * A0  (den >= 0 for
// these kinds: norms and shrink are non-negative; the regrouping moves the bound by a few ulp, far inside its
// 1e-4 safety margin).  An untouched slot holds -0.0f.
template <int KIND>
__device__ __forceinline__ unsigned survivor_mask(const KnnDev &p, const FastRow &fr, bool filter, float lo, float lc, float la,
                                                  const float4 &x, const float4 &yt, const float4 &yc, const float4 &yd) {
    const float xs[4] = {x.x, x.y, x.z, x.w};
    const float yts[4] = {yt.x, yt.y, yt.z, yt.w}, ycs[4] = {yc.x, yc.y, yc.z, yc.w}, yds[4] = {yd.x, yd.y, yd.z, yd.w};
    unsigned m = 0u;
#pragma unroll
    for (int r = 0; r < 4; r++) {
        bool rej = false;
        if (filter) {
            if (KIND == KIND_RAW) rej = xs[r] < lo;
            else if (KIND == KIND_C) rej = xs[r] < fmaf(lc, ycs[r], la);
            else rej = slot_rejected<KIND>(p, fr, xs[r], lo, yts[r], ycs[r], yds[r]);
        }
        if (!rej && __float_as_uint(xs[r]) != kSentinelBits) m |= 1u << r;
    }
    return m;
}

// raw (not yet evaluated) candidate: accumulated dot product + column
__device__ __forceinline__ u64 make_raw(float xy, int col) { return ((u64)__float_as_uint(xy) << 32) | (u64)(unsigned)col; }

// Buffer the survivors `m` of a quad RAW (dot product, column) with ONE reservation; their exact value is computed
// afterwards for the whole buffer with all lanes busy (the IEEE division of computeSimilarity would otherwise run
// with one or two lanes of a warp active).  A survivor that does not fit keeps its slot value.  Returns false
// when the buffer was full.
__device__ __forceinline__ bool push_quad(int cap, unsigned a, int col0, const float4 &x, unsigned m, u64 *cand, int *s_cnt,
                                          bool keep_rejected) {
    const float sent = __uint_as_float(kSentinelBits);
    const float xs[4] = {x.x, x.y, x.z, x.w};
    const int cnt = __popc(m);
    int pos = cap;
    if (*reinterpret_cast<volatile int *>(s_cnt) < cap) pos = atomicAdd(s_cnt, cnt);  // few lanes: they serialise on s_cnt
    const int fit = max(0, min(cnt, cap - pos));
    float ws[4] = {sent, sent, sent, sent};
    if (keep_rejected) { ws[0] = xs[0]; ws[1] = xs[1]; ws[2] = xs[2]; ws[3] = xs[3]; }  // only buffered slots are reset
#pragma unroll
    for (int r = 0; r < 4; r++)
        if (m & (1u << r)) {
            const int w = __popc(m & ((1u << r) - 1u));
            if (w < fit) { cand[pos + w] = make_raw(xs[r], col0 + r); ws[r] = sent; }
            else ws[r] = xs[r];
        }
    sts128(a, make_float4(ws[0], ws[1], ws[2], ws[3]));
    return fit == cnt;
}

// One drain pass over the current panel, in two phases.  A thread owns the quads tid + it * NT (it = 0, 1, ...: at
// most 14 steps); a warp's 32 quads of one step are the 128 consecutive columns of one block of the per-block
// minima of Y (y_block_min).
//   phase 1 (no global loads): if x < lo * (c * min_block(Y) + A0) for the four slots of a quad, nothing in it can
//           enter the result and it is reset to "untouched"; the other quads are only FLAGGED (one bit per step);
//   phase 2: the flagged quads load their Y vectors -- from L2: 227 KB of shared memory leave no L1 -- a batch at
//           a time, so that a thread (and with it the warp) waits once per batch, not once per quad; the per-slot
//           test then buffers the survivors.
// `todo` is the thread's set of steps still to drain (bit it); the pass returns what is left of it when the
// candidate buffer filled up (*s_overflow raised) -- the caller evaluates + selects, which raises tau, and calls
// again -- else 0.  No block barriers inside.
template <int NT, int KIND>
__device__ __forceinline__ unsigned drain_pass(const KnnDev &p, const FastRow &fr, unsigned acc32, int base, int width, float lo,
                                               float yv, u64 *cand, int *s_cnt, int *s_overflow, unsigned todo,
                                               bool keep_rejected) {
    const float sent = __uint_as_float(kSentinelBits);
    const float4 sent4 = make_float4(sent, sent, sent, sent);
    constexpr int S = NT * 4;
#ifndef SPY_DRAIN_BATCH
#define SPY_DRAIN_BATCH 2
#endif
    constexpr int B = (KIND == KIND_GEN) ? 1 : (KIND == KIND_T ? 2 : SPY_DRAIN_BATCH);  // quads per phase-2 batch (registers)
    const int tid = threadIdx.x, lane = tid & 31;
    const bool filter = !p.exact_only;
    const bool useT = filter && (KIND == KIND_T || (KIND == KIND_GEN && p.l1 != 0.f));
    const bool useC = filter && (KIND == KIND_C || (KIND == KIND_GEN && p.l2 != 0.f));
    const bool useD = filter && (KIND == KIND_D || (KIND == KIND_GEN && p.l3 != 0.f));
    // x < lc * y + la is the folded bound of KIND_C (y = Yc) and KIND_D (y = Yd)
    const float lc = lo * (KIND == KIND_D ? fr.cD : fr.cC), la = lo * fr.A0;
    const int wv = min(width, (p.n_cols - base) & ~3);  // quads below wv lie inside the matrix: float4 loads of Y are safe
    const int n_iter = (width + S - 1) / S;

    // ---- phase 1: coarse test against the per-block minima; lane i of the warp holds the block of step i (yv) ----
    // (a raw dot product needs no Y at all: its bound is lo itself)
    const bool coarse = filter && lo > 0.f &&
                        (KIND == KIND_RAW || ((KIND == KIND_C || KIND == KIND_D) && p.y_block_min != nullptr && lc >= 0.f));
    if (coarse) {  // block-uniform
        for (int it = 0; it < n_iter; it++) {
            const float bound = (KIND == KIND_RAW) ? lo : fmaf(lc, __shfl_sync(0xffffffffu, yv, it), la);
            const int i = tid * 4 + it * S;
            if ((todo >> it & 1u) && i < wv) {
                const unsigned a = acc32 + (unsigned)i * 4u;
                const float4 x = lds128(a);
                if ((x.x < bound) & (x.y < bound) & (x.z < bound) & (x.w < bound)) {
                    if (!keep_rejected) sts128(a, sent4);
                    todo &= ~(1u << it);
                }
            }
        }
    }

    // ---- phase 2: the per-slot test on what is left ----
    while (todo) {
        int its[B];
        float4 x[B], yt[B], yc[B], yd[B];
        unsigned rest = todo;
#pragma unroll
        for (int b = 0; b < B; b++) {
            its[b] = -1;
            yt[b] = sent4; yc[b] = sent4; yd[b] = sent4;
            if (rest) {
                const int it = __ffs(rest) - 1;
                rest &= rest - 1u;
                its[b] = it;
                const int i = tid * 4 + it * S;
                x[b] = lds128(acc32 + (unsigned)i * 4u);
                if (i < wv) {
                    if (useT) yt[b] = __ldg(reinterpret_cast<const float4 *>(p.Yt + base + i));
                    if (useC) yc[b] = __ldg(reinterpret_cast<const float4 *>(p.Yc + base + i));
                    if (useD) yd[b] = __ldg(reinterpret_cast<const float4 *>(p.Yd + base + i));
                } else {  // the last, partial quad of the matrix
                    if (useT) yt[b] = load_y4(p.Yt, base + i, p.n_cols);
                    if (useC) yc[b] = load_y4(p.Yc, base + i, p.n_cols);
                    if (useD) yd[b] = load_y4(p.Yd, base + i, p.n_cols);
                }
            }
        }
#pragma unroll
        for (int b = 0; b < B; b++) {
            if (its[b] >= 0) {
                const int i = tid * 4 + its[b] * S;
                const unsigned a = acc32 + (unsigned)i * 4u;
                const unsigned m = survivor_mask<KIND>(p, fr, filter, lo, lc, la, x[b], yt[b], yc[b], yd[b]);
                if (m == 0u) {
                    if (!keep_rejected) sts128(a, sent4);
                } else if (!push_quad(p.cap, a, base + i, x[b], m, cand, s_cnt, keep_rejected)) {
                    *s_overflow = 1;
                    return todo;  // this quad and everything after it
                }
                todo &= ~(1u << its[b]);
            }
        }
    }
    return 0u;
}

// Same for MODE_MATRIX target columns (s_plus.h:175-188): only the columns listed in the target row's sorted
// list [tlo, thi) can be candidates.  A consumed slot is reset, which also makes duplicate list entries harmless.
template <int NT>
__device__ __forceinline__ void drain_pass_list(const KnnDev &p, float *acc, int base, int width, int tlo, int thi, u64 *cand,
                                                int *s_cnt, int *s_overflow) {
    bool overflow = false;
    for (int q = tlo + threadIdx.x; q < thi; q += NT) {  // the whole (sorted) list: only columns of this panel count
        const int col = __ldg(p.t_indices + q);
        if (col < base || col >= base + width) continue;
        if (q > tlo && __ldg(p.t_indices + q - 1) == col) continue;  // one owner per column
        const float xy = acc[col - base];
        if (__float_as_uint(xy) == kSentinelBits) continue;
        int pos = p.cap;
        if (*reinterpret_cast<volatile int *>(s_cnt) < p.cap) pos = atomicAdd(s_cnt, 1);
        if (pos >= p.cap) { overflow = true; continue; }  // the slot keeps its value for the next pass
        cand[pos] = make_raw(xy, col);
        acc[col - base] = __uint_as_float(kSentinelBits);
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
#ifndef SPY_PHASE_TIMING
#define SPY_PHASE_TIMING 0
#endif
#if SPY_PHASE_TIMING
#define SPY_TICK(id) do { if (tid == 0) { const long long _t = clock64(); ph[id] += _t - t_last; t_last = _t; } } while (0)
#else
#define SPY_TICK(id) do { } while (0)
#endif
// one more sampling round right after a validated speculative bound (tighter bound for the later panels).  On cfg2
// the full library measures the same with it on (199.3 ms) and off (200.3 ms); kept on, which is the build the
// evidence under profiles/r01 was taken with
#ifndef SPY_SHARPEN
#define SPY_SHARPEN 1
#endif
#ifndef SPY_SPECULATE
#define SPY_SPECULATE 1
#endif
#ifndef SPY_PREFETCH
#define SPY_PREFETCH 1
#endif
__device__ __forceinline__ uint4 ldg128(const uint2 *ptr) {
    return __ldg(reinterpret_cast<const uint4 *>(ptr));
}

// N float adds to N shared-memory addresses; `pend` has bit r set for a live add.  red.shared.add.f32 is an
// LDS / FADD / ATOMS.CAST.SPIN loop in SASS (interleaving the loops by hand with atom.shared.cas was 3x slower).
template <int N>
__device__ __forceinline__ void smem_add_batch(const unsigned (&addr)[N], const float (&x)[N], unsigned pend) {
#pragma unroll
    for (int r = 0; r < N; r++)
        if (pend & (1u << r)) smem_add_f32(addr[r], x[r]);
}

// Bounds of a group's FIRST segment in the NEXT panel, fetched while the current panel is being accumulated
// and kept in two registers across the drain: the next accumulation then starts without the dependent
// "split point -> pairs" chain from DRAM (its pairs have been pulled into L2 as well).
struct NextFirst {
    int s, e;
    bool valid;
};

template <int NT, int G>
__device__ __forceinline__ bool accumulate_chunk(const ExpandArgs x, int n, const int *st_u, const float *st_v,
                                                 unsigned accb32, const NextFirst pre, bool want_next, NextFirst &nxt,
                                                 int *next_entry) {
    constexpr int GROUPS = NT / G;
    constexpr int GPW = 32 / G;        // groups per warp
    constexpr int U = unroll_for(NT);  // 16-byte loads (2 pairs each) in flight per lane
    const int tid = threadIdx.x;
    const int gl = tid & (G - 1);
    const int gw = (tid & 31) / G;     // the group's position inside its warp
    // A warp's first two batches of GPW entries are fixed (warp w: entries w * GPW + {0, GROUPS}); further batches are
    // claimed from a shared counter two passes ahead -- the expansion's `schedule(dynamic)` inside the CTA: warps
    // whose gathers came back late do fewer passes, so the panel barrier waits for less.
    int base0 = (tid & ~31) / G, base1 = base0 + GROUPS, base2;
    int idx = base0 + gw;
    bool any = false;
    // software pipeline over the group's entries: the bounds of entry i + 2 are being fetched and the pairs of
    // entry i + 1 are on their way into L2 while the pairs of entry i are gathered and accumulated
    int s = 0, e = 0, s1, e1, s2, e2;
    auto fetch = [&](int i, int &s_, int &e_) {
        s_ = 0; e_ = 0;
        if (i < n) {
            const int u = st_u[i];
            if (x.n_panels == 1) { s_ = __ldg(x.b_indptr + u); e_ = __ldg(x.b_indptr + u + 1); }
            else {
                const int *sp = x.b_split + (size_t)u * x.split_stride + x.pn;
                s_ = __ldg(sp); e_ = __ldg(sp + 1);
            }
        }
    };
    auto prefetch_l2 = [&](int s_, int e_) {  // lane l pulls line l of the segment (G lines cover G * 16 pairs)
#if SPY_PREFETCH == 2
        // experiment (to be measured): one bulk prefetch of exactly the segment's bytes per group instead of whole
        // 128-byte lines -- a 400-byte segment at a random offset drags in ~100 bytes it never uses with line prefetches
        if (gl == 0 && e_ > s_) {
            const uint2 *a = x.b_pairs + (s_ & ~1);
            const unsigned bytes = (unsigned)(((e_ - (s_ & ~1)) * 8 + 15) & ~15);
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a), "r"(bytes));
        }
#elif SPY_PREFETCH
        const char *nb = reinterpret_cast<const char *>(x.b_pairs + s_) + 128 * gl;
        if (nb < reinterpret_cast<const char *>(x.b_pairs + e_)) asm volatile("prefetch.global.L2 [%0];" ::"l"(nb));
#endif
    };
    if (pre.valid) { s = pre.s; e = pre.e; }
    else fetch(idx, s, e);
    fetch(base1 + gw, s1, e1);
    int e_far = 0;  // end of the first entry's segment in the next panel (its begin is this panel's end)
    if (want_next && idx < n) e_far = __ldg(x.b_split + (size_t)st_u[idx] * x.split_stride + x.pn + 2);
    nxt.s = e; nxt.valid = want_next;
    bool first_pass = true;
    while (base0 < n) {  // warp-uniform
        base2 = n;
        if (base1 < n) {  // claim the batch of the pass after next
            if ((tid & 31) == 0) base2 = atomicAdd(next_entry, GPW);
            base2 = __shfl_sync(0xffffffffu, base2, 0);
        }
        fetch(base2 + gw, s2, e2);
        if (!first_pass) prefetch_l2(s1, e1);  // its bounds were fetched a whole pass ago
        const float v = (idx < n) ? st_v[idx] : 0.f;
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
        if (first_pass) { prefetch_l2(s1, e1); first_pass = false; }
        s = s1; e = e1; s1 = s2; e1 = e2;
        base0 = base1; base1 = base2;
        idx = base0 + gw;
    }
    nxt.e = e_far;
    if (want_next) prefetch_l2(nxt.s, nxt.e);
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
__global__ void __launch_bounds__(NT, NT == 512 ? 2 : 1)
knn_flat_kernel(const __grid_constant__ KnnDev p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *acc = reinterpret_cast<float *>(smem_raw);
    unsigned char *ptr = smem_raw + (size_t)p.W * sizeof(float);
    u64 *cand;
    if (CAND_SMEM) { cand = reinterpret_cast<u64 *>(ptr); ptr += (size_t)p.cap * sizeof(u64); }
    else cand = p.cand_global + (size_t)blockIdx.x * p.cap;
    // staged chunk of the target row (8 bytes per thread); the same bytes serve as the selection's scratch
    // (`tmp`), so a selection invalidates the staged row (staged_c0)
    constexpr int CH = stage_entries(NT);                    // entries of the target row staged at a time
    int *st_u = reinterpret_cast<int *>(ptr);               // A's column ids = rows of B
    float *st_v = reinterpret_cast<float *>(st_u + CH);     // A's values
    u64 *tmp = reinterpret_cast<u64 *>(ptr);
    constexpr int kTmpCap = CH;

    __shared__ int s_cnt, s_overflow, s_live;
    __shared__ int s_entry[2];  // next unclaimed entry of the staged chunk (alternating between accumulate calls)
    __shared__ int s_next[5];  // the NEXT row of this CTA: queue slot, output position, row id, A-row begin / end
    __shared__ u64 s_tau, s_pivot;

    const int tid = threadIdx.x;
    const float sentinel = __uint_as_float(kSentinelBits);
    const float4 sentinel4 = make_float4(sentinel, sentinel, sentinel, sentinel);
    const unsigned acc32 = (unsigned)__cvta_generic_to_shared(acc);

    for (int i = tid * 4; i < p.W; i += NT * 4) *reinterpret_cast<float4 *>(acc + i) = sentinel4;
#if SPY_PHASE_TIMING
    long long ph[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long t_last = clock64();
    __shared__ unsigned long long s_busy;
    if (tid == 0) s_busy = 0ull;
#endif

    // Rows are claimed one ahead (the reference's `omp for schedule(dynamic)`, s_plus.h:337): thread 0 resolves the
    // next row's queue slot -> row id -> A-row bounds while the other warps are still accumulating, and every
    // thread loads its entry of that A row into registers before the current row's last drain, so that a row
    // starts without a chain of dependent global loads.
    auto claim_next = [&]() {
        const int slot = atomicAdd(p.work_counter, 1);
        int i_out = 0, t = 0, b = 0, e = 0;
        if (slot < p.n_targets) {
            i_out = p.row_order ? __ldg(p.row_order + slot) : slot;
            t = __ldg(p.targets + i_out);
            b = __ldg(p.a_indptr + t);
            e = __ldg(p.a_indptr + t + 1);
        }
        s_next[0] = slot; s_next[1] = i_out; s_next[2] = t; s_next[3] = b; s_next[4] = e;
    };
    int u_n = 0;
    float v_n = 0.f;
    auto load_next_entries = [&]() {  // first chunk of the next row -> registers
        if (s_next[0] < p.n_targets) {
            const int b = s_next[3], e = s_next[4];
            if (tid < e - b) { u_n = __ldg(p.a_indices + b + tid); v_n = __ldg(p.a_data + b + tid); }
        }
    };
    constexpr int kFirstClaim = 2 * (NT / G);  // entries covered by the warps' two fixed passes
    if (tid == 0) { claim_next(); s_entry[0] = kFirstClaim; s_entry[1] = kFirstClaim; }
    __syncthreads();
    load_next_entries();
    int calls = 0;  // accumulate_chunk calls so far: call c claims from s_entry[c & 1] and re-arms the other counter

    for (;;) {
        const int slot = s_next[0];
        if (slot >= p.n_targets) break;
        const int i_out = s_next[1], t = s_next[2], a0 = s_next[3], a1 = s_next[4];
        if (tid < a1 - a0) { st_u[tid] = u_n; st_v[tid] = v_n; }  // entries [0, NT) of the first chunk come from registers
        if (tid + NT < min(a1 - a0, CH)) {                        // its tail, if the row is that long, from memory
            st_u[tid + NT] = __ldg(p.a_indices + a0 + NT + tid);
            st_v[tid + NT] = __ldg(p.a_data + a0 + NT + tid);
        }
        __syncthreads();  // the staged chunk is visible; s_next and the previous row's s_cnt have been read by everyone
        if (tid == 0) { s_cnt = 0; s_tau = 0ull; s_overflow = 0; }  // first read after the accumulation's barrier
        SPY_TICK(7);
        int staged_c0 = a0;  // which chunk of the A row the staging area holds (-1: none)
        NextFirst pre = {0, 0, false};
        SimRow sr;
        sr.Xt = (p.l1 != 0.f) ? __ldg(p.Xt + t) : 0.f;
        sr.Xc = (p.l2 != 0.f) ? __ldg(p.Xc + t) : 0.f;
        sr.Xd = (p.l3 != 0.f) ? __ldg(p.Xd + t) : 0.f;
        u64 tau = 0ull;
        float lo = reject_bound(p, tau);
        int n_eval = 0;  // cand[0, n_eval) are evaluated keys, cand[n_eval, s_cnt) raw candidates (uniform)

        for (int pn = 0; pn < p.n_panels; pn++) {
            const int base = pn * p.W;
            const int width = min(p.W, p.n_cols - base);
            const unsigned accb32 = acc32 - (unsigned)base * 4u;  // &acc[col - base] == accb32 + 4 * col
            bool any = false;
#if SPY_PHASE_TIMING
            const long long t_acc0 = clock64();
#endif
            // ---------------- expand + accumulate (s_plus.h:358-403 / 418-438) ----------------
            for (int c0 = a0; c0 < a1; c0 += CH) {
                const int n = min(CH, a1 - c0);
                if (staged_c0 != c0) {
                    SPY_TICK(1);
                    __syncthreads();  // earlier readers of the staging area (previous chunk / selection scratch)
                    for (int i = tid; i < n; i += NT) { st_u[i] = __ldg(p.a_indices + c0 + i); st_v[i] = __ldg(p.a_data + c0 + i); }
                    __syncthreads();
                    staged_c0 = c0;
                    SPY_TICK(0);
                }
                const ExpandArgs x = {p.b_indptr, p.b_split, p.b_pairs, p.split_stride, p.n_panels, pn};
                const bool first = c0 == a0;  // the cross-panel prefetch covers the first chunk of the row
                NextFirst none = {0, 0, false}, got = none;
                if (tid == 0) s_entry[(calls + 1) & 1] = kFirstClaim;  // last used by call `calls - 1`: two barriers ago
                any |= accumulate_chunk<NT, G>(x, n, st_u, st_v, accb32, first ? pre : none,
                                               first && pn + 1 < p.n_panels, got, &s_entry[calls & 1]);
                calls++;
                if (first) pre = got;
            }
            if (pn == 0 && tid == 0) claim_next();  // hidden behind the other warps' accumulation
#if SPY_PHASE_TIMING
            if ((tid & 31) == 0 && a1 > a0) atomicAdd(&s_busy, (unsigned long long)(clock64() - t_acc0));
#endif
            // barrier: all adds have landed; nothing landed in this panel => accumulator still clean
            const int landed = __syncthreads_or(any ? 1 : 0);
            SPY_TICK(1);
            if (pn == p.n_panels - 1) load_next_entries();  // in flight during the last drain + final selection
            if (!landed) continue;
            SPY_TICK(13);
            FastRow fr;
            fr.A0 = p.stab + p.l1 * p.t1 * sr.Xt;
            fr.cT = p.l1 * p.t2;
            fr.cX = p.l1 * (1.f - p.t1 - p.t2);
            fr.cC = p.l2 * sr.Xc;
            fr.cD = p.l3 * sr.Xd;

            // ---------------- per-row filter matrix: erase filtered columns (s_plus.h:159-172) --
            // one coalesced sweep over the row's filter list per panel (a binary search for the panel's range would
            // be a chain of ~2 log2(n) dependent global loads in front of every drain)
            if (p.filter_mode == SPY_SEL_MATRIX) {
                const int fs = __ldg(p.f_indptr + t), fe = __ldg(p.f_indptr + t + 1);
                for (int q = fs + tid; q < fe; q += NT) {
                    const int c = __ldg(p.f_indices + q) - base;
                    if (c >= 0 && c < width) acc[c] = sentinel;
                }
                __syncthreads();
            }

            // ---------------- drain: pre-filter, similarity, threshold, top-k (s_plus.h:193-215) ----------
            int tlo = 0, thi = 0;  // the target row's list of admissible columns (matrix-mode target_cols)
            if (p.target_mode == SPY_SEL_MATRIX) { tlo = __ldg(p.t_indptr + t); thi = __ldg(p.t_indptr + t + 1); }
            // steps of this panel the thread still has to drain: bit it <=> quad tid + it * NT exists
            unsigned todo = 0u;
            for (int it = 0; tid * 4 + it * NT * 4 < width; it++) todo |= 1u << it;
            // per-block minimum of Y for the coarse test: lane i of a warp holds the block of the warp's step i
            float yv = 0.f;
            if (p.y_block_min != nullptr && (tid & 31) * NT * 4 < width) {
                const int blk = ((base + (tid & ~31) * 4) >> 7) + (tid & 31) * (NT * 4 >> 7);
                if (blk < ((p.n_cols + 127) >> 7)) yv = __ldg(p.y_block_min + blk);
            }
            // exact values of the raw candidates cand[n_eval, cnt), all lanes busy (computeSimilarity, s_plus.h:129-156, 206);
            // a candidate that cannot beat the valid running bound tau dies here
            auto evaluate = [&](int cnt) {
                for (int i = n_eval + tid; i < cnt; i += NT) {
                    const u64 raw = cand[i];
                    const int col = (int)(unsigned)(raw & 0xffffffffull);
                    const float val = similarity_value(p, sr, __uint_as_float((unsigned)(raw >> 32)),
                                                       p.l1 != 0.f ? __ldg(p.Yt + col) : 0.f, p.l2 != 0.f ? __ldg(p.Yc + col) : 0.f,
                                                       p.l3 != 0.f ? __ldg(p.Yd + col) : 0.f);
                    u64 key = 0ull;
                    if (val >= p.thr) key = make_key(val, col);
                    cand[i] = (key > tau) ? key : 0ull;
                }
                n_eval = cnt;
            };

            // ---- speculative bound for a row that has none yet (its first non-empty panel) ----
            // Without a bound the first pass floods the buffer and two selections are needed before the pre-filter
            // bites.  Instead: SAMPLE the panel (one quad per thread, spread over all blocks, slots left untouched),
            // take the r-th best sample as the bound, r chosen so that the panel holds k candidates above it with
            // overwhelming probability, and drain with it while KEEPING rejected slots.  The bound is then
            // validated -- at least k buffered candidates beat it -- before the panel is cleared; if it is not,
            // the panel is drained again with the valid bound, so results never depend on the sample.
            u64 tau_s = 0ull;
            bool spec = false;
            {
                const int n_iter = (width + NT * 4 - 1) / (NT * 4);
                // (n_eval == 0: nothing buffered yet -- a register every thread agrees on; s_cnt itself is modified below)
                if (SPY_SPECULATE && tau == 0ull && !p.exact_only && p.target_mode != SPY_SEL_MATRIX && n_iter >= 4 && n_eval == 0) {
                    const int i = tid * 4 + (tid % n_iter) * NT * 4;
                    float4 x = sentinel4;
                    if (i + 3 < width) x = lds128(acc32 + (unsigned)i * 4u);
                    const float xs[4] = {x.x, x.y, x.z, x.w};
                    unsigned m = 0u;
#pragma unroll
                    for (int r = 0; r < 4; r++)
                        if (__float_as_uint(xs[r]) != kSentinelBits) m |= 1u << r;
                    // one reservation per warp: inclusive scan of the lanes' counts
                    const int c = __popc(m);
                    int inc = c;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int v = __shfl_up_sync(0xffffffffu, inc, o);
                        if ((tid & 31) >= o) inc += v;
                    }
                    int wbase = 0;
                    if ((tid & 31) == 31 && inc > 0) wbase = atomicAdd(&s_cnt, inc);
                    wbase = __shfl_sync(0xffffffffu, wbase, 31);
                    int pos = wbase + inc - c;
#pragma unroll
                    for (int r = 0; r < 4; r++)
                        if ((m & (1u << r)) && pos < p.cap) cand[pos++] = make_raw(xs[r], base + i + r);
                        else if (m & (1u << r)) pos++;
                    __syncthreads();
                    const int attempts = s_cnt, n_s = min(attempts, p.cap);
                    evaluate(n_s);
                    __syncthreads();
                    const float f = (float)min(width, NT * 4) / (float)width * (float)n_s / (float)max(attempts, 1);
                    const float kf = (float)p.k * f;
#ifdef SPY_SPEC_FORCE_RANK  // test builds: a bound that is almost never valid, to exercise the re-drain path
                    const int r_s = SPY_SPEC_FORCE_RANK;
#else
                    const int r_s = (int)ceilf(kf + 3.5f * sqrtf(kf) + 1.5f);
#endif
                    if (f < 0.4f && n_s >= 8 * r_s) {  // uniform
                        select_topk<NT>(cand, n_s, r_s, tmp, kTmpCap, &s_cnt, &s_tau, &s_live, &s_pivot);
                        staged_c0 = -1;  // tmp overlays the staged row
                        if (s_cnt == r_s) { tau_s = s_tau; spec = true; }
                    }
                    __syncthreads();
                    if (tid == 0) { s_cnt = 0; s_tau = 0ull; }  // the sample was only read: its slots are all still there
                    n_eval = 0;
                    __syncthreads();
                }
            }

            for (int attempt = 0; attempt < 2; attempt++) {
                const float lo_use = spec ? reject_bound(p, tau_s > tau ? tau_s : tau) : lo;
                for (;;) {
                    if (p.target_mode == SPY_SEL_MATRIX) drain_pass_list<NT>(p, acc, base, width, tlo, thi, cand, &s_cnt, &s_overflow);
                    else todo = drain_pass<NT, KIND>(p, fr, acc32, base, width, spec ? fmaxf(lo_use, lo) : lo, yv, cand, &s_cnt,
                                                     &s_overflow, todo, spec);
#if SPY_PHASE_TIMING
                    if (tid == 0) { const long long _t = clock64(); ph[pn == 0 ? 11 : 12] += _t - t_last; t_last = _t; }
#endif
                    __syncthreads();
                    SPY_TICK(pn == 0 ? 3 : 8);
#if SPY_PHASE_TIMING
                    if (tid == 0) ph[pn == 0 ? 9 : 10] += 1;
#endif
                    const bool again = s_overflow != 0;
                    const int cnt = min(s_cnt, p.cap);
                    evaluate(cnt);
                    __syncthreads();
                    SPY_TICK(4);
                    if (again || cnt > p.cap / 2) {  // tighten tau while the buffer is reasonably full
                        if (tid == 0) s_overflow = 0;
                        tighten_topk<NT>(cand, cnt, p.k, tmp, kTmpCap, &s_cnt, &s_tau, &s_live, &s_pivot);
                        staged_c0 = -1;  // tmp overlays the staged row
                        tau = s_tau;
                        lo = reject_bound(p, tau);
                        n_eval = s_cnt;
                        SPY_TICK(5);
                    }
                    if (!again) break;
                }
                if (!spec) break;
                // validate the speculative bound: do k buffered candidates beat it?
                if (tid == 0) s_live = 0;
                __syncthreads();
                int above = 0;
                for (int i = tid; i < n_eval; i += NT) above += cand[i] > tau_s ? 1 : 0;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) above += __shfl_xor_sync(0xffffffffu, above, o);
                if ((tid & 31) == 0 && above) atomicAdd(&s_live, above);
                __syncthreads();
                const bool valid = s_live >= p.k;
                __syncthreads();
                spec = false;
                if (valid) {
                    if (tau_s > tau) { tau = tau_s; lo = reject_bound(p, tau); }
                    for (int i = tid * 4; i < width; i += NT * 4) sts128(acc32 + (unsigned)i * 4u, sentinel4);  // clear the panel
                    __syncthreads();  // before anyone accumulates the next panel into it
                    if (SPY_SHARPEN && n_eval > 2 * p.k) {  // sharpen the bound for the panels to come (one sampling round)
                        tighten_topk<NT>(cand, n_eval, p.k, tmp, kTmpCap, &s_cnt, &s_tau, &s_live, &s_pivot);
                        staged_c0 = -1;
                        tau = s_tau;
                        lo = reject_bound(p, tau);
                        n_eval = s_cnt;
                    }
                    break;
                }
                // not validated: everything rejected so far is still in the panel; drain it again with the valid bound
                todo = 0u;
                for (int it = 0; tid * 4 + it * NT * 4 < width; it++) todo |= 1u << it;
            }
            if (p.target_mode == SPY_SEL_MATRIX) {  // touched slots outside the target list
                for (int i = tid * 4; i < width; i += NT * 4) *reinterpret_cast<float4 *>(acc + i) = sentinel4;
                __syncthreads();  // before anyone accumulates the next panel into it
            }
        }

        // the next row's split points (one 32-byte sector per entry holds all panels) -> L2, before they are needed
        if (SPY_PREFETCH && s_next[0] < p.n_targets && tid < s_next[4] - s_next[3]) {
            const int *sp = (p.n_panels == 1) ? p.b_indptr + u_n : p.b_split + (size_t)u_n * p.split_stride;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(sp));
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
        SPY_TICK(6);
    }
#if SPY_PHASE_TIMING
    __syncthreads();
    if (tid == 0) {
        ph[2] = (long long)(s_busy / (NT / 32));  // mean over warps of "cycles from the end of staging to the warp's own end"
        for (int i = 0; i < 16; i++) atomicAdd(p.phase + i, (u64)ph[i]);
    }
#endif
}

typedef void (*knn_kernel_t)(const KnnDev);
// one translation unit per group width (knn_inst_g*.cu): the kernel for (threads, similarity kind, buffer placement)
knn_kernel_t pick_kernel_g4(int threads, int kind, bool cand_smem);
knn_kernel_t pick_kernel_g8(int threads, int kind, bool cand_smem);
knn_kernel_t pick_kernel_g16(int threads, int kind, bool cand_smem);
knn_kernel_t pick_kernel_g32(int threads, int kind, bool cand_smem);

}  // namespace spy
